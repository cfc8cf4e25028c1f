// vb_scan_bulk.cu -- the list / table scan with TMA bulk copies (cp.async.bulk + mbarrier).
//
// Same work, arithmetic and output as scan_kernel (vb_scan.cu); different data movement:
// the rows of a chunk are contiguous in HBM, so ONE producer thread streams them into a ring of
// shared-memory stages with 1-D bulk async copies (tens of KB each, completion counted on an
// mbarrier), while 8 consumer warps read rows and the query image from shared memory
// (conflict-free 128-bit LDS), accumulate in fp32 and reduce with warp shuffles.  Load issue is
// decoupled from consumption, so the bytes in flight per SM stay at STAGES x stage size
// (~190 KB) independent of register pressure or of the epilogue of a row.
//
// Roofline: HBM.  Algorithmic bytes per distance = dim x element size.
#include "vb_common.cuh"
#include "vb_distance.cuh"

#include <algorithm>

namespace vb {

constexpr int SB_CONSUMERS = 8;                       // consumer warps
constexpr int SB_THREADS = (SB_CONSUMERS + 1) * 32;   // + 1 producer warp
constexpr int SB_MAX_STAGES = 4;
constexpr int SB_STAGE_TARGET = 48 * 1024;

__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool sb_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(sb_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sb_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!sb_mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void sb_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(sb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sb_consumer_barrier() {
    // named barrier 1: the consumer warps only (the producer never joins)
    asm volatile("bar.sync 1, %0;" ::"r"(SB_CONSUMERS * 32) : "memory");
}

struct BulkShape {
    int stage_rows;   // rows per stage = SB_CONSUMERS * rows_per_warp
    int rows_per_warp;
    int stages;
    uint32_t stage_bytes;
    size_t smem;
};

template <int ELEM, int METRIC, typename OUT>
__global__ void __launch_bounds__(SB_THREADS, 1) scan_bulk_kernel(ScanArgs a, BulkShape sh) {
    extern __shared__ uint8_t sb_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sb_raw) + 127) & ~(uintptr_t)127);
    uint8_t* stage_base = smem;
    uint4* sq = reinterpret_cast<uint4*>(smem + (size_t)sh.stages * sh.stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sq) + ((a.qstride + 127) & ~(size_t)127));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + SB_MAX_STAGES;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < sh.stages; ++s) {
            sb_mbar_init(&full_bar[s], 1);
            sb_mbar_init(&empty_bar[s], SB_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int64_t total;
    if (a.chunks) total = *a.n_chunks_dev;
    else total = a.nq * a.chunks_per_q;
    const int64_t per = (total + gridDim.x - 1) / gridDim.x;
    const int64_t c_begin = per * blockIdx.x;
    const int64_t c_end = min(total, c_begin + per);
    const int V = a.vec_per_row;

    auto chunk_of = [&](int64_t c, int64_t& row_begin, int64_t& out_off, int& n_rows, int& q) {
        if (a.chunks) {
            Chunk ch = a.chunks[c];
            row_begin = ch.row_begin;
            out_off = ch.out_off;
            n_rows = ch.n_rows;
            q = ch.q;
        } else {
            q = (int)(c / a.chunks_per_q);
            int64_t r0 = (c % a.chunks_per_q) * a.rows_per_chunk;
            row_begin = r0;
            n_rows = (int)min((int64_t)a.rows_per_chunk, a.n_rows - r0);
            out_off = (int64_t)q * a.out_stride + r0;
        }
    };

    if (warp == SB_CONSUMERS) {
        // ===== producer: one bulk copy per stage, running ahead of the consumers by `stages` =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t c = c_begin; c < c_end; ++c) {
                int64_t row_begin, out_off;
                int n_rows, q;
                chunk_of(c, row_begin, out_off, n_rows, q);
                for (int r0 = 0; r0 < n_rows; r0 += sh.stage_rows, ++it) {
                    const int s = it % sh.stages;
                    const uint32_t ph = (it / sh.stages) & 1;
                    const int nr = min(sh.stage_rows, n_rows - r0);
                    sb_mbar_wait(&empty_bar[s], ph ^ 1);
                    const uint32_t bytes = (uint32_t)((size_t)nr * a.stride);
                    sb_mbar_expect_tx(&full_bar[s], bytes);
                    sb_bulk_g2s(stage_base + (size_t)s * sh.stage_bytes, a.rows + (size_t)(row_begin + r0) * a.stride, bytes, &full_bar[s]);
                }
            }
        }
    } else {
        // ===== consumers =====
        uint32_t it = 0;
        int cur_q = -1;
        for (int64_t c = c_begin; c < c_end; ++c) {
            int64_t row_begin, out_off;
            int n_rows, q;
            chunk_of(c, row_begin, out_off, n_rows, q);
            if (q != cur_q) {
                sb_consumer_barrier();      // everyone is done with the previous query image
                const uint4* gq = reinterpret_cast<const uint4*>(a.queries + (size_t)q * a.qstride);
                if (ELEM == VB_HALFVEC) {
                    // fp32 image -> packed halves (exact: the image came from halves): halves the shared-memory
                    // traffic of the query operand, which is what bounds the halfvec scan
                    for (int i = threadIdx.x; i < V; i += SB_CONSUMERS * 32) {
                        const uint4 lo = gq[2 * i], hi = gq[2 * i + 1];
                        __half2 h0 = __floats2half2_rn(__uint_as_float(lo.x), __uint_as_float(lo.y));
                        __half2 h1 = __floats2half2_rn(__uint_as_float(lo.z), __uint_as_float(lo.w));
                        __half2 h2 = __floats2half2_rn(__uint_as_float(hi.x), __uint_as_float(hi.y));
                        __half2 h3 = __floats2half2_rn(__uint_as_float(hi.z), __uint_as_float(hi.w));
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&h0);
                        pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2);
                        pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        sq[i] = pk;
                    }
                } else {
                    for (int i = threadIdx.x; i < a.qvec; i += SB_CONSUMERS * 32) sq[i] = gq[i];
                }
                sb_consumer_barrier();
                cur_q = q;
            }
            OUT* out = reinterpret_cast<OUT*>(a.out) + out_off;
            for (int r0 = 0; r0 < n_rows; r0 += sh.stage_rows, ++it) {
                const int s = it % sh.stages;
                const uint32_t ph = (it / sh.stages) & 1;
                const int nr = min(sh.stage_rows, n_rows - r0);
                sb_mbar_wait(&full_bar[s], ph);
                const uint8_t* st = stage_base + (size_t)s * sh.stage_bytes;
                // warp w owns rows w, w + 8, ... of the stage and takes them two at a time: the query vectors
                // are read from shared memory once per pair and the two reductions overlap
                for (int rr = warp; rr < nr; rr += 2 * SB_CONSUMERS) {
                    const int rr1 = rr + SB_CONSUMERS;
                    const uint4* rp0 = reinterpret_cast<const uint4*>(st + (size_t)rr * a.stride);
                    if (rr1 < nr) {
                        const uint4* rp1 = reinterpret_cast<const uint4*>(st + (size_t)rr1 * a.stride);
                        Acc<ELEM, METRIC> a0, a1;
#pragma unroll 2
                        for (int v = lane; v < V; v += 32) {
                            const uint4 x0 = rp0[v], x1 = rp1[v];
                            if (ELEM == VB_HALFVEC) {
                                const uint4 qh = sq[v];
                                a0.add_h(x0, qh);
                                a1.add_h(x1, qh);
                            } else {
                                a0.add(x0, sq, v);
                                a1.add(x1, sq, v);
                            }
                        }
                        a0.template reduce<32>();
                        a1.template reduce<32>();
                        if (lane == 0) {
                            out[r0 + rr] = (OUT)a0.value();
                            out[r0 + rr1] = (OUT)a1.value();
                        }
                    } else {
                        Acc<ELEM, METRIC> acc;
#pragma unroll 4
                        for (int v = lane; v < V; v += 32) {
                            if (ELEM == VB_HALFVEC) acc.add_h(rp0[v], sq[v]);
                            else acc.add(rp0[v], sq, v);
                        }
                        acc.template reduce<32>();
                        if (lane == 0) out[r0 + rr] = (OUT)acc.value();
                    }
                }
                __syncwarp();
                if (lane == 0) sb_mbar_arrive(&empty_bar[s]);
            }
        }
    }
}

static BulkShape bulk_shape(size_t stride, size_t qstride) {
    BulkShape sh{};
    int rpw = (int)(SB_STAGE_TARGET / (SB_CONSUMERS * stride));
    rpw = std::max(1, std::min(rpw, 8));
    sh.rows_per_warp = rpw;
    sh.stage_rows = rpw * SB_CONSUMERS;
    sh.stage_bytes = (uint32_t)(((size_t)sh.stage_rows * stride + 127) & ~(size_t)127);
    const size_t fixed = ((qstride + 127) & ~(size_t)127) + 2 * SB_MAX_STAGES * sizeof(uint64_t) + 256;
    const size_t budget = 220 * 1024;
    int stages = (int)((budget - fixed) / sh.stage_bytes);
    sh.stages = std::max(0, std::min(stages, SB_MAX_STAGES));
    sh.smem = (size_t)sh.stages * sh.stage_bytes + fixed;
    return sh;
}

bool scan_bulk_supported(int elem, size_t stride, size_t qstride) {
    (void)elem;
    if (stride < 512) return false;   // fewer than 32 vectors per row: the sub-warp LDG variant fits better
    if (qstride > 64 * 1024) return false;
    return bulk_shape(stride, qstride).stages >= 2;
}

template <int ELEM, int METRIC, typename OUT>
static int launch_bulk_t(const ScanArgs& a, int grid, cudaStream_t s) {
    BulkShape sh = bulk_shape(a.stride, a.qstride);
    auto kern = scan_bulk_kernel<ELEM, METRIC, OUT>;
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
    kern<<<grid, SB_THREADS, sh.smem, s>>>(a, sh);
    VB_CUDA(cudaGetLastError());
    count_launch();
    return VB_OK;
}

template <typename OUT>
static int launch_bulk_any(int elem, int metric, const ScanArgs& a, int grid, cudaStream_t s) {
    if (elem == VB_VECTOR) {
        switch (metric) {
            case VB_L2_SQUARED: return launch_bulk_t<VB_VECTOR, VB_L2_SQUARED, OUT>(a, grid, s);
            case VB_NEG_IP: return launch_bulk_t<VB_VECTOR, VB_NEG_IP, OUT>(a, grid, s);
            case VB_COSINE: return launch_bulk_t<VB_VECTOR, VB_COSINE, OUT>(a, grid, s);
            case VB_L1: return launch_bulk_t<VB_VECTOR, VB_L1, OUT>(a, grid, s);
        }
    } else if (elem == VB_HALFVEC) {
        switch (metric) {
            case VB_L2_SQUARED: return launch_bulk_t<VB_HALFVEC, VB_L2_SQUARED, OUT>(a, grid, s);
            case VB_NEG_IP: return launch_bulk_t<VB_HALFVEC, VB_NEG_IP, OUT>(a, grid, s);
            case VB_COSINE: return launch_bulk_t<VB_HALFVEC, VB_COSINE, OUT>(a, grid, s);
            case VB_L1: return launch_bulk_t<VB_HALFVEC, VB_L1, OUT>(a, grid, s);
        }
    } else {
        switch (metric) {
            case VB_HAMMING: return launch_bulk_t<VB_BIT, VB_HAMMING, OUT>(a, grid, s);
            case VB_JACCARD: return launch_bulk_t<VB_BIT, VB_JACCARD, OUT>(a, grid, s);
        }
    }
    set_error("unsupported metric %d for element type %d", metric, elem);
    return VB_EINVAL;
}

int launch_scan_bulk(int elem, int metric, const ScanArgs& a, bool out_f64, int max_chunks_hint) {
    // persistent: one CTA per SM (the stage ring uses most of the shared memory)
    int grid = ctx().sm_count;
    if (max_chunks_hint > 0) grid = std::min(grid, max_chunks_hint);
    if (out_f64) return launch_bulk_any<double>(elem, metric, a, grid, ctx().stream);
    return launch_bulk_any<float>(elem, metric, a, grid, ctx().stream);
}

}  // namespace vb
