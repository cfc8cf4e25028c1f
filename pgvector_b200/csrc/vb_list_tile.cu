// vb_list_tile.cu -- the IVFFlat list scan for a BATCH of queries, list-major.
//
// The reference scans the probed lists once per query (GetScanItems, src/ivfscan.c:124-180: every
// tuple of every probed list goes through the distance function).  With B queries in flight each
// list is probed by B * probes / lists of them, so the per-query formulation (vb_scan*.cu) streams
// the same rows from HBM that many times -- 131 GB per 2048-query step for config B where the table
// is 6.1 GB.  Here the (query, list) pairs of a batch are grouped by list and one CTA computes the
// distances of a 256-row tile of a list against ALL queries that probe it: rows are read from HBM
// once per batch, the arithmetic (fp32, the same (x - q)^2 / x * q FMA chain per (row, query),
// accumulated sequentially over the dimensions) moves to the FMA pipe.  Distances land in the same
// per-query candidate run as the streaming scan writes them, so everything downstream (top-k by
// (distance, position), heap-id lookup) is unchanged.
//
// Tile: 256 rows x 32 queries per CTA, 8 warps.  Warp w owns queries 4w..4w+3 (a warp whose
// queries are all past the group's end skips the arithmetic), lane l owns rows 4l..4l+3 and
// 128+4l..128+4l+3.  Operands are staged through shared memory in steps of 16 dimensions,
// k-major, double buffered with register prefetch; the query operand is stored duplicated
// ((q, q), negated for L2) so the inner loop is packed fp32x2: one FADD2 + one FFMA2 per two
// (row, query) pairs.  fp32x2 results are IEEE-identical to the scalar instructions.
//
// Bound: fp32 issue (2 packed instructions per 2 pairs for L2, 1 for inner product); HBM traffic is
// one pass over the probed lists per batch.
#include "vb_common.cuh"

#include <algorithm>

namespace vb {

constexpr int LT_ROWS = 256;
constexpr int LT_Q = 32;
constexpr int LT_KS = 16;                 // dimensions per staging step
constexpr int LT_THREADS = 256;
constexpr int LT_XP = LT_ROWS + 4;        // padded line of one dimension across the row tile (words)

struct LtArgs {
    const uint8_t* rows;
    size_t stride;
    const uint8_t* qimg;
    size_t qstride;
    const ListTile* tiles;
    const int64_t* list_off;
    const int32_t* grp_begin;
    const int32_t* grp_cnt;
    const int32_t* pair_q;
    const int64_t* pair_out;
    float* out;
    int words;   // padded dimension count (elements per row including the zero padding)
};

__device__ __forceinline__ unsigned long long lt_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void lt_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long lt_add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long lt_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// 4 consecutive elements of a row starting at element e (zero past the padded dimension count)
template <int ELEM>
__device__ __forceinline__ float4 lt_load4(const uint8_t* row, int e, int words) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < words) {
        if (ELEM == VB_VECTOR) {
            v = __ldg(reinterpret_cast<const float4*>(row) + (e >> 2));
        } else {
            const uint2 h = __ldg(reinterpret_cast<const uint2*>(row) + (e >> 2));
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            v = make_float4(a.x, a.y, b.x, b.y);
        }
    }
    return v;
}

// KIND 0: sum (x - q)^2      KIND 1: -sum x * q
template <int ELEM, int KIND>
__global__ void __launch_bounds__(LT_THREADS, 2) list_tile_kernel(LtArgs a) {
    const ListTile t = a.tiles[blockIdx.x];
    const int cnt = a.grp_cnt[t.list];
    if (cnt == 0) return;   // list not probed by this batch

    __shared__ __align__(16) float Xs[2][LT_KS][LT_XP];
    __shared__ __align__(16) float2 Qs[2][LT_KS][LT_Q];
    __shared__ int32_t s_q[LT_Q];
    __shared__ int64_t s_out[LT_Q];

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int gb = a.grp_begin[t.list];
    const int64_t row_in_list0 = t.row_begin - a.list_off[t.list];
    const int words = a.words;
    const int nsteps = (words + LT_KS - 1) / LT_KS;

    // staging roles: 4 threads per row (one 4-element piece each), 64 rows per pass, 4 passes
    const int sp = tid % 4;
    const uint8_t* xrow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = min(tid / 4 + 64 * j, t.n_rows - 1);
        xrow[j] = a.rows + (size_t)(t.row_begin + r) * a.stride;
    }

    for (int q0 = 0; q0 < cnt; q0 += LT_Q) {
        const int nqt = min(LT_Q, cnt - q0);
        __syncthreads();   // the previous query sub-tile is done with s_q / s_out and both buffers
        if (tid < LT_Q) {
            const int s = gb + q0 + min(tid, nqt - 1);
            s_q[tid] = a.pair_q[s];
            s_out[tid] = a.pair_out[s];
        }
        __syncthreads();
        const bool active = warp * 4 < nqt;
        const float* qrow = reinterpret_cast<const float*>(a.qimg + (size_t)s_q[(tid / 4) % LT_Q] * a.qstride);

        unsigned long long acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

        float4 xr[4], qr;
        auto fetch = [&](int ks) {
            const int e = ks * LT_KS + 4 * sp;
#pragma unroll
            for (int j = 0; j < 4; ++j) xr[j] = lt_load4<ELEM>(xrow[j], e, words);
            qr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < 4 * LT_Q && e < words) qr = __ldg(reinterpret_cast<const float4*>(qrow) + (e >> 2));
        };
        auto stage = [&](int buf) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = tid / 4 + 64 * j;
                Xs[buf][4 * sp + 0][r] = xr[j].x;
                Xs[buf][4 * sp + 1][r] = xr[j].y;
                Xs[buf][4 * sp + 2][r] = xr[j].z;
                Xs[buf][4 * sp + 3][r] = xr[j].w;
            }
            if (tid < 4 * LT_Q) {
                const int qi = tid / 4;
                const float s = KIND == 0 ? -1.f : 1.f;
                Qs[buf][4 * sp + 0][qi] = make_float2(s * qr.x, s * qr.x);
                Qs[buf][4 * sp + 1][qi] = make_float2(s * qr.y, s * qr.y);
                Qs[buf][4 * sp + 2][qi] = make_float2(s * qr.z, s * qr.z);
                Qs[buf][4 * sp + 3][qi] = make_float2(s * qr.w, s * qr.w);
            }
        };

        fetch(0);
        stage(0);
        __syncthreads();
        for (int ks = 0; ks < nsteps; ++ks) {
            const int buf = ks & 1;
            if (ks + 1 < nsteps) fetch(ks + 1);
            if (active) {
#pragma unroll
                for (int kk = 0; kk < LT_KS; ++kk) {
                    const ulonglong2 xa = *reinterpret_cast<const ulonglong2*>(&Xs[buf][kk][lane * 4]);
                    const ulonglong2 xb = *reinterpret_cast<const ulonglong2*>(&Xs[buf][kk][128 + lane * 4]);
                    const ulonglong2 qa = *reinterpret_cast<const ulonglong2*>(&Qs[buf][kk][warp * 4]);
                    const ulonglong2 qb = *reinterpret_cast<const ulonglong2*>(&Qs[buf][kk][warp * 4 + 2]);
                    const unsigned long long x[4] = {xa.x, xa.y, xb.x, xb.y};
                    const unsigned long long q[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (KIND == 0) {
                                const unsigned long long d = lt_add2(x[i], q[j]);   // x + (-q)
                                acc[i][j] = lt_fma2(d, d, acc[i][j]);
                            } else {
                                acc[i][j] = lt_fma2(x[i], q[j], acc[i][j]);
                            }
                        }
                }
            }
            if (ks + 1 < nsteps) stage(buf ^ 1);
            __syncthreads();
        }

        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int qi = warp * 4 + j;
                if (qi >= nqt) break;
                float* o = a.out + s_out[qi] + row_in_list0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float lo, hi;
                    lt_unpack(acc[i][j], lo, hi);
                    if (KIND == 1) {
                        lo = -lo;
                        hi = -hi;
                    }
                    const int r = (i >> 1) * 128 + lane * 4 + (i & 1) * 2;
                    if (r < t.n_rows) o[r] = lo;
                    if (r + 1 < t.n_rows) o[r + 1] = hi;
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- grouping the (query, probe) pairs by list

__global__ void lt_count_kernel(const int32_t* __restrict__ probe_lists, int64_t n_pairs, int32_t* __restrict__ cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int l = probe_lists[i];
    if (l >= 0) atomicAdd(&cnt[l], 1);
}

// exclusive prefix sum over the lists (one CTA; the list count is at most a few tens of thousands)
__global__ void __launch_bounds__(1024) lt_scan_kernel(const int32_t* __restrict__ cnt, int n, int32_t* __restrict__ begin) {
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int32_t v = i < n ? cnt[i] : 0;
        int32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) warp_sum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const int32_t before = carry + (warp ? warp_sum[warp - 1] : 0);
        if (i < n) begin[i] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + s;
        __syncthreads();
    }
}

__global__ void lt_scatter_kernel(const int32_t* __restrict__ probe_lists, int64_t n_pairs, int probes,
                                  const int32_t* __restrict__ cand_off, int64_t cap, const int32_t* __restrict__ begin,
                                  int32_t* __restrict__ cursor, int32_t* __restrict__ pair_q, int64_t* __restrict__ pair_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int l = probe_lists[i];
    if (l < 0) return;
    const int64_t q = i / probes;
    const int p = (int)(i % probes);
    const int slot = begin[l] + atomicAdd(&cursor[l], 1);
    pair_q[slot] = (int32_t)q;
    pair_out[slot] = q * cap + cand_off[q * (probes + 1) + p];
}

bool list_major_supported(int elem, int key_metric) {
    return (elem == VB_VECTOR || elem == VB_HALFVEC) && (key_metric == VB_L2_SQUARED || key_metric == VB_NEG_IP);
}

enum { WSL_GROUPS = 20 };

int launch_list_major(const Table& rows, int key_metric, const void* qimg, size_t qstride, int64_t nq, const int32_t* d_lists,
                      int probes, const int32_t* cand_off, int64_t cap, const int64_t* d_list_off, int n_lists,
                      const ListTile* d_tiles, int n_tiles, float* out) {
    VB_REQUIRE(list_major_supported(rows.elem, key_metric), "list-major scan: unsupported element type / metric");
    if (nq <= 0 || n_tiles <= 0) return VB_OK;
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const int64_t n_pairs = nq * probes;
    VB_REQUIRE(n_pairs < (int64_t)INT32_MAX, "too many (query, probe) pairs");
    void* d_ws;
    const size_t ints = (size_t)n_lists * 3 + (size_t)n_pairs;
    VB_TRY(workspace(WSL_GROUPS, sizeof(int64_t) * (size_t)n_pairs + sizeof(int32_t) * ints + 64, &d_ws));
    int64_t* pair_out = (int64_t*)d_ws;
    int32_t* pair_q = (int32_t*)(pair_out + n_pairs);
    int32_t* cnt = pair_q + n_pairs;
    int32_t* cursor = cnt + n_lists;
    int32_t* begin = cursor + n_lists;
    VB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)n_lists * 2, s));
    const unsigned gp = (unsigned)((n_pairs + 255) / 256);
    lt_count_kernel<<<gp, 256, 0, s>>>(d_lists, n_pairs, cnt);
    lt_scan_kernel<<<1, 1024, 0, s>>>(cnt, n_lists, begin);
    lt_scatter_kernel<<<gp, 256, 0, s>>>(d_lists, n_pairs, probes, cand_off, cap, begin, cursor, pair_q, pair_out);
    VB_CUDA(cudaGetLastError());
    count_launch(3);

    LtArgs a{};
    a.rows = rows.d;
    a.stride = rows.stride;
    a.qimg = (const uint8_t*)qimg;
    a.qstride = qstride;
    a.tiles = d_tiles;
    a.list_off = d_list_off;
    a.grp_begin = begin;
    a.grp_cnt = cnt;
    a.pair_q = pair_q;
    a.pair_out = pair_out;
    a.out = out;
    a.words = (int)(rows.elem == VB_HALFVEC ? rows.stride / 2 : rows.stride / 4);
    const int kind = key_metric == VB_L2_SQUARED ? 0 : 1;
    if (rows.elem == VB_VECTOR) {
        if (kind == 0) list_tile_kernel<VB_VECTOR, 0><<<n_tiles, LT_THREADS, 0, s>>>(a);
        else list_tile_kernel<VB_VECTOR, 1><<<n_tiles, LT_THREADS, 0, s>>>(a);
    } else {
        if (kind == 0) list_tile_kernel<VB_HALFVEC, 0><<<n_tiles, LT_THREADS, 0, s>>>(a);
        else list_tile_kernel<VB_HALFVEC, 1><<<n_tiles, LT_THREADS, 0, s>>>(a);
    }
    VB_CUDA(cudaGetLastError());
    count_launch();
    return VB_OK;
}

int list_tile_rows() { return LT_ROWS; }

}  // namespace vb
