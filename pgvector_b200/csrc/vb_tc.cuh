// vb_tc.cuh -- tcgen05 / TMEM / bulk-copy PTX wrappers and the split-bf16 operand packing shared by the
// tensor-core kernels (vb_assign_tc.cu, vb_list_tc.cu).
#pragma once
#include "vb_common.cuh"

#include <cuda_bf16.h>

namespace vb {

constexpr int TC_K = 64;        // bf16 elements per K block (128 bytes = one swizzle atom)

// ----------------------------------------------------------------------------- PTX wrappers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one lane of a converged warp (elect.sync): the compiler keeps code guarded by it on the uniform datapath
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp reads TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzled operand tile: 8-row atoms of 1024 bytes (SBO), one atom along K (LBO unused)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                              // leading byte offset (single atom on K)
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;        // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N = 256
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4)            // c_format = F32
           | (1u << 7)          // a_format = BF16
           | (1u << 10)         // b_format = BF16
           | (0u << 15) | (0u << 16)   // K-major A and B
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------- operand packing

// rows (fp32 or fp16) -> [tile][kblock][plane hi|lo][tile_rows x 64 bf16] in the SWIZZLE_128B image:
// byte offset of (r, kk) inside a plane = r * 128 + (((kk / 8) ^ (r & 7)) * 16) + (kk % 8) * 2
template <int ELEM>
__global__ void pack_planes_kernel(const uint8_t* __restrict__ rows, size_t stride, int64_t row0, int64_t n_valid, int dim,
                                   int tile_rows, int n_kblocks, uint8_t* __restrict__ out, float* __restrict__ sqnorm) {
    // one warp per (row, kblock) pair of chunks: thread = one 16-byte output chunk (8 elements)
    const int64_t chunk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int chunks_per_row = n_kblocks * 8;
    const int64_t r_global = chunk / chunks_per_row;    // row within this slab (padded to tile_rows multiple)
    const int cr = (int)(chunk % chunks_per_row);
    const int kb = cr / 8, c = cr % 8;
    const int64_t tile = r_global / tile_rows;
    const int r = (int)(r_global % tile_rows);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (r_global < n_valid) {
        const uint8_t* src = rows + (size_t)(row0 + r_global) * stride;
        const int e0 = kb * TC_K + c * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int e = e0 + j;
            if (e < dim) v[j] = ELEM == VB_VECTOR ? reinterpret_cast<const float*>(src)[e] : __half2float(reinterpret_cast<const __half*>(src)[e]);
        }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
        __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
        __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
        hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t plane_bytes = (size_t)tile_rows * 128;
    uint8_t* base = out + ((size_t)(tile * n_kblocks + kb) * 2) * plane_bytes;
    const size_t off = (size_t)r * 128 + (size_t)((c ^ (r & 7)) * 16);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + plane_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    (void)sqnorm;
}

// |row|^2 in fp32 (one warp per row)
template <int ELEM>
__global__ void row_sqnorm_kernel(const uint8_t* __restrict__ rows, size_t stride, int64_t n, int dim, float* __restrict__ out,
                                  int64_t n_out, float pad_value) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x % 32;
    if (r >= n_out) return;
    float s = 0.f;
    if (r < n) {
        const uint8_t* src = rows + (size_t)r * stride;
        for (int e = lane; e < dim; e += 32) {
            float x = ELEM == VB_VECTOR ? reinterpret_cast<const float*>(src)[e] : __half2float(reinterpret_cast<const __half*>(src)[e]);
            s = fmaf(x, x, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    } else {
        s = pad_value;
    }
    if (lane == 0) out[r] = s;
}

}  // namespace vb
