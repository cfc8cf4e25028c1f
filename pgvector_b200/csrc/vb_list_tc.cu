// vb_list_tc.cu -- the batched IVFFlat list scan on the tensor cores, as a FILTER in front of the exact
// fp32 arithmetic.
//
// GetScanItems (src/ivfscan.c:124-180) needs, per query, the k nearest of ~probes * rows/lists candidates.
// The list-major formulation (vb_list_tile.cu) already reads every probed list once per batch; what is left
// is 2 fp32 instructions per (row, query, dimension).  Here that arithmetic moves to tcgen05:
//
//   1. approximate pass:  d~ = |x|^2 + |q|^2 - 2 x.q   (or -x.q), x.q from three bf16 MMAs on the hi/lo split of
//      both operands (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM) -- |d~ - d| <= eps(q), a rigorous bound;
//   2. the k' = k + slack smallest d~ of every query are selected (segment_topk_kernel);
//   3. of those, the candidates with d~ <= (k-th d~) + 2 eps are re-scored with the exact scan arithmetic (Acc<>,
//      same as scan_kernel) -- nothing above that threshold can be among the k nearest;
//   4. the k nearest by (exact distance, position) are emitted, and the query is CERTIFIED when the k'-th d~ is
//      itself above the threshold (then so is every candidate that was not selected), i.e. the result equals the
//      full exact scan.  A batch with an uncertified query is re-run on the exact kernel.
//
// Layout.  The index rows are packed once per load into bf16 hi/lo planes in the 128-byte-swizzled shared-memory
// image, one 32 KB block per (128-row table tile, 64-dimension block); a unit of work is a (list, table tile)
// pair (tiles straddling a list boundary are visited by both lists, rows outside the list masked).  The queries
// of each list's group are gathered and packed per batch into 64-query B tiles (16 KB per dimension block).
// CTA = persistent, one per SM: warp 0 = bulk-copy producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (thread = row; two 128-column accumulator stages so the epilogue of one tile overlaps the
// MMAs of the next).  The query operand of a tile is ONE shared-memory tile [q_hi ; q_lo] of 2n rows (n = 32 or 64):
// x_hi . [q_hi ; q_lo] is a single UMMA 128 x 2n x 16 per K step whose two column groups the epilogue adds;
// level 2 adds x_lo . q_hi (128 x n x 16).
//
// Two filter levels.  Level 1 streams only the hi plane of the rows (16 KB + B per stage, 7 stages in flight): half
// the HBM traffic, error bound 2^-7 |x||q|.  A batch with an uncertified query is repeated at level 2 (both planes,
// 4 x 48 KB stages, bound 2^-12 |x||q|), and only then on the exact kernel; after a level-1 failure the level rests
// for 64 batches (whether it certifies is a property of the data, not of the batch).
//
// Roofline: HBM -- one pass over the packed hi planes (level 1: 2 bytes per row element) or both planes (level 2:
// 4 bytes) of the probed lists per batch.
#include "vb_tc.cuh"
#include "vb_slab_select.cuh"
#include "vb_distance.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace vb {

constexpr int LC_M = 128;
constexpr int LC_N = 64;
constexpr int LC_STAGES = 4;                             // level 2: 4 x 48 KB in flight per SM
constexpr int LC_STAGES_L1 = 7;                          // level 1: 7 x 32 KB (A hi plane + B)
constexpr int LC_MAX_STAGES = 8;
constexpr int LC_THREADS = 256;
constexpr uint32_t LC_A_PLANE = LC_M * TC_K * 2;        // 16 KB
constexpr uint32_t LC_B_PLANE = LC_N * TC_K * 2;        // 8 KB
constexpr uint32_t LC_A_STAGE = 2 * LC_A_PLANE;
constexpr uint32_t LC_B_STAGE = 2 * LC_B_PLANE;
constexpr uint32_t LC_STAGE = LC_A_STAGE + LC_B_STAGE;  // 48 KB
constexpr uint32_t LC_STAGE_L1 = LC_A_PLANE + LC_B_STAGE;  // 32 KB
constexpr size_t LC_SMEM = std::max((size_t)LC_STAGES * LC_STAGE, (size_t)LC_STAGES_L1 * LC_STAGE_L1) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int LC_ACC = 2 * LC_N;                         // TMEM columns of one accumulator stage: [x.q_hi | x.q_lo]
constexpr int LC_MAX_KP = 128;                           // candidates selected per query, at most

struct LcArgs {
    const uint8_t* A;          // table planes [tile][kb][2][128 x 64]
    const uint8_t* B;          // query-group planes [gtile][kb][2][64 x 64]
    const ListUnit* units;
    int n_units;
    const int64_t* list_off;
    const int32_t* grp_begin;
    const int32_t* grp_cnt;
    const int32_t* gt_begin;
    const int32_t* pair_q;
    const int64_t* pair_out;
    const float* xn;           // |x|^2 per table row
    const float* qn;           // |q|^2 per query of the batch
    float* out;
    float* smin;               // slab minima (slab_base(), vb_common.cuh) or nullptr
    const int32_t* pair_sbase;
    int n_kblocks;
    int is_l2;
    int hi_only;               // level 1: only the hi plane of the rows is read and multiplied (x_hi . (q_hi + q_lo))
    int uniform_nqt;           // > 0: every unit has this many query tiles and a job is one (unit, query tile) pair --
                               // the centre scan, where ONE "list" (the centre table) is probed by every query
    int n_jobs;
};

struct LcJob {
    ListUnit un;
    int cnt, q_lo, q_hi;
};
// job j of the static round-robin: a (list, table tile) unit with all its query tiles, or one query tile of it
__device__ __forceinline__ bool lc_job(const LcArgs& a, int j, LcJob& jb) {
    jb.un = a.units[a.uniform_nqt ? j / a.uniform_nqt : j];
    jb.cnt = a.grp_cnt[jb.un.list];
    if (jb.cnt == 0) return false;
    const int nqt = (jb.cnt + LC_N - 1) / LC_N;
    jb.q_lo = a.uniform_nqt ? j % a.uniform_nqt : 0;
    jb.q_hi = a.uniform_nqt ? min(jb.q_lo + 1, nqt) : nqt;
    return jb.q_lo < jb.q_hi;
}

__global__ void __launch_bounds__(LC_THREADS, 1) list_tc_kernel(LcArgs a) {
    extern __shared__ uint8_t lc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lc_smem_raw) + 1023) & ~(uintptr_t)1023);
    // stage ring: level 2 = 4 stages of (A hi+lo 32 KB | B 16 KB), level 1 = 7 stages of (A hi 16 KB | B 16 KB) --
    // the bytes in flight per SM are what keeps HBM busy
    const int n_stages = a.hi_only ? LC_STAGES_L1 : LC_STAGES;
    const uint32_t stage_bytes = a.hi_only ? LC_STAGE_L1 : LC_STAGE;
    const uint32_t a_bytes = a.hi_only ? LC_A_PLANE : LC_A_STAGE;   // the hi plane leads each 32 KB block of the image
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + LC_MAX_STAGES;
    uint64_t* tfull_bar = bars + 2 * LC_MAX_STAGES;
    uint64_t* tempty_bar = bars + 2 * LC_MAX_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * LC_MAX_STAGES + 4);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 2 * LC_ACC);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer (whole warp converged, one elected lane issues the copies) =====
        const bool leader = elect_one();
        uint32_t it = 0;
        for (int j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
            LcJob jb;
            if (!lc_job(a, j, jb)) continue;
            const ListUnit un = jb.un;
            const int gt0 = a.gt_begin[un.list];
            for (int qt = jb.q_lo; qt < jb.q_hi; ++qt)
            {
                // a query tile with at most 32 queries is multiplied as N = 32: only the first half of each B plane moves
                const bool n32 = jb.cnt - qt * LC_N <= 32;
                for (int kb = 0; kb < a.n_kblocks; ++kb, ++it) {
                    const int s = it % n_stages;
                    const uint32_t ph = (it / n_stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    const uint8_t* gb = a.B + ((size_t)(gt0 + qt) * a.n_kblocks + kb) * LC_B_STAGE;
                    if (leader) {
                        mbar_arrive_expect_tx(&full_bar[s], a_bytes + (n32 ? LC_B_STAGE / 2 : LC_B_STAGE));
                        bulk_g2s(sa, a.A + ((size_t)un.tile * a.n_kblocks + kb) * LC_A_STAGE, a_bytes, &full_bar[s]);
                        if (n32) {   // [q_hi rows 0..31 | q_lo rows 0..31] back to back = one 64-row operand
                            bulk_g2s(sb, gb, LC_B_PLANE / 2, &full_bar[s]);
                            bulk_g2s(sb + LC_B_PLANE / 2, gb + LC_B_PLANE, LC_B_PLANE / 2, &full_bar[s]);
                        } else {
                            bulk_g2s(sb, gb, LC_B_STAGE, &full_bar[s]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (warp-uniform control flow and operands), one elected lane
        // issues.  Under "lane == 0" the compiler wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop and the
        // lone thread's scalar code (~160 instructions per K block) becomes the bottleneck of the kernel. =====
        const bool leader = elect_one();
        // One UMMA costs ~130 cycles here whatever its N (it re-reads the 128 x 16 A tile from shared memory), so the
        // products are merged along N: B = [q_hi ; q_lo] is ONE operand of 2n rows, and x_hi . [q_hi ; q_lo] lands in the
        // column groups [0, n) and [n, 2n) of the accumulator with a single instruction per K step; level 2 adds
        // x_lo . q_hi into [0, n).  The epilogue sums the two groups.
        constexpr uint32_t idesc128 = make_idesc_bf16(LC_M, 128);
        constexpr uint32_t idesc64 = make_idesc_bf16(LC_M, 64);
        constexpr uint32_t idesc32 = make_idesc_bf16(LC_M, 32);
        uint32_t it = 0, tile = 0;
        for (int j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
            LcJob jb;
            if (!lc_job(a, j, jb)) continue;
            for (int qt = jb.q_lo; qt < jb.q_hi; ++qt, ++tile) {
                const int as = tile & 1;
                const uint32_t aph = (tile >> 1) & 1;
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)as * LC_ACC;
                const bool n32 = jb.cnt - qt * LC_N <= 32;
                const uint32_t idesc_both = n32 ? idesc64 : idesc128;   // N = 2n
                const uint32_t idesc_one = n32 ? idesc32 : idesc64;     // N = n
                for (int kb = 0; kb < a.n_kblocks; ++kb, ++it) {
                    const int s = it % n_stages;
                    const uint32_t ph = (it / n_stages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
                    const uint64_t da_hi = make_sw128_desc(sa), da_lo = make_sw128_desc(sa + LC_A_PLANE);
                    const uint64_t db = make_sw128_desc(sb);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < TC_K / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                            umma_bf16(tmem_d, da_hi + adv, db + adv, idesc_both, (kb | k) != 0);
                            if (!a.hi_only) umma_bf16(tmem_d, da_lo + adv, db + adv, idesc_one, 1);
                        }
                        umma_commit(&empty_bar[s]);
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(&tfull_bar[as]);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = one table row of the tile; columns = queries of the group =====
        const int qr = warp - 4;
        uint32_t tile = 0;
        for (int j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
            LcJob jb;
            if (!lc_job(a, j, jb)) continue;
            const ListUnit un = jb.un;
            const int cnt = jb.cnt;
            const int64_t lo = a.list_off[un.list], hi = a.list_off[un.list + 1];
            const int64_t r_table = (int64_t)un.tile * LC_M + qr * 32 + lane;
            const bool valid_row = r_table >= lo && r_table < hi;
            const float xnr = valid_row && a.is_l2 ? a.xn[r_table] : 0.f;
            const int gb = a.grp_begin[un.list];
            // this warp's 32 rows are one table-aligned slab of the list (if any of them belongs to it)
            const bool slabs = a.smin != nullptr && __ballot_sync(0xffffffffu, valid_row) != 0;
            const int slab_local = (int)((((int64_t)un.tile * LC_M + qr * 32) >> 5) - (lo >> 5));
            for (int qt = jb.q_lo; qt < jb.q_hi; ++qt, ++tile) {
                const int as = tile & 1;
                const uint32_t aph = (tile >> 1) & 1;
                mbar_wait(&tfull_bar[as], aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(qr * 32) << 16) + (uint32_t)as * LC_ACC;
                const int n = cnt - qt * LC_N <= 32 ? 32 : LC_N;   // queries per column group of this tile
#pragma unroll
                for (int c0 = 0; c0 < LC_N; c0 += 32) {
                    const int col0 = qt * LC_N + c0;
                    if (col0 >= cnt) break;   // warp-uniform
                    uint32_t acc[32], part[32];
                    tmem_ld32(taddr + c0, acc);          // x_hi . q_hi (+ x_lo . q_hi)
                    tmem_ld32(taddr + n + c0, part);     // x_hi . q_lo
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(part[j]));
                    // lane j fetches the bookkeeping of column j once; broadcast in the loop
                    const int myc = col0 + lane;
                    int64_t my_out = 0;
                    float my_qn = 0.f;
                    if (myc < cnt) {
                        my_out = a.pair_out[gb + myc];
                        if (a.is_l2) my_qn = a.qn[a.pair_q[gb + myc]];
                    }
                    float my_min = __int_as_float(0x7F800000);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t po = __shfl_sync(0xffffffffu, my_out, j);
                        const float qn = __shfl_sync(0xffffffffu, my_qn, j);
                        float val = __int_as_float(0x7F800000);
                        if (col0 + j < cnt && valid_row) {
                            const float dot = __uint_as_float(acc[j]);
                            val = a.is_l2 ? fmaf(-2.f, dot, xnr + qn) : -dot;
                            a.out[po + (r_table - lo)] = val;
                        }
                        if (slabs) {
                            // minimum of column j over the warp's 32 rows (one CREDUX), kept by lane j
                            float m;
                            asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(val));
                            if (lane == j) my_min = m;
                        }
                    }
                    if (slabs && myc < cnt) a.smin[a.pair_sbase[gb + myc] + slab_local] = my_min;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
            }
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * LC_ACC);
}

// gather + split the queries of every (query, list) pair into the B tiles of its list's group:
// thread = one 16-byte chunk (8 dimensions) of one pair
__global__ void pack_groups_kernel(const float* __restrict__ qimg, size_t qstride, int dim, int n_kblocks, int64_t n_pairs,
                                   const int32_t* __restrict__ pair_q, const int32_t* __restrict__ pair_list,
                                   const int32_t* __restrict__ grp_begin, const int32_t* __restrict__ gt_begin,
                                   uint8_t* __restrict__ out) {
    const int64_t chunk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int chunks_per_row = n_kblocks * 8;
    const int64_t slot = chunk / chunks_per_row;
    if (slot >= n_pairs) return;
    const int cr = (int)(chunk % chunks_per_row);
    const int kb = cr / 8, c = cr % 8;
    const int l = pair_list[slot];
    const int j = (int)(slot - grp_begin[l]);
    const int64_t gtile = gt_begin[l] + j / LC_N;
    const int r = j % LC_N;
    const float* src = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(qimg) + (size_t)pair_q[slot] * qstride);
    const int e0 = kb * TC_K + c * 8;
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = e0 + t < dim ? src[e0 + t] : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * t]), h1 = __float2bfloat16_rn(v[2 * t + 1]);
        __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * t] - __bfloat162float(h0));
        __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * t + 1] - __bfloat162float(h1));
        hi[t] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lo[t] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    uint8_t* base = out + ((size_t)(gtile * n_kblocks + kb) * 2) * LC_B_PLANE;
    const size_t off = (size_t)r * 128 + (size_t)((c ^ (r & 7)) * 16);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + LC_B_PLANE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// |d~ - d_fp32| <= eps(q) for every candidate of query q
struct LcBound {
    int is_l2;
    float c_dot, c_sum, xmax;
};
__device__ __forceinline__ float lc_eps(const LcBound& b, float qn) {
    return b.c_dot * sqrtf(qn) * b.xmax + (b.is_l2 ? b.c_sum * (b.xmax * b.xmax + qn) : 0.f);
}
// Only candidates with d~ <= (k-th smallest d~) + 2 eps can belong to the k nearest: the k candidates with the
// smallest d~ all have d <= (k-th d~) + eps, and anything above the threshold has d > (k-th d~) + eps.
__device__ __forceinline__ float lc_threshold(const LcBound& b, float qn, const float* approx_q, int k, int kp) {
    return approx_q[min(k, kp) - 1] + 2.f * lc_eps(b, qn);
}

// exact distance of the candidates under the threshold: one warp per (query, candidate), the arithmetic of the
// scan kernels; the others get +inf (they sort behind every re-scored one)
template <int ELEM, int METRIC>
__global__ void rescore_kernel(const uint8_t* __restrict__ rows, size_t stride, int V, const uint8_t* __restrict__ qimg,
                               size_t qstride, int64_t nq, int k, int kp, int probes, LcBound bound, const float* __restrict__ qn,
                               const int32_t* __restrict__ pos, const float* __restrict__ approx,
                               const int32_t* __restrict__ probe_lists, const int32_t* __restrict__ cand_off,
                               const int64_t* __restrict__ list_off, float* __restrict__ exact) {
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x % 32;
    if (w >= nq * kp) return;
    const int64_t q = w / kp;
    const int32_t ps = pos[w];
    if (ps < 0 || approx[w] > lc_threshold(bound, qn[q], approx + q * kp, k, kp)) {   // NaN compares false: re-scored
        if (lane == 0) exact[w] = __int_as_float(0x7F800000);
        return;
    }
    const int32_t* co = cand_off + q * (probes + 1);
    int lo = 0, hi = probes;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (co[mid] <= ps) lo = mid;
        else hi = mid;
    }
    while (lo + 1 < probes && co[lo + 1] <= ps) ++lo;   // empty lists share an offset
    const int l = probe_lists[q * probes + lo];
    const int64_t row = list_off[l] + (ps - co[lo]);
    const uint4* rp = reinterpret_cast<const uint4*>(rows + (size_t)row * stride);
    const uint4* sq = reinterpret_cast<const uint4*>(qimg + (size_t)q * qstride);
    Acc<ELEM, METRIC> acc;
#pragma unroll 4
    for (int v = lane; v < V; v += 32) acc.add(__ldg(rp + v), sq, v);
    acc.template reduce<32>();
    if (lane == 0) exact[w] = (float)acc.value();
}

// one warp per query: order the re-scored candidates by (exact distance, position), emit the first k, and check
// the certificate: the k'-th approximate distance lies above the threshold, so no candidate outside the k' can
// be under it either (or every candidate of the query was among the k')
__global__ void certify_kernel(int64_t nq, int k, int kp, LcBound bound,
                               const float* __restrict__ qn, const int32_t* __restrict__ seg_len,
                               const int32_t* __restrict__ pos_kp, const float* __restrict__ approx_kp,
                               const float* __restrict__ exact_kp, int32_t* __restrict__ out_pos, float* __restrict__ out_key,
                               int* __restrict__ n_failed, uint8_t* __restrict__ failed) {
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x % 32;
    if (q >= nq) return;
    __shared__ uint64_t s_key[8][LC_MAX_KP];
    uint64_t* keys = s_key[(threadIdx.x / 32) % 8];
    for (int i = lane; i < LC_MAX_KP; i += 32) {
        uint64_t key = 0xFFFFFFFF00000000ull | (uint32_t)i;   // absent entries sort last, distinct
        if (i < kp) {
            const int32_t p = pos_kp[q * kp + i];
            if (p >= 0) key = ((uint64_t)orderable_key(exact_kp[q * kp + i]) << 32) | (uint32_t)p;
        }
        keys[i] = key;
    }
    __syncwarp();
    for (int i = lane; i < kp; i += 32) {
        const uint64_t mine = keys[i];
        int rank = 0;
        for (int j = 0; j < kp; ++j) rank += keys[j] < mine;
        if (rank < k) {
            const bool present = pos_kp[q * kp + i] >= 0;
            out_pos[q * k + rank] = present ? (int32_t)(uint32_t)mine : -1;
            out_key[q * k + rank] = present ? key_to_float((uint32_t)(mine >> 32)) : __int_as_float(0x7F800000);
        }
    }
    if (lane == 0) {
        bool ok = true;
        if (seg_len[q] > kp)   // candidates beyond the k' exist: the last of the k' must already be above the threshold
            ok = approx_kp[q * kp + kp - 1] > lc_threshold(bound, qn[q], approx_kp + q * kp, k, kp);   // false for NaN
        failed[q] = ok ? 0 : 1;
        if (!ok) atomicAdd(n_failed, 1);
    }
}

// ---- the three steps after the approximate pass in ONE kernel (one warp per query) -----------------------------------
//
// select : the k' smallest approximate distances of the query's candidate run by (distance, position) -- the same
//          composite key as segment_topk_kernel.  The warp keeps the k' best as a sorted list spread over its lanes
//          (R = k' / 32 registers per lane, rank i at register i / 32, lane i % 32) and streams the run 32 candidates at
//          a time; only candidates under the current k'-th key are inserted (~k' ln(n / k') insertions).
// re-score: the candidates under the certificate threshold are a PREFIX of that sorted list; each is re-scored with
//          the scan arithmetic (Acc<>, one row per warp pass, the loop of rescore_kernel), two rows in flight.
// certify: rank by (exact distance, position), emit the first k, and check the certificate -- certify_kernel's rules.
// Results are bit-identical to segment_topk_kernel + rescore_kernel + certify_kernel (tests compare the two paths).
template <int R>
struct WarpTopList {
    uint64_t key[R];
    __device__ __forceinline__ void fill(uint64_t v) {
#pragma unroll
        for (int r = 0; r < R; ++r) key[r] = v;
    }
    // insert x (known to be smaller than the current last key); the displaced last key is dropped
    __device__ __forceinline__ void insert(uint64_t x, int lane) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned gt = __ballot_sync(0xffffffffu, key[r] > x);
            if (gt == 0) continue;                       // warp-uniform
            const int first = __ffs(gt) - 1;
            const uint64_t carry = __shfl_sync(0xffffffffu, key[r], 31);
            const uint64_t up = __shfl_up_sync(0xffffffffu, key[r], 1);
            if (lane > first) key[r] = up;
            else if (lane == first) key[r] = x;
            x = carry;
        }
    }
    __device__ __forceinline__ uint64_t at(int i) const {   // rank i, warp-uniform i
        uint64_t v = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (i / 32 == r) v = __shfl_sync(0xffffffffu, key[r], i % 32);
        return v;
    }
};

constexpr int SR_WARPS = 4;
#ifndef VB_SR_PARALLEL_ADDR
#define VB_SR_PARALLEL_ADDR 1
#endif

template <int ELEM, int METRIC, int R>
__global__ void __launch_bounds__(SR_WARPS * 32) select_refine_kernel(const uint8_t* __restrict__ rows, size_t stride, int V,
                                                                      const uint8_t* __restrict__ qimg, size_t qstride, int64_t nq, int k,
                                                                      int kp, int probes, LcBound bound, const float* __restrict__ qn,
                                                                      const float* __restrict__ dist, const int64_t* __restrict__ seg_begin,
                                                                      const int32_t* __restrict__ seg_len,
                                                                      const int32_t* __restrict__ probe_lists,
                                                                      const int32_t* __restrict__ cand_off,
                                                                      const int64_t* __restrict__ list_off, int32_t* __restrict__ out_pos,
                                                                      float* __restrict__ out_key, int* __restrict__ n_failed,
                                                                      const int32_t* __restrict__ pre_pos,
                                                                      const float* __restrict__ pre_key) {
    extern __shared__ uint4 sr_smem[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int qvec = (int)(qstride / 16);
    uint4* sq = sr_smem + (size_t)warp * qvec;
    const int64_t q = blockIdx.x * (int64_t)SR_WARPS + warp;
    if (q >= nq) return;
    const uint4* gq = reinterpret_cast<const uint4*>(qimg + (size_t)q * qstride);
    for (int i = lane; i < qvec; i += 32) sq[i] = gq[i];

    // ---- select
    const float* dp = dist + seg_begin[q];
    const int n = seg_len[q];
    WarpTopList<R> top;
    top.fill(~0ull);
    // the sentinel ~0ull sorts after every real key (position < 2^32 - 1), so the first k' candidates simply displace it
    uint64_t thr = ~0ull;                                   // current k'-th key
    const int kl = (kp - 1) % 32, kr = (kp - 1) / 32;
    constexpr int UNR = 4;
    if (pre_pos != nullptr) {
        // the k' nearest were selected by segment_topk_kernel (sorted by (distance, position), -1 padded): load them
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = r * 32 + lane;
            if (i < kp) {
                const int32_t p = pre_pos[q * kp + i];
                if (p >= 0) top.key[r] = ((uint64_t)orderable_key(pre_key[q * kp + i]) << 32) | (uint32_t)p;
            }
        }
        uint64_t t = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r == kr) t = __shfl_sync(0xffffffffu, top.key[r], kl);
        thr = t;
    }
    for (int base = 0; pre_pos == nullptr && base < n; base += 32 * UNR) {
        float v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = base + u * 32 + lane;
            v[u] = i < n ? dp[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = base + u * 32 + lane;
            const uint64_t key = i < n ? (((uint64_t)orderable_key(v[u]) << 32) | (uint32_t)i) : ~0ull;
            unsigned m = __ballot_sync(0xffffffffu, key < thr);
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const uint64_t x = __shfl_sync(0xffffffffu, key, j);
                if (x < thr) {                              // warp-uniform
                    top.insert(x, lane);
                    uint64_t t = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r == kr) t = __shfl_sync(0xffffffffu, top.key[r], kl);
                    thr = t;
                }
            }
        }
    }
    __syncwarp();

    // ---- threshold: (k-th smallest approximate distance) + 2 eps; the candidates under it are a prefix of the list
    const float qnq = qn[q];
    const int have = min(n, kp);                            // real entries in the list
    const int kth = min(k, kp) - 1;
    const uint64_t kth_key = top.at(kth);
    const float kth_approx = kth_key == ~0ull ? __int_as_float(0x7F800000) : key_to_float((uint32_t)(kth_key >> 32));
    const float T = kth_approx + 2.f * lc_eps(bound, qnq);
    // entries of rank i: approx_i <= T  <=>  not (approx_i > T)   (NaN compares false: re-scored, like rescore_kernel)
    float exact[R];
#pragma unroll
    for (int r = 0; r < R; ++r) exact[r] = __int_as_float(0x7F800000);
    const int32_t* co = cand_off + q * (probes + 1);
#if VB_SR_PARALLEL_ADDR
    // Row address of every listed candidate, one candidate per lane (R per lane): position -> probe (binary search of the
    // query's candidate offsets) -> list -> row.  That is five dependent loads; done inside the re-score loop they were
    // paid once per PAIR of candidates, ahead of the row reads, by the whole warp.
    const uint8_t* rowp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        const uint64_t ki = top.key[r];
        rowp[r] = rows;
        if (i < have && ki != ~0ull) {
            const int32_t ps = (int32_t)(uint32_t)ki;
            int lo = 0, hi = probes;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (co[mid] <= ps) lo = mid;
                else hi = mid;
            }
            while (lo + 1 < probes && co[lo + 1] <= ps) ++lo;   // empty lists share an offset
            const int l = probe_lists[q * probes + lo];
            rowp[r] = rows + (size_t)(list_off[l] + (ps - co[lo])) * stride;
        }
    }
#endif
    for (int i0 = 0; i0 < have; i0 += 2) {
        const uint64_t k0 = top.at(i0);
        const uint64_t k1 = i0 + 1 < have ? top.at(i0 + 1) : ~0ull;
        const float a0 = key_to_float((uint32_t)(k0 >> 32));
        const float a1 = k1 == ~0ull ? 0.f : key_to_float((uint32_t)(k1 >> 32));
        const bool do0 = !(a0 > T), do1 = k1 != ~0ull && !(a1 > T);
        if (!do0 && !do1) continue;                         // (no early exit: NaN distances sort last and are re-scored too)
        const uint4* rp[2];
#if VB_SR_PARALLEL_ADDR
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int it = min(i0 + t, R * 32 - 1);         // (warp-uniform; the clamp only guards the shuffle of an absent k1)
            unsigned long long a = 0;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (it / 32 == r) a = __shfl_sync(0xffffffffu, (unsigned long long)rowp[r], it % 32);
            rp[t] = reinterpret_cast<const uint4*>((t == 0 ? do0 : do1) ? (const uint8_t*)a : rows);
        }
#else
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int32_t ps = (int32_t)(uint32_t)(t == 0 ? k0 : k1);
            int lo = 0, hi = probes;
            if ((t == 0 ? do0 : do1)) {
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (co[mid] <= ps) lo = mid;
                    else hi = mid;
                }
                while (lo + 1 < probes && co[lo + 1] <= ps) ++lo;   // empty lists share an offset
                const int l = probe_lists[q * probes + lo];
                rp[t] = reinterpret_cast<const uint4*>(rows + (size_t)(list_off[l] + (ps - co[lo])) * stride);
            } else {
                rp[t] = reinterpret_cast<const uint4*>(rows);
            }
        }
#endif
        Acc<ELEM, METRIC> acc0, acc1;
        if (do0 && do1) {
#pragma unroll 4
            for (int v = lane; v < V; v += 32) {
                const uint4 x0 = __ldg(rp[0] + v), x1 = __ldg(rp[1] + v);
                acc0.add(x0, sq, v);
                acc1.add(x1, sq, v);
            }
        } else if (do0) {
#pragma unroll 4
            for (int v = lane; v < V; v += 32) acc0.add(__ldg(rp[0] + v), sq, v);
        } else {
#pragma unroll 4
            for (int v = lane; v < V; v += 32) acc1.add(__ldg(rp[1] + v), sq, v);
        }
        acc0.template reduce<32>();
        acc1.template reduce<32>();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (do0 && i0 / 32 == r && lane == i0 % 32) exact[r] = (float)acc0.value();
            if (do1 && (i0 + 1) / 32 == r && lane == (i0 + 1) % 32) exact[r] = (float)acc1.value();
        }
    }

    // ---- order by (exact distance, position), emit the first k
    uint64_t fin[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        const bool present = i < kp && top.key[r] != ~0ull;    // (ranks >= k' hold displaced leftovers, not candidates)
        fin[r] = present ? (((uint64_t)orderable_key(exact[r]) << 32) | (uint32_t)top.key[r]) : (0xFFFFFFFF00000000ull | (uint32_t)i);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int rank = 0;
#pragma unroll
        for (int r2 = 0; r2 < R; ++r2)
            for (int j = 0; j < 32; ++j) {
                const uint64_t o = __shfl_sync(0xffffffffu, fin[r2], j);
                rank += (r2 * 32 + j < kp) && o < fin[r];
            }
        const int i = r * 32 + lane;
        if (i < kp && rank < k) {
            const bool present = top.key[r] != ~0ull;
            out_pos[q * k + rank] = present ? (int32_t)(uint32_t)fin[r] : -1;
            out_key[q * k + rank] = present ? key_to_float((uint32_t)(fin[r] >> 32)) : __int_as_float(0x7F800000);
        }
    }
    // ---- certificate: candidates beyond the k' exist -> the last of the k' must already be above the threshold
    if (lane == 0) {
        bool ok = true;
        if (n > kp) ok = key_to_float((uint32_t)(thr >> 32)) > T;     // thr = the k'-th key; false for NaN
        if (!ok) atomicAdd(n_failed, 1);
    }
}

// Steps 2 + 3 + 4 with ONE CTA per query (cta_refine_kernel): the selection is CTA-wide (slab minima, or the whole run
// when it is short: the distances of a query to every centre), the candidates under the certificate threshold are then
// re-scored by the CTA's eight warps in parallel (one row per warp at a time; select_refine_kernel walks them two at a
// time on a single warp), ranked and certified.  Same arithmetic, same tie rule, same outputs as select_refine_kernel;
// a query whose selection overflows its buffer (ties by the thousand) counts as uncertified and the batch is repeated
// on the kernels above.
template <int ELEM, int METRIC>
__global__ void __launch_bounds__(SS_THREADS) cta_refine_kernel(const uint8_t* __restrict__ rows, size_t stride, int V,
                                                                const uint8_t* __restrict__ qimg, size_t qstride, int k, int kp, int probes,
                                                                LcBound bound, const float* __restrict__ qn, const float* __restrict__ dist,
                                                                const float* __restrict__ smin, int64_t cap, int64_t cap_s,
                                                                const int32_t* __restrict__ seg_len, const int32_t* __restrict__ probe_lists,
                                                                const int32_t* __restrict__ cand_off, const int64_t* __restrict__ list_off,
                                                                int32_t* __restrict__ out_pos, float* __restrict__ out_key,
                                                                int* __restrict__ n_failed) {
    extern __shared__ uint64_t cr_smem[];
    uint64_t* cand = cr_smem;                                                   // [SS_CAND]
    uint64_t* fin = cand + SS_CAND;                                             // [kp] final keys
    const uint8_t** rowp = reinterpret_cast<const uint8_t**>(fin + kp);         // [kp] row addresses
    uint4* sq = reinterpret_cast<uint4*>(rowp + kp);                            // [qvec] query image (cand, fin and rowp are 16 kp + 16384 bytes: aligned)
    const int qvec = (int)(qstride / 16);
    void* work = sq + qvec;                                                     // the selection's work area (16-byte aligned)
    float* exact = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(work) + (smin ? ss_select_smem_bytes(cap_s, probes) : (size_t)cap * 4));   // [kp]
    const int q = blockIdx.x;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const uint4* gq = reinterpret_cast<const uint4*>(qimg + (size_t)q * qstride);
    for (int i = tid; i < qvec; i += SS_THREADS) sq[i] = gq[i];
    const int n_run = seg_len[q];
    int n;
    if (smin) n = slab_select_cta(dist, smin, probes, probe_lists, cand_off, list_off, cap, cap_s, q, kp, cand, work);
    else n = direct_select_cta(dist + (int64_t)q * cap, n_run, kp, cand, work);
    if (n < 0) {
        // not selected here: the query reports as uncertified (outputs are rewritten by the repeat of the batch)
        if (tid == 0) atomicAdd(n_failed, 1);
        for (int i = tid; i < k; i += SS_THREADS) {
            out_pos[(int64_t)q * k + i] = -1;
            out_key[(int64_t)q * k + i] = __int_as_float(0x7F800000);
        }
        return;
    }
    const int have = min(n, kp);
    // ---- threshold: (k-th smallest approximate distance) + 2 eps
    const int kth = min(k, kp) - 1;
    const uint64_t kth_key = kth < have ? cand[kth] : ~0ull;
    const float kth_approx = kth_key == ~0ull ? __int_as_float(0x7F800000) : key_to_float((uint32_t)(kth_key >> 32));
    const float T = kth_approx + 2.f * lc_eps(bound, qn[q]);
    const uint64_t thr = kp - 1 < have ? cand[kp - 1] : ~0ull;                    // the k'-th key
    // ---- row address of every listed candidate (one per thread): position -> probe -> row.  After a slab selection the
    // per-probe candidate offsets and list bounds are in shared memory already.
    const int32_t* co = cand_off + (int64_t)q * (probes + 1);
    const SsWork W = ss_work_layout(work, cap_s, probes);
    for (int i = tid; i < kp; i += SS_THREADS) {
        exact[i] = __int_as_float(0x7F800000);
        rowp[i] = rows;
        if (i < have) {
            const int32_t ps = (int32_t)(uint32_t)cand[i];
            if (smin) {
                int lo = 0, hi = probes;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (W.co[mid] <= ps) lo = mid;
                    else hi = mid;
                }
                // (empty lists share an offset with their successor: the search ends on the last of them, the non-empty one)
                rowp[i] = rows + (size_t)(W.lo[lo] + (ps - W.co[lo])) * stride;
            } else {
                int lo = 0, hi = probes;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (co[mid] <= ps) lo = mid;
                    else hi = mid;
                }
                while (lo + 1 < probes && co[lo + 1] <= ps) ++lo;   // empty lists share an offset
                const int l = probe_lists[(int64_t)q * probes + lo];
                rowp[i] = rows + (size_t)(list_off[l] + (ps - co[lo])) * stride;
            }
        }
    }
    __syncthreads();
    // ---- exact re-score of the candidates with approx <= T (NaN compares false: re-scored too), one row per warp
    for (int i = warp; i < have; i += SS_THREADS / 32) {
        const float a = key_to_float((uint32_t)(cand[i] >> 32));
        if (a > T) continue;                                 // warp-uniform
        const uint4* rp = reinterpret_cast<const uint4*>(rowp[i]);
        Acc<ELEM, METRIC> acc;
#pragma unroll 4
        for (int v = lane; v < V; v += 32) acc.add(__ldg(rp + v), sq, v);
        acc.template reduce<32>();
        if (lane == 0) exact[i] = (float)acc.value();
    }
    __syncthreads();
    // ---- order by (exact distance, position), emit the first k
    for (int i = tid; i < kp; i += SS_THREADS)
        fin[i] = i < have ? (((uint64_t)orderable_key(exact[i]) << 32) | (uint32_t)cand[i]) : (0xFFFFFFFF00000000ull | (uint32_t)i);
    __syncthreads();
    for (int i = tid; i < kp; i += SS_THREADS) {
        const uint64_t mine = fin[i];
        int rank = 0;
        for (int j = 0; j < kp; ++j) rank += fin[j] < mine;
        if (rank < k) {
            const bool present = i < have;
            out_pos[(int64_t)q * k + rank] = present ? (int32_t)(uint32_t)mine : -1;
            out_key[(int64_t)q * k + rank] = present ? key_to_float((uint32_t)(mine >> 32)) : __int_as_float(0x7F800000);
        }
    }
    // ---- certificate: candidates beyond the k' exist -> the last of the k' must already be above the threshold
    if (tid == 0 && n_run > kp) {
        const bool ok = key_to_float((uint32_t)(thr >> 32)) > T;     // false for NaN
        if (!ok) atomicAdd(n_failed, 1);
    }
}

// Bytes one launch of list_tc_kernel moves, from the same job list the kernel walks (profiling only):
//   [0] bytes requested by the bulk copies (every (unit, query tile, K block) stage: A tile + B tile),
//   [1] distinct A bytes (each active (list, table tile) unit's planes once: re-reads by further query tiles of the
//       same unit come from L2),   [2] distinct B bytes (each list's query tiles once: shared by the list's units),
//   [3] launches accounted.
__global__ void lc_traffic_kernel(LcArgs a, uint32_t a_bytes, unsigned long long* __restrict__ acc) {
    unsigned long long issued = 0, a_once = 0, b_once = 0;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < a.n_units; u += gridDim.x * blockDim.x) {
        const ListUnit un = a.units[u];
        const int cnt = a.grp_cnt[un.list];
        if (cnt == 0) continue;
        const int nqt = (cnt + LC_N - 1) / LC_N;
        unsigned long long b_list = 0;
        for (int qt = 0; qt < nqt; ++qt) b_list += (unsigned long long)a.n_kblocks * (cnt - qt * LC_N <= 32 ? LC_B_STAGE / 2 : LC_B_STAGE);
        issued += (unsigned long long)nqt * a.n_kblocks * a_bytes + b_list;
        // a table tile that straddles two lists is visited by both units back to back: its second read comes from L2
        if (!(u > 0 && a.units[u - 1].tile == un.tile && a.grp_cnt[a.units[u - 1].list] > 0)) a_once += (unsigned long long)a.n_kblocks * a_bytes;
        if (u == 0 || a.units[u - 1].list != un.list) b_once += b_list;
    }
    atomicAdd(&acc[0], issued);
    atomicAdd(&acc[1], a_once);
    atomicAdd(&acc[2], b_once);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&acc[3], 1ull);
}

// ----------------------------------------------------------------------------- host side

enum { WSC_B = 21, WSC_N = 22, WSC_K = 23 };

static unsigned long long* g_traffic = nullptr;   // device accumulators of lc_traffic_kernel, per filter use (0 = lists, 1 = centres)
static bool g_traffic_on = false;

bool list_tc_supported(int elem, int key_metric, int k) {
    return (elem == VB_VECTOR || elem == VB_HALFVEC) && (key_metric == VB_L2_SQUARED || key_metric == VB_NEG_IP) && k >= 1 &&
           list_tc_kp(k, 2) <= LC_MAX_KP;
}

int list_tc_kp(int k, int level) {
    // candidates kept per query.  Only those under the threshold are re-scored, so a generous k' costs a slightly
    // larger selection, not more exact distances; the certificate fails only when ALL k' are under the threshold.
    if (level == 1) return k <= 10 ? 64 : k <= 40 ? 128 : 1 << 20;
    return k <= 10 ? 32 : k <= 24 ? 48 : k <= 40 ? 64 : 1 << 20;
}

// packed planes + norms of the whole table (once per index load)
int list_tc_prepare(const Table& rows, ListTcImage* im) {
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const int64_t n = rows.n;
    const int n_kblocks = (rows.dim + TC_K - 1) / TC_K;
    const int64_t n_tiles = (n + LC_M - 1) / LC_M;
    const size_t bytes = (size_t)n_tiles * n_kblocks * LC_A_STAGE;
    VB_CUDA(cudaMalloc(&im->planes, std::max<size_t>(bytes, 16)));
    VB_CUDA(cudaMalloc(&im->xn, sizeof(float) * (size_t)std::max<int64_t>(n_tiles * LC_M, 1)));
    im->n_kblocks = n_kblocks;
    im->n_tiles = n_tiles;
    if (n == 0) {
        im->xmax = 0.f;
        return VB_OK;
    }
    const int64_t chunks = n_tiles * LC_M * n_kblocks * 8;
    const unsigned grid = (unsigned)((chunks + 255) / 256);
    const unsigned g2 = (unsigned)((n_tiles * LC_M * 32 + 255) / 256);
    if (rows.elem == VB_VECTOR) {
        pack_planes_kernel<VB_VECTOR><<<grid, 256, 0, s>>>(rows.d, rows.stride, 0, n, rows.dim, LC_M, n_kblocks, im->planes, nullptr);
        row_sqnorm_kernel<VB_VECTOR><<<g2, 256, 0, s>>>(rows.d, rows.stride, n, rows.dim, im->xn, n_tiles * LC_M, 0.f);
    } else {
        pack_planes_kernel<VB_HALFVEC><<<grid, 256, 0, s>>>(rows.d, rows.stride, 0, n, rows.dim, LC_M, n_kblocks, im->planes, nullptr);
        row_sqnorm_kernel<VB_HALFVEC><<<g2, 256, 0, s>>>(rows.d, rows.stride, n, rows.dim, im->xn, n_tiles * LC_M, 0.f);
    }
    VB_CUDA(cudaGetLastError());
    count_launch(2);
    std::vector<float> h((size_t)n);
    VB_CUDA(cudaMemcpyAsync(h.data(), im->xn, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaStreamSynchronize(s));
    float m2 = 0.f;
    bool finite = true;
    for (float v : h) {
        if (!(v == v) || v > 3.0e38f) finite = false;
        else m2 = std::max(m2, v);
    }
    im->xmax = std::sqrt(m2);
    im->finite = finite;
    return VB_OK;
}

void list_tc_release(ListTcImage* im) {
    if (im->planes) cudaFree(im->planes);
    if (im->xn) cudaFree(im->xn);
    if (im->units) cudaFree(im->units);
    *im = ListTcImage{};
}

// approximate pass: fills `out` (the per-query candidate runs) with d~
int launch_list_tc(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                   const int32_t* d_lists, int probes, const int32_t* cand_off, int64_t cap, const int64_t* d_list_off, int n_lists,
                   float* out, const float** qn_out, bool one_list_all_queries, int level, float* smin, int64_t cap_s) {
    Context& c = ctx();
    cudaStream_t s = c.stream;
    QueryGroups g{};
    VB_TRY(build_query_groups(d_lists, nq, probes, cand_off, cap, n_lists, LC_N, &g, smin ? cap_s : 0));
    const int64_t max_gtiles = g.n_pairs / LC_N + n_lists + 1;
    void *d_B, *d_qn;
    VB_TRY(workspace(WSC_B, (size_t)max_gtiles * im.n_kblocks * LC_B_STAGE, &d_B));
    VB_TRY(workspace(WSC_N, sizeof(float) * (size_t)nq + 64, &d_qn));
    // the query image is fp32 with the rows' padded dimension count for both element types
    const int qdim = (int)(qstride / 4);
    {
        // |q|^2 of the batch: the probe selection and the list scan of one batch see the same query image -- compute once
        static const void* qn_img = nullptr;
        static int64_t qn_nq = 0;
        static uint64_t qn_epoch = ~0ull;
        static void* qn_buf = nullptr;
        if (!(qn_img == qimg && qn_nq == nq && qn_epoch == c.query_epoch && qn_buf == d_qn)) {
            row_sqnorm_kernel<VB_VECTOR><<<(unsigned)((nq * 32 + 255) / 256), 256, 0, s>>>((const uint8_t*)qimg, qstride, nq, qdim, (float*)d_qn,
                                                                                          nq, 0.f);
            count_launch();
            qn_img = qimg;
            qn_nq = nq;
            qn_epoch = c.query_epoch;
            qn_buf = d_qn;
        }
    }
    const int64_t chunks = g.n_pairs * im.n_kblocks * 8;
    pack_groups_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>((const float*)qimg, qstride, qdim, im.n_kblocks, g.n_pairs,
                                                                        g.pair_q, g.pair_list, g.begin, g.gt_begin, (uint8_t*)d_B);
    VB_CUDA(cudaGetLastError());
    count_launch();
    static bool attr_set = false;
    if (!attr_set) {
        VB_CUDA(cudaFuncSetAttribute(list_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LC_SMEM));
        attr_set = true;
    }
    LcArgs a{};
    a.A = im.planes;
    a.B = (const uint8_t*)d_B;
    a.units = im.units;
    a.n_units = im.n_units;
    a.list_off = d_list_off;
    a.grp_begin = g.begin;
    a.grp_cnt = g.cnt;
    a.gt_begin = g.gt_begin;
    a.pair_q = g.pair_q;
    a.pair_out = g.pair_out;
    a.xn = im.xn;
    a.qn = (const float*)d_qn;
    a.out = out;
    a.smin = smin;
    a.pair_sbase = g.pair_sbase;
    a.n_kblocks = im.n_kblocks;
    a.is_l2 = key_metric == VB_L2_SQUARED;
    a.hi_only = level == 1;
    a.uniform_nqt = one_list_all_queries ? (int)((nq * probes + LC_N - 1) / LC_N) : 0;
    a.n_jobs = a.uniform_nqt ? im.n_units * a.uniform_nqt : im.n_units;
    const int grid = std::max(1, std::min(a.n_jobs, c.sm_count));
    prof_begin(one_list_all_queries ? VB_PROF_CENTRE_TC : VB_PROF_LIST_TC);
    list_tc_kernel<<<grid, LC_THREADS, LC_SMEM, s>>>(a);
    prof_end(one_list_all_queries ? VB_PROF_CENTRE_TC : VB_PROF_LIST_TC);
    VB_CUDA(cudaGetLastError());
    count_launch();
    if (g_traffic_on && g_traffic) {
        lc_traffic_kernel<<<32, 256, 0, s>>>(a, level == 1 ? LC_A_PLANE : LC_A_STAGE, g_traffic + (one_list_all_queries ? 4 : 0));
        VB_CUDA(cudaGetLastError());
    }
    (void)rows;
    *qn_out = (const float*)d_qn;
    return VB_OK;
}

int list_tc_traffic(int on, int64_t* out8) {
    cudaStream_t s = ctx().stream;
    if (!g_traffic) {
        VB_CUDA(cudaMalloc(&g_traffic, 8 * sizeof(unsigned long long)));
        VB_CUDA(cudaMemsetAsync(g_traffic, 0, 8 * sizeof(unsigned long long), s));
    }
    if (out8) {
        VB_CUDA(cudaMemcpyAsync(out8, g_traffic, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaStreamSynchronize(s));
        VB_CUDA(cudaMemsetAsync(g_traffic, 0, 8 * sizeof(unsigned long long), s));
    }
    g_traffic_on = on != 0;
    return VB_OK;
}

// |d~ - d_fp32| <= eps, in units of |x||q| for the product (x2 in the L2 form):
//   representation: bf16 keeps 8 significant bits (unit roundoff 2^-8), so |x - x_hi| <= 2^-8 |x| and
//     |x - x_hi - x_lo| <= 2^-16 |x|.  Level 2 drops x_lo.q_lo and the two residuals: 3 * 2^-16.  Level 1 also
//     drops x_lo.q: 2^-8 + 2^-16.
//   accumulation: one fp32 rounding of the TMEM accumulator per UMMA, (products per K step) * dim / 16 of them,
//     <= 2^-23 each if the unit truncates; doubled to cover the alignment of the 16 products inside an UMMA.
// The constants below are the values the GPU tests and benches of round 1 ran with (dim <= 1536); the formula takes
// over for longer rows, where the accumulation term grows past them.
// The fp32 norms, the final sum and the rounding of the exact fp32 distance it is compared with: 2^-16 (|x|^2 +
// |q|^2) for L2, 2^-17 |x||q| for the inner product.
static LcBound lc_make_bound(const Table& rows, const ListTcImage& im, int key_metric, int level) {
    LcBound bound;
    bound.is_l2 = key_metric == VB_L2_SQUARED;
    const float steps = (float)(im.n_kblocks * (TC_K / 16));
    const float rep = level == 1 ? 1.0f / 256.0f + 1.0f / 65536.0f : 3.0f / 65536.0f;
    const float acc = 2.0f * (level == 1 ? 2.0f : 3.0f) * steps / 8388608.0f;
    const float ip_unit = rep + acc;
    float c_ip = level == 1 ? 1.0f / 256.0f + 1.0f / 8192.0f : 1.0f / 8192.0f;   // validated constants
    c_ip = std::max(c_ip, ip_unit);
    bound.c_dot = bound.is_l2 ? 2.0f * c_ip : c_ip + 1.0f / 131072.0f;
    // norms and the exact fp32 distance each sum dim / 32 terms per lane plus a shuffle tree: 3 * (dim / 32 + 8) * 2^-24 of
    // (|x|^2 + |q|^2) covers the two norms and the distance (<= 2 (|x|^2 + |q|^2)); 2^-16 up to ~2700 dimensions
    bound.c_sum = std::max(1.0f / 65536.0f, 3.0f * ((float)rows.dim / 32.0f + 8.0f) / 16777216.0f);
    bound.xmax = im.xmax;
    return bound;
}

// steps 3 + 4: exact re-score of the selected candidates, final order, certificate.  The number of queries whose
// certificate failed is ADDED to *fail_dev (a device counter the caller zeroes); with n_failed_host the counter is
// also read back (one stream synchronisation), otherwise the caller checks it when it synchronises anyway.
int launch_list_tc_refine(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                          int k, int kp, int probes, const int32_t* d_lists, const int32_t* cand_off, const int64_t* d_list_off,
                          const int32_t* seg_len, const float* qn, const int32_t* pos_kp, const float* approx_kp, int32_t* out_pos,
                          float* out_key, int* fail_dev, int* n_failed_host, int level) {
    Context& c = ctx();
    cudaStream_t s = c.stream;
    void* d_ws;
    VB_TRY(workspace(WSC_K, sizeof(float) * (size_t)nq * kp + (size_t)nq + 64, &d_ws));
    int* n_failed = fail_dev;
    float* exact = (float*)d_ws + 16;
    uint8_t* failed = (uint8_t*)(exact + (size_t)nq * kp);
    const int V = (int)(rows.stride / 16);
    const unsigned grid = (unsigned)((nq * kp * 32 + 255) / 256);
    const LcBound bound = lc_make_bound(rows, im, key_metric, level);
#define VB_RS(E, M) rescore_kernel<E, M><<<grid, 256, 0, s>>>(rows.d, rows.stride, V, (const uint8_t*)qimg, qstride, nq, k, kp, probes, bound, qn, pos_kp, approx_kp, d_lists, cand_off, d_list_off, exact)
    if (rows.elem == VB_VECTOR) {
        if (key_metric == VB_L2_SQUARED) VB_RS(VB_VECTOR, VB_L2_SQUARED);
        else VB_RS(VB_VECTOR, VB_NEG_IP);
    } else {
        if (key_metric == VB_L2_SQUARED) VB_RS(VB_HALFVEC, VB_L2_SQUARED);
        else VB_RS(VB_HALFVEC, VB_NEG_IP);
    }
#undef VB_RS
    certify_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, s>>>(nq, k, kp, bound, qn, seg_len, pos_kp, approx_kp, exact, out_pos, out_key,
                                                                    n_failed, failed);
    VB_CUDA(cudaGetLastError());
    count_launch(2);
    if (n_failed_host) {
        VB_CUDA(cudaMemcpyAsync(n_failed_host, n_failed, sizeof(int), cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaStreamSynchronize(s));
    }
    return VB_OK;
}

// steps 2 + 3 + 4 in one kernel (select_refine_kernel): k' select, exact re-score, final order, certificate
int launch_list_tc_select_refine(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                                 int k, int kp, int probes, const int32_t* d_lists, const int32_t* cand_off, const int64_t* d_list_off,
                                 const float* dist, const int64_t* seg_begin, const int32_t* seg_len, const float* qn, int32_t* out_pos,
                                 float* out_key, int* fail_dev, int* n_failed_host, int level, const int32_t* pre_pos,
                                 const float* pre_key) {
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const LcBound bound = lc_make_bound(rows, im, key_metric, level);
    const int V = (int)(rows.stride / 16);
    const size_t smem = qstride * SR_WARPS;
    const unsigned grid = (unsigned)((nq + SR_WARPS - 1) / SR_WARPS);
    const int R = (kp + 31) / 32;
#define VB_SR2(E, M, RR)                                                                                                              \
    do {                                                                                                                              \
        auto kern = select_refine_kernel<E, M, RR>;                                                                                   \
        if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        kern<<<grid, SR_WARPS * 32, smem, s>>>(rows.d, rows.stride, V, (const uint8_t*)qimg, qstride, nq, k, kp, probes, bound, qn, dist, \
                                              seg_begin, seg_len, d_lists, cand_off, d_list_off, out_pos, out_key, fail_dev, pre_pos, \
                                              pre_key);                                                                              \
    } while (0)
#define VB_SR(E, M)                   \
    do {                              \
        if (R <= 1) VB_SR2(E, M, 1);  \
        else if (R == 2) VB_SR2(E, M, 2); \
        else VB_SR2(E, M, 4);         \
    } while (0)
    VB_REQUIRE(kp <= 128 && smem <= 200 * 1024, "select_refine: k' = %d / query image of %zu bytes not supported", kp, qstride);
    if (rows.elem == VB_VECTOR) {
        if (key_metric == VB_L2_SQUARED) VB_SR(VB_VECTOR, VB_L2_SQUARED);
        else VB_SR(VB_VECTOR, VB_NEG_IP);
    } else {
        if (key_metric == VB_L2_SQUARED) VB_SR(VB_HALFVEC, VB_L2_SQUARED);
        else VB_SR(VB_HALFVEC, VB_NEG_IP);
    }
#undef VB_SR
#undef VB_SR2
    VB_CUDA(cudaGetLastError());
    count_launch();
    if (n_failed_host) {
        VB_CUDA(cudaMemcpyAsync(n_failed_host, fail_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaStreamSynchronize(s));
    }
    return VB_OK;
}

// steps 2 + 3 + 4 with one CTA per query; smin == nullptr: the run is short (<= SS_CAND) and selected directly
int launch_list_tc_cta_refine(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                              int k, int kp, int probes, const int32_t* d_lists, const int32_t* cand_off, const int64_t* d_list_off,
                              const float* dist, const float* smin, int64_t cap, int64_t cap_s, const int32_t* seg_len, const float* qn,
                              int32_t* out_pos, float* out_key, int* fail_dev, int level) {
    if (nq == 0) return VB_OK;
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const LcBound bound = lc_make_bound(rows, im, key_metric, level);
    const int V = (int)(rows.stride / 16);
    const size_t smem = (size_t)SS_CAND * 8 + (size_t)kp * 16 + qstride + (smin ? ss_select_smem_bytes(cap_s, probes) : (size_t)cap * 4) +
                        (size_t)kp * 4 + 16;
    VB_REQUIRE(kp <= 256 && smem <= 200 * 1024 && (smin || cap <= SS_CAND), "cta_refine: k' = %d / %zu bytes of shared memory not supported", kp, smem);
#define VB_CR(E, M)                                                                                                              \
    do {                                                                                                                         \
        auto kern = cta_refine_kernel<E, M>;                                                                                     \
        if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        kern<<<(unsigned)nq, SS_THREADS, smem, s>>>(rows.d, rows.stride, V, (const uint8_t*)qimg, qstride, k, kp, probes, bound, qn, dist, smin, \
                                                   cap, cap_s, seg_len, d_lists, cand_off, d_list_off, out_pos, out_key, fail_dev);   \
    } while (0)
    if (rows.elem == VB_VECTOR) {
        if (key_metric == VB_L2_SQUARED) VB_CR(VB_VECTOR, VB_L2_SQUARED);
        else VB_CR(VB_VECTOR, VB_NEG_IP);
    } else {
        if (key_metric == VB_L2_SQUARED) VB_CR(VB_HALFVEC, VB_L2_SQUARED);
        else VB_CR(VB_HALFVEC, VB_NEG_IP);
    }
#undef VB_CR
    VB_CUDA(cudaGetLastError());
    count_launch();
    return VB_OK;
}

}  // namespace vb
