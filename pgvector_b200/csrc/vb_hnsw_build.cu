// vb_hnsw_build.cu -- HNSW graph construction on the device (CREATE INDEX ... USING hnsw): the in-memory build of
// src/hnswbuild.c:437-480 = per element HnswFindElementNeighbors (src/hnswutils.c:1280-1357: greedy descent,
// HnswSearchLayer with ef_construction on every insertion layer, SelectNeighbors :1065-1165), duplicate folding
// (FindDuplicateInMemory, src/hnswbuild.c:343-364) and HnswUpdateConnection (:1184-1231) on every chosen neighbour.
//
// Formulation.  Elements are inserted in BATCHES, like the reference's parallel build inserts with several workers at
// once (src/hnswbuild.c:412-480: a worker sees the graph as other workers left it, not the elements in flight):
//   K1  hnsw_insert_kernel   one warp per new element: the scan's layer search (vb_hnsw.cuh, `inserting` semantics:
//                            every element counts towards ef) against the graph as of the batch start, then the
//                            neighbour-selection heuristic per layer, written to the element's own neighbour arrays;
//   K1b hnsw_finalize_kernel duplicate check against the chosen layer-0 neighbours, then one (target, layer, source,
//                            distance) record per chosen neighbour;
//       CUB radix sort of the records by (target, layer, source)  -> all updates of one neighbour array are contiguous
//                            and in insertion order;
//   K2  hnsw_update_kernel   one warp per (target, layer): HnswUpdateConnection for each incoming element in turn
//                            (append while there is room, otherwise the heuristic decides which connection goes).
// Batches grow with the graph (at most 1/64 of the elements already inserted, capped) and end at an element that
// becomes the new entry point, so the entry point every search starts from is the reference's.
//
// SelectNeighbors is evaluated EAGERLY: candidates are visited nearest first; an accepted candidate r marks every
// later candidate e with d(e, r) <= d(e, q) as pruned at once (one row image in shared memory, the surviving
// candidates scored against it in one pass).  That is the same predicate as CheckElementCloser (:1040-1060) applied
// in the same order, so the selected set, the order of the pruned candidates kept to fill up, and the connection
// HnswUpdateConnection drops (the farthest pruned candidate, or the farthest one when none is pruned) are the
// reference's.  The reference's `closer` cache (:1090-1122) only skips work; it is not reproduced.
//
// Roofline: HBM / L2 latency-bound row gathers, like the scan (n_dist * row bytes per element, ~3x the scan's because
// of the heuristic and the neighbour updates).  Parity: the graph depends on PRNG level draws and on insertion
// concurrency, which no reference test pins; parity is by the reference's recall floors (test/t/012, 020) on
// GPU-built graphs and by exact search equality of GPU and oracle on the SAME exported graph.
#include "vb_hnsw.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

namespace vb {

struct BuildDev {
    HnswDev g;
    int32_t* nbr0_w;      // same arrays as g.nbr0 / g.upper, writable
    int32_t* upper_w;
    float* nd0;           // [n][2m] distance stored with each neighbour
    float* upper_d;       // [slots][m]
    int32_t* dup_of;      // [n] element this row was folded into, or -1
    int32_t* n_heaptids;  // [n] heap tids carried by the element (HNSW_HEAPTIDS = 10 at most, src/hnsw.h:69)
    int efc;
    int b0, B;            // this batch inserts elements [b0, b0 + B)
    int qvec;             // 16-byte vectors of a row image
    uint64_t* edge_key;   // target << 26 | layer << 20 | (source - b0)
    float* edge_val;
    int* n_edges;
    int* overflow;
};

constexpr int HB_CAND = 256;     // candidates of one HnswUpdateConnection: lm + 1 <= 201, padded to a power of two

__device__ __forceinline__ float key64_to_float(uint64_t k) { return (float)key64_to_double(k); }

// shared memory of one warp of hnsw_insert_kernel (bytes, 16-aligned)
__host__ __device__ inline size_t hb_insert_smem(int qvec, int efc, int lm0) {
    size_t b = (size_t)qvec * 16 * 2;          // image of the new element, image of an accepted candidate
    b += (size_t)efc * 2 * 8 + 32 * 8;         // keys A, keys B, batch keys
    b += (size_t)efc * 2 * 4 + 32 * 4 + 32 * 4; // ids A, ids B, batch ids, batch -> candidate index
    b += (size_t)lm0 * 4;                      // selected candidate indices
    b += (size_t)efc * 2;                      // pruned candidate indices (uint16)
    b += (size_t)efc;                          // pruned flags
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t hb_update_smem(int qvec) {
    size_t b = (size_t)qvec * 16;              // image of an accepted candidate
    b += (size_t)HB_CAND * 8 + 32 * 8;         // candidate keys, batch keys
    b += 32 * 4 + 32 * 4;                      // batch ids, batch -> candidate index
    b += HB_CAND;                              // pruned flags
    return (b + 15) & ~(size_t)15;
}

// candidates cand[from..n) that are still unpruned are scored against the image `img` (the row of the candidate
// just accepted); those with d(e, r) <= d(e, q) are pruned (CheckElementCloser, src/hnswutils.c:1040-1060)
template <int ELEM, int METRIC, int LPR, typename IdOf, typename DistOf>
__device__ __forceinline__ void prune_against(const HnswDev& g, const uint4* img, int from, int n, uint8_t* dead, uint32_t* bid,
                                              int32_t* bj, uint64_t* bkey, int lane, IdOf id_of, DistOf dist_of) {
    for (int base = from; base < n; base += 32) {
        const int j = base + lane;
        const bool alive = j < n && !dead[j];
        const unsigned am = __ballot_sync(0xffffffffu, alive);
        const int cnt = __popc(am);
        if (cnt == 0) continue;
        const int pos = __popc(am & ((1u << lane) - 1u));
        if (alive) {
            bid[pos] = id_of(j);
            bj[pos] = j;
        }
        __syncwarp();
        hnsw_score_batch<ELEM, METRIC, LPR>(g, img, bid, cnt, bkey, lane);
        __syncwarp();
        if (lane < cnt) {
            const int jj = bj[lane];
            if (key64_to_float(bkey[lane]) <= dist_of(jj)) dead[jj] = 1;
        }
        __syncwarp();
    }
}

// K1: one warp = one new element
template <int ELEM, int METRIC, int LPR>
__global__ void __launch_bounds__(HN_WARPS * 32) hnsw_insert_kernel(BuildDev b, uint32_t* __restrict__ vis_all, uint32_t vis_cap,
                                                                    uint32_t vis_upper) {
    extern __shared__ uint4 smem[];
    const HnswDev& g = b.g;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int efc = b.efc, lm0 = 2 * g.m;
    uint8_t* base = reinterpret_cast<uint8_t*>(smem) + (size_t)warp * hb_insert_smem(b.qvec, efc, lm0);
    uint4* sq = reinterpret_cast<uint4*>(base);
    uint4* img = sq + b.qvec;
    uint64_t* keyA = reinterpret_cast<uint64_t*>(img + b.qvec);
    uint64_t* keyB = keyA + efc;
    uint64_t* bkey = keyB + efc;
    uint32_t* idA = reinterpret_cast<uint32_t*>(bkey + 32);
    uint32_t* idB = idA + efc;
    uint32_t* bid = idB + efc;
    int32_t* bj = reinterpret_cast<int32_t*>(bid + 32);
    int32_t* sel = bj + 32;
    uint16_t* wd = reinterpret_cast<uint16_t*>(sel + lm0);
    uint8_t* dead = reinterpret_cast<uint8_t*>(wd + efc);

    const int gwarp = blockIdx.x * HN_WARPS + warp;
    const int nwarps = gridDim.x * HN_WARPS;
    uint32_t* vis = vis_all + (size_t)gwarp * vis_cap;

    for (int w = gwarp; w < b.B; w += nwarps) {
        const int e = b.b0 + w;
        load_row_image<ELEM, METRIC>(g.rows + (size_t)e * g.stride, g.V, sq, lane);
        __syncwarp();
        const int level = g.levels[e];

        HnswWarpState S;
        S.rk = keyA;
        S.ri = idA;
        S.nk = keyB;
        S.ni = idB;
        S.vcn = 0;
        S.bkey = bkey;
        S.bid = bid;
        // entry point (HnswEntryCandidate, src/hnswutils.c:609-621)
        {
            Acc<ELEM, METRIC> acc;
            const uint4* rp = reinterpret_cast<const uint4*>(g.rows + (size_t)g.entry * g.stride);
            for (int v = lane; v < g.V; v += 32) hnsw_acc_add<ELEM, METRIC>(acc, ldg_stream(rp + v), sq, v);
            acc.template reduce<32>();
            if (lane == 0) {
                S.rk[0] = orderable_key64(acc.value());
                S.ri[0] = (uint32_t)g.entry;
            }
            S.len = 1;
            __syncwarp();
        }
        bool failed = false;
        // 1st phase: greedy search to the insert level (src/hnswutils.c:1308-1313)
        int lc = g.entry_level;
        for (; lc > level && !failed; --lc) failed = !hnsw_search_layer<ELEM, METRIC, LPR>(g, sq, lc, 1, lane, S, vis, vis_upper, nullptr);
        // 2nd phase (:1322-1354): level = min(level, entryLevel)
        for (; lc >= 0 && !failed; --lc) {
            failed = !hnsw_search_layer<ELEM, METRIC, LPR>(g, sq, lc, efc, lane, S, vis + vis_upper, vis_cap - vis_upper, nullptr);
            if (failed) break;
            const int lm = lc == 0 ? lm0 : g.m;
            int32_t* out_ids = lc == 0 ? b.nbr0_w + (size_t)e * lm : b.upper_w + ((size_t)g.upper_off[e] + (lc - 1)) * (size_t)lm;
            float* out_d = lc == 0 ? b.nd0 + (size_t)e * lm : b.upper_d + ((size_t)g.upper_off[e] + (lc - 1)) * (size_t)lm;
            const int len = S.len;
            const uint64_t* wk = S.rk;     // W, nearest first; stays intact: it is the next layer's entry list (ep = w)
            const uint32_t* wi = S.ri;
            if (len <= lm) {
                // SelectNeighbors returns the list as it is (:1077-1078): W drained from the max-heap = farthest first
                for (int i = lane; i < lm; i += 32) {
                    const int j = len - 1 - i;
                    out_ids[i] = i < len ? (int32_t)(wi[j] & 0x7fffffffu) : -1;
                    out_d[i] = i < len ? key64_to_float(wk[j]) : 0.f;
                }
            } else {
                for (int i = lane; i < len; i += 32) dead[i] = 0;
                __syncwarp();
                int nR = 0, nWd = 0;
                for (int i = 0; i < len; ++i) {
                    if (dead[i]) {   // shared memory flag: the same value for every lane
                        if (lane == 0) wd[nWd] = (uint16_t)i;
                        ++nWd;
                        continue;
                    }
                    if (lane == 0) sel[nR] = i;
                    ++nR;
                    if (nR == lm || i + 1 >= len) break;
                    load_row_image<ELEM, METRIC>(g.rows + (size_t)(wi[i] & 0x7fffffffu) * g.stride, g.V, img, lane);
                    __syncwarp();
                    prune_against<ELEM, METRIC, LPR>(
                        g, img, i + 1, len, dead, bid, bj, bkey, lane, [&](int j) { return wi[j] & 0x7fffffffu; },
                        [&](int j) { return key64_to_float(wk[j]); });
                }
                __syncwarp();
                // keep pruned connections (:1151-1153)
                for (int t = 0; t < nWd && nR < lm; ++t, ++nR)
                    if (lane == 0) sel[nR] = wd[t];
                __syncwarp();
                for (int i = lane; i < lm; i += 32) {
                    const int j = i < nR ? sel[i] : 0;
                    out_ids[i] = i < nR ? (int32_t)(wi[j] & 0x7fffffffu) : -1;
                    out_d[i] = i < nR ? key64_to_float(wk[j]) : 0.f;
                }
            }
            __syncwarp();
        }
        if (failed && lane == 0) atomicExch(b.overflow, 1);
        __syncwarp();
    }
}

// K1b: duplicates (FindDuplicateInMemory, src/hnswbuild.c:343-364) and the update records of every chosen neighbour
__global__ void __launch_bounds__(128) hnsw_finalize_kernel(BuildDev b) {
    if (*b.overflow) return;   // K1 is repeated with larger visited tables: no side effect may have happened yet
    const HnswDev& g = b.g;
    const int lane = threadIdx.x % 32;
    const int gwarp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32);
    const int nwarps = (int)((gridDim.x * (int64_t)blockDim.x) / 32);
    const int lm0 = 2 * g.m;
    for (int w = gwarp; w < b.B; w += nwarps) {
        const int e = b.b0 + w;
        const int level = g.levels[e];
        const int top = min(level, g.entry_level);
        int32_t* ids0 = b.nbr0_w + (size_t)e * lm0;
        const uint4* re = reinterpret_cast<const uint4*>(g.rows + (size_t)e * g.stride);
        bool dup = false;
        for (int i = 0; i < lm0; ++i) {
            const int t = ids0[i];
            if (t < 0) break;
            const uint4* rt = reinterpret_cast<const uint4*>(g.rows + (size_t)t * g.stride);
            bool eq = true;
            for (int v = lane; v < g.V; v += 32) {
                const uint4 x = __ldg(re + v), y = __ldg(rt + v);
                eq = eq && x.x == y.x && x.y == y.y && x.z == y.z && x.w == y.w;
            }
            if (!__all_sync(0xffffffffu, eq)) break;   // "exit early since ordered by distance"
            int old = 0;
            if (lane == 0) old = atomicAdd(&b.n_heaptids[t], 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old < 10) {
                dup = true;
                if (lane == 0) b.dup_of[e] = t;
                break;
            }
            if (lane == 0) atomicSub(&b.n_heaptids[t], 1);
        }
        if (dup) {
            // the row rides on element t: it never becomes an element (no connections in either direction)
            for (int i = lane; i < lm0; i += 32) ids0[i] = -1;
            for (int lc = 1; lc <= top; ++lc) {
                int32_t* ids = b.upper_w + ((size_t)g.upper_off[e] + (lc - 1)) * (size_t)g.m;
                for (int i = lane; i < g.m; i += 32) ids[i] = -1;
            }
            if (lane == 0) b.n_heaptids[e] = 0;
            continue;
        }
        // UpdateNeighborsInMemory (src/hnswbuild.c:381-410): one record per (neighbour, layer)
        for (int lc = top; lc >= 0; --lc) {
            const int lm = lc == 0 ? lm0 : g.m;
            const int32_t* ids = lc == 0 ? ids0 : b.upper_w + ((size_t)g.upper_off[e] + (lc - 1)) * (size_t)lm;
            const float* ds = lc == 0 ? b.nd0 + (size_t)e * lm : b.upper_d + ((size_t)g.upper_off[e] + (lc - 1)) * (size_t)lm;
            for (int off = 0; off < lm; off += 32) {
                const int t = off + lane < lm ? ids[off + lane] : -1;
                const unsigned vm = __ballot_sync(0xffffffffu, t >= 0);
                const int cnt = __popc(vm);
                if (cnt == 0) break;
                int slot = 0;
                if (lane == 0) slot = atomicAdd(b.n_edges, cnt);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (t >= 0) {
                    const int p = slot + __popc(vm & ((1u << lane) - 1u));
                    b.edge_key[p] = ((uint64_t)(uint32_t)t << 26) | ((uint64_t)lc << 20) | (uint64_t)w;
                    b.edge_val[p] = ds[off + lane];
                }
            }
        }
    }
}

// K2: one warp per (target, layer) run of the sorted records = HnswUpdateConnection (src/hnswutils.c:1184-1231) for each
// incoming element in insertion order
template <int ELEM, int METRIC, int LPR>
__global__ void __launch_bounds__(HN_WARPS * 32) hnsw_update_kernel(BuildDev b, const uint64_t* __restrict__ keys,
                                                                    const float* __restrict__ vals, int n_edges) {
    extern __shared__ uint4 smem[];
    const HnswDev& g = b.g;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* base = reinterpret_cast<uint8_t*>(smem) + (size_t)warp * hb_update_smem(b.qvec);
    uint4* img = reinterpret_cast<uint4*>(base);
    uint64_t* ck = reinterpret_cast<uint64_t*>(img + b.qvec);
    uint64_t* bkey = ck + HB_CAND;
    uint32_t* bid = reinterpret_cast<uint32_t*>(bkey + 32);
    int32_t* bj = reinterpret_cast<int32_t*>(bid + 32);
    uint8_t* dead = reinterpret_cast<uint8_t*>(bj + 32);

    const int64_t gwarp = blockIdx.x * (int64_t)HN_WARPS + warp;
    const int64_t nwarps = gridDim.x * (int64_t)HN_WARPS;
    for (int64_t i = gwarp; i < n_edges; i += nwarps) {
        const uint64_t head = keys[i] >> 20;
        if (i > 0 && (keys[i - 1] >> 20) == head) continue;   // not the first record of its (target, layer) run
        const int t = (int)(head >> 6), lc = (int)(head & 63);
        const int lm = lc == 0 ? 2 * g.m : g.m;
        int32_t* ids = lc == 0 ? b.nbr0_w + (size_t)t * lm : b.upper_w + ((size_t)g.upper_off[t] + (lc - 1)) * (size_t)lm;
        float* ds = lc == 0 ? b.nd0 + (size_t)t * lm : b.upper_d + ((size_t)g.upper_off[t] + (lc - 1)) * (size_t)lm;
        for (int64_t p = i; p < n_edges && (keys[p] >> 20) == head; ++p) {
            const int src = b.b0 + (int)(keys[p] & 0xFFFFFu);
            const float d = vals[p];
            // current length = first invalid entry
            int count = lm;
            for (int off = 0; off < lm; off += 32) {
                const int nid = off + lane < lm ? ids[off + lane] : -1;
                const unsigned inval = ~__ballot_sync(0xffffffffu, nid >= 0);
                if (inval) {
                    count = min(lm, off + __ffs(inval) - 1);
                    break;
                }
            }
            if (count < lm) {   // room left: append (:1192-1199)
                if (lane == 0) {
                    ids[count] = src;
                    ds[count] = d;
                }
                __syncwarp();
                continue;
            }
            // shrink connections (:1201-1230): candidates = the lm connections + the new element, nearest first
            const int n = lm + 1;
            int P = 2;
            while (P < n) P <<= 1;
            for (int j = lane; j < P; j += 32) {
                uint64_t key = ~0ull;
                if (j < lm) key = ((uint64_t)orderable_key(ds[j]) << 32) | (uint32_t)ids[j];
                else if (j == lm) key = ((uint64_t)orderable_key(d) << 32) | (uint32_t)src;
                ck[j] = key;
                dead[j] = 0;
            }
            __syncwarp();
            for (int size = 2; size <= P; size <<= 1)
                for (int st = size >> 1; st > 0; st >>= 1) {
                    for (int a = lane; a < P; a += 32) {
                        const int c = a ^ st;
                        if (c > a) {
                            const uint64_t x = ck[a], y = ck[c];
                            const bool up = (a & size) == 0;
                            if ((x > y) == up) {
                                ck[a] = y;
                                ck[c] = x;
                            }
                        }
                    }
                    __syncwarp();
                }
            int nAcc = 0;
            for (int c = 0; c < n; ++c) {
                if (dead[c]) continue;
                ++nAcc;
                if (nAcc == lm || c == n - 1) break;
                load_row_image<ELEM, METRIC>(g.rows + (size_t)(uint32_t)ck[c] * g.stride, g.V, img, lane);
                __syncwarp();
                prune_against<ELEM, METRIC, LPR>(
                    g, img, c + 1, n, dead, bid, bj, bkey, lane, [&](int j) { return (uint32_t)ck[j]; },
                    [&](int j) { return key_to_float((uint32_t)(ck[j] >> 32)); });
                if (dead[n - 1]) break;   // the farthest candidate is pruned: it is the one that goes
            }
            __syncwarp();
            // the connection that goes: the farthest pruned candidate, or the farthest candidate when none is pruned
            int drop = -1;
            for (int j0 = ((n - 1) / 32) * 32; j0 >= 0 && drop < 0; j0 -= 32) {
                const int j = j0 + lane;
                const unsigned dm = __ballot_sync(0xffffffffu, j < n && dead[j]);
                if (dm) drop = j0 + 31 - __clz(dm);
            }
            if (drop < 0) drop = n - 1;
            const uint32_t drop_id = (uint32_t)ck[drop];
            if (drop_id != (uint32_t)src) {
                for (int j = lane; j < lm; j += 32)
                    if ((uint32_t)ids[j] == drop_id) {
                        ids[j] = src;
                        ds[j] = d;
                    }
            }
            __syncwarp();
        }
    }
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ----------------------------------------------------------------------------- host side

enum { WSB_KEYS = 14, WSB_KEYS2 = 15, WSB_VALS = 16, WSB_VALS2 = 17, WSB_TMP = 18, WSB_FLAGS = 19 };

struct BuildLaunch {
    int (*insert)(const BuildDev&, uint32_t*, uint32_t, uint32_t, int, size_t, int*);
    int (*update)(const BuildDev&, const uint64_t*, const float*, int, int, size_t, int*);
};

template <int ELEM, int METRIC, int LPR>
static int launch_insert(const BuildDev& b, uint32_t* vis, uint32_t vis_cap, uint32_t vis_upper, int grid, size_t smem, int* occ) {
    auto kern = hnsw_insert_kernel<ELEM, METRIC, LPR>;
    if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (occ) {
        VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, HN_WARPS * 32, smem));
        return VB_OK;
    }
    kern<<<grid, HN_WARPS * 32, smem, ctx().stream>>>(b, vis, vis_cap, vis_upper);
    VB_CUDA(cudaGetLastError());
    count_launch();
    return VB_OK;
}
template <int ELEM, int METRIC, int LPR>
static int launch_update(const BuildDev& b, const uint64_t* keys, const float* vals, int n_edges, int grid, size_t smem, int* occ) {
    auto kern = hnsw_update_kernel<ELEM, METRIC, LPR>;
    if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (occ) {
        VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, HN_WARPS * 32, smem));
        return VB_OK;
    }
    kern<<<grid, HN_WARPS * 32, smem, ctx().stream>>>(b, keys, vals, n_edges);
    VB_CUDA(cudaGetLastError());
    count_launch();
    return VB_OK;
}

template <int ELEM, int METRIC>
static BuildLaunch pick_lpr(int V) {
    // lanes per row: a whole warp for rows of >= 512 bytes, 8 lanes for >= 128 bytes, one lane for tiny rows
    if (V >= 32) return BuildLaunch{launch_insert<ELEM, METRIC, 32>, launch_update<ELEM, METRIC, 32>};
    if (V >= 8) return BuildLaunch{launch_insert<ELEM, METRIC, 8>, launch_update<ELEM, METRIC, 8>};
    return BuildLaunch{launch_insert<ELEM, METRIC, 1>, launch_update<ELEM, METRIC, 1>};
}

static bool pick_kernels(const Hnsw& h, int V, BuildLaunch* out) {
    if (h.elem == VB_VECTOR) {
        if (h.metric == VB_L2_SQUARED) *out = pick_lpr<VB_VECTOR, VB_L2_SQUARED>(V);
        else if (h.metric == VB_NEG_IP) *out = pick_lpr<VB_VECTOR, VB_NEG_IP>(V);
        else if (h.metric == VB_L1) *out = pick_lpr<VB_VECTOR, VB_L1>(V);
        else return false;
    } else if (h.elem == VB_HALFVEC) {
        if (h.metric == VB_L2_SQUARED) *out = pick_lpr<VB_HALFVEC, VB_L2_SQUARED>(V);
        else if (h.metric == VB_NEG_IP) *out = pick_lpr<VB_HALFVEC, VB_NEG_IP>(V);
        else if (h.metric == VB_L1) *out = pick_lpr<VB_HALFVEC, VB_L1>(V);
        else return false;
    } else {
        if (h.metric == VB_HAMMING) *out = pick_lpr<VB_BIT, VB_HAMMING>(V);
        else if (h.metric == VB_JACCARD) *out = pick_lpr<VB_BIT, VB_JACCARD>(V);
        else return false;
    }
    return true;
}

static double build_uniform(uint64_t* st) {   // (0, 1]: -log() stays finite
    uint64_t z = (*st += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return (double)((z >> 11) + 1) * (1.0 / 9007199254740992.0);
}

static int hnsw_build_impl(Hnsw& h, const void* rows, bool rows_on_host, int64_t n, int efc, uint64_t seed, const int32_t* levels_in) {
    VB_REQUIRE(efc >= 4 && efc <= 1000, "ef_construction must be 4..1000 (src/hnsw.h:57-59)");
    VB_REQUIRE(efc >= 2 * h.m, "ef_construction must be greater than or equal to 2 * m (src/hnswbuild.c:713-716)");
    VB_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "bad row count");
    Context& c = ctx();
    cudaStream_t s = c.stream;
    hnsw_release(h);
    h.n = n;
    h.entry = -1;
    h.entry_level = -1;
    if (n == 0) {
        h.loaded = true;
        return VB_OK;
    }
    VB_REQUIRE(rows, "null rows");
    if (rows_on_host) VB_TRY(table_append_host(h.rows, rows, n));
    else VB_TRY(table_append_dev(h.rows, rows, n));

    // levels (HnswInitElement, src/hnswutils.c:248-254): (int) (-log(RandomDouble()) * ml), capped at HnswGetMaxLevel(m)
    const int m = h.m, lm0 = 2 * m;
    const double ml = 1.0 / std::log((double)m);
    const int max_level = std::min((int)((8192 - 24 - 8 - 4 - 4) / 6 / m) - 2, 63);   // src/hnsw.h:133 with BLCKSZ = 8192
    std::vector<int32_t> levels((size_t)n), uoff((size_t)n);
    uint64_t rs = seed ^ 0x2545f4914f6cdd1dULL;
    int64_t slots = 0;
    for (int64_t i = 0; i < n; ++i) {
        int lv = levels_in ? levels_in[i] : (int)(-std::log(build_uniform(&rs)) * ml);
        VB_REQUIRE(lv >= 0, "negative level");
        lv = std::min(lv, max_level);
        levels[(size_t)i] = lv;
        uoff[(size_t)i] = lv > 0 ? (int32_t)slots : -1;
        slots += lv;
        VB_REQUIRE(slots < (int64_t)0x7fffffff, "upper slot overflow");
    }
    h.upper_slots = slots;
    const size_t up_elems = (size_t)std::max<int64_t>(slots, 1) * m;
    VB_CUDA(cudaMalloc(&h.levels, sizeof(int32_t) * (size_t)n));
    VB_CUDA(cudaMalloc(&h.upper_off, sizeof(int32_t) * (size_t)n));
    VB_CUDA(cudaMalloc(&h.nbr0, sizeof(int32_t) * (size_t)n * lm0));
    VB_CUDA(cudaMalloc(&h.nd0, sizeof(float) * (size_t)n * lm0));
    VB_CUDA(cudaMalloc(&h.upper, sizeof(int32_t) * up_elems));
    VB_CUDA(cudaMalloc(&h.upper_d, sizeof(float) * up_elems));
    VB_CUDA(cudaMalloc(&h.dup_of, sizeof(int32_t) * (size_t)n));
    VB_CUDA(cudaMalloc(&h.n_heaptids, sizeof(int32_t) * (size_t)n));
    VB_CUDA(cudaMemcpyAsync(h.levels, levels.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    VB_CUDA(cudaMemcpyAsync(h.upper_off, uoff.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    VB_CUDA(cudaMemsetAsync(h.nbr0, 0xFF, sizeof(int32_t) * (size_t)n * lm0, s));
    VB_CUDA(cudaMemsetAsync(h.upper, 0xFF, sizeof(int32_t) * up_elems, s));
    VB_CUDA(cudaMemsetAsync(h.dup_of, 0xFF, sizeof(int32_t) * (size_t)n, s));
    fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h.n_heaptids, n, 1);
    VB_CUDA(cudaGetLastError());
    count_launch();

    BuildLaunch K;
    const int V = (int)(h.rows.stride / 16);
    if (!pick_kernels(h, V, &K)) {
        set_error("hnsw build: unsupported metric %d for element type %d", h.metric, h.elem);
        return VB_EINVAL;
    }
    const int qvec = h.elem == VB_HALFVEC ? 2 * V : V;
    const size_t smem_ins = hb_insert_smem(qvec, efc, lm0) * HN_WARPS;
    const size_t smem_upd = hb_update_smem(qvec) * HN_WARPS;
    VB_REQUIRE(smem_ins <= 200 * 1024 && smem_upd <= 200 * 1024,
               "ef_construction %d / m %d with this dimension need %zu bytes of shared memory per CTA", efc, m, smem_ins);

    BuildDev b{};
    b.g.rows = h.rows.d;
    b.g.stride = h.rows.stride;
    b.g.V = V;
    b.g.levels = h.levels;
    b.g.nbr0 = h.nbr0;
    b.g.upper_off = h.upper_off;
    b.g.upper = h.upper;
    b.g.m = m;
    b.g.n = n;
    b.nbr0_w = h.nbr0;
    b.upper_w = h.upper;
    b.nd0 = h.nd0;
    b.upper_d = h.upper_d;
    b.dup_of = h.dup_of;
    b.n_heaptids = h.n_heaptids;
    b.efc = efc;
    b.qvec = qvec;

    int occ_ins = 1, occ_upd = 1;
    VB_TRY(K.insert(b, nullptr, 0, 0, 0, smem_ins, &occ_ins));
    VB_TRY(K.update(b, nullptr, nullptr, 0, 0, smem_upd, &occ_upd));
    const int max_grid_ins = c.sm_count * std::max(1, occ_ins);
    const int max_grid_upd = c.sm_count * std::max(1, occ_upd) * 4;

    void* d_flags;
    VB_TRY(workspace(WSB_FLAGS, 64, &d_flags));
    b.n_edges = (int*)d_flags;
    b.overflow = b.n_edges + 1;

    // visited tables: one per resident warp; the insertion layers share the large region, the greedy layers the small one
    uint32_t cap = 1u << 14;
    while (cap < (uint32_t)(efc * m * 16) && cap < (1u << 22)) cap <<= 1;

    // Elements of one batch do not see each other (like concurrent workers), so a batch stays a small fraction of the
    // graph it is inserted into: 1/64 by default (measured on B200, 20 000 x 48-d mixture, ef_search 80: recall@10 0.65
    // at 1/8 vs 0.73 for the serial build).  Options "hnsw_build_fraction" / "hnsw_build_batch".
    const int64_t frac = std::max<int64_t>(1, c.hnsw_build_fraction);
    const int64_t b_max = std::min<int64_t>(1 << 20, std::max<int64_t>(1, c.hnsw_build_batch));
    int64_t done = 1;   // element 0 is the first entry point: no neighbours (src/hnswutils.c:1300-1302)
    h.entry = 0;
    h.entry_level = levels[0];
    while (done < n) {
        int64_t B = std::min<int64_t>(std::min<int64_t>(b_max, std::max<int64_t>(1, done / frac)), n - done);
        // a batch ends at the first element that rises above the entry point: it becomes the entry point of the next batch
        int64_t promote = -1;
        int64_t max_edges = 0;
        for (int64_t i = 0; i < B; ++i) {
            const int lv = levels[(size_t)(done + i)];
            max_edges += lm0 + (int64_t)std::min(lv, h.entry_level) * m;
            if (lv > h.entry_level) {
                promote = done + i;
                B = i + 1;
                break;
            }
        }
        VB_REQUIRE(max_edges < (int64_t)0x7fffffff, "too many connection updates in one batch");
        b.g.entry = (int)h.entry;
        b.g.entry_level = h.entry_level;
        b.b0 = (int)done;
        b.B = (int)B;
        void *d_k1, *d_k2, *d_v1, *d_v2;
        VB_TRY(workspace(WSB_KEYS, sizeof(uint64_t) * (size_t)max_edges, &d_k1));
        VB_TRY(workspace(WSB_KEYS2, sizeof(uint64_t) * (size_t)max_edges, &d_k2));
        VB_TRY(workspace(WSB_VALS, sizeof(float) * (size_t)max_edges, &d_v1));
        VB_TRY(workspace(WSB_VALS2, sizeof(float) * (size_t)max_edges, &d_v2));
        b.edge_key = (uint64_t*)d_k1;
        b.edge_val = (float*)d_v1;
        const int grid_ins = (int)std::min<int64_t>((B + HN_WARPS - 1) / HN_WARPS, max_grid_ins);
        int flags[2] = {0, 0};
        for (int attempt = 0;; ++attempt) {
            const uint32_t vis_upper = std::max<uint32_t>(2048u, cap / 8);
            const uint32_t vis_cap = cap + vis_upper;
            const size_t need = (size_t)max_grid_ins * HN_WARPS * vis_cap * sizeof(uint32_t);
            if (h.vis_bytes < need) {
                if (h.vis) {
                    VB_CUDA(cudaStreamSynchronize(s));
                    cudaFree(h.vis);
                    h.vis = nullptr;
                    h.vis_bytes = 0;
                }
                if (cudaMalloc(&h.vis, need) != cudaSuccess) {
                    set_error("hnsw build: visited tables (%zu bytes) do not fit", need);
                    return VB_ENOMEM;
                }
                h.vis_bytes = need;
            }
            VB_CUDA(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), s));
            VB_TRY(K.insert(b, h.vis, vis_cap, vis_upper, grid_ins, smem_ins, nullptr));
            hnsw_finalize_kernel<<<(unsigned)std::min<int64_t>((B * 32 + 127) / 128, (int64_t)c.sm_count * 16), 128, 0, s>>>(b);
            VB_CUDA(cudaGetLastError());
            count_launch();
            VB_CUDA(cudaMemcpyAsync(flags, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            VB_CUDA(cudaStreamSynchronize(s));
            if (!flags[1]) break;
            cap <<= 2;   // a visited table overflowed: repeat the batch's searches with larger ones (nothing was published)
            VB_REQUIRE(attempt < 5 && cap <= (1u << 26), "hnsw build: visited set overflow");
        }
        const int n_edges = flags[0];
        VB_REQUIRE(n_edges <= max_edges, "hnsw build: record overflow (%d > %lld)", n_edges, (long long)max_edges);
        if (n_edges > 0) {
            size_t tmp_bytes = 0;
            VB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)d_k1, (uint64_t*)d_k2, (const float*)d_v1,
                                                    (float*)d_v2, n_edges, 0, 57, s));
            void* d_tmp;
            VB_TRY(workspace(WSB_TMP, tmp_bytes, &d_tmp));
            VB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, (const uint64_t*)d_k1, (uint64_t*)d_k2, (const float*)d_v1,
                                                    (float*)d_v2, n_edges, 0, 57, s));
            count_launch();
            const int grid_upd = (int)std::min<int64_t>(((int64_t)n_edges + HN_WARPS - 1) / HN_WARPS, max_grid_upd);
            VB_TRY(K.update(b, (const uint64_t*)d_k2, (const float*)d_v2, n_edges, grid_upd, smem_upd, nullptr));
        }
        if (promote >= 0) {
            // UpdateGraphInMemory (src/hnswbuild.c:428-430): a duplicate never becomes the entry point
            int32_t dup = -1;
            VB_CUDA(cudaMemcpyAsync(&dup, h.dup_of + promote, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            VB_CUDA(cudaStreamSynchronize(s));
            if (dup < 0) {
                h.entry = promote;
                h.entry_level = levels[(size_t)promote];
            }
        }
        done += B;
    }
    VB_CUDA(cudaStreamSynchronize(s));
    h.loaded = true;
    return VB_OK;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_hnsw_build(vb_hnsw* p, const void* rows, int64_t n, int ef_construction, uint64_t seed, const int32_t* levels) {
    VB_TRY(require_init());
    VB_REQUIRE(p, "null index");
    return hnsw_build_impl(p->h, rows, true, n, ef_construction, seed, levels);
}

int vb_hnsw_build_dev(vb_hnsw* p, const void* rows_dev, int64_t n, int ef_construction, uint64_t seed, const int32_t* levels) {
    VB_TRY(require_init());
    VB_REQUIRE(p, "null index");
    return hnsw_build_impl(p->h, rows_dev, false, n, ef_construction, seed, levels);
}

int64_t vb_hnsw_rows(const vb_hnsw* p) { return p ? p->h.n : 0; }
int64_t vb_hnsw_upper_slots(const vb_hnsw* p) { return p ? p->h.upper_slots : 0; }

int vb_hnsw_export(vb_hnsw* p, int32_t* levels, int32_t* nbr0, int64_t* upper_off, int32_t* upper, int64_t* entry, int32_t* dup_of) {
    VB_TRY(require_init());
    VB_REQUIRE(p && p->h.loaded, "hnsw index not loaded");
    Hnsw& h = p->h;
    const int64_t n = h.n;
    cudaStream_t s = ctx().stream;
    if (entry) *entry = h.entry;
    if (n == 0) return VB_OK;
    if (levels) VB_CUDA(cudaMemcpyAsync(levels, h.levels, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (nbr0) VB_CUDA(cudaMemcpyAsync(nbr0, h.nbr0, sizeof(int32_t) * (size_t)n * 2 * h.m, cudaMemcpyDeviceToHost, s));
    if (upper && h.upper_slots > 0)
        VB_CUDA(cudaMemcpyAsync(upper, h.upper, sizeof(int32_t) * (size_t)h.upper_slots * h.m, cudaMemcpyDeviceToHost, s));
    std::vector<int32_t> uo;
    if (upper_off) {
        uo.resize((size_t)n);
        VB_CUDA(cudaMemcpyAsync(uo.data(), h.upper_off, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
    }
    if (dup_of) {
        if (h.dup_of) VB_CUDA(cudaMemcpyAsync(dup_of, h.dup_of, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        else
            for (int64_t i = 0; i < n; ++i) dup_of[i] = -1;
    }
    VB_CUDA(cudaStreamSynchronize(s));
    if (upper_off)
        for (int64_t i = 0; i < n; ++i) upper_off[i] = uo[(size_t)i];
    return VB_OK;
}

}  // extern "C"
