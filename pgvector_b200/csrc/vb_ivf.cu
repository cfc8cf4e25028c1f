// vb_ivf.cu -- C ABI for the batched distance operator, resident tables, exact
// top-k and the IVFFlat scan path (GetScanLists + GetScanItems, src/ivfscan.c:47-187).
#include "vb_common.cuh"
#include "vb_distance.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

namespace vb {

// workspace slots
enum { WS_QIMG = 0, WS_DIST = 1, WS_CDIST = 2, WS_PROBES = 3, WS_CHUNKS = 4, WS_SEG = 5, WS_POS = 6, WS_OUT = 7 };
// 8..11 are used by the CUB sort path in vb_scan.cu
enum { WS_MISC = 12, WS_OUT2 = 13, WS_SMIN = 31 };

__global__ void regular_segments_kernel(int64_t nseg, int64_t stride, int32_t len, int64_t* begin, int32_t* lens) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nseg) {
        begin[i] = i * stride;
        lens[i] = len;
    }
}

// float key -> the operator's float8 (sqrt for <->, negate for inner_product)
__device__ __forceinline__ double finish_value(int metric, float key) {
    if (metric == VB_L2) return sqrt((double)key);
    if (metric == VB_IP) return -(double)key;
    return (double)key;
}

__global__ void finish_exact_kernel(int metric, int64_t n, const int32_t* __restrict__ pos, const float* __restrict__ key,
                                    int64_t* __restrict__ out_ids, float* __restrict__ out_f, double* __restrict__ out_d) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_ids[i] = pos[i];
    double v = finish_value(metric, key[i]);
    if (out_f) out_f[i] = (float)v;
    if (out_d) out_d[i] = v;
}

// ----------------------------------------------------------------------------- IVFFlat device image

struct Ivf {
    int elem, metric, dim, lists;
    Table centers;
    Table rows;
    int64_t* d_ids = nullptr;
    int64_t* d_list_off = nullptr;
    std::vector<int64_t> h_list_off;
    std::vector<int64_t> sorted_len;  // list lengths, descending (bounds candidates per query)
    int64_t last_bytes = 0, last_cand = 0;
    int64_t* d_cand_sum = nullptr;    // device accumulator of candidates scanned
    ListTile* d_tiles = nullptr;      // static row tiles of the lists (list-major batched scan)
    int n_tiles = 0;
    ListTcImage tc;                   // packed bf16 planes + norms + (list, tile) units, built on first tensor-core scan
    ListTcImage ctc;                  // the same for the centre table (one pseudo list probed by every query)
    int64_t* d_centre_off = nullptr;  // {0, lists}
    cudaStream_t q_stream = nullptr;  // copy stream of the pipelined host path (vb_ivf_prefetch_queries)
    cudaEvent_t q_ready[2] = {nullptr, nullptr};
    void* q_buf[2] = {nullptr, nullptr};
    size_t q_bytes[2] = {0, 0};
    int64_t q_nq[2] = {0, 0};
    unsigned* d_ticket = nullptr;     // 2 x ONE_MAX_Q arrival counters of the fused one-query kernels (zero between launches)
    int* d_tc_fail = nullptr;         // device counters of uncertified queries: [0] probe selection, [1] list scan
    bool defer_tc_check = false;      // batched search: counters are read once, with the results
    bool force_exact = false;         // re-run of a batch whose certificate failed
    bool force_level2 = false;        // re-run of a batch whose level-1 (hi plane only) certificate failed
    int last_list_level = 0;          // filter level the last batched list scan ran at (0 = exact kernels)
    int l1_cooldown = 0;              // batches left before level 1 is tried again after it failed
    int64_t last_tc_failed = 0, total_tc_failed = 0, total_l1_failed = 0;
    bool loaded = false;
    // streaming load (vb_ivf_begin_load / vb_ivf_load_list / vb_ivf_end_load)
    bool loading = false;
    int next_list = 0;
    std::vector<int64_t> h_ids;
    std::vector<int64_t> pending_off;
};

// One CTA per query: candidate offsets of its probed lists and the chunk descriptors of the scan.
__global__ void __launch_bounds__(128) ivf_build_chunks_kernel(const int32_t* __restrict__ probe_lists, int probes,
                                                               const int64_t* __restrict__ list_off, int rows_per_chunk,
                                                               int64_t cap, int32_t* __restrict__ cand_off /*[nq][probes+1]*/,
                                                               int64_t* __restrict__ seg_begin, int32_t* __restrict__ seg_len,
                                                               Chunk* __restrict__ chunks, int* __restrict__ n_chunks,
                                                               int64_t* __restrict__ cand_sum) {
    const int q = blockIdx.x;
    const int32_t* pl = probe_lists + (int64_t)q * probes;
    int32_t* co = cand_off + (int64_t)q * (probes + 1);
    __shared__ int s_base;
    if (threadIdx.x == 0) {
        int32_t off = 0;
        int nch = 0;
        for (int p = 0; p < probes; ++p) {
            co[p] = off;
            int l = pl[p];
            int32_t len = l >= 0 ? (int32_t)(list_off[l + 1] - list_off[l]) : 0;
            off += len;
            nch += (len + rows_per_chunk - 1) / rows_per_chunk;
        }
        co[probes] = off;
        seg_begin[q] = (int64_t)q * cap;
        seg_len[q] = off;
        s_base = chunks ? atomicAdd(n_chunks, nch) : 0;
        atomicAdd((unsigned long long*)cand_sum, (unsigned long long)off);
    }
    __syncthreads();
    if (chunks == nullptr) return;   // the list-major kernels take the probe lists directly: no descriptors needed
    // emit descriptors; per-probe chunk base via a serial prefix held by each thread (probes is small)
    int base = s_base;
    for (int p = 0; p < probes; ++p) {
        int l = pl[p];
        if (l < 0) continue;
        int64_t lo = list_off[l];
        int32_t len = (int32_t)(list_off[l + 1] - lo);
        int nch = (len + rows_per_chunk - 1) / rows_per_chunk;
        for (int c = threadIdx.x; c < nch; c += blockDim.x) {
            Chunk ch;
            ch.row_begin = lo + (int64_t)c * rows_per_chunk;
            ch.n_rows = min(rows_per_chunk, len - c * rows_per_chunk);
            ch.q = q;
            ch.out_off = (int64_t)q * cap + co[p] + (int64_t)c * rows_per_chunk;
            chunks[base + c] = ch;
        }
        base += nch;
    }
}

// winners (position within the query's candidate run) -> heap ids and float8 distances
__global__ void ivf_finish_kernel(int metric, int64_t nq, int k, int probes, const int32_t* __restrict__ pos,
                                  const float* __restrict__ key, const int32_t* __restrict__ probe_lists,
                                  const int32_t* __restrict__ cand_off, const int64_t* __restrict__ list_off,
                                  const int64_t* __restrict__ ids, int64_t* __restrict__ out_ids,
                                  float* __restrict__ out_f, double* __restrict__ out_d) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    int64_t q = i / k;
    int32_t ps = pos[i];
    int64_t id = -1;
    const int32_t* co = cand_off + q * (probes + 1);
    // (a query the tensor-core filter could not select or certify leaves its slots unwritten -- the batch is repeated --
    // so a position is only trusted inside the query's candidate run)
    if (ps >= 0 && ps < co[probes]) {
        int lo = 0, hi = probes;  // largest p with co[p] <= ps
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (co[mid] <= ps) lo = mid;
            else hi = mid;
        }
        // skip empty lists that share the same offset
        while (lo + 1 < probes && co[lo + 1] <= ps) ++lo;
        int l = probe_lists[q * probes + lo];
        int64_t row = list_off[l] + (ps - co[lo]);
        id = ids ? ids[row] : row;
    }
    out_ids[i] = id;
    double v = finish_value(metric, key[i]);
    if (out_f) out_f[i] = (float)v;
    if (out_d) out_d[i] = v;
}

static int64_t ivf_cap(const Ivf& ix, int probes) {
    int64_t cap = 0;
    for (int i = 0; i < probes && i < (int)ix.sorted_len.size(); ++i) cap += ix.sorted_len[(size_t)i];
    return std::max<int64_t>(cap, 1);
}

// every query "probes" pseudo list 0 = the whole centre table: probe_lists[q] = 0, cand_off[q] = {0, lists}
__global__ void centre_pairs_kernel(int64_t nq, int32_t lists, int32_t* __restrict__ probe_lists, int32_t* __restrict__ cand_off) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    probe_lists[q] = 0;
    cand_off[2 * q] = 0;
    cand_off[2 * q + 1] = lists;
}

static int ivf_ensure_centre_tc(Ivf& ix) {
    if (ix.ctc.planes || !ix.ctc.finite) return VB_OK;
    VB_TRY(list_tc_prepare(ix.centers, &ix.ctc));
    std::vector<ListUnit> units;
    for (int64_t t = 0; t * 128 < ix.lists; ++t) units.push_back(ListUnit{0, (int32_t)t});
    ix.ctc.n_units = (int)units.size();
    VB_CUDA(cudaMalloc(&ix.ctc.units, sizeof(ListUnit) * units.size()));
    VB_CUDA(cudaMemcpy(ix.ctc.units, units.data(), sizeof(ListUnit) * units.size(), cudaMemcpyHostToDevice));
    if (!ix.d_centre_off) VB_CUDA(cudaMalloc(&ix.d_centre_off, 2 * sizeof(int64_t)));
    const int64_t off[2] = {0, ix.lists};
    VB_CUDA(cudaMemcpy(ix.d_centre_off, off, sizeof(off), cudaMemcpyHostToDevice));
    return VB_OK;
}

// probe selection for a batch of query images: d_probe_lists [nq x probes] ascending by (distance, list)
static int ivf_select_probes(Ivf& ix, const void* qimg, size_t qstride, int64_t nq, int probes, int32_t** d_lists,
                             float** d_ldist) {
    Context& c = ctx();
    void *d_cdist, *d_seg, *d_probe;
    VB_TRY(workspace(WS_CDIST, sizeof(float) * (size_t)nq * ix.lists, &d_cdist));
    // Query batches: the same tensor-core filter as the list scan, with the centre table as ONE list probed by every
    // query -- approximate distances to all centres, the k' nearest re-scored exactly, order (distance, list number)
    // certified; any uncertified query sends the batch through the exact tiles below.
    const int km = key_metric(ix.metric);
    bool tc = (c.scan_impl == 2 || c.scan_impl == 4) && !ix.force_exact && nq >= 256 && ix.lists >= 128 &&
              list_tc_supported(ix.elem, km, probes);
    if (tc) {
        VB_TRY(ivf_ensure_centre_tc(ix));
        tc = ix.ctc.finite && ix.ctc.planes != nullptr;
    }
    if (tc && !ix.d_tc_fail) {
        VB_CUDA(cudaMalloc(&ix.d_tc_fail, 2 * sizeof(int)));
        VB_CUDA(cudaMemsetAsync(ix.d_tc_fail, 0, 2 * sizeof(int), c.stream));
    }
    if (tc) {
        const int kp = list_tc_kp(probes);
        void *d_pairs, *d_seg2, *d_probe2;
        VB_TRY(workspace(WS_MISC, sizeof(int32_t) * (size_t)nq * 3 + 64, &d_pairs));
        int32_t* zero_lists = (int32_t*)d_pairs;
        int32_t* pair_off = zero_lists + nq;
        VB_TRY(workspace(WS_SEG, (sizeof(int64_t) + sizeof(int32_t)) * (size_t)nq * 2 + 64, &d_seg2));
        int64_t* sb = (int64_t*)d_seg2;
        int32_t* sl = (int32_t*)(sb + nq);
        VB_TRY(workspace(WS_PROBES, (sizeof(int32_t) + sizeof(float)) * (size_t)nq * (probes + kp), &d_probe2));
        int32_t* lists = (int32_t*)d_probe2;
        float* ldist = (float*)(lists + (size_t)nq * probes);
        int32_t* pos_kp = (int32_t*)(ldist + (size_t)nq * probes);
        float* key_kp = (float*)(pos_kp + (size_t)nq * kp);
        prof_begin(VB_PROF_SCAN_LISTS);
        centre_pairs_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c.stream>>>(nq, ix.lists, zero_lists, pair_off);
        regular_segments_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c.stream>>>(nq, ix.lists, ix.lists, sb, sl);
        VB_CUDA(cudaGetLastError());
        count_launch(2);
        const float* qn = nullptr;
        VB_TRY(launch_list_tc(ix.centers, ix.ctc, km, qimg, qstride, nq, zero_lists, 1, pair_off, ix.lists, ix.d_centre_off, 1,
                              (float*)d_cdist, &qn, true));
        int n_failed = 0;
        if (!ix.defer_tc_check) VB_CUDA(cudaMemsetAsync(ix.d_tc_fail, 0, sizeof(int), c.stream));
        // a run of a few thousand centre distances is selected inside the refine kernel (one warp streams it in a few
        // steps: cheaper than a launch of the CTA-per-query radix selection, 43 us for 2048 x 1000); long runs are not
        if (c.fused_refine == 3 && ix.lists <= 2048) {
            // one CTA per query: the run of centre distances is short enough to be selected directly
            VB_TRY(launch_list_tc_cta_refine(ix.centers, ix.ctc, km, qimg, qstride, nq, probes, kp, 1, zero_lists, pair_off, ix.d_centre_off,
                                             (const float*)d_cdist, nullptr, ix.lists, 0, sl, qn, lists, ldist, ix.d_tc_fail, 2));
            if (!ix.defer_tc_check) {
                VB_CUDA(cudaMemcpyAsync(&n_failed, ix.d_tc_fail, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
                VB_CUDA(cudaStreamSynchronize(c.stream));
            }
        } else if (c.fused_refine == 2 || (c.fused_refine != 0 && ix.lists <= 4096)) {
            VB_TRY(launch_list_tc_select_refine(ix.centers, ix.ctc, km, qimg, qstride, nq, probes, kp, 1, zero_lists, pair_off, ix.d_centre_off,
                                                (const float*)d_cdist, sb, sl, qn, lists, ldist, ix.d_tc_fail,
                                                ix.defer_tc_check ? nullptr : &n_failed));
        } else if (c.fused_refine != 0) {
            VB_TRY(launch_segment_topk_v((const float*)d_cdist, sb, sl, nullptr, nullptr, nq, kp, pos_kp, key_kp));
            VB_TRY(launch_list_tc_select_refine(ix.centers, ix.ctc, km, qimg, qstride, nq, probes, kp, 1, zero_lists, pair_off, ix.d_centre_off,
                                                (const float*)d_cdist, sb, sl, qn, lists, ldist, ix.d_tc_fail,
                                                ix.defer_tc_check ? nullptr : &n_failed, 2, pos_kp, key_kp));
        } else {
            VB_TRY(launch_segment_topk_v((const float*)d_cdist, sb, sl, nullptr, nullptr, nq, kp, pos_kp, key_kp));
            VB_TRY(launch_list_tc_refine(ix.centers, ix.ctc, km, qimg, qstride, nq, probes, kp, 1, zero_lists, pair_off, ix.d_centre_off, sl,
                                         qn, pos_kp, key_kp, lists, ldist, ix.d_tc_fail, ix.defer_tc_check ? nullptr : &n_failed));
        }
        prof_end(VB_PROF_SCAN_LISTS);
        ix.total_tc_failed += n_failed;
        if (n_failed == 0) {   // (deferred: optimistic, the caller checks the counter with its own synchronisation)
            *d_lists = lists;
            *d_ldist = ldist;
            return VB_OK;
        }
    }
    prof_begin(VB_PROF_SCAN_LISTS);
    // many queries at once: the query image is itself a row table with the centres' stride (vector, bit), so the
    // centre scan is computed tile-wise (both operands staged once per 128 x 128 tile).  Measured on B200 for 2048
    // queries x 1000 centres x 1536-d: 1.31 ms as one-query-at-a-time scans (L2-resident, 12.6 GB of L2 reads).
    const bool tiled = nq >= 64 && ix.elem != VB_HALFVEC && qstride == ix.centers.stride && ctx().scan_impl != 0 &&
                       (key_metric(ix.metric) == VB_L2_SQUARED || key_metric(ix.metric) == VB_NEG_IP || key_metric(ix.metric) == VB_HAMMING);
    if (tiled) {
        Table Q;
        Q.elem = ix.elem;
        Q.dim = ix.dim;
        Q.stride = qstride;
        Q.n = nq;
        Q.d = (uint8_t*)const_cast<void*>(qimg);
        VB_TRY(launch_distance_matrix(Q, ix.metric, ix.centers, ix.lists, (float*)d_cdist, ix.lists));
    } else {
        VB_TRY(launch_scan_regular(ix.centers, key_metric(ix.metric), qimg, qstride, nq, ix.lists, (float*)d_cdist, ix.lists));
    }
    prof_end(VB_PROF_SCAN_LISTS);
    VB_TRY(workspace(WS_SEG, (sizeof(int64_t) + sizeof(int32_t)) * (size_t)nq * 2 + 64, &d_seg));
    int64_t* seg_begin = (int64_t*)d_seg;
    int32_t* seg_len = (int32_t*)(seg_begin + nq);
    regular_segments_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c.stream>>>(nq, ix.lists, ix.lists, seg_begin, seg_len);
    VB_CUDA(cudaGetLastError());
    count_launch();
    VB_TRY(workspace(WS_PROBES, (sizeof(int32_t) + sizeof(float)) * (size_t)nq * probes, &d_probe));
    int32_t* lists = (int32_t*)d_probe;
    float* ldist = (float*)(lists + (size_t)nq * probes);
    std::vector<int64_t> hb;
    std::vector<int32_t> hl;
    if (probes > 2048) {
        hb.resize((size_t)nq);
        hl.assign((size_t)nq, ix.lists);
        for (int64_t i = 0; i < nq; ++i) hb[(size_t)i] = i * ix.lists;
    }
    VB_TRY(launch_segment_topk_v((const float*)d_cdist, seg_begin, seg_len, hb.empty() ? nullptr : hb.data(),
                                 hl.empty() ? nullptr : hl.data(), nq, probes, lists, ldist));
    *d_lists = lists;
    *d_ldist = ldist;
    return VB_OK;
}

// packed planes, row norms and the (list, table tile) work units of the tensor-core scan; built lazily because the
// planes double the index footprint and only batched searches use them
static int ivf_ensure_tc_image(Ivf& ix) {
    if (ix.tc.planes || !ix.tc.finite) return VB_OK;
    {
        // the planes are as large as an fp32 table: keep the exact kernels when they do not fit beside the index
        const size_t need = (size_t)((ix.rows.n + 127) / 128) * 128 * ((size_t)(ix.rows.dim + 63) / 64) * 64 * 4;
        size_t free_b = 0, total_b = 0;
        VB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (free_b < need + ((size_t)4 << 30)) {
            ix.tc.finite = false;
            return VB_OK;
        }
    }
    VB_TRY(list_tc_prepare(ix.rows, &ix.tc));
    std::vector<ListUnit> units;
    for (int l = 0; l < ix.lists; ++l) {
        const int64_t lo = ix.h_list_off[(size_t)l], hi = ix.h_list_off[(size_t)l + 1];
        if (hi <= lo) continue;
        for (int64_t t = lo / 128; t <= (hi - 1) / 128; ++t) units.push_back(ListUnit{l, (int32_t)t});
    }
    ix.tc.n_units = (int)units.size();
    if (!units.empty()) {
        VB_CUDA(cudaMalloc(&ix.tc.units, sizeof(ListUnit) * units.size()));
        VB_CUDA(cudaMemcpy(ix.tc.units, units.data(), sizeof(ListUnit) * units.size(), cudaMemcpyHostToDevice));
    }
    return VB_OK;
}

// scan the given probe lists for a batch of queries and keep the k nearest per query
static int ivf_scan_topk(Ivf& ix, const void* qimg, size_t qstride, int64_t nq, const int32_t* d_lists, int probes, int k,
                         int64_t* out_ids_dev, float* out_f_dev, double* out_d_dev, int32_t** cand_total_dev) {
    Context& c = ctx();
    const int rpc = scan_chunk_rows(ix.rows);
    const int64_t cap = ivf_cap(ix, probes);
    const int64_t max_chunks = nq * (cap / rpc + probes + 1);
    VB_REQUIRE(max_chunks < (int64_t)INT32_MAX, "too many scan chunks (%lld)", (long long)max_chunks);
    void *d_chunks, *d_seg, *d_dist, *d_pos;
    VB_TRY(workspace(WS_CHUNKS, sizeof(Chunk) * (size_t)max_chunks + sizeof(int32_t) * (size_t)nq * (probes + 1) + 64, &d_chunks));
    Chunk* chunks = (Chunk*)d_chunks;
    int32_t* cand_off = (int32_t*)(chunks + max_chunks);
    VB_TRY(workspace(WS_SEG, (sizeof(int64_t) + sizeof(int32_t)) * (size_t)nq * 2 + 64, &d_seg));
    // second half of WS_SEG (first half may still hold the probe-selection segments)
    int64_t* seg_begin = (int64_t*)d_seg;
    int32_t* seg_len = (int32_t*)(seg_begin + nq);
    int* n_chunks = (int*)(seg_len + nq);
    VB_CUDA(cudaMemsetAsync(n_chunks, 0, sizeof(int), c.stream));
    if (!ix.d_cand_sum) {
        VB_CUDA(cudaMalloc(&ix.d_cand_sum, sizeof(int64_t)));
        VB_CUDA(cudaMemsetAsync(ix.d_cand_sum, 0, sizeof(int64_t), c.stream));
    }
    // chunk descriptors are for the per-query scan kernels only; batched scans (tensor-core filter, list-major) skip them
    const bool per_query_scan = !(list_major_supported(ix.elem, key_metric(ix.metric)) && ix.n_tiles > 0 &&
                                  (c.scan_impl >= 3 || (c.scan_impl == 2 && nq * probes >= 256)));
    ivf_build_chunks_kernel<<<(unsigned)nq, per_query_scan ? 128 : 32, 0, c.stream>>>(d_lists, probes, ix.d_list_off, rpc, cap, cand_off, seg_begin,
                                                                                      seg_len, per_query_scan ? chunks : nullptr, n_chunks,
                                                                                      ix.d_cand_sum);
    VB_CUDA(cudaGetLastError());
    count_launch();
    VB_TRY(workspace(WS_DIST, sizeof(float) * (size_t)nq * cap, &d_dist));
    prof_begin(VB_PROF_SCAN_ITEMS);
    // scan_impl: 0 = per-query LDG scan, 1 = per-query bulk-copy scan, 2 = automatic, 3 = list-major fp32 wherever it
    // applies, 4 = tensor-core filter + exact re-score wherever it applies.
    // Automatic: once a batch carries enough (query, probe) pairs to fill the GPU with row tiles, group them by list
    // so each probed list is read once per batch instead of once per query; small k goes through the tensor-core
    // filter (HBM-bound), larger k through the fp32 list-major kernel (FMA-pipe bound).
    const int km = key_metric(ix.metric);
    const bool batched = nq * probes >= 256;
    bool tc = (c.scan_impl == 4 || c.scan_impl == 2) && !ix.force_exact && batched && list_tc_supported(ix.elem, km, k) && ix.rows.n > 0;
    if (tc) {
        VB_TRY(ivf_ensure_tc_image(ix));
        tc = ix.tc.finite;   // rows with Inf / NaN norms have no error bound: exact path
    }
    ix.last_list_level = 0;
    if (tc) {
        // level 1 reads only the hi plane of the rows (half the HBM traffic, error bound 2^-7 |x||q|): it certifies
        // whenever the neighbours are separated by more than that, otherwise the batch is repeated at level 2 (both
        // planes, 2^-12) and level 1 rests for a while
        const int level = (c.tc_level1 && !ix.force_level2 && ix.l1_cooldown == 0 && list_tc_kp(k, 1) <= 128) ? 1 : 2;
        if (ix.l1_cooldown > 0 && !ix.force_level2) --ix.l1_cooldown;
        ix.last_list_level = level;
        const int kp = list_tc_kp(k, level);
        const float* qn = nullptr;
        // slab minima for the selection (vb_common.cuh slab_base): with them the k' nearest are found from 32 k' candidates
        // per query instead of the whole run
        const int64_t cap_s = slab_cap(cap, probes);
        const bool slabs = c.slab_select && c.fused_refine != 2 && kp <= 128 && nq * cap_s < (int64_t)INT32_MAX &&
                           (size_t)cap_s * 4 + 20 * 1024 <= 160 * 1024;
        void* d_smin = nullptr;
        if (slabs) VB_TRY(workspace(WS_SMIN, sizeof(float) * (size_t)nq * cap_s, &d_smin));
        VB_TRY(launch_list_tc(ix.rows, ix.tc, km, qimg, qstride, nq, d_lists, probes, cand_off, cap, ix.d_list_off, ix.lists,
                              (float*)d_dist, &qn, false, level, (float*)d_smin, cap_s));
        prof_end(VB_PROF_SCAN_ITEMS);
        VB_TRY(workspace(WS_POS, (sizeof(int32_t) + sizeof(float)) * (size_t)nq * (k + kp), &d_pos));
        int32_t* pos = (int32_t*)d_pos;
        float* key = (float*)(pos + (size_t)nq * k);
        int32_t* pos_kp = (int32_t*)(key + (size_t)nq * k);
        float* key_kp = (float*)(pos_kp + (size_t)nq * kp);
        prof_begin(VB_PROF_TOPK);
        int n_failed = 0;
        if (!ix.d_tc_fail) {
            VB_CUDA(cudaMalloc(&ix.d_tc_fail, 2 * sizeof(int)));
            VB_CUDA(cudaMemsetAsync(ix.d_tc_fail, 0, 2 * sizeof(int), c.stream));
        }
        if (!ix.defer_tc_check) VB_CUDA(cudaMemsetAsync(ix.d_tc_fail + 1, 0, sizeof(int), c.stream));
        if (c.fused_refine == 3 && slabs && !ix.force_level2) {
            // one CTA per query: slab selection, re-score on eight warps, ranking, certificate (a selection that overflows
            // counts as uncertified: the repeat of the batch takes the kernels below)
            VB_TRY(launch_list_tc_cta_refine(ix.rows, ix.tc, km, qimg, qstride, nq, k, kp, probes, d_lists, cand_off, ix.d_list_off,
                                             (const float*)d_dist, (const float*)d_smin, cap, cap_s, seg_len, qn, pos, key, ix.d_tc_fail + 1,
                                             level));
            if (!ix.defer_tc_check) {
                VB_CUDA(cudaMemcpyAsync(&n_failed, ix.d_tc_fail + 1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
                VB_CUDA(cudaStreamSynchronize(c.stream));
            }
        } else if (c.fused_refine == 2) {
            VB_TRY(launch_list_tc_select_refine(ix.rows, ix.tc, km, qimg, qstride, nq, k, kp, probes, d_lists, cand_off, ix.d_list_off,
                                                (const float*)d_dist, seg_begin, seg_len, qn, pos, key, ix.d_tc_fail + 1,
                                                ix.defer_tc_check ? nullptr : &n_failed, level));
        } else if (c.fused_refine == 1 || c.fused_refine == 3) {
            if (slabs)
                VB_TRY(launch_slab_select((const float*)d_dist, (const float*)d_smin, nq, probes, d_lists, cand_off, ix.d_list_off, cap,
                                          cap_s, seg_begin, seg_len, kp, pos_kp, key_kp));
            else
                VB_TRY(launch_segment_topk_v((const float*)d_dist, seg_begin, seg_len, nullptr, nullptr, nq, kp, pos_kp, key_kp));
            VB_TRY(launch_list_tc_select_refine(ix.rows, ix.tc, km, qimg, qstride, nq, k, kp, probes, d_lists, cand_off, ix.d_list_off,
                                                (const float*)d_dist, seg_begin, seg_len, qn, pos, key, ix.d_tc_fail + 1,
                                                ix.defer_tc_check ? nullptr : &n_failed, level, pos_kp, key_kp));
        } else {
            if (slabs)
                VB_TRY(launch_slab_select((const float*)d_dist, (const float*)d_smin, nq, probes, d_lists, cand_off, ix.d_list_off, cap,
                                          cap_s, seg_begin, seg_len, kp, pos_kp, key_kp));
            else
                VB_TRY(launch_segment_topk_v((const float*)d_dist, seg_begin, seg_len, nullptr, nullptr, nq, kp, pos_kp, key_kp));
            VB_TRY(launch_list_tc_refine(ix.rows, ix.tc, km, qimg, qstride, nq, k, kp, probes, d_lists, cand_off, ix.d_list_off, seg_len, qn,
                                         pos_kp, key_kp, pos, key, ix.d_tc_fail + 1, ix.defer_tc_check ? nullptr : &n_failed, level));
        }
        prof_end(VB_PROF_TOPK);
        ix.last_tc_failed = n_failed;
        ix.total_tc_failed += n_failed;
        if (n_failed == 0) {
            ivf_finish_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, c.stream>>>(ix.metric, nq, k, probes, pos, key, d_lists, cand_off,
                                                                                      ix.d_list_off, ix.d_ids, out_ids_dev, out_f_dev,
                                                                                      out_d_dev);
            VB_CUDA(cudaGetLastError());
            count_launch();
            if (cand_total_dev) *cand_total_dev = seg_len;
            return VB_OK;
        }
        // some certificate failed: the whole batch goes through the exact kernel below (rare by construction)
        prof_begin(VB_PROF_SCAN_ITEMS);
    }
    const bool list_major = list_major_supported(ix.elem, km) && ix.n_tiles > 0 && (c.scan_impl >= 3 || (c.scan_impl == 2 && batched));
    if (list_major) {
        VB_TRY(launch_list_major(ix.rows, km, qimg, qstride, nq, d_lists, probes, cand_off, cap, ix.d_list_off, ix.lists, ix.d_tiles,
                                 ix.n_tiles, (float*)d_dist));
    } else {
        VB_TRY(launch_scan_chunks(ix.rows, km, qimg, qstride, chunks, n_chunks, (int)max_chunks, (float*)d_dist));
    }
    prof_end(VB_PROF_SCAN_ITEMS);
    VB_TRY(workspace(WS_POS, (sizeof(int32_t) + sizeof(float)) * (size_t)nq * k, &d_pos));
    int32_t* pos = (int32_t*)d_pos;
    float* key = (float*)(pos + (size_t)nq * k);
    std::vector<int64_t> hb;
    std::vector<int32_t> hl;
    if (k > 2048) {
        // "sort everything" path (reference semantics of GetScanItems): needs sizes on the host
        hb.resize((size_t)nq);
        hl.resize((size_t)nq);
        VB_CUDA(cudaMemcpyAsync(hl.data(), seg_len, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, c.stream));
        VB_CUDA(cudaStreamSynchronize(c.stream));
        for (int64_t i = 0; i < nq; ++i) hb[(size_t)i] = i * cap;
    }
    prof_begin(VB_PROF_TOPK);
    VB_TRY(launch_segment_topk_v((const float*)d_dist, seg_begin, seg_len, hb.empty() ? nullptr : hb.data(),
                                 hl.empty() ? nullptr : hl.data(), nq, k, pos, key));
    prof_end(VB_PROF_TOPK);
    ivf_finish_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, c.stream>>>(ix.metric, nq, k, probes, pos, key, d_lists, cand_off,
                                                                              ix.d_list_off, ix.d_ids, out_ids_dev, out_f_dev,
                                                                              out_d_dev);
    VB_CUDA(cudaGetLastError());
    count_launch();
    if (cand_total_dev) *cand_total_dev = seg_len;
    return VB_OK;
}

// k-way merge of the per-rank results of the list-sharded scan: gathered [world][nq][k] (ids, distances) -> the k
// nearest per query by (distance, id).  One CTA per query, bitonic sort of the world * k entries in shared memory.
__global__ void __launch_bounds__(128) merge_ranks_kernel(const int64_t* __restrict__ g_ids, const float* __restrict__ g_dist, int world,
                                                          int64_t nq, int k, int P, int64_t* __restrict__ out_ids,
                                                          float* __restrict__ out_dist) {
    extern __shared__ uint64_t mk[];            // P keys: orderable(distance) << 32 | slot
    const int64_t q = blockIdx.x;
    const int total = world * k;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < total) {
            const int r = i / k, j = i % k;
            const size_t at = ((size_t)r * nq + q) * k + j;
            if (g_ids[at] >= 0) key = ((uint64_t)orderable_key(g_dist[at]) << 32) | (uint32_t)i;
        }
        mk[i] = key;
    }
    __syncthreads();
    // equal distances: the smaller id first (slots are rank-major, so compare ids explicitly on ties)
    for (int size = 2; size <= P; size <<= 1)
        for (int st = size >> 1; st > 0; st >>= 1) {
            for (int a = threadIdx.x; a < P; a += blockDim.x) {
                const int c = a ^ st;
                if (c > a) {
                    const uint64_t x = mk[a], y = mk[c];
                    bool gt = x > y;
                    if ((x >> 32) == (y >> 32) && x != ~0ull && y != ~0ull) {
                        const int sx = (int)(uint32_t)x, sy = (int)(uint32_t)y;
                        const int64_t ix = g_ids[((size_t)(sx / k) * nq + q) * k + sx % k];
                        const int64_t iy = g_ids[((size_t)(sy / k) * nq + q) * k + sy % k];
                        gt = ix > iy;
                    }
                    const bool up = (a & size) == 0;
                    if (gt == up) {
                        mk[a] = y;
                        mk[c] = x;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = mk[i];
        int64_t id = -1;
        float d = __int_as_float(0x7F800000);
        if (key != ~0ull) {
            const int sl = (int)(uint32_t)key;
            const size_t at = ((size_t)(sl / k) * nq + q) * k + sl % k;
            id = g_ids[at];
            d = g_dist[at];
        }
        out_ids[q * k + i] = id;
        out_dist[q * k + i] = d;
    }
}

static int64_t ivf_batch_limit(const Ivf& ix, int probes) {
    // keep the candidate-distance buffer under ~1 GiB
    int64_t cap = ivf_cap(ix, probes);
    int64_t lim = (int64_t)(1ull << 30) / (4 * cap);
    return std::max<int64_t>(1, std::min<int64_t>(lim, 65535));
}

// ----------------------------------------------------------------------------- scans of one query (vb_ivf_one.cu)

static size_t ivf_qstride(const Ivf& ix) {   // stride of the query image upload_queries() builds
    const size_t pad = padded_row_bytes(ix.elem, ix.dim);
    return ix.elem == VB_HALFVEC ? pad * 2 : pad;
}

static int ivf_tickets(Ivf& ix, unsigned** probe_t, unsigned** scan_t) {
    if (!ix.d_ticket) {
        VB_CUDA(cudaMalloc(&ix.d_ticket, 2 * ONE_MAX_Q * sizeof(unsigned)));
        VB_CUDA(cudaMemsetAsync(ix.d_ticket, 0, 2 * ONE_MAX_Q * sizeof(unsigned), ctx().stream));
    }
    *probe_t = ix.d_ticket;
    *scan_t = ix.d_ticket + ONE_MAX_Q;
    return VB_OK;
}

// does a scan of nq queries (probes lists each, k results, at most cap candidates per query) take the fused kernels?
static int64_t one_cap(int64_t cap) { return (cap + 3) & ~(int64_t)3; }   // stride of a query's distance run: 16-byte aligned runs

static bool ivf_one_applies(const Ivf& ix, int64_t nq, int probes, int64_t k, int64_t cap) {
    if (!ctx().one_query || nq < 1 || nq > ONE_MAX_Q || ix.rows.n <= 0) return false;
    const size_t qs = ivf_qstride(ix);
    return one_probe_fits(ix.lists, qs, probes) && one_scan_fits(one_cap(cap), qs, probes, k);
}

// GetScanLists for nq <= ONE_MAX_Q query images: one launch
static int ivf_one_probes(Ivf& ix, const void* qimg, size_t qstride, int64_t nq, int probes, int32_t** d_lists, float** d_ldist) {
    void *d_cdist, *d_probe;
    VB_TRY(workspace(WS_CDIST, sizeof(float) * (size_t)nq * ix.lists, &d_cdist));
    VB_TRY(workspace(WS_PROBES, (sizeof(int32_t) + sizeof(float)) * (size_t)nq * probes, &d_probe));
    int32_t* lists = (int32_t*)d_probe;
    float* ldist = (float*)(lists + (size_t)nq * probes);
    unsigned *tp, *ts;
    VB_TRY(ivf_tickets(ix, &tp, &ts));
    if (!ix.d_cand_sum) VB_CUDA(cudaMalloc(&ix.d_cand_sum, sizeof(int64_t)));
    prof_begin(VB_PROF_SCAN_LISTS);
    VB_TRY(launch_one_probe(ix.centers, key_metric(ix.metric), qimg, qstride, nq, probes, (float*)d_cdist, tp, lists, ldist, ix.d_cand_sum));
    prof_end(VB_PROF_SCAN_LISTS);
    *d_lists = lists;
    *d_ldist = ldist;
    return VB_OK;
}

// GetScanItems + sort for nq <= ONE_MAX_Q query images over device-resident probe lists: one launch
static int ivf_one_items(Ivf& ix, const void* qimg, size_t qstride, int64_t nq, const int32_t* d_lists, int probes, int k, int64_t cap,
                         int64_t* out_ids_dev, float* out_f_dev, double* out_d_dev, bool cand_store) {
    cap = one_cap(cap);
    void* d_dist;
    VB_TRY(workspace(WS_DIST, sizeof(float) * (size_t)nq * cap, &d_dist));
    unsigned *tp, *ts;
    VB_TRY(ivf_tickets(ix, &tp, &ts));
    if (!ix.d_cand_sum) VB_CUDA(cudaMalloc(&ix.d_cand_sum, sizeof(int64_t)));
    prof_begin(VB_PROF_SCAN_ITEMS);
    VB_TRY(launch_one_scan(ix.rows, key_metric(ix.metric), ix.metric, ix.d_list_off, ix.d_ids, d_lists, probes, qimg, qstride, nq, k, cap,
                           (float*)d_dist, ts, out_ids_dev, out_f_dev, out_d_dev, nullptr, ix.d_cand_sum, cand_store));
    prof_end(VB_PROF_SCAN_ITEMS);
    return VB_OK;
}

// results of a fused scan to host memory: ONE copy (ids and float8 distances are adjacent in the workspace) into the pinned
// staging buffer, one synchronisation
static int ivf_one_fetch(const void* d_out, int64_t n, int64_t* out_ids, double* out_d) {
    void* pin;
    VB_TRY(pinned_buffer2(16 * (size_t)n, &pin));
    VB_CUDA(cudaMemcpyAsync(pin, d_out, 16 * (size_t)n, cudaMemcpyDeviceToHost, ctx().stream));
    VB_CUDA(cudaStreamSynchronize(ctx().stream));
    memcpy(out_ids, pin, 8 * (size_t)n);
    memcpy(out_d, (const uint8_t*)pin + 8 * (size_t)n, 8 * (size_t)n);
    return VB_OK;
}

}  // namespace vb

using namespace vb;

struct vb_ivf {
    Ivf ix;
};

extern "C" {

// ----------------------------------------------------------------------------- operator

int vb_distance_batch(int elem, int metric, int dim, const void* q, const void* rows, int64_t n, double* out) {
    VB_TRY(require_init());
    VB_REQUIRE(elem >= 0 && elem <= 2 && dim > 0 && metric_valid_for(elem, metric), "bad element type/metric/dim (%d, %d, %d)", elem,
               metric, dim);
    if (n <= 0) return VB_OK;
    if (q == nullptr) {  // ZeroDistance (src/ivfscan.c:192-196)
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
        return VB_OK;
    }
    Context& c = ctx();
    Table t;
    t.elem = elem;
    t.dim = dim;
    t.stride = padded_row_bytes(elem, dim);
    int rc = table_append_host(t, rows, n);
    if (rc != VB_OK) {
        table_free(t);
        return rc;
    }
    void* qimg;
    size_t qstride;
    void* d_out;
    rc = upload_queries(elem, dim, q, 1, true, WS_QIMG, &qimg, &qstride);
    if (rc == VB_OK) rc = workspace(WS_OUT, sizeof(double) * (size_t)n, &d_out);
    if (rc == VB_OK) rc = launch_scan_regular_f64(t, key_metric(metric), qimg, qstride, 1, n, (double*)d_out, n);
    if (rc == VB_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess) {
            set_error("copy back failed: %s", cudaGetErrorString(e));
            rc = VB_ECUDA;
        }
    }
    table_free(t);
    if (rc != VB_OK) return rc;
    // operator epilogues that are not the key metric (fp64, like the fmgr wrappers)
    if (metric == VB_L2)
        for (int64_t i = 0; i < n; ++i) out[i] = sqrt(out[i]);
    else if (metric == VB_IP)
        for (int64_t i = 0; i < n; ++i) out[i] = -out[i];
    else if (metric == VB_SPHERICAL)
        for (int64_t i = 0; i < n; ++i) {
            double d = -out[i];  // src/vector.c:714-722
            if (d > 1) d = 1;
            else if (d < -1) d = -1;
            out[i] = acos(d) / M_PI;
        }
    return VB_OK;
}

// ----------------------------------------------------------------------------- tables

int vb_table_create(int elem, int dim, vb_table** out) {
    VB_TRY(require_init());
    VB_REQUIRE(elem >= 0 && elem <= 2 && dim > 0 && out, "bad table arguments");
    vb_table* t = new vb_table();
    t->t.elem = elem;
    t->t.dim = dim;
    t->t.stride = padded_row_bytes(elem, dim);
    *out = t;
    return VB_OK;
}
int vb_table_append(vb_table* t, const void* rows, int64_t n) {
    VB_TRY(require_init());
    VB_REQUIRE(t && (rows || n == 0), "null table/rows");
    return table_append_host(t->t, rows, n);
}
int vb_table_append_dev(vb_table* t, const void* rows_dev, int64_t n) {
    VB_TRY(require_init());
    VB_REQUIRE(t && (rows_dev || n == 0), "null table/rows");
    return table_append_dev(t->t, rows_dev, n);
}
int64_t vb_table_rows(const vb_table* t) { return t ? t->t.n : 0; }
const void* vb_table_device_rows(const vb_table* t, size_t* stride_bytes) {
    if (!t) return nullptr;
    if (stride_bytes) *stride_bytes = t->t.stride;
    return t->t.d;
}
int vb_table_free(vb_table* t) {
    if (t) {
        table_free(t->t);
        delete t;
    }
    return VB_OK;
}

static int exact_topk_impl(vb_table* t, int metric, const void* queries, int64_t nq, int k, bool host, int64_t* out_ids,
                           float* out_f, double* out_d) {
    VB_TRY(require_init());
    VB_REQUIRE(t && metric_valid_for(t->t.elem, metric) && metric != VB_SPHERICAL, "bad table/metric");
    VB_REQUIRE(metric != VB_JACCARD || true, "");
    VB_REQUIRE(k > 0, "k must be positive");
    if (nq <= 0) return VB_OK;
    Context& c = ctx();
    Table& T = t->t;
    const int64_t n = T.n;
    const size_t rawq = raw_row_bytes(T.elem, T.dim);
    // sub-batch so the distance matrix stays under ~1 GiB
    int64_t bq = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)(1ull << 30) / (4 * std::max<int64_t>(n, 1))));
    for (int64_t q0 = 0; q0 < nq; q0 += bq) {
        int64_t m = std::min(bq, nq - q0);
        void *qimg, *d_dist, *d_seg, *d_pos, *d_ids, *d_of;
        size_t qstride;
        VB_TRY(upload_queries(T.elem, T.dim, (const uint8_t*)queries + (size_t)q0 * rawq, m, host, WS_QIMG, &qimg, &qstride));
        VB_TRY(workspace(WS_DIST, sizeof(float) * (size_t)m * std::max<int64_t>(n, 1), &d_dist));
        // A batch of queries against the whole table is a distance matrix: tile it (table rows read once per 128
        // queries instead of once per query) when the query image has the table's row layout.  One query at a
        // time, or a metric with a per-row epilogue, streams the table through the scan kernel instead.
        const int km = key_metric(metric);
        const bool tiled = m >= 64 && n >= 128 && n <= 65535LL * 128 && T.elem != VB_HALFVEC && qstride == T.stride && c.scan_impl != 0 &&
                           (km == VB_L2_SQUARED || km == VB_NEG_IP || km == VB_HAMMING);
        if (tiled) {
            Table Q;
            Q.elem = T.elem;
            Q.dim = T.dim;
            Q.stride = qstride;
            Q.n = m;
            Q.d = (uint8_t*)qimg;
            VB_TRY(launch_distance_matrix(Q, km, T, (int)n, (float*)d_dist, n));
        } else {
            VB_TRY(launch_scan_regular(T, km, qimg, qstride, m, n, (float*)d_dist, n));
        }
        VB_TRY(workspace(WS_SEG, (sizeof(int64_t) + sizeof(int32_t)) * (size_t)m + 64, &d_seg));
        int64_t* seg_begin = (int64_t*)d_seg;
        int32_t* seg_len = (int32_t*)(seg_begin + m);
        regular_segments_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c.stream>>>(m, n, (int32_t)n, seg_begin, seg_len);
        VB_CUDA(cudaGetLastError());
        count_launch();
        VB_TRY(workspace(WS_POS, (sizeof(int32_t) + sizeof(float)) * (size_t)m * k, &d_pos));
        int32_t* pos = (int32_t*)d_pos;
        float* key = (float*)(pos + (size_t)m * k);
        std::vector<int64_t> hb;
        std::vector<int32_t> hl;
        if (k > 2048) {
            hb.resize((size_t)m);
            hl.assign((size_t)m, (int32_t)n);
            for (int64_t i = 0; i < m; ++i) hb[(size_t)i] = i * n;
        }
        VB_TRY(launch_segment_topk_v((const float*)d_dist, seg_begin, seg_len, hb.empty() ? nullptr : hb.data(),
                                     hl.empty() ? nullptr : hl.data(), m, k, pos, key));
        int64_t* o_ids;
        float* o_f = nullptr;
        double* o_d = nullptr;
        if (host) {
            VB_TRY(workspace(WS_OUT, (sizeof(int64_t) + sizeof(double)) * (size_t)m * k, &d_ids));
            o_ids = (int64_t*)d_ids;
            o_d = (double*)(o_ids + (size_t)m * k);
        } else {
            o_ids = out_ids + q0 * k;
            o_f = out_f + q0 * k;
        }
        (void)d_of;
        finish_exact_kernel<<<(unsigned)((m * k + 255) / 256), 256, 0, c.stream>>>(metric, m * k, pos, key, o_ids, o_f, o_d);
        VB_CUDA(cudaGetLastError());
        count_launch();
        if (host) {
            VB_CUDA(cudaMemcpyAsync(out_ids + q0 * k, o_ids, sizeof(int64_t) * (size_t)m * k, cudaMemcpyDeviceToHost, c.stream));
            VB_CUDA(cudaMemcpyAsync(out_d + q0 * k, o_d, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, c.stream));
            VB_CUDA(cudaStreamSynchronize(c.stream));
        }
    }
    return VB_OK;
}

int vb_exact_topk(vb_table* t, int metric, const void* queries, int64_t nq, int k, int64_t* out_ids, double* out_dist) {
    return exact_topk_impl(t, metric, queries, nq, k, true, out_ids, nullptr, out_dist);
}
int vb_exact_topk_dev(vb_table* t, int metric, const void* queries_dev, int64_t nq, int k, int64_t* out_ids_dev,
                      float* out_dist_dev) {
    return exact_topk_impl(t, metric, queries_dev, nq, k, false, out_ids_dev, out_dist_dev, nullptr);
}

// ----------------------------------------------------------------------------- IVFFlat

int vb_ivf_create(int elem, int metric, int dim, int lists, vb_ivf** out) {
    VB_TRY(require_init());
    VB_REQUIRE(out && elem >= 0 && elem <= 2 && dim > 0 && lists >= 1 && lists <= 32768, "bad ivfflat arguments (lists 1..32768, src/ivfflat.h:56-57)");
    bool ok = elem == VB_BIT ? metric == VB_HAMMING : (metric == VB_L2_SQUARED || metric == VB_NEG_IP);
    VB_REQUIRE(ok, "ivfflat opclass proc 1 must be L2 squared / negative inner product (vector, halfvec) or Hamming (bit)");
    vb_ivf* h = new vb_ivf();
    Ivf& ix = h->ix;
    ix.elem = elem;
    ix.metric = metric;
    ix.dim = dim;
    ix.lists = lists;
    ix.centers.elem = ix.rows.elem = elem;
    ix.centers.dim = ix.rows.dim = dim;
    ix.centers.stride = ix.rows.stride = padded_row_bytes(elem, dim);
    *out = h;
    return VB_OK;
}

static int ivf_set_offsets(Ivf& ix, const int64_t* list_offsets) {
    ix.h_list_off.assign(list_offsets, list_offsets + ix.lists + 1);
    VB_REQUIRE(ix.h_list_off[0] == 0, "list_offsets[0] must be 0");
    ix.sorted_len.resize((size_t)ix.lists);
    for (int l = 0; l < ix.lists; ++l) {
        int64_t len = ix.h_list_off[(size_t)l + 1] - ix.h_list_off[(size_t)l];
        VB_REQUIRE(len >= 0 && len < (int64_t)INT32_MAX, "bad list length");
        ix.sorted_len[(size_t)l] = len;
    }
    std::sort(ix.sorted_len.begin(), ix.sorted_len.end(), std::greater<int64_t>());
    if (!ix.d_list_off) VB_CUDA(cudaMalloc(&ix.d_list_off, sizeof(int64_t) * ((size_t)ix.lists + 1)));
    VB_CUDA(cudaMemcpy(ix.d_list_off, ix.h_list_off.data(), sizeof(int64_t) * ((size_t)ix.lists + 1), cudaMemcpyHostToDevice));
    // row tiles of the list-major scan: fixed by the list boundaries, so built once per load
    std::vector<ListTile> tiles;
    const int tr = list_tile_rows();
    for (int l = 0; l < ix.lists; ++l) {
        const int64_t lo = ix.h_list_off[(size_t)l], hi = ix.h_list_off[(size_t)l + 1];
        for (int64_t r = lo; r < hi; r += tr) tiles.push_back(ListTile{r, l, (int32_t)std::min<int64_t>(tr, hi - r)});
    }
    if (ix.d_tiles) cudaFree(ix.d_tiles);
    ix.d_tiles = nullptr;
    list_tc_release(&ix.tc);   // rows are about to change: planes are rebuilt on the next tensor-core scan
    list_tc_release(&ix.ctc);
    ix.n_tiles = (int)tiles.size();
    if (ix.n_tiles) {
        VB_CUDA(cudaMalloc(&ix.d_tiles, sizeof(ListTile) * tiles.size()));
        VB_CUDA(cudaMemcpy(ix.d_tiles, tiles.data(), sizeof(ListTile) * tiles.size(), cudaMemcpyHostToDevice));
    }
    return VB_OK;
}

int vb_ivf_load(vb_ivf* h, const void* centers, const int64_t* list_offsets, const void* rows, const int64_t* ids) {
    VB_TRY(require_init());
    VB_REQUIRE(h && centers && list_offsets, "null argument");
    Ivf& ix = h->ix;
    table_free(ix.centers);
    table_free(ix.rows);
    VB_TRY(ivf_set_offsets(ix, list_offsets));
    const int64_t n = ix.h_list_off[(size_t)ix.lists];
    VB_TRY(table_append_host(ix.centers, centers, ix.lists));
    VB_TRY(table_append_host(ix.rows, rows, n));
    if (ix.d_ids) cudaFree(ix.d_ids);
    ix.d_ids = nullptr;
    if (ids && n > 0) {
        VB_CUDA(cudaMalloc(&ix.d_ids, sizeof(int64_t) * (size_t)n));
        VB_CUDA(cudaMemcpy(ix.d_ids, ids, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    ix.loaded = true;
    return VB_OK;
}

int vb_ivf_load_dev(vb_ivf* h, const void* centers_dev, const int64_t* list_offsets_host, const void* rows_dev,
                    const int64_t* ids_dev) {
    VB_TRY(require_init());
    VB_REQUIRE(h && centers_dev && list_offsets_host, "null argument");
    Ivf& ix = h->ix;
    table_free(ix.centers);
    table_free(ix.rows);
    VB_TRY(ivf_set_offsets(ix, list_offsets_host));
    const int64_t n = ix.h_list_off[(size_t)ix.lists];
    VB_TRY(table_append_dev(ix.centers, centers_dev, ix.lists));
    VB_TRY(table_append_dev(ix.rows, rows_dev, n));
    if (ix.d_ids) cudaFree(ix.d_ids);
    ix.d_ids = nullptr;
    if (ids_dev && n > 0) {
        VB_CUDA(cudaMalloc(&ix.d_ids, sizeof(int64_t) * (size_t)n));
        VB_CUDA(cudaMemcpyAsync(ix.d_ids, ids_dev, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToDevice, ctx().stream));
    }
    VB_CUDA(cudaStreamSynchronize(ctx().stream));
    ix.loaded = true;
    return VB_OK;
}

int64_t vb_ivf_rows(const vb_ivf* h) { return h ? h->ix.rows.n : 0; }

// ---- list-at-a-time loading: the packer walks one entry-page chain at a time (src/ivfscan.c:139-179) and never holds
// more than one list on the host

int vb_ivf_begin_load(vb_ivf* h, const void* centers) {
    VB_TRY(require_init());
    VB_REQUIRE(h && centers, "null argument");
    Ivf& ix = h->ix;
    table_free(ix.centers);
    table_free(ix.rows);
    VB_TRY(table_append_host(ix.centers, centers, ix.lists));
    ix.loaded = false;
    ix.loading = true;
    ix.next_list = 0;
    ix.h_ids.clear();
    ix.pending_off.assign((size_t)ix.lists + 1, 0);
    return VB_OK;
}

int vb_ivf_load_list(vb_ivf* h, int list, const void* rows, const int64_t* ids, int64_t n) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loading, "vb_ivf_load_list outside vb_ivf_begin_load / vb_ivf_end_load");
    Ivf& ix = h->ix;
    VB_REQUIRE(list >= ix.next_list && list < ix.lists, "lists must arrive in ascending order (got %d, expected >= %d)", list, ix.next_list);
    VB_REQUIRE(n >= 0 && (n == 0 || (rows && ids)), "null rows / ids");
    for (int l = ix.next_list; l <= list; ++l) ix.pending_off[(size_t)l] = ix.rows.n;
    if (n > 0) {
        VB_TRY(table_append_host(ix.rows, rows, n));
        ix.h_ids.insert(ix.h_ids.end(), ids, ids + n);
    }
    ix.next_list = list + 1;
    ix.pending_off[(size_t)list + 1] = ix.rows.n;
    return VB_OK;
}

int vb_ivf_end_load(vb_ivf* h) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loading, "vb_ivf_end_load without vb_ivf_begin_load");
    Ivf& ix = h->ix;
    for (int l = ix.next_list; l <= ix.lists; ++l) ix.pending_off[(size_t)l] = ix.rows.n;
    ix.loading = false;
    VB_TRY(ivf_set_offsets(ix, ix.pending_off.data()));
    if (ix.d_ids) cudaFree(ix.d_ids);
    ix.d_ids = nullptr;
    const int64_t n = ix.rows.n;
    if (n > 0) {
        VB_CUDA(cudaMalloc(&ix.d_ids, sizeof(int64_t) * (size_t)n));
        VB_CUDA(cudaMemcpy(ix.d_ids, ix.h_ids.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    ix.h_ids.clear();
    ix.h_ids.shrink_to_fit();
    ix.loaded = true;
    return VB_OK;
}

// One list of a loaded image changed (insert into it, vacuum of it): only that list crosses PCIe; the rows behind it
// move on the device, and the packed planes of the tensor-core filter are rebuilt on the device at the next batched scan.
int vb_ivf_replace_list(vb_ivf* h, int list, const void* rows, const int64_t* ids, int64_t n) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded, "index not loaded");
    Ivf& ix = h->ix;
    VB_REQUIRE(list >= 0 && list < ix.lists && n >= 0 && (n == 0 || (rows && ids)), "bad list / rows");
    Context& c = ctx();
    const int64_t lo = ix.h_list_off[(size_t)list], hi = ix.h_list_off[(size_t)list + 1];
    const int64_t total = ix.rows.n, tail = total - hi, new_total = lo + n + tail;
    Table T;
    T.elem = ix.rows.elem;
    T.dim = ix.rows.dim;
    T.stride = ix.rows.stride;
    VB_TRY(table_reserve(T, std::max<int64_t>(new_total, 1)));
    int64_t* new_ids = nullptr;
    if (new_total > 0) VB_CUDA(cudaMalloc(&new_ids, sizeof(int64_t) * (size_t)new_total));
    if (lo > 0) {
        VB_CUDA(cudaMemcpyAsync(T.d, ix.rows.d, (size_t)lo * T.stride, cudaMemcpyDeviceToDevice, c.stream));
        VB_CUDA(cudaMemcpyAsync(new_ids, ix.d_ids, sizeof(int64_t) * (size_t)lo, cudaMemcpyDeviceToDevice, c.stream));
    }
    T.n = lo;
    int rc = n > 0 ? table_append_host(T, rows, n) : VB_OK;   // (synchronises the stream)
    if (rc == VB_OK && n > 0 &&
        cudaMemcpyAsync(new_ids + lo, ids, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
        rc = VB_ECUDA;
    if (rc == VB_OK && tail > 0) {
        if (cudaMemcpyAsync(T.d + (size_t)(lo + n) * T.stride, ix.rows.d + (size_t)hi * T.stride, (size_t)tail * T.stride,
                            cudaMemcpyDeviceToDevice, c.stream) != cudaSuccess ||
            cudaMemcpyAsync(new_ids + lo + n, ix.d_ids + hi, sizeof(int64_t) * (size_t)tail, cudaMemcpyDeviceToDevice, c.stream) != cudaSuccess)
            rc = VB_ECUDA;
    }
    if (rc == VB_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = VB_ECUDA;
    if (rc != VB_OK) {
        table_free(T);
        if (new_ids) cudaFree(new_ids);
        if (rc == VB_ECUDA) set_error("vb_ivf_replace_list: device copy failed");
        return rc;
    }
    T.n = new_total;
    table_free(ix.rows);
    ix.rows = T;
    if (ix.d_ids) cudaFree(ix.d_ids);
    ix.d_ids = new_ids;
    std::vector<int64_t> off = ix.h_list_off;
    const int64_t delta = n - (hi - lo);
    for (int l = list + 1; l <= ix.lists; ++l) off[(size_t)l] += delta;
    return ivf_set_offsets(ix, off.data());
}

int vb_ivf_free(vb_ivf* h) {
    if (!h) return VB_OK;
    table_free(h->ix.centers);
    table_free(h->ix.rows);
    if (h->ix.d_ids) cudaFree(h->ix.d_ids);
    if (h->ix.d_list_off) cudaFree(h->ix.d_list_off);
    if (h->ix.d_cand_sum) cudaFree(h->ix.d_cand_sum);
    if (h->ix.d_tiles) cudaFree(h->ix.d_tiles);
    list_tc_release(&h->ix.tc);
    list_tc_release(&h->ix.ctc);
    if (h->ix.d_centre_off) cudaFree(h->ix.d_centre_off);
    if (h->ix.d_tc_fail) cudaFree(h->ix.d_tc_fail);
    if (h->ix.d_ticket) cudaFree(h->ix.d_ticket);
    for (int i = 0; i < 2; ++i) {
        if (h->ix.q_buf[i]) cudaFree(h->ix.q_buf[i]);
        if (h->ix.q_ready[i]) cudaEventDestroy(h->ix.q_ready[i]);
    }
    if (h->ix.q_stream) cudaStreamDestroy(h->ix.q_stream);
    delete h;
    return VB_OK;
}

int vb_ivf_scan_lists(vb_ivf* h, const void* queries, int64_t nq, int max_probes, int32_t* out_lists, double* out_dist) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded, "index not loaded");
    VB_REQUIRE(max_probes >= 1, "max_probes must be >= 1");
    Ivf& ix = h->ix;
    Context& c = ctx();
    int probes = std::min(max_probes, ix.lists);  // src/ivfscan.c:279-283
    if (nq <= 0) return VB_OK;
    if (queries == nullptr) {
        // NULL query: every centre at distance 0; with this library's tie rule the first lists win
        for (int64_t q = 0; q < nq; ++q)
            for (int p = 0; p < probes; ++p) {
                out_lists[q * max_probes + p] = p;
                if (out_dist) out_dist[q * max_probes + p] = 0.0;
            }
        return VB_OK;
    }
    void* qimg;
    size_t qstride;
    VB_TRY(upload_queries(ix.elem, ix.dim, queries, nq, true, WS_QIMG, &qimg, &qstride));
    int32_t* d_lists;
    float* d_ldist;
    if (c.one_query && nq <= ONE_MAX_Q && one_probe_fits(ix.lists, qstride, probes)) {
        // one backend, one scan: distances to the centres and the selection in a single launch; list numbers and
        // distances (adjacent in the workspace) come back with one copy
        VB_TRY(ivf_one_probes(ix, qimg, qstride, nq, probes, &d_lists, &d_ldist));
        void* pin;
        const size_t np = (size_t)nq * probes;
        VB_TRY(pinned_buffer2(8 * np, &pin));
        VB_CUDA(cudaMemcpyAsync(pin, d_lists, 8 * np, cudaMemcpyDeviceToHost, c.stream));
        VB_CUDA(cudaStreamSynchronize(c.stream));
        const int32_t* pl = (const int32_t*)pin;
        const float* pd = (const float*)(pl + np);
        for (int64_t q = 0; q < nq; ++q)
            for (int p = 0; p < max_probes; ++p) {
                const bool have = p < probes;
                out_lists[q * max_probes + p] = have ? pl[(size_t)(q * probes + p)] : -1;
                if (out_dist) out_dist[q * max_probes + p] = have ? (double)pd[(size_t)(q * probes + p)] : INFINITY;
            }
        return VB_OK;
    }
    VB_TRY(ivf_select_probes(ix, qimg, qstride, nq, probes, &d_lists, &d_ldist));
    std::vector<int32_t> hl((size_t)nq * probes);
    std::vector<float> hd((size_t)nq * probes);
    VB_CUDA(cudaMemcpyAsync(hl.data(), d_lists, sizeof(int32_t) * hl.size(), cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaMemcpyAsync(hd.data(), d_ldist, sizeof(float) * hd.size(), cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaStreamSynchronize(c.stream));
    for (int64_t q = 0; q < nq; ++q)
        for (int p = 0; p < max_probes; ++p) {
            bool have = p < probes;
            out_lists[q * max_probes + p] = have ? hl[(size_t)(q * probes + p)] : -1;
            if (out_dist) out_dist[q * max_probes + p] = have ? (double)hd[(size_t)(q * probes + p)] : INFINITY;
        }
    return VB_OK;
}

int vb_ivf_scan_items(vb_ivf* h, const void* q, const int32_t* lists, int nlists, int64_t cap, int64_t* out_ids, double* out_dist,
                      int64_t* n_out) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded, "index not loaded");
    VB_REQUIRE(nlists >= 0 && lists && n_out, "bad arguments");
    Ivf& ix = h->ix;
    Context& c = ctx();
    int64_t total = 0;
    for (int i = 0; i < nlists; ++i) {
        VB_REQUIRE(lists[i] >= 0 && lists[i] < ix.lists, "list %d out of range", lists[i]);
        total += ix.h_list_off[(size_t)lists[i] + 1] - ix.h_list_off[(size_t)lists[i]];
    }
    *n_out = total;
    int64_t k = std::min(cap, total);
    if (k <= 0) return VB_OK;
    if (q == nullptr) {
        // NULL query: all distances 0, every probed row returned in scan order (src/ivfscan.c:207-211)
        std::vector<int64_t> hid;
        int64_t w = 0;
        for (int i = 0; i < nlists && w < k; ++i) {
            int64_t lo = ix.h_list_off[(size_t)lists[i]], hi = ix.h_list_off[(size_t)lists[i] + 1];
            int64_t m = std::min(hi - lo, k - w);
            if (ix.d_ids) VB_CUDA(cudaMemcpy(out_ids + w, ix.d_ids + lo, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToHost));
            else
                for (int64_t j = 0; j < m; ++j) out_ids[w + j] = lo + j;
            for (int64_t j = 0; j < m; ++j) out_dist[w + j] = 0.0;
            w += m;
        }
        return VB_OK;
    }
    VB_REQUIRE(k < (int64_t)INT32_MAX, "too many candidates");
    void* qimg;
    size_t qstride;
    VB_TRY(upload_queries(ix.elem, ix.dim, q, 1, true, WS_QIMG, &qimg, &qstride));
    void* d_misc;
    VB_TRY(workspace(WS_MISC, sizeof(int32_t) * (size_t)nlists, &d_misc));
    void* d_out;
    VB_TRY(workspace(WS_OUT, (sizeof(int64_t) + sizeof(double)) * (size_t)k, &d_out));
    int64_t* o_ids = (int64_t*)d_out;
    double* o_d = (double*)(o_ids + k);
    if (ivf_one_applies(ix, 1, nlists, k, total)) {
        // distances, selection, heap ids and the operator's epilogue in one launch.  The list numbers go out through the
        // second pinned buffer (no synchronisation), the results come back through it with one copy.
        void* pin;
        VB_TRY(pinned_buffer2(std::max(sizeof(int32_t) * (size_t)nlists, 16 * (size_t)k), &pin));
        memcpy(pin, lists, sizeof(int32_t) * (size_t)nlists);
        VB_CUDA(cudaMemcpyAsync(d_misc, pin, sizeof(int32_t) * (size_t)nlists, cudaMemcpyHostToDevice, c.stream));
        VB_TRY(ivf_one_items(ix, qimg, qstride, 1, (const int32_t*)d_misc, nlists, (int)k, total, o_ids, nullptr, o_d, true));
        ix.last_cand = -1;
        ix.last_bytes = 1;
        return ivf_one_fetch(d_out, k, out_ids, out_dist);
    }
    VB_CUDA(cudaMemcpyAsync(d_misc, lists, sizeof(int32_t) * (size_t)nlists, cudaMemcpyHostToDevice, c.stream));
    VB_CUDA(cudaStreamSynchronize(c.stream));
    // capacity bound must cover these particular lists
    std::vector<int64_t> saved = ix.sorted_len;
    ix.sorted_len.assign(1, total);
    int rc = ivf_scan_topk(ix, qimg, qstride, 1, (const int32_t*)d_misc, nlists, (int)k, o_ids, nullptr, o_d, nullptr);
    ix.sorted_len = saved;
    VB_TRY(rc);
    VB_CUDA(cudaMemcpyAsync(out_ids, o_ids, sizeof(int64_t) * (size_t)k, cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaMemcpyAsync(out_dist, o_d, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaStreamSynchronize(c.stream));
    return VB_OK;
}

// host: results go to host memory (int64 ids + float8 distances); q_host: the queries are host memory
static int ivf_search_impl(vb_ivf* h, const void* queries, int64_t nq, int probes, int k, bool host, bool q_host, int64_t* out_ids,
                           float* out_f, double* out_d) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded, "index not loaded");
    VB_REQUIRE(queries && probes >= 1 && k >= 1, "bad search arguments");
    Ivf& ix = h->ix;
    Context& c = ctx();
    probes = std::min(probes, ix.lists);
    if (nq <= 0) return VB_OK;
    const size_t rawq = raw_row_bytes(ix.elem, ix.dim);
    if (ivf_one_applies(ix, nq, probes, k, ivf_cap(ix, probes))) {
        // a handful of queries (one backend's scan): two fused launches, no memsets, one copy back
        const int64_t cap = ivf_cap(ix, probes);
        void* qimg;
        size_t qstride;
        VB_TRY(upload_queries(ix.elem, ix.dim, queries, nq, q_host, WS_QIMG, &qimg, &qstride));
        int32_t* d_lists;
        float* d_ldist;
        VB_TRY(ivf_one_probes(ix, qimg, qstride, nq, probes, &d_lists, &d_ldist));
        if (host) {
            void* d_out;
            VB_TRY(workspace(WS_OUT, (sizeof(int64_t) + sizeof(double)) * (size_t)nq * k, &d_out));
            int64_t* o_ids = (int64_t*)d_out;
            double* o_d = (double*)(o_ids + (size_t)nq * k);
            VB_TRY(ivf_one_items(ix, qimg, qstride, nq, d_lists, probes, k, cap, o_ids, nullptr, o_d, false));
            VB_TRY(ivf_one_fetch(d_out, nq * k, out_ids, out_d));
        } else {
            VB_TRY(ivf_one_items(ix, qimg, qstride, nq, d_lists, probes, k, cap, out_ids, out_f, nullptr, false));
        }
        ix.last_cand = -1;
        ix.last_bytes = nq;
        return VB_OK;
    }
    const int64_t bq = ivf_batch_limit(ix, probes);
    if (!ix.d_cand_sum) VB_CUDA(cudaMalloc(&ix.d_cand_sum, sizeof(int64_t)));
    VB_CUDA(cudaMemsetAsync(ix.d_cand_sum, 0, sizeof(int64_t), c.stream));
    // One pass of a sub-batch.  The tensor-core filter runs optimistically: its certificate counters are read back
    // together with the results (one synchronisation per sub-batch); a sub-batch with an uncertified query is run
    // again on the exact kernels.
    auto run = [&](int64_t q0, int64_t m, int mode, int* fails) -> int {   // mode 0: automatic, 1: filter level 2, 2: exact
        const bool exact = mode == 2;
        ix.force_exact = exact;
        ix.force_level2 = mode == 1;
        ix.defer_tc_check = !exact;
        if (ix.d_tc_fail) VB_CUDA(cudaMemsetAsync(ix.d_tc_fail, 0, 2 * sizeof(int), c.stream));
        void* qimg;
        size_t qstride;
        VB_TRY(upload_queries(ix.elem, ix.dim, (const uint8_t*)queries + (size_t)q0 * rawq, m, q_host, WS_QIMG, &qimg, &qstride));
        int32_t* d_lists;
        float* d_ldist;
        VB_TRY(ivf_select_probes(ix, qimg, qstride, m, probes, &d_lists, &d_ldist));
        if (host) {
            void* d_out;
            VB_TRY(workspace(WS_OUT, (sizeof(int64_t) + sizeof(double)) * (size_t)m * k, &d_out));
            int64_t* o_ids = (int64_t*)d_out;
            double* o_d = (double*)(o_ids + (size_t)m * k);
            VB_TRY(ivf_scan_topk(ix, qimg, qstride, m, d_lists, probes, k, o_ids, nullptr, o_d, nullptr));
            VB_CUDA(cudaMemcpyAsync(out_ids + q0 * k, o_ids, sizeof(int64_t) * (size_t)m * k, cudaMemcpyDeviceToHost, c.stream));
            VB_CUDA(cudaMemcpyAsync(out_d + q0 * k, o_d, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, c.stream));
        } else {
            VB_TRY(ivf_scan_topk(ix, qimg, qstride, m, d_lists, probes, k, out_ids + q0 * k, out_f + q0 * k, nullptr, nullptr));
        }
        fails[0] = fails[1] = 0;
        const bool check = !exact && ix.d_tc_fail != nullptr;
        if (check) VB_CUDA(cudaMemcpyAsync(fails, ix.d_tc_fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        if (host || check) VB_CUDA(cudaStreamSynchronize(c.stream));
        return VB_OK;
    };
    int rc = VB_OK;
    for (int64_t q0 = 0; q0 < nq && rc == VB_OK; q0 += bq) {
        const int64_t m = std::min(bq, nq - q0);
        int fails[2];
        rc = run(q0, m, 0, fails);
        if (rc == VB_OK && fails[0] == 0 && fails[1] > 0 && ix.last_list_level == 1) {
            // the hi-plane filter could not separate the neighbours of some query: both planes, and leave level 1 alone
            // for the next batches (the data decides this, not the batch)
            ix.total_l1_failed += fails[1];
            ix.l1_cooldown = 64;
            rc = run(q0, m, 1, fails);
        }
        if (rc == VB_OK && fails[0] + fails[1] > 0) {
            ix.total_tc_failed += fails[0] + fails[1];
            rc = run(q0, m, 2, fails);
        }
    }
    ix.force_exact = false;
    ix.force_level2 = false;
    ix.defer_tc_check = false;
    VB_TRY(rc);
    ix.last_cand = -1;  // fetched lazily
    ix.last_bytes = nq;
    return VB_OK;
}

int vb_ivf_search(vb_ivf* h, const void* queries, int64_t nq, int probes, int k, int64_t* out_ids, double* out_dist) {
    return ivf_search_impl(h, queries, nq, probes, k, true, true, out_ids, nullptr, out_dist);
}
int vb_ivf_search_dev(vb_ivf* h, const void* queries_dev, int64_t nq, int probes, int k, int64_t* out_ids_dev, float* out_dist_dev) {
    return ivf_search_impl(h, queries_dev, nq, probes, k, false, false, out_ids_dev, out_dist_dev, nullptr);
}

// Pipelined host path: the queries of the NEXT call are copied to the device on a second stream while the current
// call computes.  Two slots; a slot is reusable once the search that read it has been synchronised (it has, when
// vb_ivf_search_prefetched returns).
int vb_ivf_prefetch_queries(vb_ivf* h, const void* queries, int64_t nq, int slot) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded && queries && nq > 0 && (slot == 0 || slot == 1), "bad prefetch arguments");
    Ivf& ix = h->ix;
    const size_t raw = raw_row_bytes(ix.elem, ix.dim);
    VB_REQUIRE(ix.elem == VB_VECTOR && raw == padded_row_bytes(ix.elem, ix.dim),
               "query prefetch needs vector queries whose dimension is a multiple of 4 (use vb_ivf_search otherwise)");
    if (!ix.q_stream) {
        VB_CUDA(cudaStreamCreateWithFlags(&ix.q_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) VB_CUDA(cudaEventCreateWithFlags(&ix.q_ready[i], cudaEventDisableTiming));
    }
    const size_t bytes = raw * (size_t)nq;
    if (ix.q_bytes[slot] < bytes) {
        if (ix.q_buf[slot]) {
            VB_CUDA(cudaStreamSynchronize(ctx().stream));
            VB_CUDA(cudaFree(ix.q_buf[slot]));
            ix.q_buf[slot] = nullptr;
            ix.q_bytes[slot] = 0;
        }
        VB_CUDA(cudaMalloc(&ix.q_buf[slot], bytes));
        ix.q_bytes[slot] = bytes;
    }
    VB_CUDA(cudaMemcpyAsync(ix.q_buf[slot], queries, bytes, cudaMemcpyHostToDevice, ix.q_stream));
    VB_CUDA(cudaEventRecord(ix.q_ready[slot], ix.q_stream));
    ix.q_nq[slot] = nq;
    return VB_OK;
}

int vb_ivf_search_prefetched(vb_ivf* h, int slot, int probes, int k, int64_t* out_ids, double* out_dist) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded && (slot == 0 || slot == 1) && h->ix.q_nq[slot] > 0, "no prefetched queries in this slot");
    Ivf& ix = h->ix;
    VB_CUDA(cudaStreamWaitEvent(ctx().stream, ix.q_ready[slot], 0));
    const int64_t nq = ix.q_nq[slot];
    ix.q_nq[slot] = 0;   // consumed: the slot may be refilled as soon as this call returns (it synchronises)
    return ivf_search_impl(h, ix.q_buf[slot], nq, probes, k, true, false, out_ids, nullptr, out_dist);
}

// List-sharded search: this rank's image holds its own lists under the GLOBAL list numbering (the other
// lists are empty) and all centres.  Per batch: probe selection for this rank's slice of the queries -> all-gather of
// the probe lists -> local list scan + top-k for ALL queries -> all-gather of k (distance, id) pairs per rank ->
// k-way merge.  Every rank returns the full result.  Collectives are NCCL calls on the library stream; the
// certificate counters of the tensor-core filter are summed over the ranks before they are read, so all ranks
// repeat a batch together.
int vb_ivf_search_sharded_dev(vb_ivf* h, const void* queries_dev, int64_t nq, int probes, int k, int64_t* out_ids_dev,
                              float* out_dist_dev) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded, "index not loaded");
    VB_REQUIRE(queries_dev && probes >= 1 && k >= 1 && out_ids_dev && out_dist_dev, "bad search arguments");
    Ivf& ix = h->ix;
    Context& c = ctx();
    const int world = comm_world(), rank = comm_rank();
    probes = std::min(probes, ix.lists);
    if (nq <= 0) return VB_OK;
    VB_REQUIRE(nq <= ivf_batch_limit(ix, probes), "sharded search: at most %lld queries per call for this index", (long long)ivf_batch_limit(ix, probes));
    int P = 2;
    while (P < world * k) P <<= 1;
    VB_REQUIRE(P <= 4096, "sharded search: world * k must not exceed 4096");
    const int64_t chunk = (nq + world - 1) / world;
    const int64_t q0 = std::min<int64_t>(nq, (int64_t)rank * chunk), m = std::min<int64_t>(chunk, nq - q0);
    enum { WS_SH_LISTS = 24, WS_SH_RES = 25 };
    void *d_sh, *d_res;
    VB_TRY(workspace(WS_SH_LISTS, sizeof(int32_t) * (size_t)chunk * probes * (world + 1) + 64, &d_sh));
    int32_t* my_lists = (int32_t*)d_sh;                          // [chunk x probes]
    int32_t* all_lists = my_lists + (size_t)chunk * probes;      // [world x chunk x probes] = [nq' x probes]
    const size_t res_ids = sizeof(int64_t) * (size_t)nq * k, res_dist = sizeof(float) * (size_t)nq * k;
    VB_TRY(workspace(WS_SH_RES, (res_ids + res_dist) * (size_t)(world + 1) + 256, &d_res));
    int64_t* my_ids = (int64_t*)d_res;
    float* my_dist = (float*)((uint8_t*)d_res + res_ids);
    int64_t* all_ids = (int64_t*)((uint8_t*)d_res + res_ids + res_dist);
    float* all_dist = (float*)((uint8_t*)all_ids + res_ids * (size_t)world);
    if (!ix.d_cand_sum) VB_CUDA(cudaMalloc(&ix.d_cand_sum, sizeof(int64_t)));
    VB_CUDA(cudaMemsetAsync(ix.d_cand_sum, 0, sizeof(int64_t), c.stream));
    if (!ix.d_tc_fail) VB_CUDA(cudaMalloc(&ix.d_tc_fail, 2 * sizeof(int)));

    auto run = [&](int mode, int* fails) -> int {   // mode 0: automatic, 1: filter level 2, 2: exact
        const bool exact = mode == 2;
        ix.force_exact = exact;
        ix.force_level2 = mode == 1;
        ix.defer_tc_check = !exact;
        VB_CUDA(cudaMemsetAsync(ix.d_tc_fail, 0, 2 * sizeof(int), c.stream));
        void* qimg;
        size_t qstride;
        VB_TRY(upload_queries(ix.elem, ix.dim, queries_dev, nq, false, WS_QIMG, &qimg, &qstride));
        VB_CUDA(cudaMemsetAsync(my_lists, 0xFF, sizeof(int32_t) * (size_t)chunk * probes, c.stream));
        if (m > 0) {
            int32_t* d_lists;
            float* d_ldist;
            VB_TRY(ivf_select_probes(ix, (const uint8_t*)qimg + (size_t)q0 * qstride, qstride, m, probes, &d_lists, &d_ldist));
            VB_CUDA(cudaMemcpyAsync(my_lists, d_lists, sizeof(int32_t) * (size_t)m * probes, cudaMemcpyDeviceToDevice, c.stream));
        }
        VB_TRY(comm_allgather(my_lists, all_lists, (int64_t)sizeof(int32_t) * chunk * probes));
        // (ranks hold `chunk` queries each, the last one possibly fewer: the gathered array is query-major for q < nq)
        VB_TRY(ivf_scan_topk(ix, qimg, qstride, nq, all_lists, probes, k, my_ids, my_dist, nullptr, nullptr));
        // one buffer per rank: [ids | distances]; gathered rank-major, so view it as two strided arrays
        VB_TRY(comm_allgather(my_ids, all_ids, (int64_t)res_ids));
        VB_TRY(comm_allgather(my_dist, all_dist, (int64_t)res_dist));
        merge_ranks_kernel<<<(unsigned)nq, 128, (size_t)P * 8, c.stream>>>(all_ids, all_dist, world, nq, k, P, out_ids_dev, out_dist_dev);
        VB_CUDA(cudaGetLastError());
        count_launch();
        fails[0] = fails[1] = 0;
        if (!exact) {
            VB_TRY(comm_allreduce(ix.d_tc_fail, 2, 1));
            VB_CUDA(cudaMemcpyAsync(fails, ix.d_tc_fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            VB_CUDA(cudaStreamSynchronize(c.stream));
        }
        return VB_OK;
    };
    int fails[2];
    int rc = run(0, fails);
    if (rc == VB_OK && fails[0] == 0 && fails[1] > 0 && ix.last_list_level == 1) {
        ix.total_l1_failed += fails[1];
        ix.l1_cooldown = 64;
        rc = run(1, fails);
    }
    if (rc == VB_OK && fails[0] + fails[1] > 0) {
        ix.total_tc_failed += fails[0] + fails[1];
        rc = run(2, fails);
    }
    ix.force_exact = false;
    ix.force_level2 = false;
    ix.defer_tc_check = false;
    VB_TRY(rc);
    ix.last_cand = -1;
    ix.last_bytes = nq;
    return VB_OK;
}

__global__ void add_id_offset_kernel(int64_t* ids, int64_t n, int64_t offset) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && ids[i] >= 0) ids[i] += offset;
}

// Exact (no index) top-k over a row-sharded table: every rank scans its own rows, the per-rank k nearest
// are all-gathered and merged by (distance, id).  id_offset = the global number of this rank's row 0.
int vb_exact_topk_sharded_dev(vb_table* t, int metric, const void* queries_dev, int64_t nq, int k, int64_t id_offset, int64_t* out_ids_dev,
                              float* out_dist_dev) {
    VB_TRY(require_init());
    VB_REQUIRE(t && queries_dev && out_ids_dev && out_dist_dev && k >= 1, "bad arguments");
    if (nq <= 0) return VB_OK;
    Context& c = ctx();
    const int world = comm_world();
    int P = 2;
    while (P < world * k) P <<= 1;
    VB_REQUIRE(P <= 4096, "sharded exact scan: world * k must not exceed 4096");
    enum { WS_SH_RES = 25 };
    void* d_res;
    const size_t res_ids = sizeof(int64_t) * (size_t)nq * k, res_dist = sizeof(float) * (size_t)nq * k;
    VB_TRY(workspace(WS_SH_RES, (res_ids + res_dist) * (size_t)(world + 1) + 256, &d_res));
    int64_t* my_ids = (int64_t*)d_res;
    float* my_dist = (float*)((uint8_t*)d_res + res_ids);
    int64_t* all_ids = (int64_t*)((uint8_t*)d_res + res_ids + res_dist);
    float* all_dist = (float*)((uint8_t*)all_ids + res_ids * (size_t)world);
    VB_TRY(vb_exact_topk_dev(t, metric, queries_dev, nq, k, my_ids, my_dist));
    add_id_offset_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, c.stream>>>(my_ids, nq * k, id_offset);
    VB_TRY(comm_allgather(my_ids, all_ids, (int64_t)res_ids));
    VB_TRY(comm_allgather(my_dist, all_dist, (int64_t)res_dist));
    merge_ranks_kernel<<<(unsigned)nq, 128, (size_t)P * 8, c.stream>>>(all_ids, all_dist, world, nq, k, P, out_ids_dev, out_dist_dev);
    VB_CUDA(cudaGetLastError());
    count_launch(2);
    return VB_OK;
}

int vb_ivf_search_sharded(vb_ivf* h, const void* queries, int64_t nq, int probes, int k, int64_t* out_ids, double* out_dist) {
    VB_TRY(require_init());
    VB_REQUIRE(h && h->ix.loaded && queries && out_ids && out_dist && k >= 1, "bad search arguments");
    if (nq <= 0) return VB_OK;
    Ivf& ix = h->ix;
    Context& c = ctx();
    const size_t raw = raw_row_bytes(ix.elem, ix.dim);
    enum { WS_SH_Q = 26 };
    void* d_q;
    VB_TRY(workspace(WS_SH_Q, raw * (size_t)nq + (sizeof(int64_t) + sizeof(float)) * (size_t)nq * k + 64, &d_q));
    int64_t* d_ids = (int64_t*)((uint8_t*)d_q + ((raw * (size_t)nq + 15) & ~(size_t)15));
    float* d_dist = (float*)(d_ids + (size_t)nq * k);
    VB_CUDA(cudaMemcpyAsync(d_q, queries, raw * (size_t)nq, cudaMemcpyHostToDevice, c.stream));
    VB_TRY(vb_ivf_search_sharded_dev(h, d_q, nq, probes, k, d_ids, d_dist));
    std::vector<float> hd((size_t)nq * k);
    VB_CUDA(cudaMemcpyAsync(out_ids, d_ids, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaMemcpyAsync(hd.data(), d_dist, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaStreamSynchronize(c.stream));
    for (size_t i = 0; i < hd.size(); ++i) out_dist[i] = (double)hd[i];
    return VB_OK;
}

int vb_ivf_tc_traffic(int on, int64_t* out8) {
    VB_TRY(require_init());
    return list_tc_traffic(on, out8);
}

int64_t vb_ivf_tc_fallbacks(const vb_ivf* h) { return h ? h->ix.total_tc_failed : 0; }
int64_t vb_ivf_tc_level1_fallbacks(const vb_ivf* h) { return h ? h->ix.total_l1_failed : 0; }

int64_t vb_ivf_last_candidates(const vb_ivf* h) {
    if (!h || !h->ix.d_cand_sum) return 0;
    int64_t v = 0;
    cudaStreamSynchronize(ctx().stream);
    cudaMemcpy(&v, h->ix.d_cand_sum, sizeof(int64_t), cudaMemcpyDeviceToHost);
    return v;
}
int64_t vb_ivf_last_scan_bytes(const vb_ivf* h) {
    if (!h) return 0;
    const Ivf& ix = h->ix;
    int64_t cand = vb_ivf_last_candidates(h);
    int64_t nq = ix.last_bytes;  // number of queries of the last search
    return (nq * ix.lists + cand) * (int64_t)raw_row_bytes(ix.elem, ix.dim);
}

}  // extern "C"
