// vb_common.cuh -- shared declarations for libvecb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/vecb200.h"

namespace vb {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern thread_local int g_last_status;

#define VB_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            vb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return VB_ECUDA;                                                             \
        }                                                                                \
    } while (0)

#define VB_TRY(expr)               \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != VB_OK) return rc__; \
    } while (0)

#define VB_REQUIRE(cond, ...)        \
    do {                             \
        if (!(cond)) {               \
            vb::set_error(__VA_ARGS__); \
            return VB_EINVAL;        \
        }                            \
    } while (0)

// ---------------------------------------------------------------- context
struct Context {
    bool inited = false;
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int64_t launches = 0;
    int slab_select = 1;               // selection from the filter's slab minima (0: full radix selection of every run)
    uint64_t query_epoch = 0;          // bumped by every upload_queries(): caches keyed on a query image are valid for one batch only
    int tc_level1 = 1;                 // batched list scan: try the hi-plane-only filter first (vb_set_option "tc_level1")
    int pp_filter = 1;                 // k-means++ on large fp32 sample tables: triangle-inequality + bf16 filters in front of the exact distances
    unsigned long long pp_stats[3] = {0, 0, 0};   // last seeding: samples skipped by the triangle rule / stopped by the bf16 bound / re-scored exactly
    int fused_refine = 3;              // tensor-core filter, after the k' select: 3 = select + exact re-score + certificate with one CTA per query,
                                       // 1 = re-score + certificate in one warp-per-query kernel fed by the selection kernel,
                                       // 0 = rescore_kernel + certify_kernel, 2 = the select runs inside the warp-per-query kernel too
    int one_query = 1;                 // scans of at most 16 queries: two fused distance + select kernels (vb_ivf_one.cu); 0 = the general path
    int scan_impl = 2;                 // 0 = LDG variant (vb_scan.cu), 1 = bulk-copy / TMA variant (vb_scan_bulk.cu), 2 = by table size
    int hnsw_build_fraction = 64;      // HNSW build: a batch is at most 1/fraction of the elements already inserted
    int hnsw_build_batch = 16384;      // ... and at most this many elements
    int hnsw_l2_persist = 1;           // HNSW scans: keep the visited tables in the persisting part of L2
    int64_t last_assign_flagged = -1;  // rows re-checked by the exact kernel in the last tensor-core assign (-1: exact path)
    // grow-only device workspace arenas (index = slot)
    void* ws[32] = {nullptr};
    size_t ws_bytes[32] = {0};
    // pinned staging
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    void* pinned2 = nullptr;
    size_t pinned2_bytes = 0;
};
Context& ctx();
int require_init();
// returns device pointer of at least `bytes` in slot (contents preserved only if not grown)
int workspace(int slot, size_t bytes, void** out);
int pinned_buffer(size_t bytes, void** out);
int pinned_buffer2(size_t bytes, void** out);

inline void count_launch(int n = 1) { ctx().launches += n; }

// optional event brackets around a kernel class (vb_prof_enable)
void prof_begin(int which);
void prof_end(int which);

// ---------------------------------------------------------------- communicator (vb_comm.cu)
// world size / rank of the library's NCCL communicator (1 / 0 when none was created)
int comm_world();
int comm_rank();
// in-place sum over the ranks on the library stream; dtype 0 = fp32, 1 = int32, 2 = int64, 3 = fp64, 4 = uint32
int comm_allreduce(void* buf_dev, int64_t count, int dtype);
// recv[r * bytes .. ) = rank r's send buffer, on the library stream (a plain copy when there is one rank)
int comm_allgather(const void* send_dev, void* recv_dev, int64_t bytes_per_rank);

// ---------------------------------------------------------------- layout
inline size_t raw_row_bytes(int elem, int dim) {
    return elem == VB_VECTOR ? (size_t)dim * 4 : elem == VB_HALFVEC ? (size_t)dim * 2 : ((size_t)dim + 7) / 8;
}
// device rows are padded to 16-byte multiples so every row starts 128-bit aligned
inline size_t padded_row_bytes(int elem, int dim) { return (raw_row_bytes(elem, dim) + 15) & ~(size_t)15; }

inline bool metric_valid_for(int elem, int metric) {
    if (elem == VB_BIT) return metric == VB_HAMMING || metric == VB_JACCARD;
    return metric == VB_L2_SQUARED || metric == VB_NEG_IP || metric == VB_COSINE || metric == VB_L1 ||
           metric == VB_L2 || metric == VB_IP || metric == VB_SPHERICAL;
}
// the metric the kernels order by ("key metric"); the float8 the operator returns is derived from it
inline int key_metric(int metric) {
    switch (metric) {
        case VB_L2: return VB_L2_SQUARED;
        case VB_IP:
        case VB_SPHERICAL: return VB_NEG_IP;
        default: return metric;
    }
}

// A resident row table: n rows, padded stride.
struct Table {
    int elem = 0, dim = 0;
    size_t stride = 0;      // padded row bytes
    int64_t n = 0, cap = 0;
    uint8_t* d = nullptr;   // device
};
int table_reserve(Table& t, int64_t rows);
int table_append_host(Table& t, const void* rows, int64_t n);
int table_append_dev(Table& t, const void* rows_dev, int64_t n);
void table_free(Table& t);

// Pad + (for halfvec) widen queries into the fp32 query image the kernels read.
// vector/halfvec: float[nq][qstride/4]; bit: bytes[nq][qstride]. host==true: `queries` is host memory.
int upload_queries(int elem, int dim, const void* queries, int64_t nq, bool host, int ws_slot, void** out_dev, size_t* qstride);

// ---------------------------------------------------------------- scan primitives (vb_scan.cu)
struct Chunk {           // one unit of scan work: a run of rows against one query
    int64_t row_begin;   // row index into the table
    int64_t out_off;     // where distance[0] of this run goes in the output array
    int32_t n_rows;
    int32_t q;           // query index
};

// distances of regular work: every query against rows [0, n) ; out[q * out_stride + r]
int launch_scan_regular(const Table& t, int key_metric, const void* q_dev, size_t qstride, int64_t nq,
                        int64_t n_rows, float* out, int64_t out_stride);
// distances of chunk-list work; n_chunks_dev holds the chunk count (device int32)
int launch_scan_chunks(const Table& t, int key_metric, const void* q_dev, size_t qstride,
                       const Chunk* chunks_dev, const int* n_chunks_dev, int max_chunks, float* out);
// Jaccard needs the exact float8: out as double
int launch_scan_regular_f64(const Table& t, int key_metric, const void* q_dev, size_t qstride, int64_t nq,
                            int64_t n_rows, double* out, int64_t out_stride);

// per segment top-k of float keys with position tie-break, ascending (vb_scan.cu).
// seg_begin/seg_len: device arrays (offset into keys, length).  Host copies are only needed for
// k > 2048 (full segmented sort path).  Writes out_pos[nseg*k] (position within segment, -1 pad)
// and out_key[nseg*k].
int launch_segment_topk_v(const float* keys, const int64_t* seg_begin_dev, const int32_t* seg_len_dev,
                          const int64_t* seg_begin_host, const int32_t* seg_len_host, int64_t nseg, int k,
                          int32_t* out_pos, float* out_key);
int scan_chunk_rows(const Table& t);

// exact fp32 nearest-centre assign (vb_kmeans.cu) and the default assign entry (tensor cores + exact re-check)
int launch_assign_exact(const Table& X, int metric, const Table& Cn, int k, const int32_t* row_sel_dev, int64_t n_sel,
                        int32_t* out_idx, float* out_val);
int launch_assign(const Table& X, int metric, const Table& Cn, int k, int32_t* out_idx);
// all-pairs fp32 distances X x Cn -> out[x * ld + c] (register-tiled CUDA-core kernel, vb_kmeans.cu)
int launch_distance_matrix(const Table& X, int metric, const Table& Cn, int k, float* out, int64_t ld);
void set_tc_enabled(bool on);

// list-major batched list scan (vb_list_tile.cu): the (query, probe) pairs of a batch grouped by list, one CTA
// per static row tile of the index against every query probing that list
struct ListTile {
    int64_t row_begin;   // first row of the tile in the list-ordered table
    int32_t list;
    int32_t n_rows;      // <= list_tile_rows()
};
// (query, probe) pairs of a batch grouped by list: slots [begin[l], begin[l] + cnt[l]) belong to list l
struct QueryGroups {
    int32_t* cnt;        // [lists]
    int32_t* begin;      // [lists]
    int32_t* gt_begin;   // [lists] first query tile of each list (only when built with gt_rows > 0)
    int32_t* pair_q;     // [pairs] query number
    int32_t* pair_list;  // [pairs] list number
    int64_t* pair_out;   // [pairs] offset of the pair's candidate run in the distance buffer
    int32_t* pair_sbase; // [pairs] first entry of the pair's run in the slab-minimum array (slab_base(), cap_s > 0 only)
    int64_t n_pairs;
};
// Slab minima of the tensor-core filter: a slab = 32 table-aligned rows of one (query, probe) pair's list; the filter's
// epilogue stores min d~ over the slab next to the dense d~ array, and the selection reads only the slabs that can hold
// one of the k' nearest.  Layout: query q owns cap_s = cap / 32 + 2 probes + 2 entries; probe p's run starts at
// q cap_s + cand_off[q][p] / 32 + 2 p (a list of len rows spans at most len / 32 + 2 slabs, so runs never overlap) and
// slab (r >> 5) - (list_off[l] >> 5) of the list is entry number that of the run.
__host__ __device__ inline int64_t slab_cap(int64_t cap, int probes) { return (cap >> 5) + 2 * (int64_t)probes + 2; }
__host__ __device__ inline int64_t slab_base(int64_t q, int64_t cap_s, int32_t cand_off, int p) {
    return q * cap_s + (cand_off >> 5) + 2 * p;
}
int build_query_groups(const int32_t* d_lists, int64_t nq, int probes, const int32_t* cand_off, int64_t cap, int n_lists, int gt_rows,
                       QueryGroups* g, int64_t cap_s = 0);
// tensor-core filter of the batched list scan (vb_list_tc.cu)
struct ListUnit {
    int32_t list;
    int32_t tile;        // 128-row tile of the list-ordered table that intersects the list
};
struct ListTcImage {
    uint8_t* planes = nullptr;   // bf16 hi/lo planes of the whole table, swizzled smem image per (tile, 64-dim block)
    float* xn = nullptr;         // |row|^2
    ListUnit* units = nullptr;
    int n_units = 0;
    int n_kblocks = 0;
    int64_t n_tiles = 0;
    float xmax = 0.f;            // max |row|
    bool finite = true;          // false when a row norm is Inf / NaN (no error bound -> exact path only)
};
bool list_tc_supported(int elem, int key_metric, int k);
int list_tc_kp(int k, int level = 2);
int list_tc_prepare(const Table& rows, ListTcImage* im);
void list_tc_release(ListTcImage* im);
int launch_list_tc(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                   const int32_t* d_lists, int probes, const int32_t* cand_off, int64_t cap, const int64_t* d_list_off, int n_lists,
                   float* out, const float** qn_out, bool one_list_all_queries = false, int level = 2, float* smin = nullptr,
                   int64_t cap_s = 0);
// the k' nearest of every query's candidate run from the slab minima (same output as launch_segment_topk_v)
int launch_slab_select(const float* dist, const float* smin, int64_t nq, int probes, const int32_t* probe_lists, const int32_t* cand_off,
                       const int64_t* list_off, int64_t cap, int64_t cap_s, const int64_t* seg_begin, const int32_t* seg_len, int kp,
                       int32_t* out_pos, float* out_key);
int launch_list_tc_refine(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                          int k, int kp, int probes, const int32_t* d_lists, const int32_t* cand_off, const int64_t* d_list_off,
                          const int32_t* seg_len, const float* qn, const int32_t* pos_kp, const float* approx_kp, int32_t* out_pos,
                          float* out_key, int* fail_dev, int* n_failed_host, int level = 2);
// traffic accounting of list_tc_kernel launches (profiling): enable / read-and-reset 8 counters (lists: 0-3, centres: 4-7)
int list_tc_traffic(int on, int64_t* out8);
int launch_list_tc_select_refine(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                                 int k, int kp, int probes, const int32_t* d_lists, const int32_t* cand_off, const int64_t* d_list_off,
                                 const float* dist, const int64_t* seg_begin, const int32_t* seg_len, const float* qn, int32_t* out_pos,
                                 float* out_key, int* fail_dev, int* n_failed_host, int level = 2, const int32_t* pre_pos = nullptr,
                                 const float* pre_key = nullptr);
// the same three steps with one CTA per query (selection CTA-wide, re-score on eight warps); smin == nullptr: short runs
int launch_list_tc_cta_refine(const Table& rows, const ListTcImage& im, int key_metric, const void* qimg, size_t qstride, int64_t nq,
                              int k, int kp, int probes, const int32_t* d_lists, const int32_t* cand_off, const int64_t* d_list_off,
                              const float* dist, const float* smin, int64_t cap, int64_t cap_s, const int32_t* seg_len, const float* qn,
                              int32_t* out_pos, float* out_key, int* fail_dev, int level = 2);
// vb_ivf_one.cu: the scan of one query (or a handful) as two fused distance + select kernels
bool one_probe_fits(int lists, size_t qstride, int probes);
bool one_scan_fits(int64_t cap, size_t qstride, int probes, int64_t k);
int launch_one_probe(const Table& centres, int key_metric, const void* qimg, size_t qstride, int64_t nq, int probes, float* cdist,
                     unsigned* ticket, int32_t* out_lists, float* out_ldist, int64_t* zero_me);
int launch_one_scan(const Table& rows, int key_metric, int metric, const int64_t* list_off, const int64_t* ids,
                    const int32_t* probe_lists, int probes, const void* qimg, size_t qstride, int64_t nq, int k, int64_t cap,
                    float* dist, unsigned* ticket, int64_t* out_ids, float* out_f, double* out_d, int32_t* out_total,
                    int64_t* cand_sum, bool cand_store);
constexpr int ONE_MAX_Q = 16;   // queries per call the fused path takes
int list_tile_rows();
bool list_major_supported(int elem, int key_metric);
int launch_list_major(const Table& rows, int key_metric, const void* qimg, size_t qstride, int64_t nq, const int32_t* d_lists,
                      int probes, const int32_t* cand_off, int64_t cap, const int64_t* d_list_off, int n_lists,
                      const ListTile* d_tiles, int n_tiles, float* out);

}  // namespace vb

// opaque handle types of the C ABI
struct vb_table {
    vb::Table t;
};
