/*
 * vecb200.h -- C ABI of libvecb200.so: the B200 (sm_100a) implementation of
 * pgvector's batched-distance hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain
 * pointers and sizes; there are no PostgreSQL, C++ or torch types in any
 * signature.  Each function names the reference code whose inner loop it
 * replaces (paths relative to the pgvector tree @ e48241b).  The extension-side
 * glue that calls these from the index AM is in pgvector_b200/ext/ and
 * INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (VB_OK) or a negative VB_E* code; the message is
 *     available from vb_last_error() (thread local).  The caller (a Postgres
 *     backend) turns it into ereport(ERROR) AFTER the call returns, so no device
 *     state is held across a longjmp.
 *   - there is NO CPU fallback: every call fails with VB_ENODEVICE when no
 *     sm_100 device is usable.
 *   - "host" pointers are ordinary process memory; "_dev" variants take device
 *     pointers valid on the library's current device.  Work is enqueued on the
 *     library stream (vb_stream()); host-buffer variants synchronise before
 *     returning.  _dev variants return with their work enqueued, with two stated
 *     exceptions: the batched vb_ivf_search*_dev read ONE 8-byte pair of certificate
 *     counters per sub-batch of queries (the tensor-core filter re-runs an
 *     uncertified batch exactly before the results may be used), and
 *     vb_hnsw_search_dev reads one overflow flag per call (visited-table growth).
 *   - rows are row-major and contiguous in the caller's buffers (vector: dim
 *     fp32; halfvec: dim IEEE binary16; bit: (dim+7)/8 bytes, MSB first, tail
 *     bits zero -- exactly the payload of Vector.x (src/vector.h:18-24),
 *     HalfVector.x (src/halfvec.h:67-73) and VARBITS (src/bitvec.c:16-28)).
 *     Device images pad each row to a multiple of 16 bytes (zero fill).
 *   - heap TIDs are passed opaquely as int64 ids (block << 16 | offset in the glue).
 */
#ifndef VECB200_H
#define VECB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_ABI_VERSION 1

/* status codes */
#define VB_OK 0
#define VB_EINVAL (-1)			/* bad argument (dimension mismatch, unsupported metric for type, ...) */
#define VB_ENODEVICE (-2)		/* no usable sm_100 device / CUDA failure at init */
#define VB_ECUDA (-3)			/* CUDA runtime error */
#define VB_ENOMEM (-4)			/* device or host allocation failed */
#define VB_ESTATE (-5)			/* call sequence error (index not loaded, ...) */

/* element types */
#define VB_VECTOR 0				/* vector   : fp32   (src/vector.h:18-24) */
#define VB_HALFVEC 1			/* halfvec  : fp16   (src/halfvec.h:67-73) */
#define VB_BIT 2				/* bit      : packed (src/bitvec.c:16-28) */

/* metrics: what the SQL operator / opclass support function returns */
#define VB_L2_SQUARED 0			/* vector_l2_squared_distance    src/vector.c:595-605, halfvec.c:575-585 (index proc 1 of l2 opclasses) */
#define VB_NEG_IP 1				/* vector_negative_inner_product src/vector.c:637-647, halfvec.c:605-615 (<#>; proc 1 of ip and cosine opclasses) */
#define VB_COSINE 2				/* cosine_distance               src/vector.c:671-696, halfvec.c:620-645 (<=>, sequential scan only) */
#define VB_L1 3					/* l1_distance                   src/vector.c:740-750, halfvec.c:676-686 (<+>) */
#define VB_HAMMING 4			/* hamming_distance              src/bitvec.c:45-55  (<~>) */
#define VB_JACCARD 5			/* jaccard_distance              src/bitvec.c:60-70  (<%>) */
#define VB_L2 6					/* l2_distance                   src/vector.c:579-589 (<->): sqrt((double) L2^2) */
#define VB_IP 7					/* inner_product                 src/vector.c:622-632 */
#define VB_SPHERICAL 8			/* vector_spherical_distance     src/vector.c:703-722 (k-means proc 3 of ip/cosine opclasses) */

/* ------------------------------------------------------------------ runtime */

/* Bind this process to CUDA device `device` (a backend calls it once, lazily). */
int			vb_init(int device);
/* Release every device allocation made by the library in this process. */
int			vb_shutdown(void);
const char *vb_last_error(void);
int			vb_abi_version(void);
/* cudaStream_t the library launches on (as void*), for event timing / graph capture by the host. */
void	   *vb_stream(void);
/* Kernels launched by this library since vb_init (a claim for bench.py's gpu_launches). */
int64_t		vb_launch_count(void);
int			vb_synchronize(void);
/*
 * Order the library stream after work the caller enqueued on another stream: `cuda_event` is a cudaEvent_t recorded
 * there.  Needed before a _dev call whose inputs were produced on a different stream (vb_stream() is non-blocking).
 */
int			vb_stream_wait_event(void *cuda_event);

/*
 * Optional per-kernel timing with CUDA events on vb_stream(), used by bench.py for the
 * roofline of the dominant kernel (the events add ~1 us per bracketed launch).
 * Kernels: 0 = list/candidate scan (GetScanItems) incl. query grouping / packing, 1 = centre scan (GetScanLists),
 * 2 = top-k select + exact re-score + certificate, 3 = k-means assign, 4 = HNSW search, 5 / 6 = the tensor-core filter
 * kernel alone (lists / centres).
 */
#define VB_PROF_SCAN_ITEMS 0
#define VB_PROF_SCAN_LISTS 1
#define VB_PROF_TOPK 2
#define VB_PROF_ASSIGN 3
#define VB_PROF_HNSW 4
#define VB_PROF_LIST_TC 5		/* list_tc_kernel alone (the tensor-core filter pass over the probed lists) */
#define VB_PROF_CENTRE_TC 6		/* the same kernel over the centre table (probe selection of query batches) */
int			vb_prof_enable(int on);
/* Synchronises, then returns accumulated milliseconds and bracketed launches since the last read of `kernel`. */
int			vb_prof_read(int kernel, double *total_ms, int64_t *launches);

/* ------------------------------------------------- batched distance operator */

/*
 * One query against n rows, all host buffers: out[i] = metric(rows[i], q) as the
 * float8 the fmgr wrapper returns.  Replaces n calls of
 * FunctionCall2Coll(procinfo, collation, row, q) -> l2_distance / ... / jaccard_distance
 * (src/vector.c:576-750, src/halfvec.c:557-686, src/bitvec.c:33-70).
 * dim is elements (bits for VB_BIT).  q == NULL gives all zeros (ZeroDistance, src/ivfscan.c:192-196).
 */
int			vb_distance_batch(int elem, int metric, int dim, const void *q,
							  const void *rows, int64_t n, double *out);

/*
 * Batched row transforms next to the distance path (host buffers in and out):
 *   vb_norm_batch            vector_norm / l2_norm           (src/vector.c:767-780, src/halfvec.c:703-720)
 *   vb_l2_normalize_batch    l2_normalize                    (src/vector.c:785-819, src/halfvec.c:725-759); what the cosine
 *                            opclasses apply to every indexed row and to the query (src/ivfbuild.c:174-180, src/ivfscan.c:222-229);
 *                            fails with "value out of range: overflow" like float_overflow_error()
 *   vb_binary_quantize_batch binary_quantize                 (src/vector.c:952-978): out = (dim + 7) / 8 bytes per row, MSB first
 * elem = VB_VECTOR or VB_HALFVEC; norms accumulate in fp64 as in the reference.
 */
int			vb_norm_batch(int elem, int dim, const void *rows, int64_t n, double *out);
int			vb_l2_normalize_batch(int elem, int dim, const void *rows, int64_t n, void *out);
int			vb_binary_quantize_batch(int elem, int dim, const void *rows, int64_t n, uint8_t *out);
/*
 * The casts that feed halfvec / bit indexes from vector columns (README "half-precision indexing"):
 *   vb_vector_to_halfvec_batch  vector_to_halfvec (src/halfvec.c:540-555): Float4ToHalf, round to nearest even; a finite
 *                               value that overflows fails with the reference's text: "<value>" is out of range for type halfvec
 *   vb_halfvec_to_vector_batch  halfvec_to_vector (src/vector.c halfvec_to_vector): exact widening
 */
int			vb_vector_to_halfvec_batch(int dim, const void *rows, int64_t n, void *out);
int			vb_halfvec_to_vector_batch(int dim, const void *rows, int64_t n, void *out);

/* ------------------------------------------------------ resident row tables */

typedef struct vb_table vb_table;	/* [n x dim] rows resident in HBM (exact scan, HNSW vectors, k-means samples) */

int			vb_table_create(int elem, int dim, vb_table **out);
/* Append n rows from host memory (pinned staging + async copy inside). */
int			vb_table_append(vb_table *t, const void *rows, int64_t n);
/* Append n rows that already live on the device (packed, unpadded layout). */
int			vb_table_append_dev(vb_table *t, const void *rows_dev, int64_t n);
int64_t		vb_table_rows(const vb_table *t);
/* Read-only view of the resident rows: device pointer of row 0 and the padded row stride in bytes. */
const void *vb_table_device_rows(const vb_table *t, size_t *stride_bytes);
int			vb_table_free(vb_table *t);

/*
 * Exact (no index) top-k of nq queries over the table: the sequential-scan plan
 * "ORDER BY v <op> q LIMIT k" (operator wrapper + top-N sort).
 * out_ids[q*k + j] is the row number (0-based append order), -1 padded;
 * out_dist the operator's float8.  Ties on distance: smaller row number first.
 */
int			vb_exact_topk(vb_table *t, int metric, const void *queries, int64_t nq, int k,
						  int64_t *out_ids, double *out_dist);
int			vb_exact_topk_dev(vb_table *t, int metric, const void *queries_dev, int64_t nq, int k,
							  int64_t *out_ids_dev, float *out_dist_dev);

/* ---------------------------------------------------------------- sparsevec */

/*
 * sparsevec (src/sparsevec.h:21-32): dim, nnz, indices[nnz] (0-based, ascending), values[nnz].  A batch of rows is CSR:
 * row r = entries row_off[r] .. row_off[r+1] of idx[] / val[] (row_off[0] = 0).  All host buffers.
 *
 *   vb_sparsevec_distance_batch      out[r] = metric(row r, q) as the float8 of sparsevec's l2_distance /
 *                                    l2_squared_distance / inner_product / negative_inner_product / cosine_distance /
 *                                    l1_distance (src/sparsevec.c:826-1057); metric = VB_L2, VB_L2_SQUARED, VB_IP, VB_NEG_IP,
 *                                    VB_COSINE, VB_L1.  dim != q_dim fails with CheckDims' text ("different sparsevec
 *                                    dimensions %d and %d", src/sparsevec.c:44-51); q_nnz < 0 = NULL query, all zeros.
 *   vb_sparsevec_norm_batch          l2_norm (src/sparsevec.c:1062-1077), fp64 sums
 *   vb_sparsevec_l2_normalize_batch  l2_normalize (src/sparsevec.c:1082-1150): quotients that round to zero are dropped, so
 *                                    the result has its own offsets; out_idx / out_val need room for row_off[n] entries;
 *                                    an infinite quotient fails with "value out of range: overflow"
 */
int			vb_sparsevec_distance_batch(int metric, int dim, int q_dim, int32_t q_nnz, const int32_t *q_idx, const float *q_val,
										int64_t n, const int64_t *row_off, const int32_t *idx, const float *val, double *out);
int			vb_sparsevec_norm_batch(int64_t n, const int64_t *row_off, const float *val, double *out);
int			vb_sparsevec_l2_normalize_batch(int64_t n, const int64_t *row_off, const int32_t *idx, const float *val,
											int64_t *out_row_off, int32_t *out_idx, float *out_val);

typedef struct vb_sparse_table vb_sparse_table;	/* n sparsevec rows resident in HBM as CSR */

int			vb_sparse_table_create(int dim, vb_sparse_table **out);
int			vb_sparse_table_append(vb_sparse_table *t, int64_t n, const int64_t *row_off, const int32_t *idx, const float *val);
int64_t		vb_sparse_table_rows(const vb_sparse_table *t);
int64_t		vb_sparse_table_nnz(const vb_sparse_table *t);
int			vb_sparse_table_free(vb_sparse_table *t);
/*
 * Exact (no index) top-k of nq sparsevec queries (CSR: q_off[nq+1], q_idx, q_val) over the table: the sequential-scan plan
 * "ORDER BY v <op> q LIMIT k" for <-> (VB_L2), <#> (VB_NEG_IP), <=> (VB_COSINE), <+> (VB_L1); k <= 2048.
 * out_ids = row numbers (append order, -1 padded), out_dist = the operator's float8; ties: smaller row number first.
 */
int			vb_sparse_exact_topk(vb_sparse_table *t, int metric, int q_dim, int64_t nq, const int64_t *q_off, const int32_t *q_idx,
								 const float *q_val, int k, int64_t *out_ids, double *out_dist);

/* ---------------------------------------------------------------- IVFFlat */

typedef struct vb_ivf vb_ivf;	/* device image of one ivfflat index: centres + rows grouped by list + ids */

/* metric = the opclass's proc 1: VB_L2_SQUARED, VB_NEG_IP (ip and cosine opclasses) or VB_HAMMING. */
int			vb_ivf_create(int elem, int metric, int dim, int lists, vb_ivf **out);
/*
 * Load the whole image from host memory: centres [lists], rows grouped by
 * list with list_offsets [lists+1] (row index prefix), ids [n] (heap TIDs).
 * This is what the packer produces from the list pages / entry pages
 * (src/ivfflat.h:251-277; readers src/ivfscan.c:62-107, 139-179).
 */
int			vb_ivf_load(vb_ivf *ix, const void *centers, const int64_t *list_offsets,
						const void *rows, const int64_t *ids);
/* Same, from device buffers (packed rows; ids may be NULL = row positions). */
int			vb_ivf_load_dev(vb_ivf *ix, const void *centers_dev, const int64_t *list_offsets_host,
							const void *rows_dev, const int64_t *ids_dev);
int64_t		vb_ivf_rows(const vb_ivf *ix);
/*
 * The same image, one list at a time -- the granularity the packer reads at (one entry-page chain per list,
 * src/ivfscan.c:139-179), so the host never stages more than one list: vb_ivf_begin_load (centres), vb_ivf_load_list
 * for the non-empty lists in ascending list order, vb_ivf_end_load.
 * vb_ivf_replace_list swaps one list of a loaded image for new contents (an insert into that list, a vacuum of it):
 * only that list crosses PCIe, the rows behind it are moved on the device, and the packed planes of the tensor-core
 * filter are rebuilt on the device by the next batched scan.
 */
int			vb_ivf_begin_load(vb_ivf *ix, const void *centers);
int			vb_ivf_load_list(vb_ivf *ix, int list, const void *rows, const int64_t *ids, int64_t n);
int			vb_ivf_end_load(vb_ivf *ix);
int			vb_ivf_replace_list(vb_ivf *ix, int list, const void *rows, const int64_t *ids, int64_t n);
int			vb_ivf_free(vb_ivf *ix);

/*
 * GetScanLists (src/ivfscan.c:47-118): distance from each query to every
 * centre, nearest max_probes lists, ascending.  Ties: smaller list number first.
 * out_lists [nq x max_probes] (int32), out_dist [nq x max_probes] (may be NULL).
 */
int			vb_ivf_scan_lists(vb_ivf *ix, const void *queries, int64_t nq, int max_probes,
							  int32_t *out_lists, double *out_dist);
/*
 * GetScanItems (src/ivfscan.c:123-187) for ONE query: distance to every row of
 * the given lists, fully sorted ascending (tuplesort_performsort, :182).  Writes
 * at most cap results, *n_out = number of candidates scanned.  Ties: scan order.
 * q == NULL: all distances 0 (src/ivfscan.c:207-211).
 */
int			vb_ivf_scan_items(vb_ivf *ix, const void *q, const int32_t *lists, int nlists,
							  int64_t cap, int64_t *out_ids, double *out_dist, int64_t *n_out);
/*
 * The whole first batch of ivfflatgettuple (src/ivfscan.c:360-414) for nq
 * queries at once: probe selection + list scan + top-k (k nearest of the
 * probed lists; k <= 0 is rejected here, use vb_ivf_scan_items for "all").
 * Host buffers; copies are inside the call.
 */
int			vb_ivf_search(vb_ivf *ix, const void *queries, int64_t nq, int probes, int k,
						  int64_t *out_ids, double *out_dist);
/* Same with device-resident queries and outputs (float distances), asynchronous on vb_stream(). */
int			vb_ivf_search_dev(vb_ivf *ix, const void *queries_dev, int64_t nq, int probes, int k,
							  int64_t *out_ids_dev, float *out_dist_dev);
/*
 * Pipelined host path.  vb_ivf_prefetch_queries starts the host->device copy of the NEXT batch of queries on a second
 * stream (slot 0 or 1; `queries` should be page-locked and must stay valid until the matching search returns) and
 * returns at once; vb_ivf_search_prefetched runs the search of a slot filled earlier (same results as vb_ivf_search)
 * and returns when its results are in `out_ids` / `out_dist`.  Alternating the two slots overlaps every copy with the
 * previous batch's compute.  vector queries with dim % 4 == 0 only; VB_EINVAL otherwise.
 */
int			vb_ivf_prefetch_queries(vb_ivf *ix, const void *queries, int64_t nq, int slot);
int			vb_ivf_search_prefetched(vb_ivf *ix, int slot, int probes, int k, int64_t *out_ids, double *out_dist);
/* algorithmic bytes of the last vb_ivf_search*: sum over queries of (lists + candidates) * dim * elem size */
int64_t		vb_ivf_last_scan_bytes(const vb_ivf *ix);
int64_t		vb_ivf_last_candidates(const vb_ivf *ix);
/* Queries (cumulative) whose tensor-core filter result could not be certified against the error bound and were
 * re-run on the exact fp32 kernel (option "scan_impl" = 4); 0 when that path is not in use. */
int64_t		vb_ivf_tc_fallbacks(const vb_ivf *ix);
/* Queries (cumulative) the first filter level (hi plane of the rows only, option "tc_level1") could not certify; their
 * batches were repeated with both planes.  After such a batch the level rests for 64 batches. */
int64_t		vb_ivf_tc_level1_fallbacks(const vb_ivf *ix);

/*
 * Traffic accounting of the tensor-core filter kernel (profiling, off by default): with on != 0 every launch also
 * sums, from the job list the kernel walks, the bytes its bulk copies request and the distinct bytes among them.
 * out8 (may be NULL): read and reset the counters -- list scan [0] bytes requested, [1] distinct row-plane bytes,
 * [2] distinct query-tile bytes, [3] launches; [4..7] the same for the centre scan (probe selection).
 */
int			vb_ivf_tc_traffic(int on, int64_t *out8);

/*
 * List-sharded search over the library's communicator (vb_comm_init): this rank's image holds its own lists under
 * the GLOBAL list numbering (every other list empty) and all centres.  Probe selection runs on this rank's slice of
 * the queries, the probe lists are all-gathered, every rank scans its lists for all queries, the per-rank k nearest
 * are all-gathered and merged by (distance, id).  All ranks call it with the same queries and get the full result.
 * Device pointers; asynchronous on vb_stream() apart from one read of the filter's certificate counters.
 */
int			vb_ivf_search_sharded_dev(vb_ivf *ix, const void *queries_dev, int64_t nq, int probes, int k,
									  int64_t *out_ids_dev, float *out_dist_dev);
/* The same with host buffers (queries in, int64 ids + float8 distances out; copies inside, synchronous). */
int			vb_ivf_search_sharded(vb_ivf *ix, const void *queries, int64_t nq, int probes, int k,
								  int64_t *out_ids, double *out_dist);

/*
 * Exact (no index) top-k over a row-sharded table: every rank scans its own rows (vb_exact_topk_dev), the per-rank k
 * nearest are all-gathered and merged by (distance, id).  id_offset = the global number of this rank's first row.
 * Collective; every rank gets the full result.
 */
int			vb_exact_topk_sharded_dev(vb_table *t, int metric, const void *queries_dev, int64_t nq, int k, int64_t id_offset,
									  int64_t *out_ids_dev, float *out_dist_dev);

/* --------------------------------------------------------------- communicator */

/*
 * One process per GPU.  vb_comm_unique_id (on one rank) fills the 128-byte NCCL id the host passes to the others
 * (the extension: through its DSM segment, like the reference's parallel-build state, src/ivfbuild.c:830-966);
 * vb_comm_init (every rank, collectively) creates the communicator on the library's device.  While it exists,
 * vb_kmeans / vb_kmeans_pp_init treat `samples` as this rank's slice of a row-sharded sample set (centre sums,
 * counts, the change counter, k-means++ weight sums and chosen rows are exchanged with ncclAllReduce /
 * ncclAllGather on vb_stream()), and vb_ivf_search_sharded_dev is available.  dtype: 0 fp32, 1 int32, 2 int64,
 * 3 fp64, 4 uint32.
 */
int			vb_comm_unique_id(void *out, size_t cap);
int			vb_comm_init(const void *id_bytes, int rank, int world);
int			vb_comm_free(void);
int			vb_comm_world(void);
int			vb_comm_rank(void);
int			vb_comm_allreduce(void *buf_dev, int64_t count, int dtype);
int			vb_comm_allgather(const void *send_dev, void *recv_dev, int64_t bytes_per_rank);

/* ------------------------------------------------------- IVFFlat build path */

/*
 * Collective hook for the sharded build: sum-reduce `count` elements of the
 * given device buffer in place across all ranks (dtype 0 = fp32, 1 = int32,
 * 2 = int64).  The extension passes an ncclAllReduce wrapper; tests pass a
 * torch.distributed one.  NULL = single process.
 */
typedef int (*vb_allreduce_fn) (void *buf_dev, int64_t count, int dtype, void *ctx);

/*
 * k-means on samples resident in `samples` (this rank's shard), replacing
 * ElkanKmeans (src/ivfkmeans.c:246-485) from given initial centres
 * (centres = in/out host buffer of k rows; k-means++ draws are host-side,
 * see vb_kmeans_pp_init).  kmeans_metric: VB_L2 (l2 opclasses), VB_SPHERICAL
 * (ip / cosine opclasses; samples must already be unit vectors,
 * src/ivfbuild.c:153-155) or VB_HAMMING (bit).  Lloyd iterations with a dense
 * assign step; centre update and stopping rule as src/ivfkmeans.c:179-236,
 * 482-483.  *iters_out = iterations executed.
 */
int			vb_kmeans(vb_table *samples, int kmeans_metric, void *centers, int k, int max_iter,
					  uint64_t seed, vb_allreduce_fn allreduce, void *allreduce_ctx, int *iters_out);
/* InitCenters (src/ivfkmeans.c:23-91): k-means++ seeding on the device, centres out (host). */
int			vb_kmeans_pp_init(vb_table *samples, int kmeans_metric, void *centers, int k, uint64_t seed);
/*
 * Same with the caller's random draws (the extension passes pg_prng's, the parity tests the oracle's): the first
 * centre is sample first_row (RandomInt() % numSamples, src/ivfkmeans.c:36); u[i], i < k - 1, is the RandomDouble()
 * of round i (src/ivfkmeans.c:78).  picked_out (may be NULL): the k chosen sample rows.
 */
/*
 * On large fp32 sample tables the seeding touches a sample only when its weight could change (triangle inequality over
 * the chosen centres, then a bf16 lower bound; exact fp32 re-score for the rest -- the weights and the picks are
 * those of the full pass).  out3: samples skipped by the triangle rule / stopped by the bf16 bound / re-scored
 * exactly during the last seeding of this process (all zero when the plain pass ran).  Option "pp_filter": 0 = never,
 * 1 = automatic (default), 2 = always.
 */
int			vb_kmeans_pp_stats(int64_t *out3);
int			vb_kmeans_pp_init_draws(vb_table *samples, int kmeans_metric, void *centers, int k, int64_t first_row,
									const double *u, int64_t *picked_out);
/*
 * AddTupleToSort's argmin (src/ivfbuild.c:161-219): out_list[i] = first centre
 * minimising the proc-1 distance (strict <).  metric = VB_L2_SQUARED / VB_NEG_IP / VB_HAMMING.
 */
int			vb_assign(vb_table *rows, int metric, const void *centers, int k, int32_t *out_list);
int			vb_assign_dev(vb_table *rows, int metric, const void *centers_dev, int k, int32_t *out_list_dev);
/*
 * The assign step runs on the tensor cores (tcgen05, split-bf16 GEMM with a fused row argmin)
 * and re-checks rows whose best/second-best margin is inside the error bound with the exact
 * fp32 kernel.  vb_set_tensor_cores(0) forces the exact CUDA-core kernel for everything
 * (used by the parity tests); vb_last_assign_rechecked() = rows the last assign re-checked
 * (-1 when the exact kernel did all the work).
 */
int			vb_set_tensor_cores(int on);
int64_t		vb_last_assign_rechecked(void);
/*
 * Tuning switches.  "scan_impl" selects the list / table scan: 0 = per-query LDG.128 streaming kernel,
 * 1 = per-query cp.async.bulk (TMA) + mbarrier staged kernel for rows of at least 512 bytes, 2 (default) =
 * automatic (query batches are scanned list-major: each probed list read once per batch, fp32x2 register
 * tiles; single queries stream), 3 = list-major wherever it applies, 4 = tensor-core filter (split-bf16
 * tcgen05 distances, exact fp32 re-score of k' candidates, certificate, exact fallback) wherever it applies.
 * Every setting returns the same neighbours.  "tc_level1" (default 1): the tensor-core filter first reads only the
 * hi plane of the rows (half the HBM traffic, 2^-7 relative error bound) and repeats a batch with both planes when a
 * certificate fails.  "tensor_cores" as vb_set_tensor_cores.  "one_query" (default 1): calls with at most 16 queries --
 * one backend's scan: vb_ivf_scan_lists, vb_ivf_scan_items, vb_ivf_search -- run as two fused distance + select
 * kernels (the last CTA to finish selects; csrc/vb_ivf_one.cu) instead of the general launch sequence; 0 = general path.
 */
int			vb_set_option(const char *name, int64_t value);

/* -------------------------------------------------------------------- HNSW */

typedef struct vb_hnsw vb_hnsw;	/* device image of one hnsw index: element vectors + neighbour tables */

/*
 * metric = opclass proc 1 (VB_L2_SQUARED, VB_NEG_IP, VB_L1, VB_HAMMING, VB_JACCARD).
 * Elements are numbered 0..n-1; levels[n]; nbr0 [n x 2m] layer-0 neighbour
 * element numbers in on-disk order (HnswSetNeighborTuple, src/hnswutils.c:455-486),
 * -1 terminated; upper layers as upper_off[n] (-1 when level 0) and
 * upper [slots x m] with layer lc of element e at slot upper_off[e] + lc - 1.
 * entry = entry point element (meta page, src/hnswutils.c:298-328).
 */
int			vb_hnsw_create(int elem, int metric, int dim, int m, vb_hnsw **out);
int			vb_hnsw_load(vb_hnsw *h, const void *rows, int64_t n, const int32_t *levels,
						 const int32_t *nbr0, const int64_t *upper_off, const int32_t *upper,
						 int64_t upper_slots, int64_t entry);
int			vb_hnsw_free(vb_hnsw *h);
/*
 * CREATE INDEX ... USING hnsw on the device: the in-memory build of src/hnswbuild.c:437-480 --
 * HnswFindElementNeighbors (src/hnswutils.c:1280-1357) with ef_construction, the SelectNeighbors heuristic
 * (:1065-1165), duplicate folding (src/hnswbuild.c:343-364) and HnswUpdateConnection (:1184-1231) -- for rows
 * inserted in batches the way the reference's parallel workers insert concurrently.  Row i becomes element i.
 * levels: the per-row level draws of HnswInitElement (src/hnswutils.c:248-254) when the caller owns the PRNG
 * (the extension passes pg_prng's), NULL = drawn from `seed`.  The index is searchable afterwards (vb_hnsw_search)
 * and vb_hnsw_export returns the graph in the layout vb_hnsw_load takes, for the page writer
 * (HnswSetNeighborTuple, src/hnswutils.c:455-486): levels [n], nbr0 [n x 2m], upper_off [n], upper
 * [vb_hnsw_upper_slots() x m], entry point, and dup_of [n] = the element a duplicate row was folded into (its heap
 * TID joins that element's, HNSW_HEAPTIDS = 10 at most, src/hnsw.h:69) or -1.  Any output may be NULL.
 */
int			vb_hnsw_build(vb_hnsw *h, const void *rows, int64_t n, int ef_construction, uint64_t seed,
						  const int32_t *levels);
int			vb_hnsw_build_dev(vb_hnsw *h, const void *rows_dev, int64_t n, int ef_construction, uint64_t seed,
							  const int32_t *levels);
int64_t		vb_hnsw_rows(const vb_hnsw *h);
int64_t		vb_hnsw_upper_slots(const vb_hnsw *h);
int			vb_hnsw_export(vb_hnsw *h, int32_t *levels, int32_t *nbr0, int64_t *upper_off, int32_t *upper,
						   int64_t *entry, int32_t *dup_of);
/*
 * GetScanItems (src/hnswscan.c:25-56): greedy descent with ef = 1 through the
 * upper layers, then HnswSearchLayer (src/hnswutils.c:824-987) with ef at layer 0;
 * results nearest first (src/hnswscan.c:293-326), k <= ef of them per query,
 * -1 padded.  Every distance comparison is on the total order (distance,
 * element number).  out_ndist (may be NULL) = distance evaluations per query
 * (the reference's `tuples` counter, src/hnswutils.c:872-873, 905-906).
 */
int			vb_hnsw_search(vb_hnsw *h, const void *queries, int64_t nq, int ef, int k,
						   int64_t *out_ids, double *out_dist, int64_t *out_ndist);
int			vb_hnsw_search_dev(vb_hnsw *h, const void *queries_dev, int64_t nq, int ef, int k,
							   int64_t *out_ids_dev, float *out_dist_dev, int64_t *out_ndist_dev);

/*
 * hnsw.iterative_scan (src/hnswscan.c:62-87 ResumeScanItems, :228-340 hnswgettuple) for nq queries at once.  The
 * handle owns what the reference keeps in HnswScanOpaqueData between batches: the visited set `v`, the `discarded`
 * candidates (rejected neighbours and evicted results, src/hnswutils.c:929-937, 968-973) and the `tuples` counter.
 * vb_hnsw_scan_next returns, per query, the next batch nearest first: out_ids / out_distances [nq x ef_search]
 * (-1 / +inf padded), out_counts [nq]; the first call is GetScanItems (:25-56), every later one resumes from the
 * ef_search nearest discarded candidates on the same visited set, and once a query's tuples counter has reached
 * max_scan_tuples (hnsw.max_scan_tuples, src/hnsw.c:101-105) its remaining discarded candidates come back nearest
 * first, ef_search at a time, without searching (:247-254).  out_counts[q] == 0: that scan is exhausted.
 * The sequence is hnsw.iterative_scan = relaxed_order's; strict_order is the caller's filter on it (:316-322).
 * work_mem * hnsw.scan_mem_multiplier (:247) is not modelled: map it onto max_scan_tuples.
 */
typedef struct vb_hnsw_scan vb_hnsw_scan;
int			vb_hnsw_scan_begin(vb_hnsw *h, const void *queries, int64_t nq, int ef_search, int64_t max_scan_tuples,
							   vb_hnsw_scan **out);
int			vb_hnsw_scan_next(vb_hnsw_scan *scan, int64_t *out_ids, double *out_distances, int32_t *out_counts);
int			vb_hnsw_scan_tuples(vb_hnsw_scan *scan, int64_t *out_tuples);	/* [nq] the tuples counters */
int			vb_hnsw_scan_end(vb_hnsw_scan *scan);

#ifdef __cplusplus
}
#endif
#endif							/* VECB200_H */
