/*
 * pgv_oracle.h -- CPU oracle for the pgvector distance hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker or as the
 * timed CPU baseline.  The product (libvecb200.so) never links or calls it.
 *
 * Every function is a restatement (not a copy) of the reference algorithm
 * and cites the reference file:line it follows (paths relative to the
 * pgvector tree, reference @ e48241b).
 *
 * Parity status: distance arithmetic is pinned by the reference's own
 * regression outputs (tests/golden/ *.json, transcribed from
 * test/expected/{vector_type,halfvec,bit}.out) and by the reference's
 * halfutils.c / bitutils.c compiled verbatim into oracle/_ref/.  Tie order
 * inside PostgreSQL's pairing heap / tuplesort and the PRNG streams are
 * PostgreSQL-core behaviour that is not under /root/reference: for those
 * this oracle restates the published algorithm (lib/pairingheap.c) and
 * parity on tie order / random draws is UNPINNED.
 */
#ifndef PGV_ORACLE_H
#define PGV_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types (src/vector.h:18-24, src/halfvec.h:67-73, PG VarBit) */
enum { PGV_VECTOR = 0, PGV_HALFVEC = 1, PGV_BIT = 2 };

/* metrics; values are shared with include/vecb200.h (a test checks it) */
enum {
	PGV_L2_SQUARED = 0,   /* vector_l2_squared_distance   src/vector.c:595-605  (index proc 1 of the l2 opclasses) */
	PGV_NEG_IP = 1,       /* vector_negative_inner_product src/vector.c:637-647 (index proc 1 of the ip and cosine opclasses, <#>) */
	PGV_COSINE = 2,       /* cosine_distance              src/vector.c:671-696  (<=> seq-scan operator) */
	PGV_L1 = 3,           /* l1_distance                  src/vector.c:740-750  (<+>) */
	PGV_HAMMING = 4,      /* hamming_distance             src/bitvec.c:45-55    (<~>) */
	PGV_JACCARD = 5,      /* jaccard_distance             src/bitvec.c:60-70    (<%>) */
	PGV_L2 = 6,           /* l2_distance                  src/vector.c:579-589  (<->) */
	PGV_IP = 7,           /* inner_product                src/vector.c:622-632 */
	PGV_SPHERICAL = 8     /* vector_spherical_distance    src/vector.c:703-722  (k-means proc 3 for ip/cosine) */
};

/* ---- pgv_distance.c ---------------------------------------------------- */
uint16_t pgv_float_to_half(float f);           /* Float4ToHalfUnchecked  src/halfutils.h:146-239 */
float    pgv_half_to_float(uint16_t h);        /* HalfToFloat4           src/halfutils.h:62-141 */

/* SQL-visible distance of one pair.  dim is elements (bits for PGV_BIT). */
double pgv_distance(int elem, int metric, int dim, const void *a, const void *b);
/* same value computed with fp64 accumulation ("truth" for tolerances) */
double pgv_distance_f64(int elem, int metric, int dim, const void *a, const void *b);
/* vector_norm / halfvec l2_norm  (src/vector.c:767-780, src/halfvec.c:703-720) */
double pgv_norm(int elem, int dim, const void *a);
/* l2_normalize / halfvec_l2_normalize (src/vector.c:785-819, src/halfvec.c:725-759); returns 0, or -1 on overflow */
int    pgv_l2_normalize(int elem, int dim, const void *a, void *out);
/* binary_quantize (src/vector.c:952-978): out has (dim+7)/8 bytes, MSB first */
void   pgv_binary_quantize(int elem, int dim, const void *a, uint8_t *out);
size_t pgv_row_bytes(int elem, int dim);

/* batched one-vs-many helper used by the brute-force leg of the tests */
void pgv_distance_batch(int elem, int metric, int dim, const void *q,
						const void *rows, int64_t n, double *out);

/* exact (no index) top-k: seq scan + sort. ties: smaller row id first. */
void pgv_exact_topk(int elem, int metric, int dim, const void *q, const void *rows,
					int64_t n, int k, int64_t *out_ids, double *out_dist);

/* ---- pgv_sparse.c ------------------------------------------------------ */
/* sparsevec = (dim, nnz, indices[nnz] ascending 0-based, values[nnz]) (src/sparsevec.h:21-32).  The float8 the SQL
 * function returns (src/sparsevec.c:826-1057); metrics: L2_SQUARED, L2, IP, NEG_IP, COSINE, L1 */
double pgv_sparse_distance(int metric, int a_nnz, const int32_t *a_idx, const float *a_val,
						   int b_nnz, const int32_t *b_idx, const float *b_val);
double pgv_sparse_distance_f64(int metric, int a_nnz, const int32_t *a_idx, const float *a_val,
							   int b_nnz, const int32_t *b_idx, const float *b_val);
double pgv_sparse_l2_norm(int nnz, const float *val);	/* src/sparsevec.c:1062-1077 */
/* src/sparsevec.c:1082-1150; returns the result's nnz, or -1 for float_overflow_error() */
int    pgv_sparse_l2_normalize(int nnz, const int32_t *idx, const float *val, int32_t *out_idx, float *out_val);
/* one query against n CSR rows: out[r] = distance(row r, q) */
void   pgv_sparse_distance_batch(int metric, int q_nnz, const int32_t *q_idx, const float *q_val, int64_t n,
								 const int64_t *row_off, const int32_t *idx, const float *val, double *out);

/* ---- pgv_ivfflat.c ----------------------------------------------------- */
typedef struct PgvIvfIndex
{
	int			elem;
	int			metric;			/* proc 1 metric: PGV_L2_SQUARED / PGV_NEG_IP / PGV_HAMMING */
	int			dim;
	int			lists;
	const void *centers;		/* lists rows */
	const int64_t *list_offsets;	/* lists+1 prefix offsets into rows/ids */
	const void *rows;			/* rows grouped by list */
	const int64_t *ids;			/* opaque row ids (heap TIDs) grouped by list */
}			PgvIvfIndex;

/* tie rule of the list selection: 0 = PostgreSQL pairing heap (reference), 1 = (distance, list number) */
void pgv_ivf_set_tie_mode(int total_order);
/* GetScanLists (src/ivfscan.c:47-118): out_lists[maxProbes] nearest first */
int pgv_ivf_scan_lists(const PgvIvfIndex *ix, const void *q, int max_probes,
					   int *out_lists, double *out_dist);
/*
 * GetScanItems + full sort (src/ivfscan.c:123-187): scans probe lists
 * [first_probe, first_probe+probes) of `lists`, returns count; output sorted
 * ascending by distance, ties in scan order (stable).  q == NULL => all 0.
 */
int64_t pgv_ivf_scan_items(const PgvIvfIndex *ix, const void *q, const int *lists,
						   int probes, int64_t cap, int64_t *out_ids, double *out_dist);
/* whole scan for one query: lists + items, top `k` (k<=0: all) */
int64_t pgv_ivf_search(const PgvIvfIndex *ix, const void *q, int probes, int k,
					   int64_t *out_ids, double *out_dist);
/* many queries on `threads` threads (models N concurrent backends) */
void pgv_ivf_search_batch(const PgvIvfIndex *ix, const void *queries, int64_t nq,
						  int probes, int k, int threads, int64_t *out_ids, double *out_dist);

/* AddTupleToSort's argmin (src/ivfbuild.c:161-219): strict <, first min wins */
void pgv_ivf_assign(int elem, int metric, int dim, const void *rows, int64_t n,
					const void *centers, int lists, int threads, int32_t *out_list);

/*
 * ElkanKmeans (src/ivfkmeans.c:246-485) from caller-supplied initial centres
 * (k-means++ draws are PRNG-driven and therefore shared, not reproduced).
 * kmeans_metric: PGV_L2 (l2_ops), PGV_SPHERICAL (ip/cosine ops; samples must be unit
 * vectors, centres renormalised), PGV_HAMMING (bit).  centers is in/out.
 * Returns iterations run.  closest (n) optional out.  seed drives the empty-cluster reseed.
 */
int pgv_kmeans_elkan(int elem, int kmeans_metric, int dim, const void *samples, int64_t n,
					 void *centers, int k, int max_iter, uint64_t seed, int32_t *closest);
/* plain Lloyd with identical centre update rules, same stopping rule */
int pgv_kmeans_lloyd(int elem, int kmeans_metric, int dim, const void *samples, int64_t n,
					 void *centers, int k, int max_iter, uint64_t seed, int32_t *closest);
/* InitCenters (src/ivfkmeans.c:23-91) with a splitmix/xoroshiro stream in place of pg_prng */
void pgv_kmeans_pp_init(int elem, int kmeans_metric, int dim, const void *samples, int64_t n,
						void *centers, int k, uint64_t seed);

/* same with caller-supplied draws: first row, u[k-1] uniforms; picked[k] optional */
void pgv_kmeans_pp_init_draws(int elem, int kmeans_metric, int dim, const void *samples, int64_t n,
							  void *centers, int k, int64_t first, const double *u, int64_t *picked);

/* ---- pgv_hnsw.c -------------------------------------------------------- */
typedef struct PgvHnsw PgvHnsw;

/* tie mode for the two search heaps */
enum { PGV_TIES_PG_PAIRINGHEAP = 0, PGV_TIES_TOTAL_ORDER = 1 };

PgvHnsw *pgv_hnsw_create(int elem, int metric, int dim, int m, int ef_construction, uint64_t seed);
void	 pgv_hnsw_free(PgvHnsw *g);
/* in-memory build insert (src/hnswbuild.c:437-480, hnswutils.c:1280-1357). rows must stay alive. */
void	 pgv_hnsw_build(PgvHnsw *g, const void *rows, int64_t n);
void	 pgv_hnsw_build_levels(PgvHnsw *g, const void *rows, int64_t n, const int32_t *levels);
int64_t  pgv_hnsw_count(const PgvHnsw *g);
int		 pgv_hnsw_entry(const PgvHnsw *g, int64_t *entry, int *entry_level);
/* export: levels[n]; layer-0 neighbours [n][2m] (-1 padded) */
void	 pgv_hnsw_export_layer0(const PgvHnsw *g, int32_t *levels, int32_t *nbr0);
/* upper layers: for element e with level L>=1, neighbours of layer lc in 1..L at
 * upper[(upper_off[e] + (lc-1)) * m .. +m) (-1 padded); upper_off[e] = -1 when level 0.
 * returns number of (element,layer) slots; pass NULL to size. */
int64_t  pgv_hnsw_export_upper(const PgvHnsw *g, int64_t *upper_off, int32_t *upper);
/* per element: first row index, number of heap tids (<=10, duplicates share an element,
 * src/hnswbuild.c:343-364), heap tids [n][10] (-1 padded) */
void	 pgv_hnsw_export_elements(const PgvHnsw *g, int64_t *elem_row, int32_t *n_heaptids, int64_t *heaptids);
/* import a graph exported above (so tests can share one graph between runs) */
PgvHnsw *pgv_hnsw_import(int elem, int metric, int dim, int m, const void *rows, int64_t n,
						 const int32_t *levels, const int32_t *nbr0,
						 const int64_t *upper_off, const int32_t *upper,
						 int64_t entry, int entry_level);
/* GetScanItems (src/hnswscan.c:25-56) + nearest-first drain (hnswscan.c:293-326).
 * returns number of results (<= ef); n_dist = the `tuples` counter (hnswutils.c:872,905). */
int pgv_hnsw_search(const PgvHnsw *g, const void *q, int ef, int tie_mode,
					int64_t *out_ids, double *out_dist, int64_t *n_dist);
/* iterative scan (src/hnswscan.c:62-87, 228-340; relaxed order): see pgv_hnsw.c */
int64_t pgv_hnsw_iter_scan(const PgvHnsw *g, const void *q, int ef, int tie_mode, int64_t max_scan_tuples, int64_t max_out,
						   int64_t *out_ids, double *out_dist, int32_t *out_batch, int64_t *tuples_out);
void pgv_hnsw_search_batch(const PgvHnsw *g, const void *queries, int64_t nq, int ef, int tie_mode,
						   int threads, int k, int64_t *out_ids, double *out_dist, int64_t *n_dist);

#ifdef __cplusplus
}
#endif
#endif
