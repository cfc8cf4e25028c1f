/*
 * mock_abi.c -- the subset of include/vecb200.h the extension glue calls, implemented on the CPU oracle.
 * TEST INFRASTRUCTURE (lives under tests/): lets the glue -- page packers, varlena handling, TID mapping, batching --
 * run end to end on a machine without a GPU, and records what the packers loaded so a test can compare it with the
 * arrays the pages were synthesised from.  The GPU tests link the same harness against the real libvecb200.so.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vecb200.h"
#include "../../oracle/pgv_oracle.h"

static char mock_err[256] = "";
const char *vb_last_error(void) { return mock_err; }
static int mock_fail_next_load = 0;
void mock_fail_next(int on) { mock_fail_next_load = on; }
static int mock_live_handles = 0;
int mock_live(void) { return mock_live_handles; }

struct vb_ivf
{
	PgvIvfIndex ix;
	void	   *centers, *rows;
	int64_t    *offsets, *ids;
	int64_t		n;
};

int
vb_ivf_create(int elem, int metric, int dim, int lists, vb_ivf **out)
{
	vb_ivf	   *h = calloc(1, sizeof(vb_ivf));

	h->ix.elem = elem;
	h->ix.metric = metric;
	h->ix.dim = dim;
	h->ix.lists = lists;
	*out = h;
	mock_live_handles++;
	return VB_OK;
}

int
vb_ivf_load(vb_ivf *h, const void *centers, const int64_t *list_offsets, const void *rows, const int64_t *ids)
{
	size_t		rb = pgv_row_bytes(h->ix.elem, h->ix.dim);
	int			lists = h->ix.lists;

	if (mock_fail_next_load)
	{
		mock_fail_next_load = 0;
		snprintf(mock_err, sizeof(mock_err), "mock: injected load failure");
		return VB_ENOMEM;
	}
	h->n = list_offsets[lists];
	h->centers = malloc(rb * (size_t) (lists ? lists : 1));
	memcpy(h->centers, centers, rb * (size_t) lists);
	h->offsets = malloc(sizeof(int64_t) * ((size_t) lists + 1));
	memcpy(h->offsets, list_offsets, sizeof(int64_t) * ((size_t) lists + 1));
	h->rows = malloc(rb * (size_t) (h->n ? h->n : 1));
	memcpy(h->rows, rows, rb * (size_t) h->n);
	h->ids = malloc(sizeof(int64_t) * (size_t) (h->n ? h->n : 1));
	memcpy(h->ids, ids, sizeof(int64_t) * (size_t) h->n);
	h->ix.centers = h->centers;
	h->ix.list_offsets = h->offsets;
	h->ix.rows = h->rows;
	h->ix.ids = h->ids;
	return VB_OK;
}

int
vb_ivf_free(vb_ivf *h)
{
	if (h)
	{
		free(h->centers);
		free(h->rows);
		free(h->offsets);
		free(h->ids);
		free(h);
		mock_live_handles--;
	}
	return VB_OK;
}

int64_t vb_ivf_rows(const vb_ivf *h) { return h ? h->n : 0; }

/* what the packer loaded, for direct comparison with the source arrays */
int64_t mock_ivf_rows(const vb_ivf *h, const void **centers, const int64_t **offsets, const void **rows, const int64_t **ids)
{
	*centers = h->centers;
	*offsets = h->offsets;
	*rows = h->rows;
	*ids = h->ids;
	return h->n;
}

int
vb_ivf_scan_lists(vb_ivf *h, const void *queries, int64_t nq, int max_probes, int32_t *out_lists, double *out_dist)
{
	size_t		rb = pgv_row_bytes(h->ix.elem, h->ix.dim);

	for (int64_t q = 0; q < nq; q++)
	{
		int		   *tmp = malloc(sizeof(int) * (size_t) max_probes);
		double	   *d = malloc(sizeof(double) * (size_t) max_probes);
		int			n = pgv_ivf_scan_lists(&h->ix, queries ? (const char *) queries + rb * (size_t) q : NULL, max_probes, tmp, d);

		for (int p = 0; p < max_probes; p++)
		{
			out_lists[q * max_probes + p] = p < n ? tmp[p] : -1;
			if (out_dist)
				out_dist[q * max_probes + p] = p < n ? d[p] : 1.0 / 0.0;
		}
		free(tmp);
		free(d);
	}
	return VB_OK;
}

int
vb_ivf_scan_items(vb_ivf *h, const void *q, const int32_t *lists, int nlists, int64_t cap, int64_t *out_ids, double *out_dist,
				  int64_t *n_out)
{
	int64_t		total = 0;

	for (int i = 0; i < nlists; i++)
		total += h->offsets[lists[i] + 1] - h->offsets[lists[i]];
	*n_out = total;
	if (cap > 0 && total > 0)
	{
		int64_t    *ids = malloc(sizeof(int64_t) * (size_t) total);
		double	   *d = malloc(sizeof(double) * (size_t) total);
		int64_t		k = cap < total ? cap : total;

		pgv_ivf_scan_items(&h->ix, q, (const int *) lists, nlists, total, ids, d);
		memcpy(out_ids, ids, sizeof(int64_t) * (size_t) k);
		memcpy(out_dist, d, sizeof(double) * (size_t) k);
		free(ids);
		free(d);
	}
	return VB_OK;
}

/* the first batch of ivfflatgettuple for nq queries: probes lists each, the k nearest (what vb_ivf_search returns) */
static int	mock_search_calls = 0;
int			mock_ivf_search_calls(void) { return mock_search_calls; }

int
vb_ivf_search(vb_ivf *h, const void *queries, int64_t nq, int probes, int k, int64_t *out_ids, double *out_dist)
{
	size_t		rb = pgv_row_bytes(h->ix.elem, h->ix.dim);

	__atomic_add_fetch(&mock_search_calls, 1, __ATOMIC_RELAXED);
	if (probes > h->ix.lists)
		probes = h->ix.lists;
	for (int64_t q = 0; q < nq; q++)
	{
		const char *qv = (const char *) queries + rb * (size_t) q;
		int32_t    *lists = malloc(sizeof(int32_t) * (size_t) probes);
		int64_t		n = 0;
		int			rc = vb_ivf_scan_lists(h, qv, 1, probes, lists, NULL);

		for (int i = 0; i < k; i++)
		{
			out_ids[q * k + i] = -1;
			out_dist[q * k + i] = 1.0 / 0.0;
		}
		if (rc == VB_OK)
			rc = vb_ivf_scan_items(h, qv, lists, probes, k, out_ids + q * k, out_dist + q * k, &n);
		free(lists);
		if (rc != VB_OK)
			return rc;
	}
	return VB_OK;
}

/* ---- hnsw ---- */
struct vb_hnsw
{
	int			elem, metric, dim, m;
	PgvHnsw    *g;
	void	   *rows;
	int32_t    *levels, *nbr0, *upper, *dup_of;
	int64_t    *upper_off;
	int64_t		n, entry, slots;
};

int
vb_hnsw_create(int elem, int metric, int dim, int m, vb_hnsw **out)
{
	vb_hnsw    *h = calloc(1, sizeof(vb_hnsw));

	h->elem = elem;
	h->metric = metric;
	h->dim = dim;
	h->m = m;
	*out = h;
	mock_live_handles++;
	return VB_OK;
}

int
vb_hnsw_load(vb_hnsw *h, const void *rows, int64_t n, const int32_t *levels, const int32_t *nbr0, const int64_t *upper_off,
			 const int32_t *upper, int64_t upper_slots, int64_t entry)
{
	size_t		rb = pgv_row_bytes(h->elem, h->dim);

	if (mock_fail_next_load)
	{
		mock_fail_next_load = 0;
		snprintf(mock_err, sizeof(mock_err), "mock: injected load failure");
		return VB_ENOMEM;
	}
	h->n = n;
	h->entry = n > 0 ? entry : -1;
	h->slots = upper_slots;
	h->rows = malloc(rb * (size_t) (n ? n : 1));
	memcpy(h->rows, rows, rb * (size_t) n);
	h->levels = malloc(sizeof(int32_t) * (size_t) (n ? n : 1));
	memcpy(h->levels, levels, sizeof(int32_t) * (size_t) n);
	h->nbr0 = malloc(sizeof(int32_t) * (size_t) (n ? n : 1) * 2 * h->m);
	memcpy(h->nbr0, nbr0, sizeof(int32_t) * (size_t) n * 2 * h->m);
	h->upper_off = malloc(sizeof(int64_t) * (size_t) (n ? n : 1));
	memcpy(h->upper_off, upper_off, sizeof(int64_t) * (size_t) n);
	h->upper = malloc(sizeof(int32_t) * (size_t) (upper_slots ? upper_slots : 1) * h->m);
	memcpy(h->upper, upper, sizeof(int32_t) * (size_t) upper_slots * h->m);
	if (n > 0)
		h->g = pgv_hnsw_import(h->elem, h->metric, h->dim, h->m, h->rows, n, h->levels, h->nbr0, h->upper_off, h->upper, entry, h->levels[entry]);
	return VB_OK;
}

int
vb_hnsw_free(vb_hnsw *h)
{
	if (h)
	{
		if (h->g)
			pgv_hnsw_free(h->g);
		free(h->rows);
		free(h->levels);
		free(h->nbr0);
		free(h->upper_off);
		free(h->upper);
		free(h->dup_of);
		free(h);
		mock_live_handles--;
	}
	return VB_OK;
}

int64_t mock_hnsw_graph(const vb_hnsw *h, const void **rows, const int32_t **levels, const int32_t **nbr0, const int64_t **upper_off,
						const int32_t **upper, int64_t *slots, int64_t *entry)
{
	*rows = h->rows;
	*levels = h->levels;
	*nbr0 = h->nbr0;
	*upper_off = h->upper_off;
	*upper = h->upper;
	*slots = h->slots;
	*entry = h->entry;
	return h->n;
}

int
vb_hnsw_search(vb_hnsw *h, const void *queries, int64_t nq, int ef, int k, int64_t *out_ids, double *out_dist, int64_t *out_ndist)
{
	size_t		rb = pgv_row_bytes(h->elem, h->dim);

	for (int64_t q = 0; q < nq; q++)
	{
		int64_t    *ids = malloc(sizeof(int64_t) * (size_t) (ef + 2));
		double	   *d = malloc(sizeof(double) * (size_t) (ef + 2));
		int64_t		nd = 0;
		int			n = h->g ? pgv_hnsw_search(h->g, (const char *) queries + rb * (size_t) q, ef, PGV_TIES_TOTAL_ORDER, ids, d, &nd) : 0;

		for (int i = 0; i < k; i++)
		{
			out_ids[q * k + i] = i < n ? ids[i] : -1;
			out_dist[q * k + i] = i < n ? d[i] : 1.0 / 0.0;
		}
		if (out_ndist)
			out_ndist[q] = nd;
		free(ids);
		free(d);
	}
	return VB_OK;
}

/* the iterative scan on the oracle's (single query): the whole sequence is produced up front and handed out batch by
 * batch -- the searched batches as the oracle cut them, the drain past max_scan_tuples ef_search at a time */
struct vb_hnsw_scan
{
	int			ef;
	int64_t		n,
				next,
				tuples;
	int64_t    *ids;
	double	   *dist;
	int32_t    *batch;
};

int
vb_hnsw_scan_begin(vb_hnsw *h, const void *queries, int64_t nq, int ef_search, int64_t max_scan_tuples, vb_hnsw_scan **out)
{
	vb_hnsw_scan *sc;
	int64_t		cap = h->n > 0 ? h->n : 1;

	if (nq != 1)
		return VB_EINVAL;
	sc = calloc(1, sizeof(*sc));
	sc->ef = ef_search;
	sc->ids = malloc(sizeof(int64_t) * (size_t) cap);
	sc->dist = malloc(sizeof(double) * (size_t) cap);
	sc->batch = malloc(sizeof(int32_t) * (size_t) cap);
	sc->n = h->g ? pgv_hnsw_iter_scan(h->g, queries, ef_search, PGV_TIES_TOTAL_ORDER, max_scan_tuples, cap, sc->ids, sc->dist, sc->batch,
									  &sc->tuples) : 0;
	mock_live_handles++;
	*out = sc;
	return VB_OK;
}

int
vb_hnsw_scan_next(vb_hnsw_scan *sc, int64_t *out_ids, double *out_distances, int32_t *out_counts)
{
	int			c = 0;

	while (sc->next < sc->n && c < sc->ef && (c == 0 || sc->batch[sc->next] == sc->batch[sc->next - 1]))
	{
		out_ids[c] = sc->ids[sc->next];
		out_distances[c] = sc->dist[sc->next];
		c++;
		sc->next++;
	}
	for (int i = c; i < sc->ef; i++)
	{
		out_ids[i] = -1;
		out_distances[i] = 1.0 / 0.0;
	}
	out_counts[0] = c;
	return VB_OK;
}

int
vb_hnsw_scan_tuples(vb_hnsw_scan *sc, int64_t *out_tuples)
{
	out_tuples[0] = sc->tuples;
	return VB_OK;
}

int
vb_hnsw_scan_end(vb_hnsw_scan *sc)
{
	if (sc)
	{
		free(sc->ids);
		free(sc->dist);
		free(sc->batch);
		free(sc);
		mock_live_handles--;
	}
	return VB_OK;
}

/* vb_hnsw_build on the oracle's serial build; the library numbers elements by ROW (a folded duplicate keeps its row
 * number and points at its element through dup_of), the oracle numbers elements densely: translate */
int
vb_hnsw_build(vb_hnsw *h, const void *rows, int64_t n, int ef_construction, uint64_t seed, const int32_t *levels)
{
	size_t		rb = pgv_row_bytes(h->elem, h->dim);
	PgvHnsw    *g = pgv_hnsw_create(h->elem, h->metric, h->dim, h->m, ef_construction, seed ? seed : 1);
	int64_t		ne, slots, entry = -1, uslots = 0;
	int			entry_level = 0, lm0 = 2 * h->m;
	int32_t    *el, *en0, *eup, *nht;
	int64_t    *euo, *erow, *ht;

	if (mock_fail_next_load)
	{
		mock_fail_next_load = 0;
		pgv_hnsw_free(g);
		snprintf(mock_err, sizeof(mock_err), "mock: injected load failure");
		return VB_ENOMEM;
	}
	h->rows = malloc(rb * (size_t) (n ? n : 1));
	memcpy(h->rows, rows, rb * (size_t) n);
	if (levels)
		pgv_hnsw_build_levels(g, h->rows, n, levels);
	else
		pgv_hnsw_build(g, h->rows, n);
	ne = pgv_hnsw_count(g);
	el = malloc(sizeof(int32_t) * (size_t) (ne + 1));
	en0 = malloc(sizeof(int32_t) * (size_t) (ne + 1) * lm0);
	euo = malloc(sizeof(int64_t) * (size_t) (ne + 1));
	pgv_hnsw_export_layer0(g, el, en0);
	slots = pgv_hnsw_export_upper(g, euo, NULL);
	eup = malloc(sizeof(int32_t) * (size_t) (slots + 1) * h->m);
	pgv_hnsw_export_upper(g, euo, eup);
	erow = malloc(sizeof(int64_t) * (size_t) (ne + 1));
	nht = malloc(sizeof(int32_t) * (size_t) (ne + 1));
	ht = malloc(sizeof(int64_t) * (size_t) (ne + 1) * 10);
	pgv_hnsw_export_elements(g, erow, nht, ht);
	pgv_hnsw_entry(g, &entry, &entry_level);
	/* row-indexed arrays */
	h->n = n;
	h->levels = calloc((size_t) (n ? n : 1), sizeof(int32_t));
	h->nbr0 = malloc(sizeof(int32_t) * (size_t) (n ? n : 1) * lm0);
	h->upper_off = malloc(sizeof(int64_t) * (size_t) (n ? n : 1));
	memset(h->nbr0, 0xFF, sizeof(int32_t) * (size_t) n * lm0);
	{
		int32_t    *dup = malloc(sizeof(int32_t) * (size_t) (n ? n : 1));

		memset(dup, 0xFF, sizeof(int32_t) * (size_t) n);
		for (int64_t e = 0; e < ne; e++)
			for (int j = 1; j < nht[e]; j++)
				dup[ht[e * 10 + j]] = (int32_t) erow[e];
		h->upper = (int32_t *) dup;	/* parked; swapped below */
	}
	{
		int32_t    *dup = h->upper;
		int64_t		r;

		if (levels)
			for (r = 0; r < n; r++)
				h->levels[r] = levels[r];
		for (int64_t e = 0; e < ne; e++)
			h->levels[erow[e]] = el[e];
		for (r = 0; r < n; r++)
		{
			h->upper_off[r] = h->levels[r] > 0 ? uslots : -1;
			uslots += h->levels[r];
		}
		h->upper = malloc(sizeof(int32_t) * (size_t) (uslots ? uslots : 1) * h->m);
		memset(h->upper, 0xFF, sizeof(int32_t) * (size_t) uslots * h->m);
		for (int64_t e = 0; e < ne; e++)
		{
			int64_t		row = erow[e];

			for (int j = 0; j < lm0; j++)
				h->nbr0[row * lm0 + j] = en0[e * lm0 + j] >= 0 ? (int32_t) erow[en0[e * lm0 + j]] : -1;
			for (int lc = 1; lc <= el[e]; lc++)
				for (int j = 0; j < h->m; j++)
				{
					int32_t		v = eup[(euo[e] + lc - 1) * h->m + j];

					h->upper[(h->upper_off[row] + lc - 1) * h->m + j] = v >= 0 ? (int32_t) erow[v] : -1;
				}
		}
		h->slots = uslots;
		h->entry = entry >= 0 ? erow[entry] : -1;
		h->dup_of = dup;
	}
	pgv_hnsw_free(g);
	free(el); free(en0); free(euo); free(eup); free(erow); free(nht); free(ht);
	return VB_OK;
}

int64_t vb_hnsw_rows(const vb_hnsw *h) { return h ? h->n : 0; }
int64_t vb_hnsw_upper_slots(const vb_hnsw *h) { return h ? h->slots : 0; }

int
vb_hnsw_export(vb_hnsw *h, int32_t *levels, int32_t *nbr0, int64_t *upper_off, int32_t *upper, int64_t *entry, int32_t *dup_of)
{
	int64_t		n = h->n;

	if (entry)
		*entry = h->entry;
	if (levels)
		memcpy(levels, h->levels, sizeof(int32_t) * (size_t) n);
	if (nbr0)
		memcpy(nbr0, h->nbr0, sizeof(int32_t) * (size_t) n * 2 * h->m);
	if (upper_off)
		memcpy(upper_off, h->upper_off, sizeof(int64_t) * (size_t) n);
	if (upper)
		memcpy(upper, h->upper, sizeof(int32_t) * (size_t) h->slots * h->m);
	if (dup_of)
	{
		if (h->dup_of)
			memcpy(dup_of, h->dup_of, sizeof(int32_t) * (size_t) n);
		else
			memset(dup_of, 0xFF, sizeof(int32_t) * (size_t) n);
	}
	return VB_OK;
}

/* ---- tables, k-means, assign (the build glue) ---- */
struct vb_table
{
	int			elem, dim;
	int64_t		n;
	char	   *rows;
};

int
vb_table_create(int elem, int dim, vb_table **out)
{
	vb_table   *t = calloc(1, sizeof(vb_table));

	t->elem = elem;
	t->dim = dim;
	*out = t;
	mock_live_handles++;
	return VB_OK;
}

int
vb_table_append(vb_table *t, const void *rows, int64_t n)
{
	size_t		rb = pgv_row_bytes(t->elem, t->dim);

	if (mock_fail_next_load)
	{
		mock_fail_next_load = 0;
		snprintf(mock_err, sizeof(mock_err), "mock: injected load failure");
		return VB_ENOMEM;
	}
	t->rows = realloc(t->rows, rb * (size_t) (t->n + n + 1));
	memcpy(t->rows + rb * (size_t) t->n, rows, rb * (size_t) n);
	t->n += n;
	return VB_OK;
}

int
vb_table_free(vb_table *t)
{
	if (t)
	{
		free(t->rows);
		free(t);
		mock_live_handles--;
	}
	return VB_OK;
}

int
vb_kmeans_pp_init(vb_table *t, int kmeans_metric, void *centers, int k, uint64_t seed)
{
	pgv_kmeans_pp_init(t->elem, kmeans_metric, t->dim, t->rows, t->n, centers, k, seed);
	return VB_OK;
}

int
vb_kmeans(vb_table *t, int kmeans_metric, void *centers, int k, int max_iter, uint64_t seed, vb_allreduce_fn allreduce, void *ctx,
		  int *iters_out)
{
	(void) allreduce; (void) ctx;
	*iters_out = pgv_kmeans_elkan(t->elem, kmeans_metric, t->dim, t->rows, t->n, centers, k, max_iter, seed, NULL);
	return VB_OK;
}

int
vb_assign(vb_table *t, int metric, const void *centers, int k, int32_t *out_list)
{
	pgv_ivf_assign(t->elem, metric, t->dim, t->rows, t->n, centers, k, 1, out_list);
	return VB_OK;
}
