/*
 * pgv_hnsw.c -- CPU oracle: HNSW layer search, scan driver and in-memory
 * build.  TEST INFRASTRUCTURE ONLY (see pgv_oracle.h).
 *
 * Follows src/hnswutils.c:824-987 (HnswSearchLayer), src/hnswscan.c:25-56
 * (GetScanItems), src/hnswutils.c:1040-1357 (SelectNeighbors,
 * HnswUpdateConnection, HnswFindElementNeighbors) and src/hnswbuild.c:341-480
 * (in-memory insert).  Pages/TIDs are replaced by element ids; the
 * neighbour order of an element is the order HnswSetNeighborTuple
 * (hnswutils.c:455-486) would write to disk.
 *
 * Tie order comes from PostgreSQL's pairing heap (pgv_pairingheap.h) in
 * PGV_TIES_PG_PAIRINGHEAP mode; PGV_TIES_TOTAL_ORDER replaces every distance
 * comparison by the lexicographic (distance, element id) comparison, which is
 * the deterministic instance of the same algorithm that the GPU executes.
 * Level draws use this file's own PRNG (pg_prng is PG core): parity on random
 * draws is UNPINNED; graphs are shared between oracle and GPU instead.
 */
#include "pgv_oracle.h"
#include "pgv_pairingheap.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HNSW_HEAPTIDS 10		/* src/hnsw.h:69 */
#define LAYER_M(m, lc) ((lc) == 0 ? (m) * 2 : (m))	/* src/hnsw.h:127 */

typedef struct
{
	int32_t		id;
	float		distance;
	uint8_t		closer;
}			Cand;

typedef struct
{
	int			length;
	uint8_t		closerSet;
	Cand	   *items;
}			NbrArray;

typedef struct
{
	int			level;
	int64_t		row;
	int			heaptidsLength;
	int64_t		heaptids[HNSW_HEAPTIDS];
	NbrArray   *nbr;			/* level + 1 arrays */
}			Element;

struct PgvHnsw
{
	int			elem,
				metric,
				dim,
				m,
				efc,
				maxLevel;
	double		ml;
	size_t		rb;
	const void *rows;
	Element    *el;
	int64_t		n,
				cap;
	int64_t		entry;
	uint64_t	rs0,
				rs1;
	const int32_t *levels_in;	/* caller-supplied level draws (NULL: drawn from the seed) */
	uint32_t   *visited;		/* build-time epoch set */
	uint32_t	epoch;
	int64_t		visited_cap;
};

/* --------------------------------------------------------------- rng */
static uint64_t
sm64(uint64_t *x)
{
	uint64_t	z = (*x += 0x9e3779b97f4a7c15ULL);

	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}

static double
rnd_double(PgvHnsw *g)
{
	uint64_t	s0 = g->rs0,
				s1 = g->rs1,
				r = ((s0 * 5) << 7 | (s0 * 5) >> 57) * 9;

	s1 ^= s0;
	g->rs0 = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
	g->rs1 = (s1 << 37) | (s1 >> 27);
	return (double) (r >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------- search */

typedef struct
{
	ph_node		c_node;
	ph_node		w_node;
	int32_t		element;
	double		distance;
}			SearchCand;

typedef struct
{
	SearchCand **chunks;
	int			nchunks,
				used;
}			Arena;

#define ARENA_CHUNK 1024

static SearchCand *
arena_new(Arena *a, int32_t element, double distance)
{
	SearchCand *sc;

	if (a->nchunks == 0 || a->used == ARENA_CHUNK)
	{
		a->chunks = realloc(a->chunks, sizeof(SearchCand *) * (size_t) (a->nchunks + 1));
		a->chunks[a->nchunks++] = malloc(sizeof(SearchCand) * ARENA_CHUNK);
		a->used = 0;
	}
	sc = &a->chunks[a->nchunks - 1][a->used++];
	sc->element = element;
	sc->distance = distance;
	return sc;
}

static void
arena_free(Arena *a)
{
	for (int i = 0; i < a->nchunks; i++)
		free(a->chunks[i]);
	free(a->chunks);
	a->chunks = NULL;
	a->nchunks = a->used = 0;
}

static int	g_tie_mode_dummy;

/* returns <0, 0, >0 for a before/equal/after b in "nearer" order */
static inline int
key_cmp(double da, int32_t ia, double db, int32_t ib, int total)
{
	if (da < db)
		return -1;
	if (da > db)
		return 1;
	if (total)
		return ia < ib ? -1 : (ia > ib);
	return 0;
}

/* CompareNearestCandidates (hnswutils.c:626-636) */
static int
cmp_nearest(const ph_node *a, const ph_node *b, void *arg)
{
	const SearchCand *x = ph_container(SearchCand, c_node, a);
	const SearchCand *y = ph_container(SearchCand, c_node, b);

	return -key_cmp(x->distance, x->element, y->distance, y->element, *(int *) arg);
}

/* CompareFurthestCandidates (hnswutils.c:656-666) */
static int
cmp_furthest(const ph_node *a, const ph_node *b, void *arg)
{
	const SearchCand *x = ph_container(SearchCand, w_node, a);
	const SearchCand *y = ph_container(SearchCand, w_node, b);

	return key_cmp(x->distance, x->element, y->distance, y->element, *(int *) arg);
}

/* CompareNearestDiscardedCandidates (hnswutils.c:641-651): the discarded heap links candidates through w_node */
static int
cmp_nearest_discarded(const ph_node *a, const ph_node *b, void *arg)
{
	const SearchCand *x = ph_container(SearchCand, w_node, a);
	const SearchCand *y = ph_container(SearchCand, w_node, b);

	return -key_cmp(x->distance, x->element, y->distance, y->element, *(int *) arg);
}

static inline double
elem_distance(const PgvHnsw *g, const void *q, int32_t e)
{
	/* NULL query: every distance is 0 (hnswutils.c:555-556) */
	if (q == NULL)
		return 0;
	/* HnswGetDistance(q->value, element value) (hnswutils.c:524-528, 585-591) */
	return pgv_distance(g->elem, g->metric, g->dim, q, (const char *) g->rows + (size_t) g->el[e].row * g->rb);
}

typedef struct
{
	SearchCand **items;
	int			n;
}			CandList;

/*
 * HnswSearchLayer (hnswutils.c:824-987), in-memory/on-disk semantics unified.
 * ep/out lists are furthest-first like the reference's `w`.  visited is an
 * epoch set; initVisited bumps the epoch.  tuples may be NULL.
 */
static CandList
search_layer_x(const PgvHnsw *g, const void *q, CandList ep, int ef, int lc, int tie_total,
			   uint32_t *visited, uint32_t epoch, int64_t *tuples, Arena *arena, ph_heap *discarded, int initVisited);

static CandList
search_layer(const PgvHnsw *g, const void *q, CandList ep, int ef, int lc, int tie_total,
			 uint32_t *visited, uint32_t epoch, int64_t *tuples, Arena *arena)
{
	return search_layer_x(g, q, ep, ef, lc, tie_total, visited, epoch, tuples, arena, NULL, 1);
}

/*
 * discarded (may be NULL): the iterative scan's heap of candidates that were seen but are not in W -- rejected
 * neighbours (hnswutils.c:929-937) and candidates evicted from W (:968-973).  initVisited = 0 resumes on the visited
 * set of the previous call: entry points are neither re-added nor counted (:864-873).
 */
static CandList
search_layer_x(const PgvHnsw *g, const void *q, CandList ep, int ef, int lc, int tie_total,
			   uint32_t *visited, uint32_t epoch, int64_t *tuples, Arena *arena, ph_heap *discarded, int initVisited)
{
	ph_heap		C,
				W;
	int			wlen = 0;
	int			lm = LAYER_M(g->m, lc);
	int			total = tie_total;
	int32_t    *unvisited = malloc(sizeof(int32_t) * (size_t) lm);
	CandList	w;

	ph_init(&C, cmp_nearest, &total);
	ph_init(&W, cmp_furthest, &total);

	for (int i = 0; i < ep.n; i++)
	{
		SearchCand *sc = ep.items[i];

		if (initVisited)
		{
			visited[sc->element] = epoch;
			if (tuples)
				(*tuples)++;
		}
		ph_add(&C, &sc->c_node);
		ph_add(&W, &sc->w_node);
		wlen++;
	}

	while (!ph_is_empty(&C))
	{
		SearchCand *c = ph_container(SearchCand, c_node, ph_remove_first(&C));
		SearchCand *f = ph_container(SearchCand, w_node, ph_first(&W));
		const Element *ce;
		int			unvisitedLength = 0;

		/* c->distance > f->distance (hnswutils.c:894) */
		if (key_cmp(c->distance, c->element, f->distance, f->element, total) > 0)
			break;

		ce = &g->el[c->element];
		if (lc <= ce->level)
		{
			const NbrArray *na = &ce->nbr[lc];

			/* HnswLoadUnvisitedFrom{Memory,Disk} (hnswutils.c:733-756, 796-819) */
			for (int i = 0; i < na->length; i++)
			{
				int32_t		nid = na->items[i].id;

				if (visited[nid] != epoch)
				{
					visited[nid] = epoch;
					unvisited[unvisitedLength++] = nid;
				}
			}
		}
		if (tuples)
			(*tuples) += unvisitedLength;

		for (int i = 0; i < unvisitedLength; i++)
		{
			int32_t		eid = unvisited[i];
			double		eDistance;
			int			alwaysAdd = wlen < ef;
			SearchCand *e;

			f = ph_container(SearchCand, w_node, ph_first(&W));
			eDistance = elem_distance(g, q, eid);

			/* eDistance < f->distance || alwaysAdd (hnswutils.c:913-936) */
			if (!(key_cmp(eDistance, eid, f->distance, f->element, total) < 0 || alwaysAdd))
			{
				if (discarded)
				{
					e = arena_new(arena, eid, eDistance);
					ph_add(discarded, &e->w_node);
				}
				continue;
			}
			/* hnswutils.c:949-950 */
			if (g->el[eid].level < lc)
				continue;

			e = arena_new(arena, eid, eDistance);
			ph_add(&C, &e->c_node);
			ph_add(&W, &e->w_node);
			wlen++;
			/* No need to decrement wlen (hnswutils.c:962-974) */
			if (wlen > ef)
			{
				SearchCand *d = ph_container(SearchCand, w_node, ph_remove_first(&W));

				if (discarded)
					ph_add(discarded, &d->w_node);
			}
		}
	}

	/* drain W furthest-first (hnswutils.c:979-984) */
	w.items = malloc(sizeof(SearchCand *) * (size_t) (ef + ep.n + 1));
	w.n = 0;
	while (!ph_is_empty(&W))
		w.items[w.n++] = ph_container(SearchCand, w_node, ph_remove_first(&W));
	free(unvisited);
	return w;
}

/* GetScanItems (hnswscan.c:25-56) */
static int
hnsw_search_impl(const PgvHnsw *g, const void *q, int ef, int tie_mode, uint32_t *visited, uint32_t *epoch,
				 int64_t *out_ids, double *out_dist, int64_t *n_dist)
{
	Arena		arena = {0};
	CandList	ep,
				w;
	int64_t		tuples = 0;
	int			total = tie_mode == PGV_TIES_TOTAL_ORDER;
	int			n;

	if (g->entry < 0)
	{
		if (n_dist)
			*n_dist = 0;
		return 0;
	}
	ep.items = malloc(sizeof(SearchCand *));
	ep.items[0] = arena_new(&arena, (int32_t) g->entry, elem_distance(g, q, (int32_t) g->entry));
	ep.n = 1;

	for (int lc = g->el[g->entry].level; lc >= 1; lc--)
	{
		(*epoch)++;
		w = search_layer(g, q, ep, 1, lc, total, visited, *epoch, NULL, &arena);
		free(ep.items);
		ep = w;
	}
	(*epoch)++;
	w = search_layer(g, q, ep, ef, 0, total, visited, *epoch, &tuples, &arena);
	free(ep.items);

	/* hnswgettuple pops llast(w): nearest first (hnswscan.c:293-326) */
	n = w.n;
	for (int i = 0; i < n; i++)
	{
		SearchCand *sc = w.items[n - 1 - i];

		out_ids[i] = sc->element;
		out_dist[i] = sc->distance;
	}
	free(w.items);
	arena_free(&arena);
	if (n_dist)
		*n_dist = tuples;
	return n;
}

int
pgv_hnsw_search(const PgvHnsw *g, const void *q, int ef, int tie_mode, int64_t *out_ids, double *out_dist, int64_t *n_dist)
{
	uint32_t   *visited = calloc((size_t) (g->n > 0 ? g->n : 1), sizeof(uint32_t));
	uint32_t	epoch = 0;
	int			n = hnsw_search_impl(g, q, ef, tie_mode, visited, &epoch, out_ids, out_dist, n_dist);

	free(visited);
	return n;
}

/*
 * The iterative scan as hnswgettuple drives it (hnswscan.c:228-340, relaxed order; the strict mode is a filter on
 * this sequence, :316-322): the first batch is GetScanItems with the discarded heap (:25-56), every further batch
 * ResumeScanItems (:62-87) from the ef nearest discarded candidates on the same visited set; once the tuples counter
 * has reached max_scan_tuples the remaining discarded candidates are returned nearest first (:247-254).  (The
 * work_mem bound of :247 is not modelled.)  Emits up to max_out elements; out_batch[i] = the batch that produced
 * output i (-1: the final drain).  Returns the number emitted; *tuples_out = the counter.
 */
int64_t
pgv_hnsw_iter_scan(const PgvHnsw *g, const void *q, int ef, int tie_mode, int64_t max_scan_tuples, int64_t max_out,
				   int64_t *out_ids, double *out_dist, int32_t *out_batch, int64_t *tuples_out)
{
	Arena		arena = {0};
	CandList	ep,
				w;
	int64_t		tuples = 0,
				n_out = 0;
	int			total = tie_mode == PGV_TIES_TOTAL_ORDER;
	uint32_t   *visited;
	uint32_t	epoch = 0;
	ph_heap		discarded;
	int32_t		batch = 0;

	if (tuples_out)
		*tuples_out = 0;
	if (g->entry < 0)
		return 0;
	visited = calloc((size_t) g->n, sizeof(uint32_t));
	ph_init(&discarded, cmp_nearest_discarded, &total);
	ep.items = malloc(sizeof(SearchCand *));
	ep.items[0] = arena_new(&arena, (int32_t) g->entry, elem_distance(g, q, (int32_t) g->entry));
	ep.n = 1;
	for (int lc = g->el[g->entry].level; lc >= 1; lc--)
	{
		epoch++;
		w = search_layer(g, q, ep, 1, lc, total, visited, epoch, NULL, &arena);
		free(ep.items);
		ep = w;
	}
	epoch++;
	w = search_layer_x(g, q, ep, ef, 0, total, visited, epoch, &tuples, &arena, &discarded, 1);
	free(ep.items);

	for (;;)
	{
		/* hnswgettuple pops llast(w): nearest first */
		for (int i = w.n - 1; i >= 0 && n_out < max_out; i--)
		{
			out_ids[n_out] = w.items[i]->element;
			out_dist[n_out] = w.items[i]->distance;
			out_batch[n_out] = batch;
			n_out++;
		}
		free(w.items);
		w.items = NULL;
		w.n = 0;
		if (n_out >= max_out || ph_is_empty(&discarded))
			break;
		if (tuples >= max_scan_tuples)
		{
			/* return the remaining candidates one at a time, nearest first */
			while (!ph_is_empty(&discarded) && n_out < max_out)
			{
				SearchCand *sc = ph_container(SearchCand, w_node, ph_remove_first(&discarded));

				out_ids[n_out] = sc->element;
				out_dist[n_out] = sc->distance;
				out_batch[n_out] = -1;
				n_out++;
			}
			break;
		}
		/* ResumeScanItems: the next batch_size = ef nearest discarded candidates are the entry points */
		ep.items = malloc(sizeof(SearchCand *) * (size_t) ef);
		ep.n = 0;
		for (int i = 0; i < ef && !ph_is_empty(&discarded); i++)
			ep.items[ep.n++] = ph_container(SearchCand, w_node, ph_remove_first(&discarded));
		batch++;
		w = search_layer_x(g, q, ep, ef, 0, total, visited, epoch, &tuples, &arena, &discarded, 0);
		free(ep.items);
	}
	free(w.items);
	free(visited);
	arena_free(&arena);
	if (tuples_out)
		*tuples_out = tuples;
	return n_out;
}

void
pgv_hnsw_search_batch(const PgvHnsw *g, const void *queries, int64_t nq, int ef, int tie_mode, int threads, int k,
					  int64_t *out_ids, double *out_dist, int64_t *n_dist)
{
	(void) threads;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
	{
		uint32_t   *visited = calloc((size_t) (g->n > 0 ? g->n : 1), sizeof(uint32_t));
		uint32_t	epoch = 0;
		int64_t    *ids = malloc(sizeof(int64_t) * (size_t) (ef + 2));
		double	   *dist = malloc(sizeof(double) * (size_t) (ef + 2));

#pragma omp for schedule(dynamic, 4)
		for (int64_t i = 0; i < nq; i++)
		{
			int64_t		nd = 0;
			int			n;

			if (epoch > 0xfffffff0u)
			{
				memset(visited, 0, sizeof(uint32_t) * (size_t) g->n);
				epoch = 0;
			}
			n = hnsw_search_impl(g, (const char *) queries + (size_t) i * g->rb, ef, tie_mode, visited, &epoch, ids, dist, &nd);
			for (int j = 0; j < k; j++)
			{
				out_ids[i * k + j] = j < n ? ids[j] : -1;
				out_dist[i * k + j] = j < n ? dist[j] : INFINITY;
			}
			if (n_dist)
				n_dist[i] = nd;
		}
		free(visited);
		free(ids);
		free(dist);
	}
}

/* -------------------------------------------------------------- build */

PgvHnsw *
pgv_hnsw_create(int elem, int metric, int dim, int m, int ef_construction, uint64_t seed)
{
	PgvHnsw    *g = calloc(1, sizeof(PgvHnsw));
	int			maxl;

	g->elem = elem;
	g->metric = metric;
	g->dim = dim;
	g->m = m;
	g->efc = ef_construction;
	g->ml = 1 / log((double) m);	/* HnswGetMl (hnsw.h:130) */
	/* HnswGetMaxLevel (hnsw.h:133): (8192 - 24 - 8 - 4 - 4) / 6 / m - 2, capped at 63 */
	maxl = (int) ((8192 - 24 - 8 - 4 - 4) / 6 / m) - 2;
	g->maxLevel = maxl < 63 ? maxl : 63;
	g->rb = pgv_row_bytes(elem, dim);
	g->entry = -1;
	g->rs0 = sm64(&seed);
	g->rs1 = sm64(&seed);
	return g;
}

void
pgv_hnsw_free(PgvHnsw *g)
{
	if (!g)
		return;
	for (int64_t i = 0; i < g->n; i++)
	{
		for (int lc = 0; lc <= g->el[i].level; lc++)
			free(g->el[i].nbr[lc].items);
		free(g->el[i].nbr);
	}
	free(g->el);
	free(g->visited);
	free(g);
}

static void
init_neighbors(PgvHnsw *g, Element *e)
{
	e->nbr = calloc((size_t) e->level + 1, sizeof(NbrArray));
	for (int lc = 0; lc <= e->level; lc++)
		e->nbr[lc].items = calloc((size_t) LAYER_M(g->m, lc), sizeof(Cand));
}

static inline float
pair_distance(const PgvHnsw *g, int32_t a, int32_t b)
{
	return (float) pgv_distance(g->elem, g->metric, g->dim,
								(const char *) g->rows + (size_t) g->el[a].row * g->rb,
								(const char *) g->rows + (size_t) g->el[b].row * g->rb);
}

/* CheckElementCloser (hnswutils.c:1043-1062) */
static int
check_closer(const PgvHnsw *g, const Cand *e, Cand **r, int rlen)
{
	for (int i = 0; i < rlen; i++)
	{
		float		distance = pair_distance(g, e->id, r[i]->id);

		if (distance <= e->distance)
			return 0;
	}
	return 1;
}

/* CompareCandidateDistances{,Offset} (hnswutils.c:1000-1038): sort so the LAST item is nearest, smaller id later */
static int
cmp_cand_desc(const void *pa, const void *pb)
{
	const Cand *a = *(Cand *const *) pa,
			   *b = *(Cand *const *) pb;

	if (a->distance < b->distance)
		return 1;
	if (a->distance > b->distance)
		return -1;
	if (a->id < b->id)
		return 1;
	if (a->id > b->id)
		return -1;
	return 0;
}

/*
 * SelectNeighbors (hnswutils.c:1067-1160).  c: candidate pointers, ordered
 * furthest..nearest unless sortCandidates.  Returns count written to r.
 */
static int
select_neighbors(const PgvHnsw *g, Cand **c, int clen, int lm, uint8_t *closerSet, Cand *newCandidate,
				 Cand **pruned, int sortCandidates, Cand **r)
{
	Cand	  **w,
			  **wd,
			  **added;
	int			wlen = clen,
				rlen = 0,
				wdlen = 0,
				wdoff = 0,
				addedlen = 0;
	int			mustCalculate = !(*closerSet);
	int			removedAny = 0;

	if (clen <= lm)
	{
		memcpy(r, c, sizeof(Cand *) * (size_t) clen);
		return clen;
	}
	w = malloc(sizeof(Cand *) * (size_t) clen);
	wd = malloc(sizeof(Cand *) * (size_t) clen);
	added = malloc(sizeof(Cand *) * (size_t) clen);
	memcpy(w, c, sizeof(Cand *) * (size_t) clen);
	if (sortCandidates)
		qsort(w, (size_t) wlen, sizeof(Cand *), cmp_cand_desc);

	while (wlen > 0 && rlen < lm)
	{
		Cand	   *e = w[--wlen];

		if (mustCalculate)
			e->closer = (uint8_t) check_closer(g, e, r, rlen);
		else if (addedlen > 0)
		{
			if (e->closer)
			{
				e->closer = (uint8_t) check_closer(g, e, added, addedlen);
				if (!e->closer)
					removedAny = 1;
			}
			else if (removedAny)
			{
				e->closer = (uint8_t) check_closer(g, e, r, rlen);
				if (e->closer)
					added[addedlen++] = e;
			}
		}
		else if (e == newCandidate)
		{
			e->closer = (uint8_t) check_closer(g, e, r, rlen);
			if (e->closer)
				added[addedlen++] = e;
		}

		if (e->closer)
			r[rlen++] = e;
		else
			wd[wdlen++] = e;
	}

	*closerSet = (uint8_t) sortCandidates;

	/* keep pruned connections */
	while (wdoff < wdlen && rlen < lm)
		r[rlen++] = wd[wdoff++];

	if (pruned)
	{
		if (wdoff < wdlen)
			*pruned = wd[wdoff];
		else
			*pruned = w[0];		/* linitial(w): furthest remaining */
	}
	free(w);
	free(wd);
	free(added);
	return rlen;
}

/* HnswUpdateConnection (hnswutils.c:1183-1231) */
static void
update_connection(const PgvHnsw *g, NbrArray *na, int32_t newId, float distance, int lm)
{
	Cand		newHc;

	newHc.id = newId;
	newHc.distance = distance;
	newHc.closer = 0;

	if (na->length < lm)
		na->items[na->length++] = newHc;
	else
	{
		Cand	  **c = malloc(sizeof(Cand *) * (size_t) (na->length + 1));
		Cand	  **r = malloc(sizeof(Cand *) * (size_t) (na->length + 1));
		Cand	   *pruned = NULL;

		for (int i = 0; i < na->length; i++)
			c[i] = &na->items[i];
		c[na->length] = &newHc;
		select_neighbors(g, c, na->length + 1, lm, &na->closerSet, &newHc, &pruned, 1, r);
		if (pruned)
			for (int i = 0; i < na->length; i++)
				if (na->items[i].id == pruned->id)
				{
					na->items[i] = newHc;
					break;
				}
		free(c);
		free(r);
	}
}

/* HnswFindElementNeighbors (hnswutils.c:1280-1357), existing = false */
static void
find_element_neighbors(PgvHnsw *g, int32_t eid, int64_t entryPoint)
{
	Element    *element = &g->el[eid];
	const void *q = (const char *) g->rows + (size_t) element->row * g->rb;
	int			level = element->level;
	int			entryLevel;
	Arena		arena = {0};
	CandList	ep,
				w;

	if (entryPoint < 0)
		return;

	ep.items = malloc(sizeof(SearchCand *));
	ep.items[0] = arena_new(&arena, (int32_t) entryPoint, elem_distance(g, q, (int32_t) entryPoint));
	ep.n = 1;
	entryLevel = g->el[entryPoint].level;

	/* 1st phase */
	for (int lc = entryLevel; lc >= level + 1; lc--)
	{
		g->epoch++;
		w = search_layer(g, q, ep, 1, lc, 0, g->visited, g->epoch, NULL, &arena);
		free(ep.items);
		ep = w;
	}
	if (level > entryLevel)
		level = entryLevel;

	/* 2nd phase */
	for (int lc = level; lc >= 0; lc--)
	{
		int			lm = LAYER_M(g->m, lc);
		Cand	   *lw;
		Cand	  **lwp,
				  **r;
		int			rn;
		NbrArray   *na = &element->nbr[lc];

		g->epoch++;
		w = search_layer(g, q, ep, g->efc, lc, 0, g->visited, g->epoch, NULL, &arena);

		lw = malloc(sizeof(Cand) * (size_t) (w.n + 1));
		lwp = malloc(sizeof(Cand *) * (size_t) (w.n + 1));
		r = malloc(sizeof(Cand *) * (size_t) (w.n + 1));
		for (int i = 0; i < w.n; i++)
		{
			lw[i].id = w.items[i]->element;
			lw[i].distance = (float) w.items[i]->distance;
			lw[i].closer = 0;
			lwp[i] = &lw[i];
		}
		rn = select_neighbors(g, lwp, w.n, lm, &na->closerSet, NULL, NULL, 0, r);
		/* AddConnections (hnswutils.c:1165-1174) */
		for (int i = 0; i < rn; i++)
			na->items[na->length++] = *r[i];
		free(lw);
		free(lwp);
		free(r);
		free(ep.items);
		ep = w;
	}
	free(ep.items);
	arena_free(&arena);
}

static int
rows_equal(const PgvHnsw *g, int64_t ra, int64_t rb_)
{
	return memcmp((const char *) g->rows + (size_t) ra * g->rb, (const char *) g->rows + (size_t) rb_ * g->rb, g->rb) == 0;
}

/* InsertTupleInMemory / UpdateGraphInMemory (hnswbuild.c:341-480), serial */
static void
insert_row(PgvHnsw *g, int64_t row)
{
	int32_t		eid = (int32_t) g->n;
	Element    *e = &g->el[eid];
	int			level = g->levels_in ? g->levels_in[row] : (int) (-log(rnd_double(g)) * g->ml);	/* HnswInitElement (hnswutils.c:248-254) */
	int64_t		entryPoint = g->entry;

	if (level > g->maxLevel)
		level = g->maxLevel;
	memset(e, 0, sizeof(*e));
	e->level = level;
	e->row = row;
	e->heaptids[e->heaptidsLength++] = row;
	init_neighbors(g, e);
	g->n++;						/* visible to pair_distance; rolled back on duplicate */

	find_element_neighbors(g, eid, entryPoint);

	/* FindDuplicateInMemory (hnswbuild.c:343-364) */
	{
		NbrArray   *na = &e->nbr[0];

		for (int i = 0; i < na->length; i++)
		{
			Element    *ne = &g->el[na->items[i].id];

			if (!rows_equal(g, e->row, ne->row))
				break;
			if (ne->heaptidsLength < HNSW_HEAPTIDS)
			{
				ne->heaptids[ne->heaptidsLength++] = row;
				for (int lc = 0; lc <= e->level; lc++)
					free(e->nbr[lc].items);
				free(e->nbr);
				g->n--;
				return;
			}
		}
	}

	/* UpdateNeighborsInMemory (hnswbuild.c:381-410) */
	for (int lc = e->level; lc >= 0; lc--)
	{
		int			lm = LAYER_M(g->m, lc);
		NbrArray   *na = &e->nbr[lc];

		for (int i = 0; i < na->length; i++)
		{
			Cand	   *hc = &na->items[i];

			update_connection(g, &g->el[hc->id].nbr[lc], eid, hc->distance, lm);
		}
	}
	if (entryPoint < 0 || e->level > g->el[entryPoint].level)
		g->entry = eid;
}

void
pgv_hnsw_build(PgvHnsw *g, const void *rows, int64_t n)
{
	g->rows = rows;
	g->el = realloc(g->el, sizeof(Element) * (size_t) (g->n + n + 1));
	g->visited = realloc(g->visited, sizeof(uint32_t) * (size_t) (g->n + n + 1));
	memset(g->visited, 0, sizeof(uint32_t) * (size_t) (g->n + n + 1));
	g->epoch = 0;
	for (int64_t i = 0; i < n; i++)
		insert_row(g, i);
}

/* the same build with the level of every row given by the caller (the extension draws them from pg_prng) */
void
pgv_hnsw_build_levels(PgvHnsw *g, const void *rows, int64_t n, const int32_t *levels)
{
	g->levels_in = levels;
	pgv_hnsw_build(g, rows, n);
	g->levels_in = NULL;
}

int64_t
pgv_hnsw_count(const PgvHnsw *g)
{
	return g->n;
}

int
pgv_hnsw_entry(const PgvHnsw *g, int64_t *entry, int *entry_level)
{
	*entry = g->entry;
	*entry_level = g->entry >= 0 ? g->el[g->entry].level : -1;
	return g->m;
}

void
pgv_hnsw_export_layer0(const PgvHnsw *g, int32_t *levels, int32_t *nbr0)
{
	int			lm = g->m * 2;

	for (int64_t i = 0; i < g->n; i++)
	{
		levels[i] = g->el[i].level;
		for (int j = 0; j < lm; j++)
			nbr0[i * lm + j] = j < g->el[i].nbr[0].length ? g->el[i].nbr[0].items[j].id : -1;
	}
}

int64_t
pgv_hnsw_export_upper(const PgvHnsw *g, int64_t *upper_off, int32_t *upper)
{
	int64_t		slots = 0;

	for (int64_t i = 0; i < g->n; i++)
	{
		int			L = g->el[i].level;

		if (upper_off)
			upper_off[i] = L >= 1 ? slots : -1;
		for (int lc = 1; lc <= L; lc++)
		{
			if (upper)
				for (int j = 0; j < g->m; j++)
					upper[slots * g->m + j] = j < g->el[i].nbr[lc].length ? g->el[i].nbr[lc].items[j].id : -1;
			slots++;
		}
	}
	return slots;
}

/* element -> first row and heap tid list, for result expansion (hnswscan.c:293-311) */
void
pgv_hnsw_export_elements(const PgvHnsw *g, int64_t *elem_row, int32_t *n_heaptids, int64_t *heaptids)
{
	for (int64_t i = 0; i < g->n; i++)
	{
		elem_row[i] = g->el[i].row;
		n_heaptids[i] = g->el[i].heaptidsLength;
		for (int j = 0; j < HNSW_HEAPTIDS; j++)
			heaptids[i * HNSW_HEAPTIDS + j] = j < g->el[i].heaptidsLength ? g->el[i].heaptids[j] : -1;
	}
}

PgvHnsw *
pgv_hnsw_import(int elem, int metric, int dim, int m, const void *rows, int64_t n,
				const int32_t *levels, const int32_t *nbr0, const int64_t *upper_off, const int32_t *upper,
				int64_t entry, int entry_level)
{
	PgvHnsw    *g = pgv_hnsw_create(elem, metric, dim, m, 64, 1);

	(void) entry_level;
	g->rows = rows;
	g->el = calloc((size_t) (n + 1), sizeof(Element));
	g->n = n;
	g->entry = entry;
	for (int64_t i = 0; i < n; i++)
	{
		Element    *e = &g->el[i];

		e->level = levels[i];
		e->row = i;
		e->heaptidsLength = 1;
		e->heaptids[0] = i;
		init_neighbors(g, e);
		for (int lc = 0; lc <= e->level; lc++)
		{
			int			lm = LAYER_M(m, lc);
			const int32_t *src = lc == 0 ? nbr0 + i * lm : upper + (upper_off[i] + (lc - 1)) * m;

			for (int j = 0; j < lm && src[j] >= 0; j++)
			{
				e->nbr[lc].items[j].id = src[j];
				e->nbr[lc].length++;
			}
		}
	}
	(void) g_tie_mode_dummy;
	return g;
}
