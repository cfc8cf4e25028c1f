/*
 * pgv_ivfflat.c -- CPU oracle: IVFFlat probe selection, list scan, assign pass
 * and Elkan k-means.  TEST INFRASTRUCTURE ONLY (see pgv_oracle.h).
 *
 * The on-disk page walk (ReadBuffer / PageGetItem) of the reference is
 * replaced by flat arrays (centres; rows grouped by list); the loop structure,
 * comparison operators and arithmetic types follow the cited lines.
 */
#include "pgv_oracle.h"
#include "pgv_pairingheap.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ rng */
/* stands in for pg_prng (xoroshiro128**, PG core; stream not reproduced) */
typedef struct
{
	uint64_t	s0,
				s1;
}			Rng;

static uint64_t
splitmix64(uint64_t *x)
{
	uint64_t	z = (*x += 0x9e3779b97f4a7c15ULL);

	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}

static void
rng_seed(Rng *r, uint64_t seed)
{
	r->s0 = splitmix64(&seed);
	r->s1 = splitmix64(&seed);
}

static uint64_t
rotl(uint64_t x, int k)
{
	return (x << k) | (x >> (64 - k));
}

static uint64_t
rng_next(Rng *r)
{
	uint64_t	s0 = r->s0,
				s1 = r->s1,
				res = rotl(s0 * 5, 7) * 9;

	s1 ^= s0;
	r->s0 = rotl(s0, 24) ^ s1 ^ (s1 << 16);
	r->s1 = rotl(s1, 37);
	return res;
}

static double
rng_double(Rng *r)
{
	return (double) (rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
}

/* -------------------------------------------------------- GetScanLists */

typedef struct
{
	ph_node		ph;
	int			list;
	double		distance;
}			ScanList;

/* 0: ties as PostgreSQL's pairing heap leaves them; 1: ties by list number (the deterministic instance the GPU uses) */
static int	ivf_tie_total = 0;

void
pgv_ivf_set_tie_mode(int total_order)
{
	ivf_tie_total = total_order;
}

/* CompareLists (src/ivfscan.c:32-42): furthest list at the root */
static int
compare_lists(const ph_node *a, const ph_node *b, void *arg)
{
	double		da = ph_container(ScanList, ph, a)->distance;
	double		db = ph_container(ScanList, ph, b)->distance;

	(void) arg;
	if (da > db)
		return 1;
	if (da < db)
		return -1;
	if (ivf_tie_total)
	{
		int			la = ph_container(ScanList, ph, a)->list;
		int			lb = ph_container(ScanList, ph, b)->list;

		return la > lb ? 1 : (la < lb ? -1 : 0);
	}
	return 0;
}

/* GetScanLists (src/ivfscan.c:47-118) */
int
pgv_ivf_scan_lists(const PgvIvfIndex *ix, const void *q, int max_probes, int *out_lists, double *out_dist)
{
	ph_heap		heap;
	ScanList   *slots;
	int			listCount = 0;
	double		maxDistance = DBL_MAX;
	size_t		rb = pgv_row_bytes(ix->elem, ix->dim);

	if (max_probes > ix->lists)
		max_probes = ix->lists;	/* ivfscan.c:279-283 */
	if (max_probes <= 0)
		return 0;
	slots = malloc(sizeof(ScanList) * (size_t) max_probes);
	ph_init(&heap, compare_lists, NULL);

	for (int l = 0; l < ix->lists; l++)
	{
		/* NULL query: ZeroDistance (ivfscan.c:192-196) */
		double		distance = q == NULL ? 0.0 :
			pgv_distance(ix->elem, ix->metric, ix->dim, (const char *) ix->centers + (size_t) l * rb, q);

		if (listCount < max_probes)
		{
			ScanList   *sl = &slots[listCount++];

			sl->list = l;
			sl->distance = distance;
			ph_add(&heap, &sl->ph);
			if (listCount == max_probes)
				maxDistance = ph_container(ScanList, ph, ph_first(&heap))->distance;
		}
		else if (distance < maxDistance)
		{
			ScanList   *sl = ph_container(ScanList, ph, ph_remove_first(&heap));

			sl->list = l;
			sl->distance = distance;
			ph_add(&heap, &sl->ph);
			maxDistance = ph_container(ScanList, ph, ph_first(&heap))->distance;
		}
	}

	/* ivfscan.c:114-115: pop furthest-first into the tail => nearest first */
	for (int i = listCount - 1; i >= 0; i--)
	{
		ScanList   *sl = ph_container(ScanList, ph, ph_remove_first(&heap));

		out_lists[i] = sl->list;
		if (out_dist)
			out_dist[i] = sl->distance;
	}
	free(slots);
	return listCount;
}

/* -------------------------------------------------------- GetScanItems */

typedef struct
{
	double		d;
	int64_t		id;
}			Item;

/* stable merge sort ascending on d; NaN last (float8 btree order). tuplesort itself is not stable:
 * tie order is unspecified in the reference, scan order is this oracle's deterministic choice. */
static int
item_less(const Item *a, const Item *b)
{
	int			an = isnan(a->d),
				bn = isnan(b->d);

	if (an || bn)
		return !an && bn;
	return a->d < b->d;
}

static void
merge_sort_items(Item *v, Item *tmp, int64_t n)
{
	int64_t		h,
				i,
				j,
				k;

	if (n < 2)
		return;
	h = n / 2;
	merge_sort_items(v, tmp, h);
	merge_sort_items(v + h, tmp, n - h);
	memcpy(tmp, v, sizeof(Item) * (size_t) h);
	i = 0;
	j = h;
	k = 0;
	while (i < h && j < n)
		v[k++] = item_less(&v[j], &tmp[i]) ? v[j++] : tmp[i++];
	while (i < h)
		v[k++] = tmp[i++];
}

/* GetScanItems (src/ivfscan.c:123-187) */
int64_t
pgv_ivf_scan_items(const PgvIvfIndex *ix, const void *q, const int *lists, int probes, int64_t cap, int64_t *out_ids, double *out_dist)
{
	size_t		rb = pgv_row_bytes(ix->elem, ix->dim);
	int64_t		total = 0,
				n = 0;
	Item	   *items,
			   *tmp;

	for (int p = 0; p < probes; p++)
		total += ix->list_offsets[lists[p] + 1] - ix->list_offsets[lists[p]];
	items = malloc(sizeof(Item) * (size_t) (total > 0 ? total : 1));
	tmp = malloc(sizeof(Item) * (size_t) (total > 0 ? total : 1));

	for (int p = 0; p < probes; p++)
	{
		int64_t		lo = ix->list_offsets[lists[p]],
					hi = ix->list_offsets[lists[p] + 1];

		for (int64_t r = lo; r < hi; r++)
		{
			items[n].d = q == NULL ? 0.0 :
				pgv_distance(ix->elem, ix->metric, ix->dim, (const char *) ix->rows + (size_t) r * rb, q);
			items[n].id = ix->ids ? ix->ids[r] : r;
			n++;
		}
	}
	merge_sort_items(items, tmp, n);	/* tuplesort_performsort (ivfscan.c:182) */
	if (cap > n)
		cap = n;
	for (int64_t i = 0; i < cap; i++)
	{
		out_ids[i] = items[i].id;
		out_dist[i] = items[i].d;
	}
	free(items);
	free(tmp);
	return n;
}

/* ivfflatgettuple first batch (src/ivfscan.c:360-414) */
int64_t
pgv_ivf_search(const PgvIvfIndex *ix, const void *q, int probes, int k, int64_t *out_ids, double *out_dist)
{
	int		   *lists;
	int			np;
	int64_t		n,
				total = 0;

	if (probes > ix->lists)
		probes = ix->lists;
	lists = malloc(sizeof(int) * (size_t) (probes > 0 ? probes : 1));
	np = pgv_ivf_scan_lists(ix, q, probes, lists, NULL);
	for (int p = 0; p < np; p++)
		total += ix->list_offsets[lists[p] + 1] - ix->list_offsets[lists[p]];
	n = pgv_ivf_scan_items(ix, q, lists, np, k > 0 ? k : total, out_ids, out_dist);
	if (k > 0)
		for (int64_t i = n; i < k; i++)
		{
			out_ids[i] = -1;
			out_dist[i] = INFINITY;
		}
	free(lists);
	return n;
}

void
pgv_ivf_search_batch(const PgvIvfIndex *ix, const void *queries, int64_t nq, int probes, int k, int threads, int64_t *out_ids, double *out_dist)
{
	size_t		rb = pgv_row_bytes(ix->elem, ix->dim);

	(void) threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
	for (int64_t i = 0; i < nq; i++)
		pgv_ivf_search(ix, (const char *) queries + (size_t) i * rb, probes, k,
					   out_ids + i * k, out_dist + i * k);
}

/* ------------------------------------------------------ AddTupleToSort */

/* argmin of proc-1 distance with strict < (src/ivfbuild.c:183-192) */
void
pgv_ivf_assign(int elem, int metric, int dim, const void *rows, int64_t n, const void *centers, int lists, int threads, int32_t *out_list)
{
	size_t		rb = pgv_row_bytes(elem, dim);

	(void) threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
	for (int64_t r = 0; r < n; r++)
	{
		double		minDistance = DBL_MAX;
		int			closestCenter = 0;

		for (int i = 0; i < lists; i++)
		{
			double		distance = pgv_distance(elem, metric, dim, (const char *) rows + (size_t) r * rb,
												(const char *) centers + (size_t) i * rb);

			if (distance < minDistance)
			{
				minDistance = distance;
				closestCenter = i;
			}
		}
		out_list[r] = closestCenter;
	}
}

/* ------------------------------------------------------------- k-means */

/* {Vector,Halfvec,Bit}SumCenter (src/ivfutils.c:341-370) */
static void
sum_center(int elem, int dim, const void *v, float *x)
{
	if (elem == PGV_VECTOR)
		for (int i = 0; i < dim; i++)
			x[i] += ((const float *) v)[i];
	else if (elem == PGV_HALFVEC)
		for (int i = 0; i < dim; i++)
			x[i] += pgv_half_to_float(((const uint16_t *) v)[i]);
	else
		for (int i = 0; i < dim; i++)
			x[i] += (float) ((((const uint8_t *) v)[i / 8] >> (7 - (i % 8))) & 1);
}

/* {Vector,Halfvec,Bit}UpdateCenter (src/ivfutils.c:301-339) */
static void
update_center(int elem, int dim, void *v, const float *x)
{
	if (elem == PGV_VECTOR)
		memcpy(v, x, sizeof(float) * (size_t) dim);
	else if (elem == PGV_HALFVEC)
		for (int i = 0; i < dim; i++)
			((uint16_t *) v)[i] = pgv_float_to_half(x[i]);
	else
	{
		uint8_t    *nx = v;

		memset(nx, 0, ((size_t) dim + 7) / 8);
		for (int i = 0; i < dim; i++)
			nx[i / 8] |= (uint8_t) ((x[i] > 0.5 ? 1 : 0) << (7 - (i % 8)));
	}
}

static int
kmeans_is_spherical(int kmeans_metric)
{
	return kmeans_metric == PGV_SPHERICAL;
}

/* ComputeNewCenters (src/ivfkmeans.c:179-236) */
static void
compute_new_centers(int elem, int dim, const void *samples, int64_t n, float *agg, void *newCenters,
					int k, int *centerCounts, const int32_t *closest, int spherical, Rng *rng)
{
	size_t		rb = pgv_row_bytes(elem, dim);

	for (int i = 0; i < k; i++)
	{
		float	   *x = agg + (size_t) i * dim;

		for (int j = 0; j < dim; j++)
			x[j] = 0.0f;
		centerCounts[i] = 0;
	}
	for (int64_t i = 0; i < n; i++)
		sum_center(elem, dim, (const char *) samples + (size_t) i * rb, agg + (size_t) closest[i] * dim);
	for (int64_t i = 0; i < n; i++)
		centerCounts[closest[i]] += 1;

	for (int i = 0; i < k; i++)
	{
		float	   *x = agg + (size_t) i * dim;

		if (centerCounts[i] > 0)
		{
			for (int j = 0; j < dim; j++)
				if (isinf(x[j]))
					x[j] = x[j] > 0 ? FLT_MAX : -FLT_MAX;
			for (int j = 0; j < dim; j++)
				x[j] /= (float) centerCounts[i];
		}
		else
		{
			/* empty cluster: uniform random coordinates (ivfkmeans.c:222-227) */
			for (int j = 0; j < dim; j++)
				x[j] = (float) rng_double(rng);
		}
	}
	for (int i = 0; i < k; i++)
		update_center(elem, dim, (char *) newCenters + (size_t) i * rb, agg + (size_t) i * dim);
	if (spherical)
	{
		/* NormCenters (ivfkmeans.c:96-105, 233-235) */
		void	   *tmp = malloc(rb);

		for (int i = 0; i < k; i++)
		{
			pgv_l2_normalize(elem, dim, (char *) newCenters + (size_t) i * rb, tmp);
			memcpy((char *) newCenters + (size_t) i * rb, tmp, rb);
		}
		free(tmp);
	}
}

/* InitCenters (src/ivfkmeans.c:23-91); lowerBound may be NULL */
static void
init_centers(int elem, int metric, int dim, const void *samples, int64_t n, void *centers, int k, float *lowerBound, Rng *rng)
{
	size_t		rb = pgv_row_bytes(elem, dim);
	float	   *weight = malloc(sizeof(float) * (size_t) n);

	memcpy(centers, (const char *) samples + (size_t) (rng_next(rng) % (uint64_t) n) * rb, rb);
	for (int64_t j = 0; j < n; j++)
		weight[j] = FLT_MAX;

	for (int i = 0; i < k; i++)
	{
		int64_t		j;
		double		sum = 0.0,
					choice;

		for (j = 0; j < n; j++)
		{
			double		distance = pgv_distance(elem, metric, dim, (const char *) samples + (size_t) j * rb,
												(const char *) centers + (size_t) i * rb);

			if (lowerBound)
				lowerBound[(size_t) j * k + i] = (float) distance;
			distance *= distance;
			if (distance < weight[j])
				weight[j] = (float) distance;
			sum += weight[j];
		}
		if (i + 1 == k)
			break;
		choice = sum * rng_double(rng);
		for (j = 0; j < n - 1; j++)
		{
			choice -= weight[j];
			if (choice <= 0)
				break;
		}
		memcpy((char *) centers + (size_t) (i + 1) * rb, (const char *) samples + (size_t) j * rb, rb);
	}
	free(weight);
}

/*
 * InitCenters (src/ivfkmeans.c:23-91) with the draws supplied by the caller: first = RandomInt() % numSamples
 * (:36), u[i] = RandomDouble() of round i (:78).  picked (optional) = chosen sample rows.  Same loop as init_centers.
 */
void
pgv_kmeans_pp_init_draws(int elem, int metric, int dim, const void *samples, int64_t n, void *centers, int k,
						 int64_t first, const double *u, int64_t *picked)
{
	size_t		rb = pgv_row_bytes(elem, dim);
	float	   *weight = malloc(sizeof(float) * (size_t) n);

	memcpy(centers, (const char *) samples + (size_t) first * rb, rb);
	if (picked)
		picked[0] = first;
	for (int64_t j = 0; j < n; j++)
		weight[j] = FLT_MAX;
	for (int i = 0; i + 1 < k; i++)
	{
		int64_t		j;
		double		sum = 0.0,
					choice;

		for (j = 0; j < n; j++)
		{
			double		distance = pgv_distance(elem, metric, dim, (const char *) samples + (size_t) j * rb,
												(const char *) centers + (size_t) i * rb);

			distance *= distance;
			if (distance < weight[j])
				weight[j] = (float) distance;
			sum += weight[j];
		}
		choice = sum * u[i];
		for (j = 0; j < n - 1; j++)
		{
			choice -= weight[j];
			if (choice <= 0)
				break;
		}
		memcpy((char *) centers + (size_t) (i + 1) * rb, (const char *) samples + (size_t) j * rb, rb);
		if (picked)
			picked[i + 1] = j;
	}
	free(weight);
}

void
pgv_kmeans_pp_init(int elem, int kmeans_metric, int dim, const void *samples, int64_t n, void *centers, int k, uint64_t seed)
{
	Rng			rng;

	rng_seed(&rng, seed);
	init_centers(elem, kmeans_metric, dim, samples, n, centers, k, NULL, &rng);
}

/*
 * ElkanKmeans (src/ivfkmeans.c:246-485), starting from the given centres:
 * the initial lowerBound fill that InitCenters produces as a side effect is
 * recomputed here for those centres (same values: d(x_j, c_i) as float).
 */
int
pgv_kmeans_elkan(int elem, int metric, int dim, const void *samples, int64_t n, void *centers, int k, int max_iter, uint64_t seed, int32_t *closest_out)
{
	size_t		rb = pgv_row_bytes(elem, dim);
	int			spherical = kmeans_is_spherical(metric);
	float	   *agg = malloc(sizeof(float) * (size_t) k * dim);
	int		   *centerCounts = malloc(sizeof(int) * (size_t) k);
	int32_t    *closest = malloc(sizeof(int32_t) * (size_t) n);
	float	   *lowerBound = malloc(sizeof(float) * (size_t) n * k);
	float	   *upperBound = malloc(sizeof(float) * (size_t) n);
	float	   *s = malloc(sizeof(float) * (size_t) k);
	float	   *halfcdist = malloc(sizeof(float) * (size_t) k * k);
	float	   *newcdist = malloc(sizeof(float) * (size_t) k);
	void	   *newCenters = malloc(rb * (size_t) k);
	Rng			rng;
	int			iteration;

#define SAMPLE(j) ((const char *) samples + (size_t) (j) * rb)
#define CENTER(c) ((char *) centers + (size_t) (c) * rb)
#define DIST(a, b) pgv_distance(elem, metric, dim, (a), (b))

	rng_seed(&rng, seed);
	if (max_iter <= 0 || max_iter > 500)
		max_iter = 500;			/* ivfkmeans.c:347 */

	/* lowerBound as left by InitCenters (ivfkmeans.c:62) */
#pragma omp parallel for schedule(static)
	for (int64_t j = 0; j < n; j++)
		for (int c = 0; c < k; c++)
			lowerBound[(size_t) j * k + c] = (float) DIST(SAMPLE(j), CENTER(c));

	/* initial assignment (ivfkmeans.c:324-344) */
	for (int64_t j = 0; j < n; j++)
	{
		float		minDistance = FLT_MAX;
		int			closestCenter = 0;

		for (int c = 0; c < k; c++)
		{
			float		distance = lowerBound[(size_t) j * k + c];

			if (distance < minDistance)
			{
				minDistance = distance;
				closestCenter = c;
			}
		}
		upperBound[j] = minDistance;
		closest[j] = closestCenter;
	}

	for (iteration = 0; iteration < max_iter; iteration++)
	{
		int			changes = 0;
		int			rjreset;

		/* Step 1 (ivfkmeans.c:356-367) */
		for (int j = 0; j < k; j++)
			for (int c = j + 1; c < k; c++)
			{
				float		distance = (float) (0.5 * DIST(CENTER(j), CENTER(c)));

				halfcdist[(size_t) j * k + c] = distance;
				halfcdist[(size_t) c * k + j] = distance;
			}
		/* s(c) (ivfkmeans.c:370-387) */
		for (int j = 0; j < k; j++)
		{
			float		minDistance = FLT_MAX;

			for (int c = 0; c < k; c++)
			{
				if (j == c)
					continue;
				if (halfcdist[(size_t) j * k + c] < minDistance)
					minDistance = halfcdist[(size_t) j * k + c];
			}
			s[j] = minDistance;
		}

		rjreset = iteration != 0;

		for (int64_t j = 0; j < n; j++)
		{
			int			rj;

			/* Step 2 */
			if (upperBound[j] <= s[closest[j]])
				continue;
			rj = rjreset;

			for (int c = 0; c < k; c++)
			{
				float		dxcx;

				/* Step 3 */
				if (c == closest[j])
					continue;
				if (upperBound[j] <= lowerBound[(size_t) j * k + c])
					continue;
				if (upperBound[j] <= halfcdist[(size_t) closest[j] * k + c])
					continue;

				/* Step 3a */
				if (rj)
				{
					dxcx = (float) DIST(SAMPLE(j), CENTER(closest[j]));
					lowerBound[(size_t) j * k + closest[j]] = dxcx;
					upperBound[j] = dxcx;
					rj = 0;
				}
				else
					dxcx = upperBound[j];

				/* Step 3b */
				if (dxcx > lowerBound[(size_t) j * k + c] || dxcx > halfcdist[(size_t) closest[j] * k + c])
				{
					float		dxc = (float) DIST(SAMPLE(j), CENTER(c));

					lowerBound[(size_t) j * k + c] = dxc;
					if (dxc < dxcx)
					{
						closest[j] = c;
						upperBound[j] = dxc;
						changes++;
					}
				}
			}
		}

		/* Step 4 */
		compute_new_centers(elem, dim, samples, n, agg, newCenters, k, centerCounts, closest, spherical, &rng);

		/* Step 5 */
		for (int j = 0; j < k; j++)
			newcdist[j] = (float) DIST(CENTER(j), (char *) newCenters + (size_t) j * rb);
		for (int64_t j = 0; j < n; j++)
			for (int c = 0; c < k; c++)
			{
				float		distance = lowerBound[(size_t) j * k + c] - newcdist[c];

				if (distance < 0)
					distance = 0;
				lowerBound[(size_t) j * k + c] = distance;
			}
		/* Step 6 */
		for (int64_t j = 0; j < n; j++)
			upperBound[j] += newcdist[closest[j]];
		/* Step 7 */
		memcpy(centers, newCenters, rb * (size_t) k);

		if (changes == 0 && iteration != 0)
		{
			iteration++;
			break;
		}
	}

	if (closest_out)
		memcpy(closest_out, closest, sizeof(int32_t) * (size_t) n);
	free(agg);
	free(centerCounts);
	free(closest);
	free(lowerBound);
	free(upperBound);
	free(s);
	free(halfcdist);
	free(newcdist);
	free(newCenters);
	return iteration;
#undef SAMPLE
#undef CENTER
#undef DIST
}

/*
 * Plain Lloyd with the reference's centre-update rules and stopping rule.  In
 * exact arithmetic it visits the same assignments as Elkan (Elkan only prunes
 * distance evaluations); this is the shape the GPU build executes, kept here
 * so the two can be compared iteration by iteration.
 */
int
pgv_kmeans_lloyd(int elem, int metric, int dim, const void *samples, int64_t n, void *centers, int k, int max_iter, uint64_t seed, int32_t *closest_out)
{
	size_t		rb = pgv_row_bytes(elem, dim);
	int			spherical = kmeans_is_spherical(metric);
	float	   *agg = malloc(sizeof(float) * (size_t) k * dim);
	int		   *centerCounts = malloc(sizeof(int) * (size_t) k);
	int32_t    *closest = malloc(sizeof(int32_t) * (size_t) n);
	void	   *newCenters = malloc(rb * (size_t) k);
	Rng			rng;
	int			iteration;

	rng_seed(&rng, seed);
	if (max_iter <= 0 || max_iter > 500)
		max_iter = 500;
	for (int64_t j = 0; j < n; j++)
		closest[j] = -1;

	for (iteration = 0; iteration < max_iter; iteration++)
	{
		int			changes = 0;

#pragma omp parallel for schedule(static) reduction(+:changes)
		for (int64_t j = 0; j < n; j++)
		{
			float		minDistance = FLT_MAX;
			int			best = 0;

			for (int c = 0; c < k; c++)
			{
				float		d = (float) pgv_distance(elem, metric, dim, (const char *) samples + (size_t) j * rb,
													 (char *) centers + (size_t) c * rb);

				if (d < minDistance)
				{
					minDistance = d;
					best = c;
				}
			}
			if (best != closest[j])
			{
				/* the reference counts a change only after the initial assignment */
				if (closest[j] >= 0)
					changes++;
				closest[j] = best;
			}
		}
		/*
		 * iteration 0 here = the reference's initial assignment + its first
		 * loop pass (which cannot change anything before centres move).
		 */
		compute_new_centers(elem, dim, samples, n, agg, newCenters, k, centerCounts, closest, spherical, &rng);
		memcpy(centers, newCenters, rb * (size_t) k);
		if (changes == 0 && iteration != 0)
		{
			iteration++;
			break;
		}
	}
	if (closest_out)
		memcpy(closest_out, closest, sizeof(int32_t) * (size_t) n);
	free(agg);
	free(centerCounts);
	free(closest);
	free(newCenters);
	return iteration;
}
