/*
 * pgv_sparse.c -- CPU oracle for the sparsevec distance functions.
 *
 * TEST INFRASTRUCTURE ONLY (see pgv_oracle.h).  Restatement of
 *   SparsevecL2SquaredDistance   src/sparsevec.c:826-867
 *   sparsevec_l2_distance / _l2_squared_distance   src/sparsevec.c:872-900
 *   SparsevecInnerProduct        src/sparsevec.c:905-934
 *   sparsevec_inner_product / _negative_inner_product   src/sparsevec.c:939-965
 *   sparsevec_cosine_distance    src/sparsevec.c:970-1010
 *   sparsevec_l1_distance        src/sparsevec.c:1015-1057
 *   sparsevec_l2_norm            src/sparsevec.c:1062-1077
 *   sparsevec_l2_normalize       src/sparsevec.c:1082-1150
 *
 * A sparsevec is (dim, nnz, indices[nnz] ascending and 0-based, values[nnz]) -- src/sparsevec.h:21-32.
 *
 * The reference walks `a` and, for each of its entries, advances a cursor over `b`; the terms reach the fp32
 * accumulator in ascending index order (entries of b below a[i], then the a[i] term, and the tail of b last).  This
 * file states the same thing as one two-cursor merge: identical terms in the identical order, so the fp32 sums are
 * those of the reference.  Parity is pinned by the known answers of test/expected/sparsevec.out
 * (tests/golden/distance_kat.json, type "sparsevec").
 */
#include <math.h>
#include <stdint.h>

#include "pgv_oracle.h"

/* which side(s) hold the next index of the merge */
enum { ONLY_A, ONLY_B, BOTH, DONE };

typedef struct
{
	int			i, j;
}			Cursor;

static inline int
merge_next(Cursor *c, int na, const int32_t *ai, int nb, const int32_t *bi)
{
	if (c->i >= na && c->j >= nb)
		return DONE;
	if (c->j >= nb)
		return ONLY_A;
	if (c->i >= na)
		return ONLY_B;
	if (ai[c->i] == bi[c->j])
		return BOTH;
	return ai[c->i] < bi[c->j] ? ONLY_A : ONLY_B;
}

/* src/sparsevec.c:826-867 */
static float
sparse_l2_squared(int na, const int32_t *ai, const float *ax, int nb, const int32_t *bi, const float *bx)
{
	float		d = 0.0f;
	Cursor		c = {0, 0};

	for (;;)
	{
		int			w = merge_next(&c, na, ai, nb, bi);

		if (w == DONE)
			break;
		if (w == BOTH)
		{
			float		t = ax[c.i] - bx[c.j];

			d += t * t;
			c.i++;
			c.j++;
		}
		else if (w == ONLY_A)
		{
			d += ax[c.i] * ax[c.i];
			c.i++;
		}
		else
		{
			d += bx[c.j] * bx[c.j];
			c.j++;
		}
	}
	return d;
}

/* src/sparsevec.c:905-934: only matching indices contribute */
static float
sparse_inner_product(int na, const int32_t *ai, const float *ax, int nb, const int32_t *bi, const float *bx)
{
	float		d = 0.0f;
	Cursor		c = {0, 0};

	for (;;)
	{
		int			w = merge_next(&c, na, ai, nb, bi);

		if (w == DONE || c.i >= na || c.j >= nb)
			break;
		if (w == BOTH)
		{
			d += ax[c.i] * bx[c.j];
			c.i++;
			c.j++;
		}
		else if (w == ONLY_A)
			c.i++;
		else
			c.j++;
	}
	return d;
}

/* src/sparsevec.c:1015-1057 */
static float
sparse_l1(int na, const int32_t *ai, const float *ax, int nb, const int32_t *bi, const float *bx)
{
	float		d = 0.0f;
	Cursor		c = {0, 0};

	for (;;)
	{
		int			w = merge_next(&c, na, ai, nb, bi);

		if (w == DONE)
			break;
		if (w == BOTH)
		{
			d += fabsf(ax[c.i] - bx[c.j]);
			c.i++;
			c.j++;
		}
		else if (w == ONLY_A)
			d += fabsf(ax[c.i++]);
		else
			d += fabsf(bx[c.j++]);
	}
	return d;
}

static float
sum_squares_f32(int n, const float *x)
{
	float		s = 0.0f;

	for (int i = 0; i < n; i++)
		s += x[i] * x[i];
	return s;
}

/* the float8 each SQL function returns */
double
pgv_sparse_distance(int metric, int na, const int32_t *ai, const float *ax, int nb, const int32_t *bi, const float *bx)
{
	switch (metric)
	{
		case PGV_L2_SQUARED:
			return (double) sparse_l2_squared(na, ai, ax, nb, bi, bx);
		case PGV_L2:
			return sqrt((double) sparse_l2_squared(na, ai, ax, nb, bi, bx));
		case PGV_IP:
			return (double) sparse_inner_product(na, ai, ax, nb, bi, bx);
		case PGV_NEG_IP:
			return (double) -sparse_inner_product(na, ai, ax, nb, bi, bx);
		case PGV_L1:
			return (double) sparse_l1(na, ai, ax, nb, bi, bx);
		case PGV_COSINE:
			{
				/* src/sparsevec.c:970-1010: fp32 norms, sqrt(a * b) in double, clamp to [-1, 1] */
				double		sim = sparse_inner_product(na, ai, ax, nb, bi, bx);
				float		norma = sum_squares_f32(na, ax);
				float		normb = sum_squares_f32(nb, bx);

				sim /= sqrt((double) norma * (double) normb);
				if (sim > 1)
					sim = 1.0;
				else if (sim < -1)
					sim = -1.0;
				return 1.0 - sim;
			}
		default:
			return NAN;
	}
}

/* the same value from fp64 sums ("truth" for tolerances) */
double
pgv_sparse_distance_f64(int metric, int na, const int32_t *ai, const float *ax, int nb, const int32_t *bi, const float *bx)
{
	double		l2 = 0, ip = 0, l1 = 0, sa = 0, sb = 0;
	Cursor		c = {0, 0};

	for (int i = 0; i < na; i++)
		sa += (double) ax[i] * ax[i];
	for (int j = 0; j < nb; j++)
		sb += (double) bx[j] * bx[j];
	for (;;)
	{
		int			w = merge_next(&c, na, ai, nb, bi);
		double		x, y;

		if (w == DONE)
			break;
		x = (w == ONLY_B) ? 0.0 : ax[c.i];
		y = (w == ONLY_A) ? 0.0 : bx[c.j];
		l2 += (x - y) * (x - y);
		l1 += fabs(x - y);
		ip += x * y;
		if (w != ONLY_B)
			c.i++;
		if (w != ONLY_A)
			c.j++;
	}
	switch (metric)
	{
		case PGV_L2_SQUARED:
			return l2;
		case PGV_L2:
			return sqrt(l2);
		case PGV_IP:
			return ip;
		case PGV_NEG_IP:
			return -ip;
		case PGV_L1:
			return l1;
		case PGV_COSINE:
			{
				double		sim = ip / sqrt(sa * sb);

				if (sim > 1)
					sim = 1.0;
				else if (sim < -1)
					sim = -1.0;
				return 1.0 - sim;
			}
		default:
			return NAN;
	}
}

/* src/sparsevec.c:1062-1077: fp64 sum of squares */
double
pgv_sparse_l2_norm(int nnz, const float *x)
{
	double		s = 0.0;

	for (int i = 0; i < nnz; i++)
		s += (double) x[i] * (double) x[i];
	return sqrt(s);
}

/*
 * src/sparsevec.c:1082-1150: divide by the fp64 norm, narrow to float; an infinite quotient is
 * float_overflow_error() (returns -1 here); quotients that round to zero are dropped from the result.
 * Returns the nnz of the result; a zero vector stays as it is.
 */
int
pgv_sparse_l2_normalize(int nnz, const int32_t *idx, const float *x, int32_t *out_idx, float *out_x)
{
	double		norm = pgv_sparse_l2_norm(nnz, x);
	int			kept = 0;

	if (!(norm > 0))
	{
		/* "Return zero vector for zero norm": InitSparseVector(dim, nnz) zero-filled; with nnz == 0 for every
		 * valid input (stored values are never zero), so the result is empty */
		return 0;
	}
	for (int i = 0; i < nnz; i++)
	{
		float		q = (float) (x[i] / norm);

		if (isinf(q))
			return -1;
		if (q == 0)
			continue;
		out_idx[kept] = idx[i];
		out_x[kept] = q;
		kept++;
	}
	return kept;
}

/* one query against n rows held as CSR (row r = entries row_off[r] .. row_off[r+1]); a = the row, b = the query,
 * the argument order of FunctionCall2Coll(procinfo, collation, value, q) in the index code (src/hnswutils.c:527) */
void
pgv_sparse_distance_batch(int metric, int q_nnz, const int32_t *q_idx, const float *q_val, int64_t n,
						  const int64_t *row_off, const int32_t *idx, const float *val, double *out)
{
#pragma omp parallel for schedule(static)
	for (int64_t r = 0; r < n; r++)
	{
		int64_t		b = row_off[r];

		out[r] = pgv_sparse_distance(metric, (int) (row_off[r + 1] - b), idx + b, val + b, q_nnz, q_idx, q_val);
	}
}
