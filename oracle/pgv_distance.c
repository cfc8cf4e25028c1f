/*
 * pgv_distance.c -- CPU oracle: per-pair distance kernels and their SQL-level
 * epilogues.  TEST INFRASTRUCTURE ONLY (see pgv_oracle.h).
 *
 * Build with the reference's flag set (Makefile:15,38 of the reference):
 *   -O2 -march=native -ftree-vectorize -fassociative-math -fno-signed-zeros
 *   -fno-trapping-math -ffp-contract=fast
 * so the fp32 loops reassociate / contract the way the reference's do.
 */
#include "pgv_oracle.h"

#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>

/* ------------------------------------------------------------------ half */

/* HalfToFloat4 (src/halfutils.h:62-141): exact widening of IEEE binary16 */
float
pgv_half_to_float(uint16_t h)
{
	uint32_t	sign = ((uint32_t) h & 0x8000u) << 16;
	int			e = (h >> 10) & 0x1f;
	uint32_t	mant = h & 0x3ffu;
	uint32_t	bits;
	float		f;

	if (e == 31)
		bits = sign | (mant ? 0x7fc00000u : 0x7f800000u);
	else if (e == 0)
	{
		if (mant == 0)
			bits = sign;
		else
		{
			/* subnormal half: renormalise into a float exponent */
			int			ex = -14;

			while (!(mant & 0x400u))
			{
				mant <<= 1;
				ex--;
			}
			mant &= 0x3ffu;
			bits = sign | ((uint32_t) (ex + 127) << 23) | (mant << 13);
		}
	}
	else
		bits = sign | ((uint32_t) (e - 15 + 127) << 23) | (mant << 13);

	memcpy(&f, &bits, 4);
	return f;
}

/* Float4ToHalfUnchecked (src/halfutils.h:146-239): round-to-nearest-even, overflow -> Inf */
uint16_t
pgv_float_to_half(float f)
{
	uint32_t	bits;
	uint32_t	sign;
	int			e;
	uint32_t	mant;

	memcpy(&bits, &f, 4);
	sign = (bits >> 16) & 0x8000u;
	e = (int) ((bits >> 23) & 0xff);
	mant = bits & 0x7fffffu;

	if (e == 255)
		return (uint16_t) (sign | 0x7c00u | (mant ? 0x200u : 0));	/* Inf / NaN */

	e = e - 127 + 15;

	if (e >= 31)
		return (uint16_t) (sign | 0x7c00u);	/* overflow -> Inf */

	if (e <= 0)
	{
		/* result is subnormal (or zero) in half */
		uint32_t	m;
		int			shift;
		uint32_t	halfm,
					rem,
					halfway;

		if (e < -10)
			return (uint16_t) sign; /* below half the smallest subnormal */
		m = mant | 0x800000u;
		shift = 14 - e;			/* 14..24 */
		halfm = m >> shift;
		rem = m & ((1u << shift) - 1);
		halfway = 1u << (shift - 1);
		if (rem > halfway || (rem == halfway && (halfm & 1)))
			halfm++;			/* may carry into the smallest normal: still correct encoding */
		return (uint16_t) (sign | halfm);
	}
	else
	{
		uint32_t	halfm = mant >> 13;
		uint32_t	rem = mant & 0x1fffu;
		uint32_t	out = ((uint32_t) e << 10) | halfm;

		if (rem > 0x1000u || (rem == 0x1000u && (halfm & 1)))
			out++;				/* carries propagate into the exponent; 0x7c00 = Inf on overflow */
		return (uint16_t) (sign | out);
	}
}

/* ------------------------------------------------------------ fp32 loops */

/* VectorL2SquaredDistance (src/vector.c:560-574) */
static float
vec_l2sq(int dim, const float *ax, const float *bx)
{
	float		distance = 0.0f;

	for (int i = 0; i < dim; i++)
	{
		float		diff = ax[i] - bx[i];

		distance += diff * diff;
	}
	return distance;
}

/* VectorInnerProduct (src/vector.c:607-617) */
static float
vec_ip(int dim, const float *ax, const float *bx)
{
	float		distance = 0.0f;

	for (int i = 0; i < dim; i++)
		distance += ax[i] * bx[i];
	return distance;
}

/* VectorCosineSimilarity (src/vector.c:649-666) */
static double
vec_cos(int dim, const float *ax, const float *bx)
{
	float		similarity = 0.0f,
				norma = 0.0f,
				normb = 0.0f;

	for (int i = 0; i < dim; i++)
	{
		similarity += ax[i] * bx[i];
		norma += ax[i] * ax[i];
		normb += bx[i] * bx[i];
	}
	return (double) similarity / sqrt((double) norma * (double) normb);
}

/* VectorL1Distance (src/vector.c:725-735) */
static float
vec_l1(int dim, const float *ax, const float *bx)
{
	float		distance = 0.0f;

	for (int i = 0; i < dim; i++)
		distance += fabsf(ax[i] - bx[i]);
	return distance;
}

/* ------------------------------------------------------------ fp16 loops */
/* Halfvec*Default (src/halfutils.c:29-43, 81-92, 124-145, 197-209): widen, fp32 accumulate */

#if defined(__F16C__) && defined(__AVX__) && defined(__FMA__)
#include <immintrin.h>
#define PGV_HAVE_F16C 1
/*
 * HalfvecL2SquaredDistanceF16c / HalfvecInnerProductF16c (src/halfutils.c:46-78, 94-121), the variants
 * the reference dispatches to on x86-64 with F16C: 8 halves per step widened with vcvtph2ps, fp32 FMA
 * into 8 lanes, lanes summed s[0]+...+s[7], scalar tail.
 */
static float
half_l2sq_f16c(int dim, const uint16_t *ax, const uint16_t *bx)
{
	float		s[8];
	float		distance;
	int			i;
	int			count = (dim / 8) * 8;
	__m256		dist = _mm256_setzero_ps();

	for (i = 0; i < count; i += 8)
	{
		__m256		a = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *) (ax + i)));
		__m256		b = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *) (bx + i)));
		__m256		diff = _mm256_sub_ps(a, b);

		dist = _mm256_fmadd_ps(diff, diff, dist);
	}
	_mm256_storeu_ps(s, dist);
	distance = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
	for (; i < dim; i++)
	{
		float		diff = pgv_half_to_float(ax[i]) - pgv_half_to_float(bx[i]);

		distance += diff * diff;
	}
	return distance;
}

static float
half_ip_f16c(int dim, const uint16_t *ax, const uint16_t *bx)
{
	float		s[8];
	float		distance;
	int			i;
	int			count = (dim / 8) * 8;
	__m256		dist = _mm256_setzero_ps();

	for (i = 0; i < count; i += 8)
	{
		__m256		a = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *) (ax + i)));
		__m256		b = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *) (bx + i)));

		dist = _mm256_fmadd_ps(a, b, dist);
	}
	_mm256_storeu_ps(s, dist);
	distance = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
	for (; i < dim; i++)
		distance += pgv_half_to_float(ax[i]) * pgv_half_to_float(bx[i]);
	return distance;
}
#endif

static float
half_l2sq(int dim, const uint16_t *ax, const uint16_t *bx)
{
	float		distance = 0.0f;

#ifdef PGV_HAVE_F16C
	return half_l2sq_f16c(dim, ax, bx);
#endif
	for (int i = 0; i < dim; i++)
	{
		float		diff = pgv_half_to_float(ax[i]) - pgv_half_to_float(bx[i]);

		distance += diff * diff;
	}
	return distance;
}

static float
half_ip(int dim, const uint16_t *ax, const uint16_t *bx)
{
	float		distance = 0.0f;

#ifdef PGV_HAVE_F16C
	return half_ip_f16c(dim, ax, bx);
#endif
	for (int i = 0; i < dim; i++)
		distance += pgv_half_to_float(ax[i]) * pgv_half_to_float(bx[i]);
	return distance;
}

static double
half_cos(int dim, const uint16_t *ax, const uint16_t *bx)
{
	float		similarity = 0.0f,
				norma = 0.0f,
				normb = 0.0f;

	for (int i = 0; i < dim; i++)
	{
		float		a = pgv_half_to_float(ax[i]);
		float		b = pgv_half_to_float(bx[i]);

		similarity += a * b;
		norma += a * a;
		normb += b * b;
	}
	return (double) similarity / sqrt((double) norma * (double) normb);
}

static float
half_l1(int dim, const uint16_t *ax, const uint16_t *bx)
{
	float		distance = 0.0f;

	for (int i = 0; i < dim; i++)
		distance += fabsf(pgv_half_to_float(ax[i]) - pgv_half_to_float(bx[i]));
	return distance;
}

/* -------------------------------------------------------------- bit loops */

static int
byte_ones(unsigned v)
{
	return __builtin_popcount(v & 0xffu);
}

/* BitHammingDistanceDefault (src/bitutils.c:49-73): 8-byte words then byte tail */
static uint64_t
bit_hamming(uint32_t bytes, const unsigned char *ax, const unsigned char *bx)
{
	uint64_t	distance = 0;

	for (; bytes >= 8; bytes -= 8, ax += 8, bx += 8)
	{
		uint64_t	a,
					b;

		memcpy(&a, ax, 8);
		memcpy(&b, bx, 8);
		distance += (uint64_t) __builtin_popcountll(a ^ b);
	}
	for (uint32_t i = 0; i < bytes; i++)
		distance += byte_ones(ax[i] ^ bx[i]);
	return distance;
}

/* BitJaccardDistanceDefault (src/bitutils.c:98-131) */
static double
bit_jaccard(uint32_t bytes, const unsigned char *ax, const unsigned char *bx)
{
	uint64_t	ab = 0,
				aa = 0,
				bb = 0;

	for (; bytes >= 8; bytes -= 8, ax += 8, bx += 8)
	{
		uint64_t	a,
					b;

		memcpy(&a, ax, 8);
		memcpy(&b, bx, 8);
		ab += (uint64_t) __builtin_popcountll(a & b);
		aa += (uint64_t) __builtin_popcountll(a);
		bb += (uint64_t) __builtin_popcountll(b);
	}
	for (uint32_t i = 0; i < bytes; i++)
	{
		ab += byte_ones(ax[i] & bx[i]);
		aa += byte_ones(ax[i]);
		bb += byte_ones(bx[i]);
	}
	if (ab == 0)
		return 1;
	return 1 - ((double) ab / (double) (aa + bb - ab));
}

/* ------------------------------------------------------------- epilogues */

static double
clamp_unit(double s)
{
	/* src/vector.c:690-693 (NaN falls through both tests) */
	if (s > 1)
		s = 1.0;
	else if (s < -1)
		s = -1.0;
	return s;
}

size_t
pgv_row_bytes(int elem, int dim)
{
	switch (elem)
	{
		case PGV_VECTOR:
			return (size_t) dim * 4;
		case PGV_HALFVEC:
			return (size_t) dim * 2;
		default:
			return ((size_t) dim + 7) / 8;
	}
}

double
pgv_distance(int elem, int metric, int dim, const void *a, const void *b)
{
	if (elem == PGV_BIT)
	{
		uint32_t	bytes = (uint32_t) ((dim + 7) / 8);

		if (metric == PGV_HAMMING)
			return (double) bit_hamming(bytes, a, b);	/* src/bitvec.c:45-55 */
		if (metric == PGV_JACCARD)
			return bit_jaccard(bytes, a, b);	/* src/bitvec.c:60-70 */
		return NAN;
	}

	{
		int			h = elem == PGV_HALFVEC;
		const float *fa = a,
				   *fb = b;
		const uint16_t *ha = a,
				   *hb = b;
		double		d;

		switch (metric)
		{
			case PGV_L2_SQUARED:	/* vector.c:595-605, halfvec.c:575-585 */
				return (double) (h ? half_l2sq(dim, ha, hb) : vec_l2sq(dim, fa, fb));
			case PGV_L2:		/* vector.c:579-589 */
				return sqrt((double) (h ? half_l2sq(dim, ha, hb) : vec_l2sq(dim, fa, fb)));
			case PGV_IP:		/* vector.c:622-632 */
				return (double) (h ? half_ip(dim, ha, hb) : vec_ip(dim, fa, fb));
			case PGV_NEG_IP:	/* vector.c:637-647 */
				return (double) -(h ? half_ip(dim, ha, hb) : vec_ip(dim, fa, fb));
			case PGV_COSINE:	/* vector.c:671-696 */
				return 1.0 - clamp_unit(h ? half_cos(dim, ha, hb) : vec_cos(dim, fa, fb));
			case PGV_L1:		/* vector.c:740-750 */
				return (double) (h ? half_l1(dim, ha, hb) : vec_l1(dim, fa, fb));
			case PGV_SPHERICAL: /* vector.c:703-722 */
				d = clamp_unit((double) (h ? half_ip(dim, ha, hb) : vec_ip(dim, fa, fb)));
				return acos(d) / M_PI;
			default:
				return NAN;
		}
	}
}

/* volatile-free fp64 truth; compiled in the same TU but every sum is double */
double
pgv_distance_f64(int elem, int metric, int dim, const void *a, const void *b)
{
	double		s = 0,
				na = 0,
				nb = 0;

	if (elem == PGV_BIT)
		return pgv_distance(elem, metric, dim, a, b);	/* integer metrics are exact */

	for (int i = 0; i < dim; i++)
	{
		double		x = elem == PGV_HALFVEC ? (double) pgv_half_to_float(((const uint16_t *) a)[i]) : (double) ((const float *) a)[i];
		double		y = elem == PGV_HALFVEC ? (double) pgv_half_to_float(((const uint16_t *) b)[i]) : (double) ((const float *) b)[i];

		switch (metric)
		{
			case PGV_L2_SQUARED:
			case PGV_L2:
				s += (x - y) * (x - y);
				break;
			case PGV_L1:
				s += fabs(x - y);
				break;
			default:
				s += x * y;
				na += x * x;
				nb += y * y;
		}
	}
	switch (metric)
	{
		case PGV_L2:
			return sqrt(s);
		case PGV_NEG_IP:
			return -s;
		case PGV_COSINE:
			return 1.0 - clamp_unit(s / sqrt(na * nb));
		case PGV_SPHERICAL:
			return acos(clamp_unit(s)) / M_PI;
		default:
			return s;
	}
}

/* vector_norm (src/vector.c:767-780), halfvec_l2_norm (src/halfvec.c:703-720): fp64 accumulate */
double
pgv_norm(int elem, int dim, const void *a)
{
	double		norm = 0.0;

	for (int i = 0; i < dim; i++)
	{
		double		x = elem == PGV_HALFVEC ? (double) pgv_half_to_float(((const uint16_t *) a)[i]) : (double) ((const float *) a)[i];

		norm += x * x;
	}
	return sqrt(norm);
}

/* l2_normalize (src/vector.c:785-819), halfvec_l2_normalize (src/halfvec.c:725-759) */
int
pgv_l2_normalize(int elem, int dim, const void *a, void *out)
{
	double		norm = pgv_norm(elem, dim, a);
	int			overflow = 0;

	if (elem == PGV_HALFVEC)
	{
		const uint16_t *ax = a;
		uint16_t   *rx = out;

		for (int i = 0; i < dim; i++)
			rx[i] = 0;
		if (norm > 0)
			for (int i = 0; i < dim; i++)
			{
				/* quotient in double, narrowed to float, then RNE to half */
				rx[i] = pgv_float_to_half((float) (pgv_half_to_float(ax[i]) / norm));
				if ((rx[i] & 0x7fffu) == 0x7c00u)
					overflow = 1;
			}
	}
	else
	{
		const float *ax = a;
		float	   *rx = out;

		for (int i = 0; i < dim; i++)
			rx[i] = 0;
		if (norm > 0)
			for (int i = 0; i < dim; i++)
			{
				rx[i] = (float) (ax[i] / norm);
				if (isinf(rx[i]))
					overflow = 1;
			}
	}
	return overflow ? -1 : 0;
}

/* binary_quantize (src/vector.c:952-978, src/halfvec.c twin): bit i = x[i] > 0, MSB first */
void
pgv_binary_quantize(int elem, int dim, const void *a, uint8_t *out)
{
	memset(out, 0, ((size_t) dim + 7) / 8);
	for (int i = 0; i < dim; i++)
	{
		float		x = elem == PGV_HALFVEC ? pgv_half_to_float(((const uint16_t *) a)[i]) : ((const float *) a)[i];

		out[i / 8] |= (uint8_t) ((x > 0 ? 1 : 0) << (7 - (i % 8)));
	}
}

void
pgv_distance_batch(int elem, int metric, int dim, const void *q, const void *rows, int64_t n, double *out)
{
	size_t		rb = pgv_row_bytes(elem, dim);

	for (int64_t i = 0; i < n; i++)
		out[i] = pgv_distance(elem, metric, dim, (const char *) rows + (size_t) i * rb, q);
}

typedef struct
{
	double		d;
	int64_t		id;
}			DistId;

static int
cmp_distid(const void *pa, const void *pb)
{
	const DistId *a = pa,
			   *b = pb;

	/* float8 btree order: NaN sorts after everything (PG float8_cmp_internal) */
	int			an = isnan(a->d),
				bn = isnan(b->d);

	if (an || bn)
	{
		if (an != bn)
			return an ? 1 : -1;
	}
	else if (a->d != b->d)
		return a->d < b->d ? -1 : 1;
	return a->id < b->id ? -1 : (a->id > b->id);
}

/* SeqScan -> operator -> Sort/top-N. ties resolved by row id (deterministic choice) */
void
pgv_exact_topk(int elem, int metric, int dim, const void *q, const void *rows, int64_t n, int k, int64_t *out_ids, double *out_dist)
{
	DistId	   *v = malloc(sizeof(DistId) * (size_t) (n > 0 ? n : 1));
	size_t		rb = pgv_row_bytes(elem, dim);

	for (int64_t i = 0; i < n; i++)
	{
		v[i].d = pgv_distance(elem, metric, dim, (const char *) rows + (size_t) i * rb, q);
		v[i].id = i;
	}
	qsort(v, (size_t) n, sizeof(DistId), cmp_distid);
	for (int i = 0; i < k; i++)
	{
		out_ids[i] = i < n ? v[i].id : -1;
		out_dist[i] = i < n ? v[i].d : INFINITY;
	}
	free(v);
}
