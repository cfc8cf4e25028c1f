/*
 * ref_shim.c -- thin exports over the reference's own halfutils.c / bitutils.c,
 * which the Makefile compiles UNMODIFIED from /root/reference/src into
 * oracle/_ref/libpgvref.so.  TEST INFRASTRUCTURE ONLY.  No reference source is
 * copied into this repository; this file only declares the reference's public
 * function pointers (src/halfutils.h:13-16, src/bitutils.h) and calls them.
 */
#include "postgres.h"
#include "halfutils.h"
#include "bitutils.h"

const uint8 pg_number_of_ones[256] = {
#define B2(n) n, n + 1, n + 1, n + 2
#define B4(n) B2(n), B2(n + 1), B2(n + 1), B2(n + 2)
#define B6(n) B4(n), B4(n + 1), B4(n + 1), B4(n + 2)
	B6(0), B6(1), B6(1), B6(2)
};

static int	inited = 0;

static void
ref_init(void)
{
	if (!inited)
	{
		HalfvecInit();			/* src/halfutils.c:278-300 */
		BitvecInit();			/* src/bitutils.c:207-224 */
		inited = 1;
	}
}

float ref_half_l2sq(int dim, uint16 *a, uint16 *b) { ref_init(); return HalfvecL2SquaredDistance(dim, (half *) a, (half *) b); }
float ref_half_ip(int dim, uint16 *a, uint16 *b) { ref_init(); return HalfvecInnerProduct(dim, (half *) a, (half *) b); }
double ref_half_cos(int dim, uint16 *a, uint16 *b) { ref_init(); return HalfvecCosineSimilarity(dim, (half *) a, (half *) b); }
float ref_half_l1(int dim, uint16 *a, uint16 *b) { ref_init(); return HalfvecL1Distance(dim, (half *) a, (half *) b); }
uint64 ref_bit_hamming(uint32 bytes, unsigned char *a, unsigned char *b) { ref_init(); return BitHammingDistance(bytes, a, b, 0); }
double ref_bit_jaccard(uint32 bytes, unsigned char *a, unsigned char *b) { ref_init(); return BitJaccardDistance(bytes, a, b, 0, 0, 0); }
float ref_half_to_float(uint16 h) { half x; memcpy(&x, &h, 2); return HalfToFloat4(x); }
uint16 ref_float_to_half(float f) { half x = Float4ToHalfUnchecked(f); uint16 r; memcpy(&r, &x, 2); return r; }

/* one-vs-many loops so the CPU baseline can time the reference's own kernels */
void ref_half_batch(int metric, int dim, uint16 *q, uint16 *rows, long n, float *out)
{
	ref_init();
	for (long i = 0; i < n; i++)
	{
		half	   *r = (half *) (rows + (size_t) i * dim);

		out[i] = metric == 0 ? HalfvecL2SquaredDistance(dim, r, (half *) q) :
			metric == 1 ? -HalfvecInnerProduct(dim, r, (half *) q) : HalfvecL1Distance(dim, r, (half *) q);
	}
}

void ref_bit_batch(int metric, uint32 bytes, unsigned char *q, unsigned char *rows, long n, double *out)
{
	ref_init();
	for (long i = 0; i < n; i++)
	{
		unsigned char *r = rows + (size_t) i * bytes;

		out[i] = metric == 4 ? (double) BitHammingDistance(bytes, r, q, 0) : BitJaccardDistance(bytes, r, q, 0, 0, 0);
	}
}
