/* harness_common.h -- test harness around the extension glue (pgvector_b200/ext) over synthesised index pages.
 * TEST INFRASTRUCTURE: compiled against pgstub + the reference's headers, linked with pgstub_runtime.c and either
 * the real libvecb200.so (GPU tests) or mock_abi.c (CPU tests: the C ABI backed by the oracle). */
#ifndef HARNESS_COMMON_H
#define HARNESS_COMMON_H
#include "postgres.h"
#include "access/genam.h"
#include "access/relscan.h"
#include "utils/rel.h"

extern const char *pgstub_last_error(void);

/* run `body` with an error trap: returns 0, or -1 with the message in pgstub_last_error() */
#define H_TRAP(...) \
	do { \
		volatile int h_rc__ = 0; \
		PG_TRY(); \
		{ __VA_ARGS__; } \
		PG_CATCH(); \
		{ h_rc__ = -1; } \
		PG_END_TRY(); \
		return h_rc__; \
	} while (0)

typedef struct HRelation
{
	RelationData rel;
	TupleDescData desc;
	FmgrInfo	procs[8];
	Oid			collation;
}			HRelation;

/* distance wrappers of the extension: only their addresses matter to the glue */
extern Datum vector_l2_squared_distance(PG_FUNCTION_ARGS);
extern Datum vector_negative_inner_product(PG_FUNCTION_ARGS);
extern Datum l1_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_l2_squared_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_negative_inner_product(PG_FUNCTION_ARGS);
extern Datum halfvec_l1_distance(PG_FUNCTION_ARGS);
extern Datum hamming_distance(PG_FUNCTION_ARGS);
extern Datum jaccard_distance(PG_FUNCTION_ARGS);
extern Datum l2_distance(PG_FUNCTION_ARGS);
extern Datum vector_spherical_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_l2_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_spherical_distance(PG_FUNCTION_ARGS);

extern HRelation *h_open(char *pages, uint32 nblocks, Oid relid, int dim, int proc1, int proc3);
extern void *h_make_datum(int elem, int dim, const void *payload, int short_header);
#endif
