/*
 * vb_ivfflat_build.c -- GPU bodies for the two dense phases of CREATE INDEX ... USING ivfflat:
 *
 *   VbIvfflatKmeans  replaces ElkanKmeans + InitCenters (src/ivfkmeans.c:23-91, 246-485)
 *                    inside IvfflatKmeans (src/ivfkmeans.c:553-570)
 *   VbAssignFlush    replaces the per-row centre loop of AddTupleToSort (src/ivfbuild.c:183-192)
 *                    for a batch of rows collected by the build callback
 *
 * Sampling, the sort by list and the page writer (src/ivfbuild.c:56-156, 271-331) stay as they are.
 */
#include "postgres.h"

#include "access/genam.h"
#include "executor/tuptable.h"
#include "utils/memutils.h"
#include "utils/rel.h"
#include "utils/tuplesort.h"
#include "utils/varbit.h"

#include "halfvec.h"
#include "ivfflat.h"
#include "vector.h"

#include "vb_glue.h"

extern Datum l2_distance(PG_FUNCTION_ARGS);
extern Datum vector_spherical_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_l2_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_spherical_distance(PG_FUNCTION_ARGS);
extern Datum hamming_distance(PG_FUNCTION_ARGS);

/* opclass proc 3 (IVFFLAT_KMEANS_DISTANCE_PROC, src/ivfflat.h:42) -> k-means metric of the C ABI */
static int
VbKmeansMetricFromProc(FmgrInfo *procinfo, int *elem)
{
	PGFunction	fn = procinfo->fn_addr;

	if (fn == l2_distance)
	{
		*elem = VB_VECTOR;
		return VB_L2;
	}
	if (fn == vector_spherical_distance)
	{
		*elem = VB_VECTOR;
		return VB_SPHERICAL;
	}
	if (fn == halfvec_l2_distance)
	{
		*elem = VB_HALFVEC;
		return VB_L2;
	}
	if (fn == halfvec_spherical_distance)
	{
		*elem = VB_HALFVEC;
		return VB_SPHERICAL;
	}
	if (fn == hamming_distance)
	{
		*elem = VB_BIT;
		return VB_HAMMING;
	}
	elog(ERROR, "vecb200: unsupported k-means distance function");
	return -1;
}

static inline char *
VbItemPayload(int elem, Pointer item)
{
	if (elem == VB_VECTOR)
		return (char *) ((Vector *) item)->x;
	if (elem == VB_HALFVEC)
		return (char *) ((HalfVector *) item)->x;
	return (char *) VARBITS((VarBit *) item);
}

static inline Size
VbRowBytes(int elem, int dimensions)
{
	if (elem == VB_VECTOR)
		return sizeof(float) * (Size) dimensions;
	if (elem == VB_HALFVEC)
		return sizeof(half) * (Size) dimensions;
	return ((Size) dimensions + 7) / 8;
}

/*
 * k-means on the sampled rows.  Returns false when the reference path must run
 * (no samples, or fewer samples than lists: src/ivfkmeans.c:110-133, ivfbuild.c:466-472).
 */
bool
VbIvfflatKmeans(Relation index, VectorArray samples, VectorArray centers, const IvfflatTypeInfo * typeInfo)
{
	FmgrInfo   *procinfo = index_getprocinfo(index, 1, IVFFLAT_KMEANS_DISTANCE_PROC);
	int			elem = VB_VECTOR;
	int			metric = VbKmeansMetricFromProc(procinfo, &elem);
	int			dimensions = centers->dim;
	int			k = centers->maxlen;
	Size		rowBytes = VbRowBytes(elem, dimensions);
	char	   *buf;
	char	   *cbuf;
	vb_table   *t = NULL;
	int			iters = 0;

	if (samples->length < k)
		return false;

	/* samples were normalised by SampleRows for the spherical variants (src/ivfbuild.c:153-155) */
	buf = palloc_extended(rowBytes * (Size) samples->length, MCXT_ALLOC_HUGE);
	for (int i = 0; i < samples->length; i++)
		memcpy(buf + rowBytes * (Size) i, VbItemPayload(elem, VectorArrayGet(samples, i)), rowBytes);
	cbuf = palloc(rowBytes * (Size) k);

	VB_CHECK(vb_table_create(elem, dimensions, &t));
	PG_TRY();
	{
		VB_CHECK(vb_table_append(t, buf, samples->length));
		VB_CHECK(vb_kmeans_pp_init(t, metric, cbuf, k, (uint64) RandomInt()));
		/* single process; the 8-GPU build passes an ncclAllReduce wrapper here */
		VB_CHECK(vb_kmeans(t, metric, cbuf, k, 500, (uint64) RandomInt(), NULL, NULL, &iters));
	}
	PG_FINALLY();
	{
		vb_table_free(t);
	}
	PG_END_TRY();

	/* typed centres back into the VectorArray (updateCenter's header work, src/ivfutils.c:301-339) */
	for (int i = 0; i < k; i++)
	{
		Pointer		c = VectorArrayGet(centers, i);

		if (elem == VB_VECTOR)
		{
			SET_VARSIZE(c, VECTOR_SIZE(dimensions));
			((Vector *) c)->dim = (int16) dimensions;
			((Vector *) c)->unused = 0;
		}
		else if (elem == VB_HALFVEC)
		{
			SET_VARSIZE(c, HALFVEC_SIZE(dimensions));
			((HalfVector *) c)->dim = (int16) dimensions;
			((HalfVector *) c)->unused = 0;
		}
		else
		{
			SET_VARSIZE(c, VARBITTOTALLEN((Size) dimensions));
			VARBITLEN((VarBit *) c) = dimensions;
		}
		memcpy(VbItemPayload(elem, c), cbuf + rowBytes * (Size) i, rowBytes);
	}
	centers->length = k;
	(void) typeInfo;
	pfree(buf);
	pfree(cbuf);
	return true;
}

/* rows collected by BuildCallback between flushes */
typedef struct VbAssignBatch
{
	int			elem;
	int			metric;
	int			dimensions;
	Size		rowBytes;
	int			capacity;
	int			n;
	char	   *rows;
	ItemPointerData *tids;
	Datum	   *values;			/* the (normalised) datums, copied into the batch context */
	char	   *centers;		/* packed centre payloads, built once */
	int			lists;
	MemoryContext ctx;
}			VbAssignBatch;

VbAssignBatch *
VbAssignBegin(IvfflatBuildState * buildstate, int capacity)
{
	VbAssignBatch *b = palloc0(sizeof(VbAssignBatch));

	b->metric = VbMetricFromProc(buildstate->procinfo, &b->elem);
	b->dimensions = buildstate->dimensions;
	b->rowBytes = VbRowBytes(b->elem, b->dimensions);
	b->capacity = capacity;
	b->rows = palloc_extended(b->rowBytes * (Size) capacity, MCXT_ALLOC_HUGE);
	b->tids = palloc(sizeof(ItemPointerData) * (Size) capacity);
	b->values = palloc(sizeof(Datum) * (Size) capacity);
	b->lists = buildstate->centers->length;
	b->centers = palloc(b->rowBytes * (Size) b->lists);
	for (int i = 0; i < b->lists; i++)
		memcpy(b->centers + b->rowBytes * (Size) i, VbItemPayload(b->elem, VectorArrayGet(buildstate->centers, i)), b->rowBytes);
	b->ctx = AllocSetContextCreate(CurrentMemoryContext, "vecb200 assign batch", ALLOCSET_DEFAULT_SIZES);
	return b;
}

/* what AddTupleToSort does per row after normalisation, minus the centre loop (src/ivfbuild.c:161-181) */
void
VbAssignAdd(VbAssignBatch * b, ItemPointer tid, Datum value)
{
	MemoryContext old = MemoryContextSwitchTo(b->ctx);
	Datum		copy = datumCopy(value, false, -1);

	MemoryContextSwitchTo(old);
	memcpy(b->rows + b->rowBytes * (Size) b->n, VbItemPayload(b->elem, DatumGetPointer(copy)), b->rowBytes);
	b->tids[b->n] = *tid;
	b->values[b->n] = copy;
	b->n++;
}

/* nearest centre of every buffered row in one call, then the reference's tuplesort feed (src/ivfbuild.c:200-218) */
void
VbAssignFlush(VbAssignBatch * b, IvfflatBuildState * buildstate)
{
	vb_table   *t = NULL;
	int32	   *closest;
	TupleTableSlot *slot = buildstate->slot;

	if (b->n == 0)
		return;
	closest = palloc(sizeof(int32) * (Size) b->n);
	VB_CHECK(vb_table_create(b->elem, b->dimensions, &t));
	PG_TRY();
	{
		VB_CHECK(vb_table_append(t, b->rows, b->n));
		VB_CHECK(vb_assign(t, b->metric, b->centers, b->lists, closest));
	}
	PG_FINALLY();
	{
		vb_table_free(t);
	}
	PG_END_TRY();

	for (int i = 0; i < b->n; i++)
	{
		ExecClearTuple(slot);
		slot->tts_values[0] = Int32GetDatum(closest[i]);
		slot->tts_isnull[0] = false;
		slot->tts_values[1] = PointerGetDatum(&b->tids[i]);
		slot->tts_isnull[1] = false;
		slot->tts_values[2] = b->values[i];
		slot->tts_isnull[2] = false;
		ExecStoreVirtualTuple(slot);
		tuplesort_puttupleslot(buildstate->sortstate, slot);
		buildstate->indtuples++;
	}
	pfree(closest);
	b->n = 0;
	MemoryContextReset(b->ctx);
}
