/*
 * vb_hnsw_build.c -- GPU body for the in-memory phase of CREATE INDEX ... USING hnsw:
 *
 *   VbHnswBuildAdd     replaces InsertTuple -> InsertTupleInMemory (src/hnswbuild.c:437-480, 486-575) inside
 *                      BuildCallback (:583-609): the (normalised) value and its heap TID are buffered
 *   VbHnswBuildFinish  builds the graph of every buffered row on the device (vb_hnsw_build: HnswFindElementNeighbors,
 *                      SelectNeighbors, duplicate folding, HnswUpdateConnection in batches) and materialises it as the
 *                      in-memory graph InsertTupleInMemory would have left behind -- elements linked through
 *                      graph->head, neighbour arrays per layer, duplicates as extra heap TIDs, the entry point -- so
 *                      that the reference's own FlushPages (src/hnswbuild.c:296-316: CreateMetaPage, CreateGraphPages,
 *                      WriteNeighborTuples) writes the index pages unchanged.
 *
 * The level of every element is drawn here with the reference's expression (HnswInitElement, src/hnswutils.c:248-254)
 * from the backend's PRNG and handed to the library, so the level distribution is the reference's.
 * Serial build only (base == NULL); the rows of a parallel build are gathered by the leader the same way.
 * When the buffer would exceed the graph's memory budget (maintenance_work_mem, src/hnswbuild.c:615-624) the caller
 * finishes early and lets the remaining rows take the reference's on-disk insert path, like the reference does when
 * its in-memory graph is full (src/hnswbuild.c:537-560).
 */
#include "postgres.h"

#include "access/genam.h"
#include "utils/memutils.h"
#include "utils/rel.h"
#include "utils/varbit.h"

#include "halfvec.h"
#include "hnsw.h"
#include "vector.h"

#include "vb_glue.h"

typedef struct VbHnswBuildBuffer
{
	int			elem;
	int			metric;
	int			dimensions;
	Size		rowBytes;
	int64		n;
	int64		cap;
	char	   *rows;			/* payloads, packed */
	ItemPointerData *tids;
	Size		bytesLimit;		/* graph->memoryTotal: finish early beyond it */
}			VbHnswBuildBuffer;

static inline const char *
VbHnswPayload(int elem, Datum value)
{
	if (elem == VB_VECTOR)
		return (const char *) DatumGetVector(value)->x;
	if (elem == VB_HALFVEC)
		return (const char *) DatumGetHalfVector(value)->x;
	return (const char *) VARBITS(DatumGetVarBitP(value));
}

VbHnswBuildBuffer *
VbHnswBuildBegin(HnswBuildState * buildstate)
{
	VbHnswBuildBuffer *b = palloc0(sizeof(VbHnswBuildBuffer));

	b->metric = VbMetricFromProc(buildstate->support.procinfo, &b->elem);
	b->dimensions = buildstate->dimensions;
	b->rowBytes = b->elem == VB_VECTOR ? sizeof(float) * (Size) b->dimensions :
		b->elem == VB_HALFVEC ? sizeof(half) * (Size) b->dimensions : ((Size) b->dimensions + 7) / 8;
	b->cap = 65536;
	b->rows = palloc_extended(b->rowBytes * (Size) b->cap, MCXT_ALLOC_HUGE);
	b->tids = palloc_extended(sizeof(ItemPointerData) * (Size) b->cap, MCXT_ALLOC_HUGE);
	b->bytesLimit = buildstate->graph->memoryTotal;
	return b;
}

/* false: the buffer is full (memory budget) -- finish, flush, and insert this and the following rows on disk */
bool
VbHnswBuildAdd(VbHnswBuildBuffer * b, ItemPointer tid, Datum value)
{
	if ((Size) (b->n + 1) * (b->rowBytes + sizeof(ItemPointerData)) > b->bytesLimit)
		return false;
	if (b->n == b->cap)
	{
		b->cap *= 2;
		b->rows = repalloc_huge(b->rows, b->rowBytes * (Size) b->cap);
		b->tids = repalloc_huge(b->tids, sizeof(ItemPointerData) * (Size) b->cap);
	}
	memcpy(b->rows + b->rowBytes * (Size) b->n, VbHnswPayload(b->elem, value), b->rowBytes);
	b->tids[b->n] = *tid;
	b->n++;
	return true;
}

/* the datum an element carries (HnswElementData.value): header + payload, as the heap handed it over */
static char *
VbHnswMakeValue(VbHnswBuildBuffer * b, const char *payload, HnswAllocator * allocator)
{
	Size		size = (b->elem == VB_BIT ? VARBITTOTALLEN((Size) b->dimensions) : 8 + b->rowBytes);
	char	   *v = HnswAlloc(allocator, size);

	memset(v, 0, 8);
	SET_VARSIZE(v, size);
	if (b->elem == VB_BIT)
		VARBITLEN((VarBit *) v) = b->dimensions;
	else
		((Vector *) v)->dim = (int16) b->dimensions;	/* Vector and HalfVector share the header layout */
	memcpy(v + 8, payload, b->rowBytes);
	return v;
}

void
VbHnswBuildFinish(VbHnswBuildBuffer * b, HnswBuildState * buildstate)
{
	HnswGraph  *graph = buildstate->graph;
	HnswAllocator *allocator = &buildstate->allocator;
	char	   *base = buildstate->hnswarea;	/* NULL for the serial build */
	int			m = buildstate->m;
	int64		n = b->n;
	vb_hnsw    *ix = NULL;
	int32	   *levels;
	int32	   *nbr0 = NULL;
	int64	   *upper_off = NULL;
	int32	   *upper = NULL;
	int32	   *dup_of = NULL;
	int64		entry = -1;
	int64		slots = 0;
	HnswElement *elements;
	int			rc;

	if (n == 0)
		return;

	/* level draws: (int) (-log(RandomDouble()) * ml), capped (HnswInitElement, src/hnswutils.c:248-254) */
	levels = palloc_extended(sizeof(int32) * (Size) n, MCXT_ALLOC_HUGE);
	for (int64 i = 0; i < n; i++)
	{
		int			level = (int) (-log(RandomDouble()) * buildstate->ml);

		levels[i] = Min(level, buildstate->maxLevel);
	}

	VB_CHECK(vb_hnsw_create(b->elem, b->metric, b->dimensions, m, &ix));
	rc = vb_hnsw_build(ix, b->rows, n, buildstate->efConstruction, 0, levels);
	if (rc == VB_OK)
	{
		slots = vb_hnsw_upper_slots(ix);
		nbr0 = palloc_extended(sizeof(int32) * (Size) n * 2 * m, MCXT_ALLOC_HUGE);
		upper_off = palloc_extended(sizeof(int64) * (Size) n, MCXT_ALLOC_HUGE);
		upper = palloc_extended(sizeof(int32) * (Size) Max(slots, 1) * m, MCXT_ALLOC_HUGE);
		dup_of = palloc_extended(sizeof(int32) * (Size) n, MCXT_ALLOC_HUGE);
		rc = vb_hnsw_export(ix, levels, nbr0, upper_off, upper, &entry, dup_of);
	}
	/* release the device image BEFORE any ereport: nothing is held across a longjmp */
	vb_hnsw_free(ix);
	if (rc != VB_OK)
		ereport(ERROR,
				(errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION),
				 errmsg("vecb200: %s", vb_last_error())));

	/* elements (what HnswInitElement + AddElementInMemory leave), duplicates as extra heap TIDs of their element */
	elements = palloc_extended(sizeof(HnswElement) * (Size) n, MCXT_ALLOC_HUGE);
	for (int64 i = 0; i < n; i++)
	{
		HnswElement element;

		if (dup_of[i] >= 0)
		{
			/* AddDuplicateInMemory (src/hnswbuild.c:321-337): the library never folds more than HNSW_HEAPTIDS - 1 rows */
			HnswAddHeapTid(elements[dup_of[i]], &b->tids[i]);
			elements[i] = NULL;
			continue;
		}
		element = HnswAlloc(allocator, sizeof(HnswElementData));
		element->heaptidsLength = 0;
		HnswAddHeapTid(element, &b->tids[i]);
		element->level = (uint8) levels[i];
		element->deleted = 0;
		element->version = 1;
		HnswInitNeighbors(base, element, m, allocator);
		HnswPtrStore(base, element->value, VbHnswMakeValue(b, b->rows + b->rowBytes * (Size) i, allocator));
		element->next = graph->head;
		HnswPtrStore(base, graph->head, element);
		elements[i] = element;
	}
	/* neighbour arrays, in the stored order (what AddConnections / HnswUpdateConnection left) */
	for (int64 i = 0; i < n; i++)
	{
		if (elements[i] == NULL)
			continue;
		for (int lc = 0; lc <= levels[i]; lc++)
		{
			int			lm = HnswGetLayerM(m, lc);
			const int32 *src = lc == 0 ? nbr0 + i * 2 * m : upper + (upper_off[i] + lc - 1) * m;
			HnswNeighborArray *a = HnswGetNeighbors(base, elements[i], lc);

			for (int j = 0; j < lm && src[j] >= 0; j++)
			{
				HnswCandidate *hc = &a->items[a->length++];

				HnswPtrStore(base, hc->element, elements[src[j]]);
				hc->distance = 0;	/* only the page writer reads these arrays from here on (HnswSetNeighborTuple) */
				hc->closer = false;
			}
		}
	}
	if (entry >= 0)
		HnswPtrStore(base, graph->entryPoint, elements[entry]);
	graph->indtuples += (double) n;
	b->n = 0;
}
