/*
 * vb_hnsw_scan.c -- GPU body for the HNSW scan:
 *
 *   VbHnswGetScanItems replaces GetScanItems (src/hnswscan.c:25-56), i.e. the
 *   greedy descent + HnswSearchLayer (src/hnswutils.c:824-987) on the on-disk graph,
 *
 * plus the packer that turns element / neighbour tuples (src/hnsw.h:348-394) into
 * the device image: element numbers replace index TIDs, neighbour lists keep
 * their on-disk order (HnswSetNeighborTuple, src/hnswutils.c:455-486).
 * hnswgettuple (src/hnswscan.c:189-331) keeps popping the nearest element and
 * its heap TIDs last-added-first (:293-311); see INTEGRATION.md.
 */
#include "postgres.h"

#include "access/genam.h"
#include "access/relscan.h"
#include "common/hashfn.h"
#include "storage/bufmgr.h"
#include "utils/float.h"
#include "utils/memutils.h"
#include "utils/rel.h"
#include "utils/varbit.h"

#include "halfvec.h"
#include "hnsw.h"
#include "vector.h"

#include "vb_glue.h"

typedef struct VbHnswCacheEntry
{
	VbHnswImage image;
	struct VbHnswCacheEntry *next;
}			VbHnswCacheEntry;

static VbHnswCacheEntry * hnswCache = NULL;

static void
VbHnswDropImage(VbHnswImage * img)
{
	if (img->ix != NULL)
		vb_hnsw_free(img->ix);
	img->ix = NULL;
	if (img->heaptids != NULL)
		pfree(img->heaptids);
	if (img->nheaptids != NULL)
		pfree(img->nheaptids);
	img->heaptids = NULL;
	img->nheaptids = NULL;
}

void
VbHnswInvalidate(Oid relid)
{
	for (VbHnswCacheEntry * e = hnswCache; e != NULL; e = e->next)
		if (e->image.relid == relid)
			VbHnswDropImage(&e->image);
}

/* open-addressing map index TID -> element number, built while walking the element pages */
typedef struct VbTidMap
{
	uint64	   *keys;			/* (blkno << 16 | offno) + 1, 0 = empty */
	int32	   *vals;
	uint64		mask;
}			VbTidMap;

static void
VbTidMapInit(VbTidMap * map, int64 n)
{
	uint64		cap = 16;

	while (cap < (uint64) n * 2)
		cap <<= 1;
	map->keys = palloc0(sizeof(uint64) * cap);
	map->vals = palloc(sizeof(int32) * cap);
	map->mask = cap - 1;
}

static inline uint64
VbTidKey(BlockNumber blkno, OffsetNumber offno)
{
	return (((uint64) blkno << 16) | offno) + 1;
}

static void
VbTidMapPut(VbTidMap * map, uint64 key, int32 val)
{
	uint64		h = murmurhash64(key) & map->mask;

	while (map->keys[h] != 0)
		h = (h + 1) & map->mask;
	map->keys[h] = key;
	map->vals[h] = val;
}

static int32
VbTidMapGet(VbTidMap * map, uint64 key)
{
	uint64		h = murmurhash64(key) & map->mask;

	while (map->keys[h] != 0)
	{
		if (map->keys[h] == key)
			return map->vals[h];
		h = (h + 1) & map->mask;
	}
	return -1;
}

static void
VbHnswPack(Relation index, VbHnswImage * img)
{
	BlockNumber nblocks = RelationGetNumberOfBlocks(index);
	MemoryContext packCtx = AllocSetContextCreate(CurrentMemoryContext, "vecb200 hnsw pack", ALLOCSET_DEFAULT_SIZES);
	MemoryContext oldCtx = MemoryContextSwitchTo(packCtx);
	int			m;
	HnswElement entryPoint;
	int64		n = 0;
	int64		cap = 1024;
	StringInfoData rows;
	int32	   *levels = palloc(sizeof(int32) * cap);
	ItemPointerData *neighbortids = palloc(sizeof(ItemPointerData) * cap);
	ItemPointerData *selftids = palloc(sizeof(ItemPointerData) * cap);
	ItemPointerData *heaptids = palloc(sizeof(ItemPointerData) * cap * HNSW_HEAPTIDS);
	uint8	   *nheaptids = palloc(sizeof(uint8) * cap);
	VbTidMap	map;
	int32	   *nbr0;
	int64	   *upper_off;
	int32	   *upper;
	int64		slots = 0;
	int64		entry = -1;

	HnswGetMetaPageInfo(index, &m, &entryPoint);	/* src/hnswutils.c:298-328 */
	img->m = m;
	initStringInfo(&rows);

	/* pass 1: every live element tuple, in (block, offset) order = element numbers */
	for (BlockNumber blkno = HNSW_HEAD_BLKNO; blkno < nblocks; blkno++)
	{
		Buffer		buf = ReadBufferExtended(index, MAIN_FORKNUM, blkno, RBM_NORMAL, NULL);
		Page		page;
		OffsetNumber maxoffno;

		LockBuffer(buf, BUFFER_LOCK_SHARE);
		page = BufferGetPage(buf);
		maxoffno = PageGetMaxOffsetNumber(page);
		for (OffsetNumber offno = FirstOffsetNumber; offno <= maxoffno; offno = OffsetNumberNext(offno))
		{
			HnswElementTuple etup = (HnswElementTuple) PageGetItem(page, PageGetItemId(page, offno));
			Size		bytes;
			const void *payload;

			/*
			 * Elements being deleted stay in the image: the reference still traverses them (CountElement is true for
			 * scans, src/hnswutils.c:714-727) and only withholds their heap TIDs (heaptidsLength = 0 after
			 * HnswLoadElementFromTuple) -- dropping them would cut the graph mid-vacuum.
			 */
			if (!HnswIsElementTuple(etup))
				continue;
			if (n == cap)
			{
				cap *= 2;
				levels = repalloc(levels, sizeof(int32) * cap);
				neighbortids = repalloc(neighbortids, sizeof(ItemPointerData) * cap);
				selftids = repalloc(selftids, sizeof(ItemPointerData) * cap);
				heaptids = repalloc_huge(heaptids, sizeof(ItemPointerData) * cap * HNSW_HEAPTIDS);
				nheaptids = repalloc(nheaptids, sizeof(uint8) * cap);
			}
			if (img->elem == VB_VECTOR)
			{
				bytes = sizeof(float) * (Size) img->dimensions;
				payload = ((Vector *) &etup->data)->x;
			}
			else if (img->elem == VB_HALFVEC)
			{
				bytes = sizeof(half) * (Size) img->dimensions;
				payload = ((HalfVector *) &etup->data)->x;
			}
			else
			{
				bytes = VARBITBYTES((VarBit *) &etup->data);
				payload = VARBITS((VarBit *) &etup->data);
			}
			appendBinaryStringInfo(&rows, payload, (int) bytes);
			levels[n] = etup->level;
			neighbortids[n] = etup->neighbortid;
			ItemPointerSet(&selftids[n], blkno, offno);
			nheaptids[n] = 0;
			for (int i = 0; i < HNSW_HEAPTIDS && !etup->deleted; i++)
			{
				if (!ItemPointerIsValid(&etup->heaptids[i]))
					break;
				heaptids[n * HNSW_HEAPTIDS + nheaptids[n]++] = etup->heaptids[i];
			}
			n++;
		}
		UnlockReleaseBuffer(buf);
	}

	VbTidMapInit(&map, n);
	for (int64 e = 0; e < n; e++)
	{
		VbTidMapPut(&map, VbTidKey(ItemPointerGetBlockNumber(&selftids[e]), ItemPointerGetOffsetNumber(&selftids[e])), (int32) e);
		if (levels[e] >= 1)
			slots += levels[e];
	}
	if (entryPoint != NULL)
		entry = VbTidMapGet(&map, VbTidKey(entryPoint->blkno, entryPoint->offno));

	/* pass 2: neighbour tuples -> element numbers, layer 0 table and upper-layer slots */
	nbr0 = palloc_extended(sizeof(int32) * (Size) Max(n, 1) * 2 * m, MCXT_ALLOC_HUGE);
	upper_off = palloc(sizeof(int64) * (Size) Max(n, 1));
	upper = palloc_extended(sizeof(int32) * (Size) Max(slots, 1) * m, MCXT_ALLOC_HUGE);
	slots = 0;
	for (int64 e = 0; e < n; e++)
	{
		Buffer		buf = ReadBuffer(index, ItemPointerGetBlockNumber(&neighbortids[e]));
		Page		page;
		HnswNeighborTuple ntup;
		int			level = levels[e];

		LockBuffer(buf, BUFFER_LOCK_SHARE);
		page = BufferGetPage(buf);
		ntup = (HnswNeighborTuple) PageGetItem(page, PageGetItemId(page, ItemPointerGetOffsetNumber(&neighbortids[e])));

		upper_off[e] = level >= 1 ? slots : -1;
		for (int lc = level; lc >= 0; lc--)
		{
			/* layer lc starts at (level - lc) * m (src/hnswutils.c:786) */
			int			lm = HnswGetLayerM(m, lc);
			ItemPointer tids = ntup->indextids + (level - lc) * m;
			int32	   *dst = lc == 0 ? nbr0 + e * 2 * m : upper + (upper_off[e] + (lc - 1)) * m;
			int			w = 0;

			for (int i = 0; i < lm; i++)
			{
				int32		v;

				if (!ItemPointerIsValid(&tids[i]))
					break;		/* an invalid TID terminates the list (src/hnswutils.c:809-810) */
				v = VbTidMapGet(&map, VbTidKey(ItemPointerGetBlockNumber(&tids[i]), ItemPointerGetOffsetNumber(&tids[i])));
				if (v >= 0)
					dst[w++] = v;
			}
			while (w < lm)
				dst[w++] = -1;
		}
		if (level >= 1)
			slots += level;
		UnlockReleaseBuffer(buf);
	}

	if (n > 0 && entry < 0)
	{
		/* the meta page names an entry point the element pages do not hold (concurrent repair): CPU path for now */
		MemoryContextSwitchTo(oldCtx);
		MemoryContextDelete(packCtx);
		return;
	}
	VB_CHECK(vb_hnsw_create(img->elem, img->metric, img->dimensions, m, &img->ix));
	{
		int			rc = vb_hnsw_load(img->ix, rows.data, n, levels, nbr0, upper_off, upper, slots, entry);

		if (rc != VB_OK)
		{
			VbHnswDropImage(img);	/* before raising: no device state across the longjmp */
			ereport(ERROR,
					(errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION),
					 errmsg("vecb200: %s", vb_last_error())));
		}
	}

	/* heap TIDs stay on the host: hnswgettuple expands elements into them */
	img->n = n;
	img->heaptids = MemoryContextAllocHuge(TopMemoryContext, sizeof(ItemPointerData) * (Size) Max(n, 1) * HNSW_HEAPTIDS);
	img->nheaptids = MemoryContextAlloc(TopMemoryContext, (Size) Max(n, 1));
	memcpy(img->heaptids, heaptids, sizeof(ItemPointerData) * (Size) n * HNSW_HEAPTIDS);
	memcpy(img->nheaptids, nheaptids, (Size) n);

	MemoryContextSwitchTo(oldCtx);
	MemoryContextDelete(packCtx);
}

VbHnswImage *
VbHnswGetImage(Relation index, FmgrInfo *procinfo)
{
	Oid			relid = RelationGetRelid(index);
	BlockNumber nblocks = RelationGetNumberOfBlocks(index);
	uint64		version = VbIndexVersion(index);
	VbHnswCacheEntry *e;

	for (e = hnswCache; e != NULL; e = e->next)
		if (e->image.relid == relid)
			break;
	if (e == NULL)
	{
		e = MemoryContextAllocZero(TopMemoryContext, sizeof(VbHnswCacheEntry));
		e->image.relid = relid;
		e->next = hnswCache;
		hnswCache = e;
	}
	if (e->image.ix != NULL && (e->image.nblocks != nblocks || e->image.version != version))
		VbHnswDropImage(&e->image);
	if (e->image.ix == NULL)
	{
		e->image.metric = VbMetricFromProc(procinfo, &e->image.elem);
		e->image.dimensions = TupleDescAttr(RelationGetDescr(index), 0)->atttypmod;
		e->image.nblocks = nblocks;
		e->image.version = version;
		VbHnswPack(index, &e->image);
	}
	return e->image.ix != NULL ? &e->image : NULL;
}

/* GetScanItems (src/hnswscan.c:25-56): one C ABI call, ef_search results nearest first */
bool
VbHnswGetScanItems(IndexScanDesc scan, Datum value, int ef_search, VbHnswScanState * st)
{
	HnswScanOpaque so = (HnswScanOpaque) scan->opaque;
	VbHnswImage *img;
	const void *q;

	/* NULL query: every distance is 0 (src/hnswutils.c:555-556); the reference loop serves it */
	if (DatumGetPointer(value) == NULL)
		return false;
	img = VbHnswGetImage(scan->indexRelation, so->support.procinfo);
	if (img == NULL)
		return false;

	st->image = img;
	st->elements = palloc(sizeof(int64) * (Size) ef_search);
	st->distances = palloc(sizeof(double) * (Size) ef_search);
	st->nelements = 0;
	st->cur = 0;
	st->curtid = -1;

	if (img->elem == VB_VECTOR)
		q = DatumGetVector(value)->x;
	else if (img->elem == VB_HALFVEC)
		q = DatumGetHalfVector(value)->x;
	else
		q = VARBITS(DatumGetVarBitP(value));

	st->iter = NULL;
	st->batch = ef_search;
	st->previousDistance = -get_float8_infinity();
	if (hnsw_iterative_scan != HNSW_ITERATIVE_SCAN_OFF)
	{
		int32		count = 0;

		/* the first batch is GetScanItems with the discarded heap (src/hnswscan.c:55) */
		VB_CHECK(vb_hnsw_scan_begin(img->ix, q, 1, ef_search, hnsw_max_scan_tuples, &st->iter));
		if (vb_hnsw_scan_next(st->iter, st->elements, st->distances, &count) != VB_OK)
		{
			VbHnswEndScan(st);
			ereport(ERROR, (errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION), errmsg("vecb200: %s", vb_last_error())));
		}
		st->nelements = count;
		(void) vb_hnsw_scan_tuples(st->iter, &so->tuples);
		return true;
	}
	VB_CHECK(vb_hnsw_search(img->ix, q, 1, ef_search, ef_search, st->elements, st->distances, &so->tuples));
	while (st->nelements < ef_search && st->elements[st->nelements] >= 0)
		st->nelements++;
	return true;
}

void
VbHnswEndScan(VbHnswScanState * st)
{
	if (st->iter != NULL)
		vb_hnsw_scan_end(st->iter);
	st->iter = NULL;
}

/* the loop of hnswgettuple (src/hnswscan.c:293-326): nearest element first, heap TIDs last-added first */
bool
VbHnswNextItem(IndexScanDesc scan, VbHnswScanState * st)
{
	VbHnswImage *img = st->image;

	for (;;)
	{
		int64		e;

		if (st->cur >= st->nelements)
		{
			int32		count = 0;

			/* W is consumed: resume from the discarded candidates, or stop (src/hnswscan.c:236-262) */
			if (st->iter == NULL)
				return false;
			if (vb_hnsw_scan_next(st->iter, st->elements, st->distances, &count) != VB_OK)
			{
				VbHnswEndScan(st);
				ereport(ERROR, (errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION), errmsg("vecb200: %s", vb_last_error())));
			}
			st->nelements = count;
			st->cur = 0;
			st->curtid = -1;
			(void) vb_hnsw_scan_tuples(st->iter, &((HnswScanOpaque) scan->opaque)->tuples);
			if (count == 0)
			{
				VbHnswEndScan(st);
				return false;
			}
		}
		e = st->elements[st->cur];
		if (st->curtid < 0)
		{
			st->curtid = img->nheaptids[e];
			/* strict_order: an element nearer than one already returned is dropped (src/hnswscan.c:316-322) */
			if (hnsw_iterative_scan == HNSW_ITERATIVE_SCAN_STRICT && st->curtid > 0)
			{
				if (st->distances[st->cur] < st->previousDistance)
					st->curtid = 0;
				else
					st->previousDistance = st->distances[st->cur];
			}
		}
		if (st->curtid == 0)
		{
			st->cur++;
			st->curtid = -1;
			continue;
		}
		scan->xs_heaptid = img->heaptids[e * HNSW_HEAPTIDS + (--st->curtid)];
		scan->xs_recheck = false;
		scan->xs_recheckorderby = false;
		return true;
	}
}
