/*
 * vb_ivfflat_scan.c -- GPU bodies for the two hot loops of the IVFFlat scan:
 *
 *   VbGetScanLists  replaces GetScanLists  (src/ivfscan.c:47-118)
 *   VbGetScanItems  replaces GetScanItems  (src/ivfscan.c:123-187)
 *
 * and the packer that turns an index's list pages / entry pages
 * (src/ivfflat.h:251-277) into the device image of include/vecb200.h.
 * ivfflatgettuple (src/ivfscan.c:360-414) keeps its structure; see
 * INTEGRATION.md for the five-line patch that calls these.
 */
#include "postgres.h"

#include "access/genam.h"
#include "access/itup.h"
#include "access/relscan.h"
#include "storage/bufmgr.h"
#include "utils/memutils.h"
#include "utils/rel.h"
#include "utils/varbit.h"

#include "halfvec.h"
#include "ivfflat.h"
#include "vector.h"

#include "vb_glue.h"

/* distance wrappers exported by the extension (PGDLLEXPORT in src/vector.c, halfvec.c, bitvec.c) */
extern Datum vector_l2_squared_distance(PG_FUNCTION_ARGS);
extern Datum vector_negative_inner_product(PG_FUNCTION_ARGS);
extern Datum l1_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_l2_squared_distance(PG_FUNCTION_ARGS);
extern Datum halfvec_negative_inner_product(PG_FUNCTION_ARGS);
extern Datum halfvec_l1_distance(PG_FUNCTION_ARGS);
extern Datum hamming_distance(PG_FUNCTION_ARGS);
extern Datum jaccard_distance(PG_FUNCTION_ARGS);

/*
 * Which kernel does opclass proc 1 stand for?  The index resolved it with
 * index_getprocinfo (src/ivfscan.c:288-289, src/hnswutils.c:152-158); compare the
 * C entry point instead of adding a support-function number.
 */
int
VbMetricFromProc(FmgrInfo *procinfo, int *elem)
{
	PGFunction	fn = procinfo->fn_addr;

	*elem = VB_VECTOR;
	if (fn == vector_l2_squared_distance)
		return VB_L2_SQUARED;
	if (fn == vector_negative_inner_product)
		return VB_NEG_IP;
	if (fn == l1_distance)
		return VB_L1;
	*elem = VB_HALFVEC;
	if (fn == halfvec_l2_squared_distance)
		return VB_L2_SQUARED;
	if (fn == halfvec_negative_inner_product)
		return VB_NEG_IP;
	if (fn == halfvec_l1_distance)
		return VB_L1;
	*elem = VB_BIT;
	if (fn == hamming_distance)
		return VB_HAMMING;
	if (fn == jaccard_distance)
		return VB_JACCARD;
	elog(ERROR, "vecb200: unsupported distance function");
	return -1;
}

/* payload pointer (what the C ABI calls a "row") of a detoasted datum */
static inline const void *
VbPayload(int elem, Datum d, Size *bytes, int dimensions)
{
	if (elem == VB_VECTOR)
	{
		Vector	   *v = DatumGetVector(d);

		*bytes = sizeof(float) * (Size) dimensions;
		return v->x;
	}
	else if (elem == VB_HALFVEC)
	{
		HalfVector *v = DatumGetHalfVector(d);

		*bytes = sizeof(half) * (Size) dimensions;
		return v->x;
	}
	else
	{
		VarBit	   *v = DatumGetVarBitP(d);

		*bytes = VARBITBYTES(v);
		return VARBITS(v);
	}
}

/* one cached image per backend and index (a backend scans few indexes; linear list) */
typedef struct VbIvfCacheEntry
{
	VbIvfImage	image;
	struct VbIvfCacheEntry *next;
}			VbIvfCacheEntry;

static VbIvfCacheEntry * ivfCache = NULL;

static void
VbIvfDropImage(VbIvfImage * img)
{
	if (img->ix != NULL)
		vb_ivf_free(img->ix);
	img->ix = NULL;
	if (img->startPages != NULL)
		pfree(img->startPages);
	img->startPages = NULL;
}

void
VbIvfInvalidate(Oid relid)
{
	for (VbIvfCacheEntry * e = ivfCache; e != NULL; e = e->next)
		if (e->image.relid == relid)
			VbIvfDropImage(&e->image);
}

/* the version stamp in the meta page (block 0 of both AMs), see vb_glue.h */
uint64
VbIndexVersion(Relation index)
{
	Buffer		buf = ReadBuffer(index, 0);
	uint64		version;

	LockBuffer(buf, BUFFER_LOCK_SHARE);
	memcpy(&version, BufferGetPage(buf) + MAXALIGN(SizeOfPageHeaderData) + VB_META_VERSION_OFFSET, sizeof(version));
	UnlockReleaseBuffer(buf);
	return version;
}

/*
 * What aminsert / ambulkdelete / ambuild call after they changed the index (shown without the GenericXLog
 * registration the AM wraps around every page change, src/ivfinsert.c:104-141 style).
 */
void
VbBumpIndexVersion(Relation index)
{
	Buffer		buf = ReadBuffer(index, 0);
	char	   *slot;
	uint64		version;

	LockBuffer(buf, BUFFER_LOCK_EXCLUSIVE);
	slot = BufferGetPage(buf) + MAXALIGN(SizeOfPageHeaderData) + VB_META_VERSION_OFFSET;
	memcpy(&version, slot, sizeof(version));
	version++;
	memcpy(slot, &version, sizeof(version));
	UnlockReleaseBuffer(buf);
}

/*
 * Walk the list pages and every list's entry-page chain exactly like
 * GetScanLists / GetScanItems do, but copy instead of computing:
 * centres, per-list row payloads, heap TIDs.  One pass per index version.
 */
static void
VbIvfPack(Relation index, VbIvfImage * img)
{
	TupleDesc	tupdesc = RelationGetDescr(index);
	BlockNumber nextblkno = IVFFLAT_HEAD_BLKNO;
	Size		rowBytes = 0;
	int			lists = 0;
	int			maxLists = 1024;
	char	   *centers = NULL;
	int64	   *offsets;
	StringInfoData rows;
	StringInfoData ids;
	MemoryContext packCtx = AllocSetContextCreate(CurrentMemoryContext, "vecb200 ivfflat pack", ALLOCSET_DEFAULT_SIZES);
	MemoryContext oldCtx = MemoryContextSwitchTo(packCtx);

	img->startPages = MemoryContextAlloc(TopMemoryContext, sizeof(BlockNumber) * maxLists);

	/* pass 1: centres + start pages (src/ivfscan.c:56-111) */
	while (BlockNumberIsValid(nextblkno))
	{
		Buffer		cbuf = ReadBuffer(index, nextblkno);
		Page		cpage;
		OffsetNumber maxoffno;

		LockBuffer(cbuf, BUFFER_LOCK_SHARE);
		cpage = BufferGetPage(cbuf);
		maxoffno = PageGetMaxOffsetNumber(cpage);

		for (OffsetNumber offno = FirstOffsetNumber; offno <= maxoffno; offno = OffsetNumberNext(offno))
		{
			IvfflatList list = (IvfflatList) PageGetItem(cpage, PageGetItemId(cpage, offno));
			Size		bytes;
			const void *payload = VbPayload(img->elem, PointerGetDatum(&list->center), &bytes, img->dimensions);

			if (centers == NULL)
			{
				rowBytes = bytes;
				centers = palloc(rowBytes * (Size) maxLists);
			}
			if (lists == maxLists)
			{
				maxLists *= 2;
				centers = repalloc(centers, rowBytes * (Size) maxLists);
				img->startPages = repalloc(img->startPages, sizeof(BlockNumber) * maxLists);
			}
			memcpy(centers + rowBytes * (Size) lists, payload, rowBytes);
			img->startPages[lists] = list->startPage;
			lists++;
		}
		nextblkno = IvfflatPageGetOpaque(cpage)->nextblkno;
		UnlockReleaseBuffer(cbuf);
	}
	img->lists = lists;

	/* pass 2: rows and heap TIDs, list by list (src/ivfscan.c:134-179) */
	offsets = palloc(sizeof(int64) * ((Size) lists + 1));
	initStringInfo(&rows);
	initStringInfo(&ids);
	offsets[0] = 0;
	for (int l = 0; l < lists; l++)
	{
		BlockNumber searchPage = img->startPages[l];
		int64		n = 0;

		while (BlockNumberIsValid(searchPage))
		{
			Buffer		buf = ReadBufferExtended(index, MAIN_FORKNUM, searchPage, RBM_NORMAL, NULL);
			Page		page;
			OffsetNumber maxoffno;

			LockBuffer(buf, BUFFER_LOCK_SHARE);
			page = BufferGetPage(buf);
			maxoffno = PageGetMaxOffsetNumber(page);

			for (OffsetNumber offno = FirstOffsetNumber; offno <= maxoffno; offno = OffsetNumberNext(offno))
			{
				IndexTuple	itup = (IndexTuple) PageGetItem(page, PageGetItemId(page, offno));
				bool		isnull;
				Datum		datum = index_getattr(itup, 1, tupdesc, &isnull);
				Size		bytes;
				const void *payload = VbPayload(img->elem, datum, &bytes, img->dimensions);
				int64		id = VbTidToId(&itup->t_tid);

				appendBinaryStringInfo(&rows, payload, (int) bytes);
				appendBinaryStringInfo(&ids, (const char *) &id, sizeof(int64));
				n++;
			}
			searchPage = IvfflatPageGetOpaque(page)->nextblkno;
			UnlockReleaseBuffer(buf);
		}
		offsets[l + 1] = offsets[l] + n;
	}

	/* hand the image to the device: pinned staging + DMA happen inside vb_ivf_load */
	VB_CHECK(vb_ivf_create(img->elem, img->metric, img->dimensions, lists, &img->ix));
	{
		int			rc = vb_ivf_load(img->ix, centers, offsets, rows.data, (const int64 *) ids.data);

		if (rc != VB_OK)
		{
			/* release the handle BEFORE raising: no device state is held across the longjmp */
			VbIvfDropImage(img);
			ereport(ERROR,
					(errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION),
					 errmsg("vecb200: %s", vb_last_error())));
		}
	}

	MemoryContextSwitchTo(oldCtx);
	MemoryContextDelete(packCtx);
}

VbIvfImage *
VbIvfGetImage(Relation index, FmgrInfo *procinfo, int dimensions)
{
	Oid			relid = RelationGetRelid(index);
	BlockNumber nblocks = RelationGetNumberOfBlocks(index);
	uint64		version = VbIndexVersion(index);
	VbIvfCacheEntry *e;

	for (e = ivfCache; e != NULL; e = e->next)
		if (e->image.relid == relid)
			break;
	if (e == NULL)
	{
		e = MemoryContextAllocZero(TopMemoryContext, sizeof(VbIvfCacheEntry));
		e->image.relid = relid;
		e->next = ivfCache;
		ivfCache = e;
	}
	/* stale (the index grew, or any backend inserted / vacuumed since the image was packed) or never packed */
	if (e->image.ix != NULL && (e->image.nblocks != nblocks || e->image.version != version))
		VbIvfDropImage(&e->image);
	if (e->image.ix == NULL)
	{
		e->image.metric = VbMetricFromProc(procinfo, &e->image.elem);
		e->image.dimensions = dimensions;
		e->image.nblocks = nblocks;
		e->image.version = version;
		VbIvfPack(index, &e->image);
	}
	return &e->image;
}

static inline const void *
VbQueryPayload(VbIvfImage * img, Datum value)
{
	Size		bytes;

	/* NULL query => NULL pointer => every distance is 0 (src/ivfscan.c:207-211) */
	if (DatumGetPointer(value) == NULL)
		return NULL;
	return VbPayload(img->elem, value, &bytes, img->dimensions);
}

/*
 * GetScanLists: distance(query, every centre), nearest maxProbes lists first.
 * Fills so->listPages[] exactly like the reference (src/ivfscan.c:114-115).
 */
void
VbGetScanLists(IndexScanDesc scan, Datum value, VbIvfScanState * st)
{
	IvfflatScanOpaque so = (IvfflatScanOpaque) scan->opaque;
	VbIvfImage *img = VbIvfGetImage(scan->indexRelation, so->procinfo, so->dimensions);

	st->image = img;
	st->lists = palloc(sizeof(int32) * (Size) so->maxProbes);
	VB_CHECK(vb_ivf_scan_lists(img->ix, VbQueryPayload(img, value), 1, so->maxProbes, st->lists, NULL));
	st->nlists = Min(so->maxProbes, img->lists);
	for (int i = 0; i < st->nlists; i++)
		so->listPages[i] = img->startPages[st->lists[i]];
}

/*
 * GetScanItems: the next `probes` lists, every row scored, fully sorted
 * (tuplesort_performsort, src/ivfscan.c:182) -- one C ABI call.
 */
void
VbGetScanItems(IndexScanDesc scan, Datum value, VbIvfScanState * st)
{
	IvfflatScanOpaque so = (IvfflatScanOpaque) scan->opaque;
	VbIvfImage *img = st->image;
	int			batch = 0;
	int64		total = 0;
	int			first = so->listIndex;

	/* same batching as src/ivfscan.c:134 */
	while (so->listIndex < so->maxProbes && so->listIndex < st->nlists && batch < so->probes)
	{
		so->listIndex++;
		batch++;
	}
	/* count first (cap = all), then fetch */
	VB_CHECK(vb_ivf_scan_items(img->ix, VbQueryPayload(img, value), st->lists + first, batch, 0, NULL, NULL, &total));
	st->ids = palloc(sizeof(int64) * (Size) Max(total, 1));
	st->distances = palloc(sizeof(double) * (Size) Max(total, 1));
	VB_CHECK(vb_ivf_scan_items(img->ix, VbQueryPayload(img, value), st->lists + first, batch, total, st->ids, st->distances, &total));
	st->nitems = total;
	st->next = 0;
}

/* the tail of ivfflatgettuple (src/ivfscan.c:400-413): stream the sorted batch */
bool
VbNextItem(IndexScanDesc scan, VbIvfScanState * st)
{
	if (st->next >= st->nitems)
		return false;
	VbIdToTid(st->ids[st->next++], &scan->xs_heaptid);
	scan->xs_recheck = false;
	scan->xs_recheckorderby = false;
	return true;
}
