/*
 * vb_glue.h -- extension-side glue between pgvector's index access methods and
 * libvecb200 (include/vecb200.h).  These files are compiled INSIDE the pgvector
 * source tree (added to OBJS in its Makefile); they use the
 * PostgreSQL API and the reference's own headers (ivfflat.h, hnsw.h) and keep
 * the SQL / index-AM surface unchanged.
 *
 * PostgreSQL is not present in the development image: these files are compiled
 * against pgstub/ + the reference's real headers and RUN by the test harness
 * (tests/harness, tests/test_ext_harness.py) over synthesised index pages; they
 * have not been run inside a server.
 */
#ifndef VB_GLUE_H
#define VB_GLUE_H

#include "postgres.h"

#include "access/genam.h"
#include "storage/block.h"
#include "access/relscan.h"
#include "storage/itemptr.h"
#include "utils/rel.h"

#include "vecb200.h"

/* heap TID <-> the opaque int64 id of the C ABI */
static inline int64
VbTidToId(ItemPointer tid)
{
	return ((int64) ItemPointerGetBlockNumber(tid) << 16) | (int64) ItemPointerGetOffsetNumber(tid);
}

static inline void
VbIdToTid(int64 id, ItemPointer tid)
{
	ItemPointerSet(tid, (BlockNumber) (id >> 16), (OffsetNumber) (id & 0xFFFF));
}

/* raise the library's last error as a PostgreSQL ERROR (called AFTER the C ABI call returned) */
#define VB_CHECK(call) \
	do { \
		int			vb_rc__ = (call); \
		if (vb_rc__ != VB_OK) \
			ereport(ERROR, \
					(errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION), \
					 errmsg("vecb200: %s", vb_last_error()))); \
	} while (0)

/*
 * Cross-backend invalidation stamp.  A cached device image is valid for one (relation size, index version) pair.
 * The version is a uint64 the patched AM keeps in the spare space of the index's meta page, VB_META_VERSION_OFFSET
 * bytes into the page contents (beyond IvfflatMetaPageData / HnswMetaPageData; zero on an unpatched index), and
 * bumps under the meta page's buffer lock in aminsert, ambulkdelete / amvacuumcleanup and ambuild
 * (INTEGRATION.md section 5 shows the three call sites).  Every backend compares it before a scan, so a write by
 * ANY backend -- an insert into free space, a vacuum that recycles heap TIDs -- repacks the image.
 */
#define VB_META_VERSION_OFFSET 64
extern uint64 VbIndexVersion(Relation index);
extern void VbBumpIndexVersion(Relation index);

/* ---- IVFFlat (vb_ivfflat_scan.c, vb_ivfflat_build.c) ---- */

typedef struct VbIvfImage
{
	Oid			relid;
	BlockNumber nblocks;		/* invalidation stamp: relation size ... */
	uint64		version;		/* ... and the index version the AM bumps on every insert / vacuum */
	int			elem;			/* VB_VECTOR / VB_HALFVEC / VB_BIT */
	int			metric;
	int			dimensions;
	int			lists;
	BlockNumber *startPages;	/* per list, like IvfflatScanList.startPage */
	vb_ivf	   *ix;
}			VbIvfImage;

extern VbIvfImage * VbIvfGetImage(Relation index, FmgrInfo *procinfo, int dimensions);
extern void VbIvfInvalidate(Oid relid);

/* ---- HNSW (vb_hnsw_scan.c) ---- */

typedef struct VbHnswImage
{
	Oid			relid;
	BlockNumber nblocks;
	uint64		version;
	int			elem;
	int			metric;
	int			dimensions;
	int			m;
	int64		n;
	ItemPointerData *heaptids;	/* [n][HNSW_HEAPTIDS] */
	uint8	   *nheaptids;		/* [n] */
	vb_hnsw    *ix;
}			VbHnswImage;

extern VbHnswImage * VbHnswGetImage(Relation index, FmgrInfo *procinfo);
extern void VbHnswInvalidate(Oid relid);

/* element type / metric of an opclass, from the support function the index resolved */
extern int	VbMetricFromProc(FmgrInfo *procinfo, int *elem);

/* scan-local state kept beside IvfflatScanOpaqueData (see INTEGRATION.md for where the patch hangs it) */
typedef struct VbIvfScanState
{
	VbIvfImage *image;
	int32	   *lists;			/* nearest-first list numbers [maxProbes] */
	int			nlists;
	int64	   *ids;			/* current batch, sorted by distance */
	double	   *distances;
	int64		nitems;
	int64		next;
}			VbIvfScanState;

extern void VbGetScanLists(IndexScanDesc scan, Datum value, VbIvfScanState * st);
extern void VbGetScanItems(IndexScanDesc scan, Datum value, VbIvfScanState * st);
extern bool VbNextItem(IndexScanDesc scan, VbIvfScanState * st);

/* scan-local results of the HNSW scan: elements nearest first, each expanded into its heap TIDs */
typedef struct VbHnswScanState
{
	VbHnswImage *image;
	int64	   *elements;
	double	   *distances;
	int			nelements;
	int			cur;			/* current element */
	int			curtid;			/* heap TIDs of the current element still to return (counts down) */
	/* hnsw.iterative_scan (src/hnswscan.c:62-87, 236-262): the library keeps v / discarded / tuples in this handle */
	vb_hnsw_scan *iter;
	int			batch;			/* ef_search: elements per batch */
	double		previousDistance;	/* strict_order (src/hnswscan.c:316-322) */
}			VbHnswScanState;

/* false: this scan must be served by the reference's CPU loop (NULL query, or a graph the image cannot represent) */
extern bool VbHnswGetScanItems(IndexScanDesc scan, Datum value, int ef_search, VbHnswScanState * st);
extern bool VbHnswNextItem(IndexScanDesc scan, VbHnswScanState * st);
extern void VbHnswEndScan(VbHnswScanState * st);	/* hnswrescan / hnswendscan: releases the iterative scan's handle */

#endif
