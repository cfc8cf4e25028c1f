/*
 * vb_glue.h -- extension-side glue between pgvector's index access methods and
 * libvecb200 (include/vecb200.h).  These files are compiled INSIDE the pgvector
 * source tree (added to OBJS in its Makefile); they use the
 * PostgreSQL API and the reference's own headers (ivfflat.h, hnsw.h) and keep
 * the SQL / index-AM surface unchanged.
 *
 * PostgreSQL is not present in the development image, so these files are only
 * syntax-checked there (tests/test_ext_glue.py, against pgstub/ + the reference's
 * real headers); they have not been run against a server.
 */
#ifndef VB_GLUE_H
#define VB_GLUE_H

#include "postgres.h"

#include "access/genam.h"
#include "storage/block.h"
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

/* ---- IVFFlat (vb_ivfflat_scan.c, vb_ivfflat_build.c) ---- */

typedef struct VbIvfImage
{
	Oid			relid;
	BlockNumber nblocks;		/* invalidation stamp, see INTEGRATION.md */
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

#endif
