/* harness_hnsw.c -- drives the HNSW glue the way hnswgettuple does (test infrastructure) */
#include "harness_common.h"

#include "hnsw.h"
#include "vb_glue.h"

/* HnswGetMetaPageInfo / HnswInitElementFromBlock live in the reference's hnswutils.c (src/hnswutils.c:283-328), which
 * cannot be compiled here; the two page reads they do are restated for the harness */
HnswElement
HnswInitElementFromBlock(BlockNumber blkno, OffsetNumber offno)
{
	HnswElement element = palloc0(sizeof(HnswElementData));

	element->blkno = blkno;
	element->offno = offno;
	return element;
}

void
HnswGetMetaPageInfo(Relation index, int *m, HnswElement * entryPoint)
{
	Buffer		buf = ReadBuffer(index, HNSW_METAPAGE_BLKNO);
	Page		page;
	HnswMetaPage metap;

	LockBuffer(buf, BUFFER_LOCK_SHARE);
	page = BufferGetPage(buf);
	metap = HnswPageGetMeta(page);
	if (metap->magicNumber != HNSW_MAGIC_NUMBER)
		elog(ERROR, "hnsw index is not valid");
	if (m != NULL)
		*m = metap->m;
	if (entryPoint != NULL)
	{
		if (BlockNumberIsValid(metap->entryBlkno))
		{
			*entryPoint = HnswInitElementFromBlock(metap->entryBlkno, metap->entryOffno);
			(*entryPoint)->level = metap->entryLevel;
		}
		else
			*entryPoint = NULL;
	}
	UnlockReleaseBuffer(buf);
}

/* the loop of hnswgettuple (src/hnswscan.c:189-331) with the GPU call patched in; -2 = "serve on the CPU path" */
int
h_hnsw_scan(HRelation * h, int elem, const void *query_payload, int ef_search, int64 max_items, int64 *out_tids, int64 *n_out,
			int64 *tuples)
{
	H_TRAP({
		IndexScanDescData scan;
		HnswScanOpaqueData so;
		VbHnswScanState st;
		Datum		value = (Datum) 0;
		int64		n = 0;

		memset(&scan, 0, sizeof(scan));
		memset(&so, 0, sizeof(so));
		memset(&st, 0, sizeof(st));
		scan.indexRelation = &h->rel;
		scan.opaque = &so;
		so.support.procinfo = index_getprocinfo(&h->rel, 1, HNSW_DISTANCE_PROC);
		if (query_payload != NULL)
			value = PointerGetDatum(h_make_datum(elem, h->desc.attrs[0].atttypmod, query_payload, 0));
		*n_out = -2;
		if (VbHnswGetScanItems(&scan, value, ef_search, &st))
		{
			while (n < max_items && VbHnswNextItem(&scan, &st))
				out_tids[n++] = VbTidToId(&scan.xs_heaptid);
			*n_out = n;
			*tuples = so.tuples;
		}
	});
}

void
h_hnsw_invalidate(HRelation * h)
{
	VbHnswInvalidate(h->rel.rd_id);
}
