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

/* the GUCs of src/hnsw.c:28-37, 96-105 the glue reads */
int			hnsw_iterative_scan = HNSW_ITERATIVE_SCAN_OFF;
int			hnsw_max_scan_tuples = 20000;

void
h_hnsw_set_iterative(int mode, int max_scan_tuples)
{
	hnsw_iterative_scan = mode;
	hnsw_max_scan_tuples = max_scan_tuples;
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
			VbHnswEndScan(&st);		/* hnswendscan */
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

/* ---- build: the in-memory phase through VbHnswBuildAdd / VbHnswBuildFinish ----
 * HnswAlloc / HnswInitNeighborArray / HnswInitNeighbors / HnswAddHeapTid live in the reference's hnswutils.c
 * (src/hnswutils.c:201-235, 273-277); restated for the harness like the meta page reader above */
void	   *HnswAlloc(HnswAllocator * allocator, Size size) { return allocator ? (*allocator->alloc) (size, allocator->state) : palloc(size); }

HnswNeighborArray *
HnswInitNeighborArray(int lm, HnswAllocator * allocator)
{
	HnswNeighborArray *a = HnswAlloc(allocator, HNSW_NEIGHBOR_ARRAY_SIZE(lm));

	a->length = 0;
	a->closerSet = false;
	return a;
}

void
HnswInitNeighbors(char *base, HnswElement element, int m, HnswAllocator * allocator)
{
	int			level = element->level;
	HnswNeighborArrayPtr *neighborList = (HnswNeighborArrayPtr *) HnswAlloc(allocator, sizeof(HnswNeighborArrayPtr) * ((Size) level + 1));

	HnswPtrStore(base, element->neighbors, neighborList);
	for (int lc = 0; lc <= level; lc++)
		HnswPtrStore(base, neighborList[lc], HnswInitNeighborArray(HnswGetLayerM(m, lc), allocator));
}

void		HnswAddHeapTid(HnswElement element, ItemPointer heaptid) { element->heaptids[element->heaptidsLength++] = *heaptid; }

typedef struct VbHnswBuildBuffer VbHnswBuildBuffer;
extern VbHnswBuildBuffer *VbHnswBuildBegin(HnswBuildState * buildstate);
extern bool VbHnswBuildAdd(VbHnswBuildBuffer * b, ItemPointer tid, Datum value);
extern void VbHnswBuildFinish(VbHnswBuildBuffer * b, HnswBuildState * buildstate);

static void *harness_alloc(Size size, void *state) { (void) state; return MemoryContextAllocZero(TopMemoryContext, size); }

/*
 * Feed n rows (heap TID of row i = harness numbering) through the build glue, then walk the in-memory graph it
 * left behind the way CreateGraphPages / HnswSetNeighborTuple do (graph->head chain, neighbour arrays).  Out, per
 * ELEMENT in head order (newest first): level, heap TIDs (first = the element's own row), layer-0 neighbours and
 * upper-layer neighbours as the neighbour's own first heap TID.  *n_added = rows the buffer accepted.
 */
int
h_hnsw_build(HRelation * h, int elem, int dim, const char *rows, int64 n, int m, int ef_construction, int64 memory_total, int64 *n_added,
			 int64 *n_elements, int32 *levels, int64 *heaptids /* [n][10] */ , int32 *n_heaptids, int64 *nbr0 /* [n][2m] */ ,
			 int64 *upper /* [n][max_level][m] */ , int max_level, int64 *entry_tid)
{
	H_TRAP({
		Size		rb = elem == 0 ? 4 * (Size) dim : elem == 1 ? 2 * (Size) dim : ((Size) dim + 7) / 8;
		HnswBuildState bs;
		HnswGraph	graph;
		VbHnswBuildBuffer *b;
		int64		added = 0, ne = 0;
		HnswElement e;

		memset(&bs, 0, sizeof(bs));
		memset(&graph, 0, sizeof(graph));
		bs.index = &h->rel;
		bs.dimensions = dim;
		bs.m = m;
		bs.efConstruction = ef_construction;
		bs.ml = HnswGetMl(m);
		bs.maxLevel = HnswGetMaxLevel(m);
		bs.support.procinfo = index_getprocinfo(&h->rel, 1, HNSW_DISTANCE_PROC);
		bs.graph = &graph;
		bs.hnswarea = NULL;
		bs.allocator.alloc = harness_alloc;
		bs.allocator.state = NULL;
		graph.memoryTotal = (Size) memory_total;
		b = VbHnswBuildBegin(&bs);
		for (int64 i = 0; i < n; i++)
		{
			ItemPointerData tid;
			char	   *d = h_make_datum(elem, dim, rows + rb * (Size) i, 0);

			ItemPointerSet(&tid, (BlockNumber) (i / 200), (OffsetNumber) (i % 200 + 1));
			if (!VbHnswBuildAdd(b, &tid, PointerGetDatum(d)))
			{
				pfree(d);
				break;
			}
			pfree(d);
			added++;
		}
		VbHnswBuildFinish(b, &bs);
		*n_added = added;
		for (e = graph.head.ptr; e != NULL; e = e->next.ptr)
		{
			levels[ne] = e->level;
			n_heaptids[ne] = e->heaptidsLength;
			for (int j = 0; j < e->heaptidsLength; j++)
				heaptids[ne * 10 + j] = VbTidToId(&e->heaptids[j]);
			for (int lc = 0; lc <= e->level && lc <= max_level; lc++)
			{
				HnswNeighborArray *a = HnswGetNeighbors(NULL, e, lc);
				int			lm = HnswGetLayerM(m, lc);
				int64	   *dst = lc == 0 ? nbr0 + ne * 2 * m : upper + (ne * max_level + (lc - 1)) * m;

				for (int j = 0; j < lm; j++)
					dst[j] = j < a->length ? VbTidToId(&a->items[j].element.ptr->heaptids[0]) : -1;
			}
			ne++;
		}
		*n_elements = ne;
		*entry_tid = graph.entryPoint.ptr ? VbTidToId(&graph.entryPoint.ptr->heaptids[0]) : -1;
		if ((int64) graph.indtuples != added)
			elog(ERROR, "harness: indtuples %g, rows added %ld", graph.indtuples, (long) added);
	});
}
