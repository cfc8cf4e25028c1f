/* harness_ivf.c -- drives the IVFFlat glue the way ivfflatgettuple / the build callback do (test infrastructure) */
#include "harness_common.h"

#include "ivfflat.h"
#include "vb_glue.h"

extern long pgstub_buffer_reads(void);

/* ---- scan: the body of ivfflatgettuple (src/ivfscan.c:360-414) with the two GPU calls patched in ---- */
int
h_ivf_scan(HRelation * h, int elem, const void *query_payload, int short_header, int probes, int max_probes, int64 max_items,
		   int64 *out_tids, int64 *n_out, int64 *n_batches)
{
	H_TRAP({
		IndexScanDescData scan;
		IvfflatScanOpaqueData so;
		VbIvfScanState st;
		Datum		value = (Datum) 0;
		int64		n = 0;

		memset(&scan, 0, sizeof(scan));
		memset(&so, 0, sizeof(so));
		memset(&st, 0, sizeof(st));
		scan.indexRelation = &h->rel;
		scan.opaque = &so;
		so.procinfo = index_getprocinfo(&h->rel, 1, IVFFLAT_DISTANCE_PROC);
		so.dimensions = h->desc.attrs[0].atttypmod;
		so.probes = probes;
		so.maxProbes = max_probes;
		so.listPages = palloc(sizeof(BlockNumber) * (Size) max_probes);
		so.listIndex = 0;
		if (query_payload != NULL)
			value = PointerGetDatum(h_make_datum(elem, so.dimensions, query_payload, short_header));
		*n_batches = 0;
		VbGetScanLists(&scan, value, &st);
		VbGetScanItems(&scan, value, &st);
		(*n_batches)++;
		while (n < max_items)
		{
			if (!VbNextItem(&scan, &st))
			{
				/* iterative scan (src/ivfscan.c:400-406): next batch of lists until maxProbes is reached */
				if (so.listIndex >= so.maxProbes || so.listIndex >= st.nlists)
					break;
				VbGetScanItems(&scan, value, &st);
				(*n_batches)++;
				continue;
			}
			out_tids[n++] = VbTidToId(&scan.xs_heaptid);
		}
		*n_out = n;
	});
}

/* what the packer produced: lists, start pages; forces (re)packing through VbIvfGetImage */
int
h_ivf_image(HRelation * h, int *lists, uint32 *start_pages, int max_lists, long *reads, void **image_out)
{
	H_TRAP({
		long		before = pgstub_buffer_reads();
		VbIvfImage *img = VbIvfGetImage(&h->rel, index_getprocinfo(&h->rel, 1, IVFFLAT_DISTANCE_PROC), h->desc.attrs[0].atttypmod);

		*lists = img->lists;
		for (int i = 0; i < img->lists && i < max_lists; i++)
			start_pages[i] = img->startPages[i];
		*reads = pgstub_buffer_reads() - before;
		*image_out = img->ix;
	});
}

int
h_bump_version(HRelation * h)
{
	H_TRAP({ VbBumpIndexVersion(&h->rel); });
}

void
h_ivf_invalidate(HRelation * h)
{
	VbIvfInvalidate(h->rel.rd_id);
}

/* ---- build: IvfflatKmeans' GPU body and the assign batch of the build callback ---- */
extern bool VbIvfflatKmeans(Relation index, VectorArray samples, VectorArray centers, const IvfflatTypeInfo * typeInfo);
typedef struct VbAssignBatch VbAssignBatch;
extern VbAssignBatch *VbAssignBegin(IvfflatBuildState * buildstate, int capacity);
extern void VbAssignAdd(VbAssignBatch * b, ItemPointer tid, Datum value);
extern void VbAssignFlush(VbAssignBatch * b, IvfflatBuildState * buildstate);
extern Tuplesortstate *pgstub_tuplesort_begin(void);
extern int	pgstub_tuplesort_count(Tuplesortstate *st);
extern int32 pgstub_tuplesort_list(Tuplesortstate *st, int i);
extern ItemPointer pgstub_tuplesort_tid(Tuplesortstate *st, int i);

static VectorArray
h_vector_array(int elem, int dim, const char *payloads, int n, int maxlen)
{
	Size		rb = elem == 0 ? 4 * (Size) dim : elem == 1 ? 2 * (Size) dim : ((Size) dim + 7) / 8;
	Size		itemsize = MAXALIGN(8 + rb);
	VectorArray a = palloc0(sizeof(VectorArrayData));

	a->length = n;
	a->maxlen = maxlen;
	a->dim = dim;
	a->itemsize = itemsize;
	a->items = palloc0(itemsize * (Size) maxlen);
	for (int i = 0; i < n; i++)
	{
		char	   *d = h_make_datum(elem, dim, payloads + rb * (Size) i, 0);

		memcpy(a->items + itemsize * (Size) i, d, 8 + rb);
		pfree(d);
	}
	return a;
}

int
h_ivf_kmeans(HRelation * h, int elem, int dim, const char *samples, int n, int lists, char *centers_out, int *used_gpu)
{
	H_TRAP({
		Size		rb = elem == 0 ? 4 * (Size) dim : elem == 1 ? 2 * (Size) dim : ((Size) dim + 7) / 8;
		VectorArray s = h_vector_array(elem, dim, samples, n, n);
		VectorArray c = h_vector_array(elem, dim, NULL, 0, lists);

		*used_gpu = VbIvfflatKmeans(&h->rel, s, c, NULL) ? 1 : 0;
		for (int i = 0; i < c->length; i++)
			memcpy(centers_out + rb * (Size) i, VectorArrayGet(c, i) + 8, rb);
	});
}

int
h_ivf_assign(HRelation * h, int elem, int dim, const char *rows, int n, const char *centers, int lists, int batch,
			 int32 *out_lists, int64 *out_tids)
{
	H_TRAP({
		Size		rb = elem == 0 ? 4 * (Size) dim : elem == 1 ? 2 * (Size) dim : ((Size) dim + 7) / 8;
		IvfflatBuildState bs;
		TupleTableSlot slot;
		Datum		vals[3];
		bool		nulls[3];
		VbAssignBatch *b;

		memset(&bs, 0, sizeof(bs));
		bs.index = &h->rel;
		bs.dimensions = dim;
		bs.procinfo = index_getprocinfo(&h->rel, 1, IVFFLAT_DISTANCE_PROC);
		bs.centers = h_vector_array(elem, dim, centers, lists, lists);
		bs.sortstate = pgstub_tuplesort_begin();
		slot.tts_values = vals;
		slot.tts_isnull = nulls;
		bs.slot = &slot;
		b = VbAssignBegin(&bs, batch);
		for (int i = 0; i < n; i++)
		{
			ItemPointerData tid;
			char	   *d = h_make_datum(elem, dim, rows + rb * (Size) i, 0);

			ItemPointerSet(&tid, (BlockNumber) (i / 100), (OffsetNumber) (i % 100 + 1));
			VbAssignAdd(b, &tid, PointerGetDatum(d));
			pfree(d);
			if ((i + 1) % batch == 0)
				VbAssignFlush(b, &bs);
		}
		VbAssignFlush(b, &bs);
		if (pgstub_tuplesort_count(bs.sortstate) != n)
			elog(ERROR, "harness: %d tuples reached the sort, %d were added", pgstub_tuplesort_count(bs.sortstate), n);
		for (int i = 0; i < n; i++)
		{
			out_lists[i] = pgstub_tuplesort_list(bs.sortstate, i);
			out_tids[i] = VbTidToId(pgstub_tuplesort_tid(bs.sortstate, i));
		}
	});
}
