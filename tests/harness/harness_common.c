/* harness_common.c -- relation / datum construction shared by the IVFFlat and HNSW harness drivers (test infrastructure) */
#include "harness_common.h"

#define DUMMY(name) Datum name(PG_FUNCTION_ARGS) { (void) fcinfo; return (Datum) __LINE__; }
DUMMY(vector_l2_squared_distance)
DUMMY(vector_negative_inner_product)
DUMMY(l1_distance)
DUMMY(halfvec_l2_squared_distance)
DUMMY(halfvec_negative_inner_product)
DUMMY(halfvec_l1_distance)
DUMMY(hamming_distance)
DUMMY(jaccard_distance)
DUMMY(l2_distance)
DUMMY(vector_spherical_distance)
DUMMY(halfvec_l2_distance)
DUMMY(halfvec_spherical_distance)

static PGFunction h_procs[] = {
	vector_l2_squared_distance, vector_negative_inner_product, l1_distance,
	halfvec_l2_squared_distance, halfvec_negative_inner_product, halfvec_l1_distance,
	hamming_distance, jaccard_distance,
	l2_distance, vector_spherical_distance, halfvec_l2_distance, halfvec_spherical_distance
};

/* proc1: index into h_procs for the opclass's distance proc; proc3 (>= 0): its k-means distance proc */
HRelation *
h_open(char *pages, uint32 nblocks, Oid relid, int dim, int proc1, int proc3)
{
	HRelation  *h = MemoryContextAllocZero(TopMemoryContext, sizeof(HRelation));

	h->rel.rd_id = relid;
	h->rel.rd_indcollation = &h->collation;
	h->rel.rd_att = &h->desc;
	h->rel.stub_pages = pages;
	h->rel.stub_nblocks = nblocks;
	h->desc.natts = 1;
	h->desc.attrs[0].atttypmod = dim;
	h->procs[1].fn_addr = h_procs[proc1];
	h->rel.stub_procs[1] = &h->procs[1];
	if (proc3 >= 0)
	{
		h->procs[3].fn_addr = h_procs[proc3];
		h->rel.stub_procs[3] = &h->procs[3];
	}
	return h;
}

/* a vector / halfvec / bit datum around a raw payload; short_header: the 1-byte varlena form a heap / index tuple stores
 * for values of at most 126 bytes (varatt.h) */
void *
h_make_datum(int elem, int dim, const void *payload, int short_header)
{
	Size		data = elem == 0 ? 4 + 4 * (Size) dim : elem == 1 ? 4 + 2 * (Size) dim : 4 + ((Size) dim + 7) / 8;
	char	   *full = palloc0(VARHDRSZ + data);

	SET_VARSIZE(full, VARHDRSZ + data);
	if (elem == 2)
		memcpy(full + 4, &dim, 4);			/* VarBit.bit_len */
	else
	{
		int16		d16 = (int16) dim;

		memcpy(full + 4, &d16, 2);			/* dim, unused = 0 */
	}
	memcpy(full + 8, payload, data - 4);
	if (short_header && data + 1 <= 127)
	{
		char	   *s = palloc(data + 1);

		s[0] = (char) (((data + 1) << 1) | 1);
		memcpy(s + 1, full + VARHDRSZ, data);
		return s;
	}
	return full;
}
