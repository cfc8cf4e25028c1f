/*
 * pgstub_runtime.c -- the handful of PostgreSQL server functions the glue (the pgvector_b200/ext sources) calls,
 * implemented over in-memory page images so the glue can be RUN by the test harness (tests/harness).
 * TEST INFRASTRUCTURE: not PostgreSQL code, never shipped.  Layouts are the server's on-disk ones
 * (bufpage.h PageHeaderData / ItemIdData, itup.h IndexTupleData, varatt.h varlena headers), which is the point:
 * the harness feeds byte-exact index pages to the packers.
 */
#include "postgres.h"

#include <stdarg.h>
#include <stdio.h>

/* ------------------------------------------------------------------ memory contexts (malloc-backed) */

typedef struct StubChunk
{
	struct StubChunk *prev,
			   *next;
	struct MemoryContextData *owner;
	Size		size;
}			StubChunk;

struct MemoryContextData
{
	StubChunk	head;			/* circular list of chunks */
	struct MemoryContextData *parent,
			   *firstchild,
			   *nextchild;
	Size		allocated;
};

static struct MemoryContextData topContext = {{&topContext.head, &topContext.head, &topContext, 0}, NULL, NULL, NULL, 0};
MemoryContext TopMemoryContext = &topContext;
MemoryContext CurrentMemoryContext = &topContext;

void *
MemoryContextAlloc(MemoryContext cx, Size size)
{
	StubChunk  *c = malloc(sizeof(StubChunk) + (size ? size : 1));

	if (c == NULL)
		elog(ERROR, "out of memory");
	c->owner = cx;
	c->size = size;
	c->next = cx->head.next;
	c->prev = &cx->head;
	cx->head.next->prev = c;
	cx->head.next = c;
	cx->allocated += size;
	return c + 1;
}

void *
MemoryContextAllocZero(MemoryContext cx, Size size)
{
	void	   *p = MemoryContextAlloc(cx, size);

	memset(p, 0, size);
	return p;
}

void	   *MemoryContextAllocHuge(MemoryContext cx, Size size) { return MemoryContextAlloc(cx, size); }
void	   *palloc(Size size) { return MemoryContextAlloc(CurrentMemoryContext, size); }
void	   *palloc0(Size size) { return MemoryContextAllocZero(CurrentMemoryContext, size); }
void	   *palloc_extended(Size size, int flags) { (void) flags; return palloc(size); }

void
pfree(void *p)
{
	StubChunk  *c = (StubChunk *) p - 1;

	c->prev->next = c->next;
	c->next->prev = c->prev;
	c->owner->allocated -= c->size;
	free(c);
}

void *
repalloc(void *p, Size size)
{
	StubChunk  *c = (StubChunk *) p - 1;
	void	   *n = MemoryContextAlloc(c->owner, size);

	memcpy(n, p, c->size < size ? c->size : size);
	pfree(p);
	return n;
}

void	   *repalloc_huge(void *p, Size size) { return repalloc(p, size); }

MemoryContext
AllocSetContextCreateInternal(MemoryContext parent, const char *name, Size a, Size b, Size c)
{
	MemoryContext cx = malloc(sizeof(struct MemoryContextData));

	(void) name; (void) a; (void) b; (void) c;
	cx->head.next = cx->head.prev = &cx->head;
	cx->head.owner = cx;
	cx->head.size = 0;
	cx->parent = parent;
	cx->firstchild = NULL;
	cx->nextchild = parent->firstchild;
	parent->firstchild = cx;
	cx->allocated = 0;
	return cx;
}

void
MemoryContextReset(MemoryContext cx)
{
	while (cx->head.next != &cx->head)
		pfree(cx->head.next + 1);
}

void
MemoryContextDelete(MemoryContext cx)
{
	MemoryContext *pp;

	while (cx->firstchild)
		MemoryContextDelete(cx->firstchild);
	MemoryContextReset(cx);
	for (pp = &cx->parent->firstchild; *pp; pp = &(*pp)->nextchild)
		if (*pp == cx)
		{
			*pp = cx->nextchild;
			break;
		}
	if (CurrentMemoryContext == cx)
		CurrentMemoryContext = cx->parent;
	free(cx);
}

Size		MemoryContextMemAllocated(MemoryContext cx, bool recurse) { (void) recurse; return cx->allocated; }

/* ------------------------------------------------------------------ errors */

sigjmp_buf *PG_exception_stack = NULL;
static char stub_errmsg[1024];
static int	stub_errors = 0;

const char *pgstub_last_error(void) { return stub_errmsg; }
int			pgstub_error_count(void) { return stub_errors; }

int			errcode(int c) { return c; }
int			errdetail(const char *fmt,...) { (void) fmt; return 0; }
int			errhint(const char *fmt,...) { (void) fmt; return 0; }

int
errmsg(const char *fmt,...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(stub_errmsg, sizeof(stub_errmsg), fmt, ap);
	va_end(ap);
	return 0;
}

void
pg_re_throw(void)
{
	if (PG_exception_stack != NULL)
		siglongjmp(*PG_exception_stack, 1);
	fprintf(stderr, "pgstub: unhandled ERROR: %s\n", stub_errmsg);
	abort();
}

void
vb_stub_ereport(int level,...)
{
	if (level >= ERROR)
	{
		stub_errors++;
		pg_re_throw();
	}
}

void
elog(int level, const char *fmt,...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(stub_errmsg, sizeof(stub_errmsg), fmt, ap);
	va_end(ap);
	vb_stub_ereport(level, 0);
}

/* ------------------------------------------------------------------ datums */

Datum
Float8GetDatum(float8 x)
{
	Datum		d;

	memcpy(&d, &x, sizeof(d));
	return d;
}

float8
DatumGetFloat8(Datum d)
{
	float8		x;

	memcpy(&x, &d, sizeof(x));
	return x;
}

/* a 1-byte-header varlena is expanded into a palloc'd 4-byte-header copy (detoast_attr's short-header branch) */
struct varlena *
pg_detoast_datum(struct varlena *datum)
{
	if (VARATT_IS_SHORT(datum))
	{
		Size		data_size = VARSIZE_SHORT(datum) - VARHDRSZ_SHORT;
		Size		new_size = data_size + VARHDRSZ;
		struct varlena *result = palloc(new_size);

		SET_VARSIZE(result, new_size);
		memcpy((char *) result + VARHDRSZ, VARDATA_SHORT(datum), data_size);
		return result;
	}
	return datum;
}

Datum
datumCopy(Datum value, bool typByVal, int typLen)
{
	struct varlena *v = (struct varlena *) DatumGetPointer(value);
	Size		size;
	void	   *copy;

	(void) typByVal; (void) typLen;
	size = VARSIZE_ANY(v);
	copy = palloc(size);
	memcpy(copy, v, size);
	return PointerGetDatum(copy);
}

bool
datumIsEqual(Datum a, Datum b, bool typByVal, int typLen)
{
	struct varlena *x = (struct varlena *) DatumGetPointer(a), *y = (struct varlena *) DatumGetPointer(b);

	(void) typByVal; (void) typLen;
	return VARSIZE_ANY(x) == VARSIZE_ANY(y) && memcmp(x, y, VARSIZE_ANY(x)) == 0;
}

/* common/hashfn.h */
uint64
murmurhash64(uint64 h)
{
	h ^= h >> 33;
	h *= UINT64CONST(0xff51afd7ed558ccd);
	h ^= h >> 33;
	h *= UINT64CONST(0xc4ceb9fe1a85ec53);
	h ^= h >> 33;
	return h;
}

/* ------------------------------------------------------------------ string buffers */

void
initStringInfo(StringInfo str)
{
	str->maxlen = 1024;
	str->data = palloc(str->maxlen);
	str->len = 0;
	str->cursor = 0;
	str->data[0] = '\0';
}

/* (the server caps a StringInfo at 1 GB; the glue packs larger indexes in pieces in a real build) */
void
appendBinaryStringInfo(StringInfo str, const void *data, int datalen)
{
	if ((Size) str->len + (Size) datalen + 1 > (Size) str->maxlen)
	{
		Size		n = (Size) str->maxlen;

		while (n < (Size) str->len + (Size) datalen + 1)
			n *= 2;
		if (n > 0x7fffffff)
			elog(ERROR, "out of memory: string buffer exceeds 2 GB");
		str->data = repalloc(str->data, n);
		str->maxlen = (int) n;
	}
	memcpy(str->data + str->len, data, datalen);
	str->len += datalen;
	str->data[str->len] = '\0';
}

/* ------------------------------------------------------------------ relations, buffers, pages */

#define STUB_MAX_PINS 64
static Page stub_pins[STUB_MAX_PINS];
static int	stub_pin_count = 0;
static long stub_reads = 0;

long		pgstub_buffer_reads(void) { return stub_reads; }
int			pgstub_pinned_buffers(void) { return stub_pin_count; }

BlockNumber RelationGetNumberOfBlocksInFork(Relation rel, ForkNumber fork) { (void) fork; return rel->stub_nblocks; }

Buffer
ReadBufferExtended(Relation rel, ForkNumber fork, BlockNumber blkno, ReadBufferMode mode, BufferAccessStrategy strategy)
{
	(void) fork; (void) mode; (void) strategy;
	if (blkno >= rel->stub_nblocks)
		elog(ERROR, "could not read block %u: relation has %u blocks", blkno, rel->stub_nblocks);
	for (int i = 0; i < STUB_MAX_PINS; i++)
		if (stub_pins[i] == NULL)
		{
			stub_pins[i] = rel->stub_pages + (Size) blkno * BLCKSZ;
			stub_pin_count++;
			stub_reads++;
			return i + 1;
		}
	elog(ERROR, "pgstub: too many pinned buffers (a buffer was not released)");
	return 0;
}

Buffer		ReadBuffer(Relation rel, BlockNumber blkno) { return ReadBufferExtended(rel, MAIN_FORKNUM, blkno, RBM_NORMAL, NULL); }
void		LockBuffer(Buffer buf, int mode) { (void) buf; (void) mode; }
Page		BufferGetPage(Buffer buf) { return stub_pins[buf - 1]; }

void
UnlockReleaseBuffer(Buffer buf)
{
	stub_pins[buf - 1] = NULL;
	stub_pin_count--;
}

/* bufpage.h: pd_lower at byte 12, pd_upper 14, pd_special 16; line pointers from byte 24 */
static inline uint16 page_u16(Page page, int off) { uint16 v; memcpy(&v, page + off, 2); return v; }

OffsetNumber
PageGetMaxOffsetNumber(Page page)
{
	uint16		lower = page_u16(page, 12);

	return lower <= SizeOfPageHeaderData ? 0 : (OffsetNumber) ((lower - SizeOfPageHeaderData) / sizeof(ItemIdData));
}

ItemId		PageGetItemId(Page page, OffsetNumber offno) { return (ItemId) (page + SizeOfPageHeaderData) + (offno - 1); }
void	   *PageGetItem(Page page, ItemId itemId) { return page + itemId->lp_off; }
char	   *PageGetSpecialPointer(Page page) { return page + page_u16(page, 16); }

Size
PageGetFreeSpace(Page page)
{
	int			space = (int) page_u16(page, 14) - (int) page_u16(page, 12);

	return space < (int) sizeof(ItemIdData) ? 0 : (Size) space - sizeof(ItemIdData);
}

Form_pg_attribute TupleDescAttr(TupleDesc desc, int i) { return &desc->attrs[i]; }

/*
 * itup.h index_getattr for the one-column indexes of ivfflat / hnsw: data starts at the MAXALIGN'd end of the
 * IndexTupleData header (8 bytes), after the null bitmap when INDEX_NULL_MASK (0x8000) is set.  The first attribute
 * needs no alignment padding whether it carries a 1-byte or a 4-byte varlena header.
 */
Datum
index_getattr(IndexTuple tup, int attnum, TupleDesc desc, bool *isnull)
{
	(void) attnum; (void) desc;
	if (tup->t_info & 0x8000)
	{
		uint8		bits = *((uint8 *) tup + sizeof(IndexTupleData));

		if (!(bits & 1))
		{
			*isnull = true;
			return (Datum) 0;
		}
		*isnull = false;
		return PointerGetDatum((char *) tup + MAXALIGN(sizeof(IndexTupleData) + 1));
	}
	*isnull = false;
	return PointerGetDatum((char *) tup + MAXALIGN(sizeof(IndexTupleData)));
}

FmgrInfo   *index_getprocinfo(Relation rel, AttrNumber attnum, uint16 procnum) { (void) attnum; return rel->stub_procs[procnum]; }
Oid			index_getprocid(Relation rel, AttrNumber attnum, uint16 procnum) { (void) attnum; return rel->stub_procs[procnum] ? rel->stub_procs[procnum]->fn_oid : InvalidOid; }

/* ------------------------------------------------------------------ executor / tuplesort (the build glue feeds one) */

TupleTableSlot *ExecClearTuple(TupleTableSlot *slot) { return slot; }
TupleTableSlot *ExecStoreVirtualTuple(TupleTableSlot *slot) { return slot; }

struct Tuplesortstate
{
	int			n,
				cap;
	int32	   *lists;
	ItemPointerData *tids;
	Datum	   *values;
};

Tuplesortstate *
pgstub_tuplesort_begin(void)
{
	Tuplesortstate *st = MemoryContextAllocZero(TopMemoryContext, sizeof(Tuplesortstate));

	st->cap = 1024;
	st->lists = MemoryContextAlloc(TopMemoryContext, sizeof(int32) * st->cap);
	st->tids = MemoryContextAlloc(TopMemoryContext, sizeof(ItemPointerData) * st->cap);
	st->values = MemoryContextAlloc(TopMemoryContext, sizeof(Datum) * st->cap);
	return st;
}

void
tuplesort_puttupleslot(Tuplesortstate *st, TupleTableSlot *slot)
{
	if (st->n == st->cap)
	{
		st->cap *= 2;
		st->lists = repalloc(st->lists, sizeof(int32) * st->cap);
		st->tids = repalloc(st->tids, sizeof(ItemPointerData) * st->cap);
		st->values = repalloc(st->values, sizeof(Datum) * st->cap);
	}
	st->lists[st->n] = DatumGetInt32(slot->tts_values[0]);
	st->tids[st->n] = *(ItemPointer) DatumGetPointer(slot->tts_values[1]);
	st->values[st->n] = slot->tts_values[2];
	st->n++;
}

int			pgstub_tuplesort_count(Tuplesortstate *st) { return st->n; }
int32		pgstub_tuplesort_list(Tuplesortstate *st, int i) { return st->lists[i]; }
ItemPointer pgstub_tuplesort_tid(Tuplesortstate *st, int i) { return &st->tids[i]; }

/* pg_prng: the glue only seeds the library with it */
pg_prng_state pg_global_prng_state = {0x9e3779b97f4a7c15ULL, 0xbf58476d1ce4e5b9ULL};

uint32
pg_prng_uint32(pg_prng_state *s)
{
	uint64		z = (s->s0 += 0x9e3779b97f4a7c15ULL);

	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return (uint32) ((z ^ (z >> 31)) >> 32);
}

double		pg_prng_double(pg_prng_state *s) { return (double) pg_prng_uint32(s) / 4294967296.0; }
void		pg_prng_seed(pg_prng_state *s, uint64 seed) { s->s0 = seed; s->s1 = ~seed; }
int			work_mem = 4096, maintenance_work_mem = 65536;
