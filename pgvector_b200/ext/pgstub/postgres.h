/*
 * pgstub/postgres.h -- a stand-in for the PostgreSQL server headers: complete enough to COMPILE
 * the pgvector_b200/ext sources together with the reference's own ivfflat.h / hnsw.h / vector.h / halfvec.h
 * (gcc -Ipgstub -I/root/reference/src), and -- with pgstub_runtime.c, which implements the handful of server
 * functions the glue calls over in-memory page images -- to RUN it in the test harness (tests/harness).
 * It is NOT PostgreSQL code and never ships; the real build uses the server's headers (PGXS).
 * Declarations follow the public PostgreSQL API (PG 17 signatures); the page, item-pointer, index-tuple and
 * varlena layouts are the on-disk ones (little endian), because the harness feeds the glue byte-exact page images.
 */
#ifndef PGSTUB_POSTGRES_H
#define PGSTUB_POSTGRES_H
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#define PG_VERSION_NUM 170000
#define BLCKSZ 8192
#define FLEXIBLE_ARRAY_MEMBER
#define PGDLLEXPORT
#define PGDLLIMPORT
#define pg_attribute_noreturn()
#define HAVE__GET_CPUID 1
#define HAVE__BUILTIN_POPCOUNT 1
#define HAVE_LONG_INT_64 1
#define SIZE_MAX_ (~(size_t)0)

typedef uint8_t uint8; typedef uint16_t uint16; typedef uint32_t uint32; typedef uint64_t uint64;
typedef int8_t int8; typedef int16_t int16; typedef int32_t int32; typedef int64_t int64;
typedef size_t Size; typedef uintptr_t Datum; typedef char *Pointer; typedef unsigned int Oid;
typedef uint32 BlockNumber; typedef uint16 OffsetNumber; typedef int Buffer; typedef char *Page;
typedef uint32 TransactionId; typedef int ForkNumber; typedef uint16 StrategyNumber; typedef int LOCKMODE;
typedef uint64 XLogRecPtr; typedef int ScanDirection; typedef int16 AttrNumber; typedef uint32 bits32;
typedef float float4; typedef double float8; typedef signed int Offset;
typedef struct varlena { char vl_len_[4]; char vl_dat[FLEXIBLE_ARRAY_MEMBER]; } varlena;
#define InvalidOid ((Oid) 0)
#define InvalidBlockNumber ((BlockNumber) 0xFFFFFFFF)
#define InvalidOffsetNumber ((OffsetNumber) 0)
#define FirstOffsetNumber ((OffsetNumber) 1)
#define OffsetNumberNext(o) ((OffsetNumber) (1 + (o)))
#define BlockNumberIsValid(b) ((b) != InvalidBlockNumber)
#define MAIN_FORKNUM 0
#define likely(x) __builtin_expect((x) != 0, 1)
#define unlikely(x) __builtin_expect((x) != 0, 0)
#define Min(a, b) ((a) < (b) ? (a) : (b))
#define Max(a, b) ((a) > (b) ? (a) : (b))
#define Assert(c) ((void) 0)
#define MAXALIGN(x) (((uintptr_t) (x) + 7) & ~(uintptr_t) 7)
#define UINT64CONST(x) UINT64_C(x)
#define INT64_FORMAT "%ld"

/* memory */
typedef struct MemoryContextData *MemoryContext;
extern MemoryContext CurrentMemoryContext, TopMemoryContext;
extern void *palloc(Size); extern void *palloc0(Size); extern void *repalloc(void *, Size); extern void pfree(void *);
extern void *palloc_extended(Size, int); extern void *repalloc_huge(void *, Size);
extern void *MemoryContextAlloc(MemoryContext, Size); extern void *MemoryContextAllocZero(MemoryContext, Size);
extern void *MemoryContextAllocHuge(MemoryContext, Size);
extern MemoryContext AllocSetContextCreateInternal(MemoryContext, const char *, Size, Size, Size);
#define AllocSetContextCreate(p, n, ...) AllocSetContextCreateInternal(p, n, 0, 8192, 8388608)
#define ALLOCSET_DEFAULT_SIZES 0, 8192, 8388608
extern void MemoryContextDelete(MemoryContext); extern void MemoryContextReset(MemoryContext);
extern Size MemoryContextMemAllocated(MemoryContext, bool);
static inline MemoryContext MemoryContextSwitchTo(MemoryContext c) { MemoryContext o = CurrentMemoryContext; CurrentMemoryContext = c; return o; }
#define MCXT_ALLOC_HUGE 0x01
#define palloc_object(type) ((type *) palloc(sizeof(type)))
#define palloc_array(type, n) ((type *) palloc(sizeof(type) * (n)))
static inline Size add_size(Size a, Size b) { return a + b; }
static inline Size mul_size(Size a, Size b) { return a * b; }

/* errors */
#define ERROR 21
#define NOTICE 18
#define INFO 17
#define DEBUG1 14
#define ERRCODE_EXTERNAL_ROUTINE_EXCEPTION 1
#define ERRCODE_DATA_EXCEPTION 2
#define ERRCODE_PROGRAM_LIMIT_EXCEEDED 3
#define ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE 4
extern int errcode(int); extern int errmsg(const char *, ...); extern int errdetail(const char *, ...); extern int errhint(const char *, ...);
extern void vb_stub_ereport(int, ...);
#define ereport(level, rest) vb_stub_ereport(level, rest)
extern void elog(int, const char *, ...);
/* the structure of elog.h's PG_TRY: a chain of sigjmp_bufs; ereport(ERROR) longjmps to the innermost one */
#include <setjmp.h>
extern sigjmp_buf *PG_exception_stack;
extern void pg_re_throw(void);
#define PG_TRY() do { sigjmp_buf *_save_exception_stack = PG_exception_stack; sigjmp_buf _local_sigjmp_buf; bool _do_rethrow = false; \
	if (sigsetjmp(_local_sigjmp_buf, 0) == 0) { PG_exception_stack = &_local_sigjmp_buf
#define PG_CATCH() } else { PG_exception_stack = _save_exception_stack
#define PG_FINALLY() } else _do_rethrow = true; { PG_exception_stack = _save_exception_stack
#define PG_END_TRY() } if (_do_rethrow) pg_re_throw(); PG_exception_stack = _save_exception_stack; } while (0)
#define CHECK_FOR_INTERRUPTS() ((void) 0)

/* datum / fmgr */
#define PointerGetDatum(p) ((Datum) (p))
#define DatumGetPointer(d) ((Pointer) (d))
#define Int32GetDatum(i) ((Datum) (i))
#define DatumGetInt32(d) ((int32) (d))
extern Datum Float8GetDatum(float8); extern float8 DatumGetFloat8(Datum);
typedef struct FunctionCallInfoBaseData *FunctionCallInfo;
typedef Datum (*PGFunction) (FunctionCallInfo fcinfo);
struct FmgrInfo; typedef struct FmgrInfo { PGFunction fn_addr; Oid fn_oid; short fn_nargs; } FmgrInfo;
#define PG_FUNCTION_ARGS FunctionCallInfo fcinfo
#define PG_FUNCTION_INFO_V1(f) extern int pg_finfo_##f
extern Datum PG_GETARG_DATUM_(FunctionCallInfo, int);
#define PG_GETARG_DATUM(n) PG_GETARG_DATUM_(fcinfo, n)
#define PG_GETARG_INT32(n) ((int32) PG_GETARG_DATUM(n))
#define PG_RETURN_POINTER(x) return PointerGetDatum(x)
#define PG_RETURN_FLOAT8(x) return Float8GetDatum(x)
extern struct varlena *pg_detoast_datum(struct varlena *);
#define PG_DETOAST_DATUM(d) pg_detoast_datum((struct varlena *) DatumGetPointer(d))
extern Datum FunctionCall2Coll(FmgrInfo *, Oid, Datum, Datum); extern Datum FunctionCall1Coll(FmgrInfo *, Oid, Datum);
extern Datum FunctionCall0Coll(FmgrInfo *, Oid); extern Datum DirectFunctionCall1Coll(PGFunction, Oid, Datum);
extern Datum datumCopy(Datum, bool, int); extern bool datumIsEqual(Datum, Datum, bool, int);
/* varlena headers, little endian (varatt.h): 4-byte header = length << 2; 1-byte header = (length << 1) | 1 */
#define SET_VARSIZE(p, len) (*(uint32 *) (p) = ((uint32) (len)) << 2)
#define VARSIZE(p) ((*(uint32 *) (p)) >> 2)
#define VARATT_IS_SHORT(p) ((*(const uint8 *) (p) & 0x01) == 0x01)
#define VARSIZE_SHORT(p) ((*(const uint8 *) (p) >> 1) & 0x7F)
#define VARHDRSZ 4
#define VARHDRSZ_SHORT 1
#define VARDATA_SHORT(p) ((char *) (p) + 1)
#define VARSIZE_ANY(p) (VARATT_IS_SHORT(p) ? VARSIZE_SHORT(p) : VARSIZE(p))
#define VARATT_IS_COMPRESSED(p) false
#define VARATT_IS_EXTENDED(p) VARATT_IS_SHORT(p)

/* varbit */
typedef struct { int32 vl_len_; int32 bit_len; uint8 bit_dat[FLEXIBLE_ARRAY_MEMBER]; } VarBit;
#define VARBITS(v) ((v)->bit_dat)
#define VARBITLEN(v) ((v)->bit_len)
#define VARBITBYTES(v) (VARSIZE(v) - 8)
#define VARBITTOTALLEN(n) (((n) + 7) / 8 + 8)
#define DatumGetVarBitP(d) ((VarBit *) PG_DETOAST_DATUM(d))

/* item pointers / pages */
typedef struct BlockIdData { uint16 bi_hi, bi_lo; } BlockIdData;
typedef struct ItemPointerData { BlockIdData ip_blkid; OffsetNumber ip_posid; } ItemPointerData;
typedef ItemPointerData *ItemPointer;
static inline BlockNumber ItemPointerGetBlockNumber(const ItemPointerData *p) { return ((BlockNumber) p->ip_blkid.bi_hi << 16) | p->ip_blkid.bi_lo; }
static inline OffsetNumber ItemPointerGetOffsetNumber(const ItemPointerData *p) { return p->ip_posid; }
static inline void ItemPointerSet(ItemPointerData *p, BlockNumber b, OffsetNumber o) { p->ip_blkid.bi_hi = (uint16) (b >> 16); p->ip_blkid.bi_lo = (uint16) b; p->ip_posid = o; }
static inline void ItemPointerSetInvalid(ItemPointerData *p) { ItemPointerSet(p, InvalidBlockNumber, InvalidOffsetNumber); }
static inline bool ItemPointerIsValid(const ItemPointerData *p) { return p != NULL && p->ip_posid != 0; }
typedef struct ItemIdData { unsigned lp_off:15, lp_flags:2, lp_len:15; } ItemIdData; typedef ItemIdData *ItemId;
typedef struct PageHeaderData { char pad[24]; ItemIdData pd_linp[FLEXIBLE_ARRAY_MEMBER]; } PageHeaderData;
#define SizeOfPageHeaderData 24
#define PageGetContents(page) ((char *) (page) + MAXALIGN(SizeOfPageHeaderData))
extern OffsetNumber PageGetMaxOffsetNumber(Page); extern ItemId PageGetItemId(Page, OffsetNumber);
extern void *PageGetItem(Page, ItemId); extern char *PageGetSpecialPointer(Page); extern Size PageGetFreeSpace(Page);
typedef struct IndexTupleData { ItemPointerData t_tid; unsigned short t_info; } IndexTupleData; typedef IndexTupleData *IndexTuple;
#define MaxHeapTuplesPerPage 291

/* relations, buffers */
typedef struct FormData_pg_attribute { Oid atttypid; int32 atttypmod; } FormData_pg_attribute; typedef FormData_pg_attribute *Form_pg_attribute;
typedef struct TupleDescData { int natts; FormData_pg_attribute attrs[4]; } TupleDescData; typedef struct TupleDescData *TupleDesc;
extern Form_pg_attribute TupleDescAttr(TupleDesc, int);
typedef struct RelationData { Oid rd_id; Oid *rd_indcollation; void *rd_options; TupleDesc rd_att;
	/* pgstub_runtime.c: the relation's pages (nblocks x BLCKSZ bytes) and the support functions the harness installed */
	char *stub_pages; BlockNumber stub_nblocks; struct FmgrInfo *stub_procs[8]; } RelationData; typedef RelationData *Relation;
#define RelationGetRelid(r) ((r)->rd_id)
#define RelationGetDescr(r) ((r)->rd_att)
extern BlockNumber RelationGetNumberOfBlocksInFork(Relation, ForkNumber);
#define RelationGetNumberOfBlocks(r) RelationGetNumberOfBlocksInFork(r, MAIN_FORKNUM)
typedef struct BufferAccessStrategyData *BufferAccessStrategy;
typedef enum { RBM_NORMAL } ReadBufferMode;
#define BUFFER_LOCK_SHARE 1
#define BUFFER_LOCK_EXCLUSIVE 2
extern Buffer ReadBuffer(Relation, BlockNumber); extern Buffer ReadBufferExtended(Relation, ForkNumber, BlockNumber, ReadBufferMode, BufferAccessStrategy);
extern void LockBuffer(Buffer, int); extern void UnlockReleaseBuffer(Buffer); extern Page BufferGetPage(Buffer);
extern Datum index_getattr(IndexTuple, int, TupleDesc, bool *);
extern FmgrInfo *index_getprocinfo(Relation, AttrNumber, uint16); extern Oid index_getprocid(Relation, AttrNumber, uint16);
#define OidIsValid(o) ((o) != InvalidOid)

/* scans */
typedef struct ScanKeyData { int sk_flags; Datum sk_argument; } ScanKeyData; typedef ScanKeyData *ScanKey;
#define SK_ISNULL 0x0001
typedef struct SnapshotData *Snapshot;
typedef struct IndexScanDescData { Relation indexRelation; Snapshot xs_snapshot; int numberOfKeys, numberOfOrderBys; ScanKey keyData, orderByData;
	void *opaque; ItemPointerData xs_heaptid; bool xs_recheck, xs_recheckorderby; void *instrument; } IndexScanDescData;
typedef IndexScanDescData *IndexScanDesc;
typedef struct IndexInfo IndexInfo; typedef struct IndexBuildResult IndexBuildResult; typedef struct IndexVacuumInfo IndexVacuumInfo;
typedef struct IndexBulkDeleteResult IndexBulkDeleteResult; typedef bool (*IndexBulkDeleteCallback) (ItemPointer, void *);
typedef struct PlannerInfo PlannerInfo; typedef struct IndexPath IndexPath; typedef double Cost; typedef double Selectivity;
typedef struct bytea bytea; typedef int IndexUniqueCheck;

/* executor / tuplesort */
typedef struct TupleTableSlot { Datum *tts_values; bool *tts_isnull; } TupleTableSlot;
extern TupleTableSlot *ExecClearTuple(TupleTableSlot *); extern TupleTableSlot *ExecStoreVirtualTuple(TupleTableSlot *);
typedef struct Tuplesortstate Tuplesortstate;
extern void tuplesort_puttupleslot(Tuplesortstate *, TupleTableSlot *);

/* misc types the reference headers mention */
typedef struct pairingheap_node { struct pairingheap_node *first_child, *next_sibling, *prev_or_parent; } pairingheap_node;
typedef struct pairingheap pairingheap;
typedef struct GenericXLogState GenericXLogState; typedef struct ParallelContext ParallelContext;
typedef struct dsm_segment dsm_segment; typedef struct shm_toc shm_toc; typedef struct Sharedsort Sharedsort;
typedef struct ConditionVariable { int x; } ConditionVariable; typedef unsigned char slock_t;
typedef struct LWLock { int x; } LWLock; typedef struct List List;
typedef struct BlockSamplerData { int x; } BlockSamplerData; typedef struct ReservoirStateData { int x; } ReservoirStateData;
typedef struct ParallelTableScanDescData ParallelTableScanDescData; typedef struct instr_time { int64 t; } instr_time;
typedef struct pg_prng_state { uint64 s0, s1; } pg_prng_state;
extern pg_prng_state pg_global_prng_state; extern double pg_prng_double(pg_prng_state *); extern uint32 pg_prng_uint32(pg_prng_state *);
extern void pg_prng_seed(pg_prng_state *, uint64);
extern uint64 murmurhash64(uint64);
#define relptr(type) union { type *relptr_type; Size relptr_off; }
#define relptr_declare(type, relptrtype) typedef relptr(type) relptrtype
#define relptr_access(base, rp) ((__typeof__((rp).relptr_type)) ((rp).relptr_off == 0 ? NULL : (base) + (rp).relptr_off - 1))
#define relptr_is_null(rp) ((rp).relptr_off == 0)
#define relptr_store(base, rp, val) ((rp).relptr_off = ((val) == NULL ? 0 : ((char *) (val)) - (base) + 1))
typedef struct StringInfoData { char *data; int len, maxlen, cursor; } StringInfoData; typedef StringInfoData *StringInfo;
extern void initStringInfo(StringInfo); extern void appendBinaryStringInfo(StringInfo, const void *, int);
extern int work_mem, maintenance_work_mem;
#endif
