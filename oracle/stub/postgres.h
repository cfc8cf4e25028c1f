/*
 * Stub of PostgreSQL's postgres.h -- NOT reference code, NOT PostgreSQL code.
 * Just enough declarations for /root/reference/src/halfutils.c and bitutils.c
 * to compile unmodified into oracle/_ref/ (SURVEY Appendix C).
 * TEST INFRASTRUCTURE ONLY.
 */
#ifndef STUB_POSTGRES_H
#define STUB_POSTGRES_H
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef size_t Size;
typedef uintptr_t Datum;
typedef char *Pointer;

#define PG_VERSION_NUM 170000
#define FLEXIBLE_ARRAY_MEMBER
#define HAVE__GET_CPUID 1
#define HAVE__BUILTIN_POPCOUNT 1
#define HAVE_LONG_INT_64 1
#define PGDLLEXPORT
#ifndef likely
#define likely(x) __builtin_expect((x) != 0, 1)
#define unlikely(x) __builtin_expect((x) != 0, 0)
#endif
#define palloc(sz) malloc(sz)
#define palloc0(sz) calloc(1, sz)
#define pfree(p) free(p)
#define ERROR 21
#define ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE 0
#define errcode(x) 0
#define errmsg(...) 0
#define ereport(level, rest) abort()
#define PG_FUNCTION_ARGS void *fcinfo
#define PG_DETOAST_DATUM(x) ((void *) (x))
#define PG_GETARG_DATUM(n) ((Datum) 0)
#define PG_RETURN_POINTER(x) return (Datum) (x)
static inline Size add_size(Size a, Size b) { return a + b; }
static inline Size mul_size(Size a, Size b) { return a * b; }
#endif
