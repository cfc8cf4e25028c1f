/* pgstub: syntax-check stand-in for the PostgreSQL header of the same name (NOT PostgreSQL code) */
#include "postgres.h"

#ifndef PGSTUB_FLOAT_H
#define PGSTUB_FLOAT_H
static inline double get_float8_infinity(void) { return __builtin_inf(); }
#endif
