/* stub: see stub/postgres.h */
#ifndef STUB_PG_BITUTILS_H
#define STUB_PG_BITUTILS_H
extern const uint8 pg_number_of_ones[256];
static inline int pg_popcount64(uint64 word) { return __builtin_popcountll(word); }
#endif
