/* stub: see stub/postgres.h */
#define FLOAT_SHORTEST_DECIMAL_LEN 16
static inline int float_to_shortest_decimal_buf(float f, char *result) { (void) f; result[0] = 0; return 0; }
