/* pgstub: stand-in for lib/simplehash.h -- declares only the hash type the includer names */
#define VB_SH_CAT_(a, b) a##b
#define VB_SH_CAT(a, b) VB_SH_CAT_(a, b)
struct VB_SH_CAT(SH_PREFIX, _hash);
#undef SH_PREFIX
#undef SH_ELEMENT_TYPE
#undef SH_KEY_TYPE
#undef SH_SCOPE
#undef SH_DECLARE
#undef VB_SH_CAT
#undef VB_SH_CAT_
