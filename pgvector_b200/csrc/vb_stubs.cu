// vb_stubs.cu -- entry points declared in include/vecb200.h whose kernels are not built yet.
// They fail loudly (VB_ESTATE); there is no CPU fallback behind any of them.
#include "vb_common.cuh"

#define NOT_YET(name)                                          \
    do {                                                       \
        vb::set_error(name ": not implemented in this build"); \
        return VB_ESTATE;                                      \
    } while (0)

extern "C" {
int vb_kmeans(vb_table*, int, void*, int, int, uint64_t, vb_allreduce_fn, void*, int*) { NOT_YET("vb_kmeans"); }
int vb_kmeans_pp_init(vb_table*, int, void*, int, uint64_t) { NOT_YET("vb_kmeans_pp_init"); }
int vb_assign(vb_table*, int, const void*, int, int32_t*) { NOT_YET("vb_assign"); }
int vb_assign_dev(vb_table*, int, const void*, int, int32_t*) { NOT_YET("vb_assign_dev"); }
int vb_hnsw_create(int, int, int, int, vb_hnsw**) { NOT_YET("vb_hnsw_create"); }
int vb_hnsw_load(vb_hnsw*, const void*, int64_t, const int32_t*, const int32_t*, const int64_t*, const int32_t*, int64_t, int64_t) { NOT_YET("vb_hnsw_load"); }
int vb_hnsw_free(vb_hnsw*) { return VB_OK; }
int vb_hnsw_search(vb_hnsw*, const void*, int64_t, int, int, int64_t*, double*, int64_t*) { NOT_YET("vb_hnsw_search"); }
int vb_hnsw_search_dev(vb_hnsw*, const void*, int64_t, int, int, int64_t*, float*, int64_t*) { NOT_YET("vb_hnsw_search_dev"); }
}
