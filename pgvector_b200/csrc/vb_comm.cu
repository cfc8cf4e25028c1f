// vb_comm.cu -- the library's own NCCL communicator: one process per GPU, collectives enqueued on the library stream.
//
// The reference's parallel build shares state through a dynamic shared memory segment between PostgreSQL workers
// (src/ivfbuild.c:830-966); its multi-GPU counterpart here is one backend / worker per GPU exchanging
//   - k-means partial centre sums, counts and the change counter  (ncclAllReduce, once per Lloyd iteration),
//   - k-means++ weight sums and the chosen centre row             (ncclAllGather + ncclAllReduce per new centre),
//   - probe lists and per-rank top-k of the list-sharded scan      (ncclAllGather, twice per query batch).
// NCCL is bound at run time (dlopen of libnccl.so.2: the system library in a backend, the copy torch already mapped in
// a Python process), so a single-GPU backend never loads it.  The 128-byte ncclUniqueId travels between the processes
// by whatever the host has (the extension: its DSM segment; bench.py / tests: torch.distributed's store).
#include "vb_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

namespace vb {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;

static int nccl_load() {
    if (g_nccl.lib) return VB_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return VB_ESTATE;
    }
#define VB_SYM(field, name)                                      \
    do {                                                         \
        *(void**)(&g_nccl.field) = dlsym(lib, name);             \
        if (!g_nccl.field) {                                     \
            set_error("libnccl has no symbol %s", name);         \
            dlclose(lib);                                        \
            return VB_ESTATE;                                    \
        }                                                        \
    } while (0)
    VB_SYM(GetUniqueId, "ncclGetUniqueId");
    VB_SYM(CommInitRank, "ncclCommInitRank");
    VB_SYM(CommDestroy, "ncclCommDestroy");
    VB_SYM(AllReduce, "ncclAllReduce");
    VB_SYM(AllGather, "ncclAllGather");
    VB_SYM(Broadcast, "ncclBroadcast");
    VB_SYM(GetErrorString, "ncclGetErrorString");
#undef VB_SYM
    g_nccl.lib = lib;
    return VB_OK;
}

#define VB_NCCL(call)                                                                           \
    do {                                                                                        \
        ncclResult_t r__ = (call);                                                              \
        if (r__ != ncclSuccess) {                                                               \
            set_error("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
            return VB_ECUDA;                                                                    \
        }                                                                                       \
    } while (0)

int comm_world() { return g_comm ? g_world : 1; }
int comm_rank() { return g_comm ? g_rank : 0; }

static bool comm_dtype(int dtype, ncclDataType_t* t) {
    switch (dtype) {
        case 0: *t = ncclFloat32; return true;
        case 1: *t = ncclInt32; return true;
        case 2: *t = ncclInt64; return true;
        case 3: *t = ncclFloat64; return true;
        case 4: *t = ncclUint32; return true;
    }
    return false;
}

int comm_allreduce(void* buf_dev, int64_t count, int dtype) {
    if (!g_comm || g_world == 1 || count <= 0) return VB_OK;
    ncclDataType_t t;
    VB_REQUIRE(comm_dtype(dtype, &t), "bad allreduce dtype %d", dtype);
    VB_NCCL(g_nccl.AllReduce(buf_dev, buf_dev, (size_t)count, t, ncclSum, g_comm, ctx().stream));
    return VB_OK;
}

int comm_allgather(const void* send_dev, void* recv_dev, int64_t bytes_per_rank) {
    if (bytes_per_rank <= 0) return VB_OK;
    if (!g_comm || g_world == 1) {
        if (send_dev != recv_dev) VB_CUDA(cudaMemcpyAsync(recv_dev, send_dev, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, ctx().stream));
        return VB_OK;
    }
    VB_NCCL(g_nccl.AllGather(send_dev, recv_dev, (size_t)bytes_per_rank, ncclInt8, g_comm, ctx().stream));
    return VB_OK;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_comm_unique_id(void* out, size_t cap) {
    VB_TRY(require_init());
    VB_REQUIRE(out && cap >= sizeof(ncclUniqueId), "the id buffer must hold %zu bytes", sizeof(ncclUniqueId));
    VB_TRY(nccl_load());
    ncclUniqueId id;
    VB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return VB_OK;
}

int vb_comm_init(const void* id_bytes, int rank, int world) {
    VB_TRY(require_init());
    VB_REQUIRE(id_bytes && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
    VB_TRY(nccl_load());
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    VB_CUDA(cudaSetDevice(ctx().device));
    VB_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    g_rank = rank;
    g_world = world;
    return VB_OK;
}

int vb_comm_free(void) {
    if (g_comm) {
        cudaStreamSynchronize(ctx().stream);
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_world = 1;
    return VB_OK;
}

int vb_comm_world(void) { return comm_world(); }
int vb_comm_rank(void) { return comm_rank(); }

int vb_comm_allreduce(void* buf_dev, int64_t count, int dtype) {
    VB_TRY(require_init());
    VB_REQUIRE(buf_dev || count == 0, "null buffer");
    return comm_allreduce(buf_dev, count, dtype);
}

int vb_comm_allgather(const void* send_dev, void* recv_dev, int64_t bytes_per_rank) {
    VB_TRY(require_init());
    VB_REQUIRE((send_dev && recv_dev) || bytes_per_rank == 0, "null buffer");
    return comm_allgather(send_dev, recv_dev, bytes_per_rank);
}

}  // extern "C"
