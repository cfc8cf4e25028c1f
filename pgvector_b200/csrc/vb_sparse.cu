// vb_sparse.cu -- sparsevec on the device: the distance functions of src/sparsevec.c:826-1057
// (l2_distance / l2_squared_distance / inner_product / negative_inner_product / cosine_distance / l1_distance),
// l2_norm / l2_normalize (src/sparsevec.c:1062-1150), a resident CSR row table and the exact (no index) top-k
// over it.
//
// A sparsevec is (dim, nnz, indices[nnz] ascending 0-based, values[nnz]) -- src/sparsevec.h:21-32; a batch of rows is
// CSR: row r = entries row_off[r] .. row_off[r+1] of idx[] / val[].
//
// Formulation.  The reference merges the two index lists with a moving cursor (one pair at a time, O(nnz_a + nnz_b)
// dependent steps).  Here ONE query faces many rows, so the query is staged once per CTA in shared memory
// (indices, values and a 1024-bucket directory over the index range) and every row entry LOOKS ITS INDEX UP in the
// query: bucket = index >> shift, then a binary search inside the bucket (1-2 probes for nnz <= 16000).  One warp per
// row, lanes stride the row's entries (coalesced 4-byte loads of idx / val: the row is read once, HBM bound at 8 bytes
// per stored entry).  The query entries NOT matched by the row contribute q^2 (L2) or |q| (L1): each warp keeps a bitmap
// of matched query positions in shared memory and sums the unmatched ones afterwards -- no subtraction of large sums,
// so no cancellation.  Inner product and cosine need no bitmap.  Sums are fp32 like the reference's (warp tree instead
// of index order: within 1e-5 relative of the fp64 truth, tested against the oracle).
#include "vb_common.cuh"
#include "vb_distance.cuh"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

namespace vb {

constexpr int SP_WARPS = 8;
constexpr int SP_BUCKETS = 1024;
constexpr int SP_MAX_NNZ = 16000;           // SPARSEVEC_MAX_NNZ (src/sparsevec.h:12)
constexpr int SP_MAX_DIM = 1000000000;      // SPARSEVEC_MAX_DIM (src/sparsevec.h:11)

struct SparseTable {
    int dim = 0;
    int64_t n = 0, nnz = 0;
    int64_t cap_rows = 0, cap_nnz = 0;
    int64_t* row_off = nullptr;   // [n + 1]
    int32_t* idx = nullptr;
    float* val = nullptr;
};

// queries of a batch, CSR like the rows
struct SparseQueries {
    const int64_t* off;
    const int32_t* idx;
    const float* val;
};

__device__ __forceinline__ int sp_find(const int32_t* s_idx, const int32_t* s_bucket, int shift, int32_t key) {
    const int b = key >> shift;
    int lo = s_bucket[b], hi = s_bucket[b + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int32_t v = s_idx[mid];
        if (v == key) return mid;
        if (v < key) lo = mid + 1;
        else hi = mid;
    }
    return -1;
}

// KEY: VB_L2_SQUARED, VB_NEG_IP, VB_COSINE or VB_L1.  out_d (float8 of SQL function `metric`) or out_f (ordering key).
// grid: x = row slices, y = queries.  out[(q * n + r)].
template <int KEY>
__global__ void __launch_bounds__(SP_WARPS * 32)
sparse_scan_kernel(SparseQueries Q, int shift, const int64_t* __restrict__ row_off, const int32_t* __restrict__ idx,
                   const float* __restrict__ val, int64_t n, int metric, double* __restrict__ out_d, float* __restrict__ out_f) {
    constexpr bool FLAGS = KEY == VB_L2_SQUARED || KEY == VB_L1;
    extern __shared__ __align__(16) uint8_t smem[];
    const int q = blockIdx.y;
    const int64_t qb = Q.off[q];
    const int qn = (int)(Q.off[q + 1] - qb);
    const int nwords = (qn + 31) >> 5;
    int32_t* s_idx = reinterpret_cast<int32_t*>(smem);
    float* s_val = reinterpret_cast<float*>(s_idx + qn);
    int32_t* s_bucket = reinterpret_cast<int32_t*>(s_val + qn);
    uint32_t* s_flags = reinterpret_cast<uint32_t*>(s_bucket + SP_BUCKETS + 1);
    __shared__ float s_qnorm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < qn; i += blockDim.x) {
        s_idx[i] = Q.idx[qb + i];
        s_val[i] = Q.val[qb + i];
    }
    __syncthreads();
    // bucket b starts at the first query entry with index >= b << shift
    for (int b = threadIdx.x; b <= SP_BUCKETS; b += blockDim.x) {
        const int64_t first = (int64_t)b << shift;
        int lo = 0, hi = qn;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int64_t)s_idx[mid] < first) lo = mid + 1;
            else hi = mid;
        }
        s_bucket[b] = lo;
    }
    if (KEY == VB_COSINE && warp == 0) {   // fp32 sum of squares of the query (normb, src/sparsevec.c:992-993)
        float s = 0.f;
        for (int i = lane; i < qn; i += 32) s += s_val[i] * s_val[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_qnorm = s;
    }
    __syncthreads();

    uint32_t* flags = s_flags + (size_t)warp * nwords;
    const int64_t wstride = (int64_t)gridDim.x * SP_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * SP_WARPS + warp; r < n; r += wstride) {
        const int64_t beg = row_off[r], end = row_off[r + 1];
        if (FLAGS) {
            for (int w = lane; w < nwords; w += 32) flags[w] = 0u;
            __syncwarp();
        }
        float acc = 0.f, rn = 0.f;
        for (int64_t p = beg + lane; p < end; p += 32) {
            const int32_t ri = __ldg(idx + p);
            const float rv = __ldg(val + p);
            const int pos = sp_find(s_idx, s_bucket, shift, ri);
            const float qv = pos >= 0 ? s_val[pos] : 0.f;
            if (KEY == VB_L2_SQUARED) {
                const float t = rv - qv;
                acc += t * t;
            } else if (KEY == VB_L1) {
                acc += fabsf(rv - qv);
            } else {
                acc += rv * qv;
                if (KEY == VB_COSINE) rn += rv * rv;
            }
            if (FLAGS && pos >= 0) atomicOr(&flags[pos >> 5], 1u << (pos & 31));
        }
        if (FLAGS) {
            __syncwarp();
            for (int w = lane; w < nwords; w += 32) {
                uint32_t m = ~flags[w];
                if (w == nwords - 1 && (qn & 31)) m &= (1u << (qn & 31)) - 1u;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const float qv = s_val[w * 32 + b];
                    acc += KEY == VB_L2_SQUARED ? qv * qv : fabsf(qv);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (KEY == VB_COSINE) rn += __shfl_xor_sync(0xffffffffu, rn, o);
        }
        if (lane == 0) {
            double v;
            if (KEY == VB_COSINE) {
                // src/sparsevec.c:985-1009: similarity / sqrt((double) norma * (double) normb), clamped, 1 - similarity
                double sim = (double)acc / sqrt((double)rn * (double)s_qnorm);
                if (sim > 1.0) sim = 1.0;
                else if (sim < -1.0) sim = -1.0;
                v = 1.0 - sim;
            } else if (KEY == VB_NEG_IP) {
                v = metric == VB_IP ? (double)acc : (double)-acc;
            } else if (KEY == VB_L2_SQUARED) {
                v = (metric == VB_L2 && out_d) ? sqrt((double)acc) : (double)acc;
            } else {
                v = (double)acc;
            }
            const size_t at = (size_t)q * (size_t)n + (size_t)r;
            if (out_d) out_d[at] = v;
            else out_f[at] = (float)v;
        }
    }
}

static size_t sparse_smem_bytes(int max_qnnz) {
    const size_t words = ((size_t)max_qnnz + 31) / 32;
    return (size_t)max_qnnz * 8 + (SP_BUCKETS + 1) * 4 + words * 4 * SP_WARPS + 16;
}

static int sparse_shift(int dim) {
    int shift = 0;
    while ((((int64_t)dim - 1) >> shift) >= SP_BUCKETS) ++shift;
    return shift;
}

// distances of nq queries against the n CSR rows; metric = the SQL function (out_d) or its ordering key (out_f)
static int launch_sparse_scan(int metric, int dim, SparseQueries Q, int64_t nq, int max_qnnz, const int64_t* row_off, const int32_t* idx,
                              const float* val, int64_t n, double* out_d, float* out_f) {
    if (n <= 0 || nq <= 0) return VB_OK;
    Context& c = ctx();
    const size_t smem = sparse_smem_bytes(max_qnnz);
    const int shift = sparse_shift(dim);
    int64_t gx = std::max<int64_t>(1, (2 * (int64_t)c.sm_count + nq - 1) / nq);
    gx = std::min<int64_t>(gx, (n + SP_WARPS - 1) / SP_WARPS);
    VB_REQUIRE(nq <= 65535, "at most 65535 sparse queries per launch");
    dim3 grid((unsigned)gx, (unsigned)nq);
    const int km = key_metric(metric);
#define SP_LAUNCH(KEY)                                                                                                     \
    do {                                                                                                                   \
        VB_CUDA(cudaFuncSetAttribute(sparse_scan_kernel<KEY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        sparse_scan_kernel<KEY><<<grid, SP_WARPS * 32, smem, c.stream>>>(Q, shift, row_off, idx, val, n, metric, out_d, out_f); \
    } while (0)
    switch (km) {
        case VB_L2_SQUARED: SP_LAUNCH(VB_L2_SQUARED); break;
        case VB_NEG_IP: SP_LAUNCH(VB_NEG_IP); break;
        case VB_COSINE: SP_LAUNCH(VB_COSINE); break;
        case VB_L1: SP_LAUNCH(VB_L1); break;
        default: set_error("metric %d is not defined for sparsevec", metric); return VB_EINVAL;
    }
#undef SP_LAUNCH
    VB_CUDA(cudaGetLastError());
    count_launch();
    return VB_OK;
}

// ----------------------------------------------------------------------------- norm / normalize

// mode 0: norms (fp64 sum of squares, src/sparsevec.c:1062-1077).  mode 1: quotients (float)(x / norm) into q_out,
// entries kept (quotient != 0) per row into kept, overflow flag (src/sparsevec.c:1100-1113)
__global__ void sparse_norm_kernel(const int64_t* __restrict__ row_off, const float* __restrict__ val, int64_t n, int mode,
                                   double* __restrict__ norms, float* __restrict__ q_out, int64_t* __restrict__ kept, int* __restrict__ overflow) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x % 32;
    if (r >= n) return;
    const int64_t beg = row_off[r], end = row_off[r + 1];
    double s = 0.0;
    for (int64_t p = beg + lane; p < end; p += 32) {
        const double x = (double)val[p];
        s += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double norm = sqrt(s);
    if (mode == 0) {
        if (lane == 0) norms[r] = norm;
        return;
    }
    int k = 0;
    bool inf = false;
    for (int64_t p = beg + lane; p < end; p += 32) {
        const float v = norm > 0 ? (float)((double)val[p] / norm) : 0.f;
        inf |= isinf(v);
        q_out[p] = v;
        k += v != 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    if (lane == 0) kept[r] = k;
    if (inf) atomicExch(overflow, 1);
}

// rows without their zero quotients, in index order (src/sparsevec.c:1115-1140)
__global__ void sparse_compact_kernel(const int64_t* __restrict__ row_off, const int32_t* __restrict__ idx, const float* __restrict__ q_in,
                                      int64_t n, const int64_t* __restrict__ out_off, int32_t* __restrict__ out_idx,
                                      float* __restrict__ out_val) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x % 32;
    if (r >= n) return;
    const int64_t beg = row_off[r], end = row_off[r + 1];
    int64_t w = out_off[r];
    for (int64_t p0 = beg; p0 < end; p0 += 32) {
        const int64_t p = p0 + lane;
        const float v = p < end ? q_in[p] : 0.f;
        const bool keep = v != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int64_t at = w + __popc(m & ((1u << lane) - 1u));
            out_idx[at] = idx[p];
            out_val[at] = v;
        }
        w += __popc(m);
    }
}

__global__ void sparse_segments_kernel(int64_t nseg, int64_t n, int64_t* begin, int32_t* lens) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nseg) {
        begin[i] = i * n;
        lens[i] = (int32_t)n;
    }
}

__global__ void sparse_finish_kernel(int metric, int64_t total, const int32_t* __restrict__ pos, const float* __restrict__ key,
                                     int64_t* __restrict__ out_ids, double* __restrict__ out_d) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    out_ids[i] = pos[i];
    // the ordering key is the float8 of the function except for <-> (sqrt of the fp32 L2 squared, src/sparsevec.c:872-883)
    out_d[i] = metric == VB_L2 ? sqrt((double)key[i]) : (double)key[i];
}

// workspace slots of this file
enum { WSP_Q = 20, WSP_ROWS = 21, WSP_OUT = 22, WSP_TMP = 23, WSP_SEG = 24, WSP_POS = 25, WSP_SCAN = 26 };

static bool sparse_metric_ok(int metric) {
    return metric == VB_L2_SQUARED || metric == VB_L2 || metric == VB_IP || metric == VB_NEG_IP || metric == VB_COSINE || metric == VB_L1;
}

// Validate host CSR: offsets non-decreasing from 0, row nnz <= SPARSEVEC_MAX_NNZ, indices ascending inside [0, dim)
// (what sparsevec_in / sparsevec_recv guarantee for stored values, src/sparsevec.c:88-104, 511-521)
static int check_csr(const char* what, int dim, int64_t n, const int64_t* off, const int32_t* idx, int* max_nnz) {
    VB_REQUIRE(dim >= 1 && dim <= SP_MAX_DIM, "sparsevec must have between 1 and %d dimensions", SP_MAX_DIM);
    VB_REQUIRE(off && off[0] == 0, "%s: offsets must start at 0", what);
    int mx = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t len = off[r + 1] - off[r];
        VB_REQUIRE(len >= 0, "%s: offsets must not decrease (row %lld)", what, (long long)r);
        VB_REQUIRE(len <= SP_MAX_NNZ, "sparsevec cannot have more than %d non-zero elements", SP_MAX_NNZ);
        VB_REQUIRE(len == 0 || idx, "%s: null indices", what);
        for (int64_t p = off[r]; p < off[r + 1]; ++p) {
            VB_REQUIRE(idx[p] >= 0 && idx[p] < dim, "sparsevec index out of bounds");
            VB_REQUIRE(p == off[r] || idx[p] > idx[p - 1], "sparsevec indices must be in ascending order");
        }
        mx = std::max(mx, (int)len);
    }
    if (max_nnz) *max_nnz = mx;
    return VB_OK;
}

// queries to the device: [off (nq + 1) | idx | val] in one workspace block
static int upload_sparse_queries(int64_t nq, const int64_t* off, const int32_t* idx, const float* val, SparseQueries* Q) {
    const int64_t tot = off[nq];
    const size_t b_off = sizeof(int64_t) * (size_t)(nq + 1);
    const size_t b_idx = (sizeof(int32_t) * (size_t)tot + 15) & ~(size_t)15;
    void* d;
    VB_TRY(workspace(WSP_Q, b_off + b_idx + sizeof(float) * (size_t)tot + 64, &d));
    cudaStream_t s = ctx().stream;
    uint8_t* p = (uint8_t*)d;
    VB_CUDA(cudaMemcpyAsync(p, off, b_off, cudaMemcpyHostToDevice, s));
    if (tot > 0) {
        VB_CUDA(cudaMemcpyAsync(p + b_off, idx, sizeof(int32_t) * (size_t)tot, cudaMemcpyHostToDevice, s));
        VB_CUDA(cudaMemcpyAsync(p + b_off + b_idx, val, sizeof(float) * (size_t)tot, cudaMemcpyHostToDevice, s));
    }
    Q->off = (const int64_t*)p;
    Q->idx = (const int32_t*)(p + b_off);
    Q->val = (const float*)(p + b_off + b_idx);
    return VB_OK;
}

static int sparse_reserve(SparseTable& t, int64_t rows, int64_t nnz) {
    cudaStream_t s = ctx().stream;
    if (rows + 1 > t.cap_rows) {
        const int64_t cap = std::max<int64_t>(rows + 1, t.cap_rows * 2);
        int64_t* d;
        if (cudaMalloc(&d, sizeof(int64_t) * (size_t)cap) != cudaSuccess) {
            set_error("out of device memory (sparse row offsets)");
            return VB_ENOMEM;
        }
        if (t.row_off) {
            VB_CUDA(cudaMemcpyAsync(d, t.row_off, sizeof(int64_t) * (size_t)(t.n + 1), cudaMemcpyDeviceToDevice, s));
            VB_CUDA(cudaStreamSynchronize(s));
            cudaFree(t.row_off);
        } else {
            VB_CUDA(cudaMemsetAsync(d, 0, sizeof(int64_t), s));
        }
        t.row_off = d;
        t.cap_rows = cap;
    }
    if (nnz > t.cap_nnz) {
        const int64_t cap = std::max<int64_t>(nnz, t.cap_nnz * 2);
        int32_t* di;
        float* dv;
        if (cudaMalloc(&di, sizeof(int32_t) * (size_t)cap) != cudaSuccess) {
            set_error("out of device memory (sparse indices)");
            return VB_ENOMEM;
        }
        if (cudaMalloc(&dv, sizeof(float) * (size_t)cap) != cudaSuccess) {
            cudaFree(di);
            set_error("out of device memory (sparse values)");
            return VB_ENOMEM;
        }
        if (t.nnz > 0) {
            VB_CUDA(cudaMemcpyAsync(di, t.idx, sizeof(int32_t) * (size_t)t.nnz, cudaMemcpyDeviceToDevice, s));
            VB_CUDA(cudaMemcpyAsync(dv, t.val, sizeof(float) * (size_t)t.nnz, cudaMemcpyDeviceToDevice, s));
            VB_CUDA(cudaStreamSynchronize(s));
        }
        if (t.idx) cudaFree(t.idx);
        if (t.val) cudaFree(t.val);
        t.idx = di;
        t.val = dv;
        t.cap_nnz = cap;
    }
    return VB_OK;
}

__global__ void sparse_shift_offsets_kernel(int64_t* off, int64_t n, int64_t add) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) off[i] += add;
}

}  // namespace vb

using namespace vb;

struct vb_sparse_table {
    SparseTable t;
};

extern "C" {

int vb_sparsevec_distance_batch(int metric, int dim, int q_dim, int32_t q_nnz, const int32_t* q_idx, const float* q_val, int64_t n,
                                const int64_t* row_off, const int32_t* idx, const float* val, double* out) {
    VB_TRY(require_init());
    VB_REQUIRE(sparse_metric_ok(metric), "metric %d is not defined for sparsevec", metric);
    VB_REQUIRE(n >= 0 && (n == 0 || (row_off && out)), "bad sparsevec batch arguments");
    if (n == 0) return VB_OK;
    if (q_nnz < 0) {   // NULL query: ZeroDistance (src/hnswutils.c:555-556)
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
        return VB_OK;
    }
    // CheckDims (src/sparsevec.c:44-51)
    VB_REQUIRE(dim == q_dim, "different sparsevec dimensions %d and %d", dim, q_dim);
    VB_TRY(check_csr("rows", dim, n, row_off, idx, nullptr));
    const int64_t qoff[2] = {0, q_nnz};
    int max_q = 0;
    VB_TRY(check_csr("query", dim, 1, qoff, q_idx, &max_q));
    Context& c = ctx();
    SparseQueries Q;
    VB_TRY(upload_sparse_queries(1, qoff, q_idx, q_val, &Q));
    const int64_t tot = row_off[n];
    const size_t b_off = sizeof(int64_t) * (size_t)(n + 1);
    const size_t b_idx = (sizeof(int32_t) * (size_t)tot + 15) & ~(size_t)15;
    void *d_rows, *d_out;
    VB_TRY(workspace(WSP_ROWS, b_off + b_idx + sizeof(float) * (size_t)tot + 64, &d_rows));
    VB_TRY(workspace(WSP_OUT, sizeof(double) * (size_t)n, &d_out));
    uint8_t* p = (uint8_t*)d_rows;
    VB_CUDA(cudaMemcpyAsync(p, row_off, b_off, cudaMemcpyHostToDevice, c.stream));
    if (tot > 0) {
        VB_CUDA(cudaMemcpyAsync(p + b_off, idx, sizeof(int32_t) * (size_t)tot, cudaMemcpyHostToDevice, c.stream));
        VB_CUDA(cudaMemcpyAsync(p + b_off + b_idx, val, sizeof(float) * (size_t)tot, cudaMemcpyHostToDevice, c.stream));
    }
    VB_TRY(launch_sparse_scan(metric, dim, Q, 1, max_q, (const int64_t*)p, (const int32_t*)(p + b_off), (const float*)(p + b_off + b_idx), n,
                              (double*)d_out, nullptr));
    VB_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaStreamSynchronize(c.stream));
    return VB_OK;
}

int vb_sparsevec_norm_batch(int64_t n, const int64_t* row_off, const float* val, double* out) {
    VB_TRY(require_init());
    VB_REQUIRE(n >= 0 && (n == 0 || (row_off && out && row_off[0] == 0)), "bad sparsevec norm arguments");
    if (n == 0) return VB_OK;
    for (int64_t r = 0; r < n; ++r) VB_REQUIRE(row_off[r + 1] >= row_off[r], "offsets must not decrease (row %lld)", (long long)r);
    Context& c = ctx();
    const int64_t tot = row_off[n];
    const size_t b_off = sizeof(int64_t) * (size_t)(n + 1);
    void *d_rows, *d_out;
    VB_TRY(workspace(WSP_ROWS, b_off + sizeof(float) * (size_t)tot + 64, &d_rows));
    VB_TRY(workspace(WSP_OUT, sizeof(double) * (size_t)n, &d_out));
    uint8_t* p = (uint8_t*)d_rows;
    VB_CUDA(cudaMemcpyAsync(p, row_off, b_off, cudaMemcpyHostToDevice, c.stream));
    if (tot > 0) VB_CUDA(cudaMemcpyAsync(p + b_off, val, sizeof(float) * (size_t)tot, cudaMemcpyHostToDevice, c.stream));
    sparse_norm_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, c.stream>>>((const int64_t*)p, (const float*)(p + b_off), n, 0, (double*)d_out,
                                                                                nullptr, nullptr, nullptr);
    VB_CUDA(cudaGetLastError());
    count_launch();
    VB_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
    VB_CUDA(cudaStreamSynchronize(c.stream));
    return VB_OK;
}

int vb_sparsevec_l2_normalize_batch(int64_t n, const int64_t* row_off, const int32_t* idx, const float* val, int64_t* out_row_off,
                                    int32_t* out_idx, float* out_val) {
    VB_TRY(require_init());
    VB_REQUIRE(n >= 0 && (n == 0 || (row_off && out_row_off && row_off[0] == 0)), "bad sparsevec normalize arguments");
    if (n == 0) {
        if (out_row_off) out_row_off[0] = 0;
        return VB_OK;
    }
    for (int64_t r = 0; r < n; ++r) VB_REQUIRE(row_off[r + 1] >= row_off[r], "offsets must not decrease (row %lld)", (long long)r);
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const int64_t tot = row_off[n];
    VB_REQUIRE(tot == 0 || (idx && val && out_idx && out_val), "null sparsevec buffers");
    const size_t b_off = sizeof(int64_t) * (size_t)(n + 1);
    const size_t b_idx = (sizeof(int32_t) * (size_t)tot + 15) & ~(size_t)15;
    const size_t b_val = (sizeof(float) * (size_t)tot + 15) & ~(size_t)15;
    void *d_rows, *d_tmp, *d_outb, *d_scan;
    VB_TRY(workspace(WSP_ROWS, b_off + b_idx + b_val + 64, &d_rows));
    // tmp: quotients | kept[n + 1] | out_off[n + 1] | overflow flag
    VB_TRY(workspace(WSP_TMP, b_val + 2 * b_off + 64, &d_tmp));
    VB_TRY(workspace(WSP_OUT, b_idx + b_val + 64, &d_outb));
    uint8_t* p = (uint8_t*)d_rows;
    uint8_t* t = (uint8_t*)d_tmp;
    float* d_q = (float*)t;
    int64_t* d_kept = (int64_t*)(t + b_val);
    int64_t* d_ooff = (int64_t*)(t + b_val + b_off);
    int* d_flag = (int*)(t + b_val + 2 * b_off);
    VB_CUDA(cudaMemcpyAsync(p, row_off, b_off, cudaMemcpyHostToDevice, s));
    if (tot > 0) {
        VB_CUDA(cudaMemcpyAsync(p + b_off, idx, sizeof(int32_t) * (size_t)tot, cudaMemcpyHostToDevice, s));
        VB_CUDA(cudaMemcpyAsync(p + b_off + b_idx, val, sizeof(float) * (size_t)tot, cudaMemcpyHostToDevice, s));
    }
    VB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
    VB_CUDA(cudaMemsetAsync(d_kept + n, 0, sizeof(int64_t), s));
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    sparse_norm_kernel<<<grid, 256, 0, s>>>((const int64_t*)p, (const float*)(p + b_off + b_idx), n, 1, nullptr, d_q, d_kept, d_flag);
    VB_CUDA(cudaGetLastError());
    count_launch();
    size_t scan_bytes = 0;
    VB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_kept, d_ooff, (int)(n + 1), s));
    VB_TRY(workspace(WSP_SCAN, scan_bytes + 64, &d_scan));
    VB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_kept, d_ooff, (int)(n + 1), s));
    count_launch();
    uint8_t* o = (uint8_t*)d_outb;
    sparse_compact_kernel<<<grid, 256, 0, s>>>((const int64_t*)p, (const int32_t*)(p + b_off), d_q, n, d_ooff, (int32_t*)o, (float*)(o + b_idx));
    VB_CUDA(cudaGetLastError());
    count_launch();
    int flag = 0;
    VB_CUDA(cudaMemcpyAsync(out_row_off, d_ooff, b_off, cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaStreamSynchronize(s));
    // float_overflow_error() (src/sparsevec.c:1107-1108)
    VB_REQUIRE(!flag, "value out of range: overflow");
    const int64_t kept = out_row_off[n];
    if (kept > 0) {
        VB_CUDA(cudaMemcpyAsync(out_idx, o, sizeof(int32_t) * (size_t)kept, cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaMemcpyAsync(out_val, o + b_idx, sizeof(float) * (size_t)kept, cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaStreamSynchronize(s));
    }
    return VB_OK;
}

// ----------------------------------------------------------------------------- resident CSR table + exact scan

int vb_sparse_table_create(int dim, vb_sparse_table** out) {
    VB_TRY(require_init());
    VB_REQUIRE(out && dim >= 1 && dim <= SP_MAX_DIM, "sparsevec must have between 1 and %d dimensions", SP_MAX_DIM);
    vb_sparse_table* t = new vb_sparse_table();
    t->t.dim = dim;
    *out = t;
    return VB_OK;
}

int vb_sparse_table_append(vb_sparse_table* h, int64_t n, const int64_t* row_off, const int32_t* idx, const float* val) {
    VB_TRY(require_init());
    VB_REQUIRE(h && n >= 0 && (n == 0 || row_off), "bad sparse table arguments");
    if (n == 0) return VB_OK;
    SparseTable& t = h->t;
    VB_TRY(check_csr("rows", t.dim, n, row_off, idx, nullptr));
    const int64_t tot = row_off[n];
    VB_REQUIRE(tot == 0 || val, "null sparsevec values");
    VB_TRY(sparse_reserve(t, t.n + n, t.nnz + tot));
    cudaStream_t s = ctx().stream;
    // offsets of the new rows: row_off[1..n] + nnz so far
    VB_CUDA(cudaMemcpyAsync(t.row_off + t.n + 1, row_off + 1, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    if (t.nnz > 0) {
        sparse_shift_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t.row_off + t.n + 1, n, t.nnz);
        VB_CUDA(cudaGetLastError());
        count_launch();
    }
    if (tot > 0) {
        VB_CUDA(cudaMemcpyAsync(t.idx + t.nnz, idx, sizeof(int32_t) * (size_t)tot, cudaMemcpyHostToDevice, s));
        VB_CUDA(cudaMemcpyAsync(t.val + t.nnz, val, sizeof(float) * (size_t)tot, cudaMemcpyHostToDevice, s));
    }
    VB_CUDA(cudaStreamSynchronize(s));
    t.n += n;
    t.nnz += tot;
    return VB_OK;
}

int64_t vb_sparse_table_rows(const vb_sparse_table* h) { return h ? h->t.n : 0; }
int64_t vb_sparse_table_nnz(const vb_sparse_table* h) { return h ? h->t.nnz : 0; }

int vb_sparse_table_free(vb_sparse_table* h) {
    if (h) {
        if (h->t.row_off) cudaFree(h->t.row_off);
        if (h->t.idx) cudaFree(h->t.idx);
        if (h->t.val) cudaFree(h->t.val);
        delete h;
    }
    return VB_OK;
}

int vb_sparse_exact_topk(vb_sparse_table* h, int metric, int q_dim, int64_t nq, const int64_t* q_off, const int32_t* q_idx, const float* q_val,
                         int k, int64_t* out_ids, double* out_dist) {
    VB_TRY(require_init());
    VB_REQUIRE(h && sparse_metric_ok(metric) && metric != VB_IP, "bad sparse table / ordering metric");
    VB_REQUIRE(k > 0 && k <= 2048, "k must be in 1..2048");
    if (nq <= 0) return VB_OK;
    SparseTable& t = h->t;
    VB_REQUIRE(t.dim == q_dim, "different sparsevec dimensions %d and %d", t.dim, q_dim);
    VB_REQUIRE(q_off && out_ids && out_dist, "null query / output buffers");
    int max_q = 0;
    VB_TRY(check_csr("queries", t.dim, nq, q_off, q_idx, &max_q));
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const int64_t n = t.n;
    if (n == 0) {
        for (int64_t i = 0; i < nq * k; ++i) {
            out_ids[i] = -1;
            out_dist[i] = INFINITY;
        }
        return VB_OK;
    }
    // sub-batches keep the key matrix under ~1 GiB
    const int64_t bq = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nq, 65535), (int64_t)(1ull << 30) / (4 * n)));
    for (int64_t q0 = 0; q0 < nq; q0 += bq) {
        const int64_t m = std::min(bq, nq - q0);
        // this sub-batch's queries, offsets rebased to 0
        std::vector<int64_t> off((size_t)m + 1);
        for (int64_t i = 0; i <= m; ++i) off[(size_t)i] = q_off[q0 + i] - q_off[q0];
        SparseQueries Q;
        VB_TRY(upload_sparse_queries(m, off.data(), q_idx + q_off[q0], q_val + q_off[q0], &Q));
        VB_CUDA(cudaStreamSynchronize(s));   // `off` is a local vector
        void *d_key, *d_seg, *d_pos, *d_out;
        VB_TRY(workspace(WSP_TMP, sizeof(float) * (size_t)m * (size_t)n, &d_key));
        VB_TRY(launch_sparse_scan(key_metric(metric), t.dim, Q, m, max_q, t.row_off, t.idx, t.val, n, nullptr, (float*)d_key));
        VB_TRY(workspace(WSP_SEG, (sizeof(int64_t) + sizeof(int32_t)) * (size_t)m + 64, &d_seg));
        int64_t* seg_begin = (int64_t*)d_seg;
        int32_t* seg_len = (int32_t*)(seg_begin + m);
        sparse_segments_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(m, n, seg_begin, seg_len);
        VB_CUDA(cudaGetLastError());
        count_launch();
        VB_TRY(workspace(WSP_POS, (sizeof(int32_t) + sizeof(float)) * (size_t)m * k, &d_pos));
        int32_t* pos = (int32_t*)d_pos;
        float* key = (float*)(pos + (size_t)m * k);
        VB_TRY(launch_segment_topk_v((const float*)d_key, seg_begin, seg_len, nullptr, nullptr, m, k, pos, key));
        VB_TRY(workspace(WSP_OUT, (sizeof(int64_t) + sizeof(double)) * (size_t)m * k, &d_out));
        int64_t* o_ids = (int64_t*)d_out;
        double* o_d = (double*)(o_ids + (size_t)m * k);
        sparse_finish_kernel<<<(unsigned)((m * k + 255) / 256), 256, 0, s>>>(metric, m * k, pos, key, o_ids, o_d);
        VB_CUDA(cudaGetLastError());
        count_launch();
        VB_CUDA(cudaMemcpyAsync(out_ids + q0 * k, o_ids, sizeof(int64_t) * (size_t)m * k, cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaMemcpyAsync(out_dist + q0 * k, o_d, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaStreamSynchronize(s));
    }
    return VB_OK;
}

}  // extern "C"
