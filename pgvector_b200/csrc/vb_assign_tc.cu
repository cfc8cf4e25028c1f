// vb_assign_tc.cu -- default nearest-centre assign entry.
// Round-1 state: routes to the exact fp32 CUDA-core kernel (vb_kmeans.cu).  The tcgen05
// bf16-split GEMM with fused row-argmin and exact re-check of near ties lands here.
#include "vb_common.cuh"

namespace vb {

int launch_assign(const Table& X, int metric, const Table& Cn, int k, int32_t* out_idx) {
    return launch_assign_exact(X, metric, Cn, k, nullptr, 0, out_idx, nullptr);
}

}  // namespace vb
