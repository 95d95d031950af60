// Launch-argument struct of the FE element-block kernels (fe_block.cu).
#pragma once
#include "point_solver.cuh"

namespace cmadx {

struct FeArgs {
    DevMat m;
    DevNewton nw;
    cmadx_fe_block_t b;
    // elements the J2 radial kernel hands back to the generic kernel (list mode)
    unsigned* bail_count;
    int* bail_list;
    unsigned bail_cap;
    // K6 (JVP at a given state): converged local state, input tangents
    const double* xi_state;   // [n_elems][n_ip][7]
    const double* dxi_prev;   // [n_elems][n_ip][7] or NULL (= 0)
    const double* dU;         // [n_dofs] or NULL (= 0): displacement direction
    double dp[CMADX_MAX_ACTIVE];
    int pid[CMADX_MAX_ACTIVE];
    int n_active;
};

cudaError_t launch_fe_block(const FeArgs& A, bool j2_radial, cudaStream_t stream);
cudaError_t launch_fe_block_list(const FeArgs& A, cudaStream_t stream);
cudaError_t launch_fe_block_jvp(const FeArgs& A, cudaStream_t stream);

}  // namespace cmadx
