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
    // mixed u-p formulation (small_disp_equilibrium.py:87-111): pressure dofs and shape values
    const int32_t* mix_eq_p;  // [n_elems][n_basis] equation of the pressure dof, or NULL (displacement form)
    const double* mix_N;      // [n_ip][n_basis]
    const double* mix_h;      // [n_elems] element size (K6-mixed VJP: d tau / d mu), or NULL
    double mix_stab;          // stabilization multiplier
    double dp[CMADX_MAX_ACTIVE];
    int pid[CMADX_MAX_ACTIVE];
    int n_active;
};

cudaError_t launch_fe_block(const FeArgs& A, bool j2_radial, cudaStream_t stream);
cudaError_t launch_fe_block_list(const FeArgs& A, cudaStream_t stream);
cudaError_t launch_fe_block_jvp(const FeArgs& A, cudaStream_t stream);
cudaError_t launch_fe_mixed_pressure(const cmadx_fe_block_t& b, const cmadx_fe_mixed_t& mx, double kappa,
                                     double mu, cudaStream_t stream);
cudaError_t launch_fe_mixed_pressure_jvp(const cmadx_fe_block_t& b, const cmadx_fe_mixed_t& mx, const double* dU,
                                         double kappa, double mu, double dkappa, double dmu, cudaStream_t stream);

}  // namespace cmadx
