// Launch-argument struct of the K2 kernels (mp_sens.cu, mp_sens_dt.cu, mp_hess.cu).
#pragma once
#include "point_solver.cuh"

namespace cmadx {

struct SensArgs {
    DevMat m;
    int n_active;
    int pid[CMADX_MAX_ACTIVE];
    cmadx_mp_history_t h;
    double* partials;     // [nblk][1 + n_active] (gradient) / [nblk][n_active (n_active + 1) / 2] (Hessian)
    double* phi_hist;     // [N+1][7][ld] or NULL: the adjoint pass stores phi_t (Hessian path)
    int hess_flags;       // CMADX_HESS_F_* (Hessian path)
};

}  // namespace cmadx
