// K1-J2 - radial-return specialisation of the batched constitutive update for
// J2 (von Mises) plasticity with Voce and/or linear isotropic hardening.
//
// Why it is the *same* algorithm as the reference's generic local Newton
// (cmad/models/nonlinear_solver.py:88-174 / :14-85 on the residual of
// cmad/models/small_elastic_plastic.py:238-302): started from x0 = xi_prev the
// flow-rule rows of the plastic residual vanish identically at every iterate,
// because the J2 normal depends only on the direction of the deviatoric trial
// stress, which the update  ep += dalpha * n  preserves.  The 7x7 Newton step
//     [I + b(Pdev - s^ (W s^)^T)   -n ] [dep ]   [0]
//     [      -(W n)^T           -H'/2mu] [dalp] = [f]
// then reduces exactly to  dalpha = -f / (n:n + H'/2mu) = -f / (3/2 + H'/2mu),
// dep = n dalpha, the merit of the line search to f^2/2, and the convergence
// norms to |f|.  The kernel therefore iterates on the scalar alpha with the
// reference's loop structure (same tests in the same order, same Armijo /
// quadratic-backtracking line search), and produces the same iterates up to
// rounding, the same iteration counts and the same branch flags.  The IFT
// outputs use the closed-form inverse of that Jacobian.
//
// Whenever an evaluated iterate leaves the regime in which the reduction holds
// (an iterate or line-search probe on the elastic branch, a return past the
// origin of the deviatoric plane, non-finite values, an elastic entry state that
// is not already converged) the point is appended to a "bail" list and
// re-solved from scratch by the generic kernel (mp_update.cu, list mode).
//
// One thread per point, ~13 coalesced loads and ~85 streaming stores per thread,
// a handful of FP64 operations in between: HBM-bound by construction.
#include "mp_update_j2_point.cuh"

namespace cmadx {
namespace {

__global__ void __launch_bounds__(MP_BLOCK)
mp_update_j2_kernel(const __grid_constant__ MpArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j2_point_update(A, i, i < A.b.n)) {
        // hand the point to the generic kernel; it rewrites every output
        const unsigned slot = atomicAdd(A.bail_count, 1u);
        if (slot < A.bail_cap) A.bail_list[slot] = (int)i;
    }
}

}  // namespace

cudaError_t launch_mp_update_j2(const MpArgs& A, cudaStream_t stream) {
    const int64_t nblk = (A.b.n + MP_BLOCK - 1) / MP_BLOCK;
    if (nblk == 0) return cudaSuccess;
    mp_update_j2_kernel<<<(unsigned)nblk, MP_BLOCK, 0, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace cmadx
