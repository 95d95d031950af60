// K3 / K4 - FE element-block kernels (sm_100a): gather U, interpolate grad_u at
// every integration point, run the local Newton there, and emit the element
// residual R_e, the IFT-consistent element tangent K_e (K3) and the converged
// local state, in the reference's element-major layouts.
//
// Replaces, for one mesh element block (reference file:line):
//   cmad/fem/assembly.py:616-732   assemble_element_block (vmap over elements)
//   cmad/fem/assembly.py:416-535   per_element_R_and_K_coupled (scan over IPs)
//   cmad/fem/assembly.py:538-613   per_element_R_coupled (K4: WANT_K = false)
//   cmad/global_residuals/global_residual.py:361-395   per-IP COUPLED evaluator
//   cmad/global_residuals/small_disp_equilibrium.py:112-118  R = (grad_N @ sigma) w dv
//   cmad/global_residuals/interpolation.py:49-55       grad_u = U_e^T grad_N
//
// Mapping: one thread per integration point.  tet4 (1 IP): the thread owns the
// element, everything stays in registers and K_e rows are streamed out as they
// are produced.  hex8 (8 IPs): the 8 threads of an element solve their points,
// publish (D w dv, grad_N, sigma w dv) in shared memory, then thread a sums the
// three K_e rows of node a over the 8 points in fixed order (bit-reproducible).
// Element-major arrays are moved with 256-bit vector loads/stores
// (LDG.E.256 / STG.E.256): every request is a whole 32-byte sector.
//
// The tangent: with g_kl = d u_k / d x_l = sum_b U[b,k] gN[b,l] and
// D[v(j,i)][v(k,l)] = d sigma_ji / d eps_kl (symmetric strain components),
//   K[(a,i),(b,k)] = w dv sum_{j,l} gN[a,j] D[v(j,i)][v(k,l)] (k==l ? 1 : 1/2) gN[b,l].
#include "fe_block.cuh"

namespace cmadx {

cudaError_t launch_fe_tet4(const FeArgs& A, int solver, bool list, cudaStream_t stream, int sms);
cudaError_t launch_fe_hex8(const FeArgs& A, int solver, bool list, cudaStream_t stream, int sms);
cudaError_t launch_fe_generic(const FeArgs& A, int solver, cudaStream_t stream);
cudaError_t launch_fe_generic_list(const FeArgs& A, cudaStream_t stream);
cudaError_t launch_fe_generic_barlat(const FeArgs& A, cudaStream_t stream);
cudaError_t launch_fe_tet4x4(const FeArgs& A, int solver, cudaStream_t stream);
static bool tet4x4(const FeArgs& A) { return A.b.n_basis == 4 && A.b.n_ip == 4; }

// the tuned kernels cover the reference's default rules (cmad/fem/fe_problem.py:35-38)
static bool default_rule(const FeArgs& A) {
    return (A.b.n_basis == 4 && A.b.n_ip == 1) || (A.b.n_basis == 8 && A.b.n_ip == 8);
}

// main launch: J2 radial kernel where it applies, else the generic kernel
cudaError_t launch_fe_block(const FeArgs& A, bool j2_radial, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    if (A.m.yield == CMADX_YIELD_BARLAT) return launch_fe_generic_barlat(A, stream);
    if (tet4x4(A)) return launch_fe_tet4x4(A, j2_radial ? 0 : 1 + A.m.yield, stream);
    if (!default_rule(A)) return launch_fe_generic(A, 1 + A.m.yield, stream);
    const int solver = j2_radial ? 0 : 1 + A.m.yield;
    return (A.b.n_basis == 4) ? launch_fe_tet4(A, solver, false, stream, 0)
                              : launch_fe_hex8(A, solver, false, stream, 0);
}

// K6: JVP at a given state (solver 4 + yield kind), residual-shaped outputs only
cudaError_t launch_fe_block_jvp(const FeArgs& A, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    const int solver = 4 + A.m.yield;
    if (!default_rule(A)) return launch_fe_generic(A, solver, stream);
    return (A.b.n_basis == 4) ? launch_fe_tet4(A, solver, false, stream, 0)
                              : launch_fe_hex8(A, solver, false, stream, 0);
}

// second pass: the generic solver over the elements the first pass handed back (J2 radial
// return) or deferred (generic Newton needing more than defer_request updates)
cudaError_t launch_fe_block_list(const FeArgs& A, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int solver = 1 + A.m.yield;
    if (!default_rule(A)) return launch_fe_generic_list(A, stream);
    return (A.b.n_basis == 4) ? launch_fe_tet4(A, solver, true, stream, sms)
                              : launch_fe_hex8(A, solver, true, stream, sms);
}

}  // namespace cmadx
