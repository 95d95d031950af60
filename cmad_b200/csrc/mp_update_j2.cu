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
#include "j2_radial.cuh"
#include "mp_outputs.cuh"

namespace cmadx {
namespace {

__global__ void __launch_bounds__(MP_BLOCK)
mp_update_j2_kernel(const __grid_constant__ MpArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.b.n;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    const DevNewton& nw = A.nw;

    double xp[7], em[6];
    load_point(A.b, i, live, xp, em);

    J2Radial rs;
    j2_radial_solve(m, nw, xp, em, live, rs);   // see j2_radial.cuh
    if (!live) return;
    if (rs.bail) {
        // hand the point to the generic kernel; it rewrites every output
        const unsigned slot = atomicAdd(A.bail_count, 1u);
        if (slot < A.bail_cap) A.bail_list[slot] = (int)i;
        return;
    }
    const double alpha = rs.alpha, alpha0 = rs.alpha0, f = rs.f, eD = rs.eD;
    const int ii = rs.ii, flag_entry = rs.flag_entry;
    double nc = rs.nc;
    const double (&n0v)[6] = rs.n0;

    // ---------------------------------------------------------------- outputs
    const bool pl = rs.plastic;          // branch at x*: unchanged along a valid radial solve
    const double dg = alpha - alpha0;
    double x[7];
#pragma unroll
    for (int a = 0; a < 6; ++a) x[a] = fma(dg, n0v[a], xp[a]);
    x[6] = alpha;
    double Cf[7];
#pragma unroll
    for (int a = 0; a < 6; ++a) Cf[a] = pl ? fma(-dg, n0v[a], x[a] - xp[a]) : x[a] - xp[a];
    Cf[6] = pl ? f : dg;
    if (nw.mode == CMADX_NEWTON_TRACED) nc = normN<7>(Cf);
    if (A.b.iters) A.b.iters[i] = ii;
    if (A.b.flags) A.b.flags[i] = flag_entry | ((pl ? 1 : 0) << 1);
    if (A.b.cnorm) A.b.cnorm[i] = nc;
    if (A.b.C) {
#pragma unroll
        for (int c = 0; c < 7; ++c) st(A.b.C, c, ld, i, Cf[c]);
    }
    if (A.b.xi) {
#pragma unroll
        for (int c = 0; c < 7; ++c) st(A.b.xi, c, ld, i, x[c]);
    }
    double ee[6], sig[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = em[a] - x[a];
    {
        const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
    }
    if (A.b.sigma) {
#pragma unroll
        for (int a = 0; a < 6; ++a) st(A.b.sigma, a, ld, i, sig[a]);
    }
    if (A.b.dC_dxi_prev) write_dC_dxi_prev(A.b.dC_dxi_prev, ld, i, pl, n0v);

    // yield-surface state at x*: same direction, shrunken radius
    const double snf = fma(-m.two_mu * R32, dg, rs.sn0);
    YieldFn<CMADX_YIELD_J2> yf;
    yf.sn = snf;
    yf.c = R32 / snf;
#pragma unroll
    for (int a = 0; a < 6; ++a) yf.sh[a] = rs.sh[a];
    const double beta = dg * m.two_mu * yf.c;
    const double h = j2_hardening_slope(m, eD);

    if (A.b.dC_dp && A.n_active > 0) {
        // (dn/dsigma : ee)_a = c (dev(ee)_a - s^_a (s^:ee)),  n:ee
        double see = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) see = fma(mult(a) * yf.sh[a], ee[a], see);
        const double tr3 = (ee[0] + ee[3] + ee[5]) / 3.0;
        double Mee[6];
#pragma unroll
        for (int a = 0; a < 6; ++a)
            Mee[a] = yf.c * ((is_diag(a) ? ee[a] - tr3 : ee[a]) - yf.sh[a] * see);
        write_dC_dp(A, i, pl, yf, n0v, f, eD, alpha, dg, Mee, R32 * see, sig);
    }
    if (A.b.dC_dxi) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                double v = (r == c) ? 1.0 : 0.0;
                if (pl) {
                    if (r < 6 && c < 6) v = fma(dg * m.two_mu, yf.M(r, c), v);
                    else if (r < 6) v = -n0v[r];
                    else if (c < 6) v = -mult(c) * n0v[c];
                    else v = -h;
                }
                st(A.b.dC_dxi, r * 7 + c, ld, i, v);
            }
    }
    if (!A.b.dsig_deps && !A.b.dxi_deps) return;

    // IFT with the closed-form inverse.  For a strain perturbation E:
    //   X = [A^-1]_11 E = E - g1 (dev E - s^(s^:E)) - g2 s^(s^:E),
    //   g1 = beta/(1+beta), g2 = (3/2)/(3/2 + h);  dalpha = sqrt(3/2)(s^:E)/(3/2+h)
    //   d sigma = lam tr(E) I + 2mu X ;  dx/de = [E - X ; dalpha]
    const double g1 = pl ? beta / (1.0 + beta) : 0.0;
    const double g2 = pl ? 1.5 / (1.5 + h) : 0.0;
    const double ga = pl ? R32 / (1.5 + h) : 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const double sb = mult(b) * yf.sh[b];      // s^ : E_b
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double devE = (a == b) ? 1.0 : 0.0;
            if (is_diag(a) && is_diag(b)) devE -= 1.0 / 3.0;
            // E_a - X_a
            const double emx = g1 * devE + (g2 - g1) * yf.sh[a] * sb;
            if (A.b.dxi_deps) st(A.b.dxi_deps, a * 6 + b, ld, i, emx);
            if (A.b.dsig_deps) {
                double v = m.two_mu * (((a == b) ? 1.0 : 0.0) - emx);
                if (is_diag(a) && is_diag(b)) v += m.lam;
                st(A.b.dsig_deps, a * 6 + b, ld, i, v);
            }
        }
        if (A.b.dxi_deps) st(A.b.dxi_deps, 36 + b, ld, i, ga * sb);
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
