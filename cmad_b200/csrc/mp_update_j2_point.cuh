// Per-point body of K1-J2 (see mp_update_j2.cu for the derivation).  A header so that the SAME
// source is compiled twice: by nvcc into mp_update_j2_kernel, and by the host compiler into the
// CPU baseline "port-handderived" (oracle/j2_host.cpp).
#pragma once
#include "j2_radial.cuh"
#include "mp_outputs.cuh"

namespace cmadx {

// One material point of the J2 radial-return update: solve + every requested output.
// Returns true when the point left the regime of the reduction ("bail"): the caller hands it to
// the generic Newton, which rewrites every output.  `live == false`: padding lane (takes part in
// the warp votes of the solve, writes nothing).
CMADX_DEV bool j2_point_update(const MpArgs& A, const int64_t i, const bool live) {
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    const DevNewton& nw = A.nw;

    double xp[7], em[6];
    load_point(A.b, i, live, xp, em);

    J2Radial rs;
    j2_radial_solve(m, nw, xp, em, live, rs);   // see j2_radial.cuh
    if (!live) return false;
    if (rs.bail) return true;
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
    if (!A.b.dsig_deps && !A.b.dxi_deps) return false;

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
    return false;
}

}  // namespace cmadx
