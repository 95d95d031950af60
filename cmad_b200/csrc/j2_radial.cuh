// J2 radial-return solve of one material point, shared by the material-point
// kernel K1-J2 (mp_update_j2.cu) and the FE element-block kernels (fe_block.cu).
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
// norms to |f|.  The routine iterates on the scalar alpha with the reference's
// loop structure (same tests in the same order, same Armijo / quadratic-
// backtracking line search, cmad/util/line_search.py:95-189) and produces the
// same iterates up to rounding, the same iteration counts and branch flags.
//
// Whenever an evaluated iterate leaves the regime in which the reduction holds
// (an iterate or line-search probe on the elastic branch, a return past the
// origin of the deviatoric plane, non-finite values, an elastic entry state that
// is not already converged) `bail` is set and the caller re-solves the point
// with the generic 7x7 Newton.
#pragma once
#include "point_solver.cuh"

namespace cmadx {

constexpr double R32 = 1.2247448713915890491;   // sqrt(3/2)

struct J2Scalar {
    double f, eD;
    bool ok;      // plastic branch and radial reduction valid at this alpha
};

// yield function at alpha on the radial path: s = (sn0 - 2mu*sqrt(3/2)*dgamma) s^
CMADX_DEV J2Scalar j2_eval_alpha(const DevMat& m, double alpha, double alpha0, double sn0) {
    J2Scalar r;
    const double dg = alpha - alpha0;
    const double sn = fma(-m.two_mu * R32, dg, sn0);
    const double phi = R32 * sn;
    double Hd = 0.0;
    r.eD = 0.0;
    if (m.hmask & CMADX_HARD_VOCE) { r.eD = exp(-m.D * alpha); Hd = m.S * (1.0 - r.eD); }
    if (m.hmask & CMADX_HARD_LINEAR) Hd = fma(m.K, alpha, Hd);
    r.f = (phi - (m.Y + Hd)) * m.inv_two_mu;
    const bool plastic = (r.f > m.yield_tol) || (fabs(r.f) < m.yield_tol);
    r.ok = plastic && (sn > 0.0);
    return r;
}

CMADX_DEV double j2_hardening_slope(const DevMat& m, double eD) {
    double Hp = 0.0;
    if (m.hmask & CMADX_HARD_VOCE) Hp = m.S * m.D * eD;
    if (m.hmask & CMADX_HARD_LINEAR) Hp += m.K;
    return Hp * m.inv_two_mu;
}

struct J2Radial {
    double alpha0, sn0;   // entry state
    double n0[6];         // yield normal at x0 (direction is invariant along the solve)
    double sh[6];         // s/||s|| at x0
    double alpha, f, eD;  // solution: alpha*, yield function and exp(-D alpha) there
    double nc;            // final ||C||
    int ii;               // Newton updates
    int flag_entry;       // plastic branch at x0
    bool plastic;         // branch at x* (unchanged along a valid radial solve)
    bool bail;            // hand the point to the generic solver
};

// `live` lanes take part; the loop exit is decided warp-wide by ballot, so all
// 32 lanes of the warp must call this together.
CMADX_DEV void j2_radial_solve(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                               const double (&em)[6], bool live, J2Radial& r) {
    // state at x0 = xi_prev, evaluated by the generic residual (identical arithmetic)
    SepPoint<CMADX_YIELD_J2> pt;
    double C0[7];
    pt.residual(m, xp, xp, em, C0);
    r.flag_entry = pt.plastic ? 1 : 0;
    r.plastic = pt.plastic;
    const double alpha0 = xp[6];
    const double sn0 = pt.yf.sn;
    r.alpha0 = alpha0;
    r.sn0 = sn0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        // elastic entry: the normal is unused (and NaN at a zero deviator, where the
        // reference's jnp.where masks it, paths.py:27) - keep it out of the outputs
        r.n0[a] = pt.plastic ? pt.n[a] : 0.0;
        r.sh[a] = pt.plastic ? pt.yf.sh[a] : 0.0;
    }

    double alpha = alpha0;
    double f = C0[6];            // plastic: yield function; elastic: dgamma = 0
    double eD = pt.eD;
    bool bail = false;
    int ii = 0;
    double nc = sqrt(f * f), n0 = nc;
    bool done = !live || nw.max_iters <= 0;
    bool fresh = true;           // (f, eD) belong to the current alpha
    if (live && !pt.plastic) {
        // elastic entry: C_e(x0) = 0, converged on the absolute test with ii = 0;
        // anything else (abs_tol <= 0) is left to the generic kernel
        if (!done && !(nc < nw.abs_tol)) bail = true;
        done = true;
    }
    if (live && pt.plastic && !(sn0 > 0.0)) { bail = true; done = true; }

    const unsigned full = 0xffffffffu;
    if (nw.mode == CMADX_NEWTON_TRACED) {
        while (__any_sync(full, !done)) {
            if (!done) {
                nc = sqrt(f * f);
                const double rel = nc / n0;
                if (rel < nw.rel_tol || nc < nw.abs_tol) {
                    done = true;
                } else {
                    const double h = j2_hardening_slope(m, eD);
                    const double dxa = f / -(1.5 + h);             // alpha component of solve(J, C)
                    const double CC = f * f;
                    const double phi0 = 0.5 * CC, dphi0 = -CC, armijo = nw.c1 * dphi0;
                    int ne = 0;
                    double al = 1.0, best_al = 1.0, best_phi = CUDART_INF;
                    J2Scalar best; best.f = f; best.eD = eD; best.ok = true;
                    J2Scalar tr = best;
                    bool acc = false;
                    while (ne < nw.ls_max && !acc) {
                        tr = j2_eval_alpha(m, fma(-al, dxa, alpha), alpha0, sn0);
                        if (!tr.ok) bail = true;
                        const double ph = 0.5 * (tr.f * tr.f);
                        const bool fin = isfinite(ph);
                        if (fin && ph < best_phi) { best_al = al; best_phi = ph; best = tr; }
                        acc = fin && (ph <= fma(al, armijo, phi0));
                        const double den = 2.0 * (ph - phi0 - dphi0 * al);
                        const double am = (den == 0.0) ? 0.5 * al : -dphi0 * al * al / den;
                        double ac = fmin(fmax(am, nw.bmin * al), nw.bmax * al);
                        if (am != am) ac = am;
                        if (!acc) al = fin ? ac : 0.5 * al;
                        ++ne;
                    }
                    const double ar = acc ? al : best_al;
                    alpha = fma(-ar, dxa, alpha);
                    f = acc ? tr.f : best.f;
                    eD = acc ? tr.eD : best.eD;
                    ++ii;
                    if (ii >= nw.max_iters || bail) done = true;
                }
            }
        }
        nc = sqrt(f * f);
    } else {
        while (__any_sync(full, !done)) {
            if (!done) {
                if (ii > 0) {
                    const J2Scalar cur = j2_eval_alpha(m, alpha, alpha0, sn0);
                    f = cur.f; eD = cur.eD; fresh = true;
                    if (!cur.ok) bail = true;
                }
                nc = sqrt(f * f);
                double rel = 1.0;
                if (ii == 0) n0 = nc; else rel = nc / n0;
                if (rel < nw.rel_tol || nc < nw.abs_tol || bail) {
                    done = true;
                } else {
                    const double h = j2_hardening_slope(m, eD);
                    alpha += (-f) / -(1.5 + h);                     // solve(J, -C), x += delta
                    fresh = false;
                    ++ii;
                    if (ii >= nw.max_iters) done = true;
                }
            }
        }
        if (live && !fresh) {
            const J2Scalar cur = j2_eval_alpha(m, alpha, alpha0, sn0);
            f = cur.f; eD = cur.eD;
            if (!cur.ok) bail = true;
        }
    }
    if (!isfinite(f) || !isfinite(alpha)) bail = true;
    r.alpha = alpha; r.f = f; r.eD = eD; r.nc = nc; r.ii = ii;
    r.bail = live && bail;
}

// Closed-form IFT coefficients at the radial solution.  For a strain
// perturbation E:  X = [A^-1]_11 E = E - g1 (dev E - s^(s^:E)) - g2 s^(s^:E),
//   g1 = beta/(1+beta), g2 = (3/2)/(3/2 + h), beta = dgamma 2mu sqrt(3/2)/||s*||;
//   dalpha = sqrt(3/2)(s^:E)/(3/2+h);  d sigma = lam tr(E) I + 2mu X.
struct J2Tangent {
    double g1, g2, ga;
};
CMADX_DEV J2Tangent j2_tangent_coeffs(const DevMat& m, const J2Radial& r) {
    J2Tangent t;
    const double dg = r.alpha - r.alpha0;
    const double snf = fma(-m.two_mu * R32, dg, r.sn0);
    const double beta = dg * m.two_mu * (R32 / snf);
    const double h = j2_hardening_slope(m, r.eD);
    t.g1 = r.plastic ? beta / (1.0 + beta) : 0.0;
    t.g2 = r.plastic ? 1.5 / (1.5 + h) : 0.0;
    t.ga = r.plastic ? R32 / (1.5 + h) : 0.0;
    return t;
}

// b <- A^{-1} b for the J2 Jacobian dC/dxi at the state held by `pt` (any state, not
// only a converged one): A = [[I + beta(Pdev - s^ (W s^)^T), -n], [-(W n)^T, -h]] with
// beta = dgamma 2mu sqrt(3/2)/||s||, n = sqrt(3/2) s^.  Splitting the strain-like part of
// b into its spherical part, its component y along s^ and the deviatoric remainder
// b_perp gives  alpha = -(b_a + sqrt(3/2) s^:b)/(3/2 + h),  y = s^:b + sqrt(3/2) alpha,
// e = sph(b) + y s^ + b_perp/(1 + beta).  Elastic branch: A = I.
// With TRANSPOSED the system is A^T: in the packed representation A = T W (T symmetric,
// W = diag(1,2,2,1,2,1,1)), so A^{-T} c = W A^{-1} W^{-1} c.
template <bool TRANSPOSED>
CMADX_DEV void j2_jacobian_solve(const DevMat& m, const SepPoint<CMADX_YIELD_J2>& pt, double dg,
                                 double (&b)[7]) {
    if (!pt.plastic) return;
    if (TRANSPOSED) {
#pragma unroll
        for (int a = 0; a < 6; ++a) b[a] *= is_diag(a) ? 1.0 : 0.5;
    }
    const double beta = dg * m.two_mu * pt.yf.c;
    const double h = j2_hardening_slope(m, pt.eD);
    double sb = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) sb = fma(mult(a) * pt.yf.sh[a], b[a], sb);
    const double sph = (b[0] + b[3] + b[5]) / 3.0;
    const double alpha = -(b[6] + R32 * sb) / (1.5 + h);
    const double y = fma(R32, alpha, sb);
    const double ib = 1.0 / (1.0 + beta);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const double dev = is_diag(a) ? b[a] - sph : b[a];
        const double perp = fma(-pt.yf.sh[a], sb, dev);
        b[a] = fma(pt.yf.sh[a], y, perp * ib) + (is_diag(a) ? sph : 0.0);
    }
    b[6] = alpha;
    if (TRANSPOSED) {
#pragma unroll
        for (int a = 0; a < 6; ++a) b[a] *= mult(a);
    }
}

}  // namespace cmadx
