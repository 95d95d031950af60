// Output helpers shared by the material-point kernels (generic 7x7 Newton and
// the J2 radial-return specialisation).
#pragma once
#include <type_traits>

#include "mp_update.cuh"

namespace cmadx {

// streaming store (write-once data, keep L2 for the inputs of the next launch); translation units
// whose stores are scattered over partially written sectors (mp_update_queue.cu) define
// CMADX_ST_WRITEBACK so the sectors can be completed in L2 before they are evicted
CMADX_DEV void st(double* p, int64_t c, int64_t ld, int64_t i, double v) {
#ifdef CMADX_ST_WRITEBACK
    p[c * ld + i] = v;
#else
    __stcs(p + c * ld + i, v);
#endif
}

// xi_prev (7) and the symmetric strain (6) of point i
CMADX_DEV void load_point(const cmadx_mp_buffers_t& b, int64_t i, bool live,
                          double (&xp)[7], double (&e)[6]) {
    const int64_t ld = b.ld;
    if (live) {
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + c * ld + i);
        if (b.strain_comps == 6) {
#pragma unroll
            for (int c = 0; c < 6; ++c) e[c] = __ldg(b.strain + c * ld + i);
        } else {
            double g[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) g[c] = __ldg(b.strain + c * ld + i);
            e[0] = g[0]; e[3] = g[4]; e[5] = g[8];
            e[1] = 0.5 * (g[1] + g[3]); e[2] = 0.5 * (g[2] + g[6]); e[4] = 0.5 * (g[5] + g[7]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = 0.0;
    }
}

// dC/dxi_prev: plastic rows a<6: -I and +n in the alpha column, yield row 0;
// elastic: -I
CMADX_DEV void write_dC_dxi_prev(double* out, int64_t ld, int64_t i, bool pl, const double (&n)[6]) {
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            double v = (r == c) ? -1.0 : 0.0;
            if (pl) {
                if (r == 6) v = 0.0;
                else if (c == 6) v = n[r];
            }
            st(out, r * 7 + c, ld, i, v);
        }
}

// one column of dC/dp at (x*, x_prev) for canonical parameter id `pid`.
// Elastic branch -> 0 (C_e holds no parameters).  Mee = (dn/dsigma : ee),
// nee = n : ee, sig = material cauchy (for the yield-surface parameters).
template <class YF>
CMADX_DEV void dC_dp_column(const DevMat& m, int pid, bool pl, const YF& yf, const double (&n)[6],
                            double f, double eD, double alpha, double dg,
                            const double (&Mee)[6], double nee, const double (&sig)[6],
                            double (&col)[7]) {
#pragma unroll
    for (int r = 0; r < 7; ++r) col[r] = 0.0;
    if (!pl) return;
    if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
        // only mu matters: all three surfaces are pressure-insensitive
        const double dmu = m.dmu[pid - CMADX_P_EL0];
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = -2.0 * dg * Mee[a] * dmu;
        col[6] = (nee - f) / m.mu * dmu;
    } else if (pid == CMADX_P_Y) {
        col[6] = -m.inv_two_mu;
    } else if (pid == CMADX_P_VOCE_S) {
        col[6] = -(1.0 - eD) * m.inv_two_mu;
    } else if (pid == CMADX_P_VOCE_D) {
        col[6] = -m.S * alpha * eD * m.inv_two_mu;
    } else if (pid == CMADX_P_LIN_K) {
        col[6] = -alpha * m.inv_two_mu;
    } else {
        double dphi, dn[6];
        if (yf.dparam(m, pid, sig, dphi, dn)) {
#pragma unroll
            for (int a = 0; a < 6; ++a) col[a] = -dg * dn[a];
            col[6] = dphi * m.inv_two_mu;
        }
    }
}

template <class YF>
CMADX_DEV void write_dC_dp(const MpArgs& A, int64_t i, bool pl, const YF& yf, const double (&n)[6],
                           double f, double eD, double alpha, double dg,
                           const double (&Mee)[6], double nee, const double (&sig)[6]) {
    const int na = A.n_active;
    for (int c = 0; c < na; ++c) {
        double col[7];
        dC_dp_column(A.m, A.pid[c], pl, yf, n, f, eD, alpha, dg, Mee, nee, sig, col);
#pragma unroll
        for (int r = 0; r < 7; ++r) st(A.b.dC_dp, (int64_t)r * na + c, A.b.ld, i, col[r]);
    }
}

// 6x6 maps between global and material symmetric-tensor components for a
// rotation Q (cmad/models/small_elastic_plastic.py:44-62, 318-319):
//   T[c][b] = d(Q^T e Q)_c / d e_b ,  S[a][c] = d(Q s Q^T)_a / d s_c
CMADX_DEV void rot_maps(const double* Q, double (&T)[6][6], double (&S)[6][6]) {
    const int ci[6] = {0, 0, 0, 1, 1, 2}, cj[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const int i = ci[c], j = cj[c], k = ci[b], l = cj[b];
            double t = Q[3 * k + i] * Q[3 * l + j];
            double s = Q[3 * i + k] * Q[3 * j + l];
            if (k != l) { t += Q[3 * l + i] * Q[3 * k + j]; s += Q[3 * i + l] * Q[3 * j + k]; }
            T[c][b] = t;
            S[c][b] = s;
        }
}


// --------------------------------------------------------------------------
// Everything a material-point update writes besides the Newton solve itself, at
// the converged state x* (7 comps, material axes) of point i: xi, cauchy, dC/dp,
// dC/dxi, dC/dxi_prev and the IFT products d(xi, sigma)/d(strain)
// (cmad/models/nonlinear_solver.py:158-171, cmad/models/model.py:121-166).
// `pt` must hold the state of the last residual evaluation AT x* (yield normal,
// yield-surface internals, f, exp(-D alpha)); `flags` bit 1 is the branch there.
// Shared by the one-pass kernels (mp_update.cu) and the streaming kernel
// (mp_update_stream.cu).
// --------------------------------------------------------------------------
template <int YK, bool ROT, bool REDUCED, class Pt>
CMADX_DEV void write_point_outputs(const MpArgs& A, const int64_t i, const double (&x)[7],
                                   const double alpha_prev, const double (&em)[6], const Pt& pt,
                                   const int iters, const int flags, const double cnorm) {
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    if (A.b.iters) A.b.iters[i] = iters;
    if (A.b.flags) A.b.flags[i] = flags;
    if (A.b.cnorm) A.b.cnorm[i] = cnorm;
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
        if (ROT) {
            double T[6][6], S[6][6];
            rot_maps(m.Q, T, S);
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) s = fma(S[a][c], sig[c], s);
                st(A.b.sigma, a, ld, i, s);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) st(A.b.sigma, a, ld, i, sig[a]);
        }
    }
    const double dg = x[6] - alpha_prev;
    const bool pl = (flags & 2) != 0;       // branch at x* as the Newton loop saw it

    if (A.b.dC_dxi_prev) write_dC_dxi_prev(A.b.dC_dxi_prev, ld, i, pl, pt.n);

    // dC/dp at (x*, x_prev): elastic branch -> 0 (C_e holds no parameters)
    if (A.b.dC_dp && A.n_active > 0) {
        double Mee[6];      // (dn/dsigma : ee)_a
        double nee = 0.0;   // n : ee
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s = fma(pt.yf.M(a, b), ee[b], s);
            Mee[a] = s;
            nee = fma(mult(a) * pt.n[a], ee[a], nee);
        }
        write_dC_dp(A, i, pl, pt.yf, pt.n, pt.f, pt.eD, x[6], dg, Mee, nee, sig);
        // rotation-matrix leaves (the reference's jacrev treats the 9 entries of Q as independent,
        // cmad/parameters/parameters.py:368-377): C sees Q only through e_m = Q^T eps Q
        // (small_elastic_plastic.py:44-62), so dC/dQ_ij = dC/de_m : d e_m/dQ_ij with
        // d(e_m)_kl/dQ_ij = delta_jk W_il + delta_jl W_ik, W = eps Q (= Q e_m for the orthonormal Q
        // the reference builds), and dC/de_m = -(dC/dx[:, :6] - [I6; 0]) in material axes.
        for (int c = 0; c < A.n_active; ++c) {
            const int q = A.pid[c] - CMADX_P_Q00;
            if (q < 0 || q >= 9) continue;
            double col[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (pl) {
                const int qi = q / 3, qj = q % 3;
                const double E3[3][3] = {{em[0], em[1], em[2]}, {em[1], em[3], em[4]}, {em[2], em[4], em[5]}};
                double W[3];                            // row qi of W = Q e_m
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) s = fma(m.Q[3 * qi + k], E3[k][l], s);
                    W[l] = s;
                }
                const int ck[6] = {0, 0, 0, 1, 1, 2}, cl[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    const double D = ((ck[b] == qj) ? W[cl[b]] : 0.0) + ((cl[b] == qj) ? W[ck[b]] : 0.0);
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const double dCde = ((r == b) ? 1.0 : 0.0) - full_jacobian_entry(m, pt, dg, r, b);
                        col[r] = fma(dCde, D, col[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 7; ++r) st(A.b.dC_dp, (int64_t)r * A.n_active + c, ld, i, col[r]);
        }
    }

    const bool want_ift = A.b.dsig_deps || A.b.dxi_deps;
    if (!want_ift && !A.b.dC_dxi) return;

    if (A.b.dC_dxi) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) st(A.b.dC_dxi, r * 7 + c, ld, i, full_jacobian_entry(m, pt, dg, r, c));
    }
    if (!want_ift) return;

    // IFT (nonlinear_solver.py:158-171).  In material axes dC/de = -(A[:, :6] - E),
    // E = [I6; 0], so dx/de = E - A^{-1}E and d sigma/de = Cel . (A^{-1})[0:6,0:6].
    // threshold pivoting (see RegLU): natural order unless some lane is troubled.
    // Strain components that are not unknowns of a reduced point have A^{-1} e_b = e_b.
    RegLU<N> lu;
    pt.jacobian(m, dg, lu.a);
    bool trouble = false;
    if (__any_sync(__activemask(), pl)) trouble = lu.factor_natural() && pl;
    const bool slow = __any_sync(__activemask(), trouble);
    if (slow && trouble) {
        pt.jacobian(m, dg, lu.a);
        lu.factor_pivot();
    }
    auto solve_dir = [&](int b, double (&X)[7]) {
#pragma unroll
        for (int r = 0; r < 7; ++r) X[r] = (r == b) ? 1.0 : 0.0;
        if (Tr::local(b) >= 0 && pl) {
            double Xl[N];
#pragma unroll
            for (int k = 0; k < N; ++k) Xl[k] = (k == Tr::local(b)) ? 1.0 : 0.0;
            if (slow && trouble) lu.solve_pivot(Xl); else lu.solve_natural(Xl);
#pragma unroll
            for (int k = 0; k < N; ++k) X[Tr::full(k)] = Xl[k];
        }
    };
    if (!ROT) {
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            double X[7];
            solve_dir(b, X);
            if (A.b.dxi_deps) {
#pragma unroll
                for (int r = 0; r < 7; ++r) st(A.b.dxi_deps, r * 6 + b, ld, i, ((r == b) ? 1.0 : 0.0) - X[r]);
            }
            if (A.b.dsig_deps) {
                const double ltr = m.lam * (X[0] + X[3] + X[5]);
#pragma unroll
                for (int a = 0; a < 6; ++a)
                    st(A.b.dsig_deps, a * 6 + b, ld, i, is_diag(a) ? fma(m.two_mu, X[a], ltr) : m.two_mu * X[a]);
            }
        }
    } else {
        double Dm[6][6], Xm[7][6];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            double X[7];
            solve_dir(b, X);
            const double ltr = m.lam * (X[0] + X[3] + X[5]);
#pragma unroll
            for (int a = 0; a < 6; ++a) Dm[a][b] = is_diag(a) ? fma(m.two_mu, X[a], ltr) : m.two_mu * X[a];
#pragma unroll
            for (int r = 0; r < 7; ++r) Xm[r][b] = ((r == b) ? 1.0 : 0.0) - X[r];
        }
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
        if (A.b.dxi_deps) {
#pragma unroll
            for (int r = 0; r < 7; ++r)
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) s = fma(Xm[r][c], T[c][b], s);
                    st(A.b.dxi_deps, r * 6 + b, ld, i, s);
                }
        }
        if (A.b.dsig_deps) {
            double DT[6][6];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) s = fma(Dm[a][c], T[c][b], s);
                    DT[a][b] = s;
                }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) s = fma(S[a][c], DT[c][b], s);
                    st(A.b.dsig_deps, a * 6 + b, ld, i, s);
                }
        }
    }
}

}  // namespace cmadx
