// Device routines shared by the FE element-block kernels (fe_tet4.cu, fe_hex8.cu):
// 256-bit vector memory access, the per-integration-point solvers (J2 radial
// return and the generic 7x7 Newton) returning the converged state, the global
// cauchy stress and the consistent tangent d sigma / d eps, and small helpers.
#pragma once
#include <type_traits>

#include "fe_block.cuh"
#include "j2_radial.cuh"
#include "mp_outputs.cuh"

namespace cmadx {
namespace {

constexpr int FE_BLOCK = 128;

// symmetric-tensor component of entry (i, j) in the packing xx,xy,xz,yy,yz,zz
CMADX_DEV constexpr int vix(int i, int j) {
    return (i == j) ? (i == 0 ? 0 : (i == 1 ? 3 : 5)) : ((i + j == 1) ? 1 : ((i + j == 2) ? 2 : 4));
}

CMADX_DEV void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
CMADX_DEV void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
                 : "memory");
}

// mixed u-p (cmad/global_residuals/small_disp_equilibrium.py:87-101): the momentum stress is
// dev(cauchy) - p I with p interpolated from the pressure dofs, so K_uu sees P_dev D
template <bool WANT_D>
CMADX_DEV void mixed_momentum_stress(const double p, double (&sg)[6], double (&D)[6][6]) {
    const double m = (sg[0] + sg[3] + sg[5]) / 3.0 + p;
    sg[0] -= m; sg[3] -= m; sg[5] -= m;
    if (WANT_D) {
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const double t = (D[0][b] + D[3][b] + D[5][b]) / 3.0;
            D[0][b] -= t; D[3][b] -= t; D[5][b] -= t;
        }
    }
}

struct PointOut {
    double x[7];
    double sg[6];   // global cauchy
    int iters, flags;
    bool bail;
};

// ---- J2 radial-return point (see j2_radial.cuh) ---------------------------------
template <bool WANT_D>
CMADX_DEV void point_j2(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                        const double (&e)[6], bool live, PointOut& o, double (&D)[6][6]) {
    J2Radial rs;
    j2_radial_solve(m, nw, xp, e, live, rs);
    o.bail = rs.bail;
    o.iters = rs.ii;
    o.flags = rs.flag_entry | ((rs.plastic ? 1 : 0) << 1);
    const double dg = rs.alpha - rs.alpha0;
#pragma unroll
    for (int a = 0; a < 6; ++a) o.x[a] = fma(dg, rs.n0[a], xp[a]);
    o.x[6] = rs.alpha;
    double ee[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = e[a] - o.x[a];
    const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
#pragma unroll
    for (int a = 0; a < 6; ++a) o.sg[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
    if (WANT_D) {
        const J2Tangent t = j2_tangent_coeffs(m, rs);
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const double sb = mult(b) * rs.sh[b];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double devE = (a == b) ? 1.0 : 0.0;
                if (is_diag(a) && is_diag(b)) devE -= 1.0 / 3.0;
                const double emx = t.g1 * devE + (t.g2 - t.g1) * rs.sh[a] * sb;
                double v = m.two_mu * (((a == b) ? 1.0 : 0.0) - emx);
                if (is_diag(a) && is_diag(b)) v += m.lam;
                D[a][b] = v;
            }
        }
    }
}

// 6x6 maps between global and material symmetric-tensor components for a
// rotation Q (cmad/models/small_elastic_plastic.py:44-62, 318-319)
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

// ---- generic Newton point (point_solver.cuh): 7x7, or the 4x4 reduced system for
// Hosford (the FE path always starts from xi_prev, so the reduction always applies)
template <int YK, bool ROT, bool WANT_D>
CMADX_DEV void point_generic(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                             const double (&e)[6], bool live, PointOut& o, double (&D)[6][6]) {
    constexpr bool REDUCED = (YK == CMADX_YIELD_HOSFORD);
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    double em[6];
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s = fma(T[c][b], e[b], s);
            em[c] = s;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) em[c] = e[c];
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) o.x[c] = xp[c];
    Pt pt;
    if constexpr (REDUCED) { pt.shear[0] = xp[1]; pt.shear[1] = xp[2]; pt.shear[2] = xp[4]; }
    double y[N], yp[N], Cy[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { y[k] = xp[Tr::full(k)]; yp[k] = y[k]; }
    const NewtonResult nr = local_newton<Pt, N>(m, nw, pt, y, yp, em, live, Cy);
#pragma unroll
    for (int k = 0; k < N; ++k) o.x[Tr::full(k)] = y[k];
    o.bail = nr.deferred;        // needs more than nw.defer_after updates: second (list) pass
    o.iters = nr.iters;
    o.flags = nr.flag_entry | ((pt.plastic ? 1 : 0) << 1);
    double sig[6];
    {
        double ee[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = em[a] - o.x[a];
        const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
    }
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(S[a][c], sig[c], s);
            o.sg[a] = s;
        }
    } else {
#pragma unroll
        for (int a = 0; a < 6; ++a) o.sg[a] = sig[a];
    }
    if (!WANT_D) return;

    // IFT (nonlinear_solver.py:158-171): d sigma/d eps = Cel . (A^{-1})[0:6,0:6] in material axes
    const bool pl = pt.plastic;
    const double dg = o.x[6] - xp[6];
    RegLU<N> lu;
    pt.jacobian(m, dg, lu.a);
    bool trouble = false;
    if (__any_sync(__activemask(), pl)) trouble = lu.factor_natural() && pl;
    const bool slow = __any_sync(__activemask(), trouble);
    if (slow && trouble) {
        pt.jacobian(m, dg, lu.a);
        lu.factor_pivot();
    }
    double Dm[6][6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        double X[7];
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
        const double ltr = m.lam * (X[0] + X[3] + X[5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) Dm[a][b] = is_diag(a) ? fma(m.two_mu, X[a], ltr) : m.two_mu * X[a];
    }
    if (!ROT) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) D[a][b] = Dm[a][b];
    } else {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
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
                D[a][b] = s;
            }
    }
}

// ---- K6: tangent (JVP) of the converged point w.r.t. (params, xi_prev) at fixed strain
// What jax.jvp pushes through make_newton_solve's custom_jvp rule for the FE
// sensitivities (cmad/models/nonlinear_solver.py:158-171 called from
// cmad/fem/nonlinear_solver.py:490-537):
//   dxi = -A^{-1} (dC/dp dp + dC/dxi_prev dxi_prev + dC/deps deps),
//   d sigma = d cauchy/dxi dxi + d cauchy/dp dp + d cauchy/deps deps,
// evaluated AT the given converged state; `de` is the symmetric strain direction of an
// optional displacement direction (zero for the fixed-U residual JVP).  With
// dC/deps = -(A[:, :6] - E), E = [I6; 0]:  dxi = v - A^{-1}(r0 + v), v = [deps; 0].
template <int YK, bool ROT>
CMADX_DEV void point_jvp(const FeArgs& A, const double (&xp)[7], const double (&xs)[7],
                         const double (&dxp)[7], const double (&e)[6], const double (&de)[6],
                         bool live, PointOut& o) {
    const DevMat& m = A.m;
    double em[6];
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s = fma(T[c][b], e[b], s);
            em[c] = s;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) em[c] = e[c];
    }
    double dem[6];
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s = fma(T[c][b], de[b], s);
            dem[c] = s;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) dem[c] = de[c];
    }
    SepPoint<YK> pt;
    double Cs[7];
    pt.residual(m, xs, xp, em, Cs);
    const bool pl = pt.plastic;
    const double dg = xs[6] - xp[6];
    double ee[6], sig[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = em[a] - xs[a];
    const double tre = ee[0] + ee[3] + ee[5];
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], m.lam * tre) : m.two_mu * ee[a];
    double Mee[6], nee = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) s = fma(pt.yf.M(a, b), ee[b], s);
        Mee[a] = s;
        nee = fma(mult(a) * pt.n[a], ee[a], nee);
    }
    double rhs[7], dlam = 0.0, dmu = 0.0;
#pragma unroll
    for (int r = 0; r < 7; ++r) rhs[r] = 0.0;
    for (int c = 0; c < A.n_active; ++c) {
        double col[7];
        dC_dp_column(m, A.pid[c], pl, pt.yf, pt.n, pt.f, pt.eD, xs[6], dg, Mee, nee, sig, col);
#pragma unroll
        for (int r = 0; r < 7; ++r) rhs[r] = fma(col[r], A.dp[c], rhs[r]);
        if (A.pid[c] == CMADX_P_EL0 || A.pid[c] == CMADX_P_EL1) {
            dlam = fma(m.dlam[A.pid[c] - CMADX_P_EL0], A.dp[c], dlam);
            dmu = fma(m.dmu[A.pid[c] - CMADX_P_EL0], A.dp[c], dmu);
        }
    }
    // + dC/dxi_prev dxi_prev: plastic rows a<6: -I and +n in the alpha column, yield row 0; elastic: -I
#pragma unroll
    for (int a = 0; a < 6; ++a) rhs[a] += pl ? fma(pt.n[a], dxp[6], -dxp[a]) : -dxp[a];
    rhs[6] += pl ? 0.0 : -dxp[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) rhs[a] += dem[a];                 // r0 + v
    newton_direction<SepPoint<YK>, 7>(m, pt, dg, rhs);          // rhs <- A^{-1} rhs
#pragma unroll
    for (int r = 0; r < 7; ++r) o.x[r] = live ? ((r < 6 ? dem[r] : 0.0) - rhs[r]) : 0.0;
    // d sigma (material axes) = Cel (d eps - d ep) + (d lam tr(ee) I + 2 d mu ee)
    double dee[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) dee[a] = dem[a] - o.x[a];
    const double trd = dee[0] + dee[3] + dee[5];
    double ds[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const double v = fma(2.0 * dmu, ee[a], m.two_mu * dee[a]);
        ds[a] = is_diag(a) ? v + fma(m.lam, trd, dlam * tre) : v;
    }
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(S[a][c], ds[c], s);
            o.sg[a] = s;
        }
    } else {
#pragma unroll
        for (int a = 0; a < 6; ++a) o.sg[a] = ds[a];
    }
    o.bail = false;
    o.iters = 0;
    o.flags = pl ? 3 : 0;
}

// SOLVER: 0 = J2 radial return (may bail), 1 + YK = generic Newton for yield surface YK,
// 4 + YK = K6 (JVP at a given state; handled by the kernels through point_jvp)
constexpr int FE_JVP = 4;
template <int SOLVER, bool ROT, bool WANT_D>
CMADX_DEV void solve_point(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                           const double (&e)[6], bool live, PointOut& o, double (&D)[6][6]) {
    if (SOLVER == 0) point_j2<WANT_D>(m, nw, xp, e, live, o, D);
    else point_generic<(SOLVER > 0 ? SOLVER - 1 : 0), ROT, WANT_D>(m, nw, xp, e, live, o, D);
}

// symmetric strain of grad_u[k][j] = sum_a U[a][k] gN[a][j]
template <int NB>
CMADX_DEV void strain_from_U(const double (&U)[NB][3], const double (&gN)[NB][3], double (&e)[6]) {
    double g[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < NB; ++a) s = fma(U[a][k], gN[a][j], s);
            g[k][j] = s;
        }
    e[0] = g[0][0]; e[3] = g[1][1]; e[5] = g[2][2];
    e[1] = 0.5 * (g[0][1] + g[1][0]); e[2] = 0.5 * (g[0][2] + g[2][0]); e[4] = 0.5 * (g[1][2] + g[2][1]);
}

CMADX_DEV void append_bail(const FeArgs& A, int64_t e) {
    const unsigned slot = atomicAdd(A.bail_count, 1u);
    if (slot < A.bail_cap) A.bail_list[slot] = (int)e;
}


// launch-time dispatch shared by both element families: SOLVER 0 = J2 radial,
// 1 + YK = generic; ROT only exists for the generic solvers
template <template <int, bool, bool, bool> class Launcher, bool LIST>
cudaError_t dispatch_fe(const FeArgs& A, int solver, cudaStream_t stream, int sms) {
    const bool k = A.b.K_elem != nullptr;
    const bool rot = A.m.rot != 0;
#define CMADX_FE_CASE(S)                                                                         \
    case S:                                                                                      \
        if (S != 0 && rot)                                                                       \
            return k ? Launcher<(S ? S : 1), true, true, LIST>::run(A, stream, sms)             \
                     : Launcher<(S ? S : 1), true, false, LIST>::run(A, stream, sms);           \
        return k ? Launcher<S, false, true, LIST>::run(A, stream, sms)                           \
                 : Launcher<S, false, false, LIST>::run(A, stream, sms);
    switch (solver) {
        CMADX_FE_CASE(0)
        CMADX_FE_CASE(1)
        CMADX_FE_CASE(2)
        CMADX_FE_CASE(3)
    case 4: return rot ? Launcher<4, true, false, LIST>::run(A, stream, sms) : Launcher<4, false, false, LIST>::run(A, stream, sms);
    case 5: return rot ? Launcher<5, true, false, LIST>::run(A, stream, sms) : Launcher<5, false, false, LIST>::run(A, stream, sms);
    case 6: return rot ? Launcher<6, true, false, LIST>::run(A, stream, sms) : Launcher<6, false, false, LIST>::run(A, stream, sms);
    }
#undef CMADX_FE_CASE
    return cudaErrorInvalidValue;
}

// list mode (second pass): the generic solver of the block's yield surface, never deferring
template <template <int, bool, bool, bool> class Launcher>
cudaError_t dispatch_fe_list(const FeArgs& A, int solver, cudaStream_t stream, int sms) {
    const bool k = A.b.K_elem != nullptr;
    const bool rot = A.m.rot != 0;
#define CMADX_FE_LIST_CASE(S)                                                                    \
    case S:                                                                                      \
        if (rot) return k ? Launcher<S, true, true, true>::run(A, stream, sms)                   \
                          : Launcher<S, true, false, true>::run(A, stream, sms);                 \
        return k ? Launcher<S, false, true, true>::run(A, stream, sms)                           \
                 : Launcher<S, false, false, true>::run(A, stream, sms);
    switch (solver) {
        CMADX_FE_LIST_CASE(1)
        CMADX_FE_LIST_CASE(2)
        CMADX_FE_LIST_CASE(3)
    }
#undef CMADX_FE_LIST_CASE
    return cudaErrorInvalidValue;
}

}  // namespace
}  // namespace cmadx
