// SmallRateElasticPlastic under the PLANE_STRESS / UNIAXIAL_STRESS deformation types: the point type
// shared by K1 (mp_update_rate_dt.cu) and K2 (mp_sens_rate_dt.cu).  See mp_update_rate_dt.cu for the
// model, the layout of the state and the references.
#pragma once
#include "rate_point.cuh"
#include "sep_point_dt.cuh"

namespace cmadx {
namespace {

template <int YK, int DT>
struct RatePointDT {
    static constexpr int NZ = (DT == CMADX_DEF_PLANE_STRESS) ? 1 : 2;
    static constexpr int ND = (DT == CMADX_DEF_PLANE_STRESS) ? 0 : 3;
    static constexpr int NR = NZ + ND;                  // constraint rows
    static constexpr int N = 7 + NR, ALPHA = 6;
    RatePoint<YK> b;
    bool plastic;
    double T[6][6], S[6][6];
    double dem[6];                                      // material strain increment of the last evaluation

    // global component moved by unknown 7 + r / constrained by row 7 + r
    CMADX_DEV static constexpr int ccomp(int r) {
        return (DT == CMADX_DEF_PLANE_STRESS) ? 5 : (r == 0 ? 3 : (r == 1 ? 5 : (r == 2 ? 1 : (r == 3 ? 2 : 4))));
    }

    CMADX_DEV void maps(const DevMat& m) {
        if (m.rot) {
            rot_maps(m.Q, T, S);
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c) { T[a][c] = (a == c) ? 1.0 : 0.0; S[a][c] = T[a][c]; }
        }
    }
    CMADX_DEV void qrow(const DevMat& m, int r, double (&q)[6]) const {
        const int c = ccomp(r);
        const double t = m.lam * m.inv_two_mu * (S[c][0] + S[c][3] + S[c][5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) q[a] = S[c][a] + (is_diag(a) ? t : 0.0);
    }

    CMADX_DEV void residual(const DevMat& m, const double (&x)[N], const double (&xp)[N],
                            const double (&em)[6], double (&C)[N]) {
        maps(m);
        double deg[6];
        if (DT == CMADX_DEF_PLANE_STRESS) {
            deg[0] = em[0]; deg[1] = em[1]; deg[2] = 0.0; deg[3] = em[3]; deg[4] = 0.0; deg[5] = x[7] - xp[7];
        } else {
            deg[0] = em[0]; deg[1] = x[9]; deg[2] = x[10]; deg[3] = x[7] - xp[7]; deg[4] = x[11]; deg[5] = x[8] - xp[8];
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(T[a][c], deg[c], s);
            dem[a] = s;
        }
        double x7[7], xp7[7], C7[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) { x7[c] = x[c]; xp7[c] = xp[c]; }
        b.residual(m, x7, xp7, dem, C7);
        plastic = b.plastic;
#pragma unroll
        for (int c = 0; c < 7; ++c) C[c] = C7[c];
        const double dg = x[6] - xp[6];
        double w[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) w[a] = plastic ? fma(-dg, b.n[a], dem[a]) : dem[a];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double q[6], s = 0.0;
            qrow(m, r, q);
#pragma unroll
            for (int a = 0; a < 6; ++a) s = fma(q[a], w[a], s);
            C[7 + r] = s;
        }
    }

    CMADX_DEV void jacobian(const DevMat& m, double dg, double (&J)[N][N]) const {
        double J7[7][7];
        b.jacobian(m, dg, J7);
        const double lr = m.lam * m.inv_two_mu;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
#pragma unroll
            for (int c = 0; c < 7; ++c) J[a][c] = J7[a][c];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                // d C_a / d dem_b = -(delta_ab + lr [a diag][b diag]) on the stress rows, 0 on the yield row
                double v = 0.0;
                if (a < 6) {
                    const int g = ccomp(r);
                    v = -T[a][g];
                    if (is_diag(a)) v -= lr * (T[0][g] + T[3][g] + T[5][g]);
                }
                J[a][7 + r] = v;
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double q[6];
            qrow(m, r, q);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double v = 0.0;
                if (plastic) {
#pragma unroll
                    for (int a = 0; a < 6; ++a) v = fma(q[a], b.yf.M(a, c), v);
                    v *= -dg;
                }
                J[7 + r][c] = v;
            }
            double qn = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) qn = fma(q[a], b.n[a], qn);
            J[7 + r][6] = plastic ? -qn : 0.0;
#pragma unroll
            for (int r2 = 0; r2 < NR; ++r2) {
                double v = 0.0;
#pragma unroll
                for (int a = 0; a < 6; ++a) v = fma(q[a], T[a][ccomp(r2)], v);
                J[7 + r][7 + r2] = v;
            }
        }
    }

    // dC / d(prescribed increment component bc)
    CMADX_DEV void dC_deps(const DevMat& m, int bc, double (&col)[N]) const {
        const double lr = m.lam * m.inv_two_mu;
        const double tr = T[0][bc] + T[3][bc] + T[5][bc];
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = -T[a][bc] - (is_diag(a) ? lr * tr : 0.0);
        col[6] = 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double q[6], v = 0.0;
            qrow(m, r, q);
#pragma unroll
            for (int a = 0; a < 6; ++a) v = fma(q[a], T[a][bc], v);
            col[7 + r] = v;
        }
    }
};

// B = dC/dxi_prev from A = dC/dxi (unfactored) at the same state: sigma_prev / alpha_prev as in the
// FULL_3D form; the stretches enter as z - z_prev (their columns are minus the current ones, as is the
// alpha_prev entry of the constraint rows); the delta strains have no previous value
template <class Pt>
CMADX_DEV double rate_dt_B(const DevMat& m, const Pt& pt, const double (&A)[Pt::N][Pt::N], int r, int c) {
    const bool pl = pt.plastic;
    if (c < 6) return (r == c) ? -m.inv_two_mu : 0.0;
    if (c == 6) return (r < 6) ? (pl ? -pt.b.n[r] : 0.0) : (r == 6 ? (pl ? 0.0 : -1.0) : -A[r][6]);
    if (c < 7 + Pt::NZ) return -A[r][c];
    return 0.0;
}

// one column of dC/dp (N rows) at (x, x_prev); `pt` fresh at x.  Constraint rows: the elastic constants
// through lam / 2mu, the yield-surface leaves through the normal
template <int YK, int DT>
CMADX_DEV void rate_dt_dC_dp_column(const DevMat& m, int pid, const RatePointDT<YK, DT>& pt,
                                    const double (&x)[RatePointDT<YK, DT>::N], const double (&xp)[RatePointDT<YK, DT>::N],
                                    double (&col)[RatePointDT<YK, DT>::N]) {
    using Pt = RatePointDT<YK, DT>;
    const bool pl = pt.plastic;
    const double dg = x[6] - xp[6];
    double x7[7], xp7[7], c7[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) { x7[c] = x[c]; xp7[c] = xp[c]; }
    rate_dC_dp_column<YK>(m, pid, pt.b, x7, xp7, pt.dem, c7);
#pragma unroll
    for (int r = 0; r < 7; ++r) col[r] = c7[r];
    double trw = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) if (is_diag(a)) trw += pl ? fma(-dg, pt.b.n[a], pt.dem[a]) : pt.dem[a];
    double dlr = 0.0, dn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    bool has_dn = false;
    if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
        const int k = pid - CMADX_P_EL0;
        dlr = (m.dlam[k] * m.two_mu - m.lam * 2.0 * m.dmu[k]) * m.inv_two_mu * m.inv_two_mu;
    } else if (pl && pid > CMADX_P_LIN_K) {
        double sig[6], dphi;
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = x[a];
        has_dn = pt.b.yf.dparam(m, pid, sig, dphi, dn);
    }
#pragma unroll
    for (int r = 0; r < Pt::NR; ++r) {
        const int g = Pt::ccomp(r);
        double v = dlr * trw * (pt.S[g][0] + pt.S[g][3] + pt.S[g][5]);
        if (has_dn) {
            double q[6];
            pt.qrow(m, r, q);
#pragma unroll
            for (int a = 0; a < 6; ++a) v = fma(-dg * q[a], dn[a], v);
        }
        col[7 + r] = v;
    }
}

}  // namespace
}  // namespace cmadx
