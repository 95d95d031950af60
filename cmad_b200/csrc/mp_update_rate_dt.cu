// K1 for SmallRateElasticPlastic under the PLANE_STRESS / UNIAXIAL_STRESS deformation types
// (cmad/models/small_rate_elastic_plastic.py:34-77 kinematics, :296-345 constraint rows; the
// configurations tests/models/test_elastic_plastic_models.py runs as "small rate").
//
// State x = [cauchy_m(6), alpha, z (1 | 2 stretches), d (uniaxial: 3 off-axis delta strains)], n_xi = 8 | 12.
// Global strain increment, packed xx,xy,xz,yy,yz,zz (the `strain` rows carry the INCREMENT of the
// prescribed in-plane 2x2 / axial part, or totals + strain_prev as for FULL_3D):
//   plane stress   : deg = [de_xx, de_xy, 0, de_yy, 0, z0 - z0_prev]
//   uniaxial stress: deg = [de_xx, d0, d1, z0 - z0_prev, d2, z1 - z1_prev]
// material increment dem = T deg (rot_maps; identity axes: T = S = I), FULL_3D rows from RatePoint<YK>
// (rate_point.cuh), and the constraint rows on the GLOBAL stress increment
//   R_r = (S Cel w)[c_r] / 2 mu,  w = dem - dgamma n (plastic) | dem (elastic),
//   c_r = zz (plane stress) | yy, zz, xy, xz, yz (uniaxial stress).
// Hand-derived Jacobian through G = d dem / d x (columns T[:, c] on the stretch / delta-strain
// unknowns): rows a < 6 gain -(I + (lam / 2mu) 1 1^T) G, the constraint rows are
// q_r^T (G - dgamma M [sigma cols] - n [alpha col]) with q_r = S[c_r, :] + (lam / 2mu) tr-part.
// Same Newton state machine, register LU (N = 8 | 12) and output conventions as the other K1 kernels.
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

template <int YK, int DT>
__global__ void __launch_bounds__(MP_BLOCK) mp_update_rate_dt_kernel(const __grid_constant__ MpArgs A) {
    using Pt = RatePointDT<YK, DT>;
    constexpr int N = Pt::N, NZ = Pt::NZ, NR = Pt::NR;
    constexpr int NS = (DT == CMADX_DEF_PLANE_STRESS) ? 3 : 1;
    const int scomp[3] = {0, (DT == CMADX_DEF_PLANE_STRESS) ? 1 : 0, 3};
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.b.n;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    double xp[N], x[N], em[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (live) {
#pragma unroll
        for (int c = 0; c < N; ++c) xp[c] = __ldg(A.b.xi_prev + c * ld + i);
        load_dt_strain<DT>(A.b.strain, A.b.strain_comps, ld, i, em);
        if (A.b.strain_prev) {           // totals given: form the increment here
            double ep[6];
            load_dt_strain<DT>(A.b.strain_prev, A.b.strain_comps, ld, i, ep);
#pragma unroll
            for (int c = 0; c < 6; ++c) em[c] -= ep[c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c) xp[c] = (c >= 7 && c < 7 + NZ) ? 1.0 : 0.0;
        em[0] = 1e-3;
    }
#pragma unroll
    for (int c = 0; c < N; ++c) x[c] = xp[c];
    if (live && A.b.xi_init) {
#pragma unroll
        for (int c = 0; c < N; ++c) x[c] = __ldg(A.b.xi_init + c * ld + i);
    }
    Pt pt;
    double C[N];
    const NewtonResult nr = local_newton<Pt, N>(m, A.nw, pt, x, xp, em, live, C);
    if (!live) return;
    if (A.b.iters) A.b.iters[i] = nr.iters;
    if (A.b.flags) A.b.flags[i] = nr.flag_entry | ((pt.plastic ? 1 : 0) << 1);
    if (A.b.cnorm) A.b.cnorm[i] = nr.cnorm;
    if (A.b.xi) {
#pragma unroll
        for (int c = 0; c < N; ++c) st(A.b.xi, c, ld, i, x[c]);
    }
    if (A.b.C) {
#pragma unroll
        for (int c = 0; c < N; ++c) st(A.b.C, c, ld, i, C[c]);
    }
    auto to_global = [&](const double (&sm)[6], double (&sg)[6]) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(pt.S[a][c], sm[c], s);
            sg[a] = s;
        }
    };
    if (A.b.sigma) {
        double sm[6], sg[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sm[a] = x[a];
        to_global(sm, sg);
#pragma unroll
        for (int a = 0; a < 6; ++a) st(A.b.sigma, a, ld, i, sg[a]);
    }
    const double dg = x[6] - xp[6];
    const bool pl = pt.plastic;
    RegLU<N> lu;
    pt.jacobian(m, dg, lu.a);
    if (A.b.dC_dxi) {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) st(A.b.dC_dxi, r * N + c, ld, i, lu.a[r][c]);
    }
    if (A.b.dC_dxi_prev) {
        // sigma_prev / alpha_prev as in the FULL_3D form; the stretches enter as z - z_prev (the columns
        // are minus the current ones), the delta strains have no previous value
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double v = 0.0;
                if (c < 6) v = (r == c) ? -m.inv_two_mu : 0.0;
                else if (c == 6) v = (r < 6) ? (pl ? -pt.b.n[r] : 0.0) : (r == 6 ? (pl ? 0.0 : -1.0) : -lu.a[r][6]);
                else if (c < 7 + NZ) v = -lu.a[r][c];
                st(A.b.dC_dxi_prev, r * N + c, ld, i, v);
            }
    }
    if (A.b.dC_dp && A.n_active > 0) {
        const int na = A.n_active;
        double x7[7], xp7[7], w[6];
#pragma unroll
        for (int c = 0; c < 7; ++c) { x7[c] = x[c]; xp7[c] = xp[c]; }
        double trw = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) { w[a] = pl ? fma(-dg, pt.b.n[a], pt.dem[a]) : pt.dem[a]; if (is_diag(a)) trw += w[a]; }
        for (int c = 0; c < na; ++c) {
            const int pid = A.pid[c];
            double col[7];
            rate_dC_dp_column<YK>(m, pid, pt.b, x7, xp7, pt.dem, col);
#pragma unroll
            for (int r = 0; r < 7; ++r) st(A.b.dC_dp, (int64_t)r * na + c, ld, i, col[r]);
            // constraint rows: the elastic constants through lam / 2mu, the yield-surface leaves through n
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
            for (int r = 0; r < NR; ++r) {
                const int g = Pt::ccomp(r);
                double v = dlr * trw * (pt.S[g][0] + pt.S[g][3] + pt.S[g][5]);
                if (has_dn) {
                    double q[6];
                    pt.qrow(m, r, q);
#pragma unroll
                    for (int a = 0; a < 6; ++a) v = fma(-dg * q[a], dn[a], v);
                }
                st(A.b.dC_dp, (int64_t)(7 + r) * na + c, ld, i, v);
            }
        }
    }
    if (!A.b.dsig_deps && !A.b.dxi_deps) return;
    bool trouble = lu.factor_natural();
    const bool slow = __any_sync(__activemask(), trouble);
    if (slow && trouble) { pt.jacobian(m, dg, lu.a); lu.factor_pivot(); }
#pragma unroll
    for (int bb = 0; bb < NS; ++bb) {
        double col[N];
        pt.dC_deps(m, scomp[bb], col);
        if (slow && trouble) lu.solve_pivot(col); else lu.solve_natural(col);
        double dxs[6], dsg[6];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            if (A.b.dxi_deps) st(A.b.dxi_deps, r * NS + bb, ld, i, -col[r]);
            if (r < 6) dxs[r] = -col[r];
        }
        if (A.b.dsig_deps) {
            to_global(dxs, dsg);
#pragma unroll
            for (int a = 0; a < 6; ++a) st(A.b.dsig_deps, a * NS + bb, ld, i, dsg[a]);
        }
    }
}

template <int DT>
cudaError_t launch_dt(const MpArgs& A, cudaStream_t stream) {
    const unsigned nblk = (unsigned)((A.b.n + MP_BLOCK - 1) / MP_BLOCK);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: mp_update_rate_dt_kernel<CMADX_YIELD_J2, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HILL: mp_update_rate_dt_kernel<CMADX_YIELD_HILL, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HOSFORD: mp_update_rate_dt_kernel<CMADX_YIELD_HOSFORD, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_mp_update_rate_dt(const MpArgs& A, cudaStream_t stream) {
    if (A.b.def_type == CMADX_DEF_PLANE_STRESS) return launch_dt<CMADX_DEF_PLANE_STRESS>(A, stream);
    return launch_dt<CMADX_DEF_UNIAXIAL_STRESS>(A, stream);
}

}  // namespace cmadx
