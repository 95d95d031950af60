// SmallElasticPlastic with the PLANE_STRESS / UNIAXIAL_STRESS deformation types
// (cmad/models/small_elastic_plastic.py:126-180, 274-302; kinematics.py:10-57): the FULL_3D
// point (SepPoint<YK>) bordered by the stretch unknowns and the stress-constraint rows.
// Identity material axes, uniaxial_stress_idx = 0.
//
// State x = [ep(6), alpha, z...], z = stretches (initialised to 1 by the model).  Total
// strain in material (= global) axes, packed xx,xy,xz,yy,yz,zz:
//   plane stress   : e = [e_xx, e_xy, 0, e_yy, 0, z0 - 1]
//   uniaxial stress: e = [e_xx, ep_xy, ep_xz, z0 - 1, ep_yz, z1 - 1]   (off-diagonal total
//                    strain := plastic strain, so the elastic shear strain vanishes,
//                    small_elastic_plastic.py:46-60 with Q = I)
#pragma once
#include "point_solver.cuh"

namespace cmadx {
namespace {

template <int YK, int DT>
struct SepPointDT {
    static constexpr int NZ = (DT == CMADX_DEF_PLANE_STRESS) ? 1 : 2;
    static constexpr int N = 7 + NZ, ALPHA = 6;
    SepPoint<YK> b;
    bool plastic;

    // strain component driven by stretch k
    CMADX_DEV static constexpr int zcomp(int k) { return (DT == CMADX_DEF_PLANE_STRESS) ? 5 : (k == 0 ? 3 : 5); }

    CMADX_DEV void total_strain(const double (&x)[N], const double (&em)[6], double (&et)[6]) const {
        if (DT == CMADX_DEF_PLANE_STRESS) {
            et[0] = em[0]; et[1] = em[1]; et[2] = 0.0; et[3] = em[3]; et[4] = 0.0; et[5] = x[7] - 1.0;
        } else {
            et[0] = em[0]; et[1] = x[1]; et[2] = x[2]; et[3] = x[7] - 1.0; et[4] = x[4]; et[5] = x[8] - 1.0;
        }
    }

    CMADX_DEV void residual(const DevMat& m, const double (&x)[N], const double (&xp)[N],
                            const double (&em)[6], double (&C)[N]) {
        double et[6], x7[7], xp7[7], C7[7];
        total_strain(x, em, et);
#pragma unroll
        for (int c = 0; c < 7; ++c) { x7[c] = x[c]; xp7[c] = xp[c]; }
        b.residual(m, x7, xp7, et, C7);
        plastic = b.plastic;
#pragma unroll
        for (int c = 0; c < 7; ++c) C[c] = C7[c];
        const double tre = (et[0] - x[0]) + (et[3] - x[3]) + (et[5] - x[5]);
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
            const int c = zcomp(k);
            C[7 + k] = fma(m.two_mu, et[c] - x[c], m.lam * tre) * m.inv_two_mu;     // cauchy_cc / 2mu
        }
    }

    CMADX_DEV void jacobian(const DevMat& m, double dg, double (&J)[N][N]) const {
        double J7[7][7];
        b.jacobian(m, dg, J7);
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) J[r][c] = J7[r][c];
        if (DT == CMADX_DEF_UNIAXIAL_STRESS && plastic) {
            // the stress does not depend on the plastic shear strain (elastic shear strain = 0)
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                J[r][1] = (r == 1) ? 1.0 : 0.0; J[r][2] = (r == 2) ? 1.0 : 0.0; J[r][4] = (r == 4) ? 1.0 : 0.0;
            }
        }
        const double s = dg * m.two_mu;
        const double lr = m.lam * m.inv_two_mu;
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
            const int c = zcomp(k);
            // d C[0..6] / d z_k = d C / d e_c
#pragma unroll
            for (int a = 0; a < 6; ++a) J[a][7 + k] = plastic ? -s * b.yf.M(a, c) : 0.0;
            J[6][7 + k] = plastic ? b.n[c] : 0.0;
            // stress rows: C_r = (lam tr(ee) + 2mu ee_c) / 2mu
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double v = is_diag(a) ? -lr : 0.0;
                if (a == c) v -= 1.0;
                if (DT == CMADX_DEF_UNIAXIAL_STRESS && !is_diag(a)) v = 0.0;
                J[7 + k][a] = v;
            }
            J[7 + k][6] = 0.0;
#pragma unroll
            for (int k2 = 0; k2 < NZ; ++k2) J[7 + k][7 + k2] = lr + ((k2 == k) ? 1.0 : 0.0);
        }
    }

    // dC / d e_b for a PRESCRIBED strain component b (symmetric component, both entries moving)
    CMADX_DEV void dC_deps(const DevMat& m, double dg, int bcomp, double (&col)[N]) const {
        const double s = dg * m.two_mu;
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = plastic ? -s * b.yf.M(a, bcomp) : 0.0;
        col[6] = plastic ? mult(bcomp) * b.n[bcomp] : 0.0;
#pragma unroll
        for (int k = 0; k < NZ; ++k) col[7 + k] = is_diag(bcomp) ? m.lam * m.inv_two_mu : 0.0;
    }

    // ---- what the output / sensitivity stages need besides the residual (shared interface with
    //      SepPointDTRot): material total strain, global stress, strain variation, QoI cotangent
    CMADX_DEV void material_strain(const double (&x)[N], const double (&em)[6], double (&et)[6]) const {
        total_strain(x, em, et);
    }
    CMADX_DEV void to_global(const double (&sm)[6], double (&sg)[6]) const {
#pragma unroll
        for (int a = 0; a < 6; ++a) sg[a] = sm[a];
    }
    // d(material total strain) for a unit prescribed component bc (bc < 0: none) and a state change dx
    CMADX_DEV void dmaterial_strain(int bc, const double (&dx)[N], double (&de)[6]) const {
#pragma unroll
        for (int a = 0; a < 6; ++a) de[a] = (a == bc) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < NZ; ++k) de[zcomp(k)] += dx[7 + k];
        if (DT == CMADX_DEF_UNIAXIAL_STRESS) { de[1] = dx[1]; de[2] = dx[2]; de[4] = dx[4]; }
    }
    // r = dJ/d(global stress, packed) -> material-frame cotangent rm and dJ/dx through the stress
    CMADX_DEV void stress_cotangent(const DevMat& m, const double (&r)[6], double (&rm)[6], double (&dJdx)[N]) const {
#pragma unroll
        for (int a = 0; a < 6; ++a) rm[a] = r[a];
        const double rtr = r[0] + r[3] + r[5];
#pragma unroll
        for (int bq = 0; bq < 6; ++bq) {
            double v = is_diag(bq) ? fma(-m.two_mu, r[bq], -m.lam * rtr) : -m.two_mu * r[bq];
            if (DT == CMADX_DEF_UNIAXIAL_STRESS && !is_diag(bq)) v = 0.0;   // sigma independent of ep_shear
            dJdx[bq] = v;
        }
        dJdx[6] = 0.0;
#pragma unroll
        for (int k = 0; k < NZ; ++k) dJdx[7 + k] = fma(m.two_mu, r[zcomp(k)], m.lam * rtr);
    }
};

// The same point under ROTATED MATERIAL AXES (cmad/models/small_elastic_plastic.py:44-62, 287-302):
// the kinematic constraints live in GLOBAL axes, the constitutive update in material axes.
//   global total strain  plane stress   : eg = [e_xx, e_xy, 0, e_yy, 0, z0 - 1]
//                        uniaxial stress: eg = [e_xx, og_xy, og_xz, z0 - 1, og_yz, z1 - 1],
//                                         og = off-diagonal part of Q ep Q^T (= S ep)
//   material total strain em = T eg (rot_maps), stress rows = (S sigma_m)[zz | yy, zz] / 2 mu.
// With G = d em / d x (6 x N: T O S on the plastic-strain columns for uniaxial stress, T[:, z] on the
// stretch columns) the Jacobian is the FULL_3D one chained through G:
//   dC7/dx = J7 + (E - J7[:, :6]) G (plastic; E = [I6; 0]),   stress rows = w_k^T (G - [I6 0]),
//   w_k = S[z_k, :] Cel / 2 mu.   Reduces to SepPointDT for Q = I (that type stays the identity path:
// same bits as before).
template <int YK, int DT>
struct SepPointDTRot {
    static constexpr int NZ = (DT == CMADX_DEF_PLANE_STRESS) ? 1 : 2;
    static constexpr int N = 7 + NZ, ALPHA = 6;
    SepPoint<YK> b;
    bool plastic;
    double T[6][6], S[6][6];

    CMADX_DEV static constexpr int zcomp(int k) { return (DT == CMADX_DEF_PLANE_STRESS) ? 5 : (k == 0 ? 3 : 5); }

    CMADX_DEV void material_strain(const double (&x)[N], const double (&em)[6], double (&et)[6]) const {
        double eg[6];
        if (DT == CMADX_DEF_PLANE_STRESS) {
            eg[0] = em[0]; eg[1] = em[1]; eg[2] = 0.0; eg[3] = em[3]; eg[4] = 0.0; eg[5] = x[7] - 1.0;
        } else {
            double og[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) s = fma(S[a][c], x[c], s);
                og[a] = s;
            }
            eg[0] = em[0]; eg[1] = og[1]; eg[2] = og[2]; eg[3] = x[7] - 1.0; eg[4] = og[4]; eg[5] = x[8] - 1.0;
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(T[a][c], eg[c], s);
            et[a] = s;
        }
    }
    CMADX_DEV void to_global(const double (&sm)[6], double (&sg)[6]) const {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(S[a][c], sm[c], s);
            sg[a] = s;
        }
    }
    // G = d(material total strain)/dx
    CMADX_DEV void strain_map(double (&G)[6][N]) const {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                double v = 0.0;
                if (DT == CMADX_DEF_UNIAXIAL_STRESS && c < 6)
                    v = T[a][1] * S[1][c] + T[a][2] * S[2][c] + T[a][4] * S[4][c];
                G[a][c] = v;
            }
#pragma unroll
            for (int k = 0; k < NZ; ++k) G[a][7 + k] = T[a][zcomp(k)];
        }
    }
    // w_k = S[z_k, :] Cel / 2 mu
    CMADX_DEV void stress_row(const DevMat& m, int k, double (&w)[6]) const {
        const int z = zcomp(k);
        const double lr = m.lam * m.inv_two_mu * (S[z][0] + S[z][3] + S[z][5]);
#pragma unroll
        for (int c = 0; c < 6; ++c) w[c] = S[z][c] + (is_diag(c) ? lr : 0.0);
    }

    CMADX_DEV void residual(const DevMat& m, const double (&x)[N], const double (&xp)[N],
                            const double (&em)[6], double (&C)[N]) {
        rot_maps(m.Q, T, S);
        double et[6], x7[7], xp7[7], C7[7];
        material_strain(x, em, et);
#pragma unroll
        for (int c = 0; c < 7; ++c) { x7[c] = x[c]; xp7[c] = xp[c]; }
        b.residual(m, x7, xp7, et, C7);
        plastic = b.plastic;
#pragma unroll
        for (int c = 0; c < 7; ++c) C[c] = C7[c];
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
            double w[6], s = 0.0;
            stress_row(m, k, w);
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(w[c], et[c] - x[c], s);
            C[7 + k] = s;                                                       // (Q sigma_m Q^T)_zz / 2mu
        }
    }

    CMADX_DEV void strain_sensitivity(const DevMat& m, double dg, double (&dCe)[7][6]) const {
        // d C7 / d(material strain component): -(J7[:, :6] - E) on the plastic branch, 0 on the elastic one
        const double s = dg * m.two_mu;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
#pragma unroll
            for (int a = 0; a < 6; ++a) dCe[a][q] = plastic ? -s * b.yf.M(a, q) : 0.0;
            dCe[6][q] = plastic ? mult(q) * b.n[q] : 0.0;
        }
    }

    CMADX_DEV void jacobian(const DevMat& m, double dg, double (&J)[N][N]) const {
        double J7[7][7], G[6][N], dCe[7][6];
        b.jacobian(m, dg, J7);
        strain_map(G);
        strain_sensitivity(m, dg, dCe);
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double v = (c < 7) ? J7[r][c < 7 ? c : 0] : 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) v = fma(dCe[r][q], G[q][c], v);
                J[r][c] = v;
            }
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
            double w[6];
            stress_row(m, k, w);
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double v = (c < 6) ? -w[c < 6 ? c : 0] : 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) v = fma(w[q], G[q][c], v);
                J[7 + k][c] = v;
            }
        }
    }

    CMADX_DEV void dC_deps(const DevMat& m, double dg, int bcomp, double (&col)[N]) const {
        double dCe[7][6];
        strain_sensitivity(m, dg, dCe);
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) v = fma(dCe[r][q], T[q][bcomp], v);
            col[r] = v;
        }
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
            double w[6], v = 0.0;
            stress_row(m, k, w);
#pragma unroll
            for (int q = 0; q < 6; ++q) v = fma(w[q], T[q][bcomp], v);
            col[7 + k] = v;
        }
    }

    CMADX_DEV void dmaterial_strain(int bc, const double (&dx)[N], double (&de)[6]) const {
        double G[6][N];
        strain_map(G);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double v = (bc >= 0) ? T[a][bc >= 0 ? bc : 0] : 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) v = fma(G[a][c], dx[c], v);
            de[a] = v;
        }
    }

    CMADX_DEV void stress_cotangent(const DevMat& m, const double (&r)[6], double (&rm)[6], double (&dJdx)[N]) const {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) s = fma(S[a][c], r[a], s);
            rm[c] = s;
        }
        const double rtr = rm[0] + rm[3] + rm[5];
        double qv[6], G[6][N];
#pragma unroll
        for (int q = 0; q < 6; ++q) qv[q] = is_diag(q) ? fma(m.two_mu, rm[q], m.lam * rtr) : m.two_mu * rm[q];
        strain_map(G);
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double v = (c < 6) ? -qv[c < 6 ? c : 0] : 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) v = fma(qv[q], G[q][c], v);
            dJdx[c] = v;
        }
    }
};

// prescribed symmetric strain of point i from the `strain` rows of a def-type batch
template <int DT>
CMADX_DEV void load_dt_strain(const double* strain, int comps, int64_t ld, int64_t i, double (&em)[6]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) em[c] = 0.0;
    if (DT == CMADX_DEF_PLANE_STRESS) {
        if (comps == 3) {
            em[0] = __ldg(strain + i); em[1] = __ldg(strain + ld + i); em[3] = __ldg(strain + 2 * ld + i);
        } else {                                                      // 2x2 grad_u, row-major
            em[0] = __ldg(strain + i); em[3] = __ldg(strain + 3 * ld + i);
            em[1] = 0.5 * (__ldg(strain + ld + i) + __ldg(strain + 2 * ld + i));
        }
    } else {
        em[0] = __ldg(strain + i);
    }
}

}  // namespace
}  // namespace cmadx
