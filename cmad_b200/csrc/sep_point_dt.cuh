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
