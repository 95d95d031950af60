// K1 for the PLANE_STRESS and UNIAXIAL_STRESS deformation types of SmallElasticPlastic
// (cmad/models/small_elastic_plastic.py:126-180, 274-302; cmad/models/kinematics.py:10-57):
// the local system gains the out-of-plane stretch F33 (1 unknown, residual cauchy_33 / 2mu)
// or the two off-axis stretches (2 unknowns, residual off-axis normal stresses / 2mu), i.e.
// n_xi = 8 or 9, and the prescribed kinematics shrink to the in-plane 2x2 (plane stress) or
// the axial 1x1 (uniaxial stress) part of grad_u.  Used by the material-point calibration
// path (KA5: tests/objectives/test_J2_fd_checks.py:266-349).
//
// Same machinery as the FULL_3D generic kernel: the residual / hand-derived Jacobian of
// SepPoint<YK> bordered by the stretch columns and the stress rows, the register LU with
// threshold pivoting (N = 8 / 9), the reference-identical Newton state machine
// (local_newton), IFT outputs by one solve per prescribed strain component.
// Rotated material axes: SepPointDTRot (sep_point_dt.cuh) - the constraints live in global axes,
// the update in material axes; the identity path keeps its own point type (same bits as before).
// Restriction (CMADX_EUNSUPPORTED otherwise): uniaxial_stress_idx 0.
//
// State x = [ep(6), alpha, z...], z = stretches (initialised to 1 by the model).  Total
// strain in material (= global) axes, packed xx,xy,xz,yy,yz,zz:
//   plane stress   : e = [e_xx, e_xy, 0, e_yy, 0, z0 - 1]
//   uniaxial stress: e = [e_xx, ep_xy, ep_xz, z0 - 1, ep_yz, z1 - 1]   (off-diagonal total
//                    strain := plastic strain, so the elastic shear strain vanishes,
//                    small_elastic_plastic.py:46-60 with Q = I)
#include "mp_outputs.cuh"
#include "sep_point_dt.cuh"

namespace cmadx {
namespace {

template <int YK, int DT, bool ROT = false>
__global__ void __launch_bounds__(MP_BLOCK) mp_update_dt_kernel(const __grid_constant__ MpArgs A) {
    using Pt = typename std::conditional<ROT, SepPointDTRot<YK, DT>, SepPointDT<YK, DT>>::type;
    constexpr int N = Pt::N, NZ = Pt::NZ;
    constexpr int NS = (DT == CMADX_DEF_PLANE_STRESS) ? 3 : 1;           // prescribed symmetric components
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
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c) xp[c] = (c < 7) ? 0.0 : 1.0;
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
    double et[6], ee[6], sig[6];
    pt.material_strain(x, em, et);
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = et[a] - x[a];
    const double tre = ee[0] + ee[3] + ee[5];
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], m.lam * tre) : m.two_mu * ee[a];
    if (A.b.sigma) {
        double sg[6];
        pt.to_global(sig, sg);
#pragma unroll
        for (int a = 0; a < 6; ++a) st(A.b.sigma, a, ld, i, sg[a]);
    }
    const double dg = x[6] - xp[6];
    const bool pl = pt.plastic;

    if (A.b.dC_dxi_prev) {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double v = 0.0;
                if (r < 7 && c < 7) {
                    v = (r == c) ? -1.0 : 0.0;
                    if (pl) {
                        if (r == 6) v = 0.0;
                        else if (c == 6) v = pt.b.n[r];
                    }
                }
                st(A.b.dC_dxi_prev, r * N + c, ld, i, v);
            }
    }
    if (A.b.dC_dp && A.n_active > 0) {
        double Mee[6], nee = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double sacc = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) sacc = fma(pt.b.yf.M(a, q), ee[q], sacc);
            Mee[a] = sacc;
            nee = fma(mult(a) * pt.b.n[a], ee[a], nee);
        }
        const int na = A.n_active;
        for (int c = 0; c < na; ++c) {
            const int pid = A.pid[c];
            double col[7];
            dC_dp_column(m, pid, pl, pt.b.yf, pt.b.n, pt.b.f, pt.b.eD, x[6], dg, Mee, nee, sig, col);
#pragma unroll
            for (int r = 0; r < 7; ++r) st(A.b.dC_dp, (int64_t)r * na + c, ld, i, col[r]);
            // stress rows C_r = (lam / 2mu) tr(ee) + ee_c: both branches, elastic parameters only
            double dr = 0.0;
            if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
                const int k = pid - CMADX_P_EL0;
                dr = (m.dlam[k] * m.two_mu - m.lam * 2.0 * m.dmu[k]) * m.inv_two_mu * m.inv_two_mu * tre;
            }
#pragma unroll
            for (int k = 0; k < NZ; ++k) st(A.b.dC_dp, (int64_t)(7 + k) * na + c, ld, i, dr);
        }
    }
    const bool want_ift = A.b.dsig_deps || A.b.dxi_deps;
    if (!want_ift && !A.b.dC_dxi) return;
    RegLU<N> lu;
    pt.jacobian(m, dg, lu.a);
    if (A.b.dC_dxi) {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) st(A.b.dC_dxi, r * N + c, ld, i, lu.a[r][c]);
    }
    if (!want_ift) return;
    // IFT (nonlinear_solver.py:158-171): dxi/de_b = -A^{-1} dC/de_b for the prescribed components,
    // d sigma/de_b = Cel (d e_total/de_b - d ep/de_b)
    bool trouble = lu.factor_natural();
    const bool slow = __any_sync(__activemask(), trouble);
    if (slow && trouble) {
        pt.jacobian(m, dg, lu.a);
        lu.factor_pivot();
    }
#pragma unroll
    for (int bb = 0; bb < NS; ++bb) {
        const int bc = scomp[bb];
        double col[N];
        pt.dC_deps(m, dg, bc, col);
        if (slow && trouble) lu.solve_pivot(col); else lu.solve_natural(col);
        double dx[N];
#pragma unroll
        for (int r = 0; r < N; ++r) dx[r] = -col[r];
        if (A.b.dxi_deps) {
#pragma unroll
            for (int r = 0; r < N; ++r) st(A.b.dxi_deps, r * NS + bb, ld, i, dx[r]);
        }
        if (A.b.dsig_deps) {
            double de[6];
            pt.dmaterial_strain(bc, dx, de);
            double dee[6], dsm[6], dsg[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) dee[a] = de[a] - dx[a];
            const double ltr = m.lam * (dee[0] + dee[3] + dee[5]);
#pragma unroll
            for (int a = 0; a < 6; ++a) dsm[a] = is_diag(a) ? fma(m.two_mu, dee[a], ltr) : m.two_mu * dee[a];
            pt.to_global(dsm, dsg);
#pragma unroll
            for (int a = 0; a < 6; ++a) st(A.b.dsig_deps, a * NS + bb, ld, i, dsg[a]);
        }
    }
}

template <int DT>
cudaError_t launch_dt(const MpArgs& A, cudaStream_t stream) {
    const unsigned nblk = (unsigned)((A.b.n + MP_BLOCK - 1) / MP_BLOCK);
    if (A.m.rot) {
        switch (A.m.yield) {
        case CMADX_YIELD_J2: mp_update_dt_kernel<CMADX_YIELD_J2, DT, true><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
        case CMADX_YIELD_HILL: mp_update_dt_kernel<CMADX_YIELD_HILL, DT, true><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
        case CMADX_YIELD_HOSFORD: mp_update_dt_kernel<CMADX_YIELD_HOSFORD, DT, true><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
        case CMADX_YIELD_BARLAT: mp_update_dt_kernel<CMADX_YIELD_BARLAT, DT, true><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    switch (A.m.yield) {
    case CMADX_YIELD_J2: mp_update_dt_kernel<CMADX_YIELD_J2, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HILL: mp_update_dt_kernel<CMADX_YIELD_HILL, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HOSFORD: mp_update_dt_kernel<CMADX_YIELD_HOSFORD, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_BARLAT: mp_update_dt_kernel<CMADX_YIELD_BARLAT, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_mp_update_dt(const MpArgs& A, cudaStream_t stream) {
    if (A.b.def_type == CMADX_DEF_PLANE_STRESS) return launch_dt<CMADX_DEF_PLANE_STRESS>(A, stream);
    return launch_dt<CMADX_DEF_UNIAXIAL_STRESS>(A, stream);
}

}  // namespace cmadx
