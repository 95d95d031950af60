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
#include "rate_point_dt.cuh"

namespace cmadx {
namespace {

template <int YK, int DT>
__global__ void __launch_bounds__(MP_BLOCK) mp_update_rate_dt_kernel(const __grid_constant__ MpArgs A) {
    using Pt = RatePointDT<YK, DT>;
    constexpr int N = Pt::N, NZ = Pt::NZ;
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
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) st(A.b.dC_dxi_prev, r * N + c, ld, i, rate_dt_B(m, pt, lu.a, r, c));
    }
    if (A.b.dC_dp && A.n_active > 0) {
        const int na = A.n_active;
        for (int c = 0; c < na; ++c) {
            double col[N];
            rate_dt_dC_dp_column<YK, DT>(m, A.pid[c], pt, x, xp, col);
#pragma unroll
            for (int r = 0; r < N; ++r) st(A.b.dC_dp, (int64_t)r * na + c, ld, i, col[r]);
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
    case CMADX_YIELD_BARLAT: mp_update_rate_dt_kernel<CMADX_YIELD_BARLAT, DT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
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
