// K1 for SmallRateElasticPlastic: see rate_point.cuh for the model and the references.
#include "rate_point.cuh"

namespace cmadx {
namespace {

template <int YK>
__global__ void __launch_bounds__(MP_BLOCK) mp_update_rate_kernel(const __grid_constant__ MpArgs A) {
    using Pt = RatePoint<YK>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.b.n;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    double xp[7], x[7], de[6];
    load_point(A.b, i, live, xp, de);
    if (live && A.b.strain_prev) {       // total strains given: the increment is formed here (:41-51)
        double ep[6];
        rate_load_strain(A.b.strain_prev, A.b.strain_comps, ld, i, ep);
#pragma unroll
        for (int c = 0; c < 6; ++c) de[c] -= ep[c];
    }
    if (!live) de[0] = 1e-3;
    // rotated material axes (compute_delta_strain :53-72, _cauchy_fn :351-359): the state is the
    // MATERIAL-frame stress; the increment enters as Q^T de Q, the stress leaves as Q sig Q^T
    const bool rot = m.rot != 0;
    double T[6][6], S[6][6];
    if (rot) {
        rot_maps(m.Q, T, S);
        double dm[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s = fma(T[c][b], de[b], s);
            dm[c] = s;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) de[c] = dm[c];
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) x[c] = xp[c];
    if (live && A.b.xi_init) {
#pragma unroll
        for (int c = 0; c < 7; ++c) x[c] = __ldg(A.b.xi_init + c * ld + i);
    }
    Pt pt;
    double C[7];
    const NewtonResult nr = local_newton<Pt, 7>(m, A.nw, pt, x, xp, de, live, C);
    if (!live) return;
    if (A.b.iters) A.b.iters[i] = nr.iters;
    if (A.b.flags) A.b.flags[i] = nr.flag_entry | ((pt.plastic ? 1 : 0) << 1);
    if (A.b.cnorm) A.b.cnorm[i] = nr.cnorm;
    if (A.b.xi) {
#pragma unroll
        for (int c = 0; c < 7; ++c) st(A.b.xi, c, ld, i, x[c]);
    }
    if (A.b.C) {
#pragma unroll
        for (int c = 0; c < 7; ++c) st(A.b.C, c, ld, i, C[c]);
    }
    if (A.b.sigma) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double v = x[a];
            if (rot) {
                v = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) v = fma(S[a][c], x[c], v);
            }
            st(A.b.sigma, a, ld, i, v);
        }
    }
    const double dg = x[6] - xp[6];
    const bool pl = pt.plastic;
    const double lr = m.lam * m.inv_two_mu;
    if (A.b.dC_dxi_prev) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                double v = 0.0;
                if (r < 6) v = (r == c) ? -m.inv_two_mu : ((c == 6 && pl) ? -pt.n[r] : 0.0);
                else v = (c == 6 && !pl) ? -1.0 : 0.0;
                st(A.b.dC_dxi_prev, r * 7 + c, ld, i, v);
            }
    }
    if (A.b.dC_dp && A.n_active > 0) {
        const int na = A.n_active;
        for (int c = 0; c < na; ++c) {
            double col[7];
            rate_dC_dp_column<YK>(m, A.pid[c], pt, x, xp, de, col);
#pragma unroll
            for (int r = 0; r < 7; ++r) st(A.b.dC_dp, (int64_t)r * na + c, ld, i, col[r]);
        }
    }
    const bool want_ift = A.b.dsig_deps || A.b.dxi_deps;
    if (!want_ift && !A.b.dC_dxi) return;
    RegLU<7> lu;
    pt.jacobian(m, dg, lu.a);
    if (A.b.dC_dxi) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) st(A.b.dC_dxi, r * 7 + c, ld, i, lu.a[r][c]);
    }
    if (!want_ift) return;
    const bool trouble = lu.factor_natural();
    const bool slow = __any_sync(__activemask(), trouble);
    if (slow && trouble) { pt.jacobian(m, dg, lu.a); lu.factor_pivot(); }
    double Xm[7][6];         // dxi / d(material strain increment component b)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        // dC/d(de_b) = -(delta_ab + (lam/2mu) [a diag][b diag]) on the stress rows, both branches
        double col[7];
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = ((a == b) ? 1.0 : 0.0) + ((is_diag(a) && is_diag(b)) ? lr : 0.0);
        col[6] = 0.0;
        if (slow && trouble) lu.solve_pivot(col); else lu.solve_natural(col);     // = -dxi/d(de_b)... sign folded
#pragma unroll
        for (int r = 0; r < 7; ++r) Xm[r][b] = col[r];
    }
    if (rot) {               // global strain components: columns through T, stress rows through S
        double Xg[7][6];
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) s = fma(Xm[r][c], T[c][b], s);
                Xg[r][b] = s;
            }
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int b = 0; b < 6; ++b) Xm[r][b] = Xg[r][b];
    }
    if (A.b.dxi_deps) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int b = 0; b < 6; ++b) st(A.b.dxi_deps, r * 6 + b, ld, i, Xm[r][b]);
    }
    if (A.b.dsig_deps) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double v = Xm[a][b];
                if (rot) {
                    v = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) v = fma(S[a][c], Xm[c][b], v);
                }
                st(A.b.dsig_deps, a * 6 + b, ld, i, v);
            }
    }
}

}  // namespace

cudaError_t launch_mp_update_rate(const MpArgs& A, cudaStream_t stream) {
    const unsigned nblk = (unsigned)((A.b.n + MP_BLOCK - 1) / MP_BLOCK);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: mp_update_rate_kernel<CMADX_YIELD_J2><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HILL: mp_update_rate_kernel<CMADX_YIELD_HILL><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HOSFORD: mp_update_rate_kernel<CMADX_YIELD_HOSFORD><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_BARLAT: mp_update_rate_kernel<CMADX_YIELD_BARLAT><<<nblk, MP_BLOCK, 0, stream>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace cmadx
