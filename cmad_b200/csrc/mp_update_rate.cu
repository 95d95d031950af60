// K1 for SmallRateElasticPlastic (cmad/models/small_rate_elastic_plastic.py:34-100, 250-346),
// FULL_3D, identity material axes: the rate form of the small-strain model.  State
// x = [cauchy(6), alpha]; with the strain INCREMENT de = eps - eps_prev (the `strain` rows of
// the batch carry the increment: the reference forms it from U and U_prev, :41-51):
//   trial   = Cel de
//   C_e     = [ (sig - sig_prev - trial) / 2mu ,  alpha - alpha_prev ]
//   C_p     = [ (sig - sig_prev - trial + Cel (dgamma n(sig))) / 2mu ,  f(sig, alpha) ]
//   C       = C_p if (f > tol or |f| < tol) else C_e                 (paths.py:26-27)
// Hand-derived Jacobian (plastic): d C_a/d sig_b = (delta_ab + dgamma 2mu M_ab) / 2mu
// (Cel M = 2mu M: the surfaces are pressure-insensitive), d C_a/d alpha = n_a (+ lam tr n / 2mu),
// d f/d sig_b = w_b n_b / 2mu, d f/d alpha = -H'/2mu.  Same Newton state machine, register
// LU and output conventions as the other K1 kernels; `sigma` = the stress part of the state.
// Note: the first evaluation at sig = 0 has an undefined J2 normal (0/0) exactly like the
// reference's; it only ever enters the discarded branch of the select.
#include "mp_outputs.cuh"

namespace cmadx {
namespace {

template <int YK>
struct RatePoint {
    static constexpr int N = 7, ALPHA = 6;
    YieldFn<YK> yf;
    double n[6];
    double f, eD;
    bool plastic;

    CMADX_DEV static double hard_slope(const DevMat& m, double eD_) {
        double Hp = 0.0;
        if (m.hmask & CMADX_HARD_VOCE) Hp = m.S * m.D * eD_;
        if (m.hmask & CMADX_HARD_LINEAR) Hp += m.K;
        return Hp;
    }

    CMADX_DEV void residual(const DevMat& m, const double (&x)[7], const double (&xp)[7],
                            const double (&de)[6], double (&C)[7]) {
        double sig[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = x[a];
        double phi;
        yf.eval(m, sig, phi, n);
        double Hd = 0.0;
        eD = 0.0;
        if (m.hmask & CMADX_HARD_VOCE) { eD = exp(-m.D * x[6]); Hd = m.S * (1.0 - eD); }
        if (m.hmask & CMADX_HARD_LINEAR) Hd = fma(m.K, x[6], Hd);
        f = (phi - (m.Y + Hd)) * m.inv_two_mu;
        plastic = (f > m.yield_tol) || (fabs(f) < m.yield_tol);
        const double dg = x[6] - xp[6];
        const double ltr = m.lam * (de[0] + de[3] + de[5]);
        const double lntr = m.lam * (dg * n[0] + dg * n[3] + dg * n[5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double trial = is_diag(a) ? fma(m.two_mu, de[a], ltr) : m.two_mu * de[a];
            const double ce = x[a] - xp[a] - trial;
            const double pl_ = is_diag(a) ? fma(m.two_mu, dg * n[a], lntr) : m.two_mu * (dg * n[a]);
            C[a] = (plastic ? ce + pl_ : ce) * m.inv_two_mu;
        }
        C[6] = plastic ? f : dg;
    }

    CMADX_DEV void jacobian(const DevMat& m, double dg, double (&J)[7][7]) const {
        if (plastic) {
            const double s = dg * m.two_mu;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = 0; b < 6; ++b) J[a][b] = fma(s, yf.M(a, b), (a == b) ? 1.0 : 0.0) * m.inv_two_mu;
                J[a][6] = n[a];
                J[6][a] = mult(a) * n[a] * m.inv_two_mu;
            }
            J[6][6] = -hard_slope(m, eD) * m.inv_two_mu;
        } else {
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int b = 0; b < 7; ++b) J[a][b] = (a == b) ? ((a < 6) ? m.inv_two_mu : 1.0) : 0.0;
        }
    }
};

template <int YK>
__global__ void __launch_bounds__(MP_BLOCK) mp_update_rate_kernel(const __grid_constant__ MpArgs A) {
    using Pt = RatePoint<YK>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.b.n;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    double xp[7], x[7], de[6];
    load_point(A.b, i, live, xp, de);
    if (!live) de[0] = 1e-3;
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
        for (int a = 0; a < 6; ++a) st(A.b.sigma, a, ld, i, x[a]);
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
        // C_a = (sig_a - sigp_a) / 2mu - [w_a + (lam / 2mu) tr(w) delta_a],  w = de - dgamma n (plastic) or de
        double w[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) w[a] = pl ? fma(-dg, pt.n[a], de[a]) : de[a];
        const double trw = w[0] + w[3] + w[5];
        const int na = A.n_active;
        double sig[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = x[a];
        for (int c = 0; c < na; ++c) {
            const int pid = A.pid[c];
            double col[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
                const int k = pid - CMADX_P_EL0;
                const double dinv = -2.0 * m.dmu[k] * m.inv_two_mu * m.inv_two_mu;          // d(1/2mu)
                const double dlr = (m.dlam[k] * m.two_mu - m.lam * 2.0 * m.dmu[k]) * m.inv_two_mu * m.inv_two_mu;
#pragma unroll
                for (int a = 0; a < 6; ++a) col[a] = (x[a] - xp[a]) * dinv - (is_diag(a) ? dlr * trw : 0.0);
                if (pl) col[6] = pt.f * m.two_mu * dinv;
            } else if (pl) {
                if (pid == CMADX_P_Y) col[6] = -m.inv_two_mu;
                else if (pid == CMADX_P_VOCE_S) col[6] = -(1.0 - pt.eD) * m.inv_two_mu;
                else if (pid == CMADX_P_VOCE_D) col[6] = -m.S * x[6] * pt.eD * m.inv_two_mu;
                else if (pid == CMADX_P_LIN_K) col[6] = -x[6] * m.inv_two_mu;
                else {
                    double dphi, dn[6];
                    if (pt.yf.dparam(m, pid, sig, dphi, dn)) {
                        const double trn = dn[0] + dn[3] + dn[5];
#pragma unroll
                        for (int a = 0; a < 6; ++a) col[a] = dg * (dn[a] + (is_diag(a) ? lr * trn : 0.0));
                        col[6] = dphi * m.inv_two_mu;
                    }
                }
            }
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
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        // dC/d(de_b) = -(delta_ab + (lam/2mu) [a diag][b diag]) on the stress rows, both branches
        double col[7];
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = ((a == b) ? 1.0 : 0.0) + ((is_diag(a) && is_diag(b)) ? lr : 0.0);
        col[6] = 0.0;
        if (slow && trouble) lu.solve_pivot(col); else lu.solve_natural(col);     // = -dxi/d(de_b)... sign folded
        if (A.b.dxi_deps) {
#pragma unroll
            for (int r = 0; r < 7; ++r) st(A.b.dxi_deps, r * 6 + b, ld, i, col[r]);
        }
        if (A.b.dsig_deps) {
#pragma unroll
            for (int a = 0; a < 6; ++a) st(A.b.dsig_deps, a * 6 + b, ld, i, col[a]);
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
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace cmadx
