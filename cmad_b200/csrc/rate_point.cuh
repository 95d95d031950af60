// Per-point routines of SmallRateElasticPlastic (cmad/models/small_rate_elastic_plastic.py:34-100, 250-346),
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
#pragma once
#include "mp_outputs.cuh"

namespace cmadx {

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


// symmetric strain of one row set of a batch / history (6 = symmetric components, 9 = grad_u)
CMADX_DEV void rate_load_strain(const double* s, int comps, int64_t ld, int64_t i, double (&e)[6]) {
    if (comps == 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = __ldg(s + c * ld + i);
    } else {
        double g[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) g[c] = __ldg(s + c * ld + i);
        e[0] = g[0]; e[3] = g[4]; e[5] = g[8];
        e[1] = 0.5 * (g[1] + g[3]); e[2] = 0.5 * (g[2] + g[6]); e[4] = 0.5 * (g[5] + g[7]);
    }
}

// one column of dC/dp at (x, x_prev, de) for canonical parameter id `pid`; `pt` fresh at x.
// C_a = (sig_a - sigp_a) / 2mu - [w_a + (lam / 2mu) tr(w) delta_a],  w = de - dgamma n (plastic) or de:
// the elastic parameters enter through 1/2mu and lam/2mu in BOTH branches.
template <int YK>
CMADX_DEV void rate_dC_dp_column(const DevMat& m, const int pid, const RatePoint<YK>& pt,
                                 const double (&x)[7], const double (&xp)[7], const double (&de)[6],
                                 double (&col)[7]) {
    const bool pl = pt.plastic;
    const double dg = x[6] - xp[6];
    const double lr = m.lam * m.inv_two_mu;
#pragma unroll
    for (int r = 0; r < 7; ++r) col[r] = 0.0;
    if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
        double trw = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) if (is_diag(a)) trw += pl ? fma(-dg, pt.n[a], de[a]) : de[a];
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
            double sig[6], dphi, dn[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) sig[a] = x[a];
            if (pt.yf.dparam(m, pid, sig, dphi, dn)) {
                const double trn = dn[0] + dn[3] + dn[5];
#pragma unroll
                for (int a = 0; a < 6; ++a) col[a] = dg * (dn[a] + (is_diag(a) ? lr * trn : 0.0));
                col[6] = dphi * m.inv_two_mu;
            }
        }
    }
}

}  // namespace cmadx
