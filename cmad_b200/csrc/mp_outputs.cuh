// Output helpers shared by the material-point kernels (generic 7x7 Newton and
// the J2 radial-return specialisation).
#pragma once
#include "mp_update.cuh"

namespace cmadx {

CMADX_DEV void st(double* p, int64_t c, int64_t ld, int64_t i, double v) {
    __stcs(p + c * ld + i, v);
}

// xi_prev (7) and the symmetric strain (6) of point i
CMADX_DEV void load_point(const cmadx_mp_buffers_t& b, int64_t i, bool live,
                          double (&xp)[7], double (&e)[6]) {
    const int64_t ld = b.ld;
    if (live) {
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + c * ld + i);
        if (b.strain_comps == 6) {
#pragma unroll
            for (int c = 0; c < 6; ++c) e[c] = __ldg(b.strain + c * ld + i);
        } else {
            double g[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) g[c] = __ldg(b.strain + c * ld + i);
            e[0] = g[0]; e[3] = g[4]; e[5] = g[8];
            e[1] = 0.5 * (g[1] + g[3]); e[2] = 0.5 * (g[2] + g[6]); e[4] = 0.5 * (g[5] + g[7]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = 0.0;
    }
}

// dC/dxi_prev: plastic rows a<6: -I and +n in the alpha column, yield row 0;
// elastic: -I
CMADX_DEV void write_dC_dxi_prev(double* out, int64_t ld, int64_t i, bool pl, const double (&n)[6]) {
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            double v = (r == c) ? -1.0 : 0.0;
            if (pl) {
                if (r == 6) v = 0.0;
                else if (c == 6) v = n[r];
            }
            st(out, r * 7 + c, ld, i, v);
        }
}

// one column of dC/dp at (x*, x_prev) for canonical parameter id `pid`.
// Elastic branch -> 0 (C_e holds no parameters).  Mee = (dn/dsigma : ee),
// nee = n : ee, sig = material cauchy (for the yield-surface parameters).
template <class YF>
CMADX_DEV void dC_dp_column(const DevMat& m, int pid, bool pl, const YF& yf, const double (&n)[6],
                            double f, double eD, double alpha, double dg,
                            const double (&Mee)[6], double nee, const double (&sig)[6],
                            double (&col)[7]) {
#pragma unroll
    for (int r = 0; r < 7; ++r) col[r] = 0.0;
    if (!pl) return;
    if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
        // only mu matters: all three surfaces are pressure-insensitive
        const double dmu = m.dmu[pid - CMADX_P_EL0];
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = -2.0 * dg * Mee[a] * dmu;
        col[6] = (nee - f) / m.mu * dmu;
    } else if (pid == CMADX_P_Y) {
        col[6] = -m.inv_two_mu;
    } else if (pid == CMADX_P_VOCE_S) {
        col[6] = -(1.0 - eD) * m.inv_two_mu;
    } else if (pid == CMADX_P_VOCE_D) {
        col[6] = -m.S * alpha * eD * m.inv_two_mu;
    } else if (pid == CMADX_P_LIN_K) {
        col[6] = -alpha * m.inv_two_mu;
    } else {
        double dphi, dn[6];
        if (yf.dparam(m, pid, sig, dphi, dn)) {
#pragma unroll
            for (int a = 0; a < 6; ++a) col[a] = -dg * dn[a];
            col[6] = dphi * m.inv_two_mu;
        }
    }
}

template <class YF>
CMADX_DEV void write_dC_dp(const MpArgs& A, int64_t i, bool pl, const YF& yf, const double (&n)[6],
                           double f, double eD, double alpha, double dg,
                           const double (&Mee)[6], double nee, const double (&sig)[6]) {
    const int na = A.n_active;
    for (int c = 0; c < na; ++c) {
        double col[7];
        dC_dp_column(A.m, A.pid[c], pl, yf, n, f, eD, alpha, dg, Mee, nee, sig, col);
#pragma unroll
        for (int r = 0; r < 7; ++r) st(A.b.dC_dp, (int64_t)r * na + c, A.b.ld, i, col[r]);
    }
}

}  // namespace cmadx
