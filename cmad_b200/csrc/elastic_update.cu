// K1e - batched update for the `Elastic` model (cmad/models/elastic.py:29-195):
// state xi = cauchy(6); the local residual is linear in xi so the reference's
// Newton lands in one update.  Kept behind the same entry point as the
// elastic-plastic kernel so COUPLED == CLOSED_FORM parity (KA4) can be checked
// through the same path.
#include "mp_update.cuh"

namespace cmadx {
namespace {

CMADX_DEV void st(double* p, int64_t c, int64_t ld, int64_t i, double v) {
    __stcs(p + c * ld + i, v);
}

__global__ void __launch_bounds__(MP_BLOCK)
elastic_update_kernel(const __grid_constant__ MpArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.b.n;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;
    double xp[6], x[6], e[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { xp[c] = 0.0; e[c] = 0.0; }
    if (live) {
#pragma unroll
        for (int c = 0; c < 6; ++c) xp[c] = __ldg(A.b.xi_prev + c * ld + i);
        if (A.b.strain_comps == 6) {
#pragma unroll
            for (int c = 0; c < 6; ++c) e[c] = __ldg(A.b.strain + c * ld + i);
        } else {
            double g[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) g[c] = __ldg(A.b.strain + c * ld + i);
            e[0] = g[0]; e[3] = g[4]; e[5] = g[8];
            e[1] = 0.5 * (g[1] + g[3]); e[2] = 0.5 * (g[2] + g[6]); e[4] = 0.5 * (g[5] + g[7]);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) x[c] = xp[c];
    if (live && A.b.xi_init) {
#pragma unroll
        for (int c = 0; c < 6; ++c) x[c] = __ldg(A.b.xi_init + c * ld + i);
    }
    ElasticPoint pt;
    double Cres[6];
    const NewtonResult nr = local_newton<ElasticPoint, 6>(m, A.nw, pt, x, xp, e, live, Cres);
    if (!live) return;
    if (A.b.C) {
#pragma unroll
        for (int c = 0; c < 6; ++c) st(A.b.C, c, ld, i, Cres[c]);
    }
    if (A.b.iters) A.b.iters[i] = nr.iters;
    if (A.b.flags) A.b.flags[i] = 0;
    if (A.b.cnorm) A.b.cnorm[i] = nr.cnorm;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        if (A.b.xi) st(A.b.xi, c, ld, i, x[c]);
        if (A.b.sigma) st(A.b.sigma, c, ld, i, x[c]);      // elastic.py:188-195
    }
    const double kappa = m.lam + 2.0 * m.mu / 3.0;
    // dxi/deps = -A^{-1} dC/deps = d sigma_el/d eps; cauchy = xi so the tangent is the same
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            double v = (a == b) ? m.two_mu : 0.0;
            if (is_diag(a) && is_diag(b)) v += kappa - m.two_mu / 3.0;
            if (A.b.dsig_deps) st(A.b.dsig_deps, a * 6 + b, ld, i, v);
            if (A.b.dxi_deps) st(A.b.dxi_deps, a * 6 + b, ld, i, v);
            if (A.b.dC_dxi) st(A.b.dC_dxi, a * 6 + b, ld, i, (a == b) ? m.inv_two_mu : 0.0);
            if (A.b.dC_dxi_prev) st(A.b.dC_dxi_prev, a * 6 + b, ld, i, 0.0);
        }
    if (A.b.dC_dp && A.n_active > 0) {
        const double (&C)[6] = Cres;
        const double tr = e[0] + e[3] + e[5];
        const double imu = 1.0 / m.mu;
        for (int c = 0; c < A.n_active; ++c) {
            const int pid = A.pid[c];
            const bool el = (pid == CMADX_P_EL0 || pid == CMADX_P_EL1);
            const double dl = el ? m.dlam[pid - CMADX_P_EL0] : 0.0;
            const double dm = el ? m.dmu[pid - CMADX_P_EL0] : 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                // C = (x - lam tr I - 2 mu eps)/(2 mu)
                const double dCl = is_diag(a) ? -tr * m.inv_two_mu : 0.0;
                const double dCm = -e[a] * imu - C[a] * imu;
                st(A.b.dC_dp, (int64_t)a * A.n_active + c, ld, i, dCl * dl + dCm * dm);
            }
        }
    }
}

}  // namespace

cudaError_t launch_mp_update_elastic(const MpArgs& A, cudaStream_t stream) {
    const int64_t nblk = (A.b.n + MP_BLOCK - 1) / MP_BLOCK;
    if (nblk == 0) return cudaSuccess;
    elastic_update_kernel<<<(unsigned)nblk, MP_BLOCK, 0, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace cmadx
