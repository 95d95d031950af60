// K6 (reverse mode) - VJP of the converged element block w.r.t. (params, xi_prev) at
// fixed U: the transpose of cmadx_fe_block_jvp, i.e. the per-step body of a discrete
// FE adjoint.  Given the nodal adjoint Rbar (cotangent of the assembled residual) and
// the cotangent xibar of the converged local state:
//   sbar    = w dv sym-pack(gradN^T Rbar_e)                  cotangent of cauchy
//   xbar    = xibar + (d cauchy/d xi)^T sbar
//   mu      = -A^{-T} xbar,   A = dC/dxi at (xi_state, xi_prev)
//   pbar   += (dC/dp)^T mu + (d cauchy/dp)^T sbar            summed over all points
//   xibar_prev = (dC/dxi_prev)^T mu
// This is what transposing the jax.jvp of the assembled residual (cmad/fem/
// nonlinear_solver.py:490-537 through cmad/models/nonlinear_solver.py:158-171) yields
// under jax.grad (cmad/cli/gradient.py:74-82).  One thread per integration point; pbar is
// reduced in fixed order (warp shuffles -> shared memory -> per-block partials -> one
// block), so the gradient is bit-reproducible; under torch.distributed it is then
// all-reduced with the objective.
#include "fe_common.cuh"

namespace cmadx {
cudaError_t launch_reduce_partials(const double* partials, int64_t nblk, int ncols, double* result,
                                   cudaStream_t stream);

namespace {

constexpr int VJP_BLOCK = 128;

// WANT_U: also emit the per-point cotangent of the element displacements,
//   Ubar_ip[p][(a,k)] = sum_j gbar[k][j] gN[a][j],  gbar = sym-unpack(T^T E^T (xibar + mu)),
// i.e. (d xi/dU)^T xibar + (d R/dU |total)^T Rbar restricted to this point: with
// d xi/d eps = E - A^{-1} E and d sigma/d eps |total = Cel E^T A^{-1} E the explicit stress
// terms cancel and the strain cotangent is E^T (xibar + mu).  Rbar may be NULL there.
template <int YK, bool ROT, int NB, bool WANT_U = false>
__global__ void __launch_bounds__(VJP_BLOCK, (YK == CMADX_YIELD_J2 && !ROT && !WANT_U) ? 4 : 1)
fe_vjp_kernel(const __grid_constant__ FeArgs A, const double* __restrict__ Rbar,
              const double* __restrict__ xibar, double* __restrict__ partials,
              double* __restrict__ Ubar_ip = nullptr) {
    const cmadx_fe_block_t& b = A.b;
    const DevMat& m = A.m;
    constexpr int nb = NB;
    const int nip = b.n_ip, na = A.n_active;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npts = b.n_elems * nip;
    const bool live = p < npts;
    const int64_t e = live ? p / nip : 0;
    const int ip = live ? (int)(p - e * nip) : 0;

    double g[CMADX_MAX_ACTIVE];
#pragma unroll
    for (int c = 0; c < CMADX_MAX_ACTIVE; ++c) g[c] = 0.0;

    // strain and sbar in one pass over the element's nodes
    double gu[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, sb33[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double xs[7], xp[7], xb[7];
    double wdv = 0.0;
    if (live) {
        // the point's grad_N rows and the element's equation row as whole 32- / 16-byte chunks
        const double* gN = b.grad_N + (p * nb) * 3;
        double gr[NB * 3];
        int eqr[NB * 3];
#pragma unroll
        for (int q = 0; q < NB * 3 / 4; ++q) {
            ld256(gN + 4 * q, gr[4 * q], gr[4 * q + 1], gr[4 * q + 2], gr[4 * q + 3]);
            const int4 v = __ldg(reinterpret_cast<const int4*>(b.elem_eq + e * (NB * 3)) + q);
            eqr[4 * q] = v.x; eqr[4 * q + 1] = v.y; eqr[4 * q + 2] = v.z; eqr[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int a = 0; a < nb; ++a) {
            const double g0 = gr[3 * a], g1 = gr[3 * a + 1], g2 = gr[3 * a + 2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int eq = eqr[3 * a + k];
                const double u = __ldg(b.U + eq), rb = (WANT_U && !Rbar) ? 0.0 : __ldg(Rbar + eq);
                gu[k][0] = fma(u, g0, gu[k][0]); gu[k][1] = fma(u, g1, gu[k][1]); gu[k][2] = fma(u, g2, gu[k][2]);
                // R[a][i] = sum_j gN[a][j] sigma[j][i] w dv  ->  sbar[j][i] += gN[a][j] Rbar[a][i]
                sb33[0][k] = fma(g0, rb, sb33[0][k]); sb33[1][k] = fma(g1, rb, sb33[1][k]); sb33[2][k] = fma(g2, rb, sb33[2][k]);
            }
        }
        wdv = __ldg(b.quad_w + ip) * __ldg(b.det + p);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            xs[c] = __ldg(A.xi_state + p * 7 + c);
            xp[c] = __ldg(b.xi_prev + p * 7 + c);
            xb[c] = xibar ? __ldg(xibar + p * 7 + c) : 0.0;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) { xs[c] = 0.0; xp[c] = 0.0; xb[c] = 0.0; }
        gu[0][0] = 1e-3;
    }
    double eg[6], sbg[6];      // global axes: symmetric strain, packed cotangent of cauchy
    eg[0] = gu[0][0]; eg[3] = gu[1][1]; eg[5] = gu[2][2];
    eg[1] = 0.5 * (gu[0][1] + gu[1][0]); eg[2] = 0.5 * (gu[0][2] + gu[2][0]); eg[4] = 0.5 * (gu[1][2] + gu[2][1]);
    sbg[0] = sb33[0][0] * wdv; sbg[3] = sb33[1][1] * wdv; sbg[5] = sb33[2][2] * wdv;
    sbg[1] = (sb33[0][1] + sb33[1][0]) * wdv; sbg[2] = (sb33[0][2] + sb33[2][0]) * wdv; sbg[4] = (sb33[1][2] + sb33[2][1]) * wdv;
    // mixed u-p: R_u sees dev(cauchy) - p I, and dev is self-adjoint: project the cotangent.
    // The pressure rows depend on the parameters through kappa and mu only
    // (small_disp_equilibrium.py:101-110): d/dkappa = p/kappa^2 N_a, d/dmu = tau/mu gradN_a.grad p
    double mixk = 0.0, mixm = 0.0;     // cotangents of kappa and mu from the pressure block
    if (A.mix_eq_p) {
        const double tr3 = (sbg[0] + sbg[3] + sbg[5]) / 3.0;
        sbg[0] -= tr3; sbg[3] -= tr3; sbg[5] -= tr3;
        if (live) {
            const double* gN = b.grad_N + (p * nb) * 3;
            double rbN = 0.0, pip = 0.0, rbG[3] = {0.0, 0.0, 0.0}, gp[3] = {0.0, 0.0, 0.0};
            for (int a = 0; a < nb; ++a) {
                const int eqp = __ldg(A.mix_eq_p + e * nb + a);
                const double rb = (WANT_U && !Rbar) ? 0.0 : __ldg(Rbar + eqp);
                const double pa = __ldg(b.U + eqp), Na = __ldg(A.mix_N + ip * nb + a);
                rbN = fma(rb, Na, rbN); pip = fma(pa, Na, pip);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double gk = __ldg(gN + 3 * a + k);
                    rbG[k] = fma(rb, gk, rbG[k]); gp[k] = fma(pa, gk, gp[k]);
                }
            }
            const double kappa = m.lam + 2.0 * m.mu / 3.0;
            const double h = __ldg(A.mix_h + e);
            mixk = wdv * pip * rbN / (kappa * kappa);
            mixm = wdv * A.mix_stab * 0.5 * h * h / (m.mu * m.mu) * fma(rbG[2], gp[2], fma(rbG[1], gp[1], rbG[0] * gp[0]));
        }
    }
    double em[6], sbar[6];     // material axes
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0, t = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) { s = fma(T[c][q], eg[q], s); t = fma(S[q][c], sbg[q], t); }   // S^T sbar
            em[c] = s; sbar[c] = t;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) { em[c] = eg[c]; sbar[c] = sbg[c]; }
    }

    SepPoint<YK> pt;
    double Cs[7];
    pt.residual(m, xs, xp, em, Cs);
    const bool pl = pt.plastic;
    const double dg = xs[6] - xp[6];
    double ee[6], sig[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = em[a] - xs[a];
    const double tre = ee[0] + ee[3] + ee[5];
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], m.lam * tre) : m.two_mu * ee[a];
    // xbar = xibar + (d cauchy/d ep)^T sbar,  d sigma_a/d ep_b = -(2mu delta_ab + lam [a diag][b diag])
    const double strb = sbar[0] + sbar[3] + sbar[5];
#pragma unroll
    for (int c = 0; c < 6; ++c) xb[c] += is_diag(c) ? fma(-m.two_mu, sbar[c], -m.lam * strb) : -m.two_mu * sbar[c];
    // mu = -A^{-T} xbar.  J2: the Jacobian's inverse in closed form (j2_radial.cuh), no LU;
    // other surfaces: threshold-pivoted register LU of the transposed Jacobian
    double mu[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) mu[c] = xb[c];
    constexpr bool CLOSED = (YK == CMADX_YIELD_J2);
    if constexpr (CLOSED) {
        j2_jacobian_solve<true>(m, pt, dg, mu);
    } else {
        RegLU<7> lu;
        {
            double Jm[7][7];
            pt.jacobian(m, dg, Jm);
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int c = 0; c < 7; ++c) lu.a[a][c] = Jm[c][a];
        }
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) {
            double Jm[7][7];
            pt.jacobian(m, dg, Jm);
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int c = 0; c < 7; ++c) lu.a[a][c] = Jm[c][a];
            lu.factor_pivot();
        }
        if (slow && trouble) lu.solve_pivot(mu); else lu.solve_natural(mu);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) mu[c] = -mu[c];
    if constexpr (WANT_U) {
        if (live) {
            // strain cotangent in material axes: xibar + mu on the strain-like rows (xb holds
            // xibar - Cel sbar: add the stress term back), then to global axes and to grad_u
            double ebm[6], ebg[6];
#pragma unroll
            for (int c = 0; c < 6; ++c)
                ebm[c] = xb[c] + (is_diag(c) ? fma(m.two_mu, sbar[c], m.lam * strb) : m.two_mu * sbar[c]) + mu[c];
            if (ROT) {
                double T[6][6], S[6][6];
                rot_maps(m.Q, T, S);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    double t = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) t = fma(T[c][q], ebm[c], t);     // T^T ebar_m
                    ebg[q] = t;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 6; ++c) ebg[c] = ebm[c];
            }
            // gbar[k][j]: diagonal entries take ebar, off-diagonal entries half of it each
            double gb[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int j = 0; j < 3; ++j) gb[k][j] = (k == j) ? ebg[vix(k, j)] : 0.5 * ebg[vix(k, j)];
            const double* gN = b.grad_N + (p * nb) * 3;
            double* ub = Ubar_ip + p * (NB * 3);
#pragma unroll
            for (int a = 0; a < NB; ++a) {
                const double g0 = __ldg(gN + 3 * a), g1 = __ldg(gN + 3 * a + 1), g2 = __ldg(gN + 3 * a + 2);
#pragma unroll
                for (int k = 0; k < 3; ++k) ub[3 * a + k] = fma(gb[k][2], g2, fma(gb[k][1], g1, gb[k][0] * g0));
            }
        }
    }
    // xibar_prev = B^T mu;  B = [-I, n; 0, 0] (plastic) or -I (elastic)
    if (live) {
        double nmu = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) { b.xi[p * 7 + a] = -mu[a]; nmu = fma(pt.n[a], mu[a], nmu); }
        b.xi[p * 7 + 6] = pl ? nmu : -mu[6];
    }
    // pbar
    double Mee[6], nee = 0.0, see = 0.0;
    if constexpr (CLOSED) pt.yf.Mvec(ee, Mee);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        if constexpr (!CLOSED) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) s = fma(pt.yf.M(a, q), ee[q], s);
            Mee[a] = s;
        }
        nee = fma(mult(a) * pt.n[a], ee[a], nee);
        see = fma(sbar[a], ee[a], see);
    }
    if (live) {
#pragma unroll
        for (int c = 0; c < CMADX_MAX_ACTIVE; ++c) {
            if (c < na) {
                const int pid = A.pid[c];
                double col[7];
                dC_dp_column(m, pid, pl, pt.yf, pt.n, pt.f, pt.eD, xs[6], dg, Mee, nee, sig, col);
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 7; ++q) acc = fma(mu[q], col[q], acc);
                if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
                    const double dl = m.dlam[pid - CMADX_P_EL0], dm = m.dmu[pid - CMADX_P_EL0];
                    acc += dl * tre * strb + 2.0 * dm * see;
                    acc += (dl + 2.0 * dm / 3.0) * mixk + dm * mixm;
                }
                g[c] = acc;
            }
        }
    }
    // ---- block reduction (fixed order): warp shuffles, then the warps via shared memory
    __shared__ double sm[VJP_BLOCK / 32][CMADX_MAX_ACTIVE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CMADX_MAX_ACTIVE; ++c) {
        if (c < na) {
            double v = g[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) sm[warp][c] = v;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < na) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < VJP_BLOCK / 32; ++w) v += sm[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * na + threadIdx.x] = v;
    }
}

template <int YK>
cudaError_t launch_yk(const FeArgs& A, const double* Rbar, const double* xibar, double* partials,
                      unsigned nblk, cudaStream_t s) {
    if (A.b.n_basis == 4) {
        if (A.m.rot) fe_vjp_kernel<YK, true, 4><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, partials);
        else fe_vjp_kernel<YK, false, 4><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, partials);
    } else {
        if (A.m.rot) fe_vjp_kernel<YK, true, 8><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, partials);
        else fe_vjp_kernel<YK, false, 8><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, partials);
    }
    return cudaGetLastError();
}

}  // namespace

int64_t fe_vjp_blocks(int64_t npts) { return (npts + VJP_BLOCK - 1) / VJP_BLOCK; }

// displacement cotangent only (no parameter gradient): per-point contributions Ubar_ip
template <int YK>
cudaError_t launch_yk_disp(const FeArgs& A, const double* Rbar, const double* xibar, double* Ubar_ip,
                           unsigned nblk, cudaStream_t s) {
    if (A.b.n_basis == 4) {
        if (A.m.rot) fe_vjp_kernel<YK, true, 4, true><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, nullptr, Ubar_ip);
        else fe_vjp_kernel<YK, false, 4, true><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, nullptr, Ubar_ip);
    } else {
        if (A.m.rot) fe_vjp_kernel<YK, true, 8, true><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, nullptr, Ubar_ip);
        else fe_vjp_kernel<YK, false, 8, true><<<nblk, VJP_BLOCK, 0, s>>>(A, Rbar, xibar, nullptr, Ubar_ip);
    }
    return cudaGetLastError();
}

cudaError_t launch_fe_block_vjp_disp(const FeArgs& A, const double* Rbar, const double* xibar, double* Ubar_ip,
                                     cudaStream_t s) {
    const int64_t npts = A.b.n_elems * A.b.n_ip;
    if (npts == 0) return cudaSuccess;
    const unsigned nblk = (unsigned)fe_vjp_blocks(npts);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_yk_disp<CMADX_YIELD_J2>(A, Rbar, xibar, Ubar_ip, nblk, s);
    case CMADX_YIELD_HILL: return launch_yk_disp<CMADX_YIELD_HILL>(A, Rbar, xibar, Ubar_ip, nblk, s);
    case CMADX_YIELD_HOSFORD: return launch_yk_disp<CMADX_YIELD_HOSFORD>(A, Rbar, xibar, Ubar_ip, nblk, s);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fe_block_vjp(const FeArgs& A, const double* Rbar, const double* xibar, double* partials,
                                double* pbar, cudaStream_t s) {
    const int64_t npts = A.b.n_elems * A.b.n_ip;
    if (npts == 0) return (A.n_active > 0) ? cudaMemsetAsync(pbar, 0, sizeof(double) * A.n_active, s) : cudaSuccess;
    const unsigned nblk = (unsigned)fe_vjp_blocks(npts);
    cudaError_t e;
    switch (A.m.yield) {
    case CMADX_YIELD_J2: e = launch_yk<CMADX_YIELD_J2>(A, Rbar, xibar, partials, nblk, s); break;
    case CMADX_YIELD_HILL: e = launch_yk<CMADX_YIELD_HILL>(A, Rbar, xibar, partials, nblk, s); break;
    case CMADX_YIELD_HOSFORD: e = launch_yk<CMADX_YIELD_HOSFORD>(A, Rbar, xibar, partials, nblk, s); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess || A.n_active == 0) return e;
    return launch_reduce_partials(partials, nblk, A.n_active, pbar, s);
}

}  // namespace cmadx
