// K3 / K4 for SmallRateElasticPlastic blocks (cmad/models/small_rate_elastic_plastic.py:250-346
// under assemble_element_block, cmad/fem/assembly.py:616-732): the per-point residual sees the
// strain INCREMENT eps(U) - eps(U_prev), so the block carries both displacement vectors
// (cmadx_fe_block_t::U_prev).  State per point = [cauchy(6), alpha]; the cauchy stress handed to
// the momentum residual is the state's own stress (_cauchy_fn :351-359, Q = I) and the consistent
// tangent is the stress rows of the IFT sensitivity, d cauchy / d eps = (dxi/d eps)[0:6]
// (nonlinear_solver.py:158-171).  Rotated material axes: the increment enters the point as
// Q^T de Q, the stress leaves it as Q sig Q^T (:53-72, :351-359).  Any volume rule of tet4 / hex8,
// displacement form and the MIXED u-p form (small_disp_equilibrium.py:87-111): this model's
// hydro_cauchy is tr(cauchy(xi)) / 3 (:369-376), so the pressure rows depend on the local state -
// R_p and the (u,p), (p,u), (p,p) blocks are formed here, next to the point solve, with
// d hydro / d eps = the trace rows of the IFT tangent (not by the state-free pressure kernel of
// fe_mixed.cu).  Correctness-first organisation of fe_generic.cu: one
// thread per element walks its points in the reference's scan order (bit-reproducible sums),
// K_e accumulated in place.  Not a bench path.
#include "fe_common.cuh"
#include "rate_point.cuh"

namespace cmadx {
namespace {

template <int YK, bool WANT_K, int NB, bool MIXED>
__global__ void __launch_bounds__(FE_BLOCK) fe_rate_kernel(const __grid_constant__ FeArgs A,
                                                           const __grid_constant__ cmadx_fe_mixed_t mx) {
    const cmadx_fe_block_t& b = A.b;
    const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e0 < b.n_elems;
    const int64_t e = live ? e0 : 0;
    const int nip = b.n_ip;
    constexpr int ND = NB * 3;
    int eq[ND];
    double U[NB][3], Up[NB][3];
#pragma unroll
    for (int q = 0; q < ND; ++q) eq[q] = __ldg(b.elem_eq + e * ND + q);
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            U[a][k] = __ldg(b.U + eq[3 * a + k]);
            Up[a][k] = __ldg(b.U_prev + eq[3 * a + k]);
        }
    double R[NB][3];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) R[a][i] = 0.0;
    const DevMat& m = A.m;
    DevNewton nw = A.nw;
    nw.defer_after = 0;
    const double lr = m.lam * m.inv_two_mu;
    const bool rot = m.rot != 0;
    double T[6][6], S[6][6];
    if (rot) rot_maps(m.Q, T, S);
    // mixed form: element pressures, pressure residual, constants of the pressure rows
    int eqp[NB];
    double pe[NB], Rp[NB];
    const double kappa = m.lam + 2.0 * m.mu / 3.0;                      // pressure_scale_factor (:378-380)
    double tau = 0.0;
    if constexpr (MIXED) {
        const double h = __ldg(mx.h + e);
        tau = mx.stab_mult * 0.5 * h * h / m.mu;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            eqp[a] = __ldg(mx.elem_eq_p + e * NB + a);
            pe[a] = __ldg(b.U + eqp[a]);
            Rp[a] = 0.0;
        }
    }

    for (int ip = 0; ip < nip; ++ip) {
        const int64_t p = e * nip + ip;
        double gN[NB][3], xp[7], x[7];
        const double* g = b.grad_N + p * ND;
#pragma unroll
        for (int q = 0; q < ND / 4; ++q)
            ld256(g + 4 * q, (&gN[0][0])[4 * q], (&gN[0][0])[4 * q + 1], (&gN[0][0])[4 * q + 2], (&gN[0][0])[4 * q + 3]);
#pragma unroll
        for (int c = 0; c < 7; ++c) { xp[c] = __ldg(b.xi_prev + p * 7 + c); x[c] = xp[c]; }
        const double wdv = __ldg(b.quad_w + ip) * __ldg(b.det + p);
        double de[6], ep[6];
        strain_from_U<NB>(U, gN, de);
        strain_from_U<NB>(Up, gN, ep);
#pragma unroll
        for (int c = 0; c < 6; ++c) de[c] -= ep[c];
        if (rot) {
            double dm[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) s = fma(T[c][q], de[q], s);
                dm[c] = s;
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) de[c] = dm[c];
        }
        RatePoint<YK> pt;
        double C[7];
        const NewtonResult nr = local_newton<RatePoint<YK>, 7>(m, nw, pt, x, xp, de, live, C);
        double sg[6], D[6][6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            sg[a] = x[a];
            if (rot) {
                sg[a] = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) sg[a] = fma(S[a][c], x[c], sg[a]);
            }
        }
        if constexpr (WANT_K) {
            RegLU<7> lu;
            const double dg = x[6] - xp[6];
            pt.jacobian(m, dg, lu.a);
            const bool trouble = lu.factor_natural();
            const bool slow = __any_sync(0xffffffffu, trouble);
            if (slow && trouble) { pt.jacobian(m, dg, lu.a); lu.factor_pivot(); }
#pragma unroll
            for (int bcol = 0; bcol < 6; ++bcol) {
                // -dC/d(de_b) on the stress rows, both branches (rate_point.cuh)
                double col[7];
#pragma unroll
                for (int a = 0; a < 6; ++a) col[a] = ((a == bcol) ? 1.0 : 0.0) + ((is_diag(a) && is_diag(bcol)) ? lr : 0.0);
                col[6] = 0.0;
                if (slow && trouble) lu.solve_pivot(col); else lu.solve_natural(col);
#pragma unroll
                for (int a = 0; a < 6; ++a) D[a][bcol] = col[a];
            }
            if (rot) {           // global tangent S D T
                double DT[6][6];
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        double s = 0.0;
#pragma unroll
                        for (int c = 0; c < 6; ++c) s = fma(D[a][c], T[c][q], s);
                        DT[a][q] = s;
                    }
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        double s = 0.0;
#pragma unroll
                        for (int c = 0; c < 6; ++c) s = fma(S[a][c], DT[c][q], s);
                        D[a][q] = s;
                    }
            }
        }
        if (live) {
#pragma unroll
            for (int c = 0; c < 7; ++c) b.xi[p * 7 + c] = x[c];
            if (b.iters) b.iters[p] = nr.iters;
            if (b.flags) b.flags[p] = nr.flag_entry | ((pt.plastic ? 1 : 0) << 1);
            if (b.sigma) {
#pragma unroll
                for (int a = 0; a < 6; ++a) b.sigma[p * 6 + a] = sg[a];
            }
        }
        if constexpr (MIXED) {
            double Nv[NB], gp[3] = {0.0, 0.0, 0.0}, pr = 0.0;
#pragma unroll
            for (int a = 0; a < NB; ++a) {
                Nv[a] = __ldg(mx.N + ip * NB + a);
                pr = fma(Nv[a], pe[a], pr);
#pragma unroll
                for (int l = 0; l < 3; ++l) gp[l] = fma(gN[a][l], pe[a], gp[l]);
            }
            const double hydro = (sg[0] + sg[3] + sg[5]) / 3.0;                  // hydro_cauchy (:369-376)
            double hrow[6];                                                       // d hydro / d eps (symmetric comps)
            if constexpr (WANT_K) {
#pragma unroll
                for (int q = 0; q < 6; ++q) hrow[q] = (D[0][q] + D[3][q] + D[5][q]) / 3.0;
            }
            mixed_momentum_stress<WANT_K>(pr, sg, D);                              // dev(cauchy) - p I, P_dev D
#pragma unroll
            for (int a = 0; a < NB; ++a) {
                const double gg = gN[a][0] * gp[0] + gN[a][1] * gp[1] + gN[a][2] * gp[2];
                Rp[a] = fma(-(pr + hydro) / kappa * Nv[a] - tau * gg, wdv, Rp[a]);
            }
            if constexpr (WANT_K) {
                if (live) {
                    double* Kpp = mx.K_pp ? mx.K_pp + e * (NB * NB) : nullptr;
                    double* Kup = mx.K_up ? mx.K_up + e * (ND * NB) : nullptr;
                    double* Kpu = mx.K_pu ? mx.K_pu + e * (NB * ND) : nullptr;
#pragma unroll 1
                    for (int a = 0; a < NB; ++a) {
#pragma unroll
                        for (int bb = 0; bb < NB; ++bb) {
                            if (Kpp) {
                                const double gg = gN[a][0] * gN[bb][0] + gN[a][1] * gN[bb][1] + gN[a][2] * gN[bb][2];
                                const double v = (-Nv[a] * Nv[bb] / kappa - tau * gg) * wdv;
                                Kpp[a * NB + bb] = (ip > 0 ? Kpp[a * NB + bb] : 0.0) + v;
                            }
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                if (Kup) {
                                    const double v = -gN[a][k] * Nv[bb] * wdv;
                                    Kup[(3 * a + k) * NB + bb] = (ip > 0 ? Kup[(3 * a + k) * NB + bb] : 0.0) + v;
                                }
                                if (Kpu) {
                                    double s = 0.0;
#pragma unroll
                                    for (int l = 0; l < 3; ++l)
                                        s = fma(hrow[vix(k, l)] * ((k == l) ? 1.0 : 0.5), gN[bb][l], s);
                                    const double v = -(Nv[a] / kappa) * s * wdv;
                                    Kpu[a * ND + 3 * bb + k] = (ip > 0 ? Kpu[a * ND + 3 * bb + k] : 0.0) + v;
                                }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], sg[vix(j, i)], s);
                R[a][i] = fma(s, wdv, R[a][i]);
            }
        if constexpr (WANT_K) {
            if (live) {
#pragma unroll
                for (int al = 0; al < 6; ++al)
#pragma unroll
                    for (int be = 0; be < 6; ++be) D[al][be] *= is_diag(be) ? wdv : 0.5 * wdv;
                double* Ke = b.K_elem + e * (ND * ND);
#pragma unroll 1
                for (int a = 0; a < NB; ++a)
#pragma unroll 1
                    for (int i = 0; i < 3; ++i) {
                        double P[6];
#pragma unroll
                        for (int be = 0; be < 6; ++be) {
                            double s = 0.0;
#pragma unroll
                            for (int j = 0; j < 3; ++j) s = fma(gN[a][j], D[vix(j, i)][be], s);
                            P[be] = s;
                        }
                        double* r = Ke + (3 * a + i) * ND;
#pragma unroll
                        for (int bb = 0; bb < NB; ++bb)
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                double s = 0.0;
#pragma unroll
                                for (int l = 0; l < 3; ++l) s = fma(P[vix(k, l)], gN[bb][l], s);
                                r[3 * bb + k] = (ip > 0 ? r[3 * bb + k] : 0.0) + s;
                            }
                    }
            }
        }
    }
    if (!live) return;
    if (b.R_elem) {
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i) b.R_elem[e * ND + 3 * a + i] = R[a][i];
    }
    if (b.R_global) {
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i) atomicAdd(b.R_global + eq[3 * a + i], R[a][i]);
    }
    if constexpr (MIXED) {
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            if (mx.R_p_elem) mx.R_p_elem[e * NB + a] = Rp[a];
            if (mx.R_global) atomicAdd(mx.R_global + eqp[a], Rp[a]);
        }
    }
}

template <int YK, int NB>
cudaError_t run(const FeArgs& A, const cmadx_fe_mixed_t* mix, cudaStream_t stream) {
    const unsigned nblk = (unsigned)((A.b.n_elems + FE_BLOCK - 1) / FE_BLOCK);
    cmadx_fe_mixed_t mx;
    memset(&mx, 0, sizeof mx);
    if (mix) {
        mx = *mix;
        if (A.b.K_elem) fe_rate_kernel<YK, true, NB, true><<<nblk, FE_BLOCK, 0, stream>>>(A, mx);
        else fe_rate_kernel<YK, false, NB, true><<<nblk, FE_BLOCK, 0, stream>>>(A, mx);
    } else {
        if (A.b.K_elem) fe_rate_kernel<YK, true, NB, false><<<nblk, FE_BLOCK, 0, stream>>>(A, mx);
        else fe_rate_kernel<YK, false, NB, false><<<nblk, FE_BLOCK, 0, stream>>>(A, mx);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fe_rate(const FeArgs& A, const cmadx_fe_mixed_t* mix, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    const bool tet = A.b.n_basis == 4;
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return tet ? run<CMADX_YIELD_J2, 4>(A, mix, stream) : run<CMADX_YIELD_J2, 8>(A, mix, stream);
    case CMADX_YIELD_HILL: return tet ? run<CMADX_YIELD_HILL, 4>(A, mix, stream) : run<CMADX_YIELD_HILL, 8>(A, mix, stream);
    case CMADX_YIELD_HOSFORD: return tet ? run<CMADX_YIELD_HOSFORD, 4>(A, mix, stream) : run<CMADX_YIELD_HOSFORD, 8>(A, mix, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cmadx
