// Pressure blocks of the mixed u-p (stabilised equal-order) formulation of
// SmallDispEquilibrium (cmad/global_residuals/small_disp_equilibrium.py:87-111):
//   R_p[a]         = sum_ip ( -(p + hydro)/kappa N_a - tau gradN_a . grad p ) w dv
//   K_up[(a,i), b] = d R_u[a,i] / d p_b = -sum_ip gradN[a,i] N_b w dv
//   K_pu[a, (b,k)] = d R_p[a] / d U[b,k] = -sum_ip N_a gradN[b,k] w dv      (hydro = kappa tr eps)
//   K_pp[a, b]     = -sum_ip ( N_a N_b / kappa + tau gradN_a . gradN_b ) w dv
// with tau = mult * h^2 / (2 mu).  None of these depends on the local state xi: the local
// Newton only enters the momentum block (R_u, K_uu), which the K3 kernels produce with the
// momentum stress dev(cauchy) - p I (fe_common.cuh: mixed_momentum_stress).  The four blocks
// are emitted as the reference's (r, s)-ordered COO value streams
// (cmad/fem/assembly.py:722-732).  HBM-bound, no Newton: one thread per (element, node a)
// owning row a of R_p / K_pu / K_pp and rows 3a..3a+2 of K_up; sums over the integration
// points in fixed order (bit-reproducible); 256-bit stores of whole 32-byte sectors.
#include <cstdlib>

#include "fe_common.cuh"

namespace cmadx {
namespace {

template <int NB, int NIP>
__global__ void __launch_bounds__(FE_BLOCK, 4) fe_mixed_pressure_kernel(const cmadx_fe_block_t b,
                                                                        const cmadx_fe_mixed_t mx,
                                                                        const double kappa, const double mu) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / NB;
    const int a = (int)(t - e * NB);
    const bool live = e < b.n_elems;
    if (__all_sync(0xffffffffu, !live)) return;
    const int64_t el = live ? e : 0;
    const int nip = NIP ? NIP : b.n_ip;        // NIP = 0: any quadrature rule (runtime count)
    // this thread's node: displacement and pressure dofs; the sums over the nodes of the
    // element (p, tr eps, grad p) are butterfly reductions over its NB lanes
    double Ua[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) Ua[k] = __ldg(b.U + __ldg(b.elem_eq + el * (NB * 3) + 3 * a + k));
    const double pa = __ldg(b.U + __ldg(mx.elem_eq_p + el * NB + a));
    const double h = __ldg(mx.h + el);
    const double tau = mx.stab_mult * 0.5 * h * h / mu;
    const double ik = 1.0 / kappa;
    double Rp = 0.0;
    // hex8: the element's grad_N (8 points x 24 doubles) is staged ONCE in shared memory by
    // contiguous 32-byte chunks (the 8 threads of an element cover 256 contiguous bytes per
    // request) and every later read is a broadcast within the element; the region stride of
    // 194 doubles puts the 4 elements of a warp 4 banks apart, so 128-bit reads of the same
    // offset are conflict-free.  (Before: every thread re-read the chunks through L1 - 8x the
    // wavefronts, the kernel was L1-data-pipe bound.)  tet4 keeps the direct loads (12 doubles).
    constexpr int REGION = 194;
    constexpr bool STAGED = (NB == 8 && NIP == 8);
    __shared__ __align__(16) double smem[STAGED ? (FE_BLOCK / 8) * REGION : 2];
    const double* reg = smem + (STAGED ? (threadIdx.x >> 3) * REGION : 0);
    if constexpr (STAGED) {
        double* wr = smem + (threadIdx.x >> 3) * REGION;
        const double* g = b.grad_N + el * (NIP * NB * 3);
        double c[6][4];
#pragma unroll
        for (int r = 0; r < 6; ++r) ld256(g + 4 * (a + 8 * r), c[r][0], c[r][1], c[r][2], c[r][3]);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double2* dst = reinterpret_cast<double2*>(wr + 4 * (a + 8 * r));
            dst[0] = make_double2(c[r][0], c[r][1]);
            dst[1] = make_double2(c[r][2], c[r][3]);
        }
        __syncwarp();
    }
    // column blocks of 4 nodes keep the accumulators (28 doubles) and the grad_N quarter in
    // registers at 4 blocks / SM
#pragma unroll 1
    for (int cb = 0; cb < NB / 4; ++cb) {
        double Kpu[4][3], Kup[3][4], Kpp[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            Kpp[c] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { Kpu[c][k] = 0.0; Kup[k][c] = 0.0; }
        }
#pragma unroll 1
        for (int q = 0; q < nip; ++q) {
            const double* g = b.grad_N + (el * nip + q) * (NB * 3);
            double gN[4][3], N[4];
            if constexpr (STAGED) {
                const double2* src = reinterpret_cast<const double2*>(reg + q * 24 + 12 * cb);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double2 v = src[c];
                    (&gN[0][0])[2 * c] = v.x; (&gN[0][0])[2 * c + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double v0, v1, v2, v3;
                    ld256(g + 12 * cb + 4 * c, v0, v1, v2, v3);
                    (&gN[0][0])[4 * c] = v0; (&gN[0][0])[4 * c + 1] = v1; (&gN[0][0])[4 * c + 2] = v2; (&gN[0][0])[4 * c + 3] = v3;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) N[c] = __ldg(mx.N + q * NB + 4 * cb + c);
            const double wdv = __ldg(b.quad_w + q) * __ldg(b.det + el * nip + q);
            const double Na = __ldg(mx.N + q * NB + a) * wdv;
            const double gu0 = STAGED ? reg[q * 24 + 3 * a] : __ldg(g + 3 * a);
            const double gu1 = STAGED ? reg[q * 24 + 3 * a + 1] : __ldg(g + 3 * a + 1);
            const double gu2 = STAGED ? reg[q * 24 + 3 * a + 2] : __ldg(g + 3 * a + 2);
            const double ga0 = gu0 * wdv, ga1 = gu1 * wdv, ga2 = gu2 * wdv;
            if (cb == 0) {
                double p = Na * pa, tre = fma(Ua[2], ga2, fma(Ua[1], ga1, Ua[0] * ga0));    // x w dv
                double gp[3] = {pa * ga0, pa * ga1, pa * ga2};
#pragma unroll
                for (int m = 1; m < NB; m <<= 1) {
                    p += __shfl_xor_sync(0xffffffffu, p, m);
                    tre += __shfl_xor_sync(0xffffffffu, tre, m);
#pragma unroll
                    for (int k = 0; k < 3; ++k) gp[k] += __shfl_xor_sync(0xffffffffu, gp[k], m);
                }
                // p, tre, gp carry one factor w dv; N_a and gradN_a (unweighted) = Na / wdv ...
                const double Nu = __ldg(mx.N + q * NB + a);
                Rp -= fma(p, ik, tre) * Nu + tau * fma(gu2, gp[2], fma(gu1, gp[1], gu0 * gp[0]));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                Kpp[c] -= fma(Na * ik, N[c], tau * fma(ga2, gN[c][2], fma(ga1, gN[c][1], ga0 * gN[c][0])));
#pragma unroll
                for (int k = 0; k < 3; ++k) Kpu[c][k] = fma(-Na, gN[c][k], Kpu[c][k]);
                Kup[0][c] = fma(-ga0, N[c], Kup[0][c]);
                Kup[1][c] = fma(-ga1, N[c], Kup[1][c]);
                Kup[2][c] = fma(-ga2, N[c], Kup[2][c]);
            }
        }
        if (live) {
            if (mx.K_pu) {
                double* r = mx.K_pu + (e * NB + a) * (NB * 3) + 12 * cb;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    st256(r + 4 * c, (&Kpu[0][0])[4 * c], (&Kpu[0][0])[4 * c + 1], (&Kpu[0][0])[4 * c + 2], (&Kpu[0][0])[4 * c + 3]);
            }
            if (mx.K_up) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    st256(mx.K_up + (e * NB * 3 + 3 * a + i) * NB + 4 * cb, Kup[i][0], Kup[i][1], Kup[i][2], Kup[i][3]);
            }
            if (mx.K_pp) st256(mx.K_pp + (e * NB + a) * NB + 4 * cb, Kpp[0], Kpp[1], Kpp[2], Kpp[3]);
        }
    }
    if (!live) return;
    if (mx.R_p_elem) mx.R_p_elem[e * NB + a] = Rp;
    if (mx.R_global) atomicAdd(mx.R_global + __ldg(mx.elem_eq_p + e * NB + a), Rp);
}


// hex8 x 8 points, 2-D register tiling.  The one-row-per-thread kernel above is L1-data-pipe
// bound on hex8 (ncu: 81 % LSU, 0.23 shared loads per FMA - every thread needs all 24 grad_N
// entries of every point for ITS row).  Here lane t of an element owns the 2 x 12 tile
// rows {2(t&3), 2(t&3)+1} x column nodes {4(t>>2) .. 4(t>>2)+3} of K_pu (and the matching 2 x 4
// tile of K_pp): one pass over the 8 points, every 12-double grad_N quarter feeds both rows
// (0.09 loads per FMA), no column-block loop.  K_up is the transpose of K_pu
// (K_up[(c,k), r] = K_pu[r, (c,k)] = -sum_ip N_r gradN[c,k] w dv): written from the same
// accumulators as 16-byte pieces, the 4 lanes of a column block filling one 64-byte row.
// R_p keeps the lane-owns-node-t butterfly sums (fixed order, bit-reproducible).
__global__ void __launch_bounds__(FE_BLOCK, 4) fe_mixed_pressure_hex8_kernel(const cmadx_fe_block_t b,
                                                                             const cmadx_fe_mixed_t mx,
                                                                             const double kappa, const double mu) {
    constexpr int NB = 8, NIP = 8, REGION = 194;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t >> 3;
    const int a = (int)(t & 7);
    const bool live = e < b.n_elems;
    __shared__ __align__(16) double smem[(FE_BLOCK / 8) * REGION];
    __shared__ __align__(16) double Ns[NIP * NB];
    if (threadIdx.x < NIP * NB) Ns[threadIdx.x] = __ldg(mx.N + threadIdx.x);
    const int64_t el = live ? e : 0;
    double Ua[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) Ua[k] = __ldg(b.U + __ldg(b.elem_eq + el * (NB * 3) + 3 * a + k));
    const double pa = __ldg(b.U + __ldg(mx.elem_eq_p + el * NB + a));
    const double h = __ldg(mx.h + el);
    const double tau = mx.stab_mult * 0.5 * h * h / mu;
    const double ik = 1.0 / kappa;
    const double wdv_mine = __ldg(b.quad_w + a) * __ldg(b.det + el * NIP + a);     // lane a carries point a's w dv
    double* reg = smem + (threadIdx.x >> 3) * REGION;
    {
        const double* g = b.grad_N + el * (NIP * NB * 3);
        double c[6][4];
#pragma unroll
        for (int r = 0; r < 6; ++r) ld256(g + 4 * (a + 8 * r), c[r][0], c[r][1], c[r][2], c[r][3]);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double2* dst = reinterpret_cast<double2*>(reg + 4 * (a + 8 * r));
            dst[0] = make_double2(c[r][0], c[r][1]);
            dst[1] = make_double2(c[r][2], c[r][3]);
        }
    }
    __syncthreads();
    // R_p first (its temporaries are dead before the tile accumulators become live)
    double Rp = 0.0;
#pragma unroll 1
    for (int q = 0; q < NIP; ++q) {
        const double wdv = __shfl_sync(0xffffffffu, wdv_mine, q, 8);
        const double* gq = reg + q * 24;
        const double Nu = Ns[q * NB + a];
        const double gu0 = gq[3 * a], gu1 = gq[3 * a + 1], gu2 = gq[3 * a + 2];
        const double Na = Nu * wdv, ga0 = gu0 * wdv, ga1 = gu1 * wdv, ga2 = gu2 * wdv;
        double p = Na * pa, tre = fma(Ua[2], ga2, fma(Ua[1], ga1, Ua[0] * ga0));
        double gp[3] = {pa * ga0, pa * ga1, pa * ga2};
#pragma unroll
        for (int m = 1; m < NB; m <<= 1) {
            p += __shfl_xor_sync(0xffffffffu, p, m);
            tre += __shfl_xor_sync(0xffffffffu, tre, m);
#pragma unroll
            for (int k = 0; k < 3; ++k) gp[k] += __shfl_xor_sync(0xffffffffu, gp[k], m);
        }
        Rp -= fma(p, ik, tre) * Nu + tau * fma(gu2, gp[2], fma(gu1, gp[1], gu0 * gp[0]));
    }
    const int rp = (a & 3) * 2, cb = a >> 2;
    double Kpu[2][12], Kpp[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int c = 0; c < 12; ++c) Kpu[r][c] = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) Kpp[r][c] = 0.0;
    }
#pragma unroll 1
    for (int q = 0; q < NIP; ++q) {
        const double wdv = __shfl_sync(0xffffffffu, wdv_mine, q, 8);
        const double* gq = reg + q * 24;
        double gN[12], Nc[4];
        {
            const double2* src = reinterpret_cast<const double2*>(gq + 12 * cb);
#pragma unroll
            for (int c = 0; c < 6; ++c) { const double2 v = src[c]; gN[2 * c] = v.x; gN[2 * c + 1] = v.y; }
            const double2* ns = reinterpret_cast<const double2*>(Ns + q * NB + 4 * cb);
            const double2 n0 = ns[0], n1 = ns[1];
            Nc[0] = n0.x; Nc[1] = n0.y; Nc[2] = n1.x; Nc[3] = n1.y;
        }
        const double2 nr = *reinterpret_cast<const double2*>(Ns + q * NB + rp);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const double Nr = (r ? nr.y : nr.x) * wdv;
            const double g0 = gq[3 * (rp + r)] * wdv, g1 = gq[3 * (rp + r) + 1] * wdv, g2 = gq[3 * (rp + r) + 2] * wdv;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                Kpp[r][c] -= fma(Nr * ik, Nc[c], tau * fma(g2, gN[3 * c + 2], fma(g1, gN[3 * c + 1], g0 * gN[3 * c])));
#pragma unroll
                for (int k = 0; k < 3; ++k) Kpu[r][3 * c + k] = fma(-Nr, gN[3 * c + k], Kpu[r][3 * c + k]);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (mx.K_pu) {
            double* d = mx.K_pu + (e * NB + rp + r) * (NB * 3) + 12 * cb;
#pragma unroll
            for (int c = 0; c < 3; ++c) st256(d + 4 * c, Kpu[r][4 * c], Kpu[r][4 * c + 1], Kpu[r][4 * c + 2], Kpu[r][4 * c + 3]);
        }
        if (mx.K_pp) st256(mx.K_pp + (e * NB + rp + r) * NB + 4 * cb, Kpp[r][0], Kpp[r][1], Kpp[r][2], Kpp[r][3]);
    }
    if (mx.K_up) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {      // row (node 4cb + c/3, component c%3) of K_up, columns rp, rp+1
            double* d = mx.K_up + (e * (NB * 3) + 12 * cb + c) * NB + rp;
            asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(d), "d"(Kpu[0][c]), "d"(Kpu[1][c]) : "memory");
        }
    }
    if (mx.R_p_elem) mx.R_p_elem[e * NB + a] = Rp;
    if (mx.R_global) atomicAdd(mx.R_global + __ldg(mx.elem_eq_p + e * NB + a), Rp);
}

// K6 for the pressure block: tangent of R_p at fixed local state.  R_p is linear in the
// (u, p) dofs and depends on the parameters only through kappa and mu:
//   dR_p[a] = sum_ip ( (p dkappa/kappa^2 - dp/kappa - tr(d eps)) N_a
//                      + tau_h (dmu/mu^2 gradN_a.grad p - 1/mu gradN_a.grad dp) ) w dv,
// tau_h = mult h^2 / 2, (dp, d eps) interpolated from the direction dU (or zero).  This is the
// pressure rows of what jax.jvp pushes through the assembled mixed residual
// (cmad/fem/nonlinear_solver.py:490-537 over small_disp_equilibrium.py:94-110).  One thread
// per (element, node); nodal sums by butterfly shuffles over the element's lanes.
template <int NB, int NIP>
__global__ void __launch_bounds__(FE_BLOCK) fe_mixed_pressure_jvp_kernel(const cmadx_fe_block_t b,
                                                                         const cmadx_fe_mixed_t mx,
                                                                         const double* __restrict__ dU,
                                                                         const double kappa, const double mu,
                                                                         const double dkappa, const double dmu) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / NB;
    const int a = (int)(t - e * NB);
    const bool live = e < b.n_elems;
    if (__all_sync(0xffffffffu, !live)) return;
    const int64_t el = live ? e : 0;
    const int eqp = __ldg(mx.elem_eq_p + el * NB + a);
    const double pa = __ldg(b.U + eqp);
    double dUa[3] = {0.0, 0.0, 0.0}, dpa = 0.0;
    if (dU) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dUa[k] = __ldg(dU + __ldg(b.elem_eq + el * (NB * 3) + 3 * a + k));
        dpa = __ldg(dU + eqp);
    }
    const double h = __ldg(mx.h + el);
    const double tau_h = mx.stab_mult * 0.5 * h * h;
    const double cp = dkappa / (kappa * kappa), ct = dmu / (mu * mu), ik = 1.0 / kappa, im = 1.0 / mu;
    const int nip = NIP ? NIP : b.n_ip;
    double dRp = 0.0;
#pragma unroll 1
    for (int q = 0; q < nip; ++q) {
        const double* g = b.grad_N + (el * nip + q) * (NB * 3);
        const double Na = __ldg(mx.N + q * NB + a);
        const double g0 = __ldg(g + 3 * a), g1 = __ldg(g + 3 * a + 1), g2 = __ldg(g + 3 * a + 2);
        const double wdv = __ldg(b.quad_w + q) * __ldg(b.det + el * nip + q);
        // v[0] = p, v[1..3] = grad p, v[4] = dp, v[5] = tr(d eps), v[6..8] = grad dp
        double v[9] = {Na * pa, pa * g0, pa * g1, pa * g2, Na * dpa,
                       fma(dUa[2], g2, fma(dUa[1], g1, dUa[0] * g0)), dpa * g0, dpa * g1, dpa * g2};
#pragma unroll
        for (int m = 1; m < NB; m <<= 1)
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], m);
        const double gg = fma(g2, v[3], fma(g1, v[2], g0 * v[1]));
        const double gd = fma(g2, v[8], fma(g1, v[7], g0 * v[6]));
        dRp = fma(wdv, (cp * v[0] - ik * v[4] - v[5]) * Na + tau_h * (ct * gg - im * gd), dRp);
    }
    if (!live) return;
    if (mx.R_p_elem) mx.R_p_elem[e * NB + a] = dRp;
    if (mx.R_global) atomicAdd(mx.R_global + eqp, dRp);
}

}  // namespace

cudaError_t launch_fe_mixed_pressure_jvp(const cmadx_fe_block_t& b, const cmadx_fe_mixed_t& mx, const double* dU,
                                         double kappa, double mu, double dkappa, double dmu, cudaStream_t stream) {
    if (b.n_elems == 0) return cudaSuccess;
    const int64_t nthr = b.n_elems * b.n_basis;
    const unsigned nblk = (unsigned)((nthr + FE_BLOCK - 1) / FE_BLOCK);
    if (b.n_basis == 4 && b.n_ip == 1) fe_mixed_pressure_jvp_kernel<4, 1><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, dU, kappa, mu, dkappa, dmu);
    else if (b.n_basis == 8 && b.n_ip == 8) fe_mixed_pressure_jvp_kernel<8, 8><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, dU, kappa, mu, dkappa, dmu);
    else if (b.n_basis == 4) fe_mixed_pressure_jvp_kernel<4, 0><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, dU, kappa, mu, dkappa, dmu);
    else fe_mixed_pressure_jvp_kernel<8, 0><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, dU, kappa, mu, dkappa, dmu);
    return cudaGetLastError();
}

cudaError_t launch_fe_mixed_pressure(const cmadx_fe_block_t& b, const cmadx_fe_mixed_t& mx, double kappa,
                                     double mu, cudaStream_t stream) {
    if (b.n_elems == 0) return cudaSuccess;
    const int64_t nthr = b.n_elems * b.n_basis;
    const unsigned nblk = (unsigned)((nthr + FE_BLOCK - 1) / FE_BLOCK);
    if (b.n_basis == 4 && b.n_ip == 1) fe_mixed_pressure_kernel<4, 1><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, kappa, mu);
    else if (b.n_basis == 8 && b.n_ip == 8) {
        static const bool row_kernel = std::getenv("CMADX_PRESSURE_ROW_KERNEL") != nullptr;    // A/B measurements
        if (row_kernel) fe_mixed_pressure_kernel<8, 8><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, kappa, mu);
        else fe_mixed_pressure_hex8_kernel<<<nblk, FE_BLOCK, 0, stream>>>(b, mx, kappa, mu);
    }
    else if (b.n_basis == 4) fe_mixed_pressure_kernel<4, 0><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, kappa, mu);   // any rule
    else fe_mixed_pressure_kernel<8, 0><<<nblk, FE_BLOCK, 0, stream>>>(b, mx, kappa, mu);
    return cudaGetLastError();
}

}  // namespace cmadx
