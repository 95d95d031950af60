// K3 / K4 / K6-JVP for ANY volume quadrature rule of the two element families: tet4 and hex8
// with n_ip != the defaults (tet4 x 1, hex8 x 8 have their own tuned kernels, fe_tet4.cu /
// fe_hex8.cu).  The reference lets a deck override the rule (`discretization.quadrature.volume
// degree`, cmad/cli/common.py:497-540: tet_quadrature 1..6 -> 1, 4, 5, 11, 15, 24 points,
// hex_quadrature d -> ceil((d+1)/2)^3 points) and REQUIRES degree >= 2 for the mixed u-p
// formulation (cmad/cli/common.py:379-391), i.e. tet4 x 4 points for a mixed tet deck.
//
// Correctness-first fallback: one thread owns one element and walks its integration points
// in order (the order of the reference's scan, cmad/fem/assembly.py:477-535, so the sums are
// bit-reproducible); the local solve is the generic 7x7 Newton of the block's yield surface
// (same iterates as the reference's loop; no radial-return specialisation, no deferral);
// K_e is accumulated in place in the caller's buffer (first point stores, later points add).
// Not tuned: n_ip passes over K_e instead of one - the default rules are the bench paths.
#include "fe_common.cuh"

namespace cmadx {
namespace {

// LIST: the elements a tuned first pass handed back (bail list), one thread per list entry
// YKX >= 0: primal solve with the generic Newton of yield surface YKX, for surfaces outside the
// SOLVER numbering shared with the tuned kernels (Yld2004-18p: this kernel is its only element kernel)
template <int SOLVER, bool ROT, bool WANT_K, int NB, bool LIST = false, int YKX = -1>
__global__ void __launch_bounds__(FE_BLOCK) fe_generic_kernel(const __grid_constant__ FeArgs A) {
    const cmadx_fe_block_t& b = A.b;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (LIST) {
        const unsigned cnt = *A.bail_count;
        const bool all = cnt > A.bail_cap;
        if (e >= (all ? b.n_elems : (int64_t)cnt)) return;
        if (!all) e = A.bail_list[e];
    }
    if (e >= b.n_elems) return;
    const int nip = b.n_ip;
    constexpr int ND = NB * 3;
    int eq[ND];
    double U[NB][3], dUe[NB][3];
#pragma unroll
    for (int q = 0; q < ND; ++q) eq[q] = __ldg(b.elem_eq + e * ND + q);
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            U[a][k] = __ldg(b.U + eq[3 * a + k]);
            dUe[a][k] = (SOLVER >= FE_JVP && A.dU) ? __ldg(A.dU + eq[3 * a + k]) : 0.0;
        }
    double R[NB][3];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) R[a][i] = 0.0;

    for (int ip = 0; ip < nip; ++ip) {
        const int64_t p = e * nip + ip;
        double gN[NB][3], xp[7];
        const double* g = b.grad_N + p * ND;
#pragma unroll
        for (int q = 0; q < ND / 4; ++q)
            ld256(g + 4 * q, (&gN[0][0])[4 * q], (&gN[0][0])[4 * q + 1], (&gN[0][0])[4 * q + 2], (&gN[0][0])[4 * q + 3]);
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + p * 7 + c);
        const double wdv = __ldg(b.quad_w + ip) * __ldg(b.det + p);
        double eps[6];
        strain_from_U<NB>(U, gN, eps);
        PointOut o;
        double D[6][6];
        if constexpr (SOLVER >= FE_JVP) {
            double xs[7], dxp[7], de[6];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                xs[c] = __ldg(A.xi_state + p * 7 + c);
                dxp[c] = A.dxi_prev ? __ldg(A.dxi_prev + p * 7 + c) : 0.0;
            }
            strain_from_U<NB>(dUe, gN, de);
            point_jvp<SOLVER - FE_JVP, ROT>(A, xp, xs, dxp, eps, de, true, o);
        } else {
            DevNewton nw = A.nw;
            nw.defer_after = 0;
            if constexpr (YKX >= 0) point_generic<YKX, ROT, WANT_K>(A.m, nw, xp, eps, true, o, D);
            else solve_point<SOLVER, ROT, WANT_K>(A.m, nw, xp, eps, true, o, D);
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) b.xi[p * 7 + c] = o.x[c];
        if (b.iters) b.iters[p] = o.iters;
        if (b.flags) b.flags[p] = o.flags;
        if (b.sigma) {
#pragma unroll
            for (int a = 0; a < 6; ++a) b.sigma[p * 6 + a] = o.sg[a];
        }
        if (A.mix_eq_p) {
            const double* pv = (SOLVER >= FE_JVP) ? A.dU : b.U;
            double pr = 0.0;
            if (pv) {
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    pr = fma(__ldg(A.mix_N + ip * NB + a), __ldg(pv + __ldg(A.mix_eq_p + e * NB + a)), pr);
            }
            mixed_momentum_stress<WANT_K>(pr, o.sg, D);
        }
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], o.sg[vix(j, i)], s);
                R[a][i] = fma(s, wdv, R[a][i]);
            }
        if constexpr (WANT_K) {
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
                    double row[ND];
#pragma unroll
                    for (int bb = 0; bb < NB; ++bb)
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            double s = 0.0;
#pragma unroll
                            for (int l = 0; l < 3; ++l) s = fma(P[vix(k, l)], gN[bb][l], s);
                            row[3 * bb + k] = s;
                        }
                    // whole 32-byte sectors: first point stores the row, later points add to it
                    double* r = Ke + (3 * a + i) * ND;
#pragma unroll
                    for (int q = 0; q < ND / 4; ++q) {
                        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
                        if (ip > 0)
                            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                                         : "=d"(v0), "=d"(v1), "=d"(v2), "=d"(v3) : "l"(r + 4 * q) : "memory");
                        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(r + 4 * q), "d"(v0 + row[4 * q]),
                                     "d"(v1 + row[4 * q + 1]), "d"(v2 + row[4 * q + 2]), "d"(v3 + row[4 * q + 3])
                                     : "memory");
                    }
                }
        }
    }
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
}

template <int NB>
struct GenericLauncher {
    template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
    struct L {
        static cudaError_t run(const FeArgs& A, cudaStream_t stream, int) {
            const int64_t nblk = (A.b.n_elems + FE_BLOCK - 1) / FE_BLOCK;
            fe_generic_kernel<(SOLVER == 0 ? 1 : SOLVER), ROT, WANT_K, NB><<<(unsigned)nblk, FE_BLOCK, 0, stream>>>(A);
            return cudaGetLastError();
        }
    };
};

template <int NB>
struct GenericListLauncher {
    template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
    struct L {
        static cudaError_t run(const FeArgs& A, cudaStream_t stream, int) {
            // sized for the whole block (the count lives on the device); surplus threads exit at once
            const int64_t nblk = (A.b.n_elems + FE_BLOCK - 1) / FE_BLOCK;
            fe_generic_kernel<SOLVER, ROT, WANT_K, NB, true><<<(unsigned)nblk, FE_BLOCK, 0, stream>>>(A);
            return cudaGetLastError();
        }
    };
};

}  // namespace

// second pass over a bail list with the generic solver of the block's yield surface
cudaError_t launch_fe_generic_list(const FeArgs& A, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    const int solver = 1 + A.m.yield;
    if (A.b.n_basis == 4) return dispatch_fe_list<GenericListLauncher<4>::L>(A, solver, stream, 0);
    return dispatch_fe_list<GenericListLauncher<8>::L>(A, solver, stream, 0);
}

// K3 / K4 of a Yld2004-18p block, any rule (the default rules included)
template <int NB>
static cudaError_t launch_barlat_nb(const FeArgs& A, cudaStream_t stream) {
    constexpr int B = CMADX_YIELD_BARLAT;
    const unsigned nblk = (unsigned)((A.b.n_elems + FE_BLOCK - 1) / FE_BLOCK);
    const bool k = A.b.K_elem != nullptr;
    if (A.m.rot) {
        if (k) fe_generic_kernel<1, true, true, NB, false, B><<<nblk, FE_BLOCK, 0, stream>>>(A);
        else fe_generic_kernel<1, true, false, NB, false, B><<<nblk, FE_BLOCK, 0, stream>>>(A);
    } else {
        if (k) fe_generic_kernel<1, false, true, NB, false, B><<<nblk, FE_BLOCK, 0, stream>>>(A);
        else fe_generic_kernel<1, false, false, NB, false, B><<<nblk, FE_BLOCK, 0, stream>>>(A);
    }
    return cudaGetLastError();
}
cudaError_t launch_fe_generic_barlat(const FeArgs& A, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    return (A.b.n_basis == 4) ? launch_barlat_nb<4>(A, stream) : launch_barlat_nb<8>(A, stream);
}

// solver: 1 + yield (primal, generic Newton) or 4 + yield (JVP)
cudaError_t launch_fe_generic(const FeArgs& A, int solver, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    if (A.b.n_basis == 4) return dispatch_fe<GenericLauncher<4>::L, false>(A, solver, stream, 0);
    return dispatch_fe<GenericLauncher<8>::L, false>(A, solver, stream, 0);
}

}  // namespace cmadx
