// K3 / K4 for tet4 elements with the 4-point degree-2 rule - the rule the reference's deck
// driver forces on the mixed u-p formulation (cmad/cli/common.py:379-391) and the first
// `volume degree` override a tet deck would use.  Four lanes own one element, one integration
// point each (8 elements per warp): every lane gathers U_e, interpolates grad_u at its point,
// runs the local Newton there (J2 radial return with the element-level hand-back list, or the
// generic solver), and the element sums - R_e and the twelve rows of K_e - are butterfly
// reductions over the 4 lanes in fixed order (bit-reproducible; the same association for every
// element).  Row r of K_e ends up on lane r mod 4, so after every group of four rows each lane
// streams one 96-byte row out as three 256-bit stores: K_e is written exactly once (the
// any-rule kernel of fe_generic.cu makes n_ip passes over it).
#include <cstdlib>

#include "fe_common.cuh"

namespace cmadx {
namespace {

CMADX_DEV double quad_sum(double v) {          // sum over the 4 lanes of an element
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// Reduce-scatter over the 4 lanes of an element: every lane contributes v[0..3], lane l returns
// sum over the lanes of v[l] - 3 shuffles for 4 sums (the butterfly all-reduce of each value takes
// 8), with the butterfly's association ((v_l + v_{l^1}) + (v_{l^2} + v_{l^3})): bit-identical.
CMADX_DEV double quad_reduce_scatter(const double (&v)[4], const int ip) {
    const bool b0 = ip & 1, b1 = ip & 2;
    double ka = b0 ? v[1] : v[0], kb = b0 ? v[3] : v[2];
    ka += __shfl_xor_sync(0xffffffffu, b0 ? v[0] : v[1], 1);
    kb += __shfl_xor_sync(0xffffffffu, b0 ? v[2] : v[3], 1);
    const double keep = b1 ? kb : ka;
    return keep + __shfl_xor_sync(0xffffffffu, b1 ? ka : kb, 2);
}

// FACTORED (the J2 radial kernel, which always runs with the element hand-back list): only the
// factored path below is compiled in - an element whose grad_N is not the same at its 4 points is
// handed to the any-rule kernel like a radial-return bail - so the kernel stays lean.  Measured
// (3.07 M tets, B200): per-point products, rolled row groups 2.93 ms, unrolled 2.69 ms
// (profiles/r2p_fe.jsonl); factored: see DESIGN.md.  CMADX_TET4X4_GENERAL=1 selects the per-point kernel.
template <int SOLVER, bool ROT, bool WANT_K, bool FACTORED = false>
__global__ void __launch_bounds__(FE_BLOCK, (SOLVER == 0) ? (FACTORED ? 4 : 3) : 1) fe_tet4x4_kernel(const __grid_constant__ FeArgs A) {
    const cmadx_fe_block_t& b = A.b;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t >> 2;
    const int ip = (int)(t & 3);
    const bool live = e < b.n_elems;
    const int64_t el = live ? e : 0;
    const int64_t p = el * 4 + ip;
    double gN[4][3], U[4][3], xp[7];
    int eq[12];
    {
        const int4* q = reinterpret_cast<const int4*>(b.elem_eq + el * 12);
        const int4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        eq[0] = q0.x; eq[1] = q0.y; eq[2] = q0.z; eq[3] = q0.w;
        eq[4] = q1.x; eq[5] = q1.y; eq[6] = q1.z; eq[7] = q1.w;
        eq[8] = q2.x; eq[9] = q2.y; eq[10] = q2.z; eq[11] = q2.w;
        const double* g = b.grad_N + p * 12;
        ld256(g, gN[0][0], gN[0][1], gN[0][2], gN[1][0]);
        ld256(g + 4, gN[1][1], gN[1][2], gN[2][0], gN[2][1]);
        ld256(g + 8, gN[2][2], gN[3][0], gN[3][1], gN[3][2]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) U[a][k] = __ldg(b.U + eq[3 * a + k]);
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + p * 7 + c);
    }
    const double wdv = live ? __ldg(b.quad_w + ip) * __ldg(b.det + p) : 0.0;
    double eps[6];
    strain_from_U<4>(U, gN, eps);
    if (!live) { eps[0] = 1e-3; eps[1] = eps[2] = eps[3] = eps[4] = eps[5] = 0.0; }

    PointOut o;
    double D[6][6];
    DevNewton nw = A.nw;
    nw.defer_after = 0;
    solve_point<SOLVER, ROT, WANT_K>(A.m, nw, xp, eps, live, o, D);

    // an element is handed to the generic second pass as a whole
    // one grad_N per element?  (bit-equal across the element's 4 points)
    bool uniform = true;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            uniform = uniform && (gN[a][k] == __shfl_sync(0xffffffffu, gN[a][k], (threadIdx.x & 31) & ~3));
    bool ebail = false;
    if (SOLVER == 0) {
        const unsigned bal = __ballot_sync(0xffffffffu, live && (o.bail || (FACTORED && !uniform)));
        ebail = ((bal >> ((threadIdx.x & 31) & ~3)) & 0xfu) != 0u;
        if (ebail && live && ip == 0 && A.bail_count) append_bail(A, e);
    }
    const bool emit = live && !ebail;
    if (emit) {
#pragma unroll
        for (int c = 0; c < 7; ++c) b.xi[p * 7 + c] = o.x[c];
        if (b.iters) b.iters[p] = o.iters;
        if (b.flags) b.flags[p] = o.flags;
        if (b.sigma) {
#pragma unroll
            for (int a = 0; a < 6; ++a) b.sigma[p * 6 + a] = o.sg[a];
        }
    }
    if (A.mix_eq_p) {
        const int4 qp = __ldg(reinterpret_cast<const int4*>(A.mix_eq_p + el * 4));
        const double* Np = A.mix_N + ip * 4;
        const double pr = fma(__ldg(Np + 3), __ldg(b.U + qp.w),
                              fma(__ldg(Np + 2), __ldg(b.U + qp.z),
                                  fma(__ldg(Np + 1), __ldg(b.U + qp.y), __ldg(Np) * __ldg(b.U + qp.x))));
        mixed_momentum_stress<WANT_K>(pr, o.sg, D);
    }
    // ---- Linear tets have ONE grad_N per element (constant Jacobian): the per-point sums then factor,
    //   R_e = B^T (sum_ip sigma w dv),   K_e = B^T (sum_ip Dh w dv) B,
    // so the 4 lanes all-reduce sigma w dv (6 values) and Dh w dv (36 values) and lane a forms rows
    // 3a..3a+2 of R_e / K_e from the sums: a quarter of the FP64 work of the per-point products and
    // no reduction of the 144 K_e entries (ncu on the per-point version: 3072 instructions per warp,
    // 38 % of the issue slots at 8 warps / SM - latency-bound).  The kernel checks the premise on the
    // data it was given (bit-equal grad_N across the 4 points of every element of the warp) and keeps
    // the general per-point path otherwise.
    if (FACTORED || __all_sync(0xffffffffu, uniform || !live)) {
        double ga[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) ga[j] = (ip == 0) ? gN[0][j] : ((ip == 1) ? gN[1][j] : ((ip == 2) ? gN[2][j] : gN[3][j]));
        if (b.R_elem || b.R_global) {
            double sw[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) sw[a] = quad_sum(o.sg[a] * wdv);
            if (emit) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double r = fma(ga[2], sw[vix(2, i)], fma(ga[1], sw[vix(1, i)], ga[0] * sw[vix(0, i)]));
                    if (b.R_elem) b.R_elem[e * 12 + 3 * ip + i] = r;
                    if (b.R_global) atomicAdd(b.R_global + eq[3 * ip + i], r);
                }
            }
        }
        if constexpr (WANT_K) {
#pragma unroll
            for (int al = 0; al < 6; ++al)
#pragma unroll
                for (int be = 0; be < 6; ++be) D[al][be] = quad_sum(D[al][be] * (is_diag(be) ? wdv : 0.5 * wdv));
            double* Ke = b.K_elem + el * 144 + 36 * ip;          // rows 3 ip .. 3 ip + 2
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double P[6], row[12];
#pragma unroll
                for (int be = 0; be < 6; ++be)
                    P[be] = fma(ga[2], D[vix(2, i)][be], fma(ga[1], D[vix(1, i)][be], ga[0] * D[vix(0, i)][be]));
#pragma unroll
                for (int bb = 0; bb < 4; ++bb)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        row[3 * bb + k] = fma(P[vix(k, 2)], gN[bb][2], fma(P[vix(k, 1)], gN[bb][1], P[vix(k, 0)] * gN[bb][0]));
                if (emit) {
                    st256(Ke + 12 * i, row[0], row[1], row[2], row[3]);
                    st256(Ke + 12 * i + 4, row[4], row[5], row[6], row[7]);
                    st256(Ke + 12 * i + 8, row[8], row[9], row[10], row[11]);
                }
            }
        }
        return;
    }
    if constexpr (!FACTORED) {
    // ---- R_e: this point's contribution, summed over the element's 4 lanes; lane a keeps node a
    if (b.R_elem || b.R_global) {
        double mine[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double v[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], o.sg[vix(j, i)], s);
                v[a] = s * wdv;
            }
            mine[i] = quad_reduce_scatter(v, ip);
        }
        if (emit) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (b.R_elem) b.R_elem[e * 12 + 3 * ip + i] = mine[i];
                if (b.R_global) atomicAdd(b.R_global + eq[3 * ip + i], mine[i]);
            }
        }
    }
    if constexpr (WANT_K) {
#pragma unroll
        for (int al = 0; al < 6; ++al)
#pragma unroll
            for (int be = 0; be < 6; ++be) D[al][be] *= is_diag(be) ? wdv : 0.5 * wdv;
        double* Ke = b.K_elem + el * 144;
#pragma unroll 1
        for (int g4 = 0; g4 < 3; ++g4) {             // rows 4 g4 .. 4 g4 + 3
            double keep[12], P[4][6];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = 4 * g4 + rr, a = r / 3, i = r - 3 * a;
#pragma unroll
                for (int be = 0; be < 6; ++be) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < 3; ++j) s = fma(gN[a][j], D[vix(j, i)][be], s);
                    P[rr][be] = s;
                }
            }
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double v[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        double s = 0.0;
#pragma unroll
                        for (int l = 0; l < 3; ++l) s = fma(P[rr][vix(k, l)], gN[bb][l], s);
                        v[rr] = s;
                    }
                    keep[3 * bb + k] = quad_reduce_scatter(v, ip);
                }
            if (emit) {
                double* r = Ke + (4 * g4 + ip) * 12;
                st256(r, keep[0], keep[1], keep[2], keep[3]);
                st256(r + 4, keep[4], keep[5], keep[6], keep[7]);
                st256(r + 8, keep[8], keep[9], keep[10], keep[11]);
            }
        }
    }
    }   // !FACTORED
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
struct Tet4x4Launcher {
    static cudaError_t run(const FeArgs& A, cudaStream_t stream, int) {
        if constexpr (SOLVER >= FE_JVP) {
            return cudaErrorInvalidValue;            // K6-JVP of this rule runs the any-rule kernel
        } else {
            const int64_t nthr = A.b.n_elems * 4;
            const int64_t nblk = (nthr + FE_BLOCK - 1) / FE_BLOCK;
            static const bool factored = std::getenv("CMADX_TET4X4_GENERAL") == nullptr;
            if (SOLVER == 0 && !ROT && factored && A.bail_count) fe_tet4x4_kernel<SOLVER, ROT, WANT_K, SOLVER == 0 && !ROT><<<(unsigned)nblk, FE_BLOCK, 0, stream>>>(A);
            else fe_tet4x4_kernel<SOLVER, ROT, WANT_K><<<(unsigned)nblk, FE_BLOCK, 0, stream>>>(A);
            return cudaGetLastError();
        }
    }
};

}  // namespace

// solver: 0 = J2 radial return (hand-backs go to the any-rule kernel in list mode), 1 + yield
cudaError_t launch_fe_tet4x4(const FeArgs& A, int solver, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    return dispatch_fe<Tet4x4Launcher, false>(A, solver, stream, 0);
}

}  // namespace cmadx
