// K3 / K4 for tet4 elements with the degree-1 rule (1 integration point): one
// thread owns one element; see fe_block.cu for the contract and the references.
#include "fe_common.cuh"

namespace cmadx {
namespace {

// ==================================================================== tet4, 1 IP
template <int SOLVER, bool ROT, bool WANT_K>
CMADX_DEV void tet4_element(const FeArgs& A, const int64_t e, const bool live, const bool allow_defer) {
    const cmadx_fe_block_t& b = A.b;
    double gN[4][3], U[4][3], xp[7];
    int eq[12];
    double wdv = 0.0;
    if (live) {
        const int4* q = reinterpret_cast<const int4*>(b.elem_eq + e * 12);
        const int4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        eq[0] = q0.x; eq[1] = q0.y; eq[2] = q0.z; eq[3] = q0.w;
        eq[4] = q1.x; eq[5] = q1.y; eq[6] = q1.z; eq[7] = q1.w;
        eq[8] = q2.x; eq[9] = q2.y; eq[10] = q2.z; eq[11] = q2.w;
        const double* g = b.grad_N + e * 12;
        ld256(g, gN[0][0], gN[0][1], gN[0][2], gN[1][0]);
        ld256(g + 4, gN[1][1], gN[1][2], gN[2][0], gN[2][1]);
        ld256(g + 8, gN[2][2], gN[3][0], gN[3][1], gN[3][2]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) U[a][k] = __ldg(b.U + eq[3 * a + k]);
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + e * 7 + c);
        wdv = __ldg(b.quad_w) * __ldg(b.det + e);
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) { U[a][k] = 0.0; gN[a][k] = 0.0; eq[3 * a + k] = 0; }
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = 0.0;
    }
    double eps[6];
    strain_from_U<4>(U, gN, eps);

    PointOut o;
    double D[6][6];
    if constexpr (SOLVER >= FE_JVP) {
        double xs[7], dxp[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            xs[c] = live ? __ldg(A.xi_state + e * 7 + c) : 0.0;
            dxp[c] = (live && A.dxi_prev) ? __ldg(A.dxi_prev + e * 7 + c) : 0.0;
        }
        double de[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (live && A.dU) {
            double dUe[4][3];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int k = 0; k < 3; ++k) dUe[a][k] = __ldg(A.dU + eq[3 * a + k]);
            strain_from_U<4>(dUe, gN, de);
        }
        point_jvp<SOLVER - FE_JVP, ROT>(A, xp, xs, dxp, eps, de, live, o);
    } else {
        DevNewton nw = A.nw;
        nw.defer_after = allow_defer ? A.nw.defer_request : 0;
        solve_point<SOLVER, ROT, WANT_K>(A.m, nw, xp, eps, live, o, D);
    }
    if (SOLVER < FE_JVP && A.bail_count)
        list_append(live && o.bail, A.bail_count, A.bail_list, A.bail_cap, (int)e);
    if (!live || (SOLVER < FE_JVP && o.bail)) return;

#pragma unroll
    for (int c = 0; c < 7; ++c) b.xi[e * 7 + c] = o.x[c];
    if (b.iters) b.iters[e] = o.iters;
    if (b.flags) b.flags[e] = o.flags;
    if (b.sigma) {
#pragma unroll
        for (int a = 0; a < 6; ++a) b.sigma[e * 6 + a] = o.sg[a];
    }
    if (A.mix_eq_p) {
        // primal: p from U; K6: the momentum-stress direction dev(d cauchy) - dp I, dp from dU
        const double* pv = (SOLVER >= FE_JVP) ? A.dU : b.U;
        double p = 0.0;
        if (pv) {
            const int4 qp = __ldg(reinterpret_cast<const int4*>(A.mix_eq_p + e * 4));
            p = fma(__ldg(A.mix_N + 3), __ldg(pv + qp.w),
                    fma(__ldg(A.mix_N + 2), __ldg(pv + qp.z),
                        fma(__ldg(A.mix_N + 1), __ldg(pv + qp.y), __ldg(A.mix_N) * __ldg(pv + qp.x))));
        }
        mixed_momentum_stress<WANT_K>(p, o.sg, D);
    }
    // R[a][i] = sum_j gN[a][j] sigma[j][i] w dv
    if (b.R_elem || b.R_global) {
        double R[4][3];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], o.sg[vix(j, i)], s);
                R[a][i] = s * wdv;
            }
        if (b.R_elem) {
            double* r = b.R_elem + e * 12;
            st256(r, R[0][0], R[0][1], R[0][2], R[1][0]);
            st256(r + 4, R[1][1], R[1][2], R[2][0], R[2][1]);
            st256(r + 8, R[2][2], R[3][0], R[3][1], R[3][2]);
        }
        if (b.R_global) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int i = 0; i < 3; ++i) atomicAdd(b.R_global + eq[3 * a + i], R[a][i]);
        }
    }
    if constexpr (WANT_K) {
    // Dh = D (k==l ? 1 : 1/2) w dv
#pragma unroll
    for (int al = 0; al < 6; ++al)
#pragma unroll
        for (int be = 0; be < 6; ++be) D[al][be] *= is_diag(be) ? wdv : 0.5 * wdv;
    double* Ke = b.K_elem + e * 144;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double P[6];
#pragma unroll
            for (int be = 0; be < 6; ++be) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], D[vix(j, i)][be], s);
                P[be] = s;
            }
            double row[12];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < 3; ++l) s = fma(P[vix(k, l)], gN[bb][l], s);
                    row[3 * bb + k] = s;
                }
            double* r = Ke + (3 * a + i) * 12;
            st256(r, row[0], row[1], row[2], row[3]);
            st256(r + 4, row[4], row[5], row[6], row[7]);
            st256(r + 8, row[8], row[9], row[10], row[11]);
        }
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
__global__ void __launch_bounds__(FE_BLOCK, (SOLVER == 0) ? (WANT_K ? 3 : 4) : 1) fe_tet4_kernel(const __grid_constant__ FeArgs A) {
    if (!LIST) {
        const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        tet4_element<SOLVER, ROT, WANT_K>(A, e, e < A.b.n_elems, A.bail_count != nullptr);
    } else {
        // list mode: a small grid walks the elements the J2 kernel handed back
        const unsigned cnt = *A.bail_count;
        if (cnt == 0u) return;
        const bool all = cnt > A.bail_cap;
        const int64_t total = all ? A.b.n_elems : (int64_t)cnt;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        const int lane = threadIdx.x & 31;
        for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < total;
             base += stride) {
            const int64_t j = base + lane;
            const bool live = j < total;
            const int64_t e = live ? (all ? j : (int64_t)A.bail_list[j]) : 0;
            tet4_element<SOLVER, ROT, WANT_K>(A, e, live, false);
        }
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
struct Tet4Launcher {
    static cudaError_t run(const FeArgs& A, cudaStream_t stream, int sms) {
        const int64_t nblk = LIST ? 2 * sms : (A.b.n_elems + FE_BLOCK - 1) / FE_BLOCK;
        fe_tet4_kernel<SOLVER, ROT, WANT_K, LIST><<<(unsigned)nblk, FE_BLOCK, 0, stream>>>(A);
        return cudaGetLastError();
    }
};

}  // namespace

cudaError_t launch_fe_tet4(const FeArgs& A, int solver, bool list, cudaStream_t stream, int sms) {
    // list mode: the generic solver over the elements handed back / deferred by the first pass
    if (list) return dispatch_fe_list<Tet4Launcher>(A, solver, stream, sms);
    return dispatch_fe<Tet4Launcher, false>(A, solver, stream, sms);
}

}  // namespace cmadx
