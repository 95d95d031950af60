// K3 / K4 for SmallRateElasticPlastic blocks (cmad/models/small_rate_elastic_plastic.py:250-346
// under assemble_element_block, cmad/fem/assembly.py:616-732): the per-point residual sees the
// strain INCREMENT eps(U) - eps(U_prev), so the block carries both displacement vectors
// (cmadx_fe_block_t::U_prev).  State per point = [cauchy(6), alpha]; the cauchy stress handed to
// the momentum residual is the state's own stress (_cauchy_fn :351-359, Q = I) and the consistent
// tangent is the stress rows of the IFT sensitivity, d cauchy / d eps = (dxi/d eps)[0:6]
// (nonlinear_solver.py:158-171).  Any volume rule of tet4 / hex8, displacement form (the mixed
// u-p form of this model has state-dependent pressure rows, :361-376 - not carried).  Correctness-first organisation of fe_generic.cu: one
// thread per element walks its points in the reference's scan order (bit-reproducible sums),
// K_e accumulated in place.  Not a bench path.
#include "fe_common.cuh"
#include "rate_point.cuh"

namespace cmadx {
namespace {

template <int YK, bool WANT_K, int NB>
__global__ void __launch_bounds__(FE_BLOCK) fe_rate_kernel(const __grid_constant__ FeArgs A) {
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
        RatePoint<YK> pt;
        double C[7];
        const NewtonResult nr = local_newton<RatePoint<YK>, 7>(m, nw, pt, x, xp, de, live, C);
        double sg[6], D[6][6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sg[a] = x[a];
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
}

template <int YK, int NB>
cudaError_t run(const FeArgs& A, cudaStream_t stream) {
    const unsigned nblk = (unsigned)((A.b.n_elems + FE_BLOCK - 1) / FE_BLOCK);
    if (A.b.K_elem) fe_rate_kernel<YK, true, NB><<<nblk, FE_BLOCK, 0, stream>>>(A);
    else fe_rate_kernel<YK, false, NB><<<nblk, FE_BLOCK, 0, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fe_rate(const FeArgs& A, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    const bool tet = A.b.n_basis == 4;
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return tet ? run<CMADX_YIELD_J2, 4>(A, stream) : run<CMADX_YIELD_J2, 8>(A, stream);
    case CMADX_YIELD_HILL: return tet ? run<CMADX_YIELD_HILL, 4>(A, stream) : run<CMADX_YIELD_HILL, 8>(A, stream);
    case CMADX_YIELD_HOSFORD: return tet ? run<CMADX_YIELD_HOSFORD, 4>(A, stream) : run<CMADX_YIELD_HOSFORD, 8>(A, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cmadx
