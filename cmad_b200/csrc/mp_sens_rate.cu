// K2 for SmallRateElasticPlastic: the calibration objective and its gradient over whole load
// histories, adjoint (cmad/objectives/mp_objective.py:92-147) and direct (:150-215), with the
// Calibration QoI (cmad/qois/calibration.py:56-66).  Same organisation as mp_sens.cu - one thread
// per material point walks its stored history in registers, fixed-order block partials, final
// reduction by reduce_partials_kernel - with the rate model's blocks (rate_point.cuh):
//   state x = [cauchy(6), alpha], de_t = eps_t - eps_{t-1} (the history carries total strains),
//   A = dC/dxi (RatePoint::jacobian), B = dC/dxi_prev = [-I/2mu, -n (plastic); 0, -1 (elastic)],
//   the QoI reads the state's own stress: dJ/dxi = [r, 0], dJ/dp = 0.
// FULL_3D, identity material axes.
#include "mp_sens.cuh"
#include "rate_point.cuh"

namespace cmadx {

cudaError_t launch_reduce_partials(const double* partials, int64_t nblk, int ncols, double* result,
                                   cudaStream_t stream);
int64_t sens_blocks(int64_t n);

namespace {

constexpr int SENS_BLOCK = 128;      // = mp_sens.cu's block (sens_blocks() sizes the partials)

template <int YK, bool ADJOINT>
__global__ void __launch_bounds__(SENS_BLOCK) mp_sens_rate_kernel(const __grid_constant__ SensArgs A) {
    constexpr int NA_MAX = CMADX_MAX_ACTIVE;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int N = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;
    const int comp[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};      // tensor entry -> packed component

    double g[NA_MAX], X[NA_MAX][7];
#pragma unroll
    for (int c = 0; c < NA_MAX; ++c) {
        g[c] = 0.0;
#pragma unroll
        for (int r = 0; r < 7; ++r) X[c][r] = 0.0;
    }
    double Jacc = 0.0, hist[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    // rotated material axes: de_m = T de, the QoI reads the global stress S sig (rate_point.cuh)
    const bool rot = m.rot != 0;
    double T[6][6], S[6][6];
    if (rot) rot_maps(m.Q, T, S);
    for (int s = 0; s < N; ++s) {
        const int t = ADJOINT ? N - s : s + 1;
        double x[7], xp[7], de[6], d[9];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)t * 7 * ld + i;
#pragma unroll
            for (int c = 0; c < 7; ++c) { x[c] = __ldg(xs + c * ld); xp[c] = __ldg(xs + (c - 7) * ld); }
            double ep[6];
            rate_load_strain(A.h.strain + (int64_t)t * sc * ld, sc, ld, i, de);
            rate_load_strain(A.h.strain + (int64_t)(t - 1) * sc * ld, sc, ld, i, ep);
#pragma unroll
            for (int c = 0; c < 6; ++c) de[c] -= ep[c];
            if (rot) {
                double dm[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < 6; ++b) s = fma(T[c][b], de[b], s);
                    dm[c] = s;
                }
#pragma unroll
                for (int c = 0; c < 6; ++c) de[c] = dm[c];
            }
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < 7; ++c) { x[c] = (c == 0) ? 1.0 : 0.0; xp[c] = 0.0; }
#pragma unroll
            for (int c = 0; c < 6; ++c) de[c] = 1e-3 * (c == 0);
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
        }
        RatePoint<YK> pt;
        double C[7];
        pt.residual(m, x, xp, de, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        // Calibration QoI on the state's stress
        double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        double sgl[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            sgl[a] = x[a];
            if (rot) {
                sgl[a] = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) sgl[a] = fma(S[a][c], x[c], sgl[a]);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double mis = A.h.weight[k] * (sgl[comp[k]] - d[k]);
            Jacc = fma(0.5 * mis, mis, Jacc);
            r[comp[k]] = fma(A.h.weight[k], mis, r[comp[k]]);
        }
        if (rot) {               // dJ/dsig_m = S^T dJ/dsig_g
            double rm[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < 6; ++a) s = fma(S[a][c], r[a], s);
                rm[c] = s;
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) r[c] = rm[c];
        }
        RegLU<7> lu;
        auto load = [&]() {
            if (ADJOINT) {
                double Jm[7][7];
                pt.jacobian(m, dg, Jm);
#pragma unroll
                for (int a = 0; a < 7; ++a)
#pragma unroll
                    for (int b = 0; b < 7; ++b) lu.a[a][b] = Jm[b][a];
            } else {
                pt.jacobian(m, dg, lu.a);
            }
        };
        load();
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) { load(); lu.factor_pivot(); }
        auto solve7 = [&](double (&v)[7]) { if (slow && trouble) lu.solve_pivot(v); else lu.solve_natural(v); };
        if (ADJOINT) {
            double phi[7];
#pragma unroll
            for (int c = 0; c < 6; ++c) phi[c] = hist[c] - r[c];
            phi[6] = hist[6];
            solve7(phi);
            if (A.phi_hist && live) {      // kept for the direct-adjoint Hessian pass (mp_hess.cu)
                double* ph = A.phi_hist + (int64_t)t * 7 * ld + i;
#pragma unroll
                for (int c = 0; c < 7; ++c) ph[c * ld] = phi[c];
            }
            // h <- -B^T phi
            double nphi = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) { hist[a] = m.inv_two_mu * phi[a]; nphi = fma(pt.n[a], phi[a], nphi); }
            hist[6] = pl ? nphi : phi[6];
            for (int c = 0; c < na; ++c) {
                double col[7], acc = 0.0;
                rate_dC_dp_column<YK>(m, A.pid[c], pt, x, xp, de, col);
#pragma unroll
                for (int q = 0; q < 7; ++q) acc = fma(phi[q], col[q], acc);
                g[c] += acc;
            }
        } else {
            for (int c = 0; c < na; ++c) {
                double col[7], rhs[7];
                rate_dC_dp_column<YK>(m, A.pid[c], pt, x, xp, de, col);
                const double x6 = X[c][6];
#pragma unroll
                for (int q = 0; q < 6; ++q) rhs[q] = -col[q] + m.inv_two_mu * X[c][q] + (pl ? pt.n[q] * x6 : 0.0);
                rhs[6] = -col[6] + (pl ? 0.0 : x6);
                solve7(rhs);
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 7; ++q) { X[c][q] = rhs[q]; if (q < 6) acc = fma(r[q], rhs[q], acc); }
                g[c] += acc;
            }
        }
    }
    if (!live) {
        Jacc = 0.0;
#pragma unroll
        for (int c = 0; c < NA_MAX; ++c) g[c] = 0.0;
    } else if (A.h.J_point) {
        A.h.J_point[i] = Jacc;
    }
    __shared__ double sm[SENS_BLOCK / 32][1 + CMADX_MAX_ACTIVE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c <= na; ++c) {
        double v = (c == 0) ? Jacc : g[c - 1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x <= na) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < SENS_BLOCK / 32; ++w) v += sm[w][threadIdx.x];
        A.partials[(int64_t)blockIdx.x * (1 + na) + threadIdx.x] = v;
    }
}

template <bool ADJOINT>
cudaError_t launch_t(const SensArgs& A, cudaStream_t stream) {
    const unsigned nblk = (unsigned)sens_blocks(A.h.n);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: mp_sens_rate_kernel<CMADX_YIELD_J2, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HILL: mp_sens_rate_kernel<CMADX_YIELD_HILL, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HOSFORD: mp_sens_rate_kernel<CMADX_YIELD_HOSFORD, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_BARLAT: mp_sens_rate_kernel<CMADX_YIELD_BARLAT, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_reduce_partials(A.partials, nblk, 1 + A.n_active, A.h.result, stream);
}

}  // namespace

cudaError_t launch_mp_sens_rate(const SensArgs& A, bool adjoint, cudaStream_t stream) {
    if (A.h.n == 0) return cudaMemsetAsync(A.h.result, 0, sizeof(double) * (1 + A.n_active), stream);
    return adjoint ? launch_t<true>(A, stream) : launch_t<false>(A, stream);
}

}  // namespace cmadx
