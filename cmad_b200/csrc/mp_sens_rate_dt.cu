// K2 for SmallRateElasticPlastic under the PLANE_STRESS / UNIAXIAL_STRESS deformation types: the
// calibration objective and its gradient over stored load histories, adjoint
// (cmad/objectives/mp_objective.py:92-147) and direct (:150-215), Calibration QoI
// (cmad/qois/calibration.py:56-66) on the global stress Q sig Q^T - the setting of the reference's
// tests/objectives/test_J2_fd_checks.py (plane stress, both small-strain models) and
// tests/objectives/test_calibrations.py.  Organisation of mp_sens_rate.cu (one thread walks its
// point's history, register LU of A = dC/dxi - transposed for the adjoint -, fixed-order block
// partials), on the bordered systems of rate_point_dt.cuh (n_xi = 8 | 12):
//   B = dC/dxi_prev from A (rate_dt_B),  dC/dp with the constraint rows (rate_dt_dC_dp_column),
//   dJ/dxi = [S^T r, 0, ...]: the QoI reads the state's own stress; dJ/dp = 0.
// The history carries TOTAL prescribed strains; the increment is formed here.
#include "mp_sens.cuh"
#include "rate_point_dt.cuh"

namespace cmadx {

cudaError_t launch_reduce_partials(const double* partials, int64_t nblk, int ncols, double* result,
                                   cudaStream_t stream);
int64_t sens_blocks(int64_t n);

namespace {

constexpr int SENS_BLOCK = 128;      // = mp_sens.cu's block (sens_blocks() sizes the partials)

template <int YK, int DT, bool ADJOINT>
__global__ void __launch_bounds__(SENS_BLOCK) mp_sens_rate_dt_kernel(const __grid_constant__ SensArgs A) {
    using Pt = RatePointDT<YK, DT>;
    constexpr int N = Pt::N, NZ = Pt::NZ, NA_MAX = CMADX_MAX_ACTIVE;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int NT = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;
    const int comp[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};      // tensor entry -> packed component

    double g[NA_MAX], X[NA_MAX][N];
#pragma unroll
    for (int c = 0; c < NA_MAX; ++c) g[c] = 0.0;
    if (!ADJOINT) {
        for (int c = 0; c < na; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) X[c][r] = 0.0;
    }
    double Jacc = 0.0, hist[N];
#pragma unroll
    for (int c = 0; c < N; ++c) hist[c] = 0.0;
    for (int s = 0; s < NT; ++s) {
        const int t = ADJOINT ? NT - s : s + 1;
        double x[N], xp[N], em[6] = {1e-3, 0.0, 0.0, 0.0, 0.0, 0.0}, d[9];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)t * N * ld + i;
#pragma unroll
            for (int c = 0; c < N; ++c) { x[c] = __ldg(xs + c * ld); xp[c] = __ldg(xs + (c - N) * ld); }
            double ep[6];
            load_dt_strain<DT>(A.h.strain + (int64_t)t * sc * ld, sc, ld, i, em);
            load_dt_strain<DT>(A.h.strain + (int64_t)(t - 1) * sc * ld, sc, ld, i, ep);
#pragma unroll
            for (int c = 0; c < 6; ++c) em[c] -= ep[c];
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < N; ++c) { x[c] = (c == 0 || (c >= 7 && c < 7 + NZ)) ? 1.0 : 0.0; xp[c] = (c >= 7 && c < 7 + NZ) ? 1.0 : 0.0; }
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
        }
        Pt pt;
        double C[N];
        pt.residual(m, x, xp, em, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        // Calibration QoI on the global stress S sig; cotangent back to the state through S^T
        double sgl[6], r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, dJdx[N];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) v = fma(pt.S[a][c], x[c], v);
            sgl[a] = v;
        }
        double dJdz[NZ];                   // direct dependence of the QoI on the stretch dofs
#pragma unroll
        for (int k = 0; k < NZ; ++k) dJdz[k] = 0.0;
        if (DT == CMADX_DEF_UNIAXIAL_STRESS && A.h.qoi_kind == CMADX_QOI_UNIAXIAL_CALIBRATION) {
            // UniaxialCalibration (cmad/qois/uniaxial_calibration.py:69-85): pred = [sigma_axial,
            // lambda_2 - 1, lambda_3 - 1], weights of this step, data rows 0..2 (as in mp_sens_dt.cu)
            const double* ws = A.h.weight_steps + (int64_t)t * 3;
            const double w0 = __ldg(ws), mis0 = w0 * (sgl[0] - d[0]);
            Jacc = fma(0.5 * mis0, mis0, Jacc);
            r[0] = w0 * mis0;
#pragma unroll
            for (int k = 0; k < NZ; ++k) {
                const double wk = __ldg(ws + 1 + k), mis = wk * (x[7 + k] - 1.0 - d[1 + k]);
                Jacc = fma(0.5 * mis, mis, Jacc);
                dJdz[k] = wk * mis;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double mis = A.h.weight[k] * (sgl[comp[k]] - d[k]);
                Jacc = fma(0.5 * mis, mis, Jacc);
                r[comp[k]] = fma(A.h.weight[k], mis, r[comp[k]]);
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double v = 0.0;
            if (c < 6) {
#pragma unroll
                for (int a = 0; a < 6; ++a) v = fma(pt.S[a][c], r[a], v);
            } else if (c >= 7 && c < 7 + NZ) {
                v = dJdz[c - 7];
            }
            dJdx[c] = v;
        }
        double Am[N][N];
        pt.jacobian(m, dg, Am);
        RegLU<N> lu;
        auto load = [&]() {
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int b = 0; b < N; ++b) lu.a[a][b] = ADJOINT ? Am[b][a] : Am[a][b];
        };
        load();
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) { load(); lu.factor_pivot(); }
        auto solveN = [&](double (&v)[N]) { if (slow && trouble) lu.solve_pivot(v); else lu.solve_natural(v); };
        if (ADJOINT) {
            double phi[N];
#pragma unroll
            for (int c = 0; c < N; ++c) phi[c] = hist[c] - dJdx[c];
            solveN(phi);
            if (A.phi_hist && live) {      // kept for the direct-adjoint Hessian pass (mp_hess.cu)
                double* ph = A.phi_hist + (int64_t)t * N * ld + i;
#pragma unroll
                for (int c = 0; c < N; ++c) ph[c * ld] = phi[c];
            }
            // h <- -B^T phi
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < N; ++q) v = fma(rate_dt_B(m, pt, Am, q, c), phi[q], v);
                hist[c] = -v;
            }
            for (int c = 0; c < na; ++c) {
                double col[N], acc = 0.0;
                rate_dt_dC_dp_column<YK, DT>(m, A.pid[c], pt, x, xp, col);
#pragma unroll
                for (int q = 0; q < N; ++q) acc = fma(phi[q], col[q], acc);
                g[c] += acc;
            }
        } else {
            for (int c = 0; c < na; ++c) {
                double col[N], rhs[N];
                rate_dt_dC_dp_column<YK, DT>(m, A.pid[c], pt, x, xp, col);
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    double v = -col[q];
#pragma unroll
                    for (int k = 0; k < N; ++k) v = fma(-rate_dt_B(m, pt, Am, q, k), X[c][k], v);
                    rhs[q] = v;
                }
                solveN(rhs);
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; ++q) { X[c][q] = rhs[q]; acc = fma(dJdx[q], rhs[q], acc); }
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

template <int DT, bool ADJOINT>
cudaError_t launch_t(const SensArgs& A, cudaStream_t stream) {
    const unsigned nblk = (unsigned)sens_blocks(A.h.n);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: mp_sens_rate_dt_kernel<CMADX_YIELD_J2, DT, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HILL: mp_sens_rate_dt_kernel<CMADX_YIELD_HILL, DT, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HOSFORD: mp_sens_rate_dt_kernel<CMADX_YIELD_HOSFORD, DT, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_BARLAT: mp_sens_rate_dt_kernel<CMADX_YIELD_BARLAT, DT, ADJOINT><<<nblk, SENS_BLOCK, 0, stream>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_reduce_partials(A.partials, nblk, 1 + A.n_active, A.h.result, stream);
}

}  // namespace

cudaError_t launch_mp_sens_rate_dt(const SensArgs& A, int def_type, bool adjoint, cudaStream_t stream) {
    if (A.h.n == 0) return cudaMemsetAsync(A.h.result, 0, sizeof(double) * (1 + A.n_active), stream);
    if (def_type == CMADX_DEF_PLANE_STRESS)
        return adjoint ? launch_t<CMADX_DEF_PLANE_STRESS, true>(A, stream) : launch_t<CMADX_DEF_PLANE_STRESS, false>(A, stream);
    return adjoint ? launch_t<CMADX_DEF_UNIAXIAL_STRESS, true>(A, stream) : launch_t<CMADX_DEF_UNIAXIAL_STRESS, false>(A, stream);
}

}  // namespace cmadx
