// K2 - calibration-objective sensitivities over whole load histories.
//
// One thread per material point walks that point's stored history in
// registers: the adjoint variant backwards (phi_t = A_t^-T(-dJ/dxi_t + h),
// h <- -B_t^T phi_t, g += phi_t^T dC/dp_t + dJ/dp_t), the direct variant forwards
// (X_t = A_t^-1(-dC/dp_t - B_t X_{t-1}), g += dJ/dxi_t X_t + dJ/dp_t) with
// A = dC/dxi, B = dC/dxi_prev evaluated at the stored (xi_t, xi_{t-1}).  Every
// slab access is coalesced across the warp.  (J, grad) are reduced block-wise
// (shuffle + shared memory) into per-block partials and then by a single block
// in fixed order, so the result is bit-reproducible; the cross-GPU sum is one
// NCCL allreduce of 1 + n_active doubles issued by the host layer.
//
// Replaces (reference file:line): cmad/objectives/mp_objective.py:92-147
// (MPAdjointObjective), :150-215 (MPDirectObjective), with the QoI of
// cmad/qois/calibration.py:56-66.
#include "j2_radial.cuh"
#include "mp_outputs.cuh"
#include "mp_sens.cuh"

namespace cmadx {


namespace {

constexpr int SENS_BLOCK = 128;

// Calibration QoI at one step: J, r_a = dJ/d sigma_a (both tensor entries of an
// off-diagonal component summed)
CMADX_DEV double qoi_terms(const double (&w)[9], const double (&sig)[6], const double (&d)[9],
                           double (&r)[6]) {
    // tensor entry (i,j) -> packed component
    const int comp[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    double J = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) r[a] = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double mis = w[k] * (sig[comp[k]] - d[k]);
        J = fma(0.5 * mis, mis, J);
        r[comp[k]] = fma(w[k], mis, r[comp[k]]);
    }
    return J;
}

// 6x6 maps between global and material symmetric-tensor components for a rotation Q
// (cmad/models/small_elastic_plastic.py:44-62, 318-319): T[c][b] = d(Q^T e Q)_c / d e_b,
// S[a][c] = d(Q s Q^T)_a / d s_c (packed components, both tensor entries moving)
CMADX_DEV void sens_rot_maps(const double* Q, double (&T)[6][6], double (&S)[6][6]) {
    const int ci[6] = {0, 0, 0, 1, 1, 2}, cj[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const int i = ci[c], j = cj[c], k = ci[b], l = cj[b];
            double t = Q[3 * k + i] * Q[3 * l + j];
            double s = Q[3 * i + k] * Q[3 * j + l];
            if (k != l) { t += Q[3 * l + i] * Q[3 * k + j]; s += Q[3 * i + l] * Q[3 * j + k]; }
            T[c][b] = t;
            S[c][b] = s;
        }
}

// NA_MAX: compile-time bound of the active-parameter loops (6 covers the usual calibration
// sets such as [E, nu, D, S, Y]; 16 = CMADX_MAX_ACTIVE), which sizes the register-resident
// gradient accumulators.
// ROT: rotated material axes ("rotation matrix" != I): the strain is taken to material axes, the
// state lives there, and the QoI compares the GLOBAL cauchy Q sigma_m Q^T with the data.
template <int YK, bool ADJOINT, int NA_MAX, bool ROT = false>
__global__ void __launch_bounds__(SENS_BLOCK, (YK == CMADX_YIELD_J2 && NA_MAX <= 6 && !ROT) ? 4 : 1)
mp_sens_kernel(const __grid_constant__ SensArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int N = A.h.nsteps;
    const int na = A.n_active;
    const int sc = A.h.strain_comps;

    double g[NA_MAX];
#pragma unroll
    for (int c = 0; c < NA_MAX; ++c) g[c] = 0.0;
    double Jacc = 0.0;
    double hist[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) hist[c] = 0.0;
    // direct: dxi/dp (7 x n_active per point) carried forward in SHARED memory, entry (c, r) of
    // thread t at (c * 7 + r) * SENS_BLOCK + t (conflict-free).  It used to live in local memory:
    // 336 B per thread that the streaming history loads kept evicting from L1 (direct 7.56 ms
    // against 4.0 ms for the adjoint on the same history, 0.30 of HBM).
    extern __shared__ double Xs[];
    if (!ADJOINT) {
        for (int c = 0; c < na * 7; ++c) Xs[c * SENS_BLOCK + threadIdx.x] = 0.0;
    }

    // the state pair (xi_t, xi_{t-1}) shares one member with the next step's pair: it is
    // carried in registers, so every stored state is read from HBM exactly once
    double x[7], xp[7];
    {
        const double* x0 = A.h.xi_hist + (int64_t)(ADJOINT ? N : 0) * 7 * ld + i;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const double v = live ? __ldg(x0 + c * ld) : 0.0;
            x[c] = v; xp[c] = v;
        }
    }
    for (int s = 0; s < N; ++s) {
        const int t = ADJOINT ? N - s : s + 1;
        double em[6], d[9];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)(ADJOINT ? t - 1 : t) * 7 * ld + i;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const double v = __ldg(xs + c * ld);
                if (ADJOINT) xp[c] = v; else x[c] = v;
            }
            const double* es = A.h.strain + (int64_t)t * sc * ld + i;
            if (sc == 6) {
#pragma unroll
                for (int c = 0; c < 6; ++c) em[c] = __ldg(es + c * ld);
            } else {
                double gq[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) gq[c] = __ldg(es + c * ld);
                em[0] = gq[0]; em[3] = gq[4]; em[5] = gq[8];
                em[1] = 0.5 * (gq[1] + gq[3]); em[2] = 0.5 * (gq[2] + gq[6]); em[4] = 0.5 * (gq[5] + gq[7]);
            }
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < 6; ++c) em[c] = 1e-3 * (c == 0);
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
        }
        if constexpr (ROT) {                       // global strain -> material axes
            double T[6][6], S[6][6], eg[6];
            sens_rot_maps(m.Q, T, S);
#pragma unroll
            for (int c = 0; c < 6; ++c) eg[c] = em[c];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < 6; ++b) sacc = fma(T[c][b], eg[b], sacc);
                em[c] = sacc;
            }
        }
        SepPoint<YK> pt;
        double C[7];
        pt.residual(m, x, xp, em, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        double ee[6], sig[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = em[a] - x[a];
        const double tree = ee[0] + ee[3] + ee[5];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], m.lam * tree) : m.two_mu * ee[a];
        double r[6];
        if constexpr (ROT) {
            // J on the global cauchy S sigma_m; cotangent back to material axes: r_m = S^T r_g
            double T[6][6], S[6][6], sg[6], rg[6];
            sens_rot_maps(m.Q, T, S);
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double sacc = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) sacc = fma(S[a][c], sig[c], sacc);
                sg[a] = sacc;
            }
            Jacc += qoi_terms(A.h.weight, sg, d, rg);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int a = 0; a < 6; ++a) sacc = fma(S[a][c], rg[a], sacc);
                r[c] = sacc;
            }
        } else {
            Jacc += qoi_terms(A.h.weight, sig, d, r);
        }
        // dJ/dxi (row) : d sigma_a/d ep_b = -(2mu delta_ab + lam [a diag][b diag])
        const double rtr = r[0] + r[3] + r[5];
        double dJdx[7];
#pragma unroll
        for (int b = 0; b < 6; ++b) dJdx[b] = is_diag(b) ? fma(-m.two_mu, r[b], -m.lam * rtr) : -m.two_mu * r[b];
        dJdx[6] = 0.0;
        // dJ/dp: only through lambda, mu (sigma = lam tr(ee) I + 2 mu ee)
        double ree = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) ree = fma(r[a], ee[a], ree);
        const double dJdlam = tree * rtr, dJdmu = 2.0 * ree;
        // (dn/dsigma : ee), n : ee for the dC/dp columns
        double Mee[6], nee = 0.0;
        if constexpr (YK == CMADX_YIELD_J2) {
            pt.yf.Mvec(ee, Mee);
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < 6; ++b) sacc = fma(pt.yf.M(a, b), ee[b], sacc);
                Mee[a] = sacc;
            }
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) nee = fma(mult(a) * pt.n[a], ee[a], nee);
        // J2: closed-form solve with the Jacobian (j2_radial.cuh) - no LU at all.
        // Other surfaces: threshold-pivoted register LU as in the Newton kernels.
        constexpr bool CLOSED = (YK == CMADX_YIELD_J2);
        RegLU<CLOSED ? 1 : 7> lu;
        bool trouble = false, slow = false;
        if constexpr (!CLOSED) {
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
            // natural order is stable for these (row-/column-scaled SPD + border) matrices
            trouble = lu.factor_natural();
            slow = __any_sync(__activemask(), trouble);
            if (slow && trouble) {
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
                lu.factor_pivot();
            }
        }
        auto solve7 = [&](double (&v)[7]) {
            if constexpr (CLOSED) {
                j2_jacobian_solve<ADJOINT>(m, pt, dg, v);
            } else {
                if (slow && trouble) lu.solve_pivot(v); else lu.solve_natural(v);
            }
        };
        if (ADJOINT) {
            double phi[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) phi[c] = hist[c] - dJdx[c];
            solve7(phi);
            if (A.phi_hist && live) {      // kept for the direct-adjoint Hessian pass (mp_hess.cu)
                double* ph = A.phi_hist + (int64_t)t * 7 * ld + i;
#pragma unroll
                for (int c = 0; c < 7; ++c) ph[c * ld] = phi[c];
            }
            // h <- -B^T phi
            double nphi = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) { hist[a] = phi[a]; nphi = fma(pt.n[a], phi[a], nphi); }
            hist[6] = pl ? -nphi : phi[6];
#pragma unroll
            for (int c = 0; c < NA_MAX; ++c) {
                if (c < na) {
                    const int pid = A.pid[c];
                    double col[7];
                    dC_dp_column(m, pid, pl, pt.yf, pt.n, pt.f, pt.eD, x[6], dg, Mee, nee, sig, col);
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < 7; ++q) acc = fma(phi[q], col[q], acc);
                    if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1)
                        acc += dJdlam * m.dlam[pid - CMADX_P_EL0] + dJdmu * m.dmu[pid - CMADX_P_EL0];
                    g[c] += acc;
                }
            }
        } else {
            // rhs = -dC/dp - B X_prev ;  B = [-I, n; 0, 0] (plastic) or -I (elastic)
            auto column = [&](const int pid, double* Xc, double& gc) {
                double col[7], rhs[7];
                dC_dp_column(m, pid, pl, pt.yf, pt.n, pt.f, pt.eD, x[6], dg, Mee, nee, sig, col);
                const double x6 = Xc[6 * SENS_BLOCK];
#pragma unroll
                for (int q = 0; q < 6; ++q) rhs[q] = -col[q] + Xc[q * SENS_BLOCK] - (pl ? pt.n[q] * x6 : 0.0);
                rhs[6] = -col[6] + (pl ? 0.0 : x6);
                solve7(rhs);
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 7; ++q) { Xc[q * SENS_BLOCK] = rhs[q]; acc = fma(dJdx[q], rhs[q], acc); }
                if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1)
                    acc += dJdlam * m.dlam[pid - CMADX_P_EL0] + dJdmu * m.dmu[pid - CMADX_P_EL0];
                gc += acc;
            };
#pragma unroll
            for (int c = 0; c < NA_MAX; ++c)
                if (c < na) column(A.pid[c], Xs + (c * 7) * SENS_BLOCK + threadIdx.x, g[c]);
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (ADJOINT) x[c] = xp[c]; else xp[c] = x[c];
        }
    }
    if (!live) {
        Jacc = 0.0;
#pragma unroll
        for (int c = 0; c < NA_MAX; ++c) g[c] = 0.0;
    } else if (A.h.J_point) {
        A.h.J_point[i] = Jacc;
    }
    // ---- block reduction (fixed order): warp shuffles, then 4 warps via smem
    __shared__ double sm[SENS_BLOCK / 32][1 + CMADX_MAX_ACTIVE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c <= NA_MAX; ++c) {
        if (c <= na) {
            double v = (c == 0) ? Jacc : g[c - 1];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) sm[warp][c] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x <= na) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < SENS_BLOCK / 32; ++w) v += sm[w][threadIdx.x];
        A.partials[(int64_t)blockIdx.x * (1 + na) + threadIdx.x] = v;
    }
}

// final reduction of the per-block partials in fixed order: one block per column (the columns
// are independent), 1024 threads striding over the partials, then a shared-memory tree - the
// summation order depends only on (nblk, thread count): bit-reproducible.  (The first version
// used ONE block of 256 threads for all columns: 116 us for the 32 000 partial rows of a 4 M-point
// FE adjoint step - 15 % of the step.)
constexpr int REDUCE_THREADS = 1024;
__global__ void __launch_bounds__(REDUCE_THREADS)
reduce_partials_kernel(const double* partials, int64_t nblk, int ncols, double* result) {
    __shared__ double sm[REDUCE_THREADS];
    const int c = blockIdx.x;
    double v = 0.0;
    for (int64_t b = threadIdx.x; b < nblk; b += REDUCE_THREADS) v += partials[b * ncols + c];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = REDUCE_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) result[c] = sm[0];
}

// dynamic shared memory of the direct variant: dxi/dp, 7 x n_active doubles per thread
template <class K>
cudaError_t sens_launch(K kern, const SensArgs& A, bool adjoint, int64_t nblk, cudaStream_t stream) {
    const size_t smem = adjoint ? 0 : sizeof(double) * 7 * (size_t)A.n_active * SENS_BLOCK;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)nblk, SENS_BLOCK, smem, stream>>>(A);
    return cudaGetLastError();
}

template <bool ADJOINT>
cudaError_t launch_sens_rot(const SensArgs& A, cudaStream_t stream) {
    const int64_t nblk = (A.h.n + SENS_BLOCK - 1) / SENS_BLOCK;
    constexpr int NA = CMADX_MAX_ACTIVE;
    cudaError_t e0;
    switch (A.m.yield) {
    case CMADX_YIELD_J2:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_J2, ADJOINT, NA, true>, A, ADJOINT, nblk, stream); break;
    case CMADX_YIELD_HILL:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_HILL, ADJOINT, NA, true>, A, ADJOINT, nblk, stream); break;
    case CMADX_YIELD_HOSFORD:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_HOSFORD, ADJOINT, NA, true>, A, ADJOINT, nblk, stream); break;
    case CMADX_YIELD_BARLAT:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_BARLAT, ADJOINT, NA, true>, A, ADJOINT, nblk, stream); break;
    default: return cudaErrorInvalidValue;
    }
    if (e0 != cudaSuccess) return e0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    reduce_partials_kernel<<<1 + A.n_active, REDUCE_THREADS, 0, stream>>>(A.partials, nblk, 1 + A.n_active, A.h.result);
    return cudaGetLastError();
}

template <bool ADJOINT, int NA_MAX>
cudaError_t launch_sens_t(const SensArgs& A, cudaStream_t stream) {
    const int64_t nblk = (A.h.n + SENS_BLOCK - 1) / SENS_BLOCK;
    cudaError_t e0;
    switch (A.m.yield) {
    case CMADX_YIELD_J2:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_J2, ADJOINT, NA_MAX>, A, ADJOINT, nblk, stream); break;
    case CMADX_YIELD_HILL:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_HILL, ADJOINT, NA_MAX>, A, ADJOINT, nblk, stream); break;
    case CMADX_YIELD_HOSFORD:
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_HOSFORD, ADJOINT, NA_MAX>, A, ADJOINT, nblk, stream); break;
    case CMADX_YIELD_BARLAT:
        if (NA_MAX < CMADX_MAX_ACTIVE) return cudaErrorInvalidValue;       // one instantiation: see launch_mp_sens
        e0 = sens_launch(mp_sens_kernel<CMADX_YIELD_BARLAT, ADJOINT, CMADX_MAX_ACTIVE>, A, ADJOINT, nblk, stream); break;
    default: return cudaErrorInvalidValue;
    }
    if (e0 != cudaSuccess) return e0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    reduce_partials_kernel<<<1 + A.n_active, REDUCE_THREADS, 0, stream>>>(A.partials, nblk, 1 + A.n_active, A.h.result);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_reduce_partials(const double* partials, int64_t nblk, int ncols, double* result,
                                   cudaStream_t stream) {
    if (ncols <= 0) return cudaSuccess;
    reduce_partials_kernel<<<ncols, REDUCE_THREADS, 0, stream>>>(partials, nblk, ncols, result);
    return cudaGetLastError();
}

cudaError_t launch_mp_sens(const SensArgs& A, bool adjoint, cudaStream_t stream) {
    if (A.h.n == 0) return cudaMemsetAsync(A.h.result, 0, sizeof(double) * (1 + A.n_active), stream);
    if (A.m.rot) return adjoint ? launch_sens_rot<true>(A, stream) : launch_sens_rot<false>(A, stream);
    if (A.n_active <= 6 && A.m.yield != CMADX_YIELD_BARLAT) return adjoint ? launch_sens_t<true, 6>(A, stream) : launch_sens_t<false, 6>(A, stream);
    return adjoint ? launch_sens_t<true, CMADX_MAX_ACTIVE>(A, stream) : launch_sens_t<false, CMADX_MAX_ACTIVE>(A, stream);
}

int64_t sens_blocks(int64_t n) { return (n + SENS_BLOCK - 1) / SENS_BLOCK; }

}  // namespace cmadx
