// K2 for the PLANE_STRESS / UNIAXIAL_STRESS deformation types: adjoint / direct calibration
// gradients over stored load histories (cmad/objectives/mp_objective.py:92-215 with the
// Calibration QoI, cmad/qois/calibration.py:56-66) on the bordered (n_xi = 8 / 9) local
// systems of sep_point_dt.cuh - KA5 of the reference (tests/objectives/
// test_J2_fd_checks.py:303-349 runs its gradient checks in plane stress).
// Same structure as mp_sens.cu: one thread walks its point's history, the state pair is
// carried in registers, A = dC/dxi factored by the threshold-pivoted register LU (transposed
// for the adjoint), fixed-order block reduction -> partials -> reduce_partials_kernel.
#include "mp_outputs.cuh"
#include "sep_point_dt.cuh"
#include "mp_sens.cuh"

namespace cmadx {

cudaError_t launch_reduce_partials(const double* partials, int64_t nblk, int ncols, double* result,
                                   cudaStream_t stream);

namespace {

constexpr int SENS_DT_BLOCK = 128;

template <int YK, int DT, bool ADJOINT, bool ROT = false>
__global__ void __launch_bounds__(SENS_DT_BLOCK) mp_sens_dt_kernel(const __grid_constant__ SensArgs A) {
    using Pt = typename std::conditional<ROT, SepPointDTRot<YK, DT>, SepPointDT<YK, DT>>::type;
    constexpr int N = Pt::N, NZ = Pt::NZ;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int NT = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;

    double g[CMADX_MAX_ACTIVE];
#pragma unroll
    for (int c = 0; c < CMADX_MAX_ACTIVE; ++c) g[c] = 0.0;
    double Jacc = 0.0, hist[N];
#pragma unroll
    for (int c = 0; c < N; ++c) hist[c] = 0.0;
    double X[CMADX_MAX_ACTIVE][N];          // direct: dxi/dp carried forward (local memory)
    if (!ADJOINT) {
        for (int c = 0; c < na; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) X[c][r] = 0.0;
    }
    double x[N], xp[N];
    {
        const double* x0 = A.h.xi_hist + (int64_t)(ADJOINT ? NT : 0) * N * ld + i;
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const double v = live ? __ldg(x0 + c * ld) : ((c < 7) ? 0.0 : 1.0);
            x[c] = v; xp[c] = v;
        }
    }
    const double lr = m.lam * m.inv_two_mu;
    for (int s = 0; s < NT; ++s) {
        const int t = ADJOINT ? NT - s : s + 1;
        double em[6] = {1e-3, 0.0, 0.0, 0.0, 0.0, 0.0}, d[9];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)(ADJOINT ? t - 1 : t) * N * ld + i;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                const double v = __ldg(xs + c * ld);
                if (ADJOINT) xp[c] = v; else x[c] = v;
            }
            load_dt_strain<DT>(A.h.strain + (int64_t)t * sc * ld, sc, ld, i, em);
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
        }
        Pt pt;
        double C[N];
        pt.residual(m, x, xp, em, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        double et[6], ee[6], sig[6], sgl[6];      // sig: material-frame stress, sgl: global (what the QoI reads)
        pt.material_strain(x, em, et);
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = et[a] - x[a];
        const double tree = ee[0] + ee[3] + ee[5];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], m.lam * tree) : m.two_mu * ee[a];
        pt.to_global(sig, sgl);
        // Calibration QoI: J, r_a = dJ/d sigma_a (both entries of an off-diagonal component summed)
        double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        double dJdz[NZ];                   // direct dependence of the QoI on the stretch dofs
#pragma unroll
        for (int k = 0; k < NZ; ++k) dJdz[k] = 0.0;
        if (DT == CMADX_DEF_UNIAXIAL_STRESS && A.h.qoi_kind == CMADX_QOI_UNIAXIAL_CALIBRATION) {
            // UniaxialCalibration (cmad/qois/uniaxial_calibration.py:69-85): pred = [sigma_axial,
            // lambda_2 - 1, lambda_3 - 1], weights of this step, data rows 0..2
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
            const int comp[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double mis = A.h.weight[k] * (sgl[comp[k]] - d[k]);
                Jacc = fma(0.5 * mis, mis, Jacc);
                r[comp[k]] = fma(A.h.weight[k], mis, r[comp[k]]);
            }
        }
        // cotangent of the stress in material axes, and dJ/dx through the stress (sep_point_dt.cuh)
        double rm[6], dJdx[N];
        pt.stress_cotangent(m, r, rm, dJdx);
#pragma unroll
        for (int k = 0; k < NZ; ++k) dJdx[7 + k] += dJdz[k];
        const double rtr = rm[0] + rm[3] + rm[5];
        double ree = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) ree = fma(rm[a], ee[a], ree);
        const double dJdlam = tree * rtr, dJdmu = 2.0 * ree;
        double Mee[6], nee = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double sacc = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) sacc = fma(pt.b.yf.M(a, b), ee[b], sacc);
            Mee[a] = sacc;
            nee = fma(mult(a) * pt.b.n[a], ee[a], nee);
        }
        RegLU<N> lu;
        auto load_A = [&]() {
            if (ADJOINT) {
                double Jm[N][N];
                pt.jacobian(m, dg, Jm);
#pragma unroll
                for (int a = 0; a < N; ++a)
#pragma unroll
                    for (int b = 0; b < N; ++b) lu.a[a][b] = Jm[b][a];
            } else {
                pt.jacobian(m, dg, lu.a);
            }
        };
        load_A();
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) { load_A(); lu.factor_pivot(); }
        auto solveN = [&](double (&v)[N]) {
            if (slow && trouble) lu.solve_pivot(v); else lu.solve_natural(v);
        };
        // one column of dC/dp (N rows): the FULL_3D rows + the stress rows (elastic parameters only)
        auto dCdp = [&](int pid, double (&col)[N]) {
            double c7[7];
            dC_dp_column(m, pid, pl, pt.b.yf, pt.b.n, pt.b.f, pt.b.eD, x[6], dg, Mee, nee, sig, c7);
#pragma unroll
            for (int q = 0; q < 7; ++q) col[q] = c7[q];
            double dr = 0.0;
            if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
                const int k = pid - CMADX_P_EL0;
                dr = (m.dlam[k] * m.two_mu - m.lam * 2.0 * m.dmu[k]) * m.inv_two_mu * m.inv_two_mu * tree;
            }
#pragma unroll
            for (int k = 0; k < NZ; ++k) col[7 + k] = dr;
        };
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
            // h <- -B^T phi, B = dC/dxi_prev = [[-I, n, 0], [0, 0, 0], [0, 0, 0]] (plastic) or diag(-I7, 0)
            double nphi = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) { hist[a] = phi[a]; nphi = fma(pt.b.n[a], phi[a], nphi); }
            hist[6] = pl ? -nphi : phi[6];
#pragma unroll
            for (int k = 0; k < NZ; ++k) hist[7 + k] = 0.0;
            for (int c = 0; c < na; ++c) {
                const int pid = A.pid[c];
                double col[N];
                dCdp(pid, col);
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; ++q) acc = fma(phi[q], col[q], acc);
                if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1)
                    acc += dJdlam * m.dlam[pid - CMADX_P_EL0] + dJdmu * m.dmu[pid - CMADX_P_EL0];
                g[c] += acc;
            }
        } else {
            for (int c = 0; c < na; ++c) {
                const int pid = A.pid[c];
                double col[N], rhs[N];
                dCdp(pid, col);
                const double x6 = X[c][6];
#pragma unroll
                for (int q = 0; q < 6; ++q) rhs[q] = -col[q] + X[c][q] - (pl ? pt.b.n[q] * x6 : 0.0);
                rhs[6] = -col[6] + (pl ? 0.0 : x6);
#pragma unroll
                for (int k = 0; k < NZ; ++k) rhs[7 + k] = -col[7 + k];
                solveN(rhs);
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; ++q) { X[c][q] = rhs[q]; acc = fma(dJdx[q], rhs[q], acc); }
                if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1)
                    acc += dJdlam * m.dlam[pid - CMADX_P_EL0] + dJdmu * m.dmu[pid - CMADX_P_EL0];
                g[c] += acc;
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) {
            if (ADJOINT) x[c] = xp[c]; else xp[c] = x[c];
        }
        (void)lr;
    }
    if (!live) {
        Jacc = 0.0;
#pragma unroll
        for (int c = 0; c < CMADX_MAX_ACTIVE; ++c) g[c] = 0.0;
    } else if (A.h.J_point) {
        A.h.J_point[i] = Jacc;
    }
    __shared__ double sm[SENS_DT_BLOCK / 32][1 + CMADX_MAX_ACTIVE];
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
        for (int w = 0; w < SENS_DT_BLOCK / 32; ++w) v += sm[w][threadIdx.x];
        A.partials[(int64_t)blockIdx.x * (1 + na) + threadIdx.x] = v;
    }
}

template <int DT, bool ADJOINT>
cudaError_t launch_t(const SensArgs& A, cudaStream_t stream) {
    const int64_t nblk = (A.h.n + SENS_DT_BLOCK - 1) / SENS_DT_BLOCK;
    if (A.m.rot) {
        switch (A.m.yield) {
        case CMADX_YIELD_J2:
            mp_sens_dt_kernel<CMADX_YIELD_J2, DT, ADJOINT, true><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
        case CMADX_YIELD_HILL:
            mp_sens_dt_kernel<CMADX_YIELD_HILL, DT, ADJOINT, true><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
        case CMADX_YIELD_HOSFORD:
            mp_sens_dt_kernel<CMADX_YIELD_HOSFORD, DT, ADJOINT, true><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
        case CMADX_YIELD_BARLAT:
            mp_sens_dt_kernel<CMADX_YIELD_BARLAT, DT, ADJOINT, true><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
        default: return cudaErrorInvalidValue;
        }
        cudaError_t er = cudaGetLastError();
        if (er != cudaSuccess) return er;
        return launch_reduce_partials(A.partials, nblk, 1 + A.n_active, A.h.result, stream);
    }
    switch (A.m.yield) {
    case CMADX_YIELD_J2:
        mp_sens_dt_kernel<CMADX_YIELD_J2, DT, ADJOINT><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HILL:
        mp_sens_dt_kernel<CMADX_YIELD_HILL, DT, ADJOINT><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_HOSFORD:
        mp_sens_dt_kernel<CMADX_YIELD_HOSFORD, DT, ADJOINT><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
    case CMADX_YIELD_BARLAT:
        mp_sens_dt_kernel<CMADX_YIELD_BARLAT, DT, ADJOINT><<<(unsigned)nblk, SENS_DT_BLOCK, 0, stream>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_reduce_partials(A.partials, nblk, 1 + A.n_active, A.h.result, stream);
}

}  // namespace

cudaError_t launch_mp_sens_dt(const SensArgs& A, int def_type, bool adjoint, cudaStream_t stream) {
    if (A.h.n == 0) return cudaMemsetAsync(A.h.result, 0, sizeof(double) * (1 + A.n_active), stream);
    if (def_type == CMADX_DEF_PLANE_STRESS)
        return adjoint ? launch_t<CMADX_DEF_PLANE_STRESS, true>(A, stream) : launch_t<CMADX_DEF_PLANE_STRESS, false>(A, stream);
    return adjoint ? launch_t<CMADX_DEF_UNIAXIAL_STRESS, true>(A, stream) : launch_t<CMADX_DEF_UNIAXIAL_STRESS, false>(A, stream);
}

}  // namespace cmadx
