// The raw AD products of the reference's Model object at a GIVEN state, for a batch of points:
//   dC/dU             cmad/models/model.py:126-133 (jacfwd over the GlobalFieldsAtPoint argument)
//   dcauchy/dxi       cmad/models/model.py:150-152 (jacfwd of cauchy_fun over the state blocks)
//   dcauchy/dU        the partial the IFT tangent is built from
//   dcauchy/dparams   cmad/models/model.py:152      (jacrev over the parameter leaves)
// in closed form for SmallElasticPlastic / FULL_3D (small_elastic_plastic.py:238-331):
//   C sees U only through the material-frame strain e_m = Q^T sym(grad u) Q,
//     plastic branch: dC_a/de_m,b = -dgamma 2 mu M(a, b),  dC_alpha/de_m,b = mult(b) n_b;  elastic branch: 0
//   cauchy = Q [lambda tr(e_m - ep) I + 2 mu (e_m - ep)] Q^T: linear in ep and e_m, independent of alpha;
//   of the parameters only the elastic constants enter it (and the rotation-matrix entries, which this
//   entry point does not differentiate).
// dC/dU_prev and dcauchy/dxi_prev are identically zero for this model (the reference's AD returns
// zeros as well, tests/golden/ref_model_partials.npz) and have no output.
// Derivatives with respect to U are with respect to the SYMMETRIC strain components (both tensor
// entries moving); the reference's single-entry columns are half of them off the diagonal.
// One thread per point, component-major arrays; not a hot path (Model.evaluate-style inspection).
#include <atomic>

#include "mp_update.cuh"
#include "mp_outputs.cuh"

namespace cmadx {
int cuda_fail(cudaError_t e);
extern std::atomic<int64_t> g_launches;

struct PartialArgs {
    DevMat m;
    int n_active;
    int pid[CMADX_MAX_ACTIVE];
    cmadx_mp_partials_t p;
};

namespace {

template <int YK, bool ROT>
__global__ void __launch_bounds__(128)
mp_partials_kernel(const __grid_constant__ PartialArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.p.n) return;
    const int64_t ld = A.p.ld;
    const DevMat& m = A.m;
    double x[7], xp[7], e[6];
#pragma unroll
    for (int c = 0; c < 7; ++c) { x[c] = __ldg(A.p.xi + c * ld + i); xp[c] = __ldg(A.p.xi_prev + c * ld + i); }
    if (A.p.strain_comps == 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = __ldg(A.p.strain + c * ld + i);
    } else {
        double g[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) g[c] = __ldg(A.p.strain + c * ld + i);
        e[0] = g[0]; e[3] = g[4]; e[5] = g[8];
        e[1] = 0.5 * (g[1] + g[3]); e[2] = 0.5 * (g[2] + g[6]); e[4] = 0.5 * (g[5] + g[7]);
    }
    double T[6][6], S[6][6], em[6];
    if (ROT) {
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s = fma(T[c][b], e[b], s);
            em[c] = s;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) em[c] = e[c];
    }
    SepPoint<YK> pt;
    double C[7];
    pt.residual(m, x, xp, em, C);
    const double dg = x[6] - xp[6];
    const bool pl = pt.plastic;

    if (A.p.dC_deps) {
        double Dm[7][6];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
#pragma unroll
            for (int a = 0; a < 6; ++a) Dm[a][b] = pl ? -dg * m.two_mu * pt.yf.M(a, b) : 0.0;
            Dm[6][b] = pl ? mult(b) * pt.n[b] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double v = Dm[r][b];
                if (ROT) {
                    v = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) v = fma(Dm[r][c], T[c][b], v);
                }
                st(A.p.dC_deps, r * 6 + b, ld, i, v);
            }
    }
    // elastic stiffness in material axes, by symmetric component: Cel[a][c]
    auto cel = [&](int a, int c) { return ((a == c) ? m.two_mu : 0.0) + ((is_diag(a) && is_diag(c)) ? m.lam : 0.0); };
    if (A.p.dsig_dxi) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double v = -cel(a, b);
                if (ROT) {
                    v = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) v = fma(S[a][c], -cel(c, b), v);
                }
                st(A.p.dsig_dxi, a * 7 + b, ld, i, v);
            }
            st(A.p.dsig_dxi, a * 7 + 6, ld, i, 0.0);
        }
    }
    if (A.p.dsig_deps) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double v = cel(a, b);
                if (ROT) {
                    v = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c)
#pragma unroll
                        for (int d = 0; d < 6; ++d) v = fma(S[a][c] * cel(c, d), T[d][b], v);
                }
                st(A.p.dsig_deps, a * 6 + b, ld, i, v);
            }
    }
    if (A.p.dsig_dp && A.n_active > 0) {
        double ee[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = em[a] - x[a];
        const double tr = ee[0] + ee[3] + ee[5];
        for (int c = 0; c < A.n_active; ++c) {
            const int pid = A.pid[c];
            double dsm[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
                const double dl = m.dlam[pid - CMADX_P_EL0], dm = m.dmu[pid - CMADX_P_EL0];
#pragma unroll
                for (int a = 0; a < 6; ++a) dsm[a] = 2.0 * dm * ee[a] + (is_diag(a) ? dl * tr : 0.0);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double v = dsm[a];
                if (ROT) {
                    v = 0.0;
#pragma unroll
                    for (int d = 0; d < 6; ++d) v = fma(S[a][d], dsm[d], v);
                }
                st(A.p.dsig_dp, (int64_t)a * A.n_active + c, ld, i, v);
            }
        }
    }
}

template <int YK>
cudaError_t launch_yk(const PartialArgs& A, cudaStream_t s) {
    const unsigned nblk = (unsigned)((A.p.n + 127) / 128);
    if (A.m.rot) mp_partials_kernel<YK, true><<<nblk, 128, 0, s>>>(A);
    else mp_partials_kernel<YK, false><<<nblk, 128, 0, s>>>(A);
    return cudaGetLastError();
}

}  // namespace
}  // namespace cmadx

extern "C" int cmadx_mp_model_partials(const cmadx_material_t* mat, const int32_t* active_pid, int32_t n_active,
                                       const cmadx_mp_partials_t* p, void* stream) {
    using namespace cmadx;
    if (!p) return CMADX_EINVAL;
    PartialArgs A;
    if (int rc = make_dev_mat(mat, &A.m)) return rc;
    if (A.m.model != CMADX_MODEL_SMALL_ELASTIC_PLASTIC) return CMADX_EUNSUPPORTED;
    if (p->n < 0 || p->ld < p->n || (p->strain_comps != 6 && p->strain_comps != 9)) return CMADX_EINVAL;
    if (n_active < 0 || n_active > CMADX_MAX_ACTIVE || (n_active > 0 && !active_pid)) return CMADX_EINVAL;
    if (p->n > 0 && (!p->xi || !p->xi_prev || !p->strain)) return CMADX_EINVAL;
    for (int c = 0; c < n_active; ++c) {
        const int pid = active_pid[c];
        if (pid < 0 || pid >= CMADX_NUM_PARAM_IDS) return CMADX_EINVAL;
        // d cauchy / d(rotation-matrix entry) is not carried
        if (p->dsig_dp && pid >= CMADX_P_Q00 && pid < CMADX_P_BARLAT_C0) return CMADX_EUNSUPPORTED;
        A.pid[c] = pid;
    }
    A.n_active = n_active;
    A.p = *p;
    if (p->n == 0) return CMADX_OK;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    switch (A.m.yield) {
    case CMADX_YIELD_J2: e = launch_yk<CMADX_YIELD_J2>(A, s); break;
    case CMADX_YIELD_HILL: e = launch_yk<CMADX_YIELD_HILL>(A, s); break;
    case CMADX_YIELD_HOSFORD: e = launch_yk<CMADX_YIELD_HOSFORD>(A, s); break;
    case CMADX_YIELD_BARLAT: e = launch_yk<CMADX_YIELD_BARLAT>(A, s); break;
    default: return CMADX_EINVAL;
    }
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CMADX_OK;
}
