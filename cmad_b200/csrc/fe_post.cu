// Post-processing companions of the FE path that re-enter the constitutive model at a
// STORED state (no Newton):
//   fe_cauchy_kernel   - evaluate_cauchy_at_ips for COUPLED blocks
//                        (cmad/fem/postprocess.py:35-185): model.cauchy(xi, .., U_ip) at every
//                        (element, IP) from the converged xi history, global axes, packed
//                        xx,xy,xz,yy,yz,zz.  One thread per integration point; HBM-bound
//                        (grad_N 24 n_b + xi 56 read, 48 written per point).
//   embedded-BC kernels - _embedded_bc_enforce / _embedded_residual
//                        (cmad/fem/sparse_solve.py:1058-1174) on the deduplicated COO data:
//                        see cmadx_embedded_bc_* in the header.
#include <vector>

#include "fe_common.cuh"

namespace cmadx {
namespace {

template <int NB, int NIP, bool ROT>
__global__ void __launch_bounds__(FE_BLOCK) fe_cauchy_kernel(const DevMat m, const cmadx_fe_block_t b,
                                                             const double* __restrict__ xi_state,
                                                             double* __restrict__ sigma) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // point = e * n_ip + ip
    const int nip = NIP ? NIP : b.n_ip;                                    // NIP = 0: any rule
    if (p >= b.n_elems * nip) return;
    const int64_t e = p / nip;
    const double* g = b.grad_N + p * (NB * 3);
    double gN[NB][3], U[NB][3];
#pragma unroll
    for (int c = 0; c < NB * 3 / 4; ++c) {
        double v0, v1, v2, v3;
        ld256(g + 4 * c, v0, v1, v2, v3);
        (&gN[0][0])[4 * c] = v0; (&gN[0][0])[4 * c + 1] = v1; (&gN[0][0])[4 * c + 2] = v2; (&gN[0][0])[4 * c + 3] = v3;
    }
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) U[a][k] = __ldg(b.U + __ldg(b.elem_eq + e * (NB * 3) + 3 * a + k));
    double eps[6], em[6], sig[6];
    strain_from_U<NB>(U, gN, eps);
    double T[6][6], S[6][6];
    if (ROT) {
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) s = fma(T[c][q], eps[q], s);
            em[c] = s;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) em[c] = eps[c];
    }
    double ee[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = em[a] - __ldg(xi_state + p * 7 + a);
    const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        double s = sig[a];
        if (ROT) {
            s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(S[a][c], sig[c], s);
        }
        sigma[p * 6 + a] = s;
    }
}

// ---- embedded Dirichlet BCs on the deduplicated COO tangent ---------------------------------
// K_emb[e] = K[e] if both indices are free, or the entry is a prescribed diagonal; else 0
__global__ void __launch_bounds__(256) embedded_mask_kernel(const double* __restrict__ K,
                                                            const unsigned char* __restrict__ keep,
                                                            int64_t nnz, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) out[i] = keep[i] ? K[i] : 0.0;
}

// r[i] = R[i] + sum_{entries (i, j), j prescribed} K[e] (val_j - U_j)      (free rows, fixed order)
// r[i] = K_ii (U_i - val_i)                                                (prescribed rows)
__global__ void __launch_bounds__(256) embedded_residual_kernel(
        const double* __restrict__ K, const double* __restrict__ R, const double* __restrict__ U,
        const double* __restrict__ presc_vals, const int* __restrict__ slot_of_dof,
        const int64_t* __restrict__ diag_pos, const int64_t* __restrict__ cpl_ptr,
        const int64_t* __restrict__ cpl_entry, const int* __restrict__ cpl_slot,
        const int* __restrict__ presc_idx, int64_t n, double* __restrict__ r) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slot_of_dof[i];
    if (s >= 0) {
        r[i] = K[diag_pos[s]] * (U[i] - presc_vals[s]);
        return;
    }
    double acc = R[i];
    for (int64_t k = cpl_ptr[i]; k < cpl_ptr[i + 1]; ++k) {
        const int sj = cpl_slot[k];
        acc = fma(K[cpl_entry[k]], presc_vals[sj] - U[presc_idx[sj]], acc);
    }
    r[i] = acc;
}

}  // namespace

cudaError_t launch_fe_cauchy(const DevMat& m, const cmadx_fe_block_t& b, const double* xi_state, double* sigma,
                             cudaStream_t stream) {
    if (b.n_elems == 0) return cudaSuccess;
    const int64_t npts = b.n_elems * b.n_ip;
    const unsigned nblk = (unsigned)((npts + FE_BLOCK - 1) / FE_BLOCK);
    const bool rot = m.rot != 0;
    if (b.n_basis == 4) {
        if (rot) fe_cauchy_kernel<4, 0, true><<<nblk, FE_BLOCK, 0, stream>>>(m, b, xi_state, sigma);
        else fe_cauchy_kernel<4, 0, false><<<nblk, FE_BLOCK, 0, stream>>>(m, b, xi_state, sigma);
    } else {
        if (rot) fe_cauchy_kernel<8, 0, true><<<nblk, FE_BLOCK, 0, stream>>>(m, b, xi_state, sigma);
        else fe_cauchy_kernel<8, 0, false><<<nblk, FE_BLOCK, 0, stream>>>(m, b, xi_state, sigma);
    }
    return cudaGetLastError();
}

struct EmbeddedPlan {
    int64_t n = 0, nnz = 0, n_presc = 0;
    unsigned char* keep = nullptr;
    int* slot_of_dof = nullptr;
    int* presc_idx = nullptr;
    int64_t* diag_pos = nullptr;
    int64_t* cpl_ptr = nullptr;
    int64_t* cpl_entry = nullptr;
    int* cpl_slot = nullptr;
    void* all = nullptr;
};

cudaError_t embedded_plan_build(const int64_t* rows, const int64_t* cols, int64_t nnz, int64_t n,
                                const int64_t* presc, int64_t n_presc, EmbeddedPlan** out) {
    std::vector<int> slot(n, -1);
    for (int64_t s = 0; s < n_presc; ++s) {
        if (presc[s] < 0 || presc[s] >= n) return cudaErrorInvalidValue;
        slot[presc[s]] = (int)s;
    }
    std::vector<unsigned char> keep(nnz);
    std::vector<int64_t> diag(n_presc, -1), ptr(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t i = rows[e], j = cols[e];
        if (i < 0 || i >= n || j < 0 || j >= n) return cudaErrorInvalidValue;
        const bool pi = slot[i] >= 0, pj = slot[j] >= 0;
        keep[e] = (!pi && !pj) || (pi && i == j);
        if (pi && i == j) diag[slot[i]] = e;
        if (!pi && pj) ptr[i + 1]++;
    }
    for (int64_t s = 0; s < n_presc; ++s)
        if (diag[s] < 0) return cudaErrorInvalidValue;          // a prescribed dof without a diagonal entry
    for (int64_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    const int64_t ncpl = ptr[n];
    std::vector<int64_t> centry(ncpl), fill(ptr.begin(), ptr.end() - 1);
    std::vector<int> cslot(ncpl), pidx(n_presc);
    for (int64_t e = 0; e < nnz; ++e) {                          // increasing entry order within a row
        const int64_t i = rows[e], j = cols[e];
        if (slot[i] < 0 && slot[j] >= 0) { const int64_t k = fill[i]++; centry[k] = e; cslot[k] = slot[j]; }
    }
    for (int64_t s = 0; s < n_presc; ++s) pidx[s] = (int)presc[s];
    auto* P = new EmbeddedPlan;
    P->n = n; P->nnz = nnz; P->n_presc = n_presc;
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t sz[7] = {al(nnz), al(sizeof(int) * n), al(sizeof(int) * n_presc), al(8 * n_presc),
                          al(8 * (n + 1)), al(8 * ncpl), al(sizeof(int) * ncpl)};
    size_t tot = 0;
    for (size_t s : sz) tot += s;
    cudaError_t err = cudaMalloc(&P->all, tot ? tot : 256);
    if (err != cudaSuccess) { delete P; return err; }
    char* base = static_cast<char*>(P->all);
    const void* src[7] = {keep.data(), slot.data(), pidx.data(), diag.data(), ptr.data(), centry.data(), cslot.data()};
    const size_t bytes[7] = {(size_t)nnz, sizeof(int) * n, sizeof(int) * n_presc, (size_t)8 * n_presc,
                             (size_t)8 * (n + 1), (size_t)8 * ncpl, sizeof(int) * ncpl};
    void* dst[7];
    for (int k = 0; k < 7; ++k) {
        dst[k] = base;
        if (bytes[k]) {
            err = cudaMemcpy(base, src[k], bytes[k], cudaMemcpyHostToDevice);
            if (err != cudaSuccess) { cudaFree(P->all); delete P; return err; }
        }
        base += sz[k];
    }
    P->keep = (unsigned char*)dst[0]; P->slot_of_dof = (int*)dst[1]; P->presc_idx = (int*)dst[2];
    P->diag_pos = (int64_t*)dst[3]; P->cpl_ptr = (int64_t*)dst[4]; P->cpl_entry = (int64_t*)dst[5];
    P->cpl_slot = (int*)dst[6];
    *out = P;
    return cudaSuccess;
}

void embedded_plan_free(EmbeddedPlan* P) {
    if (!P) return;
    cudaFree(P->all);
    delete P;
}

cudaError_t embedded_apply(const EmbeddedPlan* P, const double* K, const double* R, const double* U,
                           const double* presc_vals, double* r_out, double* K_emb_out, cudaStream_t stream) {
    if (P->n > 0 && r_out) {
        embedded_residual_kernel<<<(unsigned)((P->n + 255) / 256), 256, 0, stream>>>(
            K, R, U, presc_vals, P->slot_of_dof, P->diag_pos, P->cpl_ptr, P->cpl_entry, P->cpl_slot,
            P->presc_idx, P->n, r_out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (P->nnz > 0 && K_emb_out) {
        embedded_mask_kernel<<<(unsigned)((P->nnz + 255) / 256), 256, 0, stream>>>(K, P->keep, P->nnz, K_emb_out);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

int64_t embedded_plan_dims(const EmbeddedPlan* P, int which) {
    return which == 0 ? P->n : (which == 1 ? P->nnz : P->n_presc);
}

}  // namespace cmadx
