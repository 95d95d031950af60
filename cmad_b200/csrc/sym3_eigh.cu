// Batched eigen-decomposition of symmetric 3x3 tensors (principal stresses / stretches).
//
// Replaces (reference file:line): cmad/util/jax_eigen_decomposition.py:86-171
// (`compute_eigen_decomposition` / `sorted_eigen_decomposition`: the trigonometric closed form of
// Harari-Albocher / Scherzinger-Dohrmann with deflation for the vectors) - same results
// (ascending eigenvalues, orthonormal eigenvectors as columns, defined up to sign), obtained with
// the cyclic Jacobi iteration the Yld2004-18p kernels use (barlat.cuh): backward stable also for
// close eigenvalues, no trigonometric calls, a few sweeps of 3 rotations in registers.
//
// One thread per tensor, component-major arrays (row c of A6 / w / V at c * ld): every access is
// coalesced.  HBM-bound: 48 B in, 96 B out per tensor.
#include <atomic>

#include "point_solver.cuh"

namespace cmadx {
int cuda_fail(cudaError_t e);
extern std::atomic<int64_t> g_launches;

namespace {

__global__ void __launch_bounds__(256)
sym3_eigh_kernel(int64_t n, int64_t ld, const double* __restrict__ A6, double* __restrict__ w_out,
                 double* __restrict__ V_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double S[6], w[3], V[3][3];
#pragma unroll
    for (int c = 0; c < 6; ++c) S[c] = __ldg(A6 + c * ld + i);
    eig3_jacobi(S, w, V);
    // ascending order, vectors follow (three compare-exchanges)
    auto cswap = [&](int a, int b) {
        if (w[a] > w[b]) {
            const double t = w[a]; w[a] = w[b]; w[b] = t;
#pragma unroll
            for (int m = 0; m < 3; ++m) { const double v = V[m][a]; V[m][a] = V[m][b]; V[m][b] = v; }
        }
    };
    cswap(0, 1); cswap(1, 2); cswap(0, 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) __stcs(w_out + k * ld + i, w[k]);
    if (V_out) {
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int k = 0; k < 3; ++k) __stcs(V_out + (3 * m + k) * ld + i, V[m][k]);
    }
}

}  // namespace
}  // namespace cmadx

extern "C" int cmadx_sym3_eigh(int64_t n, int64_t ld, const double* A6, double* w, double* V, void* stream) {
    using namespace cmadx;
    if (n < 0 || ld < n || (n > 0 && (!A6 || !w))) return CMADX_EINVAL;
    if (n == 0) return CMADX_OK;
    sym3_eigh_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, ld, A6, w, V);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CMADX_OK;
}
