// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the closed-form Yld2004-18p routine of the
// CUDA kernels - cmad_b200/csrc/barlat.cuh, the very source nvcc compiles - built for the host so
// that its value, normal, Hessian and parameter derivatives can be checked against automatic
// differentiation of the oracle's restatement on a machine without a GPU.  Never linked into the
// product library.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cuda_runtime.h>

#include "host_shims.h"

#include "point_solver.cuh"

using namespace cmadx;

extern "C" {
// the Jacobi eigen-solver of the kernels (barlat.cuh, sym3_eigh.cu) on n packed tensors (row-major
// n x 6: xx,xy,xz,yy,yz,zz): eigenvalues in ascending order (n x 3), vectors as columns (n x 3 x 3)
void eig3_host(int64_t n, const double* A6, double* w, double* V) {
    for (int64_t i = 0; i < n; ++i) {
        double S[6], ww[3], VV[3][3];
        for (int c = 0; c < 6; ++c) S[c] = A6[6 * i + c];
        eig3_jacobi(S, ww, VV);
        int ord[3] = {0, 1, 2};
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b)
                if (ww[ord[b]] < ww[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
        for (int k = 0; k < 3; ++k) {
            w[3 * i + k] = ww[ord[k]];
            for (int m = 0; m < 3; ++m) V[9 * i + 3 * m + k] = VV[m][ord[k]];
        }
    }
}

// coeffs: 18 tensor coefficients + exponent; sig: xx,xy,xz,yy,yz,zz.
// out: phi, n[6], M[36] (row-major a,b), then per parameter (19, header order): dphi, dn[6]
int barlat_host_eval(const double* coeffs, const double* sig, double* out) {
    DevMat m;
    memset(&m, 0, sizeof m);
    for (int i = 0; i < 18; ++i) m.barlat[i] = coeffs[i];
    m.a = coeffs[18];
    m.inv_a = 1.0 / coeffs[18];
    YieldFn<CMADX_YIELD_BARLAT> yf;
    double s[6], n[6], phi;
    for (int i = 0; i < 6; ++i) s[i] = sig[i];
    yf.eval(m, s, phi, n);
    out[0] = phi;
    for (int a = 0; a < 6; ++a) out[1 + a] = n[a];
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) out[7 + 6 * a + b] = yf.M(a, b);
    for (int k = 0; k < 19; ++k) {
        double dphi = 0.0, dn[6] = {0, 0, 0, 0, 0, 0};
        if (!yf.dparam(m, CMADX_P_BARLAT_C0 + k, s, dphi, dn)) return 1;
        out[43 + 7 * k] = dphi;
        for (int a = 0; a < 6; ++a) out[44 + 7 * k + a] = dn[a];
    }
    return 0;
}
}
