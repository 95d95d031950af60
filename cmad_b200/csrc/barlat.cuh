// Yld2004-18p ("barlat") effective stress with closed-form first and second derivatives.
//
// Reference semantics (sandialabs/cmad, file:line relative to its tree; nothing is translated -
// the reference differentiates `jnp.linalg.eigh` by JAX AD, this file derives the derivatives):
//   cmad/models/effective_stress.py:55-84          parameter leaves sp_*, dp_*, a
//   cmad/verification/functions.py:71-99           the two linear maps L', L'' of the stress
//   cmad/verification/functions.py:129-154         phi = (1/4 sum_ij |S'_i - S''_j|^a)^(1/a),
//                                                  S'_i / S''_j eigenvalues of L' sigma / L'' sigma
//
// With e = (S'_1..3, S''_1..3), d_ij = S'_i - S''_j, r_ij = |d_ij| / phi (1/4 sum r^a = 1):
//   g = d phi / d e :  g'_i = 1/4 sum_j sgn(d_ij) r_ij^(a-1),   g''_j = -1/4 sum_i sgn(d_ij) r_ij^(a-1)
//   d2 phi / d e d e = (a-1)/phi (W - g g^T),  W = 1/4 [[diag(sum_j w_ij), -w], [-w^T, diag(sum_i w_ij)]],
//                                              w_ij = r_ij^(a-2)
// and for a symmetric tensor T with eigenpairs (t_i, v_i) and a symmetric perturbation dT:
//   d t_i = v_i^T dT v_i,     d2 t_i = 2 sum_{k != i} (v_i^T dT v_k)^2 / (t_i - t_k)
// so in the coordinates z_b = [diag(V'^T dS'_b V'), diag(V''^T dS''_b V''), offdiag(..'), offdiag(..'')]
// of the unit symmetric stress perturbation b (dS_b = L E_b):
//   d phi / d sym_b         = g . z_b[0..5]
//   d2 phi / d sym_a d sym_b = z_a^T blockdiag(d2phi/de de, 2 theta) z_b,
//   theta_ik = (g_i - g_k) / (t_i - t_k)   per tensor (pairs 01, 02, 12).
// At coincident eigenvalues JAX's eigh rule returns inf / NaN; here theta takes its limit
// d2phi/de_i de_i - d2phi/de_i de_k (phi is a symmetric function of each triple), so the Jacobian
// stays finite on the states every uniaxial test along a symmetry axis passes through.
// The eigen-decomposition is a cyclic Jacobi iteration in registers (quadratic convergence, a few
// sweeps): eigenvalues to an ulp of the norm, orthonormal vectors also for close eigenvalues.
#pragma once

namespace cmadx {

// A = V diag(w) V^T for the symmetric 3x3 with packed entries xx,xy,xz,yy,yz,zz; V[m][i] is
// component m of eigenvector i
template <int P, int Q, int R>
CMADX_DEV void jacobi_rotate(double (&A)[3][3], double (&V)[3][3]) {
    const double apq = A[P][Q];
    if (apq == 0.0) return;
    const double th = (A[Q][Q] - A[P][P]) / (2.0 * apq);
    const double t = copysign(1.0, th) / (fabs(th) + sqrt(fma(th, th, 1.0)));
    const double c = 1.0 / sqrt(fma(t, t, 1.0)), s = t * c;
    A[P][P] = fma(-t, apq, A[P][P]);
    A[Q][Q] = fma(t, apq, A[Q][Q]);
    A[P][Q] = 0.0; A[Q][P] = 0.0;
    const double arp = A[R][P], arq = A[R][Q];
    A[R][P] = c * arp - s * arq; A[P][R] = A[R][P];
    A[R][Q] = s * arp + c * arq; A[Q][R] = A[R][Q];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const double vp = V[m][P], vq = V[m][Q];
        V[m][P] = c * vp - s * vq;
        V[m][Q] = s * vp + c * vq;
    }
}

CMADX_DEV void eig3_jacobi(const double (&S)[6], double (&w)[3], double (&V)[3][3]) {
    double A[3][3] = {{S[0], S[1], S[2]}, {S[1], S[3], S[4]}, {S[2], S[4], S[5]}};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        const double dia = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (!(off > 1e-36 * dia)) break;
        jacobi_rotate<0, 1, 2>(A, V);
        jacobi_rotate<0, 2, 1>(A, V);
        jacobi_rotate<1, 2, 0>(A, V);
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

// rows of the normal block of L (times 3) per unit coefficient: c12 c13 | c21 c23 | c31 c32
CMADX_DEV constexpr double barlat_wv(int k, int col) {
    // k = 0: [1,-2,1]  1: [1,1,-2]  2: [-2,1,1]  3: [1,1,-2]  4: [-2,1,1]  5: [1,-2,1]
    return (k == 0 || k == 5) ? (col == 1 ? -2.0 : 1.0)
         : (k == 1 || k == 3) ? (col == 2 ? -2.0 : 1.0)
                              : (col == 0 ? -2.0 : 1.0);
}

template <> struct YieldFn<CMADX_YIELD_BARLAT> {
    // state of the last evaluation
    double V[2][3][3];     // eigenvectors of S' and S''
    double t[2][3];        // eigenvalues
    double g[6];           // d phi / d e
    double He[6][6];       // d2 phi / d e d e
    double th[6];          // theta: S' pairs 01 02 12, S'' pairs 01 02 12
    double tq[9];          // sgn(d_ij) r_ij^(a-1)
    double lr[9];          // ln r_ij (0 where r = 0)
    double Lr;             // 1/4 sum r^a ln r
    double phi_;
    double U[2][3][3];     // normal blocks of L', L''
    double cs[2][3];       // shear coefficients: xy (c44), yz (c55), xz (c66)
    double Z[6][12];       // z_b
    double Mm[6][6];       // d n_a / d sym_b

    CMADX_DEV static constexpr int dslot(int a) { return a == 0 ? 0 : (a == 3 ? 1 : 2); }   // diagonal comp -> 0..2
    // off-diagonal component a couples tensor axes (p, q); its coefficient slot in cs
    CMADX_DEV static constexpr int op(int a) { return a == 4 ? 1 : 0; }
    CMADX_DEV static constexpr int oq(int a) { return a == 1 ? 1 : 2; }
    CMADX_DEV static constexpr int oslot(int a) { return a == 1 ? 0 : (a == 4 ? 1 : 2); }

    CMADX_DEV void maps(const DevMat& m) {
#pragma unroll
        for (int T = 0; T < 2; ++T) {
            const double* c = m.barlat + 9 * T;
            const double k3 = 1.0 / 3.0;
            U[T][0][0] = (c[0] + c[1]) * k3;        U[T][0][1] = (-2.0 * c[0] + c[1]) * k3; U[T][0][2] = (c[0] - 2.0 * c[1]) * k3;
            U[T][1][0] = (-2.0 * c[2] + c[3]) * k3; U[T][1][1] = (c[2] + c[3]) * k3;        U[T][1][2] = (c[2] - 2.0 * c[3]) * k3;
            U[T][2][0] = (-2.0 * c[4] + c[5]) * k3; U[T][2][1] = (c[4] - 2.0 * c[5]) * k3;  U[T][2][2] = (c[4] + c[5]) * k3;
            cs[T][0] = c[6]; cs[T][1] = c[7]; cs[T][2] = c[8];
        }
    }

    // V^T D V for D = diag(d0,d1,d2) (OFF = false) or D = c (e_p e_q^T + e_q e_p^T): the three diagonal
    // entries, then the pairs 01, 02, 12
    CMADX_DEV void proj_diag(int T, double d0, double d1, double d2, double (&out)[6]) const {
        const int pi[6] = {0, 1, 2, 0, 0, 1}, pk[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            const int i = pi[e], k = pk[e];
            out[e] = d0 * V[T][0][i] * V[T][0][k] + d1 * V[T][1][i] * V[T][1][k] + d2 * V[T][2][i] * V[T][2][k];
        }
    }
    CMADX_DEV void proj_off(int T, int p, int q, double c, double (&out)[6]) const {
        const int pi[6] = {0, 1, 2, 0, 0, 1}, pk[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            const int i = pi[e], k = pk[e];
            out[e] = c * (V[T][p][i] * V[T][q][k] + V[T][q][i] * V[T][p][k]);
        }
    }
    // z^T K y for 12-vectors in the layout [diag', diag'', off', off'']
    CMADX_DEV double quad(const double (&za)[12], const double (&zb)[12]) const {
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            double r = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) r = fma(He[p][q], zb[q], r);
            s = fma(za[p], r, s);
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) s = fma(2.0 * th[r] * za[6 + r], zb[6 + r], s);
        return s;
    }

    __device__ __noinline__ void eval_impl(const DevMat& m, const double* sig, double* phi_out, double* n) {
        const double a = m.a;
        maps(m);
        // the two images of the stress and their spectra
#pragma unroll
        for (int T = 0; T < 2; ++T) {
            double S[6];
            S[0] = U[T][0][0] * sig[0] + U[T][0][1] * sig[3] + U[T][0][2] * sig[5];
            S[3] = U[T][1][0] * sig[0] + U[T][1][1] * sig[3] + U[T][1][2] * sig[5];
            S[5] = U[T][2][0] * sig[0] + U[T][2][1] * sig[3] + U[T][2][2] * sig[5];
            S[1] = cs[T][0] * sig[1]; S[4] = cs[T][1] * sig[4]; S[2] = cs[T][2] * sig[2];
            eig3_jacobi(S, t[T], V[T]);
        }
        // phi, scaled by the largest difference so large exponents neither overflow nor vanish
        double d[9], dmax = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { d[3 * i + j] = t[0][i] - t[1][j]; dmax = fmax(dmax, fabs(d[3 * i + j])); }
        if (!(dmax > 0.0)) {                 // zero image: phi = 0, derivatives undefined (NaN, as JAX AD)
            phi_ = 0.0; *phi_out = 0.0;
            const double q = nan("");
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                n[b] = q;
#pragma unroll
                for (int c = 0; c < 6; ++c) Mm[b][c] = q;
            }
            return;
        }
        double pw[9], psi = 0.0;
#pragma unroll
        for (int e = 0; e < 9; ++e) { pw[e] = pow(fabs(d[e]) / dmax, a); psi += pw[e]; }
        psi *= 0.25;
        const double phi = dmax * pow(psi, m.inv_a);
        phi_ = phi;
        *phi_out = phi;
        const double iphi = 1.0 / phi;
        double w2[9];
        Lr = 0.0;
#pragma unroll
        for (int e = 0; e < 9; ++e) {
            const double r = fabs(d[e]) * iphi;
            const double ra = pw[e] / psi;                   // r^a
            const bool nz = r > 0.0;
            tq[e] = nz ? copysign(ra / r, d[e]) : 0.0;
            w2[e] = nz ? ra / (r * r) : (a == 2.0 ? 1.0 : 0.0);     // r^(a-2): 0^0 = 1
            lr[e] = nz ? log(r) : 0.0;
            Lr = fma(0.25 * ra, lr[e], Lr);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g[i] = 0.25 * (tq[3 * i] + tq[3 * i + 1] + tq[3 * i + 2]);
            g[3 + i] = -0.25 * (tq[i] + tq[3 + i] + tq[6 + i]);
        }
        const double k = (a - 1.0) * iphi;
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double wv = 0.0;
                if (p == q) wv = (p < 3) ? 0.25 * (w2[3 * p] + w2[3 * p + 1] + w2[3 * p + 2])
                                         : 0.25 * (w2[p - 3] + w2[p] + w2[p + 3]);
                else if (p < 3 && q >= 3) wv = -0.25 * w2[3 * p + (q - 3)];
                else if (p >= 3 && q < 3) wv = -0.25 * w2[3 * q + (p - 3)];
                He[p][q] = k * (wv - g[p] * g[q]);
            }
        // theta, with its limit at (numerically) coincident eigenvalues
        const double scale = fmax(fabs(t[0][0]) + fabs(t[0][1]) + fabs(t[0][2]), fabs(t[1][0]) + fabs(t[1][1]) + fabs(t[1][2]));
        const int pi[3] = {0, 0, 1}, pk[3] = {1, 2, 2};
#pragma unroll
        for (int T = 0; T < 2; ++T)
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const int i = 3 * T + pi[e], kk = 3 * T + pk[e];
                const double dt = t[T][pi[e]] - t[T][pk[e]];
                th[3 * T + e] = (fabs(dt) > 1e-8 * scale) ? (g[i] - g[kk]) / dt
                                                          : 0.5 * (He[i][i] + He[kk][kk]) - He[i][kk];
            }
        // z_b of the six unit symmetric stress perturbations
#pragma unroll
        for (int b = 0; b < 6; ++b) {
#pragma unroll
            for (int T = 0; T < 2; ++T) {
                double o[6];
                if (is_diag(b)) proj_diag(T, U[T][0][dslot(b)], U[T][1][dslot(b)], U[T][2][dslot(b)], o);
                else proj_off(T, op(b), oq(b), cs[T][oslot(b)], o);
#pragma unroll
                for (int e = 0; e < 3; ++e) { Z[b][3 * T + e] = o[e]; Z[b][6 + 3 * T + e] = o[3 + e]; }
            }
        }
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            double s = 0.0;
#pragma unroll
            for (int p = 0; p < 6; ++p) s = fma(g[p], Z[b][p], s);
            n[b] = s / mult(b);
        }
#pragma unroll
        for (int a2 = 0; a2 < 6; ++a2)
#pragma unroll
            for (int b = a2; b < 6; ++b) {
                const double h = quad(Z[a2], Z[b]);
                Mm[a2][b] = h / mult(a2);
                Mm[b][a2] = h / mult(b);
            }
    }

    CMADX_DEV void eval(const DevMat& m, const double (&sig)[6], double& phi, double (&n)[6]) {
        eval_impl(m, sig, &phi, n);
    }
    CMADX_DEV double M(int a, int b) const { return Mm[a][b]; }

    // d(phi, n)/d(theta) for the 18 tensor coefficients and the exponent, at the last evaluated state
    __device__ __noinline__ bool dparam_impl(const DevMat& m, int pid, const double* sig, double* dphi_out, double* dn) const {
        if (pid < CMADX_P_BARLAT_C0 || pid > CMADX_P_BARLAT_A) return false;
        if (pid == CMADX_P_BARLAT_A) {
            const double a = m.a;
            *dphi_out = phi_ * Lr / a;
            const double shift = (a - 1.0) * Lr / a;
            double dg[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                dg[i] = 0.25 * (tq[3 * i] * (lr[3 * i] - shift) + tq[3 * i + 1] * (lr[3 * i + 1] - shift)
                                + tq[3 * i + 2] * (lr[3 * i + 2] - shift));
                dg[3 + i] = -0.25 * (tq[i] * (lr[i] - shift) + tq[3 + i] * (lr[3 + i] - shift)
                                     + tq[6 + i] * (lr[6 + i] - shift));
            }
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double s = 0.0;
#pragma unroll
                for (int p = 0; p < 6; ++p) s = fma(dg[p], Z[b][p], s);
                dn[b] = s / mult(b);
            }
            return true;
        }
        const int idx = pid - CMADX_P_BARLAT_C0, T = idx / 9, k = idx % 9;
        double zt[12], o[6];
#pragma unroll
        for (int e = 0; e < 12; ++e) zt[e] = 0.0;
        // D = (dL/dc_k) sigma and its projection
        const int row = k >> 1;                       // k < 6: the normal-block row the coefficient sits in
        double wv[3] = {0.0, 0.0, 0.0};
        int p = 0, q = 1, oc = 1;                     // k >= 6: tensor axes and stress component of the shear entry
        if (k < 6) {
#pragma unroll
            for (int c = 0; c < 3; ++c) wv[c] = barlat_wv(k, c) * (1.0 / 3.0);
            const double dv = wv[0] * sig[0] + wv[1] * sig[3] + wv[2] * sig[5];
            proj_diag(T, row == 0 ? dv : 0.0, row == 1 ? dv : 0.0, row == 2 ? dv : 0.0, o);
        } else {
            if (k == 7) { p = 1; q = 2; oc = 4; } else if (k == 8) { p = 0; q = 2; oc = 2; }
            proj_off(T, p, q, sig[oc], o);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) { zt[3 * T + e] = o[e]; zt[6 + 3 * T + e] = o[3 + e]; }
        double s0 = 0.0;
#pragma unroll
        for (int e = 0; e < 6; ++e) s0 = fma(g[e], zt[e], s0);
        *dphi_out = s0;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            double s = quad(Z[b], zt);
            // explicit dependence of dS_b = L E_b on the coefficient: g . diag(V^T (dL/dc E_b) V)
            if (k < 6) {
                if (is_diag(b)) {
                    const double dv = wv[dslot(b)];
#pragma unroll
                    for (int i = 0; i < 3; ++i) s = fma(g[3 * T + i] * dv, V[T][row][i] * V[T][row][i], s);
                }
            } else if (b == oc) {
#pragma unroll
                for (int i = 0; i < 3; ++i) s = fma(g[3 * T + i], 2.0 * V[T][p][i] * V[T][q][i], s);
            }
            dn[b] = s / mult(b);
        }
        return true;
    }
    CMADX_DEV bool dparam(const DevMat& m, int pid, const double (&sig)[6], double& dphi, double (&dn)[6]) const {
        return dparam_impl(m, pid, sig, &dphi, dn);
    }

    // hand-over of a point between threads (block hand-off / parking kernels): not carried
    static constexpr int NS = 0;
    CMADX_DEV void save_n(double*, int, const double (&)[6]) const {}
    CMADX_DEV void save(double*, int) const {}
    CMADX_DEV void load(const DevMat&, const double*, int, double (&)[6]) {}
};

}  // namespace cmadx
