// Per-material-point device routines of the constitutive update: residual,
// hand-derived Jacobian, register-resident 7x7 LU with partial pivoting, the
// reference-identical local Newton (traced + imperative flavours) with its
// quadratic backtracking line search, and the IFT / parameter-sensitivity
// outputs.  Everything lives in registers: one thread owns one point.
//
// Semantics follow the reference (sandialabs/cmad), file:line relative to its
// tree; nothing here is translated from it - the reference obtains every
// derivative by tracing JAX AD, this file derives them in closed form:
//   residual / branch select  cmad/models/small_elastic_plastic.py:238-302,
//                             cmad/models/paths.py:26-27
//   effective stresses        cmad/models/effective_stress.py:30-52,168-177
//   hardening                 cmad/models/hardening.py:9-34
//   traced Newton + IFT       cmad/models/nonlinear_solver.py:88-174
//   imperative Newton         cmad/models/nonlinear_solver.py:14-85
//   line search               cmad/util/line_search.py:74-85,95-189
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

#include "cmad_b200.h"

namespace cmadx {

struct DevMat {
    double lam, mu, two_mu, inv_two_mu;
    double Y, S, D, K;
    double hill[6];
    double a, inv_a;          // Hosford exponent and 1 / a (host-computed: the same correctly rounded quotient)
    double Q[9];
    double yield_tol;
    double barlat[18];        // Yld2004-18p: sp_12 .. sp_66, dp_12 .. dp_66 (a / inv_a hold its exponent)
    double dlam[2], dmu[2];   // d(lambda, mu)/d(elastic[0..1])
    double d2lam[3], d2mu[3]; // second derivatives (00, 01, 11): Hessian path only
    int hmask, rot, model, yield;
    int a_int;                // Hosford exponent when it is a small positive integer, else 0
    int root_int;             // = a_int when the outer root may skip libm pow (hosford_root), else 0
};

struct DevNewton {
    int mode, max_iters, ls_max, flags;
    int defer_after;   // > 0: a lane that needs more than this many Newton updates stops and is
                       // handed to the second (compacted) pass; 0: off.  Only kernels that own a
                       // second pass set it (from defer_request); everywhere else it stays 0
    int defer_request; // what the caller asked for (cmadx_newton_t flags bits 8..15)
    int defer_min;     // >= 0: a lane that needs a Newton direction after this many updates stops with its
                       // state intact (block-level hand-off, mp_update_cta.cu); -1: off
    double abs_tol, rel_tol, c1, bmin, bmax;
};

#define CMADX_DEV __device__ __forceinline__

// component bookkeeping for the packing xx,xy,xz,yy,yz,zz
CMADX_DEV constexpr bool is_diag(int a) { return a == 0 || a == 3 || a == 5; }
CMADX_DEV constexpr double mult(int a) { return is_diag(a) ? 1.0 : 2.0; }

// --------------------------------------------------------------------------
// Effective stresses.  Each provides, for a symmetric sigma (6 comps):
//   eval   : phi and the normal n_a = d phi / d sigma_a (single tensor entry,
//            as jax.grad over the full 3x3 gives it), keeping what hess needs
//   M(a,b) : d n_a / d(symmetric perturbation of component b)
//   dparam : d phi / d theta and d n / d theta for a yield-surface parameter
// All three surfaces are pressure-insensitive (n : I = 0, M : I = 0), which the
// Jacobian assembly below uses.
// --------------------------------------------------------------------------
template <int YK> struct YieldFn;

template <> struct YieldFn<CMADX_YIELD_J2> {
    double c;       // sqrt(3/2)/||s||
    double sn;      // ||s||
    double sh[6];   // s/||s||
    CMADX_DEV void eval(const DevMat&, const double (&sig)[6], double& phi, double (&n)[6]) {
        const double h = (sig[0] + sig[3] + sig[5]) / 3.0;
        double s[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) s[a] = is_diag(a) ? sig[a] - h : sig[a];
        double ss = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) ss = fma(mult(a) * s[a], s[a], ss);
        sn = sqrt(ss);
        const double r32 = 1.2247448713915890491;   // sqrt(3/2)
        phi = r32 * sn;
        const double inv = 1.0 / sn;                // inf at zero deviator -> NaN normal (as JAX)
        c = r32 * inv;
#pragma unroll
        for (int a = 0; a < 6; ++a) { sh[a] = s[a] * inv; n[a] = r32 * sh[a]; }
    }
    CMADX_DEV double M(int a, int b) const {
        double v = -sh[a] * sh[b] * mult(b);
        if (a == b) v += 1.0;
        if (is_diag(a) && is_diag(b)) v -= 1.0 / 3.0;
        return c * v;
    }
    // (dn/dsigma) : v in closed form: c (dev v - s^ (s^ : v))
    CMADX_DEV void Mvec(const double (&v)[6], double (&out)[6]) const {
        const double tr3 = (v[0] + v[3] + v[5]) * (1.0 / 3.0);
        double sv = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) sv = fma(mult(a) * sh[a], v[a], sv);
#pragma unroll
        for (int a = 0; a < 6; ++a) out[a] = c * fma(-sh[a], sv, is_diag(a) ? v[a] - tr3 : v[a]);
    }
    CMADX_DEV bool dparam(const DevMat&, int, const double (&)[6], double&, double (&)[6]) const {
        return false;
    }
    // state of the last evaluation as NS doubles (stride apart), and back - with the normal it implies
    static constexpr int NS = 8;
    CMADX_DEV void save_n(double*, int, const double (&)[6]) const {}
    CMADX_DEV void save(double* p, int st) const {
        p[0] = c; p[st] = sn;
#pragma unroll
        for (int a = 0; a < 6; ++a) p[(2 + a) * st] = sh[a];
    }
    CMADX_DEV void load(const DevMat&, const double* p, int st, double (&n)[6]) {
        c = p[0]; sn = p[st];
#pragma unroll
        for (int a = 0; a < 6; ++a) { sh[a] = p[(2 + a) * st]; n[a] = 1.2247448713915890491 * sh[a]; }
    }
};

template <> struct YieldFn<CMADX_YIELD_HILL> {
    double iphi;     // 1/phi
    double nn[6];    // normal
    double F, G, H, L, Mm, N;
    double d12, d20, d01;
    CMADX_DEV void eval(const DevMat& m, const double (&sig)[6], double& phi, double (&n)[6]) {
        F = m.hill[0]; G = m.hill[1]; H = m.hill[2]; L = m.hill[3]; Mm = m.hill[4]; N = m.hill[5];
        d12 = sig[3] - sig[5]; d20 = sig[5] - sig[0]; d01 = sig[0] - sig[3];
        const double q = F * d12 * d12 + G * d20 * d20 + H * d01 * d01
                         + L * (2.0 * sig[4] * sig[4]) + Mm * (2.0 * sig[2] * sig[2])
                         + N * (2.0 * sig[1] * sig[1]);
        phi = sqrt(q);
        iphi = 1.0 / phi;
        n[0] = (H * d01 - G * d20) * iphi;
        n[3] = (F * d12 - H * d01) * iphi;
        n[5] = (G * d20 - F * d12) * iphi;
        n[1] = N * sig[1] * iphi;
        n[2] = Mm * sig[2] * iphi;
        n[4] = L * sig[4] * iphi;
#pragma unroll
        for (int a = 0; a < 6; ++a) nn[a] = n[a];
    }
    CMADX_DEV double hq(int a, int b) const {   // (d2 q / d sigma_a d sym(b)) / 2
        if (a == 0 && b == 0) return G + H;
        if (a == 3 && b == 3) return F + H;
        if (a == 5 && b == 5) return F + G;
        if ((a == 0 && b == 3) || (a == 3 && b == 0)) return -H;
        if ((a == 0 && b == 5) || (a == 5 && b == 0)) return -G;
        if ((a == 3 && b == 5) || (a == 5 && b == 3)) return -F;
        if (a == 1 && b == 1) return N;
        if (a == 2 && b == 2) return Mm;
        if (a == 4 && b == 4) return L;
        return 0.0;
    }
    CMADX_DEV double M(int a, int b) const {
        return (hq(a, b) - nn[a] * nn[b] * mult(b)) * iphi;
    }
    CMADX_DEV bool dparam(const DevMat&, int pid, const double (&sig)[6], double& dphi,
                          double (&dn)[6]) const {
        if (pid < CMADX_P_HILL_F || pid > CMADX_P_HILL_N) return false;
        double dq = 0.0, dg[6] = {0, 0, 0, 0, 0, 0};   // dq/dtheta, d(grad q)/dtheta
        switch (pid) {
        case CMADX_P_HILL_F: dq = d12 * d12; dg[3] = 2.0 * d12; dg[5] = -2.0 * d12; break;
        case CMADX_P_HILL_G: dq = d20 * d20; dg[5] = 2.0 * d20; dg[0] = -2.0 * d20; break;
        case CMADX_P_HILL_H: dq = d01 * d01; dg[0] = 2.0 * d01; dg[3] = -2.0 * d01; break;
        case CMADX_P_HILL_L: dq = 2.0 * sig[4] * sig[4]; dg[4] = 2.0 * sig[4]; break;
        case CMADX_P_HILL_M: dq = 2.0 * sig[2] * sig[2]; dg[2] = 2.0 * sig[2]; break;
        default:             dq = 2.0 * sig[1] * sig[1]; dg[1] = 2.0 * sig[1]; break;
        }
        dphi = 0.5 * dq * iphi;
#pragma unroll
        for (int a = 0; a < 6; ++a) dn[a] = (0.5 * dg[a] - nn[a] * dphi) * iphi;
        return true;
    }
    static constexpr int NS = 10;
    CMADX_DEV void save_n(double*, int, const double (&)[6]) const {}
    CMADX_DEV void save(double* p, int st) const {
        p[0] = iphi; p[st] = d12; p[2 * st] = d20; p[3 * st] = d01;
#pragma unroll
        for (int a = 0; a < 6; ++a) p[(4 + a) * st] = nn[a];
    }
    CMADX_DEV void load(const DevMat& m, const double* p, int st, double (&n)[6]) {
        F = m.hill[0]; G = m.hill[1]; H = m.hill[2]; L = m.hill[3]; Mm = m.hill[4]; N = m.hill[5];
        iphi = p[0]; d12 = p[st]; d20 = p[2 * st]; d01 = p[3 * st];
#pragma unroll
        for (int a = 0; a < 6; ++a) { nn[a] = p[(4 + a) * st]; n[a] = nn[a]; }
    }
};

// r^a for r >= 0: square-and-multiply when the exponent is a small positive integer
// (a = 4 in the reference's tests, 100 in examples/notch_hosford.yaml; the branch is
// uniform over the launch), the libm pow otherwise.  Agrees with pow to a few ulp.
CMADX_DEV double hosford_pow(double r, double a, int a_int) {
    if (a_int <= 0) return pow(r, a);
    double res = 1.0, b = r;
    for (int e = a_int; e != 0; e >>= 1) {
        if (e & 1) res *= b;
        b *= b;
    }
    return res;
}

// s^(1/a) for s > 0, the outer root of the Hosford norm.  libm's pow was 13 % of all instructions
// of the a = 4 kernel (ncu, profiles/r2g_k1_hosford_4_flops.txt): two correctly rounded square
// roots for a = 4, one for a = 2, exp(log(s) / a) for the other integer exponents (|log s| < 15 and
// 1/a <= 1/3: the argument error is far below one ulp of the result, which then carries exp's own
// rounding), libm pow otherwise.  Same value as pow to an ulp or two; iteration counts of every
// parity fixture unchanged.
CMADX_DEV double hosford_root(double s, double inv_a, int a_int) {
    if (a_int == 4) return sqrt(sqrt(s));
    if (a_int == 2) return sqrt(s);
    if (a_int > 2) return exp(inv_a * log(s));
    return pow(s, inv_a);
}

// Hosford: phi = vm * (1/2 sum_i |Delta_i/vm|^a)^(1/a) on the *diagonal* stress
// entries only (the reference's documented limitation).  The vm scaling cancels
// analytically (phi is the plain a-norm of the differences); it is kept for the
// value so large exponents do not overflow, and dropped from the derivatives.
template <> struct YieldFn<CMADX_YIELD_HOSFORD> {
    double g[3];       // d phi / d Delta_i
    double w[3];       // 1/2 r_i^(a-2)
    double iphi, am1;
    CMADX_DEV void eval(const DevMat& m, const double (&sig)[6], double& phi, double (&n)[6]) {
        const double a = m.a;
        const double h = (sig[0] + sig[3] + sig[5]) / 3.0;
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double s = is_diag(k) ? sig[k] - h : sig[k];
            ss = fma(mult(k) * s, s, ss);
        }
        const double vm = 1.2247448713915890491 * sqrt(ss);
        const double ivm = 1.0 / vm;
        const double dl[3] = {sig[0] - sig[3], sig[3] - sig[5], sig[5] - sig[0]};
        double q[3], sq = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { q[i] = hosford_pow(fabs(dl[i] * ivm), a, m.a_int); sq += q[i]; }
        sq *= 0.5;
        phi = vm * hosford_root(sq, m.inv_a, m.root_int);
        iphi = 1.0 / phi;
        am1 = a - 1.0;
        const double isq = 1.0 / sq;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double r = fabs(dl[i]) * iphi;            // |Delta_i|/phi
            const double ra = q[i] * isq;                   // r^a
            const double t = (dl[i] > 0.0) - (dl[i] < 0.0); // sign, 0 at 0
            const double ir = (r > 0.0) ? 1.0 / r : 0.0;
            g[i] = 0.5 * t * ra * ir;                       // 1/2 t r^(a-1)
            // 1/2 r^(a-2).  Rounded here once and for all (__dmul_rn is never contracted into the
            // adds of Gd()): the value is part of the state another thread may restore from shared
            // memory (mp_update_cta.cu), and both flows must see the same bits
            w[i] = __dmul_rn(0.5 * ra * ir, ir);
        }
        n[0] = g[0] - g[2]; n[3] = g[1] - g[0]; n[5] = g[2] - g[1];
        n[1] = 0.0; n[2] = 0.0; n[4] = 0.0;
    }
    // G_ij = d2 phi / dDelta_i dDelta_j = (a-1)/phi (delta_ij w_i - g_i g_j)
    CMADX_DEV double Gd(int i, int j) const {
        double v = -g[i] * g[j];
        if (i == j) v += w[i];
        return am1 * iphi * v;
    }
    CMADX_DEV double M(int a, int b) const {
        if (!is_diag(a) || !is_diag(b)) return 0.0;
        // Delta = B sigma_diag, B rows: (1,-1,0),(0,1,-1),(-1,0,1); M = B^T G B
        const int ia = (a == 0) ? 0 : (a == 3 ? 1 : 2);
        const int ib = (b == 0) ? 0 : (b == 3 ? 1 : 2);
        // column ia of B: +1 at row ia, -1 at row (ia+2)%3
        const int pa = ia, ma = (ia + 2) % 3, pb = ib, mb = (ib + 2) % 3;
        return Gd(pa, pb) - Gd(pa, mb) - Gd(ma, pb) + Gd(ma, mb);
    }
    // d(phi, n)/d(exponent a) at the last evaluated state (cmad/models/effective_stress.py:168-177
    // differentiated by jacrev over the `a` leaf, cmad/models/model.py:126-133).  With
    // r_i = |Delta_i| / phi (so 1/2 sum r_i^a = 1) and L = 1/2 sum r_i^a ln r_i:
    //   d phi / da = phi L / a,    d g_i / da = g_i (ln r_i - (a - 1) L / a),   g_i = 1/2 t_i r_i^(a-1)
    // (the vm scaling of the reference's formula cancels analytically; terms with r_i = 0 vanish).
    CMADX_DEV bool dparam(const DevMat& m, int pid, const double (&sig)[6], double& dphi, double (&dn)[6]) const {
        if (pid != CMADX_P_HOSFORD_A) return false;
        const double dl[3] = {sig[0] - sig[3], sig[3] - sig[5], sig[5] - sig[0]};
        double lr[3], L = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double r = fabs(dl[i]) * iphi;
            lr[i] = (r > 0.0) ? log(r) : 0.0;
            L = fma(fabs(g[i]) * r, lr[i], L);          // 1/2 r^a ln r = |g| r ln r
        }
        const double a = m.a;
        dphi = L / (a * iphi);
        double dgi[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) dgi[i] = g[i] * (lr[i] - am1 * L / a);
        dn[0] = dgi[0] - dgi[2]; dn[3] = dgi[1] - dgi[0]; dn[5] = dgi[2] - dgi[1];
        dn[1] = 0.0; dn[2] = 0.0; dn[4] = 0.0;
        return true;
    }
    // the normal is saved as evaluated (n = g_i - g_j may have been contracted with the product
    // that defines g_i: recomputing it from the rounded g would differ in the last bit)
    static constexpr int NS = 10;
    CMADX_DEV void save_n(double* p, int st, const double (&n)[6]) const {
        p[7 * st] = n[0]; p[8 * st] = n[3]; p[9 * st] = n[5];
    }
    CMADX_DEV void save(double* p, int st) const {
#pragma unroll
        for (int i = 0; i < 3; ++i) { p[i * st] = g[i]; p[(3 + i) * st] = w[i]; }
        p[6 * st] = iphi;
    }
    CMADX_DEV void load(const DevMat& m, const double* p, int st, double (&n)[6]) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { g[i] = p[i * st]; w[i] = p[(3 + i) * st]; }
        iphi = p[6 * st];
        am1 = m.a - 1.0;
        n[0] = p[7 * st]; n[3] = p[8 * st]; n[5] = p[9 * st];
        n[1] = 0.0; n[2] = 0.0; n[4] = 0.0;
    }
};

}  // namespace cmadx
#include "barlat.cuh"
namespace cmadx {

// --------------------------------------------------------------------------
// register-resident dense LU.
//
// The reference solves the local system with LAPACK-style partial pivoting
// (jnp.linalg.solve, cmad/models/nonlinear_solver.py:123).  With the matrix
// in registers a row exchange costs a sweep of predicated selects per
// candidate row, which tripled the instruction count of the whole kernel
// (profiles/r1_k1_v1: 37 % FSEL).  The plastic Jacobian
// [[I + dgamma*2mu*H*W, -n], [-(W n)^T, -H'/2mu]] (H = yield-surface Hessian,
// PSD; W = diag(1,2,2,1,2,1)) is a column-scaled SPD block bordered by the
// consistency row, for which elimination in natural order is backward stable
// whenever dgamma >= 0.  So: *threshold* pivoting - eliminate in natural order,
// flag a lane as "troubled" if any natural pivot is more than 10x smaller than
// an entry below it (or is NaN), and only troubled lanes re-do the solve with
// full partial pivoting (factor_pivot / solve_pivot).  Both paths are stable
// solves of the same system; they differ at rounding level only.
// a[k][k] holds 1/pivot after factorisation.
// --------------------------------------------------------------------------
// partial-pivoting LU of an n x n matrix in (local) memory; returns the swap bits in
// the order RegLU::solve_pivot replays them; a[k][k] holds 1/pivot afterwards
// (n <= 11: 55 decisions fit one word; n = 12 - the rate form under uniaxial stress - has 66: the
// decisions past bit 63 go to *hi)
static __device__ __noinline__ unsigned long long lu_factor_pivot_mem(double* a, int n, unsigned long long* hi = nullptr) {
    unsigned long long swaps = 0ull;
    if (hi) *hi = 0ull;
    int bit = 0;
    for (int k = 0; k < n; ++k) {
        for (int i = k + 1; i < n; ++i) {
            const bool sw = fabs(a[i * n + k]) > fabs(a[k * n + k]);
            if (sw) {
                for (int j = 0; j < n; ++j) {
                    const double u = a[k * n + j];
                    a[k * n + j] = a[i * n + j];
                    a[i * n + j] = u;
                }
            }
            if (bit < 64) swaps |= (sw ? 1ull : 0ull) << bit;
            else if (hi) *hi |= (sw ? 1ull : 0ull) << (bit - 64);
            ++bit;
        }
        const double rp = 1.0 / a[k * n + k];
        a[k * n + k] = rp;
        for (int i = k + 1; i < n; ++i) {
            const double l = a[i * n + k] * rp;
            a[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) a[i * n + j] = fma(-l, a[k * n + j], a[i * n + j]);
        }
    }
    return swaps;
}

template <int N>
struct RegLU {
    double a[N][N];
    unsigned long long swaps;   // N (N - 1) / 2 row-exchange decisions (36 for N = 9)
    unsigned long long swaps_hi;   // decisions 64.. (N = 12: 66 in all); unused for N <= 11
    static_assert(N * (N - 1) / 2 <= 128, "RegLU: swap record holds 128 decisions");

    // natural-order elimination; returns true when this lane needs pivoting
    CMADX_DEV bool factor_natural() {
        bool trouble = false;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double thr = 10.0 * fabs(a[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) trouble = trouble || !(fabs(a[i][k]) <= thr);
            const double rp = 1.0 / a[k][k];
            a[k][k] = rp;
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double l = a[i][k] * rp;
                a[i][k] = l;
#pragma unroll
                for (int j = k + 1; j < N; ++j) a[i][j] = fma(-l, a[k][j], a[i][j]);
            }
        }
        return trouble;
    }
    CMADX_DEV void solve_natural(double (&b)[N]) const {
#pragma unroll
        for (int i = 1; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < i; ++j) b[i] = fma(-a[i][j], b[j], b[i]);
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            double s = b[i];
#pragma unroll
            for (int j = i + 1; j < N; ++j) s = fma(-a[i][j], b[j], s);
            b[i] = s * a[i][i];
        }
    }

    // full partial pivoting (rows bubble so that row k holds the column
    // maximum, first maximum wins); the swap decisions are recorded so
    // right-hand sides can be permuted later.  Rare path: done in local memory by
    // one out-of-line routine shared by every kernel (keeps the hot code small).
    CMADX_DEV void factor_pivot() {
        double t[N * N];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) t[i * N + j] = a[i][j];
        swaps_hi = 0ull;
        swaps = (N * (N - 1) / 2 > 64) ? lu_factor_pivot_mem(t, N, &swaps_hi) : lu_factor_pivot_mem(t, N);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) a[i][j] = t[i * N + j];
    }
    CMADX_DEV void solve_pivot(double (&b)[N]) const {
        int bit = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (bit < 64) ? ((swaps >> bit) & 1ull) : ((swaps_hi >> (bit - 64)) & 1ull);
                const double u = b[k], v = b[i];
                b[k] = sw ? v : u;
                b[i] = sw ? u : v;
                ++bit;
            }
        }
        solve_natural(b);
    }
};

// delta = J(x)^{-1} rhs for the point's current state
template <class Pt, int N>
CMADX_DEV void newton_direction(const DevMat& m, const Pt& pt, double dg, double (&dx)[N]) {
    double rhs[N];
#pragma unroll
    for (int i = 0; i < N; ++i) rhs[i] = dx[i];
    bool trouble;
    {
        RegLU<N> lu;
        pt.jacobian(m, dg, lu.a);
        trouble = lu.factor_natural();
        lu.solve_natural(dx);
    }
    if (__any_sync(__activemask(), trouble)) {
        if (trouble) {
            RegLU<N> lu;
            pt.jacobian(m, dg, lu.a);
            lu.factor_pivot();
#pragma unroll
            for (int i = 0; i < N; ++i) dx[i] = rhs[i];
            lu.solve_pivot(dx);
        }
    }
}

// --------------------------------------------------------------------------
// SmallElasticPlastic, FULL_3D.  State x = [ep(6), alpha], strain `em` already
// in material axes.
// --------------------------------------------------------------------------
template <int YK>
struct SepPoint {
    static constexpr int N = 7, ALPHA = 6;
    YieldFn<YK> yf;
    double n[6];      // yield normal at the last evaluated state
    double f, eD;     // yield function, exp(-D alpha)
    bool plastic;

    CMADX_DEV void residual(const DevMat& m, const double (&x)[7], const double (&xp)[7],
                            const double (&em)[6], double (&C)[7]) {
        double ee[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = em[a] - x[a];
        const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
        double sig[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
        double phi;
        yf.eval(m, sig, phi, n);
        double Hd = 0.0;
        eD = 0.0;
        if (m.hmask & CMADX_HARD_VOCE) { eD = exp(-m.D * x[6]); Hd = m.S * (1.0 - eD); }
        if (m.hmask & CMADX_HARD_LINEAR) Hd = fma(m.K, x[6], Hd);
        f = (phi - (m.Y + Hd)) * m.inv_two_mu;
        const double dg = x[6] - xp[6];
        plastic = (f > m.yield_tol) || (fabs(f) < m.yield_tol);   // paths.py:26
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double ce = x[a] - xp[a];
            C[a] = plastic ? fma(-dg, n[a], ce) : ce;
        }
        C[6] = plastic ? f : dg;
    }

    // dC/dx at the last evaluated state (dg = alpha - alpha_prev there)
    CMADX_DEV void jacobian(const DevMat& m, double dg, double (&J)[7][7]) const {
        if (plastic) {
            const double s = dg * m.two_mu;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = 0; b < 6; ++b) J[a][b] = fma(s, yf.M(a, b), (a == b) ? 1.0 : 0.0);
                J[a][6] = -n[a];
                J[6][a] = -mult(a) * n[a];
            }
            double Hp = 0.0;
            if (m.hmask & CMADX_HARD_VOCE) Hp = m.S * m.D * eD;
            if (m.hmask & CMADX_HARD_LINEAR) Hp += m.K;
            J[6][6] = -Hp * m.inv_two_mu;
        } else {
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int b = 0; b < 7; ++b) J[a][b] = (a == b) ? 1.0 : 0.0;
        }
    }
};

// the yield-surface state of a point (what the Jacobian and the derivative outputs read after a
// residual evaluation) as NS doubles `st` apart: lets another thread continue or finish the point
template <class Pt>
CMADX_DEV void save_point_state(const Pt& pt, double* p, int st) {
    pt.yf.save(p, st);
    pt.yf.save_n(p, st, pt.n);
    p[decltype(pt.yf)::NS * st] = pt.f;
    p[(decltype(pt.yf)::NS + 1) * st] = pt.eD;
}
template <class Pt>
CMADX_DEV void load_point_state(const DevMat& m, Pt& pt, const double* p, int st, bool plastic) {
    pt.yf.load(m, p, st, pt.n);
    pt.f = p[decltype(pt.yf)::NS * st];
    pt.eD = p[(decltype(pt.yf)::NS + 1) * st];
    pt.plastic = plastic;
}

// position in the full 7-vector [ep(6), alpha] of local unknown k
template <int YK> struct SepPointTraits {
    CMADX_DEV static constexpr int full(int k) { return k; }
    CMADX_DEV static constexpr int local(int c) { return c; }      // -1: not an unknown
};

// --------------------------------------------------------------------------
// Hosford, reduced: the reference's Hosford surface only sees the DIAGONAL stress
// entries (effective_stress.py:167-177), so the yield normal and its derivative
// vanish on the shear components: started from x0 = xi_prev the shear rows of the
// residual are identically zero and those rows/columns of the Jacobian are the
// identity.  Natural-order elimination of the 7x7 system then performs exactly the
// operations of the 4x4 system in [ep_xx, ep_yy, ep_zz, alpha] (the other
// multipliers are exact zeros), so this point type yields the same iterates, norms,
// iteration counts and flags as SepPoint<HOSFORD> - with a 16-entry LU.
// Valid when the starting iterate equals xi_prev on the shear components.
// --------------------------------------------------------------------------
struct HosfordPoint {
    static constexpr int N = 4, ALPHA = 3;
    YieldFn<CMADX_YIELD_HOSFORD> yf;
    double n[6];
    double f, eD;
    double shear[3];   // ep_xy, ep_xz, ep_yz: frozen
    bool plastic;

    CMADX_DEV void residual(const DevMat& m, const double (&x)[4], const double (&xp)[4],
                            const double (&em)[6], double (&C)[4]) {
        double ee[6];
        ee[0] = em[0] - x[0]; ee[3] = em[3] - x[1]; ee[5] = em[5] - x[2];
        ee[1] = em[1] - shear[0]; ee[2] = em[2] - shear[1]; ee[4] = em[4] - shear[2];
        const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
        double sig[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
        double phi;
        yf.eval(m, sig, phi, n);
        double Hd = 0.0;
        eD = 0.0;
        if (m.hmask & CMADX_HARD_VOCE) { eD = exp(-m.D * x[3]); Hd = m.S * (1.0 - eD); }
        if (m.hmask & CMADX_HARD_LINEAR) Hd = fma(m.K, x[3], Hd);
        f = (phi - (m.Y + Hd)) * m.inv_two_mu;
        const double dg = x[3] - xp[3];
        plastic = (f > m.yield_tol) || (fabs(f) < m.yield_tol);   // paths.py:26
        const double nd[3] = {n[0], n[3], n[5]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double ce = x[k] - xp[k];
            C[k] = plastic ? fma(-dg, nd[k], ce) : ce;
        }
        C[3] = plastic ? f : dg;
    }

    CMADX_DEV void jacobian(const DevMat& m, double dg, double (&J)[4][4]) const {
        constexpr int D[3] = {0, 3, 5};
        if (plastic) {
            const double s = dg * m.two_mu;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int b = 0; b < 3; ++b) J[a][b] = fma(s, yf.M(D[a], D[b]), (a == b) ? 1.0 : 0.0);
                J[a][3] = -n[D[a]];
                J[3][a] = -n[D[a]];
            }
            double Hp = 0.0;
            if (m.hmask & CMADX_HARD_VOCE) Hp = m.S * m.D * eD;
            if (m.hmask & CMADX_HARD_LINEAR) Hp += m.K;
            J[3][3] = -Hp * m.inv_two_mu;
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) J[a][b] = (a == b) ? 1.0 : 0.0;
        }
    }
};

struct HosfordTraits {
    CMADX_DEV static constexpr int full(int k) { return k == 0 ? 0 : (k == 1 ? 3 : (k == 2 ? 5 : 6)); }
    CMADX_DEV static constexpr int local(int c) { return c == 0 ? 0 : (c == 3 ? 1 : (c == 5 ? 2 : (c == 6 ? 3 : -1))); }
};

// full 7x7 dC/dx from a point's state (any point type with yf, n, eD, plastic)
template <class Pt>
CMADX_DEV double full_jacobian_entry(const DevMat& m, const Pt& pt, double dg, int r, int c) {
    if (!pt.plastic) return (r == c) ? 1.0 : 0.0;
    if (r < 6 && c < 6) return fma(dg * m.two_mu, pt.yf.M(r, c), (r == c) ? 1.0 : 0.0);
    if (r < 6) return -pt.n[r];
    if (c < 6) return -mult(c) * pt.n[c];
    double Hp = 0.0;
    if (m.hmask & CMADX_HARD_VOCE) Hp = m.S * m.D * pt.eD;
    if (m.hmask & CMADX_HARD_LINEAR) Hp += m.K;
    return -Hp * m.inv_two_mu;
}

template <int N> CMADX_DEV double normN(const double (&v)[N]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) s = fma(v[i], v[i], s);
    return (s == 0.0) ? 0.0 : sqrt(s);       // sqrt(0) = 0 without the special-case subroutine (elastic lanes)
}
template <int N> CMADX_DEV double dotN(const double (&u)[N], const double (&v)[N]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) s = fma(u[i], v[i], s);
    return s;
}

// Elastic model (cmad/models/elastic.py:139-173): state x = cauchy(6),
// C = vec6(x - sigma_el(eps)) / (2 mu), sigma_el = kappa tr(eps) I + 2 mu dev(eps).
struct ElasticPoint {
    static constexpr int N = 6, ALPHA = 5;
    bool plastic;
    CMADX_DEV void residual(const DevMat& m, const double (&x)[6], const double (&)[6],
                            const double (&em)[6], double (&C)[6]) {
        const double tr = em[0] + em[3] + em[5];
        const double kappa = m.lam + 2.0 * m.mu / 3.0;
        const double kt = kappa * tr;
        const double t3 = tr / 3.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double se = is_diag(a) ? fma(m.two_mu, em[a] - t3, kt) : m.two_mu * em[a];
            C[a] = (x[a] - se) * m.inv_two_mu;
        }
        plastic = false;
    }
    CMADX_DEV void jacobian(const DevMat& m, double, double (&J)[6][6]) const {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) J[a][b] = (a == b) ? m.inv_two_mu : 0.0;
    }
};

struct NewtonResult {
    int iters;
    int flag_entry;
    double cnorm;
    bool deferred;   // stopped by DevNewton::defer_after before converging: outputs are not valid
};

// warp-aggregated append to a device list: one atomicAdd per warp instead of one per lane
// (a third of an a = 100 Hosford batch is deferred: millions of appends to one counter)
CMADX_DEV void list_append(bool want, unsigned* count, int* list, unsigned cap, int value) {
    const unsigned active = __activemask();
    const unsigned m = __ballot_sync(active, want);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0u;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(active, base, leader);
    if (want) {
        const unsigned slot = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        if (slot < cap) list[slot] = value;
    }
}

// Local Newton of one point as a resumable per-lane state machine.
//
// Both Newton flavours are written as ONE routine around a SINGLE residual call
// site, driven by a per-lane phase: the residual (with its pow / exp / sqrt) is
// by far the largest piece of code, and five inlined copies of it made the
// kernels instruction-cache bound (ncu: 26 % of the stalls "no instruction").
// Lanes in different phases share the evaluation instead of serialising.
// `trip()` = one residual evaluation plus everything the reference does between
// that evaluation and the next one; the sequence of evaluations, tests and
// updates of a lane is exactly that of the reference loops
// (nonlinear_solver.py:102-155 and :14-85, line_search.py:125-181).
// Two drivers use it: local_newton() below (a warp iterates until its slowest
// lane is done) and the streaming kernel (mp_update_stream.cu), where a lane
// that finishes takes the next point while its neighbours keep iterating.
//
// State carried between trips is kept minimal (it is what a lane holds in
// registers while its neighbours work): the iterate x, the Newton direction dx
// and six scalars.  In particular
//   * the trial point is recomputed as x - al dx instead of being stored;
//   * the residual at x is never stored: every trip that needs it has just
//     evaluated it (a line search that runs out of probes moves x to its best
//     probe and RE-EVALUATES there - the reference's best_aux, bit for bit, since
//     the same routine sees the same arguments - which is also the evaluation that
//     makes the yield-surface state fresh at x for the next Jacobian);
//   * phi0, phi0' and the Armijo slope are functions of C.C alone.
// The one visible difference to the reference: if NO probe of a line search is
// finite, the reference continues with the stale residual of the previous
// iterate at a non-finite x; here the residual is the (non-finite) one at x.
// Either way the point runs to max_iters and returns a non-finite state.
// Pt provides residual(m, x, xp, em, C), jacobian(m, dgamma, J) and `plastic`.
template <class Pt, int N>
struct NewtonLane {
    // PH_DIR: resume at "take a direction"; PH_LEGACY: a probe of the imperative flavour's legacy
    // line search (newton_solve(max_ls_evals > 0), cmad/models/nonlinear_solver.py:55-81)
    enum { PH_INIT = 0, PH_PROBE = 1, PH_EVAL = 2, PH_DIR = 3, PH_LEGACY = 4 };
    double x[N];                 // current iterate
    double dx[N];                // Newton direction of the running line search / last step
    double n0, nc;               // ||C|| at x0 and at the last convergence test
    double al, best_al, best_phi, CC;   // line search: trial step, best so far, C.C at x
    int phase, ii, ne, flag_entry;
    bool active;                 // this lane wants another residual evaluation
    bool deferred;               // stopped by DevNewton::defer_after before converging

    CMADX_DEV void start(const double (&x0)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) { x[i] = x0[i]; dx[i] = 0.0; }
        n0 = 0.0; nc = 0.0;
        al = 1.0; best_al = 1.0; best_phi = CUDART_INF; CC = 0.0;
        phase = PH_INIT; ii = 0; ne = 0; flag_entry = 0;
        active = true; deferred = false;
    }

    // One evaluation.  `Ct` returns the residual just evaluated; when the lane finishes in this
    // trip (`active` turns false) it is the residual at the returned x and `pt` is fresh there.
    // `live == false`: evaluate the entry state only (padding lanes of the one-pass kernels).
    CMADX_DEV void trip(const DevMat& m, const DevNewton& nw, Pt& pt, const double (&xp)[N],
                        const double (&em)[6], bool live, double (&Ct)[N]) {
        const bool traced = (nw.mode == CMADX_NEWTON_TRACED);
        if (phase != PH_DIR) {   // PH_DIR: (Ct, pt) at x were restored by the caller, nothing to evaluate
            double xt[N];
#pragma unroll
            for (int i = 0; i < N; ++i) xt[i] = (phase == PH_PROBE) ? fma(-al, dx[i], x[i]) : x[i];
            pt.residual(m, xt, xp, em, Ct);                // the only call site
        }
        bool need_dir = (phase == PH_DIR);   // (x, Ct) current and pt fresh at x: take a Newton step
        if (phase == PH_DIR) {
        } else if (phase == PH_INIT) {
            flag_entry = pt.plastic ? 1 : 0;
            n0 = normN<N>(Ct);
            nc = n0;
            if (!live || nw.max_iters <= 0) {
                active = false;
            } else {
                // nc == n0 here: n0 / n0 is 1 for finite non-zero n0 and NaN otherwise (0/0, inf/inf:
                // the test is then false) - spelled out so elastic lanes (n0 = 0) skip the
                // division's special-case subroutine
                const double rel = !traced ? 1.0 : ((n0 > 0.0 && n0 < CUDART_INF) ? 1.0 : CUDART_NAN);
                if (rel < nw.rel_tol || nc < nw.abs_tol) active = false; else need_dir = true;
            }
        } else {
            bool test = (phase == PH_EVAL);                  // x is where Ct was evaluated
            if (phase == PH_LEGACY) {
                // x already sits at x_k + al dx.  psi_j >= (1 - 2 beta al) psi_0 (false for NaN):
                // next al = max(eta al, -al^2 psi_0' / (2 (psi_j - psi_0 - al psi_0'))), psi_0' = -2 psi_0,
                // beta = 1e-4, eta = 0.5; at jj == max_ls_evals the loop breaks WITHOUT moving x.
                // The accepted (or last) evaluation is the one the next convergence test would
                // repeat at the same x, so it is reused (identical values).
                const double psi0 = 0.5 * CC, dpsi0 = -2.0 * psi0;          // CC = ||C(x_k)||^2 via the norm
                const double cj = normN<N>(Ct);
                const double psij = 0.5 * cj * cj;
                if (psij >= (1.0 - 2.0 * 1e-4 * al) * psi0) {
                    const double an = fmax(0.5 * al, -(al * al * dpsi0) / (2.0 * (psij - psi0 - al * dpsi0)));
                    if (ne == nw.ls_max) {
                        test = true;                                         // "reached max ls evals"
                    } else {
                        ++ne;
                        const double da = an - al;
#pragma unroll
                        for (int i = 0; i < N; ++i) x[i] = fma(da, dx[i], x[i]);
                        al = an;
                    }
                } else {
                    test = true;
                }
            }
            if (phase == PH_PROBE) {
                // ---- line search (quadratic model), line_search.py:125-181
                const double phi0 = 0.5 * CC, dphi0 = -CC, armijo = nw.c1 * dphi0;
                const double ph = 0.5 * dotN<N>(Ct, Ct);
                const bool fin = isfinite(ph);
                if (fin && ph < best_phi) { best_al = al; best_phi = ph; }
                const bool acc = fin && (ph <= fma(al, armijo, phi0));
                const double den = 2.0 * (ph - phi0 - dphi0 * al);
                const double am = (den == 0.0) ? 0.5 * al : -dphi0 * al * al / den;
                double ac = fmin(fmax(am, nw.bmin * al), nw.bmax * al);
                if (am != am) ac = am;                    // clip propagates NaN
                ++ne;
                if (acc) {
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] = fma(-al, dx[i], x[i]);
                    ++ii;
                    test = true;                          // the accepted probe is the new (x, C)
                } else if (ne < nw.ls_max) {
                    al = fin ? ac : 0.5 * al;             // next probe
                } else {
                    // out of probes: the lowest-merit step; its residual (the reference's best_aux)
                    // and the yield-surface state at it come from re-evaluating there
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] = fma(-best_al, dx[i], x[i]);
                    ++ii;
                    phase = PH_EVAL;
                }
            }
            if (test) {
                if (phase == PH_LEGACY) phase = PH_EVAL;
                const bool by_iters = ii >= nw.max_iters;
                // imperative flavour at max_iters: newton_solve returns the norm of its last test
                if (traced || !by_iters) nc = normN<N>(Ct);
                bool stop = by_iters;
                if (!stop) {
                    const double rel = nc / n0;
                    stop = (rel < nw.rel_tol || nc < nw.abs_tol);
                }
                if (stop) active = false; else need_dir = true;
            }
        }
        // Two-pass divergence control of the one-pass kernels: a lane that still needs a
        // direction after `defer_after` updates stops here; the kernel appends it to a list and
        // a second pass re-solves those points in warps made of hard points only.
        if (need_dir && phase != PH_DIR &&
            ((nw.defer_after > 0 && ii >= nw.defer_after) || (nw.defer_min >= 0 && ii >= nw.defer_min))) {
            need_dir = false;
            active = false;
            deferred = true;
        }
        if (need_dir) {
            if (traced) {
#pragma unroll
                for (int i = 0; i < N; ++i) dx[i] = Ct[i];
                newton_direction<Pt, N>(m, pt, x[Pt::ALPHA] - xp[Pt::ALPHA], dx);   // solve(J, C)
                CC = dotN<N>(Ct, Ct);
                ne = 0; al = 1.0; best_al = 1.0; best_phi = CUDART_INF;
                phase = PH_PROBE;
            } else {
                // imperative newton_solve (no line search)
#pragma unroll
                for (int i = 0; i < N; ++i) dx[i] = -Ct[i];
                newton_direction<Pt, N>(m, pt, x[Pt::ALPHA] - xp[Pt::ALPHA], dx);   // solve(J, -C)
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] += dx[i];
                ++ii;
                if (nw.ls_max > 0) {          // legacy line search: probes start at the full step
                    CC = nc * nc;
                    al = 1.0; ne = 1;
                    phase = PH_LEGACY;
                } else {
                    phase = PH_EVAL;
                }
            }
        }
    }
};

// Local Newton for one point; `live` lanes take part, the loop exit is decided
// warp-wide by ballot so the whole warp leaves together.  On return x is the
// solution, C the residual there, and pt holds the state (n, f, plastic, yield
// internals) at x.
// CTA_SYNC: the loop exit is voted by the whole thread block (every thread of the block must call):
// the warps of a block then execute the same stretch of code at the same time, which is what the
// instruction caches want (the generic kernels are 70-120 KB of straight-line code and were
// losing a quarter of their issue slots to "no instruction" stalls with unsynchronised warps).
template <class Pt, int N, bool CTA_SYNC = false>
CMADX_DEV NewtonResult local_newton(const DevMat& m, const DevNewton& nw, Pt& pt,
                                    double (&x)[N], const double (&xp)[N],
                                    const double (&em)[6], bool live, double (&C)[N]) {
    NewtonLane<Pt, N> L;
    L.start(x);
#pragma unroll
    for (int i = 0; i < N; ++i) C[i] = 0.0;
    const unsigned full = 0xffffffffu;
    while (CTA_SYNC ? (__syncthreads_or(L.active ? 1 : 0) != 0) : (__any_sync(full, L.active) != 0)) {
        if (L.active) L.trip(m, nw, pt, xp, em, live, C);   // the finishing trip leaves C at x
    }
    NewtonResult r;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = L.x[i];
    r.flag_entry = L.flag_entry;
    r.deferred = L.deferred;
    r.iters = L.ii;
    r.cnorm = L.nc;
    return r;
}

}  // namespace cmadx
