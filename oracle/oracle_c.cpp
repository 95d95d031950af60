// CPU oracle (C++17 + OpenMP) for CMAD's per-integration-point constitutive
// update.  TEST INFRASTRUCTURE ONLY: linked/executed only by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
//
// It restates the reference algorithm (file:line below are relative to the
// sandialabs/cmad tree) with *forward-mode dual numbers* standing in for
// jax.jacfwd / jacrev / grad, so every derivative here is AD-derived and
// independent of the hand-derived CUDA kernels it checks.
//
// Parity status: validated against oracle/cmad_oracle.py (torch.func AD) and
// the reference's known answers KA1..KA4 (tests/test_oracle_*.py).  Parity
// against actual JAX output is unpinned (JAX cannot be installed here).
//
// Reference map:
//   residual            cmad/models/small_elastic_plastic.py:33-92,238-302
//   branch select       cmad/models/paths.py:26-27
//   effective stress    cmad/models/effective_stress.py:30-52,168-177
//   hardening           cmad/models/hardening.py:9-34
//   elasticity          cmad/models/elastic_stress.py:14-21,24-40,71-72
//                       cmad/models/elastic_constants.py:54-104
//   Elastic model       cmad/models/elastic.py:139-173,188-195
//   traced Newton+IFT   cmad/models/nonlinear_solver.py:88-174
//   imperative Newton   cmad/models/nonlinear_solver.py:14-85
//   line search         cmad/util/line_search.py:74-85,95-189
//   packing             cmad/models/var_types.py:43-47,75-76
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------ duals
template <class T, int N>
struct Dual {
    T v;
    T d[N];
    Dual() : v(T(0)) { for (int i = 0; i < N; ++i) d[i] = T(0); }
    Dual(double c) : v(T(c)) { for (int i = 0; i < N; ++i) d[i] = T(0); }
    template <class U = T, class = std::enable_if_t<!std::is_same<U, double>::value>>
    Dual(const T& c) : v(c) { for (int i = 0; i < N; ++i) d[i] = T(0); }
};

inline double primal(double x) { return x; }
template <class T, int N> inline double primal(const Dual<T, N>& x) { return primal(x.v); }

#define DUAL_BIN(op, expr_v, expr_d)                                              \
    template <class T, int N>                                                     \
    inline Dual<T, N> operator op(const Dual<T, N>& a, const Dual<T, N>& b) {     \
        Dual<T, N> r; r.v = expr_v;                                               \
        for (int i = 0; i < N; ++i) r.d[i] = expr_d;                              \
        return r;                                                                 \
    }
DUAL_BIN(+, a.v + b.v, a.d[i] + b.d[i])
DUAL_BIN(-, a.v - b.v, a.d[i] - b.d[i])
DUAL_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
#undef DUAL_BIN
template <class T, int N>
inline Dual<T, N> operator/(const Dual<T, N>& a, const Dual<T, N>& b) {
    Dual<T, N> r; r.v = a.v / b.v;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r;
}
// mixed with plain double constants
template <class T, int N> inline Dual<T, N> operator+(const Dual<T, N>& a, double b) { return a + Dual<T, N>(b); }
template <class T, int N> inline Dual<T, N> operator+(double a, const Dual<T, N>& b) { return Dual<T, N>(a) + b; }
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a, double b) { return a - Dual<T, N>(b); }
template <class T, int N> inline Dual<T, N> operator-(double a, const Dual<T, N>& b) { return Dual<T, N>(a) - b; }
template <class T, int N> inline Dual<T, N> operator*(const Dual<T, N>& a, double b) { return a * Dual<T, N>(b); }
template <class T, int N> inline Dual<T, N> operator*(double a, const Dual<T, N>& b) { return Dual<T, N>(a) * b; }
template <class T, int N> inline Dual<T, N> operator/(const Dual<T, N>& a, double b) { return a / Dual<T, N>(b); }
template <class T, int N> inline Dual<T, N> operator/(double a, const Dual<T, N>& b) { return Dual<T, N>(a) / b; }

inline double dsqrt(double x) { return std::sqrt(x); }
inline double dexp(double x) { return std::exp(x); }
inline double dabs(double x) { return std::fabs(x); }
inline double dlog(double x) { return std::log(x); }
inline double dsign(double x) { return (x > 0.) - (x < 0.); }
inline double dpow(double x, double y) { return std::pow(x, y); }

template <class T, int N> inline Dual<T, N> dsqrt(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = dsqrt(a.v);
    T g = T(0.5) / r.v;                       // inf at 0 -> NaN tangents, as in JAX
    for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i];
    return r;
}
template <class T, int N> inline Dual<T, N> dexp(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = dexp(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i];
    return r;
}
template <class T, int N> inline Dual<T, N> dabs(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = dabs(a.v);
    double s = dsign(primal(a.v));            // jnp.abs tangent: sign(x), 0 at 0
    for (int i = 0; i < N; ++i) r.d[i] = T(s) * a.d[i];
    return r;
}
template <class T, int N> inline Dual<T, N> dlog(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = dlog(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / a.v;
    return r;
}
// x ** y with JAX's pow JVP conventions (lhs: y*x^(y-1) with y==0 guarded;
// rhs: log(where(x==0,1,x)) * ans).
template <class T, int N> inline Dual<T, N> dpow(const Dual<T, N>& x, const Dual<T, N>& y) {
    Dual<T, N> r; r.v = dpow(x.v, y.v);
    T ym1 = (primal(y.v) == 0.) ? T(1.) : y.v - T(1.);
    T gl = y.v * dpow(x.v, ym1);
    T xs = (primal(x.v) == 0.) ? T(1.) : x.v;
    T gr = dlog(xs) * r.v;
    for (int i = 0; i < N; ++i) r.d[i] = gl * x.d[i] + gr * y.d[i];
    return r;
}

// ------------------------------------------------------------ problem desc
enum { YIELD_J2 = 0, YIELD_HILL = 1, YIELD_HOSFORD = 2 };
enum { MODEL_SEP = 0, MODEL_ELASTIC = 1 };
// elastic pairs, values given in sorted-key order ("E"<"kappa"<"lambda"<"mu"<"nu")
enum { EP_E_NU = 0, EP_E_MU, EP_E_KAPPA, EP_E_LAMBDA, EP_KAPPA_MU, EP_KAPPA_NU,
       EP_KAPPA_LAMBDA, EP_LAMBDA_MU, EP_LAMBDA_NU, EP_MU_NU };
// canonical parameter ids (columns of dC/dp are requested by id)
enum { PID_EL0 = 0, PID_EL1, PID_Y, PID_VOCE_S, PID_VOCE_D, PID_LIN_K,
       PID_HILL_F, PID_HILL_G, PID_HILL_H, PID_HILL_L, PID_HILL_M, PID_HILL_N,
       PID_HOSFORD_A, PID_Q00, NUM_PID = PID_Q00 + 9 };

// flat material vector layout (doubles), indexed by canonical parameter id
// mat[NUM_PID] = yield_tol
// integer config: cfg[0]=model cfg[1]=yield cfg[2]=elastic pair cfg[3]=hardening
// mask (1 voce | 2 linear) cfg[4]=newton mode (0 traced, 1 imperative)
// cfg[5]=max_iters cfg[6]=ls max evals cfg[7]=strain components (6|9)
// solver doubles: sol[0]=abs_tol sol[1]=rel_tol sol[2]=c1 sol[3]=bmin sol[4]=bmax

template <class T>
struct Mat {
    T el0, el1, Y, S, D, K, hill[6], a, Q[9];
    int yield, pair, hmask;
    double yield_tol;
};

template <class T> inline void lame(const Mat<T>& m, T& lam, T& mu) {
    const T& p = m.el0; const T& q = m.el1;   // sorted-key order
    switch (m.pair) {                           // elastic_constants.py:68-103
    case EP_E_NU:        lam = p * q / ((1. + q) * (1. - 2. * q)); mu = p / (2. * (1. + q)); break;
    case EP_E_MU:        mu = q; lam = q * (p - 2. * q) / (3. * q - p); break;
    case EP_E_KAPPA:     mu = 3. * q * p / (9. * q - p); lam = 3. * q * (3. * q - p) / (9. * q - p); break;
    case EP_E_LAMBDA:  { lam = q; T R = dsqrt(p * p + 9. * q * q + 2. * p * q); mu = (p - 3. * q + R) / 4.; break; }
    case EP_KAPPA_MU:    mu = q; lam = p - 2. * q / 3.; break;
    case EP_KAPPA_NU:    mu = 3. * p * (1. - 2. * q) / (2. * (1. + q)); lam = 3. * p * q / (1. + q); break;
    case EP_KAPPA_LAMBDA: lam = q; mu = 3. * (p - q) / 2.; break;
    case EP_LAMBDA_MU:   lam = p; mu = q; break;
    case EP_LAMBDA_NU:   lam = p; mu = p * (1. - 2. * q) / (2. * q); break;
    case EP_MU_NU:       mu = p; lam = 2. * p * q / (1. - 2. * q); break;
    default:             lam = T(0.); mu = T(0.);
    }
}

template <class T> inline T trace3(const T* A) { return A[0] + A[4] + A[8]; }

// J2 effective stress, effective_stress.py:30-37
template <class T> inline T phi_j2(const T* c) {
    T h = trace3(c) / 3.;
    T s[9];
    for (int i = 0; i < 9; ++i) s[i] = c[i];
    s[0] = s[0] - h; s[4] = s[4] - h; s[8] = s[8] - h;
    T ss = T(0.);
    for (int i = 0; i < 9; ++i) ss = ss + s[i] * s[i];
    return std::sqrt(3. / 2.) * dsqrt(ss);
}
// Hill, effective_stress.py:40-52
template <class T, class P> inline T phi_hill(const T* c, const P* h) {
    T d12 = c[4] - c[8], d20 = c[8] - c[0], d01 = c[0] - c[4];
    return dsqrt(T(h[0]) * d12 * d12 + T(h[1]) * d20 * d20 + T(h[2]) * d01 * d01
                 + T(h[3]) * (c[7] * c[7] + c[5] * c[5])
                 + T(h[4]) * (c[6] * c[6] + c[2] * c[2])
                 + T(h[5]) * (c[3] * c[3] + c[1] * c[1]));
}
// Hosford, effective_stress.py:167-177 (diagonal entries only)
template <class T, class P> inline T phi_hosford(const T* c, const P& a) {
    T vm = phi_j2(c);
    T s0 = c[0] / vm, s1 = c[4] / vm, s2 = c[8] / vm;
    T A = T(a);
    T d01 = dpow(dabs(s0 - s1), A);
    T d12 = dpow(dabs(s1 - s2), A);
    T d20 = dpow(dabs(s2 - s0), A);
    return vm * dpow(0.5 * (d01 + d12 + d20), dpow(A, T(-1.)));
}

template <class T> inline T phi_of(const Mat<T>& m, const T* c) {
    if (m.yield == YIELD_J2) return phi_j2(c);
    if (m.yield == YIELD_HILL) return phi_hill(c, m.hill);
    return phi_hosford(c, m.a);
}

// yield normal = jax.grad(effective_stress)(cauchy) (small_elastic_plastic.py:90).
// J2 and Hill: the first derivative is written out (it is validated against the
// torch.func AD oracle in tests/test_oracle_cross.py); all *second* derivatives
// still come from the dual numbers carried by T.  Hosford: one 9-direction dual
// pass over the cauchy entries, nested over T.
template <class T> inline void phi_and_normal(const Mat<T>& m, const T* c, T& phi, T* n) {
    if (m.yield == YIELD_J2) {
        T h = trace3(c) / 3.;
        T s[9];
        for (int i = 0; i < 9; ++i) s[i] = c[i];
        s[0] = s[0] - h; s[4] = s[4] - h; s[8] = s[8] - h;
        T ss = T(0.);
        for (int i = 0; i < 9; ++i) ss = ss + s[i] * s[i];
        T sn = dsqrt(ss);
        phi = std::sqrt(3. / 2.) * sn;
        T g = std::sqrt(3. / 2.) / sn;
        T tr = T(0.);
        for (int i = 0; i < 9; ++i) n[i] = g * s[i];
        tr = (n[0] + n[4] + n[8]) / 3.;          // chain rule through s = c - tr(c)/3 I
        n[0] = n[0] - tr; n[4] = n[4] - tr; n[8] = n[8] - tr;
        return;
    }
    if (m.yield == YIELD_HILL) {
        const T* h = m.hill;
        T d12 = c[4] - c[8], d20 = c[8] - c[0], d01 = c[0] - c[4];
        phi = phi_hill(c, m.hill);
        T ip = 1. / phi;
        n[0] = (h[2] * d01 - h[1] * d20) * ip;
        n[4] = (h[0] * d12 - h[2] * d01) * ip;
        n[8] = (h[1] * d20 - h[0] * d12) * ip;
        n[1] = h[5] * c[1] * ip; n[3] = h[5] * c[3] * ip;
        n[2] = h[4] * c[2] * ip; n[6] = h[4] * c[6] * ip;
        n[5] = h[3] * c[5] * ip; n[7] = h[3] * c[7] * ip;
        return;
    }
    typedef Dual<T, 9> D9;
    D9 cc[9];
    for (int i = 0; i < 9; ++i) { cc[i].v = c[i]; cc[i].d[i] = T(1.); }
    D9 p = phi_hosford(cc, m.a);
    phi = p.v;
    for (int i = 0; i < 9; ++i) n[i] = p.d[i];
}

template <class T> inline T hardening(const Mat<T>& m, const T& alpha) {   // hardening.py:9-34
    T H = T(0.);
    if (m.hmask & 1) H = H + m.S * (1. - dexp(-(m.D * alpha)));
    if (m.hmask & 2) H = H + m.K * alpha;
    return H;
}

// sym(grad_u) rotated to material axes minus plastic strain -> stress
template <class T>
inline void material_cauchy(const Mat<T>& m, const T* x, const T* gu, T* sig) {
    T eps[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) eps[3 * i + j] = 0.5 * (gu[3 * i + j] + gu[3 * j + i]);
    // Q^T eps Q  (small_elastic_plastic.py:61-62)
    T tmp[9], em[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0.);
            for (int k = 0; k < 3; ++k) s = s + eps[3 * i + k] * m.Q[3 * k + j];
            tmp[3 * i + j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0.);
            for (int k = 0; k < 3; ++k) s = s + m.Q[3 * k + i] * tmp[3 * k + j];
            em[3 * i + j] = s;
        }
    T ep[9] = {x[0], x[1], x[2], x[1], x[3], x[4], x[2], x[4], x[5]};      // var_types.py:43-47
    T ee[9];
    for (int i = 0; i < 9; ++i) ee[i] = em[i] - ep[i];
    T lam, mu; lame(m, lam, mu);
    T tr = trace3(ee);
    for (int i = 0; i < 9; ++i) sig[i] = 2. * mu * ee[i];
    sig[0] = sig[0] + lam * tr; sig[4] = sig[4] + lam * tr; sig[8] = sig[8] + lam * tr;
}

// SmallElasticPlastic FULL_3D residual; returns branch flag
template <class T>
inline bool sep_residual(const Mat<T>& m, const T* x, const T* xp, const T* gu, T* C) {
    T sig[9]; material_cauchy(m, x, gu, sig);
    T phi, n[9]; phi_and_normal(m, sig, phi, n);
    T lam, mu; lame(m, lam, mu);
    T f = (phi - (m.Y + hardening(m, x[6]))) / (2. * mu);
    T dg = x[6] - xp[6];
    double fv = primal(f);
    bool plastic = (fv > m.yield_tol) || (std::fabs(fv) < m.yield_tol);   // paths.py:26
    static const int up[6] = {0, 1, 2, 4, 5, 8};                           // var_types.py:75-76
    if (plastic) {
        for (int a = 0; a < 6; ++a) C[a] = (x[a] - xp[a]) - dg * n[up[a]];
        C[6] = f;
    } else {
        for (int a = 0; a < 6; ++a) C[a] = x[a] - xp[a];
        C[6] = dg;
    }
    return plastic;
}

template <class T>
inline void sep_cauchy(const Mat<T>& m, const T* x, const T* gu, T* sg) {  // :308-321
    T sig[9]; material_cauchy(m, x, gu, sig);
    T tmp[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0.);
            for (int k = 0; k < 3; ++k) s = s + m.Q[3 * i + k] * sig[3 * k + j];
            tmp[3 * i + j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0.);
            for (int k = 0; k < 3; ++k) s = s + tmp[3 * i + k] * m.Q[3 * j + k];
            sg[3 * i + j] = s;
        }
}

// Elastic model (xi = cauchy vec6), elastic.py:139-173
template <class T>
inline bool elastic_residual(const Mat<T>& m, const T* x, const T* xp, const T* gu, T* C) {
    T lam, mu; lame(m, lam, mu);
    T kappa = lam + 2. * mu / 3.;
    T eps[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) eps[3 * i + j] = 0.5 * (gu[3 * i + j] + gu[3 * j + i]);
    T tr = trace3(eps);
    T se[9];
    for (int i = 0; i < 9; ++i) se[i] = 2. * mu * eps[i];
    for (int i = 0; i < 3; ++i) se[4 * i] = 2. * mu * (eps[4 * i] - tr / 3.) + kappa * tr;
    static const int up[6] = {0, 1, 2, 4, 5, 8};
    for (int a = 0; a < 6; ++a) C[a] = (x[a] - se[up[a]]) / (2. * mu);
    (void)xp;
    return false;
}

template <class T>
inline bool residual(int model, const Mat<T>& m, const T* x, const T* xp, const T* gu, T* C) {
    return model == MODEL_SEP ? sep_residual(m, x, xp, gu, C) : elastic_residual(m, x, xp, gu, C);
}
template <class T>
inline void cauchy(int model, const Mat<T>& m, const T* x, const T* gu, T* sg) {
    if (model == MODEL_SEP) { sep_cauchy(m, x, gu, sg); return; }
    T t[9] = {x[0], x[1], x[2], x[1], x[3], x[4], x[2], x[4], x[5]};        // elastic.py:188-195
    for (int i = 0; i < 9; ++i) sg[i] = t[i];
}

template <class T> inline Mat<T> make_mat(const double* mat, const int* cfg) {
    Mat<T> m;
    m.el0 = T(mat[PID_EL0]); m.el1 = T(mat[PID_EL1]); m.Y = T(mat[PID_Y]);
    m.S = T(mat[PID_VOCE_S]); m.D = T(mat[PID_VOCE_D]); m.K = T(mat[PID_LIN_K]);
    for (int i = 0; i < 6; ++i) m.hill[i] = T(mat[PID_HILL_F + i]);
    m.a = T(mat[PID_HOSFORD_A]);
    for (int i = 0; i < 9; ++i) m.Q[i] = T(mat[PID_Q00 + i]);
    m.yield = cfg[1]; m.pair = cfg[2]; m.hmask = cfg[3];
    m.yield_tol = mat[NUM_PID];
    return m;
}
template <class T> inline T& mat_slot(Mat<T>& m, int pid) {
    switch (pid) {
    case PID_EL0: return m.el0; case PID_EL1: return m.el1; case PID_Y: return m.Y;
    case PID_VOCE_S: return m.S; case PID_VOCE_D: return m.D; case PID_LIN_K: return m.K;
    case PID_HOSFORD_A: return m.a;
    default:
        if (pid >= PID_HILL_F && pid <= PID_HILL_N) return m.hill[pid - PID_HILL_F];
        return m.Q[pid - PID_Q00];
    }
}

// dense LU with partial pivoting (LAPACK getrf/getrs order of operations)
inline void lu_solve(int n, double* A, double* b, int nrhs) {
    // A row-major n x n, b row-major n x nrhs; in-place solve
    for (int k = 0; k < n; ++k) {
        int p = k; double best = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = std::fabs(A[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        if (p != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int j = 0; j < nrhs; ++j) { double t = b[k * nrhs + j]; b[k * nrhs + j] = b[p * nrhs + j]; b[p * nrhs + j] = t; }
        }
        double piv = A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double l = A[i * n + k] / piv;
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        }
    }
    for (int c = 0; c < nrhs; ++c) {
        for (int i = 1; i < n; ++i) {
            double s = b[i * nrhs + c];
            for (int j = 0; j < i; ++j) s -= A[i * n + j] * b[j * nrhs + c];
            b[i * nrhs + c] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = b[i * nrhs + c];
            for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j * nrhs + c];
            b[i * nrhs + c] = s / A[i * n + i];
        }
    }
}

inline double norm2(int n, const double* v) {
    double s = 0.; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return std::sqrt(s);
}
inline double dot(int n, const double* a, const double* b) {
    double s = 0.; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s;
}

struct PointOut {
    double x[7]; int iters; int flag_entry, flag_exit; double cnorm; int ls_evals;
};

// Jacobian dC/dx by n-direction forward duals  (jacfwd, nonlinear_solver.py:122)
inline void jac_x(int model, int n, const Mat<double>& md, const double* mat, const int* cfg,
                  const double* x, const double* xp, const double* gu, double* J) {
    typedef Dual<double, 7> D7;
    Mat<D7> m = make_mat<D7>(mat, cfg);
    D7 xx[7], xxp[7], g[9], C[7];
    for (int i = 0; i < 7; ++i) { xx[i] = D7(x[i]); xxp[i] = D7(xp[i]); }
    for (int i = 0; i < n; ++i) xx[i].d[i] = 1.;
    for (int i = 0; i < 9; ++i) g[i] = D7(gu[i]);
    residual(model, m, xx, xxp, g, C);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) J[r * n + c] = C[r].d[c];
    (void)md;
}

inline void solve_point(const double* mat, const int* cfg, const double* sol,
                        const double* xprev, const double* xinit, const double* gu, PointOut& out) {
    const int model = cfg[0], mode = cfg[4], max_iters = cfg[5], ls_max = cfg[6];
    const int n = (model == MODEL_SEP) ? 7 : 6;
    const double abs_tol = sol[0], rel_tol = sol[1], c1 = sol[2], bmin = sol[3], bmax = sol[4];
    Mat<double> m = make_mat<double>(mat, cfg);
    double x[7] = {0}, xp[7] = {0}, C[7] = {0};
    for (int i = 0; i < n; ++i) { x[i] = xinit ? xinit[i] : xprev[i]; xp[i] = xprev[i]; }
    out.flag_entry = residual(model, m, x, xp, gu, C);
    out.ls_evals = 0;
    int ii = 0; bool converged = false;
    double n0 = norm2(n, C), nc = n0;
    if (mode == 0) {
        // traced Newton, nonlinear_solver.py:102-155
        while (ii < max_iters && !converged) {
            nc = norm2(n, C);
            double rel = nc / n0;                         // 0/0 -> NaN
            if (rel < rel_tol || nc < abs_tol) { converged = true; continue; }
            double J[49], delta[7];
            jac_x(model, n, m, mat, cfg, x, xp, gu, J);
            for (int i = 0; i < n; ++i) delta[i] = C[i];
            lu_solve(n, J, delta, 1);                    // delta = solve(J, C)
            // line search, line_search.py:125-181 (quadratic model)
            double CC = dot(n, C, C);
            double phi0 = 0.5 * CC, dphi0 = -CC, armijo = c1 * dphi0;
            int ne = 0; double a = 1.; bool acc = false;
            double aux[7], best_a = 1., best_phi = std::numeric_limits<double>::infinity(), best_aux[7];
            for (int i = 0; i < n; ++i) { aux[i] = C[i]; best_aux[i] = C[i]; }
            while (ne < ls_max && !acc) {
                double xt[7] = {0}, Ct[7] = {0};
                for (int i = 0; i < n; ++i) xt[i] = x[i] - a * delta[i];
                residual(model, m, xt, xp, gu, Ct);
                double phi = 0.5 * dot(n, Ct, Ct);
                bool finite = std::isfinite(phi);
                if (finite && phi < best_phi) { best_a = a; best_phi = phi; for (int i = 0; i < n; ++i) best_aux[i] = Ct[i]; }
                acc = finite && (phi <= phi0 + a * armijo);
                double denom = 2.0 * (phi - phi0 - dphi0 * a);
                double am = (denom == 0.0) ? 0.5 * a : -dphi0 * a * a / denom;
                double ac = std::fmin(std::fmax(am, bmin * a), bmax * a);
                if (am != am) ac = am;                    // jnp.clip propagates NaN
                if (!acc) a = finite ? ac : 0.5 * a;
                for (int i = 0; i < n; ++i) aux[i] = Ct[i];
                ++ne;
            }
            out.ls_evals += ne;
            double ar = acc ? a : best_a;
            const double* Cn = acc ? aux : best_aux;
            for (int i = 0; i < n; ++i) { x[i] = x[i] - ar * delta[i]; C[i] = Cn[i]; }
            ++ii;
        }
        nc = norm2(n, C);
    } else {
        // imperative newton_solve, nonlinear_solver.py:14-85 (ls_max = its max_ls_evals, 0 = none)
        while (ii < max_iters && !converged) {
            residual(model, m, x, xp, gu, C);
            nc = norm2(n, C);
            double rel;
            if (ii == 0) { n0 = nc; rel = 1.; } else rel = nc / n0;
            if (rel < rel_tol || nc < abs_tol) { converged = true; break; }
            double J[49], delta[7];
            jac_x(model, n, m, mat, cfg, x, xp, gu, J);
            for (int i = 0; i < n; ++i) delta[i] = -C[i];
            lu_solve(n, J, delta, 1);
            for (int i = 0; i < n; ++i) x[i] += delta[i];
            if (ls_max > 0) {
                // legacy line search, nonlinear_solver.py:55-81 (beta = 1e-4, eta = 0.5)
                double Cj[7] = {0};
                residual(model, m, x, xp, gu, Cj);
                const double psi0 = 0.5 * nc * nc, dpsi0 = -2.0 * psi0;
                int jj = 1;
                double aj = 1.0, cj = norm2(n, Cj), psij = 0.5 * cj * cj;
                while (psij >= (1.0 - 2.0 * 1e-4 * aj) * psi0) {
                    const double ap = aj;
                    aj = std::fmax(0.5 * aj, -(aj * aj * dpsi0) / (2.0 * (psij - psi0 - aj * dpsi0)));
                    if (jj == ls_max) break;
                    ++jj;
                    const double da = aj - ap;
                    for (int i = 0; i < n; ++i) x[i] += da * delta[i];
                    residual(model, m, x, xp, gu, Cj);
                    cj = norm2(n, Cj);
                    psij = 0.5 * cj * cj;
                    ++out.ls_evals;
                }
            }
            ++ii;
        }
    }
    for (int i = 0; i < 7; ++i) out.x[i] = x[i];
    out.iters = ii; out.cnorm = nc;
    double Ctmp[7];
    out.flag_exit = residual(model, m, x, xp, gu, Ctmp);
}

}  // namespace

extern "C" {

// All point arrays are component-major ("SoA"): comp c of point i at [c*ld + i].
// strain: 6 comps (xx,xy,xz,yy,yz,zz tensor components, grad_u := sym strain) or
// 9 comps (grad_u row-major [k*3+j] = du_k/dx_j) per cfg[7].
// Optional outputs may be NULL.
//   xi (n_xi), sigma (6, global cauchy upper triangle), dsig_deps (36: row a col b,
//   derivative wrt the symmetric strain component b with both (k,l),(l,k) entries
//   moving), dxi_deps (n_xi*6), dC_dp (n_xi*n_active, at (xi*, xi_prev)),
//   dC_dxi (n_xi*n_xi), dC_dxi_prev (n_xi*n_xi), iters, flags (bit0 = plastic at
//   entry x0, bit1 = plastic at exit x*), cnorm, ls_evals.
int oracle_mp_update(const double* mat, const int* cfg, const double* sol,
                     const int* active_pid, int n_active,
                     int64_t npts, int64_t ld,
                     const double* xi_prev, const double* strain, const double* xi_init,
                     double* xi, double* sigma, double* dsig_deps, double* dxi_deps,
                     double* dC_dp, double* dC_dxi, double* dC_dxi_prev,
                     int* iters, int* flags, double* cnorm, int* ls_evals,
                     int nthreads, double* dsig_dxi, double* dsig_dp) {
    const int model = cfg[0];
    const int n = (model == MODEL_SEP) ? 7 : 6;
    const int ncomp = cfg[7];
    if (ncomp != 6 && ncomp != 9) return 1;
    if (n_active > 16) return 2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256)
#endif
    for (int64_t i = 0; i < npts; ++i) {
        double xp[7] = {0}, gu[9];
        for (int c = 0; c < n; ++c) xp[c] = xi_prev[c * ld + i];
        if (ncomp == 6) {
            double e[6];
            for (int c = 0; c < 6; ++c) e[c] = strain[c * ld + i];
            gu[0] = e[0]; gu[1] = e[1]; gu[2] = e[2]; gu[3] = e[1]; gu[4] = e[3];
            gu[5] = e[4]; gu[6] = e[2]; gu[7] = e[4]; gu[8] = e[5];
        } else {
            for (int c = 0; c < 9; ++c) gu[c] = strain[c * ld + i];
        }
        PointOut po;
        double x0[7] = {0};
        if (xi_init) for (int c = 0; c < n; ++c) x0[c] = xi_init[c * ld + i];
        solve_point(mat, cfg, sol, xp, xi_init ? x0 : nullptr, gu, po);
        if (xi) for (int c = 0; c < n; ++c) xi[c * ld + i] = po.x[c];
        if (iters) iters[i] = po.iters;
        if (flags) flags[i] = po.flag_entry | (po.flag_exit << 1);
        if (cnorm) cnorm[i] = po.cnorm;
        if (ls_evals) ls_evals[i] = po.ls_evals;
        Mat<double> md = make_mat<double>(mat, cfg);
        static const int up[6] = {0, 1, 2, 4, 5, 8};
        if (sigma) {
            double sg[9]; cauchy(model, md, po.x, gu, sg);
            for (int a = 0; a < 6; ++a) sigma[a * ld + i] = sg[up[a]];
        }
        const bool need_A = dsig_deps || dxi_deps || dC_dxi;
        double A[49];
        if (need_A) {
            jac_x(model, n, md, mat, cfg, po.x, xp, gu, A);
            if (dC_dxi) for (int k = 0; k < n * n; ++k) dC_dxi[k * ld + i] = A[k];
        }
        if (dC_dxi_prev) {
            typedef Dual<double, 7> D7;
            Mat<D7> m = make_mat<D7>(mat, cfg);
            D7 xx[7], xxp[7], g[9], C[7];
            for (int k = 0; k < 7; ++k) { xx[k] = D7(po.x[k]); xxp[k] = D7(xp[k]); }
            for (int k = 0; k < n; ++k) xxp[k].d[k] = 1.;
            for (int k = 0; k < 9; ++k) g[k] = D7(gu[k]);
            residual(model, m, xx, xxp, g, C);
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) dC_dxi_prev[(r * n + c) * ld + i] = C[r].d[c];
        }
        if (dsig_deps || dxi_deps) {
            // IFT (nonlinear_solver.py:158-171): dx/de_b = -A^{-1} dC/de_b, six
            // symmetric strain directions e_b (both tensor entries move).
            typedef Dual<double, 6> D6;
            Mat<D6> m = make_mat<D6>(mat, cfg);
            D6 xx[7], xxp[7], g[9], C[7];
            for (int k = 0; k < 7; ++k) { xx[k] = D6(po.x[k]); xxp[k] = D6(xp[k]); }
            for (int k = 0; k < 9; ++k) g[k] = D6(gu[k]);
            static const int lo[6] = {0, 3, 6, 4, 7, 8};
            for (int b = 0; b < 6; ++b) { g[up[b]].d[b] = 1.; g[lo[b]].d[b] = 1.; }
            residual(model, m, xx, xxp, g, C);
            double rhs[42];
            for (int r = 0; r < n; ++r)
                for (int b = 0; b < 6; ++b) rhs[r * 6 + b] = -C[r].d[b];
            double Ac[49];
            std::memcpy(Ac, A, sizeof(double) * n * n);
            lu_solve(n, Ac, rhs, 6);
            if (dxi_deps)
                for (int r = 0; r < n; ++r)
                    for (int b = 0; b < 6; ++b) dxi_deps[(r * 6 + b) * ld + i] = rhs[r * 6 + b];
            if (dsig_deps) {
                // total derivative of cauchy: seed x with dx/de_b and grad_u with e_b
                D6 sg[9];
                for (int r = 0; r < n; ++r)
                    for (int b = 0; b < 6; ++b) xx[r].d[b] = rhs[r * 6 + b];
                cauchy(model, m, xx, g, sg);
                for (int a = 0; a < 6; ++a)
                    for (int b = 0; b < 6; ++b) dsig_deps[(a * 6 + b) * ld + i] = sg[up[a]].d[b];
            }
        }
        if (dsig_dxi) {        // dcauchy/dxi (model.py:150), (a*n + c)
            typedef Dual<double, 7> D7;
            Mat<D7> m = make_mat<D7>(mat, cfg);
            D7 xx[7], g[9], sg[9];
            for (int k = 0; k < 7; ++k) xx[k] = D7(po.x[k]);
            for (int k = 0; k < n; ++k) xx[k].d[k] = 1.;
            for (int k = 0; k < 9; ++k) g[k] = D7(gu[k]);
            cauchy(model, m, xx, g, sg);
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < n; ++c) dsig_dxi[(a * n + c) * ld + i] = sg[up[a]].d[c];
        }
        if (dsig_dp && n_active > 0) {   // dcauchy/dparams (model.py:152), (a*n_active + c)
            typedef Dual<double, 16> DP;
            Mat<DP> m = make_mat<DP>(mat, cfg);
            for (int c = 0; c < n_active; ++c) mat_slot(m, active_pid[c]).d[c] = 1.;
            DP xx[7], g[9], sg[9];
            for (int k = 0; k < 7; ++k) xx[k] = DP(po.x[k]);
            for (int k = 0; k < 9; ++k) g[k] = DP(gu[k]);
            cauchy(model, m, xx, g, sg);
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < n_active; ++c) dsig_dp[(a * n_active + c) * ld + i] = sg[up[a]].d[c];
        }
        if (dC_dp && n_active > 0) {
            typedef Dual<double, 16> DP;
            Mat<DP> m = make_mat<DP>(mat, cfg);
            for (int c = 0; c < n_active; ++c) mat_slot(m, active_pid[c]).d[c] = 1.;
            DP xx[7], xxp[7], g[9], C[7];
            for (int k = 0; k < 7; ++k) { xx[k] = DP(po.x[k]); xxp[k] = DP(xp[k]); }
            for (int k = 0; k < 9; ++k) g[k] = DP(gu[k]);
            residual(model, m, xx, xxp, g, C);
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n_active; ++c) dC_dp[(r * n_active + c) * ld + i] = C[r].d[c];
        }
    }
    return 0;
}

// lambda, mu and their 2x2 Jacobian wrt the given pair (sorted-key order)
void oracle_lame(int pair, double p, double q, double* out6) {
    typedef Dual<double, 2> D2;
    Mat<D2> m; m.pair = pair; m.el0 = D2(p); m.el1 = D2(q); m.el0.d[0] = 1.; m.el1.d[1] = 1.;
    D2 lam, mu; lame(m, lam, mu);
    out6[0] = lam.v; out6[1] = mu.v; out6[2] = lam.d[0]; out6[3] = lam.d[1];
    out6[4] = mu.d[0]; out6[5] = mu.d[1];
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
