// K2-H - second-order (Hessian) pass of the material-point calibration objective: the
// direct-adjoint recurrence of MPDirectAdjointObjective (cmad/objectives/mp_objective.py:
// 218-343, arXiv:2501.04584) over whole stored load histories.
//
// With the adjoint states phi_t of the reverse pass (stored by the K2 adjoint kernel), the
// forward sensitivities X_t = dxi_t/dp (the K2 direct recurrence) and
//   L_t(z) = J_t(xi_t, p) + phi_t . C(xi_t, xi_{t-1}, p),   z = (p, xi_t, xi_{t-1}),
// the reference's thirteen einsum terms (:317-336) are exactly
//   H += Z^T (d2 L_t / dz2) Z,   Z = [I; X_t; X_{t-1}]:
// H_ij = D2 L_t [Z_i, Z_j], a MIXED second directional derivative.  The reference builds the
// tensors d2C/d(.)d(.) by jax.hessian / jacrev(jacfwd) (cmad/models/model.py:134-148,
// cmad/qois/qoi.py:47-58); here every (i <= j) entry is ONE evaluation of L_t in hyper-dual
// arithmetic (value, d/ds, d/du, d2/ds du) along (Z_i, Z_j): templated forward-mode dual
// numbers instead of traced AD - no third-derivative formulas of the yield surfaces are
// written down, the yield normal n = d phi/d sigma is evaluated from its closed form in
// hyper-dual arithmetic.  The branch (plastic / elastic) is the one of the stored state, as
// jnp.where differentiates (cmad/models/paths.py:26-27).
//
// One thread per material point walks its history forward; X_t / X_{t-1} and the pair
// accumulators live in local memory (this is the small-batch calibration path, not a bench
// line).  The pair sums are reduced in fixed order (bit-reproducible).  FULL_3D, PLANE_STRESS and
// UNIAXIAL_STRESS with identity material axes; d/d(hosford a) and d/d(rotation) are not
// provided (as in K2).
#include "mp_outputs.cuh"
#include "mp_sens.cuh"
#include "sep_point_dt.cuh"
#include "rate_point_dt.cuh"

namespace cmadx {
cudaError_t launch_reduce_partials(const double* partials, int64_t nblk, int ncols, double* result,
                                   cudaStream_t stream);

namespace {

constexpr int HESS_BLOCK = 64;
constexpr int HESS_MAX_PAIRS = CMADX_MAX_ACTIVE * (CMADX_MAX_ACTIVE + 1) / 2;

// hyper-dual number: f(z + s v + u w) = v + a s + b u + ab s u  (s^2 = u^2 = 0)
struct HD {
    double v, a, b, ab;
};
CMADX_DEV HD hd(double c) { return {c, 0.0, 0.0, 0.0}; }
CMADX_DEV HD operator+(const HD& x, const HD& y) { return {x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; }
CMADX_DEV HD operator-(const HD& x, const HD& y) { return {x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; }
CMADX_DEV HD operator-(const HD& x) { return {-x.v, -x.a, -x.b, -x.ab}; }
CMADX_DEV HD operator*(const HD& x, const HD& y) {
    return {x.v * y.v, fma(x.a, y.v, x.v * y.a), fma(x.b, y.v, x.v * y.b),
            fma(x.ab, y.v, fma(x.a, y.b, fma(x.b, y.a, x.v * y.ab)))};
}
CMADX_DEV HD operator*(double c, const HD& x) { return {c * x.v, c * x.a, c * x.b, c * x.ab}; }
CMADX_DEV HD operator+(const HD& x, double c) { return {x.v + c, x.a, x.b, x.ab}; }
CMADX_DEV HD operator-(const HD& x, double c) { return {x.v - c, x.a, x.b, x.ab}; }
CMADX_DEV HD operator-(double c, const HD& x) { return {c - x.v, -x.a, -x.b, -x.ab}; }
// g(u): value g, first and second derivative g1, g2 at u.v
CMADX_DEV HD chain(const HD& u, double g, double g1, double g2) {
    return {g, g1 * u.a, g1 * u.b, fma(g1, u.ab, g2 * u.a * u.b)};
}
CMADX_DEV HD inv(const HD& y) {
    const double r = 1.0 / y.v;
    return chain(y, r, -r * r, 2.0 * r * r * r);
}
CMADX_DEV HD operator/(const HD& x, const HD& y) { return x * inv(y); }
CMADX_DEV HD hsqrt(const HD& x) {
    const double s = sqrt(x.v);
    return chain(x, s, 0.5 / s, -0.25 / (s * x.v));
}
CMADX_DEV HD hexp(const HD& x) {
    const double e = exp(x.v);
    return chain(x, e, e, e);
}
// u^r for u.v > 0
CMADX_DEV HD hpow(const HD& u, double r) {
    const double p2 = pow(u.v, r - 2.0);
    return chain(u, p2 * u.v * u.v, r * p2 * u.v, r * (r - 1.0) * p2);
}
CMADX_DEV HD habs(const HD& x) { return (x.v < 0.0) ? -x : x; }

struct HessParams {
    HD lam, mu, Y, S, D, K, hill[6];
};

// phi(sigma) and the yield normal n_a = d phi/d sigma_a (single tensor entry) in hyper-dual
// arithmetic (cmad/models/effective_stress.py:30-52, 168-177; the normal is what
// jax.grad gives, small_elastic_plastic.py:90)
template <int YK>
CMADX_DEV void yield_hd(const DevMat& m, const HessParams& P, const HD (&sig)[6], HD& phi, HD (&n)[6]) {
    if constexpr (YK == CMADX_YIELD_J2) {
        const HD h = (1.0 / 3.0) * (sig[0] + sig[3] + sig[5]);
        HD s[6], ss = hd(0.0);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            s[a] = is_diag(a) ? sig[a] - h : sig[a];
            ss = ss + mult(a) * (s[a] * s[a]);
        }
        const double r32 = 1.2247448713915890491;
        const HD sn = hsqrt(ss);
        phi = r32 * sn;
        const HD isn = inv(sn);
#pragma unroll
        for (int a = 0; a < 6; ++a) n[a] = r32 * (s[a] * isn);
    } else if constexpr (YK == CMADX_YIELD_HILL) {
        const HD d12 = sig[3] - sig[5], d20 = sig[5] - sig[0], d01 = sig[0] - sig[3];
        const HD& F = P.hill[0]; const HD& G = P.hill[1]; const HD& H = P.hill[2];
        const HD& L = P.hill[3]; const HD& Mm = P.hill[4]; const HD& N = P.hill[5];
        const HD q = F * d12 * d12 + G * d20 * d20 + H * d01 * d01
                     + 2.0 * (L * sig[4] * sig[4]) + 2.0 * (Mm * sig[2] * sig[2]) + 2.0 * (N * sig[1] * sig[1]);
        phi = hsqrt(q);
        const HD ip = inv(phi);
        n[0] = (H * d01 - G * d20) * ip;
        n[3] = (F * d12 - H * d01) * ip;
        n[5] = (G * d20 - F * d12) * ip;
        n[1] = N * sig[1] * ip;
        n[2] = Mm * sig[2] * ip;
        n[4] = L * sig[4] * ip;
    } else {
        // Hosford on the diagonal entries: phi = (1/2 sum |Delta_i|^a)^(1/a); the reference's
        // von Mises scaling is a positive homogeneity factor - applied here as a CONSTANT c0
        // (phi(sigma) = c0 phi(sigma / c0) exactly), which keeps large exponents in range
        const double a = m.a;
        const double hv = (sig[0].v + sig[3].v + sig[5].v) / 3.0;
        double ssv = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double s = is_diag(k) ? sig[k].v - hv : sig[k].v;
            ssv = fma(mult(k) * s, s, ssv);
        }
        const double c0 = 1.2247448713915890491 * sqrt(ssv), ic0 = 1.0 / c0;
        const HD dl[3] = {sig[0] - sig[3], sig[3] - sig[5], sig[5] - sig[0]};
        HD sq = hd(0.0);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (dl[i].v != 0.0) sq = sq + hpow(ic0 * habs(dl[i]), a);
        sq = 0.5 * sq;
        phi = c0 * hpow(sq, 1.0 / a);
        const HD ip = inv(phi);
        HD g[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            // d phi/d Delta_i = 1/2 sign(Delta_i) (|Delta_i|/phi)^(a-1)
            if (dl[i].v != 0.0) {
                const HD r = hpow(habs(dl[i]) * ip, a - 1.0);
                g[i] = (dl[i].v > 0.0 ? 0.5 : -0.5) * r;
            } else {
                g[i] = hd(0.0);
            }
        }
        n[0] = g[0] - g[2]; n[3] = g[1] - g[0]; n[5] = g[2] - g[1];
        n[1] = hd(0.0); n[2] = hd(0.0); n[4] = hd(0.0);
    }
}

CMADX_DEV void stress_hd(const HD& lam, const HD& mu, const HD (&x)[7], const double (&em)[6], HD (&sig)[6]) {
    HD ee[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = em[a] - x[a];
    const HD ltr = lam * (ee[0] + ee[3] + ee[5]);
    const HD two_mu = 2.0 * mu;
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? two_mu * ee[a] + ltr : two_mu * ee[a];
}

// Calibration QoI (cmad/qois/calibration.py:56-66)
CMADX_DEV HD qoi_hd(const HD (&sig)[6], const double (&w)[9], const double (&d)[9]) {
    const int comp[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    HD J = hd(0.0);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const HD mis = w[k] * (sig[comp[k]] - d[k]);
        J = J + 0.5 * (mis * mis);
    }
    return J;
}

// rotated material axes: the QoI compares the GLOBAL cauchy S sigma_m with the data
// (S = d(Q s Q^T)/ds on packed components, rot_maps_hess below); S == nullptr: identity axes
CMADX_DEV void to_global_hd(const double* S, HD (&sig)[6]) {
    if (!S) return;
    HD g[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        HD acc = hd(0.0);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = acc + S[a * 6 + c] * sig[c];
        g[a] = acc;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = g[a];
}

// The (parameter, state) cross terms of the QoI's second derivative along the two directions:
// D2 J [(p_i, 0), (0, X_j)] + D2 J [(0, X_i), (p_j, 0)].  The reference omits exactly these
// (its QoI takes jacrev(jacfwd(., DXI_PREV), DPARAMS), cmad/qois/qoi.py:53-55, which is zero for
// a QoI that does not depend on xi_prev); the reference-compatible mode subtracts them.
__device__ __noinline__ double qoi_cross_terms(const HD& lam, const HD& mu, const HD (&x)[7],
                                               const double (&em)[6], const double (&w)[9],
                                               const double (&d)[9], const double* S = nullptr) {
    double acc = 0.0;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        // side 0: parameters keep slot a, the state keeps slot b; side 1: the other way round
        const HD l = side ? HD{lam.v, 0.0, lam.b, 0.0} : HD{lam.v, lam.a, 0.0, 0.0};
        const HD u = side ? HD{mu.v, 0.0, mu.b, 0.0} : HD{mu.v, mu.a, 0.0, 0.0};
        HD xs[7], sig[6];
#pragma unroll
        for (int r = 0; r < 7; ++r) xs[r] = side ? HD{x[r].v, x[r].a, 0.0, 0.0} : HD{x[r].v, 0.0, x[r].b, 0.0};
        stress_hd(l, u, xs, em, sig);
        to_global_hd(S, sig);
        acc += qoi_hd(sig, w, d).ab;
    }
    return acc;
}

// d2/ds du of L_t = J_t + phi . C along the two directions carried by the hyper-duals
template <int YK>
__device__ __forceinline__ double lagrangian_mixed(const DevMat& m, const HessParams& P, const HD (&x)[7],
                                                const HD (&xp)[7], const double (&em)[6],
                                                const double (&phi)[7], const double (&w)[9],
                                                const double (&d)[9], bool plastic, const double* S = nullptr) {
    HD sig[6];
    stress_hd(P.lam, P.mu, x, em, sig);
    const HD two_mu = 2.0 * P.mu;
    HD L;
    {
        HD sg[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sg[a] = sig[a];
        to_global_hd(S, sg);
        L = qoi_hd(sg, w, d);
    }
    if (plastic) {
        HD pe, n[6];
        yield_hd<YK>(m, P, sig, pe, n);
        HD hard = P.Y;
        if (m.hmask & CMADX_HARD_VOCE) hard = hard + P.S * (1.0 - hexp(-(P.D * x[6])));
        if (m.hmask & CMADX_HARD_LINEAR) hard = hard + P.K * x[6];
        const HD f = (pe - hard) * inv(two_mu);
        const HD dg = x[6] - xp[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) L = L + phi[a] * (x[a] - xp[a] - dg * n[a]);
        L = L + phi[6] * f;
    } else {
#pragma unroll
        for (int a = 0; a < 7; ++a) L = L + phi[a] * (x[a] - xp[a]);
    }
    return L.ab;
}

// seeds of parameter `pid` for the pass along (direction slot a: parameter pi, slot b: pj)
CMADX_DEV HD seed(double value, int pid, int pi, int pj) {
    return {value, pid == pi ? 1.0 : 0.0, pid == pj ? 1.0 : 0.0, 0.0};
}

template <int YK>
__global__ void __launch_bounds__(HESS_BLOCK) mp_hess_kernel(const __grid_constant__ SensArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int N = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;
    const int npairs = na * (na + 1) / 2;

    // X_t, X_{t-1} and the pair sums live in SHARED memory, [slot][thread] (conflict-free): as
    // per-thread local arrays (4.3 KB of stack) they overflowed L1 and L2 and the pass was bound
    // by local-memory round trips to DRAM (ncu: long-scoreboard 7.4 per issue, FP64 pipe 18 %)
    extern __shared__ double hs[];
    constexpr int NX = 7;
    auto X = [&](int c, int r) -> double& { return hs[(c * NX + r) * HESS_BLOCK + threadIdx.x]; };
    auto Xp = [&](int c, int r) -> double& { return hs[((na + c) * NX + r) * HESS_BLOCK + threadIdx.x]; };
    auto Hacc = [&](int q) -> double& { return hs[(2 * na * NX + q) * HESS_BLOCK + threadIdx.x]; };
    for (int q = 0; q < npairs; ++q) Hacc(q) = 0.0;
    for (int c = 0; c < na; ++c)
#pragma unroll
        for (int r = 0; r < 7; ++r) { X(c, r) = 0.0; Xp(c, r) = 0.0; }
    // rotated material axes (cmad/models/small_elastic_plastic.py:44-62, 318-319): the strain goes to
    // material axes, the state lives there, the QoI reads the global cauchy.  T[c][b] =
    // d(Q^T e Q)_c / d e_b, S[a][c] = d(Q s Q^T)_a / d s_c on packed components.
    double Trot[36], Srot[36];
    const double* Sq = nullptr;
    if (m.rot) {
        const int ci_[6] = {0, 0, 0, 1, 1, 2}, cj_[6] = {0, 1, 2, 1, 2, 2};
        for (int c = 0; c < 6; ++c)
            for (int b = 0; b < 6; ++b) {
                const int ii = ci_[c], jj = cj_[c], kk = ci_[b], ll = cj_[b];
                double tt = m.Q[3 * kk + ii] * m.Q[3 * ll + jj];
                double ss = m.Q[3 * ii + kk] * m.Q[3 * jj + ll];
                if (kk != ll) { tt += m.Q[3 * ll + ii] * m.Q[3 * kk + jj]; ss += m.Q[3 * ii + ll] * m.Q[3 * jj + kk]; }
                Trot[c * 6 + b] = tt;
                Srot[c * 6 + b] = ss;
            }
        Sq = Srot;
    }
    double x[7], xp[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const double v = live ? __ldg(A.h.xi_hist + c * ld + i) : 0.0;
        x[c] = v; xp[c] = v;
    }
    for (int t = 1; t <= N; ++t) {
        double em[6], d[9], phi[7];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)t * 7 * ld + i;
            const double* ph = A.phi_hist + (int64_t)t * 7 * ld + i;
#pragma unroll
            for (int c = 0; c < 7; ++c) { x[c] = __ldg(xs + c * ld); phi[c] = ph[c * ld]; }
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
#pragma unroll
            for (int c = 0; c < 7; ++c) phi[c] = 0.0;
        }
        if (m.rot) {
            double eg[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) eg[c] = em[c];
            for (int c = 0; c < 6; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < 6; ++b) sacc = fma(Trot[c * 6 + b], eg[b], sacc);
                em[c] = sacc;
            }
        }
        // ---- forward sensitivities X_t = A^{-1}(-dC/dp - B X_{t-1})  (mp_objective.py:300-301)
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
        double Mee[6], nee = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double sacc = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) sacc = fma(pt.yf.M(a, b), ee[b], sacc);
            Mee[a] = sacc;
            nee = fma(mult(a) * pt.n[a], ee[a], nee);
        }
        RegLU<7> lu;
        pt.jacobian(m, dg, lu.a);
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) {
            pt.jacobian(m, dg, lu.a);
            lu.factor_pivot();
        }
        for (int c = 0; c < na; ++c) {
            double col[7], rhs[7];
            dC_dp_column(m, A.pid[c], pl, pt.yf, pt.n, pt.f, pt.eD, x[6], dg, Mee, nee, sig, col);
            const double x6 = Xp(c, 6);
#pragma unroll
            for (int q = 0; q < 6; ++q) rhs[q] = -col[q] + Xp(c, q) - (pl ? pt.n[q] * x6 : 0.0);
            rhs[6] = -col[6] + (pl ? 0.0 : x6);
            if (slow && trouble) lu.solve_pivot(rhs); else lu.solve_natural(rhs);
#pragma unroll
            for (int q = 0; q < 7; ++q) X(c, q) = rhs[q];
        }
        // ---- H_ij += D2 L_t [Z_i, Z_j], one hyper-dual evaluation per pair
        int q = 0;
#pragma unroll 1
        for (int ci = 0; ci < na; ++ci) {
#pragma unroll 1
            for (int cj = ci; cj < na; ++cj, ++q) {
                const int pi = A.pid[ci], pj = A.pid[cj];
                HessParams P;
                {
                    const int ki = pi - CMADX_P_EL0, kj = pj - CMADX_P_EL0;
                    const bool ei = (ki == 0 || ki == 1), ej = (kj == 0 || kj == 1);
                    const int k2 = ki + kj;     // (0,0) -> 0, (0,1)/(1,0) -> 1, (1,1) -> 2
                    P.lam = {m.lam, ei ? m.dlam[ki] : 0.0, ej ? m.dlam[kj] : 0.0, (ei && ej) ? m.d2lam[k2] : 0.0};
                    P.mu = {m.mu, ei ? m.dmu[ki] : 0.0, ej ? m.dmu[kj] : 0.0, (ei && ej) ? m.d2mu[k2] : 0.0};
                }
                P.Y = seed(m.Y, CMADX_P_Y, pi, pj);
                P.S = seed(m.S, CMADX_P_VOCE_S, pi, pj);
                P.D = seed(m.D, CMADX_P_VOCE_D, pi, pj);
                P.K = seed(m.K, CMADX_P_LIN_K, pi, pj);
#pragma unroll
                for (int k = 0; k < 6; ++k) P.hill[k] = seed(m.hill[k], CMADX_P_HILL_F + k, pi, pj);
                HD xh[7], xph[7];
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    xh[r] = {x[r], X(ci, r), X(cj, r), 0.0};
                    xph[r] = {xp[r], Xp(ci, r), Xp(cj, r), 0.0};
                }
                double hij = lagrangian_mixed<YK>(m, P, xh, xph, em, phi, A.h.weight, d, pl, Sq);
                if (A.hess_flags & CMADX_HESS_F_REFERENCE_QOI_CROSS)
                    hij -= qoi_cross_terms(P.lam, P.mu, xh, em, A.h.weight, d, Sq);
                Hacc(q) += hij;
            }
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = x[c];
        for (int c = 0; c < na; ++c)
#pragma unroll
            for (int r = 0; r < 7; ++r) Xp(c, r) = X(c, r);
    }
    // ---- block reduction of the pair sums (fixed order)
    __shared__ double sm[HESS_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < npairs; ++q) {
        double v = live ? Hacc(q) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int wq = 0; wq < HESS_BLOCK / 32; ++wq) s += sm[wq];
            A.partials[(int64_t)blockIdx.x * npairs + q] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// SmallRateElasticPlastic, FULL_3D (cmad/models/small_rate_elastic_plastic.py:250-346): state
// x = [cauchy(6), alpha] in material axes, de = T (eps_t - eps_{t-1}); the QoI reads the state's own
// stress (S x for rotated axes), so J_t has no parameter dependence and the reference's Hessian
// (qoi.py:53-55) is complete for this model.  Same scheme as mp_hess_kernel: X_t from the K2 direct
// recurrence of mp_sens_rate.cu, one hyper-dual evaluation of J_t + phi_t . C per pair.
template <int YK>
__device__ __forceinline__ double lagrangian_mixed_rate(const DevMat& m, const HessParams& P, const HD (&x)[7],
                                                     const HD (&xp)[7], const double (&de)[6],
                                                     const double (&phi)[7], const double (&w)[9],
                                                     const double (&d)[9], bool plastic, const double* S) {
    HD sig[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = x[a];
    HD L;
    {
        HD sg[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sg[a] = sig[a];
        to_global_hd(S, sg);
        L = qoi_hd(sg, w, d);
    }
    const HD two_mu = 2.0 * P.mu;
    const HD i2mu = inv(two_mu);
    const HD dg = x[6] - xp[6];
    HD wv[6];                        // the strain the elastic operator acts on: de (- dgamma n)
#pragma unroll
    for (int a = 0; a < 6; ++a) wv[a] = hd(de[a]);
    if (plastic) {
        HD pe, n[6];
        yield_hd<YK>(m, P, sig, pe, n);
        HD hard = P.Y;
        if (m.hmask & CMADX_HARD_VOCE) hard = hard + P.S * (1.0 - hexp(-(P.D * x[6])));
        if (m.hmask & CMADX_HARD_LINEAR) hard = hard + P.K * x[6];
        L = L + phi[6] * ((pe - hard) * i2mu);
#pragma unroll
        for (int a = 0; a < 6; ++a) wv[a] = wv[a] - dg * n[a];
    } else {
        L = L + phi[6] * dg;
    }
    const HD ltr = P.lam * (wv[0] + wv[3] + wv[5]);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const HD inc = is_diag(a) ? two_mu * wv[a] + ltr : two_mu * wv[a];
        L = L + phi[a] * ((x[a] - xp[a] - inc) * i2mu);
    }
    return L.ab;
}

template <int YK>
__global__ void __launch_bounds__(HESS_BLOCK) mp_hess_rate_kernel(const __grid_constant__ SensArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int N = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;
    const int npairs = na * (na + 1) / 2;
    extern __shared__ double hs[];
    constexpr int NX = 7;
    auto X = [&](int c, int r) -> double& { return hs[(c * NX + r) * HESS_BLOCK + threadIdx.x]; };
    auto Xp = [&](int c, int r) -> double& { return hs[((na + c) * NX + r) * HESS_BLOCK + threadIdx.x]; };
    auto Hacc = [&](int q) -> double& { return hs[(2 * na * NX + q) * HESS_BLOCK + threadIdx.x]; };
    for (int q = 0; q < npairs; ++q) Hacc(q) = 0.0;
    for (int c = 0; c < na; ++c)
#pragma unroll
        for (int r = 0; r < 7; ++r) { X(c, r) = 0.0; Xp(c, r) = 0.0; }
    double Trot[36], Srot[36];
    const double* Sq = nullptr;
    if (m.rot) {
        const int ci_[6] = {0, 0, 0, 1, 1, 2}, cj_[6] = {0, 1, 2, 1, 2, 2};
        for (int c = 0; c < 6; ++c)
            for (int b = 0; b < 6; ++b) {
                const int ii = ci_[c], jj = cj_[c], kk = ci_[b], ll = cj_[b];
                double tt = m.Q[3 * kk + ii] * m.Q[3 * ll + jj];
                double ss = m.Q[3 * ii + kk] * m.Q[3 * jj + ll];
                if (kk != ll) { tt += m.Q[3 * ll + ii] * m.Q[3 * kk + jj]; ss += m.Q[3 * ii + ll] * m.Q[3 * jj + kk]; }
                Trot[c * 6 + b] = tt;
                Srot[c * 6 + b] = ss;
            }
        Sq = Srot;
    }
    double x[7], xp[7], ep[6];
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const double v = live ? __ldg(A.h.xi_hist + c * ld + i) : 0.0;
        x[c] = v; xp[c] = v;
    }
    if (live) rate_load_strain(A.h.strain, sc, ld, i, ep);
    else {
#pragma unroll
        for (int c = 0; c < 6; ++c) ep[c] = 0.0;
    }
    for (int t = 1; t <= N; ++t) {
        double de[6], et[6], d[9], phi[7];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)t * 7 * ld + i;
            const double* ph = A.phi_hist + (int64_t)t * 7 * ld + i;
#pragma unroll
            for (int c = 0; c < 7; ++c) { x[c] = __ldg(xs + c * ld); phi[c] = ph[c * ld]; }
            rate_load_strain(A.h.strain + (int64_t)t * sc * ld, sc, ld, i, et);
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < 7; ++c) { x[c] = (c == 0) ? 1.0 : 0.0; phi[c] = 0.0; }
#pragma unroll
            for (int c = 0; c < 6; ++c) et[c] = ep[c] + 1e-3 * (c == 0);
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) { de[c] = et[c] - ep[c]; ep[c] = et[c]; }
        if (m.rot) {
            double dgl[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) dgl[c] = de[c];
            for (int c = 0; c < 6; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < 6; ++b) sacc = fma(Trot[c * 6 + b], dgl[b], sacc);
                de[c] = sacc;
            }
        }
        // ---- forward sensitivities X_t = A^{-1}(-dC/dp - B X_{t-1}),  B = [-I/2mu, -n; 0, -1]
        RatePoint<YK> pt;
        double C[7];
        pt.residual(m, x, xp, de, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        RegLU<7> lu;
        pt.jacobian(m, dg, lu.a);
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) {
            pt.jacobian(m, dg, lu.a);
            lu.factor_pivot();
        }
        for (int c = 0; c < na; ++c) {
            double col[7], rhs[7];
            rate_dC_dp_column<YK>(m, A.pid[c], pt, x, xp, de, col);
            const double x6 = Xp(c, 6);
#pragma unroll
            for (int q = 0; q < 6; ++q) rhs[q] = -col[q] + m.inv_two_mu * Xp(c, q) + (pl ? pt.n[q] * x6 : 0.0);
            rhs[6] = -col[6] + (pl ? 0.0 : x6);
            if (slow && trouble) lu.solve_pivot(rhs); else lu.solve_natural(rhs);
#pragma unroll
            for (int q = 0; q < 7; ++q) X(c, q) = rhs[q];
        }
        // ---- H_ij += D2 L_t [Z_i, Z_j]
        int q = 0;
#pragma unroll 1
        for (int ci = 0; ci < na; ++ci) {
#pragma unroll 1
            for (int cj = ci; cj < na; ++cj, ++q) {
                const int pi = A.pid[ci], pj = A.pid[cj];
                HessParams P;
                {
                    const int ki = pi - CMADX_P_EL0, kj = pj - CMADX_P_EL0;
                    const bool ei = (ki == 0 || ki == 1), ej = (kj == 0 || kj == 1);
                    const int k2 = ki + kj;
                    P.lam = {m.lam, ei ? m.dlam[ki] : 0.0, ej ? m.dlam[kj] : 0.0, (ei && ej) ? m.d2lam[k2] : 0.0};
                    P.mu = {m.mu, ei ? m.dmu[ki] : 0.0, ej ? m.dmu[kj] : 0.0, (ei && ej) ? m.d2mu[k2] : 0.0};
                }
                P.Y = seed(m.Y, CMADX_P_Y, pi, pj);
                P.S = seed(m.S, CMADX_P_VOCE_S, pi, pj);
                P.D = seed(m.D, CMADX_P_VOCE_D, pi, pj);
                P.K = seed(m.K, CMADX_P_LIN_K, pi, pj);
#pragma unroll
                for (int k = 0; k < 6; ++k) P.hill[k] = seed(m.hill[k], CMADX_P_HILL_F + k, pi, pj);
                HD xh[7], xph[7];
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    xh[r] = {x[r], X(ci, r), X(cj, r), 0.0};
                    xph[r] = {xp[r], Xp(ci, r), Xp(cj, r), 0.0};
                }
                Hacc(q) += lagrangian_mixed_rate<YK>(m, P, xh, xph, de, phi, A.h.weight, d, pl, Sq);
            }
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = x[c];
        for (int c = 0; c < na; ++c)
#pragma unroll
            for (int r = 0; r < 7; ++r) Xp(c, r) = X(c, r);
    }
    __shared__ double sm[HESS_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < npairs; ++q) {
        double v = live ? Hacc(q) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int wq = 0; wq < HESS_BLOCK / 32; ++wq) s += sm[wq];
            A.partials[(int64_t)blockIdx.x * npairs + q] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// PLANE_STRESS / UNIAXIAL_STRESS (n_xi = 8 / 9): the reference's Hessian known answer (KA5,
// tests/objectives/test_jvp_vs_original.py:75-97) lives in plane stress.  Same scheme; the
// bordered residual (sep_point_dt.cuh: stretch unknowns, stress-constraint rows cauchy_cc/2mu)
// is evaluated in hyper-dual arithmetic, the linear algebra of X_t reuses SepPointDT.
template <int DT, int N>
CMADX_DEV void stress_hd_dt(const HD& lam, const HD& mu, const HD (&x)[N], const double (&em)[6],
                            HD (&ee)[6], HD (&sig)[6]) {
    HD et[6];
    if (DT == CMADX_DEF_PLANE_STRESS) {
        et[0] = hd(em[0]); et[1] = hd(em[1]); et[2] = hd(0.0); et[3] = hd(em[3]); et[4] = hd(0.0); et[5] = x[7] - 1.0;
    } else {
        et[0] = hd(em[0]); et[1] = x[1]; et[2] = x[2]; et[3] = x[7] - 1.0; et[4] = x[4]; et[5] = x[8] - 1.0;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = et[a] - x[a];
    const HD ltr = lam * (ee[0] + ee[3] + ee[5]);
    const HD two_mu = 2.0 * mu;
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? two_mu * ee[a] + ltr : two_mu * ee[a];
}

// rotated material axes (SepPointDTRot): the kinematic constraints live in GLOBAL axes
// (uniaxial stress: the off-diagonal global strains are those of S ep), the update in material axes;
// sig = material-frame stress, sg = S sig = what the QoI and the constraint rows read
template <int DT, int N>
__device__ __noinline__ void stress_hd_dt_rot(const HD& lam, const HD& mu, const HD (&x)[N], const double (&em)[6],
                                const double (&T)[6][6], const double (&S)[6][6], HD (&sig)[6], HD (&sg)[6]) {
    HD eg[6];
    if (DT == CMADX_DEF_PLANE_STRESS) {
        eg[0] = hd(em[0]); eg[1] = hd(em[1]); eg[2] = hd(0.0); eg[3] = hd(em[3]); eg[4] = hd(0.0); eg[5] = x[7] - 1.0;
    } else {
        HD og[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            HD acc = hd(0.0);
#pragma unroll
            for (int c = 0; c < 6; ++c) acc = acc + S[a][c] * x[c];
            og[a] = acc;
        }
        eg[0] = hd(em[0]); eg[1] = og[1]; eg[2] = og[2]; eg[3] = x[7] - 1.0; eg[4] = og[4]; eg[5] = x[N > 8 ? 8 : 7] - 1.0;
    }
    HD ee[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        HD acc = hd(0.0);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = acc + T[a][c] * eg[c];
        ee[a] = acc - x[a];
    }
    const HD ltr = lam * (ee[0] + ee[3] + ee[5]);
    const HD two_mu = 2.0 * mu;
#pragma unroll
    for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? two_mu * ee[a] + ltr : two_mu * ee[a];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        HD acc = hd(0.0);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = acc + S[a][c] * sig[c];
        sg[a] = acc;
    }
}

template <int DT, int N>
__device__ __noinline__ double qoi_cross_terms_dt_rot(const HD& lam, const HD& mu, const HD (&x)[N],
                                                      const double (&em)[6], const double (&w)[9],
                                                      const double (&d)[9], const double (&T)[6][6],
                                                      const double (&S)[6][6]) {
    double acc = 0.0;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const HD l = side ? HD{lam.v, 0.0, lam.b, 0.0} : HD{lam.v, lam.a, 0.0, 0.0};
        const HD u = side ? HD{mu.v, 0.0, mu.b, 0.0} : HD{mu.v, mu.a, 0.0, 0.0};
        HD xs[N], sig[6], sg[6];
#pragma unroll
        for (int r = 0; r < N; ++r) xs[r] = side ? HD{x[r].v, x[r].a, 0.0, 0.0} : HD{x[r].v, 0.0, x[r].b, 0.0};
        stress_hd_dt_rot<DT, N>(l, u, xs, em, T, S, sig, sg);
        acc += qoi_hd(sg, w, d).ab;
    }
    return acc;
}

template <int DT, int N>
__device__ __noinline__ double qoi_cross_terms_dt(const HD& lam, const HD& mu, const HD (&x)[N],
                                                  const double (&em)[6], const double (&w)[9],
                                                  const double (&d)[9]) {
    double acc = 0.0;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const HD l = side ? HD{lam.v, 0.0, lam.b, 0.0} : HD{lam.v, lam.a, 0.0, 0.0};
        const HD u = side ? HD{mu.v, 0.0, mu.b, 0.0} : HD{mu.v, mu.a, 0.0, 0.0};
        HD xs[N], ee[6], sig[6];
#pragma unroll
        for (int r = 0; r < N; ++r) xs[r] = side ? HD{x[r].v, x[r].a, 0.0, 0.0} : HD{x[r].v, 0.0, x[r].b, 0.0};
        stress_hd_dt<DT, N>(l, u, xs, em, ee, sig);
        acc += qoi_hd(sig, w, d).ab;
    }
    return acc;
}

template <int YK, int DT, int N, bool ROT = false>
__device__ __forceinline__ double lagrangian_mixed_dt(const DevMat& m, const HessParams& P, const HD (&x)[N],
                                                   const HD (&xp)[N], const double (&em)[6],
                                                   const double (&phi)[N], const double (&w)[9],
                                                   const double (&d)[9], bool plastic,
                                                   const double (*T)[6] = nullptr, const double (*S)[6] = nullptr) {
    constexpr int NZ = N - 7;
    HD sig[6], sg[6];
    if constexpr (ROT) {
        stress_hd_dt_rot<DT, N>(P.lam, P.mu, x, em, *reinterpret_cast<const double (*)[6][6]>(T),
                                *reinterpret_cast<const double (*)[6][6]>(S), sig, sg);
    } else {
        HD ee[6];
        stress_hd_dt<DT, N>(P.lam, P.mu, x, em, ee, sig);
#pragma unroll
        for (int a = 0; a < 6; ++a) sg[a] = sig[a];
    }
    const HD two_mu = 2.0 * P.mu;
    const HD i2m = inv(two_mu);
    HD L = qoi_hd(sg, w, d);
    if (plastic) {
        HD pe, n[6];
        yield_hd<YK>(m, P, sig, pe, n);
        HD hard = P.Y;
        if (m.hmask & CMADX_HARD_VOCE) hard = hard + P.S * (1.0 - hexp(-(P.D * x[6])));
        if (m.hmask & CMADX_HARD_LINEAR) hard = hard + P.K * x[6];
        const HD f = (pe - hard) * i2m;
        const HD dg = x[6] - xp[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) L = L + phi[a] * (x[a] - xp[a] - dg * n[a]);
        L = L + phi[6] * f;
    } else {
#pragma unroll
        for (int a = 0; a < 7; ++a) L = L + phi[a] * (x[a] - xp[a]);
    }
    // stress-constraint rows (both branches): cauchy_cc / 2mu for the stretch-driven components
#pragma unroll
    for (int k = 0; k < NZ; ++k) L = L + phi[7 + k] * (sg[SepPointDT<YK, DT>::zcomp(k)] * i2m);
    return L.ab;
}

template <int YK, int DT, bool ROT = false>
__global__ void __launch_bounds__(HESS_BLOCK) mp_hess_dt_kernel(const __grid_constant__ SensArgs A) {
    using Pt = typename std::conditional<ROT, SepPointDTRot<YK, DT>, SepPointDT<YK, DT>>::type;
    constexpr int N = Pt::N, NZ = Pt::NZ;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int NT = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;
    const int npairs = na * (na + 1) / 2;

    extern __shared__ double hs[];          // X_t, X_{t-1}, pair sums: [slot][thread], see mp_hess_kernel
    auto X = [&](int c, int r) -> double& { return hs[(c * N + r) * HESS_BLOCK + threadIdx.x]; };
    auto Xp = [&](int c, int r) -> double& { return hs[((na + c) * N + r) * HESS_BLOCK + threadIdx.x]; };
    auto Hacc = [&](int q) -> double& { return hs[(2 * na * N + q) * HESS_BLOCK + threadIdx.x]; };
    for (int q = 0; q < npairs; ++q) Hacc(q) = 0.0;
    for (int c = 0; c < na; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) { X(c, r) = 0.0; Xp(c, r) = 0.0; }
    double x[N], xp[N];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const double v = live ? __ldg(A.h.xi_hist + c * ld + i) : ((c < 7) ? 0.0 : 1.0);
        x[c] = v; xp[c] = v;
    }
    for (int t = 1; t <= NT; ++t) {
        double em[6] = {1e-3, 0.0, 0.0, 0.0, 0.0, 0.0}, d[9], phi[N];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)t * N * ld + i;
            const double* ph = A.phi_hist + (int64_t)t * N * ld + i;
#pragma unroll
            for (int c = 0; c < N; ++c) { x[c] = __ldg(xs + c * ld); phi[c] = ph[c * ld]; }
            load_dt_strain<DT>(A.h.strain + (int64_t)t * sc * ld, sc, ld, i, em);
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) phi[c] = 0.0;
        }
        // ---- forward sensitivities X_t = A^{-1}(-dC/dp - B X_{t-1}) as in mp_sens_dt.cu
        Pt pt;
        double C[N];
        pt.residual(m, x, xp, em, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        double et[6], ee[6], sig[6];
        pt.material_strain(x, em, et);
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = et[a] - x[a];
        const double tree = ee[0] + ee[3] + ee[5];
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], m.lam * tree) : m.two_mu * ee[a];
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
        pt.jacobian(m, dg, lu.a);
        const bool trouble = lu.factor_natural();
        const bool slow = __any_sync(__activemask(), trouble);
        if (slow && trouble) { pt.jacobian(m, dg, lu.a); lu.factor_pivot(); }
        for (int c = 0; c < na; ++c) {
            const int pid = A.pid[c];
            double c7[7], rhs[N];
            dC_dp_column(m, pid, pl, pt.b.yf, pt.b.n, pt.b.f, pt.b.eD, x[6], dg, Mee, nee, sig, c7);
            double dr = 0.0;
            if (pid == CMADX_P_EL0 || pid == CMADX_P_EL1) {
                const int k = pid - CMADX_P_EL0;
                dr = (m.dlam[k] * m.two_mu - m.lam * 2.0 * m.dmu[k]) * m.inv_two_mu * m.inv_two_mu * tree;
            }
            const double x6 = Xp(c, 6);
#pragma unroll
            for (int q = 0; q < 6; ++q) rhs[q] = -c7[q] + Xp(c, q) - (pl ? pt.b.n[q] * x6 : 0.0);
            rhs[6] = -c7[6] + (pl ? 0.0 : x6);
#pragma unroll
            for (int k = 0; k < NZ; ++k) rhs[7 + k] = -dr;
            if (slow && trouble) lu.solve_pivot(rhs); else lu.solve_natural(rhs);
#pragma unroll
            for (int q = 0; q < N; ++q) X(c, q) = rhs[q];
        }
        int q = 0;
#pragma unroll 1
        for (int ci = 0; ci < na; ++ci) {
#pragma unroll 1
            for (int cj = ci; cj < na; ++cj, ++q) {
                const int pi = A.pid[ci], pj = A.pid[cj];
                HessParams P;
                {
                    const int ki = pi - CMADX_P_EL0, kj = pj - CMADX_P_EL0;
                    const bool ei = (ki == 0 || ki == 1), ej = (kj == 0 || kj == 1);
                    const int k2 = ki + kj;
                    P.lam = {m.lam, ei ? m.dlam[ki] : 0.0, ej ? m.dlam[kj] : 0.0, (ei && ej) ? m.d2lam[k2] : 0.0};
                    P.mu = {m.mu, ei ? m.dmu[ki] : 0.0, ej ? m.dmu[kj] : 0.0, (ei && ej) ? m.d2mu[k2] : 0.0};
                }
                P.Y = seed(m.Y, CMADX_P_Y, pi, pj);
                P.S = seed(m.S, CMADX_P_VOCE_S, pi, pj);
                P.D = seed(m.D, CMADX_P_VOCE_D, pi, pj);
                P.K = seed(m.K, CMADX_P_LIN_K, pi, pj);
#pragma unroll
                for (int k = 0; k < 6; ++k) P.hill[k] = seed(m.hill[k], CMADX_P_HILL_F + k, pi, pj);
                HD xh[N], xph[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    xh[r] = {x[r], X(ci, r), X(cj, r), 0.0};
                    xph[r] = {xp[r], Xp(ci, r), Xp(cj, r), 0.0};
                }
                double hij;
                if constexpr (ROT) {
                    hij = lagrangian_mixed_dt<YK, DT, N, true>(m, P, xh, xph, em, phi, A.h.weight, d, pl, pt.T, pt.S);
                    if (A.hess_flags & CMADX_HESS_F_REFERENCE_QOI_CROSS)
                        hij -= qoi_cross_terms_dt_rot<DT, N>(P.lam, P.mu, xh, em, A.h.weight, d, pt.T, pt.S);
                } else {
                    hij = lagrangian_mixed_dt<YK, DT, N>(m, P, xh, xph, em, phi, A.h.weight, d, pl);
                    if (A.hess_flags & CMADX_HESS_F_REFERENCE_QOI_CROSS)
                        hij -= qoi_cross_terms_dt<DT, N>(P.lam, P.mu, xh, em, A.h.weight, d);
                }
                Hacc(q) += hij;
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) xp[c] = x[c];
        for (int c = 0; c < na; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) Xp(c, r) = X(c, r);
    }
    __shared__ double sm[HESS_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < npairs; ++q) {
        double v = live ? Hacc(q) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int wq = 0; wq < HESS_BLOCK / 32; ++wq) s += sm[wq];
            A.partials[(int64_t)blockIdx.x * npairs + q] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// SmallRateElasticPlastic under PLANE_STRESS / UNIAXIAL_STRESS (n_xi 8 / 12; rate_point_dt.cuh,
// small_rate_elastic_plastic.py:189-199, 312-343): the stretch unknowns enter the global strain
// increment as z - z_prev, the off-axis delta strains directly; the constraint rows act on the GLOBAL
// stress increment S Cel (dem - dgamma n) / 2mu.  Identity and rotated axes share the code path
// (T = S = I).  X_t as in mp_sens_rate_dt.cu's direct recurrence.
template <int YK, int DT, int N>
__device__ __forceinline__ double lagrangian_mixed_rate_dt(const DevMat& m, const HessParams& P, const HD (&x)[N],
                                                        const HD (&xp)[N], const double (&em)[6],
                                                        const double (&phi)[N], const double (&w)[9],
                                                        const double (&d)[9], bool plastic,
                                                        const double (&T)[6][6], const double (&S)[6][6]) {
    using Pt = RatePointDT<YK, DT>;
    HD deg[6];
    if (DT == CMADX_DEF_PLANE_STRESS) {
        deg[0] = hd(em[0]); deg[1] = hd(em[1]); deg[2] = hd(0.0); deg[3] = hd(em[3]); deg[4] = hd(0.0);
        deg[5] = x[7] - xp[7];
    } else {
        deg[0] = hd(em[0]); deg[1] = x[N > 9 ? 9 : 0]; deg[2] = x[N > 10 ? 10 : 0]; deg[3] = x[7] - xp[7];
        deg[4] = x[N > 11 ? 11 : 0]; deg[5] = x[N > 8 ? 8 : 0] - xp[N > 8 ? 8 : 0];
    }
    HD wv[6], sig[6], sg[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        HD acc = hd(0.0), gl = hd(0.0);
#pragma unroll
        for (int c = 0; c < 6; ++c) { acc = acc + T[a][c] * deg[c]; gl = gl + S[a][c] * x[c]; }
        wv[a] = acc;
        sig[a] = x[a];
        sg[a] = gl;
    }
    HD L = qoi_hd(sg, w, d);
    const HD two_mu = 2.0 * P.mu;
    const HD i2mu = inv(two_mu);
    const HD dg = x[6] - xp[6];
    if (plastic) {
        HD pe, n[6];
        yield_hd<YK>(m, P, sig, pe, n);
        HD hard = P.Y;
        if (m.hmask & CMADX_HARD_VOCE) hard = hard + P.S * (1.0 - hexp(-(P.D * x[6])));
        if (m.hmask & CMADX_HARD_LINEAR) hard = hard + P.K * x[6];
        L = L + phi[6] * ((pe - hard) * i2mu);
#pragma unroll
        for (int a = 0; a < 6; ++a) wv[a] = wv[a] - dg * n[a];
    } else {
        L = L + phi[6] * dg;
    }
    const HD ltr = P.lam * (wv[0] + wv[3] + wv[5]);
    HD inc[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        inc[a] = is_diag(a) ? two_mu * wv[a] + ltr : two_mu * wv[a];
        L = L + phi[a] * ((x[a] - xp[a] - inc[a]) * i2mu);
    }
#pragma unroll
    for (int r = 0; r < Pt::NR; ++r) {
        const int c = Pt::ccomp(r);
        HD gi = hd(0.0);
#pragma unroll
        for (int a = 0; a < 6; ++a) gi = gi + S[c][a] * inc[a];
        L = L + phi[7 + r] * (gi * i2mu);
    }
    return L.ab;
}

template <int YK, int DT>
__global__ void __launch_bounds__(HESS_BLOCK) mp_hess_rate_dt_kernel(const __grid_constant__ SensArgs A) {
    using Pt = RatePointDT<YK, DT>;
    constexpr int N = Pt::N, NZ = Pt::NZ;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < A.h.n;
    const int64_t ld = A.h.ld;
    const DevMat& m = A.m;
    const int NT = A.h.nsteps, na = A.n_active, sc = A.h.strain_comps;
    const int npairs = na * (na + 1) / 2;
    extern __shared__ double hs[];
    auto X = [&](int c, int r) -> double& { return hs[(c * N + r) * HESS_BLOCK + threadIdx.x]; };
    auto Xp = [&](int c, int r) -> double& { return hs[((na + c) * N + r) * HESS_BLOCK + threadIdx.x]; };
    auto Hacc = [&](int q) -> double& { return hs[(2 * na * N + q) * HESS_BLOCK + threadIdx.x]; };
    for (int q = 0; q < npairs; ++q) Hacc(q) = 0.0;
    for (int c = 0; c < na; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) { X(c, r) = 0.0; Xp(c, r) = 0.0; }
    double x[N], xp[N], ep[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const double v = live ? __ldg(A.h.xi_hist + c * ld + i) : ((c >= 7 && c < 7 + NZ) ? 1.0 : 0.0);
        x[c] = v; xp[c] = v;
    }
    if (live) load_dt_strain<DT>(A.h.strain, sc, ld, i, ep);
    for (int t = 1; t <= NT; ++t) {
        double em[6] = {1e-3, 0.0, 0.0, 0.0, 0.0, 0.0}, d[9], phi[N];
        if (live) {
            const double* xs = A.h.xi_hist + (int64_t)t * N * ld + i;
            const double* ph = A.phi_hist + (int64_t)t * N * ld + i;
#pragma unroll
            for (int c = 0; c < N; ++c) { x[c] = __ldg(xs + c * ld); phi[c] = ph[c * ld]; }
            double et[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            load_dt_strain<DT>(A.h.strain + (int64_t)t * sc * ld, sc, ld, i, et);
#pragma unroll
            for (int c = 0; c < 6; ++c) { em[c] = et[c] - ep[c]; ep[c] = et[c]; }
            const double* ds = A.h.data + (int64_t)t * 9 * ld + i;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = __ldg(ds + c * ld);
        } else {
            x[0] = 1.0;
#pragma unroll
            for (int c = 0; c < 9; ++c) d[c] = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) phi[c] = 0.0;
        }
        Pt pt;
        double C[N];
        pt.residual(m, x, xp, em, C);
        const bool pl = pt.plastic;
        const double dg = x[6] - xp[6];
        {
            double Am[N][N];
            pt.jacobian(m, dg, Am);
            RegLU<N> lu;
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int b = 0; b < N; ++b) lu.a[a][b] = Am[a][b];
            const bool trouble = lu.factor_natural();
            const bool slow = __any_sync(__activemask(), trouble);
            if (slow && trouble) {
#pragma unroll
                for (int a = 0; a < N; ++a)
#pragma unroll
                    for (int b = 0; b < N; ++b) lu.a[a][b] = Am[a][b];
                lu.factor_pivot();
            }
            for (int c = 0; c < na; ++c) {
                double col[N], rhs[N], xpv[N];
                rate_dt_dC_dp_column<YK, DT>(m, A.pid[c], pt, x, xp, col);
#pragma unroll
                for (int k = 0; k < N; ++k) xpv[k] = Xp(c, k);
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    double v = -col[q];
#pragma unroll
                    for (int k = 0; k < N; ++k) v = fma(-rate_dt_B(m, pt, Am, q, k), xpv[k], v);
                    rhs[q] = v;
                }
                if (slow && trouble) lu.solve_pivot(rhs); else lu.solve_natural(rhs);
#pragma unroll
                for (int q = 0; q < N; ++q) X(c, q) = rhs[q];
            }
        }
        int q = 0;
#pragma unroll 1
        for (int ci = 0; ci < na; ++ci) {
#pragma unroll 1
            for (int cj = ci; cj < na; ++cj, ++q) {
                const int pi = A.pid[ci], pj = A.pid[cj];
                HessParams P;
                {
                    const int ki = pi - CMADX_P_EL0, kj = pj - CMADX_P_EL0;
                    const bool ei = (ki == 0 || ki == 1), ej = (kj == 0 || kj == 1);
                    const int k2 = ki + kj;
                    P.lam = {m.lam, ei ? m.dlam[ki] : 0.0, ej ? m.dlam[kj] : 0.0, (ei && ej) ? m.d2lam[k2] : 0.0};
                    P.mu = {m.mu, ei ? m.dmu[ki] : 0.0, ej ? m.dmu[kj] : 0.0, (ei && ej) ? m.d2mu[k2] : 0.0};
                }
                P.Y = seed(m.Y, CMADX_P_Y, pi, pj);
                P.S = seed(m.S, CMADX_P_VOCE_S, pi, pj);
                P.D = seed(m.D, CMADX_P_VOCE_D, pi, pj);
                P.K = seed(m.K, CMADX_P_LIN_K, pi, pj);
#pragma unroll
                for (int k = 0; k < 6; ++k) P.hill[k] = seed(m.hill[k], CMADX_P_HILL_F + k, pi, pj);
                HD xh[N], xph[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    xh[r] = {x[r], X(ci, r), X(cj, r), 0.0};
                    xph[r] = {xp[r], Xp(ci, r), Xp(cj, r), 0.0};
                }
                Hacc(q) += lagrangian_mixed_rate_dt<YK, DT, N>(m, P, xh, xph, em, phi, A.h.weight, d, pl, pt.T, pt.S);
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) xp[c] = x[c];
        for (int c = 0; c < na; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) Xp(c, r) = X(c, r);
    }
    __shared__ double sm[HESS_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < npairs; ++q) {
        double v = live ? Hacc(q) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int wq = 0; wq < HESS_BLOCK / 32; ++wq) s += sm[wq];
            A.partials[(int64_t)blockIdx.x * npairs + q] = s;
        }
        __syncthreads();
    }
}

// dynamic shared memory of a block: (2 na n_xi + npairs) doubles per thread
template <class K>
cudaError_t launch_hess_kernel(K kernel, const SensArgs& A, int n_xi, unsigned nblk, cudaStream_t stream) {
    const int na = A.n_active;
    const size_t smem = sizeof(double) * HESS_BLOCK * (size_t)(2 * na * n_xi + na * (na + 1) / 2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<nblk, HESS_BLOCK, smem, stream>>>(A);
    return cudaGetLastError();
}

template <int DT>
cudaError_t launch_hess_dt(const SensArgs& A, unsigned nblk, cudaStream_t stream) {
    constexpr int NX = (DT == CMADX_DEF_PLANE_STRESS) ? 8 : 9;
    if (A.m.model == CMADX_MODEL_SMALL_RATE_ELASTIC_PLASTIC) {
        constexpr int NR = (DT == CMADX_DEF_PLANE_STRESS) ? 8 : 12;
        switch (A.m.yield) {
        case CMADX_YIELD_J2: return launch_hess_kernel(mp_hess_rate_dt_kernel<CMADX_YIELD_J2, DT>, A, NR, nblk, stream);
        case CMADX_YIELD_HILL: return launch_hess_kernel(mp_hess_rate_dt_kernel<CMADX_YIELD_HILL, DT>, A, NR, nblk, stream);
        case CMADX_YIELD_HOSFORD: return launch_hess_kernel(mp_hess_rate_dt_kernel<CMADX_YIELD_HOSFORD, DT>, A, NR, nblk, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    if (A.m.rot) {
        switch (A.m.yield) {
        case CMADX_YIELD_J2: return launch_hess_kernel(mp_hess_dt_kernel<CMADX_YIELD_J2, DT, true>, A, NX, nblk, stream);
        case CMADX_YIELD_HILL: return launch_hess_kernel(mp_hess_dt_kernel<CMADX_YIELD_HILL, DT, true>, A, NX, nblk, stream);
        case CMADX_YIELD_HOSFORD: return launch_hess_kernel(mp_hess_dt_kernel<CMADX_YIELD_HOSFORD, DT, true>, A, NX, nblk, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_hess_kernel(mp_hess_dt_kernel<CMADX_YIELD_J2, DT>, A, NX, nblk, stream);
    case CMADX_YIELD_HILL: return launch_hess_kernel(mp_hess_dt_kernel<CMADX_YIELD_HILL, DT>, A, NX, nblk, stream);
    case CMADX_YIELD_HOSFORD: return launch_hess_kernel(mp_hess_dt_kernel<CMADX_YIELD_HOSFORD, DT>, A, NX, nblk, stream);
    default: return cudaErrorInvalidValue;
    }
}

// upper-triangle pair sums -> full symmetric na x na matrix
__global__ void expand_pairs_kernel(const double* pairs, int na, double* H) {
    const int ci = threadIdx.x / na, cj = threadIdx.x % na;
    if (ci >= na) return;
    const int lo = ci < cj ? ci : cj, hi = ci < cj ? cj : ci;
    const int q = lo * na - lo * (lo - 1) / 2 + (hi - lo);
    H[ci * na + cj] = pairs[q];
}

}  // namespace

int64_t hess_blocks(int64_t n) { return (n + HESS_BLOCK - 1) / HESS_BLOCK; }

// partials: [nblk][npairs], pair_sums: [npairs] scratch, H_out: [na*na]
cudaError_t launch_mp_hess(const SensArgs& A, int def_type, double* pair_sums, double* H_out, cudaStream_t stream) {
    const int na = A.n_active, npairs = na * (na + 1) / 2;
    if (na == 0) return cudaSuccess;
    if (A.h.n == 0) return cudaMemsetAsync(H_out, 0, sizeof(double) * na * na, stream);
    const int64_t nblk = hess_blocks(A.h.n);
    cudaError_t e;
    if (def_type == CMADX_DEF_PLANE_STRESS) {
        e = launch_hess_dt<CMADX_DEF_PLANE_STRESS>(A, (unsigned)nblk, stream);
    } else if (def_type == CMADX_DEF_UNIAXIAL_STRESS) {
        e = launch_hess_dt<CMADX_DEF_UNIAXIAL_STRESS>(A, (unsigned)nblk, stream);
    } else if (A.m.model == CMADX_MODEL_SMALL_RATE_ELASTIC_PLASTIC) {
        switch (A.m.yield) {
        case CMADX_YIELD_J2: e = launch_hess_kernel(mp_hess_rate_kernel<CMADX_YIELD_J2>, A, 7, (unsigned)nblk, stream); break;
        case CMADX_YIELD_HILL: e = launch_hess_kernel(mp_hess_rate_kernel<CMADX_YIELD_HILL>, A, 7, (unsigned)nblk, stream); break;
        case CMADX_YIELD_HOSFORD: e = launch_hess_kernel(mp_hess_rate_kernel<CMADX_YIELD_HOSFORD>, A, 7, (unsigned)nblk, stream); break;
        default: return cudaErrorInvalidValue;
        }
    } else {
        switch (A.m.yield) {
        case CMADX_YIELD_J2: e = launch_hess_kernel(mp_hess_kernel<CMADX_YIELD_J2>, A, 7, (unsigned)nblk, stream); break;
        case CMADX_YIELD_HILL: e = launch_hess_kernel(mp_hess_kernel<CMADX_YIELD_HILL>, A, 7, (unsigned)nblk, stream); break;
        case CMADX_YIELD_HOSFORD: e = launch_hess_kernel(mp_hess_kernel<CMADX_YIELD_HOSFORD>, A, 7, (unsigned)nblk, stream); break;
        default: return cudaErrorInvalidValue;
        }
    }
    if (e != cudaSuccess) return e;
    e = launch_reduce_partials(A.partials, nblk, npairs, pair_sums, stream);
    if (e != cudaSuccess) return e;
    expand_pairs_kernel<<<1, na * na, 0, stream>>>(pair_sums, na, H_out);
    return cudaGetLastError();
}

}  // namespace cmadx
