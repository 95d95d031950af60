// K3 / K4 - FE element-block kernels (sm_100a): gather U, interpolate grad_u at
// every integration point, run the local Newton there, and emit the element
// residual R_e, the IFT-consistent element tangent K_e (K3) and the converged
// local state, in the reference's element-major layouts.
//
// Replaces, for one mesh element block (reference file:line):
//   cmad/fem/assembly.py:616-732   assemble_element_block (vmap over elements)
//   cmad/fem/assembly.py:416-535   per_element_R_and_K_coupled (scan over IPs)
//   cmad/fem/assembly.py:538-613   per_element_R_coupled (K4: WANT_K = false)
//   cmad/global_residuals/global_residual.py:361-395   per-IP COUPLED evaluator
//   cmad/global_residuals/small_disp_equilibrium.py:112-118  R = (grad_N @ sigma) w dv
//   cmad/global_residuals/interpolation.py:49-55       grad_u = U_e^T grad_N
//
// Mapping: one thread per integration point.  tet4 (1 IP): the thread owns the
// element, everything stays in registers and K_e rows are streamed out as they
// are produced.  hex8 (8 IPs): the 8 threads of an element solve their points,
// publish (D w dv, grad_N, sigma w dv) in shared memory, then thread a sums the
// three K_e rows of node a over the 8 points in fixed order (bit-reproducible).
// Element-major arrays are moved with 256-bit vector loads/stores
// (LDG.E.256 / STG.E.256): every request is a whole 32-byte sector.
//
// The tangent: with g_kl = d u_k / d x_l = sum_b U[b,k] gN[b,l] and
// D[v(j,i)][v(k,l)] = d sigma_ji / d eps_kl (symmetric strain components),
//   K[(a,i),(b,k)] = w dv sum_{j,l} gN[a,j] D[v(j,i)][v(k,l)] (k==l ? 1 : 1/2) gN[b,l].
#include "fe_block.cuh"
#include "j2_radial.cuh"

namespace cmadx {
namespace {

constexpr int FE_BLOCK = 128;

// symmetric-tensor component of entry (i, j) in the packing xx,xy,xz,yy,yz,zz
CMADX_DEV constexpr int vix(int i, int j) {
    return (i == j) ? (i == 0 ? 0 : (i == 1 ? 3 : 5)) : ((i + j == 1) ? 1 : ((i + j == 2) ? 2 : 4));
}

CMADX_DEV void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
CMADX_DEV void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
                 : "memory");
}

struct PointOut {
    double x[7];
    double sg[6];   // global cauchy
    int iters, flags;
    bool bail;
};

// ---- J2 radial-return point (see j2_radial.cuh) ---------------------------------
template <bool WANT_D>
CMADX_DEV void point_j2(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                        const double (&e)[6], bool live, PointOut& o, double (&D)[6][6]) {
    J2Radial rs;
    j2_radial_solve(m, nw, xp, e, live, rs);
    o.bail = rs.bail;
    o.iters = rs.ii;
    o.flags = rs.flag_entry | ((rs.plastic ? 1 : 0) << 1);
    const double dg = rs.alpha - rs.alpha0;
#pragma unroll
    for (int a = 0; a < 6; ++a) o.x[a] = fma(dg, rs.n0[a], xp[a]);
    o.x[6] = rs.alpha;
    double ee[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ee[a] = e[a] - o.x[a];
    const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
#pragma unroll
    for (int a = 0; a < 6; ++a) o.sg[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
    if (WANT_D) {
        const J2Tangent t = j2_tangent_coeffs(m, rs);
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const double sb = mult(b) * rs.sh[b];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double devE = (a == b) ? 1.0 : 0.0;
                if (is_diag(a) && is_diag(b)) devE -= 1.0 / 3.0;
                const double emx = t.g1 * devE + (t.g2 - t.g1) * rs.sh[a] * sb;
                double v = m.two_mu * (((a == b) ? 1.0 : 0.0) - emx);
                if (is_diag(a) && is_diag(b)) v += m.lam;
                D[a][b] = v;
            }
        }
    }
}

// 6x6 maps between global and material symmetric-tensor components for a
// rotation Q (cmad/models/small_elastic_plastic.py:44-62, 318-319)
CMADX_DEV void rot_maps(const double* Q, double (&T)[6][6], double (&S)[6][6]) {
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

// ---- generic 7x7 Newton point (point_solver.cuh) --------------------------------
template <int YK, bool ROT, bool WANT_D>
CMADX_DEV void point_generic(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                             const double (&e)[6], bool live, PointOut& o, double (&D)[6][6]) {
    double em[6];
    if (ROT) {
        double T[6][6], S[6][6];
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
#pragma unroll
    for (int c = 0; c < 7; ++c) o.x[c] = xp[c];
    SepPoint<YK> pt;
    double Cres[7];
    const NewtonResult nr = local_newton<SepPoint<YK>, 7>(m, nw, pt, o.x, xp, em, live, Cres);
    o.bail = false;
    o.iters = nr.iters;
    o.flags = nr.flag_entry | ((pt.plastic ? 1 : 0) << 1);
    double sig[6];
    {
        double ee[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) ee[a] = em[a] - o.x[a];
        const double ltr = m.lam * (ee[0] + ee[3] + ee[5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) sig[a] = is_diag(a) ? fma(m.two_mu, ee[a], ltr) : m.two_mu * ee[a];
    }
    if (ROT) {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) s = fma(S[a][c], sig[c], s);
            o.sg[a] = s;
        }
    } else {
#pragma unroll
        for (int a = 0; a < 6; ++a) o.sg[a] = sig[a];
    }
    if (!WANT_D) return;

    // IFT (nonlinear_solver.py:158-171): d sigma/d eps = Cel . (A^{-1})[0:6,0:6] in material axes
    const bool pl = pt.plastic;
    const double dg = o.x[6] - xp[6];
    RegLU<7> lu;
    pt.jacobian(m, dg, lu.a);
    bool trouble = false;
    if (__any_sync(__activemask(), pl)) trouble = lu.factor_natural() && pl;
    const bool slow = __any_sync(__activemask(), trouble);
    if (slow && trouble) {
        pt.jacobian(m, dg, lu.a);
        lu.factor_pivot();
    }
    double Dm[6][6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        double X[7];
#pragma unroll
        for (int r = 0; r < 7; ++r) X[r] = (r == b) ? 1.0 : 0.0;
        if (pl) { if (slow && trouble) lu.solve_pivot(X); else lu.solve_natural(X); }
        const double ltr = m.lam * (X[0] + X[3] + X[5]);
#pragma unroll
        for (int a = 0; a < 6; ++a) Dm[a][b] = is_diag(a) ? fma(m.two_mu, X[a], ltr) : m.two_mu * X[a];
    }
    if (!ROT) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) D[a][b] = Dm[a][b];
    } else {
        double T[6][6], S[6][6];
        rot_maps(m.Q, T, S);
        double DT[6][6];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) s = fma(Dm[a][c], T[c][b], s);
                DT[a][b] = s;
            }
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) s = fma(S[a][c], DT[c][b], s);
                D[a][b] = s;
            }
    }
}

// SOLVER: 0 = J2 radial return (may bail), 1 + YK = generic Newton for yield surface YK
template <int SOLVER, bool ROT, bool WANT_D>
CMADX_DEV void solve_point(const DevMat& m, const DevNewton& nw, const double (&xp)[7],
                           const double (&e)[6], bool live, PointOut& o, double (&D)[6][6]) {
    if (SOLVER == 0) point_j2<WANT_D>(m, nw, xp, e, live, o, D);
    else point_generic<(SOLVER > 0 ? SOLVER - 1 : 0), ROT, WANT_D>(m, nw, xp, e, live, o, D);
}

// symmetric strain of grad_u[k][j] = sum_a U[a][k] gN[a][j]
template <int NB>
CMADX_DEV void strain_from_U(const double (&U)[NB][3], const double (&gN)[NB][3], double (&e)[6]) {
    double g[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < NB; ++a) s = fma(U[a][k], gN[a][j], s);
            g[k][j] = s;
        }
    e[0] = g[0][0]; e[3] = g[1][1]; e[5] = g[2][2];
    e[1] = 0.5 * (g[0][1] + g[1][0]); e[2] = 0.5 * (g[0][2] + g[2][0]); e[4] = 0.5 * (g[1][2] + g[2][1]);
}

CMADX_DEV void append_bail(const FeArgs& A, int64_t e) {
    const unsigned slot = atomicAdd(A.bail_count, 1u);
    if (slot < A.bail_cap) A.bail_list[slot] = (int)e;
}

// ==================================================================== tet4, 1 IP
template <int SOLVER, bool ROT, bool WANT_K>
CMADX_DEV void tet4_element(const FeArgs& A, const int64_t e, const bool live) {
    const cmadx_fe_block_t& b = A.b;
    double gN[4][3], U[4][3], xp[7];
    int eq[12];
    double wdv = 0.0;
    if (live) {
        const int4* q = reinterpret_cast<const int4*>(b.elem_eq + e * 12);
        const int4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        eq[0] = q0.x; eq[1] = q0.y; eq[2] = q0.z; eq[3] = q0.w;
        eq[4] = q1.x; eq[5] = q1.y; eq[6] = q1.z; eq[7] = q1.w;
        eq[8] = q2.x; eq[9] = q2.y; eq[10] = q2.z; eq[11] = q2.w;
        const double* g = b.grad_N + e * 12;
        ld256(g, gN[0][0], gN[0][1], gN[0][2], gN[1][0]);
        ld256(g + 4, gN[1][1], gN[1][2], gN[2][0], gN[2][1]);
        ld256(g + 8, gN[2][2], gN[3][0], gN[3][1], gN[3][2]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) U[a][k] = __ldg(b.U + eq[3 * a + k]);
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + e * 7 + c);
        wdv = __ldg(b.quad_w) * __ldg(b.det + e);
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) { U[a][k] = 0.0; gN[a][k] = 0.0; eq[3 * a + k] = 0; }
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = 0.0;
    }
    double eps[6];
    strain_from_U<4>(U, gN, eps);

    PointOut o;
    double D[6][6];
    solve_point<SOLVER, ROT, WANT_K>(A.m, A.nw, xp, eps, live, o, D);
    if (!live) return;
    if (SOLVER == 0 && o.bail) { append_bail(A, e); return; }

#pragma unroll
    for (int c = 0; c < 7; ++c) b.xi[e * 7 + c] = o.x[c];
    if (b.iters) b.iters[e] = o.iters;
    if (b.flags) b.flags[e] = o.flags;
    if (b.sigma) {
#pragma unroll
        for (int a = 0; a < 6; ++a) b.sigma[e * 6 + a] = o.sg[a];
    }
    // R[a][i] = sum_j gN[a][j] sigma[j][i] w dv
    if (b.R_elem || b.R_global) {
        double R[4][3];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], o.sg[vix(j, i)], s);
                R[a][i] = s * wdv;
            }
        if (b.R_elem) {
            double* r = b.R_elem + e * 12;
            st256(r, R[0][0], R[0][1], R[0][2], R[1][0]);
            st256(r + 4, R[1][1], R[1][2], R[2][0], R[2][1]);
            st256(r + 8, R[2][2], R[3][0], R[3][1], R[3][2]);
        }
        if (b.R_global) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int i = 0; i < 3; ++i) atomicAdd(b.R_global + eq[3 * a + i], R[a][i]);
        }
    }
    if constexpr (WANT_K) {
    // Dh = D (k==l ? 1 : 1/2) w dv
#pragma unroll
    for (int al = 0; al < 6; ++al)
#pragma unroll
        for (int be = 0; be < 6; ++be) D[al][be] *= is_diag(be) ? wdv : 0.5 * wdv;
    double* Ke = b.K_elem + e * 144;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double P[6];
#pragma unroll
            for (int be = 0; be < 6; ++be) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fma(gN[a][j], D[vix(j, i)][be], s);
                P[be] = s;
            }
            double row[12];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < 3; ++l) s = fma(P[vix(k, l)], gN[bb][l], s);
                    row[3 * bb + k] = s;
                }
            double* r = Ke + (3 * a + i) * 12;
            st256(r, row[0], row[1], row[2], row[3]);
            st256(r + 4, row[4], row[5], row[6], row[7]);
            st256(r + 8, row[8], row[9], row[10], row[11]);
        }
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
__global__ void __launch_bounds__(FE_BLOCK) fe_tet4_kernel(const __grid_constant__ FeArgs A) {
    if (!LIST) {
        const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        tet4_element<SOLVER, ROT, WANT_K>(A, e, e < A.b.n_elems);
    } else {
        // list mode: a small grid walks the elements the J2 kernel handed back
        const unsigned cnt = *A.bail_count;
        if (cnt == 0u) return;
        const bool all = cnt > A.bail_cap;
        const int64_t total = all ? A.b.n_elems : (int64_t)cnt;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        const int lane = threadIdx.x & 31;
        for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < total;
             base += stride) {
            const int64_t j = base + lane;
            const bool live = j < total;
            const int64_t e = live ? (all ? j : (int64_t)A.bail_list[j]) : 0;
            tet4_element<SOLVER, ROT, WANT_K>(A, e, live);
        }
    }
}

// =================================================================== hex8, 8 IPs
// shared-memory record of one integration point
constexpr int HEX_REC = 66;                       // 36 Dh + 24 gN + 6 sigma*w*dv
constexpr int HEX_ESTRIDE = 8 * HEX_REC + 2;      // +2 doubles: the 4 elements of a warp hit distinct banks
constexpr int HEX_EPB = FE_BLOCK / 8;             // elements per block
constexpr int HEX_SMEM_DOUBLES = HEX_EPB * HEX_ESTRIDE + HEX_EPB * 24;

template <int SOLVER, bool ROT, bool WANT_K>
CMADX_DEV void hex8_point(const FeArgs& A, const int64_t e, const bool live, double* smem) {
    const cmadx_fe_block_t& b = A.b;
    const int lane = threadIdx.x & 31;
    const int ip = lane & 7;                      // also the node this thread owns in phase C
    const int eloc = threadIdx.x >> 3;
    double* rec_e = smem + eloc * HEX_ESTRIDE;
    double* Ue = smem + HEX_EPB * HEX_ESTRIDE + eloc * 24;

    // ---- phase A: gather U_e (thread t fetches node t), load this point's geometry/state
    int eq3[3] = {0, 0, 0};
    double gN[8][3], xp[7];
    double wdv = 0.0;
    if (live) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            eq3[k] = __ldg(b.elem_eq + e * 24 + 3 * ip + k);
            Ue[3 * ip + k] = __ldg(b.U + eq3[k]);
        }
        const double* g = b.grad_N + (e * 8 + ip) * 24;
        double* gf = &gN[0][0];
#pragma unroll
        for (int q = 0; q < 6; ++q) ld256(g + 4 * q, gf[4 * q], gf[4 * q + 1], gf[4 * q + 2], gf[4 * q + 3]);
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = __ldg(b.xi_prev + (e * 8 + ip) * 7 + c);
        wdv = __ldg(b.quad_w + ip) * __ldg(b.det + e * 8 + ip);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) Ue[3 * ip + k] = 0.0;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) gN[a][k] = 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) xp[c] = 0.0;
    }
    __syncwarp();
    double eps[6];
    {
        double U[8][3];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) U[a][k] = Ue[3 * a + k];
        strain_from_U<8>(U, gN, eps);
    }

    // ---- phase B: local Newton at this point
    PointOut o;
    double D[6][6];
    solve_point<SOLVER, ROT, WANT_K>(A.m, A.nw, xp, eps, live, o, D);

    // an element is handed to the generic kernel as a whole
    bool ebail = false;
    if (SOLVER == 0) {
        const unsigned bal = __ballot_sync(0xffffffffu, o.bail);
        ebail = ((bal >> (lane & ~7)) & 0xffu) != 0u;
        if (ebail && live && ip == 0) append_bail(A, e);
    }
    const bool emit = live && !ebail;
    if (emit) {
        const int64_t p = e * 8 + ip;
#pragma unroll
        for (int c = 0; c < 7; ++c) b.xi[p * 7 + c] = o.x[c];
        if (b.iters) b.iters[p] = o.iters;
        if (b.flags) b.flags[p] = o.flags;
        if (b.sigma) {
#pragma unroll
            for (int a = 0; a < 6; ++a) b.sigma[p * 6 + a] = o.sg[a];
        }
    }
    // publish this point's record
    {
        double* rec = rec_e + ip * HEX_REC;
        if (WANT_K) {
#pragma unroll
            for (int al = 0; al < 6; ++al)
#pragma unroll
                for (int be = 0; be < 6; ++be) rec[al * 6 + be] = D[al][be] * (is_diag(be) ? wdv : 0.5 * wdv);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) rec[36 + 3 * a + k] = gN[a][k];
#pragma unroll
        for (int a = 0; a < 6; ++a) rec[60 + a] = o.sg[a] * wdv;
    }
    __syncwarp();
    if (!emit) return;

    // ---- phase C: thread a sums rows 3a..3a+2 of R_e / K_e over the 8 points, in order
    const int a = ip;
    double Racc[3] = {0.0, 0.0, 0.0};
    if (b.R_elem || b.R_global) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double* rec = rec_e + q * HEX_REC;
            const double ga0 = rec[36 + 3 * a], ga1 = rec[36 + 3 * a + 1], ga2 = rec[36 + 3 * a + 2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                Racc[i] += fma(ga2, rec[60 + vix(2, i)], fma(ga1, rec[60 + vix(1, i)], ga0 * rec[60 + vix(0, i)]));
        }
        if (b.R_elem) {
#pragma unroll
            for (int i = 0; i < 3; ++i) b.R_elem[e * 24 + 3 * a + i] = Racc[i];
        }
        if (b.R_global) {
#pragma unroll
            for (int i = 0; i < 3; ++i) atomicAdd(b.R_global + eq3[i], Racc[i]);
        }
    }
    if constexpr (WANT_K) {
    double acc[3][24];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 24; ++c) acc[i][c] = 0.0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const double* rec = rec_e + q * HEX_REC;
        const double ga[3] = {rec[36 + 3 * a], rec[36 + 3 * a + 1], rec[36 + 3 * a + 2]};
        double P[3][6];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int be = 0; be < 6; ++be)
                P[i][be] = fma(ga[2], rec[vix(2, i) * 6 + be],
                               fma(ga[1], rec[vix(1, i) * 6 + be], ga[0] * rec[vix(0, i) * 6 + be]));
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const double g0 = rec[36 + 3 * bb], g1 = rec[36 + 3 * bb + 1], g2 = rec[36 + 3 * bb + 2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    acc[i][3 * bb + k] += fma(P[i][vix(k, 2)], g2, fma(P[i][vix(k, 1)], g1, P[i][vix(k, 0)] * g0));
        }
    }
    double* Ke = b.K_elem + e * 576 + (3 * a) * 24;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 6; ++c)
            st256(Ke + i * 24 + 4 * c, acc[i][4 * c], acc[i][4 * c + 1], acc[i][4 * c + 2], acc[i][4 * c + 3]);
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
__global__ void __launch_bounds__(FE_BLOCK) fe_hex8_kernel(const __grid_constant__ FeArgs A) {
    extern __shared__ __align__(16) double smem[];
    if (!LIST) {
        const int64_t e = (int64_t)blockIdx.x * HEX_EPB + (threadIdx.x >> 3);
        hex8_point<SOLVER, ROT, WANT_K>(A, e, e < A.b.n_elems, smem);
    } else {
        const unsigned cnt = *A.bail_count;
        if (cnt == 0u) return;
        const bool all = cnt > A.bail_cap;
        const int64_t total = all ? A.b.n_elems : (int64_t)cnt;
        const int64_t stride = (int64_t)gridDim.x * HEX_EPB;
        for (int64_t base = (int64_t)blockIdx.x * HEX_EPB; base < total; base += stride) {
            const int64_t j = base + (threadIdx.x >> 3);
            const bool live = j < total;
            const int64_t e = live ? (all ? j : (int64_t)A.bail_list[j]) : 0;
            hex8_point<SOLVER, ROT, WANT_K>(A, e, live, smem);
            __syncwarp();
        }
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
cudaError_t launch_one(const FeArgs& A, cudaStream_t stream, int sms) {
    const int64_t n = A.b.n_elems;
    if (A.b.n_basis == 4) {
        const int64_t nblk = LIST ? 2 * sms : (n + FE_BLOCK - 1) / FE_BLOCK;
        fe_tet4_kernel<SOLVER, ROT, WANT_K, LIST><<<(unsigned)nblk, FE_BLOCK, 0, stream>>>(A);
    } else {
        auto kern = fe_hex8_kernel<SOLVER, ROT, WANT_K, LIST>;
        const size_t smem = sizeof(double) * HEX_SMEM_DOUBLES;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t nblk = LIST ? 2 * sms : (n + HEX_EPB - 1) / HEX_EPB;
        kern<<<(unsigned)nblk, FE_BLOCK, smem, stream>>>(A);
    }
    return cudaGetLastError();
}

template <int SOLVER, bool LIST>
cudaError_t launch_solver(const FeArgs& A, cudaStream_t stream, int sms) {
    const bool k = A.b.K_elem != nullptr;
    if (SOLVER != 0 && A.m.rot) {
        return k ? launch_one<(SOLVER ? SOLVER : 1), true, true, LIST>(A, stream, sms)
                 : launch_one<(SOLVER ? SOLVER : 1), true, false, LIST>(A, stream, sms);
    }
    return k ? launch_one<SOLVER, false, true, LIST>(A, stream, sms)
             : launch_one<SOLVER, false, false, LIST>(A, stream, sms);
}

}  // namespace

// main launch: J2 radial kernel where it applies, else the generic kernel
cudaError_t launch_fe_block(const FeArgs& A, bool j2_radial, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    if (j2_radial) return launch_solver<0, false>(A, stream, 0);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_solver<1, false>(A, stream, 0);
    case CMADX_YIELD_HILL: return launch_solver<2, false>(A, stream, 0);
    case CMADX_YIELD_HOSFORD: return launch_solver<3, false>(A, stream, 0);
    }
    return cudaErrorInvalidValue;
}

// generic J2 kernel over the elements the radial kernel handed back
cudaError_t launch_fe_block_list(const FeArgs& A, cudaStream_t stream) {
    if (A.b.n_elems == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return launch_solver<1, true>(A, stream, sms);
}

}  // namespace cmadx
