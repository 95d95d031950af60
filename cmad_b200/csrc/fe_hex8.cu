// K3 / K4 for hex8 elements with the degree-2 rule (8 integration points): the 8
// threads of an element cooperate through shared memory; see fe_block.cu for the
// contract and the references.
//
// Data movement.  Everything element-major in HBM is moved by the 8 threads of the
// element as contiguous runs (grad_N: 48 x 32 B, xi_prev / xi: 56 doubles, K_e:
// 144 x 32 B), lane t taking chunk q*8 + t, so one warp instruction touches 8
// 128-byte lines instead of 32 (ncu on the first version: 0.65 L1 wavefronts per
// 32-byte sector moved; the L1 data pipe, not HBM or FP64, was the limiter).
// The transposition "thread owns a point" <-> "thread owns a chunk" goes through
// the element's shared-memory region.
//
// Shared memory of one element (HEX_REGION doubles), reused in time:
//   [0,24)    U_e
//   [24,456)  8 integration-point records of HEX_REC = 54 doubles
//             [0,21)  Dh = D[al][be] h(be) w dv, upper triangle; h = 1 for the diagonal
//                     strain components, 1/2 for the off-diagonal ones (D is the
//                     derivative w.r.t. a symmetric component, both tensor entries
//                     moving), which makes Dh symmetric: 21 entries do
//             [22,46) grad_N (8 nodes x 3), loaded straight from HBM into place
//             [46,52) sigma w dv
//   [456,512) xi_prev, then xi, of the 8 points (7 doubles each)
//   [512,536) dU_e (K6 JVP with a displacement direction only)
//   phase C end: the 24x24 K_e tile (8 node blocks x 74 doubles) over the whole region
//
// Bank layout (ncu: the first version spent half its shared-memory wavefronts on
// conflicts): 128-bit accesses are served per quarter-warp (= one element), 64-bit
// ones per half-warp (= two elements).  HEX_REC*8 B = 27 x 16 B, odd: the 8 records of
// an element start in distinct 16-byte bank groups.  The two elements of a half-warp
// are 600 doubles = 8 mod 16 apart (16 banks): their node-strided 64-bit grad_N reads
// fall in complementary banks.  The tile keeps the three rows of node a 74 doubles
// apart (37 x 16 B, odd) so the block-wise 64-bit writes are conflict-free (simulated:
// 162 wavefronts per warp, the minimum).
#include "fe_common.cuh"

namespace cmadx {
namespace {

// K_e leaves the SM as ONE bulk copy per element (cp.async.bulk shared -> global, 4608 B, issued
// by the element's first lane; SASS: UBLKCP): the tile is dense (node blocks 72 doubles apart -
// the block-wise 64-bit writes are conflict-free at this stride too, 162 wavefronts per warp;
// it was the 128-bit READ-BACK of the register path that needed the 74-double padding) and the
// 36 LDS.128 + 18 STG.256 per thread of the read-back loop, with their staging registers, are
// gone.  -DCMADX_HEX8_NO_BULK_STORE restores the register path (A/B measurements).
#ifdef CMADX_HEX8_NO_BULK_STORE
constexpr bool HEX_BULK_STORE = false;
constexpr int HEX_TILE_STRIDE = 74;               // 72 + 2: 37 x 16 B, odd
#else
constexpr bool HEX_BULK_STORE = true;
constexpr int HEX_TILE_STRIDE = 72;
#endif
constexpr int HEX_EPB = FE_BLOCK / 8;             // elements per block

// Layout constants.  K3 (WANT_K): as described above, 76 KB per block -> 3 blocks / SM.
// K4 / K6 (no tangent): records shrink to grad_N + sigma (30 doubles = 15 x 16 B, odd) and
// there is no tile: 344 doubles per element (= 8 mod 16, see above), 44 KB per block ->
// the register file (<= 128 regs) allows 4 blocks / SM.
template <bool WANT_K> struct HexLayout;
template <> struct HexLayout<true> {
    static constexpr int REC = 54, RECS = 24, GN = 22, SW = 46, XI = 456, DU = 512;
    static constexpr int SMEM_DOUBLES = (HEX_EPB / 2) * 1192;      // element pairs: 600 + 592
    CMADX_DEV static int region(int eloc) { return (eloc >> 1) * 1192 + (eloc & 1) * 600; }
};
template <> struct HexLayout<false> {
    static constexpr int REC = 30, RECS = 24, GN = 0, SW = 24, XI = 264, DU = 320;
    static constexpr int SMEM_DOUBLES = HEX_EPB * 344;
    CMADX_DEV static int region(int eloc) { return eloc * 344; }
};

// index of (al, be) in the packed upper triangle of a symmetric 6x6
CMADX_DEV constexpr int sidx(int al, int be) {
    return (al <= be) ? (al * 6 - al * (al - 1) / 2 + (be - al)) : (be * 6 - be * (be - 1) / 2 + (al - be));
}
// position of entry `o` (0..71: row i*24 + column) of node a's row block in the tile
CMADX_DEV int tile_pos(int a, int o) { return a * HEX_TILE_STRIDE + o; }

template <int SOLVER, bool ROT, bool WANT_K>
CMADX_DEV void hex8_point(const FeArgs& A, const int64_t e, const bool live, double* smem,
                           const bool allow_defer) {
    const cmadx_fe_block_t& b = A.b;
    const int lane = threadIdx.x & 31;
    const int ip = lane & 7;                      // this thread's point; also its node in phase C
    const int eloc = threadIdx.x >> 3;
    using LY = HexLayout<WANT_K>;
    constexpr int HEX_REC = LY::REC, HEX_GN = LY::GN, HEX_SW = LY::SW, HEX_XI = LY::XI, HEX_DU = LY::DU;
    double* reg = smem + LY::region(eloc);
    double* recs = reg + LY::RECS;

    // ---- phase A: element data -> shared memory, in contiguous runs
    int eq3[3] = {0, 0, 0};
    double wdv = 0.0;
    if (live) {
        // all global loads first (independent, in flight together), then the stores
#pragma unroll
        for (int k = 0; k < 3; ++k) eq3[k] = __ldg(b.elem_eq + e * 24 + 3 * ip + k);
        const double* g = b.grad_N + e * 192;     // 8 points x 24 doubles = 48 chunks of 32 B
        double c[6][4], xs[7], Un[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) ld256(g + 4 * (q * 8 + ip), c[q][0], c[q][1], c[q][2], c[q][3]);
#pragma unroll
        for (int r = 0; r < 7; ++r) xs[r] = __ldg(b.xi_prev + e * 56 + r * 8 + ip);
        wdv = __ldg(b.quad_w + ip) * __ldg(b.det + e * 8 + ip);
#pragma unroll
        for (int k = 0; k < 3; ++k) Un[k] = __ldg(b.U + eq3[k]);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int m = q * 8 + ip;             // chunk: point m / 6, part m % 6
            double2* dst = reinterpret_cast<double2*>(recs + (m / 6) * HEX_REC + HEX_GN + (m % 6) * 4);
            if ((m % 6) & 4) { dst[1] = make_double2(c[q][2], c[q][3]); dst[0] = make_double2(c[q][0], c[q][1]); }   // bank spread
            else { dst[0] = make_double2(c[q][0], c[q][1]); dst[1] = make_double2(c[q][2], c[q][3]); }
        }
#pragma unroll
        for (int r = 0; r < 7; ++r) reg[HEX_XI + r * 8 + ip] = xs[r];
#pragma unroll
        for (int k = 0; k < 3; ++k) reg[3 * ip + k] = Un[k];
        if constexpr (SOLVER >= FE_JVP) {     // displacement direction of the JVP, next to xi
#pragma unroll
            for (int k = 0; k < 3; ++k) reg[HEX_DU + 3 * ip + k] = A.dU ? __ldg(A.dU + eq3[k]) : 0.0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) reg[3 * ip + k] = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int m = q * 8 + ip;
            double2* dst = reinterpret_cast<double2*>(recs + (m / 6) * HEX_REC + HEX_GN + (m % 6) * 4);
            dst[0] = make_double2(0.0, 0.0);
            dst[1] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int r = 0; r < 7; ++r) reg[HEX_XI + r * 8 + ip] = 0.0;
    }
    __syncwarp();
    double xp[7], eps[6], deps[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 7; ++c) xp[c] = reg[HEX_XI + ip * 7 + c];
    {
        double gN[8][3], U[8][3];
        const double* mine = recs + ip * HEX_REC + HEX_GN;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) { gN[a][k] = mine[3 * a + k]; U[a][k] = reg[3 * a + k]; }
        strain_from_U<8>(U, gN, eps);
        if constexpr (SOLVER >= FE_JVP) {
            if (live) {
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int k = 0; k < 3; ++k) U[a][k] = reg[HEX_DU + 3 * a + k];
                strain_from_U<8>(U, gN, deps);
            }
        }
    }

    // ---- phase B: local Newton at this point
    PointOut o;
    double D[6][6];
    if constexpr (SOLVER >= FE_JVP) {
        double xs[7], dxp[7];
        const int64_t p = e * 8 + ip;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            xs[c] = live ? __ldg(A.xi_state + p * 7 + c) : 0.0;
            dxp[c] = (live && A.dxi_prev) ? __ldg(A.dxi_prev + p * 7 + c) : 0.0;
        }
        point_jvp<SOLVER - FE_JVP, ROT>(A, xp, xs, dxp, eps, deps, live, o);
    } else {
        DevNewton nw = A.nw;
        nw.defer_after = allow_defer ? A.nw.defer_request : 0;
        solve_point<SOLVER, ROT, WANT_K>(A.m, nw, xp, eps, live, o, D);
    }

    // an element is handed to the second pass as a whole (J2 radial hand-backs, deferred points)
    bool ebail = false;
    if (SOLVER < FE_JVP) {
        const unsigned bal = __ballot_sync(0xffffffffu, o.bail);
        ebail = ((bal >> (lane & ~7)) & 0xffu) != 0u;
        if (ebail && live && ip == 0) append_bail(A, e);
    }
    const bool emit = live && !ebail;
    if (emit) {
        const int64_t p = e * 8 + ip;
        if (b.iters) b.iters[p] = o.iters;
        if (b.flags) b.flags[p] = o.flags;
        if (b.sigma) {
#pragma unroll
            for (int a = 0; a < 6; ++a) b.sigma[p * 6 + a] = o.sg[a];
        }
    }
    if (A.mix_eq_p) {
        // primal: p from U; K6: the momentum-stress direction dev(d cauchy) - dp I, dp from dU
        const double* pv = (SOLVER >= FE_JVP) ? A.dU : b.U;
        double p = 0.0;
        if (live && pv) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
                p = fma(__ldg(A.mix_N + ip * 8 + a), __ldg(pv + __ldg(A.mix_eq_p + e * 8 + a)), p);
        }
        mixed_momentum_stress<WANT_K>(p, o.sg, D);
    }
    {   // own record slots and own xi slot: no other thread has touched or read them yet
        double* rec = recs + ip * HEX_REC;
        if (WANT_K) {
#pragma unroll
            for (int al = 0; al < 6; ++al)
#pragma unroll
                for (int be = al; be < 6; ++be)
                    rec[sidx(al, be)] = D[al][be] * (is_diag(be) ? wdv : 0.5 * wdv);
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) rec[HEX_SW + a] = o.sg[a] * wdv;
#pragma unroll
        for (int c = 0; c < 7; ++c) reg[HEX_XI + ip * 7 + c] = o.x[c];
    }
    __syncwarp();
    if (emit) {
        double* xd = b.xi + e * 56;
#pragma unroll
        for (int r = 0; r < 7; ++r) xd[r * 8 + ip] = reg[HEX_XI + r * 8 + ip];
    }

    // ---- phase C: thread a owns the rows 3a..3a+2 of R_e / K_e; sums run over the 8
    // points in fixed order (bit-reproducible)
    const int a = ip;
    const bool want_R = b.R_elem || b.R_global;
    double Racc[3] = {0.0, 0.0, 0.0};
    if constexpr (!WANT_K) {
        if (emit && want_R) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double* rec = recs + q * HEX_REC;
                const double ga0 = rec[HEX_GN + 3 * a], ga1 = rec[HEX_GN + 3 * a + 1], ga2 = rec[HEX_GN + 3 * a + 2];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    Racc[i] = fma(ga2, rec[HEX_SW + vix(2, i)], fma(ga1, rec[HEX_SW + vix(1, i)], fma(ga0, rec[HEX_SW + vix(0, i)], Racc[i])));
            }
        }
    } else {
        // The consistent tangent of an associative model is symmetric (Dh = Dh^T), hence
        // K_ba = K_ab^T: thread a computes the 3x3 node blocks (a, (a+d) mod 8), d = 0..3,
        // plus d = 4 for a < 4 - 36 of the 64 blocks - and the mirror images are filled
        // in through the shared-memory tile.  P[i][be] = sum_j gN[a][j] Dh[v(j,i)][be].
        double acc[5][3][3];
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[d][i][k] = 0.0;
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const double* rec = recs + q * HEX_REC;
            const double ga[3] = {rec[HEX_GN + 3 * a], rec[HEX_GN + 3 * a + 1], rec[HEX_GN + 3 * a + 2]};
#pragma unroll
            for (int i = 0; i < 3; ++i)
                Racc[i] = fma(ga[2], rec[HEX_SW + vix(2, i)], fma(ga[1], rec[HEX_SW + vix(1, i)], fma(ga[0], rec[HEX_SW + vix(0, i)], Racc[i])));
            double P[3][6];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int be = 0; be < 6; ++be)
                    P[i][be] = fma(ga[2], rec[sidx(vix(2, i), be)],
                                   fma(ga[1], rec[sidx(vix(1, i), be)], ga[0] * rec[sidx(vix(0, i), be)]));
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int bb = (a + d) & 7;
                double g0, g1, g2;
                if (d == 0) { g0 = ga[0]; g1 = ga[1]; g2 = ga[2]; }
                else { g0 = rec[HEX_GN + 3 * bb]; g1 = rec[HEX_GN + 3 * bb + 1]; g2 = rec[HEX_GN + 3 * bb + 2]; }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        acc[d][i][k] = fma(P[i][vix(k, 2)], g2, fma(P[i][vix(k, 1)], g1, fma(P[i][vix(k, 0)], g0, acc[d][i][k])));
            }
        }
        __syncwarp();                             // all records consumed: the tile may overwrite them
        double* tile = reg;
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            if (d < 4 || a < 4) {
                const int bb = (a + d) & 7;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        tile[tile_pos(a, i * 24 + 3 * bb + k)] = acc[d][i][k];
                        if (d > 0) tile[tile_pos(bb, k * 24 + 3 * a + i)] = acc[d][i][k];
                    }
            }
        }
        if constexpr (HEX_BULK_STORE) {
            // generic-proxy writes of the tile -> visible to the async proxy, then one bulk copy
            // per element; the group is waited for (source read) before the region is reused
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (emit && ip == 0) {
                const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(b.K_elem + e * 576), "r"(src), "n"(576 * 8) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
        __syncwarp();
        if (emit) {
            // K_e = 144 chunks of 32 B; lane t stores chunk c*8 + t.  Chunk s lives in node
            // block s / 18 at 16-byte units 2(s % 18), +1; the lanes of the upper half read
            // the units in swapped order so that a quarter-warp load hits 8 distinct groups.
            double* Ke = b.K_elem + e * 576;
            const int hi = ip >> 2;
#pragma unroll
            for (int c = 0; c < 18; ++c) {
                const int s = c * 8 + ip;
                const double2* src = reinterpret_cast<const double2*>(tile + (s / 18) * HEX_TILE_STRIDE + (s % 18) * 4);
                const double2 u = src[hi], v = src[hi ^ 1];
                st256(Ke + 4 * s, hi ? v.x : u.x, hi ? v.y : u.y, hi ? u.x : v.x, hi ? u.y : v.y);
            }
        }
        }
    }
    if (emit && want_R) {
        if (b.R_elem) {
#pragma unroll
            for (int i = 0; i < 3; ++i) b.R_elem[e * 24 + 3 * a + i] = Racc[i];
        }
        if (b.R_global) {
#pragma unroll
            for (int i = 0; i < 3; ++i) atomicAdd(b.R_global + eq3[i], Racc[i]);
        }
    }
    if constexpr (WANT_K && HEX_BULK_STORE) {
        // the tile must stay intact until the copy engine has read it (block exit / next list item)
        if (emit && ip == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
__global__ void __launch_bounds__(FE_BLOCK, (SOLVER == 0) ? (WANT_K ? 3 : 4) : 1) fe_hex8_kernel(const __grid_constant__ FeArgs A) {
    extern __shared__ __align__(16) double smem[];
    if (!LIST) {
        const int64_t e = (int64_t)blockIdx.x * HEX_EPB + (threadIdx.x >> 3);
        hex8_point<SOLVER, ROT, WANT_K>(A, e, e < A.b.n_elems, smem, A.bail_count != nullptr);
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
            hex8_point<SOLVER, ROT, WANT_K>(A, e, live, smem, false);
            __syncwarp();
        }
    }
}

template <int SOLVER, bool ROT, bool WANT_K, bool LIST>
struct Hex8Launcher {
    static cudaError_t run(const FeArgs& A, cudaStream_t stream, int sms) {
        auto kern = fe_hex8_kernel<SOLVER, ROT, WANT_K, LIST>;
        const size_t smem = sizeof(double) * HexLayout<WANT_K>::SMEM_DOUBLES;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t nblk = LIST ? 2 * sms : (A.b.n_elems + HEX_EPB - 1) / HEX_EPB;
        kern<<<(unsigned)nblk, FE_BLOCK, smem, stream>>>(A);
        return cudaGetLastError();
    }
};

}  // namespace

cudaError_t launch_fe_hex8(const FeArgs& A, int solver, bool list, cudaStream_t stream, int sms) {
    if (list) return dispatch_fe_list<Hex8Launcher>(A, solver, stream, sms);
    return dispatch_fe<Hex8Launcher, false>(A, solver, stream, sms);
}

}  // namespace cmadx
