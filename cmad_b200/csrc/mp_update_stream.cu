// K1-stream - generic-Newton material-point update with LANE REFILL (sm_100a).
//
// The one-pass kernels (mp_update.cu) give every thread one point and a warp leaves
// its Newton loop when its slowest lane is done.  The Newton counts of a batch are
// multi-modal (elastic points: 0 updates, plastic: 2-10 and more with long line
// searches for near-Tresca Hosford exponents), so a warp spends 1.4x (Hosford a = 4)
// to 3x (a = 100) the trips its lanes need on average (ncu, profiles/r2a_k1_*: 22.5 and
// 18.4 of 32 lanes active per instruction), and the two-pass deferral that softened this
// wrote partially filled sectors (5x the algorithmic DRAM reads from read-modify-write).
//
// Here a warp is a 32-lane Newton engine fed from a ring of G groups of 32 consecutive
// points in shared memory:
//   * a lane whose point converged parks the result in the ring and immediately takes the
//     next unassigned point of the window, so every residual evaluation (the single call
//     site of NewtonLane::trip) runs with all lanes busy;
//   * groups are loaded asynchronously (cp.async, one 256-byte coalesced row per
//     component) one to G-1 groups ahead of their use;
//   * when the OLDEST group of the window is complete the warp drains it: lane l writes
//     every output of point 32 g + l - xi, cauchy, the IFT tangent, dC/dp - as whole
//     256-byte rows (the outputs are 87 % of the traffic and stay fully coalesced), frees
//     the slot and issues the next load.  The drain is an out-of-line routine: the live
//     Newton state of the lanes is parked on the stack around it by the call ABI and the
//     Newton loop stays small in the instruction cache.
// Work is handed to the warps in chunks of consecutive groups from a global counter
// (persistent grid, one CTA slot per SM resident CTA).  Every lane runs exactly the
// evaluation sequence of the reference loops (cmad/models/nonlinear_solver.py:102-155,
// :14-85; cmad/util/line_search.py:125-181), so iterates, Newton counts and branch flags
// are those of the one-pass kernels bit for bit; the derivative outputs are computed from
// a re-evaluation of the residual at the converged state.
#include <cstdio>
#include <cstdlib>

#include "mp_outputs.cuh"

namespace cmadx {

namespace {

constexpr int ST_ROWS = 16;          // ring rows (of 32 doubles) per group slot
constexpr int ROW_XP = 0;            // 0..6  xi_prev; rows 0..5 become ep* when the point retires
constexpr int ROW_E = 7;             // 7..12 symmetric strain
constexpr int ROW_ALPHA = 13;        // alpha*
constexpr int ROW_CNORM = 14;        // ||C|| at the returned state
constexpr int ROW_META = 15;         // int2 {Newton updates, flags}
constexpr int SLOT_DOUBLES = ST_ROWS * 32;
constexpr unsigned FULL = 0xffffffffu;

CMADX_DEV void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
CMADX_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// wait until at most `pending` of this thread's committed copy groups are in flight
CMADX_DEV void cp_async_wait_pending(int pending) {
    switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;\n" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 5;\n" ::: "memory"); break;
    }
}

// start the copy of group g (points 32 g .. 32 g + 31) into a ring slot; one commit group
CMADX_DEV void load_group(const cmadx_mp_buffers_t& b, double* slot, int g, int lane) {
    int64_t i = (int64_t)g * 32 + lane;
    if (i >= b.n) i = b.n - 1;                 // tail lanes duplicate the last point (never assigned)
    const int64_t ld = b.ld;
    double* col = slot + lane;
#pragma unroll
    for (int c = 0; c < 7; ++c) cp_async8(col + (ROW_XP + c) * 32, b.xi_prev + c * ld + i);
    if (b.strain_comps == 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) cp_async8(col + (ROW_E + c) * 32, b.strain + c * ld + i);
    } else {
        // grad_u (9 rows, cmad/global_residuals/interpolation.py:36-39): symmetrise on the way in
        double gu[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) gu[c] = __ldg(b.strain + c * ld + i);
        col[(ROW_E + 0) * 32] = gu[0]; col[(ROW_E + 3) * 32] = gu[4]; col[(ROW_E + 5) * 32] = gu[8];
        col[(ROW_E + 1) * 32] = 0.5 * (gu[1] + gu[3]);
        col[(ROW_E + 2) * 32] = 0.5 * (gu[2] + gu[6]);
        col[(ROW_E + 4) * 32] = 0.5 * (gu[5] + gu[7]);
    }
    cp_async_commit();
}

template <bool ROT>
CMADX_DEV void material_strain(const DevMat& m, const double (&e)[6], double (&em)[6]) {
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
}

// every output of the 32 points of a completed group: lane l owns point 32 g + l (coalesced
// 256-byte rows).  Out of line on purpose (see the header comment).
template <int YK, bool ROT, bool REDUCED>
__device__ __noinline__ void drain_group(const MpArgs& A, const double* slot, int g) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)g * 32 + lane;
    if (i >= A.b.n) return;
    const double* col = slot + lane;
    double x[7], e[6], em[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { x[c] = col[(ROW_XP + c) * 32]; e[c] = col[(ROW_E + c) * 32]; }
    x[6] = col[ROW_ALPHA * 32];
    const double alpha_prev = col[(ROW_XP + 6) * 32];
    const double cnorm = col[ROW_CNORM * 32];
    const int2 meta = *reinterpret_cast<const int2*>(col + ROW_META * 32);
    material_strain<ROT>(A.m, e, em);
    // yield-surface state (normal, its derivative data, f, exp(-D alpha)) at x*: one residual
    // evaluation; the previous plastic strain only enters C, which the Newton loop already wrote
    Pt pt;
    if constexpr (REDUCED) { pt.shear[0] = x[1]; pt.shear[1] = x[2]; pt.shear[2] = x[4]; }
    double y[N], yp[N], Cy[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { y[k] = x[Tr::full(k)]; yp[k] = y[k]; }
    yp[Pt::ALPHA] = alpha_prev;
    pt.residual(A.m, y, yp, em, Cy);
    pt.plastic = (meta.y & 2) != 0;            // the branch the Newton loop returned with
    write_point_outputs<YK, ROT, REDUCED>(A, i, x, alpha_prev, em, pt, meta.x, meta.y, cnorm);
}

template <int YK, bool ROT, bool REDUCED, int G>
__global__ void __launch_bounds__(MP_BLOCK, REDUCED ? 4 : 2)
mp_update_stream_kernel(const __grid_constant__ MpArgs A, unsigned* __restrict__ chunk_counter,
                        const int chunk_groups) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_gidx[MP_BLOCK / 32][G];     // global group index held by each ring slot
    __shared__ int s_done[MP_BLOCK / 32][G];     // retired points of that group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* ring = smem + (size_t)warp * G * SLOT_DOUBLES;
    int* gidx = s_gidx[warp];
    int* done = s_done[warp];
    const unsigned lt_mask = (1u << lane) - 1u;

    const int64_t n = A.b.n;
    const int total_groups = (int)((n + 31) >> 5);
    // warp-uniform cursors over this warp's sequence of groups (slot = seq % G)
    int seq_head = 0;        // oldest group not yet drained
    int seq_load = 0;        // next sequence number to load
    int seq_assign = 0;      // group points are currently handed out from
    int col_assign = 0;      //   ... next column of it
    int seq_ready = 0;       // groups below this have landed in shared memory
    int grp_next = 0, grp_end = 0;   // the chunk of consecutive groups this warp is working through
    bool more = true;        // the global counter may still hold chunks

    // the point this lane is solving
    bool has = false;
    int my_slot = 0, my_col = 0;
    double xp[N], em[6];
    Pt pt;
    NewtonLane<Pt, N> L;
    L.active = false;
    DevNewton nw = A.nw;
    nw.defer_after = 0;

    for (;;) {
        // ---- (a) drain the oldest groups while they are complete (in order: coalesced rows)
        while (seq_head < seq_load) {
            const int hs = seq_head % G;
            const int g = gidx[hs];
            const int cnt = (int)min((int64_t)32, n - (int64_t)g * 32);
            if (done[hs] < cnt) break;
            drain_group<YK, ROT, REDUCED>(A, ring + (size_t)hs * SLOT_DOUBLES, g);
            ++seq_head;
            __syncwarp();
        }
        // ---- (b) refill the free slots of the window
        while (seq_load - seq_head < G) {
            if (grp_next >= grp_end) {
                if (!more) break;
                unsigned c = 0u;
                if (lane == 0) c = atomicAdd(chunk_counter, 1u);
                c = __shfl_sync(FULL, c, 0);
                const int64_t first = (int64_t)c * chunk_groups;
                if (first >= total_groups) { more = false; break; }
                grp_next = (int)first;
                grp_end = (int)min((int64_t)total_groups, first + chunk_groups);
            }
            const int s = seq_load % G;
            load_group(A.b, ring + (size_t)s * SLOT_DOUBLES, grp_next, lane);
            if (lane == 0) { gidx[s] = grp_next; done[s] = 0; }
            ++grp_next;
            ++seq_load;
        }
        __syncwarp();
        // ---- (c) idle lanes take the next unassigned points of the window
        unsigned idle = __ballot_sync(FULL, !has);
        while (idle != 0u && seq_assign < seq_load) {
            if (seq_assign >= seq_ready) {          // first use of this group: its copy must have landed
                cp_async_wait_pending(seq_load - 1 - seq_assign);
                __syncwarp();
                seq_ready = seq_assign + 1;
            }
            const int s = seq_assign % G;
            const int g = gidx[s];
            const int cnt = (int)min((int64_t)32, n - (int64_t)g * 32);
            const int avail = cnt - col_assign;
            const int rank = __popc(idle & lt_mask);
            if (!has && rank < avail) {
                my_slot = s;
                my_col = col_assign + rank;
                const double* col = ring + (size_t)s * SLOT_DOUBLES + my_col;
                double xf[7], e[6];
#pragma unroll
                for (int c = 0; c < 7; ++c) xf[c] = col[(ROW_XP + c) * 32];
#pragma unroll
                for (int c = 0; c < 6; ++c) e[c] = col[(ROW_E + c) * 32];
                material_strain<ROT>(A.m, e, em);
                if constexpr (REDUCED) { pt.shear[0] = xf[1]; pt.shear[1] = xf[2]; pt.shear[2] = xf[4]; }
#pragma unroll
                for (int k = 0; k < N; ++k) xp[k] = xf[Tr::full(k)];
                L.start(xp);                        // x0 = xi_prev
                has = true;
            }
            col_assign += min(avail, __popc(idle));
            if (col_assign >= cnt) { ++seq_assign; col_assign = 0; }
            idle = __ballot_sync(FULL, !has);
        }
        // ---- (d) nothing in any lane: the window is drained and the counter is exhausted
        if (!__any_sync(FULL, has)) break;
        // ---- (e) one residual evaluation + what follows it, for every lane that has work; a lane
        //      that finishes parks its result in the ring (its slot / column) right away
        bool fin = false;
        if (has) {
            double Ct[N];
            L.trip(A.m, nw, pt, xp, em, true, Ct);
            if (!L.active) {
                double* col = ring + (size_t)my_slot * SLOT_DOUBLES + my_col;
#pragma unroll
                for (int k = 0; k < N; ++k) col[(Tr::full(k) < 6 ? ROW_XP + Tr::full(k) : ROW_ALPHA) * 32] = L.x[k];
                col[ROW_CNORM * 32] = L.nc;
                *reinterpret_cast<int2*>(col + ROW_META * 32) =
                    make_int2(L.ii, L.flag_entry | ((pt.plastic ? 1 : 0) << 1));
                if (A.b.C) {       // residual at the returned state (diagnostic output): straight to HBM
                    const int64_t i = (int64_t)gidx[my_slot] * 32 + my_col;
#pragma unroll
                    for (int c = 0; c < 7; ++c)
                        st(A.b.C, c, A.b.ld, i, (Tr::local(c) >= 0) ? Ct[Tr::local(c) >= 0 ? Tr::local(c) : 0] : 0.0);
                }
                has = false;
                fin = true;
            }
        }
#pragma unroll
        for (int s = 0; s < G; ++s) {
            const unsigned mk = __ballot_sync(FULL, fin && my_slot == s);
            if (mk != 0u && lane == 0) done[s] += __popc(mk);
        }
        __syncwarp();
    }
}

template <int YK, bool ROT, bool REDUCED>
cudaError_t launch_stream_inst(const MpArgs& A, unsigned* counter, cudaStream_t stream, int sms) {
    constexpr int G = REDUCED ? 3 : 6;
    constexpr int CTAS = REDUCED ? 4 : 2;
    const size_t smem = (size_t)(MP_BLOCK / 32) * G * SLOT_DOUBLES * sizeof(double);
    auto kern = mp_update_stream_kernel<YK, ROT, REDUCED, G>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // the rings want most of the SM's shared memory: without the carveout hint the driver keeps its
    // default L1 / shared split and fewer CTAs than the register budget allows become resident
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int resident = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, MP_BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (resident < 1) return cudaErrorLaunchOutOfResources;
    if (std::getenv("CMADX_DEBUG_STREAM"))
        fprintf(stderr, "[cmadx] mp_update_stream<%d,%d,%d>: %d CTAs/SM resident (want %d), %zu B smem/CTA\n", YK,
                (int)ROT, (int)REDUCED, resident, CTAS, smem);
    const int64_t total_groups = (A.b.n + 31) >> 5;
    const int64_t warps = (int64_t)sms * resident * (MP_BLOCK / 32);
    int64_t chunk = total_groups / (warps * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > 16) chunk = 16;
    int64_t blocks = (int64_t)sms * resident;
    const int64_t needed = (total_groups + (MP_BLOCK / 32) - 1) / (MP_BLOCK / 32);
    if (blocks > needed) blocks = needed;
    kern<<<(unsigned)blocks, MP_BLOCK, smem, stream>>>(A, counter, (int)chunk);
    return cudaGetLastError();
}

template <int YK>
cudaError_t launch_stream_yk(const MpArgs& A, unsigned* counter, cudaStream_t stream, int sms) {
    if (YK == CMADX_YIELD_HOSFORD && !(A.nw.flags & CMADX_NEWTON_F_GENERIC)) {
        constexpr int H = CMADX_YIELD_HOSFORD;
        return A.m.rot ? launch_stream_inst<H, true, true>(A, counter, stream, sms)
                       : launch_stream_inst<H, false, true>(A, counter, stream, sms);
    }
    return A.m.rot ? launch_stream_inst<YK, true, false>(A, counter, stream, sms)
                   : launch_stream_inst<YK, false, false>(A, counter, stream, sms);
}

}  // namespace

bool mp_update_stream_supported(const MpArgs& A) {
    return A.b.def_type == CMADX_DEF_FULL_3D && A.m.model == CMADX_MODEL_SMALL_ELASTIC_PLASTIC &&
           !A.b.xi_init && A.b.n > 0 && A.b.n < (int64_t)0x7fffff00;
}

// `counter`: one zeroed unsigned in device memory (the chunk dispenser of this launch)
cudaError_t launch_mp_update_stream(const MpArgs& A, unsigned* counter, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_stream_yk<CMADX_YIELD_J2>(A, counter, stream, sms);
    case CMADX_YIELD_HILL: return launch_stream_yk<CMADX_YIELD_HILL>(A, counter, stream, sms);
    case CMADX_YIELD_HOSFORD: return launch_stream_yk<CMADX_YIELD_HOSFORD>(A, counter, stream, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cmadx
