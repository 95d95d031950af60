// K1-queue - generic-Newton material-point update with WARP-LEVEL PARKING (sm_100a).
//
// The one-pass kernels (mp_update.cu) give every thread one point; a warp leaves its
// Newton loop when its slowest lane is done.  Newton counts of a batch are multi-modal
// (elastic points: 0 updates; plastic: 2-10 and more), so the Jacobian / LU / line-search
// trips and the derivative outputs run with half-empty warps (ncu, profiles/r2a_k1_*:
// 22.5 / 18.4 of 32 lanes per instruction for Hosford a = 4 / 100, and that average
// includes the fully populated load / first-evaluation / store code).
//
// Here a warp works in ROUNDS.  A round starts either from a fresh tile of 32 consecutive
// points (coalesced loads) or from 32 lanes' worth of PARKED Newton states popped from the
// warp's own queue in shared memory.  Inside a round the warp runs NewtonLane::trip (the
// single residual call site shared with the one-pass kernels) while enough lanes are busy;
// as soon as fewer than `park_below` lanes still iterate - and there is other work to
// combine them with - the complete NewtonLane state of those lanes (iterate, direction,
// line-search scalars, counters: 2 N + 6 doubles, 5 ints) is parked in the queue and the
// round ends.  Finished lanes write all their outputs from registers, once per round.
// A resumed lane continues exactly where it stopped: every lane executes the evaluation
// sequence of the reference loops (cmad/models/nonlinear_solver.py:102-155, :14-85;
// cmad/util/line_search.py:125-181) on the same arguments as in the one-pass kernels, so
// iterates, Newton counts, flags, ||C|| AND the derivative outputs are those of the
// one-pass kernels bit for bit - unlike the streaming kernel (mp_update_stream.cu) no
// residual is re-evaluated.
//
// Stores of a round that started from the queue are scattered over the few tiles the
// parked points came from; they are ordinary write-back stores (not .cs) so that the
// partially written 32-byte sectors are completed in L2 by the other rounds of the same
// warp before they are evicted.  Tiles are handed to the warps in chunks of consecutive
// tiles from a global counter (persistent grid).
#define CMADX_ST_WRITEBACK 1
#include <cstdio>
#include <cstdlib>

#include "mp_outputs.cuh"

namespace cmadx {

namespace {

constexpr int QCAP = 64;             // queue slots per warp (pop at >= 32, a round parks < 32)
constexpr unsigned FULL = 0xffffffffu;

template <int N> struct QueueLayout {
    static constexpr int ND = 2 * N + 6;     // x, dx, n0, nc, al, best_al, best_phi, CC
    static constexpr int NI = 5;             // point, phase, ii, ne, flag_entry
    static constexpr size_t BYTES_PER_WARP = (size_t)QCAP * (ND * sizeof(double) + NI * sizeof(int));
};

template <class Pt, int N>
CMADX_DEV void park_lane(const NewtonLane<Pt, N>& L, int point, double* qd, int* qi, int slot) {
#pragma unroll
    for (int k = 0; k < N; ++k) { qd[k * QCAP + slot] = L.x[k]; qd[(N + k) * QCAP + slot] = L.dx[k]; }
    double* s = qd + 2 * N * QCAP + slot;
    s[0 * QCAP] = L.n0; s[1 * QCAP] = L.nc; s[2 * QCAP] = L.al;
    s[3 * QCAP] = L.best_al; s[4 * QCAP] = L.best_phi; s[5 * QCAP] = L.CC;
    qi[0 * QCAP + slot] = point; qi[1 * QCAP + slot] = L.phase; qi[2 * QCAP + slot] = L.ii;
    qi[3 * QCAP + slot] = L.ne; qi[4 * QCAP + slot] = L.flag_entry;
}

template <class Pt, int N>
CMADX_DEV int unpark_lane(NewtonLane<Pt, N>& L, const double* qd, const int* qi, int slot) {
#pragma unroll
    for (int k = 0; k < N; ++k) { L.x[k] = qd[k * QCAP + slot]; L.dx[k] = qd[(N + k) * QCAP + slot]; }
    const double* s = qd + 2 * N * QCAP + slot;
    L.n0 = s[0 * QCAP]; L.nc = s[1 * QCAP]; L.al = s[2 * QCAP];
    L.best_al = s[3 * QCAP]; L.best_phi = s[4 * QCAP]; L.CC = s[5 * QCAP];
    L.phase = qi[1 * QCAP + slot]; L.ii = qi[2 * QCAP + slot];
    L.ne = qi[3 * QCAP + slot]; L.flag_entry = qi[4 * QCAP + slot];
    L.active = true; L.deferred = false;
    return qi[0 * QCAP + slot];
}

template <int YK, bool ROT, bool REDUCED>
__global__ void __launch_bounds__(MP_BLOCK, REDUCED ? 4 : 2)
mp_update_queue_kernel(const __grid_constant__ MpArgs A, unsigned* __restrict__ chunk_counter,
                       const int chunk_tiles, const int park_below) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    using QL = QueueLayout<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* qd = reinterpret_cast<double*>(smem_raw) + (size_t)warp * QL::ND * QCAP;
    int* qi = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw) + (size_t)(MP_BLOCK / 32) * QL::ND * QCAP) +
              (size_t)warp * QL::NI * QCAP;
    const unsigned lt_mask = (1u << lane) - 1u;

    const int64_t n = A.b.n, ld = A.b.ld;
    const int64_t total_tiles = (n + 31) >> 5;
    int64_t tile_next = 0, tile_end = 0;     // the chunk of consecutive tiles this warp is working through
    bool more = true;                        // the global counter may still hold chunks
    int qn = 0;                              // parked lanes in this warp's queue (warp-uniform)
    DevNewton nw = A.nw;
    nw.defer_after = 0;
    const DevMat& m = A.m;

    for (;;) {
        // ---- source of this round: 32 parked states, else a fresh tile, else the queue's rest
        bool from_q = qn >= 32;
        if (!from_q) {
            if (tile_next >= tile_end && more) {
                unsigned c = 0u;
                if (lane == 0) c = atomicAdd(chunk_counter, 1u);
                c = __shfl_sync(FULL, c, 0);
                const int64_t first = (int64_t)c * chunk_tiles;
                if (first >= total_tiles) {
                    more = false;
                } else {
                    tile_next = first;
                    tile_end = min(total_tiles, first + (int64_t)chunk_tiles);
                }
            }
            if (tile_next >= tile_end) {
                if (qn == 0) break;
                from_q = true;
            }
        }
        NewtonLane<Pt, N> L;
        int64_t i = 0;
        bool live;
        if (from_q) {
            const int take = min(32, qn);
            qn -= take;
            live = lane < take;
            if (live) i = (int64_t)unpark_lane<Pt, N>(L, qd, qi, qn + lane);
            __syncwarp();
        } else {
            i = tile_next * 32 + lane;
            ++tile_next;
            live = i < n;
        }
        double xp[7], e[6], em[6];
        load_point(A.b, i, live, xp, e);
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
        Pt pt;
        if constexpr (REDUCED) { pt.shear[0] = xp[1]; pt.shear[1] = xp[2]; pt.shear[2] = xp[4]; }
        double yp[N];
#pragma unroll
        for (int k = 0; k < N; ++k) yp[k] = xp[Tr::full(k)];
        if (!from_q) {
            double y0[N];
#pragma unroll
            for (int k = 0; k < N; ++k) y0[k] = yp[k];
            if (!REDUCED && live && A.b.xi_init) {
#pragma unroll
                for (int k = 0; k < N; ++k) y0[k] = __ldg(A.b.xi_init + (int64_t)Tr::full(k) * ld + i);
            }
            L.start(y0);
        }
        if (!live) L.active = false;

        // ---- Newton trips while enough lanes are busy
        const bool other_work = (tile_next < tile_end) || more || qn >= 16;
        bool fin = false;
        double Ct[N];
#pragma unroll
        for (int k = 0; k < N; ++k) Ct[k] = 0.0;
        for (;;) {
            if (L.active) {
                L.trip(m, nw, pt, yp, em, true, Ct);     // the finishing trip leaves Ct at x, pt fresh there
                fin = !L.active;
            }
            const unsigned act = __ballot_sync(FULL, L.active);
            if (act == 0u) break;
            if (other_work && __popc(act) < park_below) {
                if (L.active) park_lane<Pt, N>(L, (int)i, qd, qi, qn + __popc(act & lt_mask));
                qn += __popc(act);
                __syncwarp();
                break;
            }
        }

        // ---- every output of the lanes that finished in this round
        if (fin) {
            double x[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) x[c] = xp[c];
#pragma unroll
            for (int k = 0; k < N; ++k) x[Tr::full(k)] = L.x[k];
            if (A.b.C) {
#pragma unroll
                for (int c = 0; c < 7; ++c)
                    st(A.b.C, c, ld, i, (Tr::local(c) >= 0) ? Ct[Tr::local(c) >= 0 ? Tr::local(c) : 0] : 0.0);
            }
            write_point_outputs<YK, ROT, REDUCED>(A, i, x, xp[6], em, pt, L.ii,
                                                  L.flag_entry | ((pt.plastic ? 1 : 0) << 1), L.nc);
        }
        __syncwarp();
    }
}

template <int YK, bool ROT, bool REDUCED>
cudaError_t launch_queue_inst(const MpArgs& A, unsigned* counter, cudaStream_t stream, int sms) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    const size_t smem = (size_t)(MP_BLOCK / 32) * QueueLayout<Pt::N>::BYTES_PER_WARP;
    auto kern = mp_update_queue_kernel<YK, ROT, REDUCED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int resident = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, MP_BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (resident < 1) return cudaErrorLaunchOutOfResources;
    int park_below = 24;
    if (const char* s = std::getenv("CMADX_QUEUE_PARK_BELOW")) park_below = atoi(s);
    if (park_below > 32) park_below = 32;     // a round parks at most 31 lanes: the queue cannot overflow
    if (std::getenv("CMADX_DEBUG_QUEUE"))
        fprintf(stderr, "[cmadx] mp_update_queue<%d,%d,%d>: %d CTAs/SM resident, %zu B smem/CTA, park below %d\n", YK,
                (int)ROT, (int)REDUCED, resident, smem, park_below);
    const int64_t total_tiles = (A.b.n + 31) >> 5;
    const int64_t warps = (int64_t)sms * resident * (MP_BLOCK / 32);
    int64_t chunk = total_tiles / (warps * 8);
    if (chunk < 1) chunk = 1;
    if (chunk > 8) chunk = 8;
    int64_t blocks = (int64_t)sms * resident;
    const int64_t needed = (total_tiles + (MP_BLOCK / 32) - 1) / (MP_BLOCK / 32);
    if (blocks > needed) blocks = needed;
    kern<<<(unsigned)blocks, MP_BLOCK, smem, stream>>>(A, counter, (int)chunk, park_below);
    return cudaGetLastError();
}

template <int YK>
cudaError_t launch_queue_yk(const MpArgs& A, unsigned* counter, cudaStream_t stream, int sms) {
    if (YK == CMADX_YIELD_HOSFORD && !A.b.xi_init && !(A.nw.flags & CMADX_NEWTON_F_GENERIC)) {
        constexpr int H = CMADX_YIELD_HOSFORD;
        return A.m.rot ? launch_queue_inst<H, true, true>(A, counter, stream, sms)
                       : launch_queue_inst<H, false, true>(A, counter, stream, sms);
    }
    return A.m.rot ? launch_queue_inst<YK, true, false>(A, counter, stream, sms)
                   : launch_queue_inst<YK, false, false>(A, counter, stream, sms);
}

}  // namespace

bool mp_update_queue_supported(const MpArgs& A) {
    return A.b.def_type == CMADX_DEF_FULL_3D && A.m.model == CMADX_MODEL_SMALL_ELASTIC_PLASTIC &&
           A.b.n > 0 && A.b.n < (int64_t)0x7fffff00;
}

// `counter`: one zeroed unsigned in device memory (the chunk dispenser of this launch)
cudaError_t launch_mp_update_queue(const MpArgs& A, unsigned* counter, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_queue_yk<CMADX_YIELD_J2>(A, counter, stream, sms);
    case CMADX_YIELD_HILL: return launch_queue_yk<CMADX_YIELD_HILL>(A, counter, stream, sms);
    case CMADX_YIELD_HOSFORD: return launch_queue_yk<CMADX_YIELD_HOSFORD>(A, counter, stream, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cmadx
