// K1-cta - generic-Newton material-point update with BLOCK-LEVEL HAND-OFF of the points that
// still iterate (sm_100a).
//
// Why.  In the one-pass kernels (mp_update.cu) a warp's Newton trips, Jacobian / LU and
// derivative outputs run with the lanes of the easy points idle (ncu: 22.5 / 18.4 of 32 lanes
// per instruction for Hosford a = 4 / 100).  The two remedies tried before write partially
// filled 32-byte sectors - every scattered 8-byte store costs a DRAM read-modify-write:
//   * two-pass deferral (second launch over a list): measured for a = 100 at 2^23 points
//     9.1 GB read + 9.6 GB written for 6.6 GB of algorithmic traffic (profiles/r2l_a100_passes.csv);
//   * warp-level parking / lane refill (mp_update_queue.cu, mp_update_stream.cu): 10x the DRAM reads.
//
// Here a 256-thread block owns a tile of 512 consecutive points and EVERY global store is a full,
// coalesced row of the tile:
//   phase 1  each thread runs its two points up to the first Newton direction they need after
//            `defer_min` updates (0: every plastic point; 2: the hard points of near-Tresca Hosford)
//            - lock-step, the whole block votes the loop (see local_newton) - and leaves a RECORD of
//            the point in shared memory: iterate, residual, yield-surface state, norms, counters.
//            Points that stopped go on the block's hard list;
//   phase 2  the hard list is solved 256 at a time by full warps: a lane restores a record, gathers
//            the point's inputs (L1 / L2 hits) and RESUMES exactly where the owner stopped
//            (NewtonLane::PH_DIR: no residual is re-evaluated), then overwrites the record with the
//            converged one;
//   phase 3  each thread restores the records of its own two points and writes all their outputs.
// Every lane executes the evaluation sequence of the reference loops
// (cmad/models/nonlinear_solver.py:102-155, :14-85; cmad/util/line_search.py:125-181) on the same
// arguments as in the one-pass kernels, and the outputs are computed by the same routine from the
// same state: results are bit-identical to the one-pass kernels'.
#include "mp_outputs.cuh"

namespace cmadx {

namespace {

constexpr int CTA = 256;            // threads per block
constexpr int SUB = 2;              // points per thread in phases 1 and 3
constexpr int TILE = CTA * SUB;     // points per block

template <class Pt, int N> struct RecordLayout {
    using YF = decltype(Pt::yf);
    // x[N], C[N], yield-surface state, f, eD, n0, nc
    static constexpr int X = 0, C = N, YS = 2 * N, N0 = 2 * N + YF::NS + 2, NC = N0 + 1, ND = NC + 1;
    static constexpr size_t BYTES = (size_t)TILE * (ND * sizeof(double) + 2 * sizeof(int)) + (size_t)TILE * sizeof(int);
};

// meta word: bit 0 flag at entry, bit 1 plastic at the recorded state, bit 2 stopped (on the hard list)
CMADX_DEV int pack_meta(int flag_entry, bool plastic, bool stopped) {
    return (flag_entry & 1) | (plastic ? 2 : 0) | (stopped ? 4 : 0);
}

template <int YK, bool ROT>
CMADX_DEV void material_axes(const DevMat& m, const double (&e)[6], double (&em)[6]) {
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

template <int YK, bool ROT, bool REDUCED>
__global__ void __launch_bounds__(CTA, REDUCED ? 2 : 1)
mp_update_cta_kernel(const __grid_constant__ MpArgs A, const int defer_min) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    using RL = RecordLayout<Pt, N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* rec = reinterpret_cast<double*>(smem_raw);                  // [RL::ND][TILE]
    int* rec_ii = reinterpret_cast<int*>(rec + (size_t)RL::ND * TILE);  // [TILE] Newton updates
    int* rec_meta = rec_ii + TILE;                                      // [TILE]
    int* hard = rec_meta + TILE;                                        // [TILE] tile-local slots
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t n = A.b.n, ld = A.b.ld;
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    const DevMat& m = A.m;
    if (tid == 0) s_count = 0;
    __syncthreads();

    DevNewton nw = A.nw;
    nw.defer_after = 0;
    int cnt = 0, n_rounds = SUB;
    for (int r = 0; r < n_rounds; ++r) {
        const bool fresh = r < SUB;
        int slot;
        bool live;
        if (fresh) {
            slot = r * CTA + tid;
            live = tile0 + slot < n;
        } else {
            const int j = (r - SUB) * CTA + tid;
            live = j < cnt;
            slot = live ? hard[j] : 0;
        }
        const int64_t i = tile0 + slot;
        double xp[7], e[6], em[6];
        load_point(A.b, i, live, xp, e);
        material_axes<YK, ROT>(m, e, em);
        Pt pt;
        if constexpr (REDUCED) { pt.shear[0] = xp[1]; pt.shear[1] = xp[2]; pt.shear[2] = xp[4]; }
        double yp[N], Ct[N];
#pragma unroll
        for (int k = 0; k < N; ++k) { yp[k] = xp[Tr::full(k)]; Ct[k] = 0.0; }
        NewtonLane<Pt, N> L;
        {
            double y0[N];
#pragma unroll
            for (int k = 0; k < N; ++k) y0[k] = yp[k];
            if (!REDUCED && fresh && live && A.b.xi_init) {
#pragma unroll
                for (int k = 0; k < N; ++k) y0[k] = __ldg(A.b.xi_init + (int64_t)Tr::full(k) * ld + i);
            }
            L.start(y0);
        }
        if (!fresh && live) {        // resume where the owner stopped: needs a direction at (x, Ct, pt)
            const double* p = rec + slot;
#pragma unroll
            for (int k = 0; k < N; ++k) { L.x[k] = p[(RL::X + k) * TILE]; Ct[k] = p[(RL::C + k) * TILE]; }
            const int meta = rec_meta[slot];
            load_point_state(m, pt, p + RL::YS * TILE, TILE, (meta & 2) != 0);
            L.n0 = p[RL::N0 * TILE];
            L.nc = p[RL::NC * TILE];
            L.ii = rec_ii[slot];
            L.flag_entry = meta & 1;
            L.phase = NewtonLane<Pt, N>::PH_DIR;
        }
        if (!live) L.active = false;
        nw.defer_min = fresh ? defer_min : -1;
        // phase 1 votes block-wide (lock-step: instruction-cache locality); the hard rounds vote per
        // warp - their lanes' counts differ widely and no warp reads another's records before phase 3
        while (fresh ? (__syncthreads_or(L.active ? 1 : 0) != 0) : (__any_sync(0xffffffffu, L.active) != 0)) {
            if (L.active) L.trip(m, nw, pt, yp, em, true, Ct);     // the finishing trip leaves Ct at x, pt fresh there
        }
        if (live) {
            double* p = rec + slot;
#pragma unroll
            for (int k = 0; k < N; ++k) { p[(RL::X + k) * TILE] = L.x[k]; p[(RL::C + k) * TILE] = Ct[k]; }
            save_point_state(pt, p + RL::YS * TILE, TILE);
            p[RL::N0 * TILE] = L.n0;
            p[RL::NC * TILE] = L.nc;
            rec_ii[slot] = L.ii;
            rec_meta[slot] = pack_meta(L.flag_entry, pt.plastic, L.deferred);
        }
        if (fresh) {
            // hard list: warp-aggregated append (the order does not matter, records are per point)
            const bool stop = live && L.deferred;
            const unsigned mk = __ballot_sync(0xffffffffu, stop);
            int base = 0;
            if (mk != 0u && lane == 0) base = atomicAdd(&s_count, __popc(mk));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (stop) hard[base + __popc(mk & ((1u << lane) - 1u))] = slot;
            if (r == SUB - 1) {
                __syncthreads();
                cnt = s_count;
                n_rounds = SUB + (cnt + CTA - 1) / CTA;
            }
        }
    }
    __syncthreads();

    // ---- phase 3: every output of this thread's own points, full rows of the tile
    for (int r = 0; r < SUB; ++r) {
        const int slot = r * CTA + tid;
        const int64_t i = tile0 + slot;
        if (i >= n) continue;
        double xp[7], e[6], em[6];
        load_point(A.b, i, true, xp, e);
        material_axes<YK, ROT>(m, e, em);
        Pt pt;
        if constexpr (REDUCED) { pt.shear[0] = xp[1]; pt.shear[1] = xp[2]; pt.shear[2] = xp[4]; }
        const double* p = rec + slot;
        const int meta = rec_meta[slot];
        load_point_state(m, pt, p + RL::YS * TILE, TILE, (meta & 2) != 0);
        double x[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) x[c] = xp[c];
#pragma unroll
        for (int k = 0; k < N; ++k) x[Tr::full(k)] = p[(RL::X + k) * TILE];
        if (A.b.C) {
#pragma unroll
            for (int c = 0; c < 7; ++c)
                st(A.b.C, c, ld, i, (Tr::local(c) >= 0) ? p[(RL::C + (Tr::local(c) >= 0 ? Tr::local(c) : 0)) * TILE] : 0.0);
        }
        write_point_outputs<YK, ROT, REDUCED>(A, i, x, xp[6], em, pt, rec_ii[slot], meta & 3, p[RL::NC * TILE]);
    }
}

template <int YK, bool ROT, bool REDUCED>
cudaError_t launch_cta_inst(const MpArgs& A, cudaStream_t stream, int defer_min) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    const size_t smem = RecordLayout<Pt, Pt::N>::BYTES;
    auto kern = mp_update_cta_kernel<YK, ROT, REDUCED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (A.b.n + TILE - 1) / TILE;
    kern<<<(unsigned)blocks, CTA, smem, stream>>>(A, defer_min);
    return cudaGetLastError();
}

template <int YK>
cudaError_t launch_cta_yk(const MpArgs& A, cudaStream_t stream, int defer_min) {
    if (YK == CMADX_YIELD_HOSFORD && !A.b.xi_init && !(A.nw.flags & CMADX_NEWTON_F_GENERIC)) {
        constexpr int H = CMADX_YIELD_HOSFORD;
        return A.m.rot ? launch_cta_inst<H, true, true>(A, stream, defer_min)
                       : launch_cta_inst<H, false, true>(A, stream, defer_min);
    }
    return A.m.rot ? launch_cta_inst<YK, true, false>(A, stream, defer_min)
                   : launch_cta_inst<YK, false, false>(A, stream, defer_min);
}

}  // namespace

bool mp_update_cta_supported(const MpArgs& A) {
    return A.b.def_type == CMADX_DEF_FULL_3D && A.m.model == CMADX_MODEL_SMALL_ELASTIC_PLASTIC && A.b.n > 0 &&
           (A.b.n + TILE - 1) / TILE < (int64_t)0x7fffffff;
}

// defer_min: Newton updates a point may take in phase 1 before it is handed to the block's hard list
cudaError_t launch_mp_update_cta(const MpArgs& A, int defer_min, cudaStream_t stream) {
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_cta_yk<CMADX_YIELD_J2>(A, stream, defer_min);
    case CMADX_YIELD_HILL: return launch_cta_yk<CMADX_YIELD_HILL>(A, stream, defer_min);
    case CMADX_YIELD_HOSFORD: return launch_cta_yk<CMADX_YIELD_HOSFORD>(A, stream, defer_min);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cmadx
