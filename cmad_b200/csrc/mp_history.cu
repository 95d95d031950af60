// Fused forward history: ONE launch walks every material point through all N load steps
// (newton_solve per step, state carried in registers), instead of N launches of K1 that
// re-read xi_{t-1} from HBM.  Replaces the forward time-step loop of the material-point
// objectives (cmad/objectives/mp_objective.py:62-90, cmad/cli/primal.py:158-175) for a batch.
// Per point-step: reads the strain slab (48 / 72 B), writes xi_t (56 B) and the Newton count
// (4 B): 108 B instead of K1's 164 B, and the launch count drops from 2-3 N to 1-2 - the
// small-batch calibration loop (a few experiments x 100 steps) was launch-bound.
// Same per-step routines as K1 (j2_radial.cuh / point_solver.cuh through fe_common.cuh's
// solve_point), hence the same iterates, counts and states (asserted exactly in tests).  A
// point whose J2 radial-return solve leaves its regime at some step is re-walked from step 1
// by the generic solver in a second, list-mode launch.  FULL_3D, identity material axes;
// everything else keeps the per-step path (cmadx_mp_forward_history in api.cu).
#include "fe_common.cuh"

namespace cmadx {

struct HistArgs {
    DevMat m;
    DevNewton nw;
    cmadx_mp_history_t h;
    unsigned* bail_count;
    int* bail_list;
    unsigned bail_cap;
};

namespace {

template <int SOLVER, bool LIST>
__global__ void __launch_bounds__(MP_BLOCK, (SOLVER == 0) ? 4 : 1)
mp_history_kernel(const __grid_constant__ HistArgs A) {
    const int64_t ld = A.h.ld;
    const int N = A.h.nsteps, sc = A.h.strain_comps;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < A.h.n;
    if (LIST) {
        const unsigned cnt = *A.bail_count;
        if (cnt == 0u) return;
        const bool all = cnt > A.bail_cap;
        const int64_t total = all ? A.h.n : (int64_t)cnt;
        // one pass is enough: the list kernel is launched with >= total threads
        live = i < total;
        i = live ? (all ? i : (int64_t)A.bail_list[i]) : 0;
    }
    double xp[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) xp[c] = live ? __ldg(A.h.xi_hist + c * ld + i) : 0.0;
    bool dead = false;                       // J2 radial: handed to the generic re-walk
    // software pipeline: the strain of step t + 1 is in flight while step t is solved
    double raw[9];
    auto fetch = [&](int t) {
        const double* es = A.h.strain + (int64_t)t * sc * ld + i;
        if (sc == 6) {
#pragma unroll
            for (int c = 0; c < 6; ++c) raw[c] = __ldg(es + c * ld);
        } else {
#pragma unroll
            for (int c = 0; c < 9; ++c) raw[c] = __ldg(es + c * ld);
        }
    };
#pragma unroll
    for (int c = 0; c < 9; ++c) raw[c] = 0.0;
    if (live && N >= 1) fetch(1);
    for (int t = 1; t <= N; ++t) {
        const bool on = live && !dead;
        double e[6] = {1e-3, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (on) {
            if (sc == 6) {
#pragma unroll
                for (int c = 0; c < 6; ++c) e[c] = raw[c];
            } else {
                e[0] = raw[0]; e[3] = raw[4]; e[5] = raw[8];
                e[1] = 0.5 * (raw[1] + raw[3]); e[2] = 0.5 * (raw[2] + raw[6]); e[4] = 0.5 * (raw[5] + raw[7]);
            }
            if (t < N) fetch(t + 1);
        }
        PointOut o;
        double D[6][6];
        solve_point<SOLVER, false, false>(A.m, A.nw, xp, e, on, o, D);
        if (SOLVER == 0 && on && o.bail) dead = true;
        if (on && !dead) {
            double* xs = A.h.xi_hist + (int64_t)t * 7 * ld + i;
#pragma unroll
            for (int c = 0; c < 7; ++c) { __stcs(xs + c * ld, o.x[c]); xp[c] = o.x[c]; }
            if (A.h.iters_hist) A.h.iters_hist[(int64_t)t * ld + i] = o.iters;
        }
        if (__all_sync(0xffffffffu, !live || dead)) break;      // warp-uniform exit
    }
    if (SOLVER == 0) list_append(live && dead, A.bail_count, A.bail_list, A.bail_cap, (int)i);
}

template <bool LIST>
cudaError_t launch_generic(const HistArgs& A, unsigned nblk, cudaStream_t s) {
    switch (A.m.yield) {
    case CMADX_YIELD_J2: mp_history_kernel<1, LIST><<<nblk, MP_BLOCK, 0, s>>>(A); break;
    case CMADX_YIELD_HILL: mp_history_kernel<2, LIST><<<nblk, MP_BLOCK, 0, s>>>(A); break;
    case CMADX_YIELD_HOSFORD: mp_history_kernel<3, LIST><<<nblk, MP_BLOCK, 0, s>>>(A); break;
    case CMADX_YIELD_BARLAT: mp_history_kernel<4, LIST><<<nblk, MP_BLOCK, 0, s>>>(A); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

// radial: J2 radial-return first pass + generic list pass; else one generic pass
cudaError_t launch_mp_history(const HistArgs& A, bool radial, cudaStream_t s) {
    if (A.h.n == 0 || A.h.nsteps == 0) return cudaSuccess;
    const unsigned nblk = (unsigned)((A.h.n + MP_BLOCK - 1) / MP_BLOCK);
    if (!radial) return launch_generic<false>(A, nblk, s);
    mp_history_kernel<0, false><<<nblk, MP_BLOCK, 0, s>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the list may hold up to n entries (bail_cap >= n is guaranteed by the caller)
    return launch_generic<true>(A, nblk, s);
}

}  // namespace cmadx
