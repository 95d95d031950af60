// K1 - batched material-point constitutive update (sm_100a).
//
// One thread per point.  Inputs/outputs are component-major (SoA) so every
// load/store of a warp is one fully coalesced 256-byte transaction; the local
// 7x7 system, its LU factors and all derivative blocks stay in registers.  The
// Newton loop exits warp-wide by ballot.  Outputs are written with streaming
// stores (write-once data, keep L2 for the inputs of the next launch).
//
// Replaces, for a batch of points (reference file:line):
//   cmad/models/nonlinear_solver.py:88-174 (make_newton_solve + IFT rule) or
//   :14-85 (newton_solve), cmad/models/model.py:121-166 (AD Jacobians),
//   cmad/parameters/parameters.py:368-377 (active-column selection).
#include <type_traits>

#include "mp_outputs.cuh"

namespace cmadx {

namespace {

// REDUCED: Hosford through the 4-unknown HosfordPoint (see point_solver.cuh); needs
// the starting iterate to be xi_prev (no xi_init)
template <int YK, bool ROT, bool REDUCED, bool CTA_SYNC = false>
CMADX_DEV void process_point(const MpArgs& A, const int64_t i, const bool live, const bool allow_defer) {
    using Pt = typename std::conditional<REDUCED, HosfordPoint, SepPoint<YK>>::type;
    using Tr = typename std::conditional<REDUCED, HosfordTraits, SepPointTraits<YK>>::type;
    constexpr int N = Pt::N;
    const int64_t ld = A.b.ld;
    const DevMat& m = A.m;

    double xp[7], x[7], e[6];
    load_point(A.b, i, live, xp, e);
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
    for (int c = 0; c < 7; ++c) x[c] = xp[c];
    if (!REDUCED && live && A.b.xi_init) {
#pragma unroll
        for (int c = 0; c < 7; ++c) x[c] = __ldg(A.b.xi_init + c * ld + i);
    }

    Pt pt;
    if constexpr (REDUCED) { pt.shear[0] = xp[1]; pt.shear[1] = xp[2]; pt.shear[2] = xp[4]; }
    double y[N], yp[N], Cy[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { y[k] = x[Tr::full(k)]; yp[k] = xp[Tr::full(k)]; }
    DevNewton nw = A.nw;
    nw.defer_after = allow_defer ? A.nw.defer_request : 0;
    const NewtonResult nr = local_newton<Pt, N, CTA_SYNC>(m, nw, pt, y, yp, em, live, Cy);
    if (CTA_SYNC) __syncthreads();     // outputs start together as well
    if (allow_defer && nw.defer_after > 0)    // second pass (list mode) re-solves the deferred points
        list_append(live && nr.deferred, A.bail_count, A.bail_list, A.bail_cap, (int)i);
    if (!live || nr.deferred) return;
#pragma unroll
    for (int k = 0; k < N; ++k) x[Tr::full(k)] = y[k];
    if (A.b.C) {
#pragma unroll
        for (int c = 0; c < 7; ++c) st(A.b.C, c, ld, i, (Tr::local(c) >= 0) ? Cy[Tr::local(c) >= 0 ? Tr::local(c) : 0] : 0.0);
    }

    write_point_outputs<YK, ROT, REDUCED>(A, i, x, xp[6], em, pt, nr.iters,
                                          nr.flag_entry | ((pt.plastic ? 1 : 0) << 1), nr.cnorm);
}

// one thread per point over the whole batch
template <int YK, bool ROT, bool REDUCED>
__global__ void __launch_bounds__(MP_BLOCK, REDUCED ? 4 : 1)
mp_update_kernel(const __grid_constant__ MpArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    process_point<YK, ROT, REDUCED>(A, i, i < A.b.n, A.bail_count != nullptr);
}

// lock-step variant of the reduced Hosford kernel: 256-thread blocks whose warps vote the Newton
// loop together (see local_newton).  Measured on B200 (2^23 points, a = 4, profiles/r2g_k1.jsonl,
// r2h_k1.jsonl): 2.04 -> 1.80 ms; "no instruction" stalls 2.0 -> 0.2 per issue.  512-thread blocks
// lose the gain to barrier waits and to every warp hitting memory at the same time; tighter
// register budgets (5 / 6 blocks of 128, 3 of 256) gain nothing; the 7x7 kernels (8 warps / SM)
// do not benefit.
constexpr int LOCKSTEP_BLOCK = 256;
template <bool ROT>
__global__ void __launch_bounds__(LOCKSTEP_BLOCK, 2)
mp_update_hosford_lockstep_kernel(const __grid_constant__ MpArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    process_point<CMADX_YIELD_HOSFORD, ROT, true, true>(A, i, i < A.b.n, A.bail_count != nullptr);
}

// list mode: a small grid walks the points the J2 radial kernel handed back
// (bail_count == 0: nothing to do; > bail_cap: the list overflowed, redo all)
template <int YK, bool ROT, bool REDUCED>
__global__ void __launch_bounds__(MP_BLOCK, REDUCED ? 4 : 1)
mp_update_list_kernel(const __grid_constant__ MpArgs A) {
    const unsigned cnt = *A.bail_count;
    if (cnt == 0u) return;
    const bool all = cnt > A.bail_cap;
    const int64_t total = all ? A.b.n : (int64_t)cnt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < total; base += stride) {
        const int64_t j = base + lane;
        const bool live = j < total;
        const int64_t i = live ? (all ? j : (int64_t)A.bail_list[j]) : 0;
        process_point<YK, ROT, REDUCED>(A, i, live, false);
    }
}

template <int YK>
cudaError_t launch_yk(const MpArgs& A, cudaStream_t stream) {
    const int64_t nblk = (A.b.n + MP_BLOCK - 1) / MP_BLOCK;
    if (nblk == 0) return cudaSuccess;
    if (YK == CMADX_YIELD_HOSFORD && !A.b.xi_init && !(A.nw.flags & CMADX_NEWTON_F_GENERIC)) {
        constexpr int H = CMADX_YIELD_HOSFORD;
        if (!A.bail_count && !(A.nw.flags & CMADX_NEWTON_F_ONE_PASS)) {      // single pass: the lock-step blocks
            const unsigned nb = (unsigned)((A.b.n + LOCKSTEP_BLOCK - 1) / LOCKSTEP_BLOCK);
            if (A.m.rot) mp_update_hosford_lockstep_kernel<true><<<nb, LOCKSTEP_BLOCK, 0, stream>>>(A);
            else mp_update_hosford_lockstep_kernel<false><<<nb, LOCKSTEP_BLOCK, 0, stream>>>(A);
            return cudaGetLastError();
        }
        if (A.m.rot) mp_update_kernel<H, true, true><<<(unsigned)nblk, MP_BLOCK, 0, stream>>>(A);
        else mp_update_kernel<H, false, true><<<(unsigned)nblk, MP_BLOCK, 0, stream>>>(A);
        return cudaGetLastError();
    }
    if (A.m.rot) mp_update_kernel<YK, true, false><<<(unsigned)nblk, MP_BLOCK, 0, stream>>>(A);
    else mp_update_kernel<YK, false, false><<<(unsigned)nblk, MP_BLOCK, 0, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_mp_update_sep(const MpArgs& A, cudaStream_t stream) {
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_yk<CMADX_YIELD_J2>(A, stream);
    case CMADX_YIELD_HILL: return launch_yk<CMADX_YIELD_HILL>(A, stream);
    case CMADX_YIELD_HOSFORD: return launch_yk<CMADX_YIELD_HOSFORD>(A, stream);
    case CMADX_YIELD_BARLAT: return launch_yk<CMADX_YIELD_BARLAT>(A, stream);
    }
    return cudaErrorInvalidValue;
}

template <int YK>
cudaError_t launch_list_yk(const MpArgs& A, cudaStream_t stream, int sms) {
    if (YK == CMADX_YIELD_HOSFORD && !A.b.xi_init && !(A.nw.flags & CMADX_NEWTON_F_GENERIC)) {
        constexpr int H = CMADX_YIELD_HOSFORD;
        if (A.m.rot) mp_update_list_kernel<H, true, true><<<(unsigned)(4 * sms), MP_BLOCK, 0, stream>>>(A);
        else mp_update_list_kernel<H, false, true><<<(unsigned)(4 * sms), MP_BLOCK, 0, stream>>>(A);
        return cudaGetLastError();
    }
    if (A.m.rot) mp_update_list_kernel<YK, true, false><<<(unsigned)(2 * sms), MP_BLOCK, 0, stream>>>(A);
    else mp_update_list_kernel<YK, false, false><<<(unsigned)(2 * sms), MP_BLOCK, 0, stream>>>(A);
    return cudaGetLastError();
}

// second pass over the listed points: the J2 radial kernel's hand-backs, or the points the
// generic kernel deferred (DevNewton::defer_after); never defers itself
cudaError_t launch_mp_update_sep_list(const MpArgs& A, cudaStream_t stream) {
    if (A.b.n == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    switch (A.m.yield) {
    case CMADX_YIELD_J2: return launch_list_yk<CMADX_YIELD_J2>(A, stream, sms);
    case CMADX_YIELD_HILL: return launch_list_yk<CMADX_YIELD_HILL>(A, stream, sms);
    case CMADX_YIELD_HOSFORD: return launch_list_yk<CMADX_YIELD_HOSFORD>(A, stream, sms);
    case CMADX_YIELD_BARLAT: return launch_list_yk<CMADX_YIELD_BARLAT>(A, stream, sms);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cmadx
