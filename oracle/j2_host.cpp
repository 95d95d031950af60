// CPU baseline "port-handderived"  - TEST / BENCH INFRASTRUCTURE ONLY.
//
// The J2 radial-return routine of the CUDA kernel - cmad_b200/csrc/j2_radial.cuh and the
// closed-form outputs of cmad_b200/csrc/mp_update_j2_point.cuh, i.e. the very source nvcc
// compiles into mp_update_j2_kernel - compiled for the host and run over the points with
// OpenMP.  It is the honest multi-core CPU number to put beside the GPU: the AD oracle
// (oracle_c.cpp) is a fidelity restatement of the reference's traced-AD algorithm at
// ~6 us per update per core, not a tuned CPU code.  Never linked into the product library.
//
// The device intrinsics the per-point routine touches are shimmed for one "lane" per call:
// warp votes degenerate to the lane's own predicate, read-only / streaming accesses to plain
// loads and stores.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>

#include <cuda_runtime.h>

#include "host_shims.h"

#include "mp_update_j2_point.cuh"

using namespace cmadx;

extern "C" {

// same argument structs as cmadx_mp_update (host pointers); `dev_mat` / `dev_newton` are the
// converted structs the product library hands back through cmadx_debug_dev_structs.
// Returns the number of points that bailed (they keep their xi_prev; the bench workload has none).
int64_t j2_host_mp_update(const void* dev_mat, const void* dev_newton, const int32_t* active_pid,
                          int32_t n_active, const cmadx_mp_buffers_t* b, int nthreads) {
    MpArgs A;
    memset(&A, 0, sizeof A);
    memcpy(&A.m, dev_mat, sizeof(DevMat));
    memcpy(&A.nw, dev_newton, sizeof(DevNewton));
    A.n_active = n_active;
    for (int c = 0; c < n_active; ++c) A.pid[c] = active_pid[c];
    A.b = *b;
    int64_t bails = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(+ : bails)
    for (int64_t i = 0; i < b->n; ++i) bails += j2_point_update(A, i, true) ? 1 : 0;
    return bails;
}

int j2_host_struct_sizes(int64_t* out2) {
    out2[0] = sizeof(DevMat);
    out2[1] = sizeof(DevNewton);
    return 0;
}

int j2_host_max_threads(void) { return omp_get_max_threads(); }
}
