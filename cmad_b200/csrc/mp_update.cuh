// Shared launch-argument structs of the material-point kernels.
#pragma once
#include "point_solver.cuh"

namespace cmadx {

constexpr int MP_BLOCK = 128;

struct MpArgs {
    DevMat m;
    DevNewton nw;
    int n_active;
    int pid[CMADX_MAX_ACTIVE];
    cmadx_mp_buffers_t b;
    // fallback ("bail") list written by the J2 radial kernel and consumed by the
    // generic kernel in list mode; all NULL/0 for a plain full-batch launch
    unsigned* bail_count;
    int* bail_list;
    unsigned bail_cap;
};

// host-side: validate + convert the C-ABI structs (defined in api.cu)
int make_dev_mat(const cmadx_material_t* mat, DevMat* out);
int make_dev_newton(const cmadx_newton_t* nw, DevNewton* out);

// SmallRateElasticPlastic under PLANE_STRESS / UNIAXIAL_STRESS (mp_update_rate_dt.cu)
cudaError_t launch_mp_update_rate_dt(const MpArgs& A, cudaStream_t stream);
cudaError_t launch_mp_update_sep(const MpArgs& A, cudaStream_t stream);
// J2 radial-return kernel (valid for yield J2, no rotation, no xi_init)
cudaError_t launch_mp_update_j2(const MpArgs& A, cudaStream_t stream);
// generic kernel over the bail list (small persistent grid)
cudaError_t launch_mp_update_sep_list(const MpArgs& A, cudaStream_t stream);
// generic Newton with lane refill (mp_update_stream.cu): persistent grid, points handed out in
// chunks from `counter` (one zeroed unsigned in device memory)
bool mp_update_stream_supported(const MpArgs& A);
cudaError_t launch_mp_update_stream(const MpArgs& A, unsigned* counter, cudaStream_t stream);
// generic Newton with warp-level parking of unfinished lanes (mp_update_queue.cu): persistent grid,
// tiles handed out in chunks from `counter` (one zeroed unsigned in device memory)
bool mp_update_queue_supported(const MpArgs& A);
cudaError_t launch_mp_update_queue(const MpArgs& A, unsigned* counter, cudaStream_t stream);
// generic Newton with block-level hand-off of the points that still iterate (mp_update_cta.cu)
bool mp_update_cta_supported(const MpArgs& A);
cudaError_t launch_mp_update_cta(const MpArgs& A, int defer_min, cudaStream_t stream);
cudaError_t launch_mp_update_elastic(const MpArgs& A, cudaStream_t stream);
// SmallRateElasticPlastic (rate form; `strain` = strain increment)
cudaError_t launch_mp_update_rate(const MpArgs& A, cudaStream_t stream);
// PLANE_STRESS / UNIAXIAL_STRESS deformation types (n_xi = 8 / 9)
cudaError_t launch_mp_update_dt(const MpArgs& A, cudaStream_t stream);

}  // namespace cmadx
