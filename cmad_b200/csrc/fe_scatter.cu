// K5 - deterministic segment sums: the R scatter-add and the COO dedup of the FE
// assembly, as gathers over a precomputed CSR of item lists per segment.
//
// Replaces (reference file:line):
//   cmad/fem/assembly.py:715-720    R_block.at[eq].add(R_flat)
//   cmad/fem/assembly.py:906-909    unique_data.at[coo_dedup_scatter].add(vals)
//   cmad/fem/assembly.py:1026-1070  assembled_coo_dedup (the scatter map this plan inverts)
//
// One thread per segment walks its item list in increasing item order, so the
// result is bit-reproducible and equals a sequential scatter-add.  Reads of
// `vals` are gathers (8 useful bytes per 32-byte sector in the worst case); the
// item lists of neighbouring segments come from the same few elements, so the
// sectors are shared through L2.  HBM-bound: 8 B/item vals + 4|8 B/item index +
// 8 B/segment out + 8 B/segment offsets.
#include <atomic>
#include <cstring>
#include <new>
#include <vector>

#include "point_solver.cuh"

struct cmadx_segment_plan {
    int device;
    int64_t n_items, n_segments;
    int wide;            // items indexed with int64 (n_items >= 2^31)
    int64_t* offsets;    // [n_segments + 1] device
    void* items;         // [n_items] device, int32 or int64
};

namespace cmadx {
int cuda_fail(cudaError_t e);
extern std::atomic<int64_t> g_launches;

namespace {

template <class IT>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const int64_t* __restrict__ offsets, const IT* __restrict__ items,
                   const double* __restrict__ vals, double* __restrict__ out, int64_t n_segments,
                   int accumulate) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segments) return;
    const int64_t j0 = __ldg(offsets + s), j1 = __ldg(offsets + s + 1);
    double acc = accumulate ? out[s] : 0.0;
    for (int64_t j = j0; j < j1; ++j) acc += __ldg(vals + (int64_t)__ldg(items + j));
    out[s] = acc;
}

__global__ void index_gather_kernel(const int64_t* __restrict__ index, int64_t n, const double* __restrict__ src,
                                    double* __restrict__ packed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) packed[i] = src[index[i]];
}
__global__ void index_scatter_kernel(const int64_t* __restrict__ index, int64_t n, const double* __restrict__ packed,
                                     double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[index[i]] = packed[i];
}

}  // namespace
}  // namespace cmadx

using namespace cmadx;

extern "C" {

int cmadx_segment_plan_create(const int64_t* seg, int64_t n_items, int64_t n_segments,
                              cmadx_segment_plan_t** plan) {
    if (!plan || n_items < 0 || n_segments < 0 || (n_items > 0 && !seg)) return CMADX_EINVAL;
    *plan = nullptr;
    // counting sort (stable: items of one segment stay in increasing order)
    std::vector<int64_t> off;
    try { off.assign((size_t)n_segments + 1, 0); } catch (const std::bad_alloc&) { return CMADX_ENOMEM; }
    for (int64_t i = 0; i < n_items; ++i) {
        const int64_t s = seg[i];
        if (s < 0 || s >= n_segments) return CMADX_EINVAL;
        ++off[(size_t)s + 1];
    }
    for (int64_t s = 0; s < n_segments; ++s) off[(size_t)s + 1] += off[(size_t)s];
    const int wide = n_items >= ((int64_t)1 << 31);
    const size_t isz = wide ? 8 : 4;
    std::vector<char> items;
    std::vector<int64_t> cur;
    try { items.resize((size_t)n_items * isz); cur.assign(off.begin(), off.end() - 1); }
    catch (const std::bad_alloc&) { return CMADX_ENOMEM; }
    for (int64_t i = 0; i < n_items; ++i) {
        const int64_t pos = cur[(size_t)seg[i]]++;
        if (wide) reinterpret_cast<int64_t*>(items.data())[pos] = i;
        else reinterpret_cast<int32_t*>(items.data())[pos] = (int32_t)i;
    }
    cmadx_segment_plan* p = new (std::nothrow) cmadx_segment_plan();
    if (!p) return CMADX_ENOMEM;
    p->n_items = n_items; p->n_segments = n_segments; p->wide = wide;
    p->offsets = nullptr; p->items = nullptr;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaMalloc(&p->offsets, sizeof(int64_t) * ((size_t)n_segments + 1));
    if (e == cudaSuccess) e = cudaMalloc(&p->items, items.size() > 0 ? items.size() : 8);
    if (e == cudaSuccess)
        e = cudaMemcpy(p->offsets, off.data(), sizeof(int64_t) * ((size_t)n_segments + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !items.empty())
        e = cudaMemcpy(p->items, items.data(), items.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (p->offsets) cudaFree(p->offsets);
        if (p->items) cudaFree(p->items);
        delete p;
        return (e == cudaErrorMemoryAllocation) ? CMADX_ENOMEM : cuda_fail(e);
    }
    *plan = p;
    return CMADX_OK;
}

int cmadx_segment_plan_destroy(cmadx_segment_plan_t* p) {
    if (!p) return CMADX_OK;
    cudaFree(p->offsets);
    cudaFree(p->items);
    delete p;
    return CMADX_OK;
}

int cmadx_segment_sum(const cmadx_segment_plan_t* p, const double* vals, double* out, int accumulate,
                      void* stream) {
    if (!p || !out || (p->n_items > 0 && !vals)) return CMADX_EINVAL;
    if (p->n_segments == 0) return CMADX_OK;
    const unsigned nblk = (unsigned)((p->n_segments + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (p->wide)
        segment_sum_kernel<int64_t><<<nblk, 256, 0, s>>>(p->offsets, (const int64_t*)p->items, vals, out,
                                                         p->n_segments, accumulate);
    else
        segment_sum_kernel<int32_t><<<nblk, 256, 0, s>>>(p->offsets, (const int32_t*)p->items, vals, out,
                                                         p->n_segments, accumulate);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CMADX_OK;
}

// Interface pack / unpack of the halo exchange (the element partition's shared dofs): the two
// index kernels around the NCCL all-reduce, so the step path launches no framework kernels.
int cmadx_index_gather(const int64_t* index, int64_t n, const double* src, double* packed, void* stream) {
    if (n < 0 || (n > 0 && (!index || !src || !packed))) return CMADX_EINVAL;
    if (n == 0) return CMADX_OK;
    index_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(index, n, src, packed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CMADX_OK;
}

int cmadx_index_scatter(const int64_t* index, int64_t n, const double* packed, double* dst, void* stream) {
    if (n < 0 || (n > 0 && (!index || !packed || !dst))) return CMADX_EINVAL;
    if (n == 0) return CMADX_OK;
    index_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(index, n, packed, dst);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CMADX_OK;
}

}  // extern "C"
