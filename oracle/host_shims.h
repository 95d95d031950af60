// TEST / BENCH INFRASTRUCTURE ONLY: the device intrinsics the per-point routines of
// cmad_b200/csrc touch, shimmed for one "lane" per call so that the same source builds with the
// host compiler - warp votes degenerate to the lane's own predicate, read-only / streaming accesses
// to plain loads and stores.
#pragma once
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int) { return v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline double __ldg(const double* p) { return *p; }
static inline void __stcs(double* p, double v) { *p = v; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }   // one rounding, never fused
static inline int __syncthreads_or(int p) { return p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, sizeof d); return d; }
static const struct { unsigned x, y, z; } threadIdx = {0u, 0u, 0u};

