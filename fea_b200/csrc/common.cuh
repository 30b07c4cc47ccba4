// Shared device/host helpers for libfea_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "fea_b200.h"

// Checked build (python -m fea_b200.build --checked -> libfea_b200_checked.so): device-side assertions on
// every index the kernels derive from the pattern -- staging offsets and capacities of the bulk-copy ring,
// slot searches of the assembly, partial-sum slots, peer-slot indices.  compute-sanitizer is closed on
// the GPU pool this was developed on, so this is the memcheck of the project: tests/test_gpu_parity.py::
// test_checked_build runs tools/sanitize.py (every kernel family, small cases, results compared with the
// oracle) against it; a violated assertion prints file:line and fails the CUDA context.
#ifdef FEA_CHECKED
#include <cassert>
#define FEA_ASSERT(cond) assert(cond)
#else
#define FEA_ASSERT(cond) ((void)0)
#endif

namespace fea {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Last CUDA error seen by any entry point (reported by fea_last_cuda_error()).
void set_last_error(cudaError_t e);

// Profiling counters (read through fea_profile_read): kernels launched by this library, and a
// sampled CUDA-event timing of the PCG SpMV kernel taken inside fea_pcg_solve.
struct Profile {  // every field is atomic or guarded: entry points may be called from several host threads
  std::atomic<long long> launches{0};
  std::atomic<int> enabled{0};
  std::atomic<long long> spmv_samples{0};
  std::atomic<long long> spmv_ns{0};  // sum of sampled SpMV durations, nanoseconds
  std::atomic<long long> pcg_iterations{0};
  void add_spmv_sample(float ms) {
    spmv_ns += (long long)((double)ms * 1e6);
    spmv_samples += 1;
  }
};
Profile& profile();

// Checks the launch(es) just made and counts them (n = kernels launched since the last check).
inline int check_launch(int n = 1) {
  profile().launches += n;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(e);
    return FEA_ERR_CUDA;
  }
  return FEA_OK;
}

inline int check(cudaError_t e) {
  if (e != cudaSuccess) {
    set_last_error(e);
    return FEA_ERR_CUDA;
  }
  return FEA_OK;
}

// Per-thread pinned host scratch for the solvers' state snapshots: allocated once and kept
// (cudaMallocHost / cudaFreeHost cost 0.1-300 ms per call and cudaFreeHost synchronises the device,
// which dominated short solves).  Returns nullptr on failure.  `slot` 0..3 are independent buffers
// (a solver uses one; nesting solvers in one thread is not supported by the C ABI anyway).
void* pinned_scratch(int slot, size_t bytes);

#define FEA_TRY(expr)                    \
  do {                                   \
    int _rc = (expr);                    \
    if (_rc != FEA_OK) return _rc;       \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32).  Result valid in thread 0.
// `scratch` is >= 32 doubles of shared memory.  Fixed reduction tree => deterministic.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    t = lane < nw ? scratch[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;
}

// Record a data-dependent error in a zero-initialised slot: status[0] = max code seen,
// status[1] = 0x7fffffff - (smallest offending index)  (atomicMax keeps the smallest index).
__device__ __forceinline__ void raise_status(int32_t* status, int code, int index) {
  if (status == nullptr) return;
  atomicMax(&status[0], code);
  atomicMax(&status[1], 0x7fffffff - index);
}

// System-scope release / acquire on 64-bit flags: the cross-GPU tag protocol (pcg_common.cuh) and the
// in-kernel halo gate of the SpMV (spmv_tma.cuh).
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// Bounded spin: true when *p reaches `want` (exactly, or at least with `at_least`).  2^24 polls are
// about 2 s; the first exchange of a solve passes a larger bound (ranks arrive with host-side skew).
__device__ __forceinline__ bool spin_until(const long long* p, long long want, bool at_least, int polls = 1 << 24) {
  for (int i = 0; i < polls; ++i) {
    const long long v = ld_acquire_sys(p);
    if (at_least ? v >= want : v == want) return true;
    __nanosleep(100);
  }
  return false;
}
// The same with exponential back-off (0.2 .. 3.2 us), for flags that MANY threads wait on at once: the
// halo tags are polled by every CTA of the SpMV, and ~900 pollers at 10 MHz each saturate the one L2
// slice that the neighbour's NVLink write of the very same tag has to get through.  ~2 s bound.
__device__ __forceinline__ bool spin_until_backoff(const long long* p, long long want) {
  unsigned ns = 200;
  for (int i = 0; i < (1 << 20); ++i) {
    if (ld_acquire_sys(p) >= want) return true;
    __nanosleep(ns);
    if (ns < 3200) ns *= 2;
  }
  return false;
}

// x^3 rounded once (to within a double-rounding tie of the correctly rounded cube), like libm's pow(x, 3)
// that the reference's `element_length**3` calls (euler_bernoulli.py:22).  x*x*x rounds twice and lands an
// ulp away about every third argument; harmless at 1e-10 -- but the Hermite beam matrix has cond ~ 5 n^4,
// and at 100 k elements the LAST BIT of its entries decides whether the exact solution of the assembled
// matrix is 1e-5 or 3e-2 away from the analytic deflection (tests: test_beam_chain_solver_at_size).
__device__ __forceinline__ double cube_rn(double x) {
  const double p = __dmul_rn(x, x), pe = __fma_rn(x, x, -p);      // x^2 = p + pe exactly
  const double q = __dmul_rn(p, x), qe = __fma_rn(p, x, -q);      // p x = q + qe exactly
  return __dadd_rn(q, __dadd_rn(qe, __dmul_rn(pe, x)));
}

// Programmatic dependent launch (sm_90+).  A kernel of the PCG iteration lets the NEXT kernel of the stream be
// scheduled as soon as all of its own CTAs are running (launch_dependents, first instruction), and waits for the
// PREVIOUS kernel to have completed and flushed (wait) before it touches anything that kernel wrote: launch
// latency, CTA scheduling and each kernel's prologue then overlap the tail of the kernel before.  Both are no-ops
// in a kernel that was launched without cudaLaunchAttributeProgrammaticStreamSerialization.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Kernel launch with or without the programmatic-serialisation attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Streaming (evict-first) loads for data read exactly once per kernel: keeps L2 for the vectors.
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }

// Device-wide exclusive scan of int32 (out has n+1 entries, out[n] = total).
// `block_sums` needs ceil(n / kScanChunk) + 1 ints.  Also writes max(in) to *max_out if non-null.
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* block_sums,
                       int32_t* max_out, cudaStream_t stream);
inline size_t scan_workspace_ints(int64_t n) { return (size_t)ceil_div(n, kScanChunk) + 2; }

}  // namespace fea
