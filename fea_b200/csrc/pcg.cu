// SpMV entry points and the Jacobi-preconditioned conjugate-gradient solver.
//
// Replaces `np.linalg.solve(reduced_K, reduced_forces)` (cubebeam.py:98, fea.py:105,
// euler_bernoulli.py:69 -- the author's "# TODO iterative solver", cubebeam.py:99) and the
// reaction product `K @ u` (cubebeam.py:106).
//
// Dirichlet elimination costs nothing at solve time: constrained DOF carry dinv = 0, hence
// z = dinv r, p and x stay exactly 0 there, their columns contribute nothing to K p, and their
// rows never enter a dot product (every dot is weighted by dinv or masked by dinv != 0).  The
// iteration is therefore the PCG of the reference's reduced system K_ff u_f = f_f.
//
// One iteration = 3 kernels, all HBM-bound, no host synchronisation:
//   step 1  ap = K p                       + partial sums of p.ap          (fused in the SpMV)
//   step 2  x += a p;  r -= a ap           + partial sums of r.dinv.r, r.r (fused)
//   step 3  p = dinv r + (rz_new/rz) p     + convergence test / bookkeeping in the last block
// Cross-block reductions: per-block partial -> last block to finish sums them in a fixed order
// (deterministic; no floating-point atomics).  Every kernel starts with `if (done) return`, so
// the host enqueues iterations in chunks and polls a pinned copy of the state without ever
// stalling the GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "pcg_common.cuh"
#include "pcg_fused.cuh"

namespace fea {

// Per-device facts and launch configurations.  cudaFuncSetAttribute and the occupancy numbers are
// per device: a process that drives several GPUs (the C ABI allows it) must not reuse device 0's.
constexpr int kMaxDevices = 64;
struct DeviceInfo {
  int sm_count = 0, smem_optin = 0;
};
static const DeviceInfo* device_info() {
  static DeviceInfo info[kMaxDevices];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  DeviceInfo& di = info[dev];
  if (di.sm_count == 0) {
    cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return &di;
}
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}

TmaPlan tma_plan(int d, int max_coupled, const void* values, const void* node_colidx, int64_t n_nodes,
                 bool small_problem) {
  TmaPlan plan;
  plan.ok = false;
  plan.max_grid = 0;
  plan.groups = 0;
  plan.layout = TmaLayout{0, 0, 0, 0};
  if (max_coupled < 1 || d < 1 || d > 3) return plan;
  if ((reinterpret_cast<uintptr_t>(values) & 15u) || (reinterpret_cast<uintptr_t>(node_colidx) & 15u)) return plan;
  const DeviceInfo* di = device_info();
  if (di == nullptr || di->sm_count <= 0 || di->smem_optin <= 0) return plan;
  if (n_nodes >= (int64_t)INT32_MAX / 4) return plan;
  // Shape of the pipeline: consumer groups per CTA, ring stages per CTA, CTAs (rings) per SM.
  int cfg_groups = 2, cfg_stages = 2, cfg_ctas = 3;  // best of the sweep on B200: profiles/tma_sweep_r01.log
  // A few tiles per consumer group (a 133 k-DOF mesh has 2784 tiles for 888 groups) leave the 2-stage rings no
  // prefetch depth and a 4-against-3 tile imbalance: 3 groups x 3 stages on 2 CTAs per SM measured -7 % per PCG
  // iteration at 31 k / 133 k DOF and -2 % at 0.4 M / 1 M DOF (profiles/small_mesh_sweep_r02.txt).
  if (small_problem) cfg_groups = 3, cfg_stages = 3, cfg_ctas = 2;
  if (const char* env = std::getenv("FEA_TMA_CFG")) {
    int g = 0, s = 0, c = 0;
    if (std::sscanf(env, "%d,%d,%d", &g, &s, &c) == 3 && g >= 1 && g <= kTmaMaxGroups && s >= 1 &&
        s <= kTmaMaxStages && c >= 1 && c <= 8)
      cfg_groups = g, cfg_stages = s, cfg_ctas = c;
  }
  // shared memory per SM is split between the CTAs; keep ~1 KB per CTA for static smem + reserve
  const size_t per_cta = ((size_t)di->smem_optin + 1024) / cfg_ctas - 2048;
  // The ring length is a multiple of the group count (a stage is always consumed by the same
  // group); drop groups until one ring fits.
  int groups = cfg_groups;
  plan.layout = tma_layout(d, groups, max_coupled, cfg_stages, per_cta);
  while (plan.layout.stages < 1 && groups > 1) {
    --groups;
    plan.layout = tma_layout(d, groups, max_coupled, cfg_stages, per_cta);
  }
  if (plan.layout.stages < 1) return plan;  // tile wider than shared memory
  const int64_t tiles = ceil_div(n_nodes, kTileNodes);
  plan.max_grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)di->sm_count * cfg_ctas));
  plan.groups = groups;
  plan.ok = true;
  // L2 policy of the matrix stream: a matrix that fits in L2 next to the solver's vectors is kept there between the
  // SpMVs of a solve (evict-last: -5 % per iteration at 133 k DOF / 85 MB); anything larger is streamed evict-first so
  // that L2 keeps the gathered vector instead (evict-last costs +9 % at 0.4 M DOF / 280 MB).  Upper bound of the
  // matrix bytes from the widest row.  FEA_TMA_L2=0|1|2 overrides (evict-first / normal / evict-last).
  const double matrix_bytes = (double)n_nodes * max_coupled * (8.0 * d * d + 4.0);
  plan.l2_hint = matrix_bytes <= 100e6 ? 2 : 0;
  if (const char* env = std::getenv("FEA_TMA_L2")) {
    if (env[0] >= '0' && env[0] <= '2') plan.l2_hint = env[0] - '0';
  }
  return plan;
}

// Block partial -> global slot; returns true in the last block to finish (all threads).
// `sys`: the block also stored rows into a peer GPU's memory; its leader fences at system scope (one
// fence per block, after the block-wide barrier of block_sum, instead of one per thread) so that the
// tag the LAST block releases afterwards is ordered behind every block's NVLink stores.
__device__ __forceinline__ bool publish_partials(double* partials, int n_scalars, const double* vals,
                                                 uint32_t* counter, bool sys = false) {
  __shared__ bool s_last;
  FEA_ASSERT(gridDim.x <= (unsigned)kMaxPartials && n_scalars <= 2);
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_scalars; ++s) partials[s * kMaxPartials + blockIdx.x] = vals[s];
    if (sys)
      __threadfence_system();
    else
      __threadfence();
    const uint32_t ticket = atomicAdd(counter, 1u);
    s_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  return s_last;
}

// Sum of partials[0..gridDim.x) in a fixed order; result valid in thread 0.
__device__ __forceinline__ double reduce_partials(const double* partials, double* scratch) {
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partials + i);
  return block_sum(t, scratch);
}

// Last block of a PCG SpMV: p.Ap of this rank -> state; with peers, warp 0 also publishes it.
__device__ __forceinline__ void finish_pap(const double* partials, double* s_red, PcgState* st, const PeerView* pv,
                                           const PeerKey& key) {
  const double s = reduce_partials(partials, s_red);
  if (threadIdx.x == 0) {
    st->pap = s;
    st->counter[0] = 0;
    s_red[0] = s;
  }
  if (pv != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0 && key.own->error != 0) peer_failure(key, st);  // a halo gate timed out
    if (threadIdx.x == 0) dbg_stamp(key.own, st->iter, 5);
    if (threadIdx.x < 32) peer_publish(*pv, 1, st->iter, s_red[0], 0.0);
  }
}

// Multi-GPU: hold the calling thread until the neighbours' halo rows of this iteration have landed.
__device__ __forceinline__ bool wait_halo(const PeerKey& key, long long iter) {
  const long long tag = peer_tag(key, iter);
  bool ok = true;
  if (key.has_lower) ok = spin_until(&key.own->halo_tag[0], tag, true) && ok;
  if (key.has_upper) ok = spin_until(&key.own->halo_tag[1], tag, true) && ok;
  return ok;
}

template <int D>
__global__ void __launch_bounds__(kSpmvThreads) spmv_kernel(int64_t n_nodes, const int32_t* __restrict__ node_rowptr,
                                                            const int32_t* __restrict__ node_colidx,
                                                            const double* __restrict__ values,
                                                            const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t node = (int64_t)blockIdx.x * kSpmvWarps + warp; node < n_nodes;
       node += (int64_t)gridDim.x * kSpmvWarps) {
    const int lo = node_rowptr[node];
    const int cnt = node_rowptr[node + 1] - lo;
    double out[D];
    spmv_node<D>(node_colidx, values, x, lo, cnt, lane, out);
    if (lane < D) {
      double o = out[0];
#pragma unroll
      for (int a = 1; a < D; ++a) o = lane == a ? out[a] : o;
      y[node * D + lane] = o;
    }
  }
}

// step 1: ap = K p over the owned nodes, fused with the partial p.ap.
template <int D>
__global__ void __launch_bounds__(kSpmvThreads)
pcg_spmv_kernel(int64_t n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                const double* __restrict__ values, const double* __restrict__ p, double* __restrict__ ap,
                int64_t p_row_offset, PcgState* st, double* partials, const PeerView* pv, PeerKey key) {
  __shared__ double s_red[32];
  if (st->done) return;
  if (pv != nullptr) {  // no tile ordering here: every block waits for the halo before its first gather
    if (threadIdx.x == 0 && !wait_halo(key, st->iter)) key.own->error = 1;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* p_own = p + p_row_offset * D;
  double dot = 0.0;
  for (int64_t node = (int64_t)blockIdx.x * kSpmvWarps + warp; node < n_nodes;
       node += (int64_t)gridDim.x * kSpmvWarps) {
    const int lo = node_rowptr[node];
    const int cnt = node_rowptr[node + 1] - lo;
    double out[D];
    spmv_node<D>(node_colidx, values, p, lo, cnt, lane, out);
    if (lane < D) {
      double o = out[0];
#pragma unroll
      for (int a = 1; a < D; ++a) o = lane == a ? out[a] : o;
      ap[node * D + lane] = o;
      dot = fma(o, p_own[node * D + lane], dot);
    }
  }
  const double total = block_sum(dot, s_red);
  if (publish_partials(partials, 1, &total, &st->counter[0])) finish_pap(partials, s_red, st, pv, key);
}

// step 1, bulk-copy pipeline variant (spmv_tma.cuh)
// The gated variant is held to the ungated one's register budget (3 CTAs per SM at <3,2>): left alone
// ptxas gives it 126 registers and one CTA per SM.
template <int D, int G, bool GATED>
__global__ void __launch_bounds__(tma_threads(D, G))
pcg_spmv_tma_kernel(int n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                    const double* __restrict__ values, const double* p, double* __restrict__ ap,
                    const double* p_own, int stages, int val_cap, int col_cap, PcgState* st,
                    double* partials, const PeerView* pv, PeerKey key) {
  extern __shared__ __align__(128) unsigned char s_tma[];
  __shared__ double s_red[32];
  pdl_launch_dependents();
  pdl_wait();
  if (st->done) return;
  double dot = 0.0;
  HaloGate gate{};
  if (GATED) {  // multi-GPU: interior tiles first, the halo tags are only looked at before a face tile
    gate.tag_lower = key.has_lower ? &key.own->halo_tag[0] : nullptr;
    gate.tag_upper = key.has_upper ? &key.own->halo_tag[1] : nullptr;
    gate.iter = &st->iter;
    gate.epoch = key.epoch;
    gate.lower_tiles = key.lower_tiles;
    gate.upper_tiles = key.upper_tiles;
    gate.error = &key.own->error;
    gate.aligned = key.aligned;
    if (blockIdx.x == 0 && threadIdx.x == 0) dbg_stamp(key.own, st->iter, 0);
  }
  spmv_tma_body<D, G, true, GATED>(n_nodes, node_rowptr, node_colidx, values, p, ap, p_own, stages, val_cap, col_cap,
                                   s_tma, dot, gate);
  const double total = block_sum(dot, s_red);
  if (publish_partials(partials, 1, &total, &st->counter[0])) finish_pap(partials, s_red, st, pv, key);
}

// Launch configuration of one kernel instance, cached per device.
struct GridCache {
  int grid[kMaxDevices] = {};
  long long key[kMaxDevices];
  std::mutex mu;
  GridCache() {
    for (auto& k : key) k = -1;
  }
};

template <typename Kernel>
static int tma_grid(GridCache& cache, Kernel kernel, const TmaPlan& plan, int threads, int* grid_out) {
  const TmaLayout& L = plan.layout;
  const long long key = ((long long)plan.max_grid << 32) | (long long)L.smem_bytes;
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(cache.mu);
  if (cache.key[dev] != key) {
    FEA_TRY(check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes)));
    // Persistent kernel: never launch more CTAs than can be resident at once (a partial second
    // wave would idle most of the chip), whatever registers / shared memory allow for this variant.
    int per_sm = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, L.smem_bytes) != cudaSuccess ||
        per_sm < 1)
      per_sm = 1;
    cache.grid[dev] = std::min(plan.max_grid, per_sm * sms);
    cache.key[dev] = key;
  }
  *grid_out = cache.grid[dev];
  return FEA_OK;
}

template <int D, int G>
static int launch_tma(const TmaPlan& plan, bool dot, int64_t n_nodes, const int32_t* rp, const int32_t* ci,
                      const double* values, const double* x, double* y, int64_t off, PcgState* st, double* partials,
                      cudaStream_t stream, const PeerLaunch* peer) {
  const TmaLayout& L = plan.layout;
  const int n = (int)n_nodes;
  int grid = 0;
  if (dot && peer != nullptr) {
    static GridCache cache;
    FEA_TRY(tma_grid(cache, pcg_spmv_tma_kernel<D, G, true>, plan, tma_threads(D, G), &grid));
    pcg_spmv_tma_kernel<D, G, true><<<grid, tma_threads(D, G), L.smem_bytes, stream>>>(
        n, rp, ci, values, x, y, x + off * D, L.stages | (plan.l2_hint << kTmaHintShift), L.val_cap, L.col_cap, st, partials, peer->view, peer->key);
  } else if (dot) {
    static GridCache cache;
    FEA_TRY(tma_grid(cache, pcg_spmv_tma_kernel<D, G, false>, plan, tma_threads(D, G), &grid));
    FEA_TRY(check(launch_kernel(pcg_spmv_tma_kernel<D, G, false>, dim3(grid), dim3(tma_threads(D, G)), L.smem_bytes,
                                stream, plan.pdl, n, rp, ci, values, x, y, x + off * D, L.stages | (plan.l2_hint << kTmaHintShift), L.val_cap, L.col_cap,
                                st, partials, (const PeerView*)nullptr, PeerKey{})));
  } else {
    static GridCache cache;
    FEA_TRY(tma_grid(cache, spmv_tma_kernel<D, G>, plan, tma_threads(D, G), &grid));
    spmv_tma_kernel<D, G><<<grid, tma_threads(D, G), L.smem_bytes, stream>>>(n, rp, ci, values, x, y, L.stages | (plan.l2_hint << kTmaHintShift),
                                                                        L.val_cap, L.col_cap);
  }
  return FEA_OK;
}

template <int D>
static int launch_tma_g(const TmaPlan& plan, bool dot, int64_t n_nodes, const int32_t* rp, const int32_t* ci,
                        const double* values, const double* x, double* y, int64_t off, PcgState* st,
                        double* partials, cudaStream_t stream, const PeerLaunch* peer) {
  switch (plan.groups) {
    case 1: return launch_tma<D, 1>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    case 2: return launch_tma<D, 2>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    case 3: return launch_tma<D, 3>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    case 4: return launch_tma<D, 4>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    default: return FEA_ERR_INVALID;
  }
}

static int dispatch_tma(int d, const TmaPlan& plan, bool dot, int64_t n_nodes, const int32_t* rp, const int32_t* ci,
                        const double* values, const double* x, double* y, int64_t off, PcgState* st,
                        double* partials, cudaStream_t stream, const PeerLaunch* peer = nullptr) {
  switch (d) {
    case 1: return launch_tma_g<1>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    case 2: return launch_tma_g<2>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    case 3: return launch_tma_g<3>(plan, dot, n_nodes, rp, ci, values, x, y, off, st, partials, stream, peer);
    default: return FEA_ERR_INVALID;
  }
}

// Stagnation guard (one thread, once per iteration, identical on every rank: rr is the world's sum).
// A singular reduced system with an inconsistent right-hand side -- an under-constrained body -- never
// breaks down in exact terms (p.Ap stays > 0) and would iterate to max_iter = 10 n; its residual stops
// improving instead.  If the best ||r||^2 seen has not dropped by 1 % within `window` iterations the
// solve ends with FEA_ERR_STAGNATION.  window = max(10000, max_iter / 1000): far longer than any plateau
// of a convergent run (config 4 needs 8.5 k iterations in total), 1000x shorter than max_iter.
__device__ __forceinline__ bool stall_improved(const PcgState* st, double rr) {
  return !(rr >= 0.99 * st->spare[1]) || st->spare[1] == 0.0;
}
__device__ __forceinline__ bool stall_check(const PcgState* st, double rr, int iter) {  // read-only
  if (stall_improved(st, rr)) return false;
  return iter - (int)st->spare[2] >= max(10000, st->max_iter / 1000);
}
__device__ __forceinline__ void stall_note(PcgState* st, double rr, int iter) {  // one thread per iteration
  if (stall_improved(st, rr)) {
    st->spare[1] = rr;
    st->spare[2] = (double)iter;
  }
}
__device__ __forceinline__ bool stagnated(PcgState* st, double rr, int iter) {
  const bool stalled = stall_check(st, rr, iter);
  stall_note(st, rr, iter);
  return stalled;
}

// step 2
__global__ void __launch_bounds__(256)
pcg_update_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ p,
                  const double* __restrict__ ap, double* __restrict__ x, double* __restrict__ r, PcgState* st,
                  double* partials, const PeerView* pv, PeerKey key) {
  __shared__ double s_red[32];
  __shared__ double s_glob[2];
  __shared__ int s_ok;
  pdl_launch_dependents();
  pdl_wait();
  if (st->done) return;
  double pap = st->pap;
  const double rz = st->rz;
  const int32_t iter = st->iter;
  if (pv != nullptr) {  // p.Ap of all ranks: every CTA forms the same rank-ordered sum
    if (threadIdx.x < 32) {
      double a, b;
      const bool ok = peer_collect(key, 1, iter, a, b);
      if (threadIdx.x == 0) {
        s_glob[0] = a;
        s_ok = ok;
      }
    }
    __syncthreads();
    if (!s_ok) {
      if (blockIdx.x == 0 && threadIdx.x == 0) peer_failure(key, st);
      return;
    }
    pap = s_glob[0];
  }
  if (st->bnorm2 == 0.0 || !(pap > 0.0)) {  // zero right-hand side, or K_ff not positive definite
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->done = 1;
      st->status = st->bnorm2 == 0.0 ? FEA_OK : FEA_ERR_BREAKDOWN;
      st->rr_final = st->rr;
    }
    return;
  }
  const double alpha = rz / pap;
  double s_rz = 0.0, s_rr = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // 4 independent elements per trip, every load issued before the first use (HBM-bound, 7 streams)
  for (; i + 3 * stride < n; i += 4 * stride) {
    double di[4], pi[4], ai[4], ri[4], xi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = i + u * stride;
      di[u] = dinv[j];
      pi[u] = p[j];
      ai[u] = __ldcs(ap + j);  // ap and x are not needed again before they are rewritten
      ri[u] = r[j];
      xi[u] = __ldcs(x + j);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = i + u * stride;
      const double rn = fma(-alpha, ai[u], ri[u]);
      __stcs(x + j, fma(alpha, pi[u], xi[u]));
      r[j] = rn;
      if (di[u] != 0.0) {
        s_rz = fma(rn * di[u], rn, s_rz);
        s_rr = fma(rn, rn, s_rr);
      }
    }
  }
  for (; i < n; i += stride) {
    const double di = dinv[i];
    const double pi = p[i];
    const double ri = fma(-alpha, ap[i], r[i]);
    x[i] = fma(alpha, pi, x[i]);
    r[i] = ri;
    if (di != 0.0) {
      s_rz = fma(ri * di, ri, s_rz);
      s_rr = fma(ri, ri, s_rr);
    }
  }
  double tot[2];
  tot[0] = block_sum(s_rz, s_red);
  tot[1] = block_sum(s_rr, s_red);
  if (publish_partials(partials, 2, tot, &st->counter[1])) {
    const double a = reduce_partials(partials, s_red);
    const double b = reduce_partials(partials + kMaxPartials, s_red);
    if (threadIdx.x == 0) {
      st->rz_new = a;  // with peers: this rank's partial sums; the direction kernel stores the global ones
      st->rr = b;
      st->iter = iter + 1;
      st->counter[1] = 0;
      s_glob[0] = a;
      s_glob[1] = b;
    }
    if (pv != nullptr) {
      __syncthreads();
      if (threadIdx.x < 32) peer_publish(*pv, 2, iter, s_glob[0], s_glob[1]);
    }
  }
}

// step 3
__global__ void __launch_bounds__(256)
pcg_direction_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                     double* __restrict__ p, PcgState* st, double* history, const PeerView* pv, PeerKey key) {
  __shared__ bool s_last;
  __shared__ double s_glob[2];
  __shared__ int s_ok;
  pdl_launch_dependents();
  pdl_wait();
  if (st->done) return;
  const double rz = st->rz;
  double rz_new = st->rz_new, rr = st->rr;
  const int32_t iter = st->iter;  // already counts the iteration being finished
  if (pv != nullptr) {  // r.z and r.r of all ranks
    if (threadIdx.x < 32) {
      double a, b;
      const bool ok = peer_collect(key, 2, iter - 1, a, b);
      if (threadIdx.x == 0) {
        s_glob[0] = a;
        s_glob[1] = b;
        s_ok = ok;
      }
    }
    __syncthreads();
    if (!s_ok) {
      if (blockIdx.x == 0 && threadIdx.x == 0) peer_failure(key, st);
      return;
    }
    rz_new = s_glob[0];
    rr = s_glob[1];
  }
  const bool converged = rr <= st->tol2 * st->bnorm2;
  if (!converged) {
    const double beta = rz_new / rz;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
      double di[4], ri[4], pi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t j = i + u * stride;
        di[u] = dinv[j];
        ri[u] = r[j];
        pi[u] = p[j];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) p[i + u * stride] = fma(beta, pi[u], di[u] * ri[u]);
    }
    for (; i < n; i += stride) p[i] = fma(beta, p[i], dinv[i] * r[i]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&st->counter[2], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    if (pv != nullptr) {
      st->rz_new = rz_new;
      st->rr = rr;
    }
    if (history != nullptr && iter >= 1 && iter <= st->max_iter) history[iter - 1] = sqrt(rr / st->bnorm2);
    bool finished = false;
    if (converged) {
      st->done = 1;
      st->rr_final = rr;
      finished = true;
    } else {
      st->rz = rz_new;
      const bool stalled = stagnated(st, rr, iter);
      if (iter >= st->max_iter || stalled) {
        st->done = 1;
        st->status = stalled ? FEA_ERR_STAGNATION : FEA_ERR_MAXITER;
        st->rr_final = rr;
        finished = true;
      }
    }
    st->counter[2] = 0;
    (void)finished;  // multi-GPU: the boundary rows of the new p travel in p2p_halo_kernel, on a side stream
  }
}

// ---------------------------------------------------------------------------------------------
// Single-reduction variant (Chronopoulos & Gear 1989), FEA_PCG_ALGO=1: the same iterates as the
// classical recurrence in exact arithmetic, with TWO kernels and ONE dependent reduction per
// iteration instead of three and two:
//   step 1  w = K u            + partial delta = u.w          (the SpMV kernels above; u = dinv r)
//   step 2  beta = gamma/gamma_old, alpha = gamma / (delta - beta gamma / alpha_old)
//           p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = dinv r
//           + partial gamma' = r.u and r.r                              (this kernel)
// Convergence of iteration k is known once its r.r has been reduced, i.e. at the start of step 2 of
// iteration k+1: the solve spends one extra SpMV.  Small meshes are launch-bound (3 kernels cost
// ~33 us per iteration at 133 k DOF, the SpMV itself 17 us), and across GPUs every reduction is a
// cross-rank synchronisation point; at 7.9 M DOF on one GPU the variant is neutral (12 instead of
// 11 vector passes, one launch less).  State: rz = gamma_old, rz_new = gamma, pap = delta,
// spare[0] = alpha_old.
__global__ void __launch_bounds__(256, 3)
pcg_cgcg_kernel(int64_t n, const double* __restrict__ dinv, double* __restrict__ u, const double* __restrict__ w,
                double* __restrict__ p, double* __restrict__ s, double* __restrict__ x, double* __restrict__ r,
                PcgState* st, double* partials, double* history, const PeerView* pv, PeerKey key) {
  __shared__ double s_red[32];
  __shared__ double s_glob[3];
  __shared__ int s_ok;
  pdl_launch_dependents();
  pdl_wait();
  if (st->done) return;
  const int32_t iter = st->iter;  // iterations completed so far
  double gamma = iter == 0 ? st->rz : st->rz_new;
  double rr = st->rr, delta = st->pap;
  if (pv != nullptr) {  // world sums: u.w of this iteration, (r.u, r.r) of the previous one
    if (threadIdx.x < 32) {
      double a, c, d;
      const bool ok = peer_collect2(key, 1, iter, 2, (long long)iter - 1, a, c, d);
      if (threadIdx.x == 0) {
        s_glob[0] = a;
        s_glob[1] = c;
        s_glob[2] = d;
        s_ok = ok;
      }
    }
    __syncthreads();
    if (!s_ok) {
      if (blockIdx.x == 0 && threadIdx.x == 0) peer_failure(key, st);
      return;
    }
    delta = s_glob[0];
    if (iter > 0) {
      gamma = s_glob[1];
      rr = s_glob[2];
    }
  }
  const double bnorm2 = st->bnorm2;
  const bool converged = rr <= st->tol2 * bnorm2;  // also covers a zero right-hand side
  const bool stalled = !converged && iter > 0 && stall_check(st, rr, iter);
  const bool exhausted = !converged && (iter >= st->max_iter || stalled);
  double beta = 0.0, alpha = 0.0;
  bool breakdown = false;
  if (!converged && !exhausted) {
    if (iter == 0) {
      breakdown = !(delta > 0.0);
      alpha = gamma / delta;
    } else {
      beta = gamma / st->rz;
      const double denom = delta - beta * gamma / st->spare[0];
      breakdown = !(denom > 0.0);
      alpha = gamma / denom;
    }
  }
  if (converged || exhausted || breakdown) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      if (history != nullptr && iter >= 1 && iter <= st->max_iter) history[iter - 1] = sqrt(rr / bnorm2);
      st->done = 1;
      st->status = converged ? FEA_OK
                             : (breakdown ? FEA_ERR_BREAKDOWN : (stalled ? FEA_ERR_STAGNATION : FEA_ERR_MAXITER));
      st->rr = rr;
      st->rr_final = rr;
    }
    return;
  }
  double s_ru = 0.0, s_rr = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto finish = [&](int64_t j, double di, double uj, double wj, double pj, double sj, double xj, double rj) {
    const double pn = iter == 0 ? uj : fma(beta, pj, uj);
    const double sn = iter == 0 ? wj : fma(beta, sj, wj);
    const double rn = fma(-alpha, sn, rj);
    const double un = di * rn;
    p[j] = pn;
    s[j] = sn;
    __stcs(x + j, fma(alpha, pn, xj));
    r[j] = rn;
    u[j] = un;
    if (di != 0.0) {
      s_ru = fma(rn, un, s_ru);
      s_rr = fma(rn, rn, s_rr);
    }
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {  // two elements in flight: 14 independent loads
    const int64_t j0 = i, j1 = i + stride;
    const double d0 = dinv[j0], u0 = u[j0], w0 = __ldcs(w + j0), r0 = r[j0], x0 = __ldcs(x + j0);
    const double d1 = dinv[j1], u1 = u[j1], w1 = __ldcs(w + j1), r1 = r[j1], x1 = __ldcs(x + j1);
    double p0 = 0.0, q0 = 0.0, p1 = 0.0, q1 = 0.0;
    if (iter != 0) {
      p0 = p[j0];
      q0 = s[j0];
      p1 = p[j1];
      q1 = s[j1];
    }
    finish(j0, d0, u0, w0, p0, q0, x0, r0);
    finish(j1, d1, u1, w1, p1, q1, x1, r1);
  }
  for (; i < n; i += stride)
    finish(i, dinv[i], u[i], w[i], iter != 0 ? p[i] : 0.0, iter != 0 ? s[i] : 0.0, x[i], r[i]);
  double tot[2];
  tot[0] = block_sum(s_ru, s_red);
  tot[1] = block_sum(s_rr, s_red);
  if (publish_partials(partials, 2, tot, &st->counter[1])) {
    const double a = reduce_partials(partials, s_red);
    const double b = reduce_partials(partials + kMaxPartials, s_red);
    if (threadIdx.x == 0) {
      if (history != nullptr && iter >= 1 && iter <= st->max_iter) history[iter - 1] = sqrt(rr / bnorm2);
      if (iter > 0) stall_note(st, rr, iter);
      st->rz = gamma;    // gamma_old of the next iteration
      st->rz_new = a;    // with peers: this rank's partial sums (the next step 2 collects the world's)
      st->rr = b;
      st->spare[0] = alpha;
      st->iter = iter + 1;
      st->counter[1] = 0;
      s_glob[0] = a;
      s_glob[1] = b;
    }
    if (pv != nullptr) {
      __syncthreads();
      if (threadIdx.x < 32) peer_publish(*pv, 2, iter, s_glob[0], s_glob[1]);
    }
  }
}

// x = 0, r = b on free DOF (0 on constrained), p = dinv r; partial r.z and ||b||^2.
__global__ void __launch_bounds__(256)
pcg_init_kernel(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ p, double tol, int max_iter, PcgState* st,
                double* partials) {
  __shared__ double s_red[32];
  double s_rz = 0.0, s_bb = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double di = dinv[i];
    const double ri = di != 0.0 ? b[i] : 0.0;
    const double zi = di * ri;
    x[i] = 0.0;
    r[i] = ri;
    p[i] = zi;
    s_rz = fma(ri, zi, s_rz);
    s_bb = fma(ri, ri, s_bb);
  }
  double tot[2];
  tot[0] = block_sum(s_rz, s_red);
  tot[1] = block_sum(s_bb, s_red);
  if (publish_partials(partials, 2, tot, &st->counter[3])) {
    const double a = reduce_partials(partials, s_red);
    const double c = reduce_partials(partials + kMaxPartials, s_red);
    if (threadIdx.x == 0) {
      st->rz = a;
      st->bnorm2 = c;
      st->rz_new = 0.0;
      st->rr = c;
      st->pap = 0.0;
      st->tol2 = tol * tol;
      st->iter = 0;
      st->done = 0;
      st->status = FEA_OK;
      st->max_iter = max_iter;
      st->counter[0] = st->counter[1] = st->counter[2] = st->counter[3] = 0;
    }
  }
}

inline unsigned spmv_blocks(int64_t n_nodes) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_nodes, kSpmvWarps), 148LL * 8));
}

struct PcgWork {
  PcgState* state;
  double* partials;  // 2 * kMaxPartials
  double* r;
  double* p;   // classical: search direction; single-reduction variant: u = dinv r (the SpMV input)
  double* ap;  // SpMV output
  double* p2;  // single-reduction variant only: search direction p and s = K p
  double* s;
};

static PcgWork carve_pcg(void* work, int64_t n) {
  PcgWork w;
  char* c = static_cast<char*>(work);
  w.state = reinterpret_cast<PcgState*>(c);
  c += FEA_PCG_STATE_BYTES;
  w.partials = reinterpret_cast<double*>(c);
  c += sizeof(double) * 2 * kMaxPartials;
  const size_t vec = align_up(sizeof(double) * (size_t)n, 256);
  w.r = reinterpret_cast<double*>(c);
  c += vec;
  w.p = reinterpret_cast<double*>(c);
  c += vec;
  w.ap = reinterpret_cast<double*>(c);
  c += vec;
  w.p2 = reinterpret_cast<double*>(c);
  c += vec;
  w.s = reinterpret_cast<double*>(c);
  return w;
}

int pcg_algorithm(int64_t n_dof, bool multi_gpu) {  // read per solve, so that a process can switch (tests do)
  const char* env = std::getenv("FEA_PCG_ALGO");
  if (env != nullptr && (env[0] == '0' || env[0] == '1')) return env[0] - '0';
  // The single-reduction variant trades one kernel boundary / one cross-rank synchronisation per
  // iteration for one more vector pass (8 n bytes); n is the per-rank size.  Measured on B200
  // (400x80x80 unless noted): one GPU 133 k DOF -12 %, one GPU 7.9 M DOF +2 %; two GPUs (3.9 M DOF per
  // rank) +0.5 %; eight GPUs (1 M DOF per rank) -8.9 % (1.417 -> 1.291 s).  The vector kernel must run
  // as ONE resident wave: with 4 waves of CTAs each polling the peer slots it lost 19 % on 8 GPUs.
  (void)multi_gpu;
  return n_dof < kSingleReductionBelowDof ? 1 : 0;
}

template <int D>
static int launch_step_spmv(int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                            const double* p, double* ap, int64_t off, PcgState* st, double* partials,
                            cudaStream_t stream, const PeerLaunch* peer) {
  pcg_spmv_kernel<D><<<spmv_blocks(n_nodes), kSpmvThreads, 0, stream>>>(
      n_nodes, rp, ci, values, p, ap, off, st, partials, peer ? peer->view : nullptr, peer ? peer->key : PeerKey{});
  return FEA_OK;
}

int pcg_step_spmv(int d, int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                  const double* p, double* ap, int64_t off, PcgState* st, double* partials, cudaStream_t stream,
                  const TmaPlan* plan, const PeerLaunch* peer) {
  if (plan != nullptr && plan->ok)
    return dispatch_tma(d, *plan, true, n_nodes, rp, ci, values, p, ap, off, st, partials, stream, peer);
  switch (d) {
    case 1: return launch_step_spmv<1>(n_nodes, rp, ci, values, p, ap, off, st, partials, stream, peer);
    case 2: return launch_step_spmv<2>(n_nodes, rp, ci, values, p, ap, off, st, partials, stream, peer);
    case 3: return launch_step_spmv<3>(n_nodes, rp, ci, values, p, ap, off, st, partials, stream, peer);
    default: return FEA_ERR_INVALID;
  }
}

}  // namespace fea

using namespace fea;

extern "C" int fea_spmv(int64_t n_nodes, int32_t d, const int32_t* node_rowptr, const int32_t* node_colidx,
                        const double* values, int32_t max_coupled, const double* x, double* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !values || !x || !y || n_nodes <= 0) return FEA_ERR_INVALID;
  if (d < 1 || d > 3) return FEA_ERR_INVALID;
  const TmaPlan plan = tma_plan(d, max_coupled, values, node_colidx, n_nodes);
  if (plan.ok) {
    FEA_TRY(dispatch_tma(d, plan, false, n_nodes, node_rowptr, node_colidx, values, x, y, 0, nullptr, nullptr,
                         stream));
    return check_launch();
  }
  const unsigned blocks = spmv_blocks(n_nodes);
  switch (d) {
    case 1: spmv_kernel<1><<<blocks, kSpmvThreads, 0, stream>>>(n_nodes, node_rowptr, node_colidx, values, x, y); break;
    case 2: spmv_kernel<2><<<blocks, kSpmvThreads, 0, stream>>>(n_nodes, node_rowptr, node_colidx, values, x, y); break;
    case 3: spmv_kernel<3><<<blocks, kSpmvThreads, 0, stream>>>(n_nodes, node_rowptr, node_colidx, values, x, y); break;
    default: return FEA_ERR_INVALID;
  }
  return check_launch();
}

extern "C" size_t fea_pcg_workspace(int64_t n_dof) {
  return FEA_PCG_STATE_BYTES + sizeof(double) * 2 * kMaxPartials + 5 * align_up(sizeof(double) * (size_t)n_dof, 256);
}

// The SpMV kernel wants (almost) the whole SM array as shared memory; the vector kernels of the
// same iteration would default to a large L1.  Alternating carve-outs makes every launch wait for
// an SM reconfiguration (~25 us per launch measured at 100x20x20, where a whole iteration is
// otherwise ~35 us), so the vector kernels ask for the SpMV's carve-out: they stream and do not
// need L1.
void fea::pcg_match_carveout() {
  static bool done[kMaxDevices] = {};
  static std::mutex mu;
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  if (done[dev]) return;
  done[dev] = true;
  cudaFuncSetAttribute(pcg_update_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(pcg_direction_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(pcg_init_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(pcg_cgcg_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

extern "C" int fea_pcg_init(int64_t n_dof, const double* b, const double* dinv, double* x, double* r, double* p,
                            double tol, int32_t max_iter, void* state, void* partials, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!b || !dinv || !x || !r || !p || !state || !partials || n_dof <= 0) return FEA_ERR_INVALID;
  pcg_match_carveout();
  FEA_TRY(check(cudaMemsetAsync(state, 0, FEA_PCG_STATE_BYTES, stream)));
  pcg_init_kernel<<<vec_blocks(n_dof), 256, 0, stream>>>(n_dof, b, dinv, x, r, p, tol, max_iter,
                                                        static_cast<PcgState*>(state), static_cast<double*>(partials));
  return check_launch();
}

extern "C" int fea_pcg_step_spmv(int64_t n_owned_nodes, int32_t d, const int32_t* node_rowptr,
                                 const int32_t* node_colidx, const double* values, int32_t max_coupled,
                                 const double* p, double* ap, int64_t p_row_offset, void* state, void* partials,
                                 void* stream_) {
  if (!node_rowptr || !node_colidx || !values || !p || !ap || !state || !partials || n_owned_nodes <= 0)
    return FEA_ERR_INVALID;
  const TmaPlan plan = tma_plan(d, max_coupled, values, node_colidx, n_owned_nodes);
  FEA_TRY(pcg_step_spmv(d, n_owned_nodes, node_rowptr, node_colidx, values, p, ap, p_row_offset,
                    static_cast<PcgState*>(state), static_cast<double*>(partials),
                    static_cast<cudaStream_t>(stream_), &plan));
  return check_launch();
}

extern "C" int fea_pcg_step_update(int64_t n_dof, const double* dinv, const double* p, const double* ap, double* x,
                                   double* r, void* state, void* partials, void* stream_) {
  if (!dinv || !p || !ap || !x || !r || !state || !partials || n_dof <= 0) return FEA_ERR_INVALID;
  pcg_update_kernel<<<vec_blocks(n_dof), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      n_dof, dinv, p, ap, x, r, static_cast<PcgState*>(state), static_cast<double*>(partials), nullptr, PeerKey{});
  return check_launch();
}

extern "C" int fea_pcg_step_direction(int64_t n_dof, const double* dinv, const double* r, double* p, void* state,
                                      double* history, void* stream_) {
  if (!dinv || !r || !p || !state || n_dof <= 0) return FEA_ERR_INVALID;
  pcg_direction_kernel<<<vec_blocks(n_dof), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      n_dof, dinv, r, p, static_cast<PcgState*>(state), history, nullptr, PeerKey{});
  return check_launch();
}

// One launch of the persistent solver (pcg_fused.cuh): `args.iters` iterations of the single-reduction recurrence.
template <int D, int G>
static int launch_fused(const TmaPlan& plan, FusedArgs args, cudaStream_t stream) {
  static GridCache cache;
  int grid = 0;
  FEA_TRY(tma_grid(cache, pcg_fused_kernel<D, G>, plan, tma_threads(D, G), &grid));
  if (5 * grid > 2 * kMaxPartials) return FEA_ERR_INVALID;
  FEA_TRY(check(cudaMemsetAsync(&args.st->counter[3], 0, sizeof(uint32_t), stream)));  // the grid barrier's ticket
  void* kargs[] = {&args};
  return check(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(pcg_fused_kernel<D, G>), dim3(grid),
                                           dim3(tma_threads(D, G)), kargs, plan.layout.smem_bytes, stream));
}
template <int D>
static int launch_fused_g(const TmaPlan& plan, const FusedArgs& args, cudaStream_t stream) {
  switch (plan.groups) {
    case 1: return launch_fused<D, 1>(plan, args, stream);
    case 2: return launch_fused<D, 2>(plan, args, stream);
    case 3: return launch_fused<D, 3>(plan, args, stream);
    case 4: return launch_fused<D, 4>(plan, args, stream);
    default: return FEA_ERR_INVALID;
  }
}
static int dispatch_fused(int d, const TmaPlan& plan, const FusedArgs& args, cudaStream_t stream) {
  switch (d) {
    case 1: return launch_fused_g<1>(plan, args, stream);
    case 2: return launch_fused_g<2>(plan, args, stream);
    case 3: return launch_fused_g<3>(plan, args, stream);
    default: return FEA_ERR_INVALID;
  }
}
// The persistent kernel needs every CTA resident at once (grid barrier): cooperative launch.
static bool fused_supported() {
  int dev = 0, coop = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) return false;
  const char* env = std::getenv("FEA_PCG_FUSED");
  return coop != 0 && !(env != nullptr && env[0] == '0');
}

extern "C" int fea_pcg_solve(int64_t n_nodes, int32_t d, const int32_t* node_rowptr, const int32_t* node_colidx,
                             const double* values, int32_t max_coupled, const double* dinv, const double* b,
                             double* x, double tol, int32_t max_iter, void* work, size_t work_bytes,
                             double* history, fea_pcg_result* result_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !values || !dinv || !b || !x || !work || !result_host) return FEA_ERR_INVALID;
  if (n_nodes <= 0 || d < 1 || d > 3 || max_iter < 1) return FEA_ERR_INVALID;
  const int64_t n = n_nodes * d;
  if (work_bytes < fea_pcg_workspace(n)) return FEA_ERR_WORKSPACE;
  PcgWork w = carve_pcg(work, n);
  TmaPlan plan = tma_plan(d, max_coupled, values, node_colidx, n_nodes, /*small_problem=*/n < kSingleReductionBelowDof);
  // programmatic dependent launch between the kernels of an iteration (common.cuh); FEA_PCG_PDL=0 switches it off
  const char* pdl_env = std::getenv("FEA_PCG_PDL");
  const bool pdl = pdl_env == nullptr || pdl_env[0] != '0';
  plan.pdl = pdl;

  // two pinned snapshots of the state, polled one chunk behind the GPU
  PcgState* snap = static_cast<PcgState*>(pinned_scratch(0, 2 * sizeof(PcgState)));
  if (snap == nullptr) return FEA_ERR_CUDA;
  cudaEvent_t ev[2];
  int rc = FEA_OK;
  if ((rc = check(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming))) != FEA_OK) return rc;
  if ((rc = check(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming))) != FEA_OK) {
    cudaEventDestroy(ev[0]);
    return rc;
  }

  // The solve runs on a private stream ordered after `stream` (and `stream` is ordered after it
  // at the end): the caller's stream may be the legacy default stream, which cannot be captured.
  cudaStream_t caller = stream;
  cudaStream_t own = nullptr;
  cudaEvent_t ev_order = nullptr;
  if (std::getenv("FEA_PCG_NO_GRAPH") == nullptr &&
      cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) == cudaSuccess &&
      cudaEventCreateWithFlags(&ev_order, cudaEventDisableTiming) == cudaSuccess &&
      cudaEventRecord(ev_order, caller) == cudaSuccess && cudaStreamWaitEvent(own, ev_order, 0) == cudaSuccess) {
    stream = own;
  } else if (own != nullptr) {
    cudaStreamDestroy(own);
    own = nullptr;
    cudaGetLastError();
  }

  rc = fea_pcg_init(n, b, dinv, x, w.r, w.p, tol, max_iter, w.state, w.partials, stream);
  const unsigned vb = vec_blocks(n);
  const int chunk = 32;
  const int algo = pcg_algorithm(n, false);
  // L2-sized problems on the single-reduction recurrence: one persistent kernel per chunk of iterations
  bool fused = algo == 1 && plan.ok && n_nodes < (int64_t)INT32_MAX / 4 && fused_supported();
  FusedArgs fargs{};
  if (fused) {
    fargs.n_nodes = (int)n_nodes;
    fargs.node_rowptr = node_rowptr;
    fargs.node_colidx = node_colidx;
    fargs.values = values;
    fargs.dinv = dinv;
    fargs.u = w.p;
    fargs.w = w.ap;
    fargs.p = w.p2;
    fargs.s = w.s;
    fargs.x = x;
    fargs.r = w.r;
    fargs.stages_arg = plan.layout.stages | (plan.l2_hint << kTmaHintShift);
    fargs.val_cap = plan.layout.val_cap;
    fargs.col_cap = plan.layout.col_cap;
    fargs.st = w.state;
    fargs.partials = w.partials;
    fargs.history = history;
  }
  const int per_iter = algo == 1 ? 2 : 3;  // kernels per iteration
  // the single-reduction variant learns about convergence / max_iter one step later
  const int64_t enqueue_limit = (int64_t)max_iter + (algo == 1 ? 1 : 0);
  int64_t enqueued = 0;
  int slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  // measurement hook: CUDA-event pairs around the first SpMV launch of each chunk
  constexpr int kMaxSamples = 256;
  cudaEvent_t* sample_ev = nullptr;
  int n_samples = 0;
  if (profile().enabled) {
    sample_ev = new cudaEvent_t[2 * kMaxSamples];
    for (int i = 0; i < 2 * kMaxSamples; ++i) cudaEventCreate(&sample_ev[i]);
  }
  int sample_iter[kMaxSamples];
  auto enqueue_iteration = [&](bool sample) -> int {
    if (sample) sample_iter[n_samples] = (int)enqueued;  // 0-based index of the iteration being timed
    if (sample) cudaEventRecord(sample_ev[2 * n_samples], stream);
    const int r = pcg_step_spmv(d, n_nodes, node_rowptr, node_colidx, values, w.p, w.ap, 0, w.state, w.partials, stream,
                            &plan);
    if (sample) cudaEventRecord(sample_ev[2 * n_samples++ + 1], stream);
    const PeerView* no_peer = nullptr;
    if (algo == 1) {
      launch_kernel(pcg_cgcg_kernel, dim3(cgcg_blocks(n)), dim3(256), 0, stream, pdl, n, dinv, w.p, w.ap, w.p2, w.s, x, w.r,
                    w.state, w.partials, history, no_peer, PeerKey{});
    } else {
      launch_kernel(pcg_update_kernel, dim3(vb), dim3(256), 0, stream, pdl, n, dinv, w.p, w.ap, x, w.r, w.state,
                    w.partials, no_peer, PeerKey{});
      launch_kernel(pcg_direction_kernel, dim3(vb), dim3(256), 0, stream, pdl, n, dinv, w.r, w.p, w.state, history,
                    no_peer, PeerKey{});
    }
    return r;
  };
  // One iteration is launched directly (it carries the timing sample), the other chunk-1 are one
  // CUDA-graph launch: the kernel arguments never change, and small problems (100x20x20: ~35 us
  // of kernels per iteration) are otherwise bound by the host's launch rate.
  cudaGraphExec_t graph_exec = nullptr;
  if (own != nullptr && rc == FEA_OK && max_iter >= chunk && !fused) {
    enqueue_iteration(false);  // warm: every cudaFuncSetAttribute / occupancy query happens outside capture
    rc = check_launch(per_iter);
    enqueued += 1;
    cudaGraph_t graph = nullptr;
    if (rc == FEA_OK && cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      for (int it = 0; it < chunk - 1; ++it) enqueue_iteration(false);
      if (cudaStreamEndCapture(stream, &graph) != cudaSuccess || graph == nullptr ||
          cudaGraphInstantiate(&graph_exec, graph, 0) != cudaSuccess)
        graph_exec = nullptr;
      if (graph != nullptr) cudaGraphDestroy(graph);
    }
    cudaGetLastError();  // a failed capture falls back to plain launches
  }
  while (rc == FEA_OK && !finished) {
    const int todo = (int)std::min<int64_t>(fused ? 4 * chunk : chunk, enqueue_limit - enqueued);
    if (fused) {
      fargs.iters = todo;
      rc = dispatch_fused(d, plan, fargs, stream);
      if (rc != FEA_OK && enqueued == 0) {
        // the cooperative launch was refused (e.g. the device is partitioned and cannot hold the whole grid):
        // nothing has run yet, take the two-kernel path for this solve
        cudaGetLastError();
        rc = FEA_OK;
        fused = false;
        continue;
      }
      if (rc == FEA_OK) profile().launches += 1;
    } else if (graph_exec != nullptr && todo == chunk) {
      rc = enqueue_iteration(sample_ev != nullptr && n_samples < kMaxSamples);
      if (rc == FEA_OK) rc = check(cudaGraphLaunch(graph_exec, stream));
    } else {
      for (int it = 0; it < todo && rc == FEA_OK; ++it)
        rc = enqueue_iteration(sample_ev != nullptr && it == 0 && n_samples < kMaxSamples);
    }
    if (rc == FEA_OK && !fused) rc = check_launch(per_iter * todo);
    if (rc != FEA_OK) break;
    enqueued += todo;
    rc = check(cudaMemcpyAsync(&snap[slot], w.state, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
    if (rc != FEA_OK) break;
    rc = check(cudaEventRecord(ev[slot], stream));
    if (rc != FEA_OK) break;
    pending[slot] = true;
    const int prev = slot ^ 1;
    if (pending[prev]) {  // look at the chunk before the one just enqueued
      rc = check(cudaEventSynchronize(ev[prev]));
      pending[prev] = false;
      if (rc == FEA_OK && snap[prev].done) finished = true;
    }
    if (!finished && enqueued >= enqueue_limit) {
      rc = check(cudaEventSynchronize(ev[slot]));
      pending[slot] = false;
      finished = true;
    }
    slot ^= 1;
  }
  if (rc == FEA_OK) {
    rc = check(cudaMemcpyAsync(&snap[0], w.state, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
    if (rc == FEA_OK) rc = check(cudaStreamSynchronize(stream));
  }
  if (rc == FEA_OK) {
    const PcgState& s = snap[0];
    result_host->iterations = s.iter;
    result_host->status = s.status;
    if (!s.done && s.status == FEA_OK) result_host->status = FEA_ERR_MAXITER;
    result_host->bnorm = std::sqrt(s.bnorm2);
    result_host->rel_residual = s.bnorm2 > 0.0 ? std::sqrt(s.rr / s.bnorm2) : 0.0;
    profile().pcg_iterations += s.iter;
    for (int i = 0; i < n_samples; ++i) {
      if (sample_iter[i] >= s.iter) break;  // launches after convergence are no-ops: not SpMV work
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sample_ev[2 * i], sample_ev[2 * i + 1]) == cudaSuccess) {
        profile().add_spmv_sample(ms);
      }
    }
  } else {
    cudaStreamSynchronize(stream);
  }
  if (sample_ev != nullptr) {
    for (int i = 0; i < 2 * kMaxSamples; ++i) cudaEventDestroy(sample_ev[i]);
    delete[] sample_ev;
  }
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  if (graph_exec != nullptr) cudaGraphExecDestroy(graph_exec);
  if (own != nullptr) {
    // the solve has been synchronised above; order the caller's stream after it anyway
    if (ev_order != nullptr && cudaEventRecord(ev_order, own) == cudaSuccess) cudaStreamWaitEvent(caller, ev_order, 0);
    cudaStreamDestroy(own);
  }
  if (ev_order != nullptr) cudaEventDestroy(ev_order);
  return rc;
}
