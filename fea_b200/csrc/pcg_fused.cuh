// Persistent single-kernel PCG for problems whose matrix is (nearly) L2-resident: BASELINE config 3.
//
// Two kernels per iteration cost such a solve more than the arithmetic: at 133 k DOF the SpMV alone takes
// 13.6 us inside a CUDA graph, the iteration 26-28 us -- two kernel boundaries, two last-block reductions (fence,
// ticket, one block re-reading every partial), the ramp of the bulk-copy rings at every launch
// (profiles/small_mesh_sweep_r02.txt).  Here ONE cooperative launch runs a chunk of iterations of the
// single-reduction recurrence (Chronopoulos & Gear; the same iterates as pcg_cgcg_kernel, pcg.cu):
//
//   producer warp    streams the CTA's tiles of K through the mbarrier ring, iteration after iteration; the ring
//                    never drains, so the first tiles of iteration k+1 arrive during the vector phase of k
//   consumer warps   phase 1  w = K u on the CTA's tiles (spmv_tma.cuh's (row, b)-per-lane scheme; u is gathered with
//                             ordinary loads -- other SMs rewrite it every iteration, the read-only path would be stale)
//                             + partial delta = u.w                                   -> grid barrier
//                    phase 2  every CTA adds the partials of delta, gamma = r.u, r.r in the same fixed order
//                             (identical alpha, beta and convergence decision everywhere, no last block),
//                             updates p, s, x, r, u on ITS slice of the vectors
//                             + partial gamma', r.r                                   -> grid barrier
//
// The grid barrier is a ticket counter in the solver state polled by one thread per CTA (bounded: a CTA that
// waits ~2 s gives up and the solve ends with FEA_ERR_CUDA instead of hanging the device).  The state block
// (PcgState) has the same meaning as in the two-kernel path at every launch boundary, so the host loop, the
// snapshots and the result are shared.  (Exchanging the partials as tagged words polled by every CTA -- barrier and
// reduction in one, as between GPUs -- was measured too: 27.6 us per iteration against 21.2, 87 k pollers on L2.)  When the solve ends inside a chunk the consumers leave at once; the producer
// sees their flag, waits for the copies it has already issued to land, and leaves too.
#pragma once
#include "pcg_common.cuh"

namespace fea {

constexpr int kFusedBarrierId = 15;  // named barrier of all consumer threads of a CTA (1 .. G are the groups')

struct FusedArgs {  // by value; scalars and pointers only
  int n_nodes;
  const int32_t* node_rowptr;
  const int32_t* node_colidx;
  const double* values;
  const double* dinv;
  double* u;  // dinv r: the SpMV's input
  double* w;  // K u
  double* p;
  double* s;
  double* x;
  double* r;
  int stages_arg, val_cap, col_cap;
  PcgState* st;
  double* partials;  // 5 * gridDim.x doubles: delta | gamma'[0] | rr'[0] | gamma'[1] | rr'[1]
  double* history;
  int iters;  // iterations of this launch
};

// Sum over the consumer threads of a CTA, returned to every one of them; fixed order.
template <int NW, int NV>
__device__ __forceinline__ void consumer_sum(double (&v)[NV], double* scratch, int cwarp, int lane, int nthreads) {
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  group_barrier(kFusedBarrierId, nthreads);  // scratch free again
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[k * NW + cwarp] = v[k];
  }
  group_barrier(kFusedBarrierId, nthreads);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += scratch[k * NW + i];
    v[k] = t;
  }
}

// row_part_dot (spmv_tma.cuh) with ORDINARY loads of x: cached in L1 like the read-only path, but -- unlike it --
// covered by the grid barrier's acquire, which invalidates the SM's L1 (the cooperative-groups grid.sync contract):
// x is rewritten by other SMs between two sweeps of the same kernel.  Gathering through L2 instead (ld.global.cg)
// doubles the L2 traffic of an L2-resident SpMV: measured 39 against 26 us per iteration at 133 k DOF.
template <int D>
__device__ __forceinline__ double row_part_dot_plain(const double* vrow, const int32_t* cols, int cnt, int b,
                                                     const double* x) {
  double acc = 0.0;
  const double* xb = x + b;
  const double* vb = vrow + b;
  for (int k0 = 0; k0 < cnt; k0 += kTmaUnroll) {
    double xv[kTmaUnroll];
#pragma unroll
    for (int u = 0; u < kTmaUnroll; ++u) {
      const int k = k0 + u;
      xv[u] = k < cnt ? xb[(int64_t)D * cols[k]] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kTmaUnroll; ++u) {
      const int k = k0 + u;
      if (k < cnt) acc = fma(vb[D * k], xv[u], acc);
    }
  }
  return acc;
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// All consumer threads of all CTAs.  Returns false when the other CTAs never arrived.
// One thread per CTA: release fence, ticket, RELAXED polls, one acquire fence at the end (MEMBAR.ALL.GPU +
// CCTL.IVALL: the invalidation of this SM's L1 that the ordinary loads of u and w rely on).  An acquire load per poll
// (LDG.STRONG.GPU + CCTL.IVALL each time) would keep wiping the L1 under the other CTA of the SM while it still sweeps.
__device__ __forceinline__ bool fused_grid_barrier(unsigned* counter, unsigned target, int ctid, int nthreads,
                                                   int* s_fail) {
  group_barrier(kFusedBarrierId, nthreads);
  if (ctid == 0) {
    fence_acq_rel_gpu();
    atomicAdd(counter, 1u);
    int polls = 0;
    while (ld_relaxed_gpu_u32(counter) < target) {
      if (++polls > (1 << 24)) {
        *s_fail = 1;
        break;
      }
    }
    fence_acq_rel_gpu();
  }
  group_barrier(kFusedBarrierId, nthreads);
  return *s_fail == 0;
}

// Phase 1 of one iteration for a consumer thread: the CTA's tiles of w = K u (ring tiles it * Qr + q, then the direct
// ones), returns this thread's part of u.w.  NOT inlined: the gather loop wants the whole register budget to itself
// (inlined into the persistent kernel it spilled inside the loop: 36 against 26 us per iteration at 133 k DOF); what
// the kernel keeps alive across the call is saved once per iteration instead.
template <int D, int G>
__device__ __noinline__ double fused_sweep(int n_nodes, const int32_t* __restrict__ node_rowptr,
                                           const int32_t* __restrict__ node_colidx, const double* __restrict__ values,
                                           const double* u, double* w, unsigned char* smem, int stages_arg, int val_cap,
                                           int col_cap, int it, int Q, int Qr) {
  constexpr int DD = D * D;
  constexpr int ROWS = D * kTileNodes;
  constexpr int ITEMS = tma_items(D);
  constexpr int GW = tma_group_warps(D);
  const int stages = stages_arg & ((1 << kTmaHintShift) - 1);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kTmaMaxStages;
  double* parts = reinterpret_cast<double*>(smem + kTmaBarrierBytes);  // [2][G][ITEMS]
  unsigned char* stage0 = smem + tma_fixed_bytes(D, G);
  const int stage_bytes = (int)(sizeof(double) * val_cap + sizeof(int32_t) * col_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = (int)gridDim.x;
  const int total_cols = node_rowptr[n_nodes];
  const int group = warp / GW;
  const int item = (warp - group * GW) * 32 + lane;  // b-major: item = b * ROWS + row
  const bool has_item = item < ITEMS;
  const int b = has_item ? item / ROWS : 0;
  const int row = has_item ? item - b * ROWS : 0;
  const int node_in_tile = row / D, arow = row - node_in_tile * D;
  double* parts_g = parts + group * ITEMS;
  int buf_sel = 0;
  // row pointers of this lane's node in local tile q
  auto my_rows = [&](int q, int& r0, int& r1, int& lo, int& hi) {
    r0 = r1 = lo = hi = 0;
    if (q < Q) {
      const int n0 = ((int)blockIdx.x + stride * q) * kTileNodes;
      const int n1 = n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes;
      r0 = node_rowptr[n0];
      r1 = node_rowptr[n1];
      if (n0 + node_in_tile < n1) {
        lo = node_rowptr[n0 + node_in_tile];
        hi = node_rowptr[n0 + node_in_tile + 1];
      }
    }
  };
  int r0, r1, lo, hi;
  my_rows(group, r0, r1, lo, hi);
  double dot = 0.0;
  for (int q = group; q < Q; q += G) {
    const int n0 = ((int)blockIdx.x + stride * q) * kTileNodes;
    const int n1 = n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes;
    const int node = n0 + node_in_tile;
    const bool active = has_item && node < n1;
    int nr0, nr1, nlo, nhi;  // next tile of this group, in flight during this one
    my_rows(q + G, nr0, nr1, nlo, nhi);
    const TileRange t = tile_range<D>(r0, r1, total_cols);
    const int cnt = hi - lo;
    double part = 0.0;
    if (q >= Qr) {
      if (active) {
        const double* vg = values + (int64_t)DD * lo + arow * D * cnt;
        part = row_part_dot_plain<D>(vg, node_colidx + lo, cnt, b, u);
      }
    } else {
      const long long rq = (long long)it * Qr + q;
      const int s = (int)(rq % stages);
      const uint32_t ph = (uint32_t)(rq / stages) & 1u;
      mbar_wait(&full[s], ph);
      if (active) {
        const unsigned char* buf = stage0 + (size_t)s * stage_bytes;
        const double* vs = reinterpret_cast<const double*>(buf) + (int)((int64_t)DD * lo - t.v_lo);
        const int32_t* cs = reinterpret_cast<const int32_t*>(buf + sizeof(double) * val_cap) + (lo - t.c_lo);
        FEA_ASSERT((int64_t)DD * lo >= t.v_lo && (int64_t)DD * hi <= t.v_hi && lo >= t.c_lo && hi <= t.c_hi);
        part = row_part_dot_plain<D>(vs + arow * D * cnt, cs, cnt, b, u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    double* my_parts = parts_g + buf_sel * (G * ITEMS);
    if (has_item) my_parts[item] = part;
    group_barrier(1 + group, GW * 32);
    if (active && b == 0) {  // items 0 .. ROWS-1 finish their row
      double out = part;
#pragma unroll
      for (int bb = 1; bb < D; ++bb) out += my_parts[bb * ROWS + row];
      const int64_t j = (int64_t)node * D + arow;
      w[j] = out;
      dot = fma(out, u[j], dot);
    }
    buf_sel ^= 1;
    r0 = nr0;
    r1 = nr1;
    lo = nlo;
    hi = nhi;
  }
  return dot;
}

template <int D, int G>
__global__ void __launch_bounds__(tma_threads(D, G), tma_min_blocks(D, G)) pcg_fused_kernel(FusedArgs a) {
  constexpr int GW = tma_group_warps(D);
  constexpr int NW = G * GW;     // consumer warps
  constexpr int CT = NW * 32;    // consumer threads
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double s_scratch[3 * NW];
  __shared__ int s_fail;
  __shared__ int s_stop;  // consumers -> producer: the solve has ended, issue nothing more
  __shared__ double s_sc[8];  // [6] ||b||^2  [7] tol^2: constants of the solve
  __shared__ int s_iter;      // iterations completed so far
  __shared__ int s_max_iter;
  const int stages = a.stages_arg & ((1 << kTmaHintShift) - 1), l2_hint = a.stages_arg >> kTmaHintShift;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kTmaMaxStages;
  unsigned char* stage0 = smem + tma_fixed_bytes(D, G);
  const int val_cap = a.val_cap, col_cap = a.col_cap;
  const int stage_bytes = (int)(sizeof(double) * val_cap + sizeof(int32_t) * col_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PcgState* st = a.st;
  if (st->done) return;  // uniform: nothing has been issued yet

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GW);
    }
    mbar_fence_init();
    s_fail = 0;
    s_stop = 0;
  }
  __syncthreads();

  const int n_nodes = a.n_nodes;
  const int32_t* __restrict__ node_rowptr = a.node_rowptr;
  const int32_t* __restrict__ node_colidx = a.node_colidx;
  const double* __restrict__ values = a.values;
  const int n_tiles = (n_nodes + kTileNodes - 1) / kTileNodes;
  const int total_cols = node_rowptr[n_nodes];
  const int stride = (int)gridDim.x;
  // Local tiles q = 0 .. Q-1 are the matrix tiles blockIdx + stride q.  Tiles whose 16-byte rounded byte range
  // would run past the arrays ("direct": the very last tile(s) of the matrix) are read with plain loads and do not
  // take a ring slot; they are a suffix of the local sequence, so ring tile number it * Qr + q is known to both
  // sides without communication.
  const int Q = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + stride - 1) / stride : 0;
  auto tile_rows = [&](int q, int& r0, int& r1) {
    const int n0 = ((int)blockIdx.x + stride * q) * kTileNodes;
    r0 = node_rowptr[n0];
    r1 = node_rowptr[n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes];
  };
  int Qr = Q;
  while (Qr > 0) {
    int r0, r1;
    tile_rows(Qr - 1, r0, r1);
    if (!tile_range<D>(r0, r1, total_cols).direct) break;
    --Qr;
  }

  if (warp == NW) {
    // ------------------------------------------------------------------ producer
    const uint64_t policy = l2_hint == 0 ? l2_evict_first_policy()
                                         : (l2_hint == 1 ? l2_evict_normal_policy() : l2_evict_last_policy());
    long long issued = 0;  // ring tiles issued so far (lane 0)
    bool stop = false;
    for (int it = 0; it < a.iters && !stop; ++it) {
      for (int q0 = 0; q0 < Qr && !stop; q0 += 32) {
        int my_r0 = 0, my_r1 = 0;
        if (q0 + lane < Qr) tile_rows(q0 + lane, my_r0, my_r1);
        const int nq = Qr - q0 < 32 ? Qr - q0 : 32;
        for (int j = 0; j < nq && !stop; ++j) {
          const int r0 = __shfl_sync(kFull, my_r0, j), r1 = __shfl_sync(kFull, my_r1, j);
          if (lane == 0) {
            const TileRange t = tile_range<D>(r0, r1, total_cols);
            const long long rq = (long long)it * Qr + q0 + j;
            const int s = (int)(rq % stages);
            const uint32_t ph = (uint32_t)(rq / stages) & 1u;
            while (!mbar_try_wait(&empty[s], ph ^ 1u)) {
              if (*reinterpret_cast<volatile int*>(&s_stop)) {
                stop = true;
                break;
              }
            }
            if (!stop) {
              unsigned char* buf = stage0 + (size_t)s * stage_bytes;
              const uint32_t vbytes = (uint32_t)(t.v_hi - t.v_lo) * 8u, cbytes = (uint32_t)(t.c_hi - t.c_lo) * 4u;
              FEA_ASSERT(!t.direct && t.v_hi - t.v_lo <= val_cap && t.c_hi - t.c_lo <= col_cap);
              mbar_expect_tx(&full[s], vbytes + cbytes);
              if (vbytes) bulk_g2s(buf, values + t.v_lo, vbytes, &full[s], policy);
              if (cbytes) bulk_g2s(buf + sizeof(double) * val_cap, node_colidx + t.c_lo, cbytes, &full[s], policy);
              issued = rq + 1;
            }
          }
          stop = __shfl_sync(kFull, (int)stop, 0) != 0;
        }
      }
    }
    if (lane == 0 && stop) {
      // the consumers are gone: the latest copy into each stage must have landed before the CTA may end
      const long long first_live = issued > stages ? issued - stages : 0;
      for (long long rq = first_live; rq < issued; ++rq) mbar_wait(&full[(int)(rq % stages)], (uint32_t)(rq / stages) & 1u);
    }
    return;  // every copy issued here has been consumed (normal end) or has landed (early end)
  }

  // -------------------------------------------------------------------- consumers
  const int ctid = threadIdx.x;  // consumer warps come first
  const int64_t n = (int64_t)n_nodes * D;
  const int64_t slice = (n + gridDim.x - 1) / gridDim.x;
  const int64_t e_lo = (int64_t)blockIdx.x * slice < n ? (int64_t)blockIdx.x * slice : n;
  const int64_t e_hi = e_lo + slice < n ? e_lo + slice : n;
  const int nblk = (int)gridDim.x;
  double* p_delta = a.partials;

  // Solver scalars live in shared memory between iterations (phase 1 needs every register for its gathers):
  // [0] gamma (r.u of the current iterate)  [1] r.r  [2] gamma_old  [3] alpha_old  [4] best r.r  [5] its iteration
  // Every consumer thread computes the same new values in phase 2; thread 0 stores them.
  if (ctid == 0) {
    const int it0 = st->iter;
    s_sc[0] = it0 == 0 ? st->rz : st->rz_new;
    s_sc[1] = st->rr;
    s_sc[2] = st->rz;
    s_sc[3] = st->spare[0];
    s_sc[4] = st->spare[1];
    s_sc[5] = st->spare[2];
    s_sc[6] = st->bnorm2;
    s_sc[7] = st->tol2;
    s_iter = it0;
    s_max_iter = st->max_iter;
  }
  group_barrier(kFusedBarrierId, CT);
  unsigned* counter = &st->counter[3];
  unsigned barriers = 0;
  bool stopped = false, first = true;

  for (int it = 0; it < a.iters; ++it) {
    // ---------------------------------------------------------------- phase 1: w = K u, delta = u.w
    const double dot = fused_sweep<D, G>(n_nodes, node_rowptr, node_colidx, values, a.u, a.w, smem, a.stages_arg, val_cap,
                                         col_cap, it, Q, Qr);
    // The operands of this thread's first vector element that nobody else writes (everything but w) are fetched now:
    // their L2 latency hides behind the reduction and the grid barrier (whose fence empties L1 every time).
    const int64_t j0 = e_lo + ctid;
    const bool own0 = j0 < e_hi;
    double di0 = 0.0, u0 = 0.0, rv0 = 0.0, xv0 = 0.0, pv0 = 0.0, sv0 = 0.0;
    if (own0) {
      di0 = a.dinv[j0];
      u0 = __ldcg(a.u + j0);
      rv0 = a.r[j0];
      xv0 = a.x[j0];
      if (s_iter != 0) {
        pv0 = a.p[j0];
        sv0 = a.s[j0];
      }
    }
    {
      double v[1] = {dot};
      consumer_sum<NW, 1>(v, s_scratch, warp, lane, CT);
      if (ctid == 0) __stcg(p_delta + blockIdx.x, v[0]);
    }
    barriers += (unsigned)nblk;
    if (!fused_grid_barrier(counter, barriers, ctid, CT, &s_fail)) {
      if (blockIdx.x == 0 && ctid == 0) {
        st->status = FEA_ERR_CUDA;
        st->done = 1;
      }
      stopped = true;
      break;
    }

    // ---------------------------------------------------------------- phase 2: scalars, vector update
    const int iter = s_iter;
    const double bnorm2 = s_sc[6], tol2 = s_sc[7];
    const int max_iter = s_max_iter;
    double gamma = s_sc[0], rr = s_sc[1];
    const double gamma_old = s_sc[2], alpha_old = s_sc[3], best_rr = s_sc[4], best_it = s_sc[5];
    double delta;
    {
      // partials of this iteration's delta and -- unless this is the first iteration of the launch, whose gamma and
      // r.r come from the state block -- of the previous iteration's gamma' and r.r
      const int par = (iter - 1) & 1;
      const double* pg = a.partials + (size_t)nblk * (1 + 2 * par);
      const double* pr = pg + nblk;
      double v[3] = {0.0, 0.0, 0.0};
      for (int i = ctid; i < nblk; i += CT) {
        v[0] += __ldcg(p_delta + i);
        if (!first) {
          v[1] += __ldcg(pg + i);
          v[2] += __ldcg(pr + i);
        }
      }
      consumer_sum<NW, 3>(v, s_scratch, warp, lane, CT);  // (its barriers also order the s_sc reads above)
      delta = v[0];
      if (!first) {
        gamma = v[1];
        rr = v[2];
      }
      first = false;
    }
    const bool converged = rr <= tol2 * bnorm2;  // also covers a zero right-hand side
    const bool improved = !(rr >= 0.99 * best_rr) || best_rr == 0.0;
    const bool stalled = !converged && iter > 0 && !improved && iter - (int)best_it >= max(10000, max_iter / 1000);
    const bool exhausted = !converged && (iter >= max_iter || stalled);
    double beta = 0.0, alpha = 0.0;
    bool breakdown = false;
    if (!converged && !exhausted) {
      if (iter == 0) {
        breakdown = !(delta > 0.0);
        alpha = gamma / delta;
      } else {
        beta = gamma / gamma_old;
        const double denom = delta - beta * gamma / alpha_old;
        breakdown = !(denom > 0.0);
        alpha = gamma / denom;
      }
    }
    if (blockIdx.x == 0 && ctid == 0 && a.history != nullptr && iter >= 1 && iter <= max_iter)
      a.history[iter - 1] = sqrt(rr / bnorm2);
    if (converged || exhausted || breakdown) {
      if (blockIdx.x == 0 && ctid == 0) {
        st->status = converged ? FEA_OK
                               : (breakdown ? FEA_ERR_BREAKDOWN : (stalled ? FEA_ERR_STAGNATION : FEA_ERR_MAXITER));
        st->rr = rr;
        st->rr_final = rr;
        st->iter = iter;
        st->done = 1;
      }
      stopped = true;
      break;
    }
    double s_ru = 0.0, s_rr = 0.0;
    auto update = [&](int64_t j, double di, double uj, double wj, double rj, double xj, double pj, double sj) {
      const double pn = iter == 0 ? uj : fma(beta, pj, uj);
      const double sn = iter == 0 ? wj : fma(beta, sj, wj);
      const double rn = fma(-alpha, sn, rj);
      const double un = di * rn;
      a.p[j] = pn;
      a.s[j] = sn;
      a.x[j] = fma(alpha, pn, xj);
      a.r[j] = rn;
      a.u[j] = un;
      if (di != 0.0) {
        s_ru = fma(rn, un, s_ru);
        s_rr = fma(rn, rn, s_rr);
      }
    };
    if (own0) update(j0, di0, u0, __ldcg(a.w + j0), rv0, xv0, pv0, sv0);
    for (int64_t j = j0 + CT; j < e_hi; j += CT) {
      const double di = a.dinv[j], uj = __ldcg(a.u + j), wj = __ldcg(a.w + j), rj = a.r[j], xj = a.x[j];
      double pj = 0.0, sj = 0.0;
      if (iter != 0) {
        pj = a.p[j];
        sj = a.s[j];
      }
      update(j, di, uj, wj, rj, xj, pj, sj);
    }
    {
      double v[2] = {s_ru, s_rr};
      consumer_sum<NW, 2>(v, s_scratch, warp, lane, CT);
      if (ctid == 0) {
        double* pg = a.partials + (size_t)nblk * (1 + 2 * (iter & 1));
        __stcg(pg + blockIdx.x, v[0]);
        __stcg(pg + nblk + blockIdx.x, v[1]);
      }
    }
    if (ctid == 0) {  // read again only after the barriers of the next iteration's consumer_sum
      s_sc[0] = gamma;  // replaced by the reduced gamma' in the next phase 2
      s_sc[1] = rr;
      s_sc[2] = gamma;
      s_sc[3] = alpha;
      if (iter > 0 && improved) {
        s_sc[4] = rr;
        s_sc[5] = (double)iter;
      }
      s_iter = iter + 1;
    }
    barriers += (unsigned)nblk;
    if (!fused_grid_barrier(counter, barriers, ctid, CT, &s_fail)) {
      if (blockIdx.x == 0 && ctid == 0) {
        st->status = FEA_ERR_CUDA;
        st->done = 1;
      }
      stopped = true;
      break;
    }
  }
  if (stopped && ctid == 0) {
    *reinterpret_cast<volatile int*>(&s_stop) = 1;
    __threadfence_block();
  }

  // launch boundary: the state block gets the meaning it has between two kernels of the two-kernel path
  if (!stopped && blockIdx.x == 0) {
    group_barrier(kFusedBarrierId, CT);
    const int iter = s_iter;
    const double* pg = a.partials + (size_t)nblk * (1 + 2 * ((iter - 1) & 1));
    double v[2] = {0.0, 0.0};
    if (!first) {
      for (int i = ctid; i < nblk; i += CT) {
        v[0] += __ldcg(pg + i);
        v[1] += __ldcg(pg + nblk + i);
      }
    }
    consumer_sum<NW, 2>(v, s_scratch, warp, lane, CT);
    if (ctid == 0 && !first) {
      st->rz = s_sc[2];
      st->rz_new = v[0];
      st->rr = v[1];
      st->spare[0] = s_sc[3];
      st->spare[1] = s_sc[4];
      st->spare[2] = s_sc[5];
      st->iter = iter;
    }
  }
}

}  // namespace fea
