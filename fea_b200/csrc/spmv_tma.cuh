// Node-block CSR SpMV with a TMA bulk-copy pipeline (sm_90+/sm_100a: cp.async.bulk + mbarrier).
//
// Why: the matrix stream (values + node-level column indices) of a run of consecutive nodes is
// ONE contiguous byte range of each array, so a single elected thread can stream it into shared
// memory with 1-D bulk copies (SASS: UBLKCP) through a ring of stages, completely decoupled
// from the warps that consume it.  That keeps kStages x ~33 KB of HBM reads in flight per SM
// regardless of register pressure or instruction scheduling -- the plain-load kernel (spmv.cuh)
// only reaches ~2 KB per warp and ptxas serialises its loads when registers get tight.
//
// Once a tile sits in shared memory, random access is cheap, so the consumers work
// THREAD-PER-ROW ("CSR-stream"): lane <-> one DOF row of the tile.  Consecutive lanes read
// shared memory at a stride of one row (81 doubles on a hex8 mesh: odd, hence conflict-free),
// the D lanes of a node gather the same x entries (broadcast) and neighbouring nodes gather
// neighbouring entries (few sectors per request), there is no cross-lane reduction at all, and
// y is written with unit stride.
//
//   producer (1 thread)      for tile q = 0, 1, ... of this CTA: wait empty[q % S]; expect_tx;
//                            bulk-copy values[D*D*rp[n0] .. D*D*rp[n1]) and
//                            node_colidx[rp[n0] .. rp[n1]) into stage q % S
//   consumer groups (G x 2 warps)  group g takes tiles q = g, g+G, ...: wait full[q % S];
//                            row-per-lane products with 9 node columns (27 gathers) in flight;
//                            arrive on empty[q % S]
// Persistent grid (one CTA per SM); tile = blockIdx + gridDim * q, so the chip sweeps one
// contiguous window of the matrix and of x at a time.
#pragma once
#include "spmv.cuh"

namespace fea {

constexpr int kTileNodes = 16;
constexpr int kTmaGroups = 4;
constexpr int kTmaGroupWarps = 2;  // 64 lanes >= D * kTileNodes rows for D <= 3 (static_assert below)
constexpr int kTmaConsumerWarps = kTmaGroups * kTmaGroupWarps;
constexpr int kTmaThreads = (kTmaConsumerWarps + 1) * 32;  // + 1 producer warp
constexpr int kTmaMaxStages = 5;       // ~164 KB on a hex8 mesh: leaves ~90 KB of the SM array to L1 for the x gathers
constexpr int kTmaBarrierBytes = 128;  // full[kTmaMaxStages], empty[kTmaMaxStages], padded
constexpr int kTmaUnroll = 9;          // node columns per round: D * 9 gathers in flight per lane

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion reported to an mbarrier (bytes % 16 == 0, 16 B aligned).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TmaLayout {
  int stages;
  int val_cap;  // doubles per stage (even)
  int col_cap;  // int32 per stage (multiple of 4)
  size_t smem_bytes;
};

// Stage capacity for the worst tile: kTileNodes nodes of `maxc` coupled nodes each, plus the
// slack that rounding the byte range out to 16 B needs.
inline TmaLayout tma_layout(int d, int maxc, size_t smem_limit) {
  TmaLayout L;
  L.val_cap = kTileNodes * d * d * maxc + 2;
  L.val_cap += L.val_cap & 1;
  L.col_cap = (kTileNodes * maxc + 8 + 3) & ~3;
  const size_t per_stage = sizeof(double) * L.val_cap + sizeof(int32_t) * L.col_cap;
  const size_t fixed = kTmaBarrierBytes;
  int s = (int)((smem_limit - fixed) / per_stage);
  L.stages = s > kTmaMaxStages ? kTmaMaxStages : s;
  L.smem_bytes = fixed + (size_t)(L.stages > 0 ? L.stages : 0) * per_stage;
  return L;
}

// Byte ranges of one tile, rounded out to 16 B; `direct` when the rounding would run past the end
// of the arrays (only the last tile(s) of the matrix): those are read with plain loads instead.
struct TileRange {
  int64_t v_lo, v_hi, c_lo, c_hi;
  bool direct;
};
template <int D>
__device__ __forceinline__ TileRange tile_range(int64_t r0, int64_t r1, int64_t total_vals, int64_t total_cols) {
  TileRange t;
  t.v_lo = (D * D * r0) & ~1LL;
  t.v_hi = (D * D * r1 + 1) & ~1LL;
  t.c_lo = r0 & ~3LL;
  t.c_hi = (r1 + 3) & ~3LL;
  t.direct = t.v_hi > total_vals || t.c_hi > total_cols;
  return t;
}

// One DOF row against x: `vrow` points at the row's values, `cols` at the node's column list
// (shared or global memory).  kTmaUnroll node columns per round, all gathers issued first.
template <int D>
__device__ __forceinline__ double row_dot(const double* vrow, const int32_t* cols, int cnt,
                                          const double* __restrict__ x) {
  double acc[D];
#pragma unroll
  for (int b = 0; b < D; ++b) acc[b] = 0.0;
  for (int k0 = 0; k0 < cnt; k0 += kTmaUnroll) {
    double xv[kTmaUnroll][D];
#pragma unroll
    for (int u = 0; u < kTmaUnroll; ++u) {
      const int k = k0 + u;
      if (k < cnt) {
        const int64_t col = (int64_t)D * cols[k];
#pragma unroll
        for (int b = 0; b < D; ++b) xv[u][b] = __ldg(x + col + b);
      }
    }
#pragma unroll
    for (int u = 0; u < kTmaUnroll; ++u) {
      const int k = k0 + u;
      if (k < cnt) {
#pragma unroll
        for (int b = 0; b < D; ++b) acc[b] = fma(vrow[D * k + b], xv[u][b], acc[b]);
      }
    }
  }
  double s = acc[0];
#pragma unroll
  for (int b = 1; b < D; ++b) s += acc[b];
  return s;
}

// DOT: also accumulate sum_owned x_own[row] * y[row] (the PCG p.Ap); the caller reduces `dot`.
template <int D, bool DOT>
__device__ __forceinline__ void spmv_tma_body(int64_t n_nodes, const int32_t* __restrict__ node_rowptr,
                                              const int32_t* __restrict__ node_colidx,
                                              const double* __restrict__ values, const double* __restrict__ x,
                                              double* __restrict__ y, const double* __restrict__ x_own, int stages,
                                              int val_cap, int col_cap, unsigned char* smem, double& dot) {
  static_assert(D * kTileNodes <= kTmaGroupWarps * 32, "one lane per row of a tile");
  constexpr int DD = D * D;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kTmaMaxStages;
  unsigned char* stage0 = smem + kTmaBarrierBytes;
  const size_t stage_bytes = sizeof(double) * val_cap + sizeof(int32_t) * col_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaGroupWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int64_t n_tiles = (n_nodes + kTileNodes - 1) / kTileNodes;
  const int64_t total_cols = node_rowptr[n_nodes];
  const int64_t total_vals = (int64_t)DD * total_cols;

  if (warp == kTmaConsumerWarps) {
    // ------------------------------------------------------------------ producer
    // The whole warp fetches the row pointers of the next 32 tiles at once (one memory latency
    // per 32 tiles instead of one per tile -- a single thread chasing rp[] tile by tile caps the
    // CTA at one tile per DRAM round trip); lane 0 then issues the copies.
    auto fetch = [&](int64_t q, int& a0, int& a1) {
      const int64_t t = blockIdx.x + (int64_t)gridDim.x * q;
      a0 = a1 = 0;
      if (t < n_tiles) {
        const int64_t n0 = t * kTileNodes;
        a0 = node_rowptr[n0];
        a1 = node_rowptr[n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes];
      }
    };
    int nxt_r0, nxt_r1;
    fetch(lane, nxt_r0, nxt_r1);
    for (int64_t q0 = 0; blockIdx.x + (int64_t)gridDim.x * q0 < n_tiles; q0 += 32) {
      const int my_r0 = nxt_r0, my_r1 = nxt_r1;
      fetch(q0 + 32 + lane, nxt_r0, nxt_r1);  // the batch after this one, in flight while this one is issued
      for (int j = 0; j < 32; ++j) {
        const int r0 = __shfl_sync(kFull, my_r0, j), r1 = __shfl_sync(kFull, my_r1, j);
        const int64_t q = q0 + j;
        if (blockIdx.x + (int64_t)gridDim.x * q >= n_tiles) break;  // warp-uniform
        const TileRange t = tile_range<D>(r0, r1, total_vals, total_cols);
        if (lane == 0 && !t.direct) {
          const int s = (int)(q % stages);
          const uint32_t ph = (uint32_t)(q / stages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          unsigned char* buf = stage0 + (size_t)s * stage_bytes;
          const uint32_t vbytes = (uint32_t)(t.v_hi - t.v_lo) * 8u, cbytes = (uint32_t)(t.c_hi - t.c_lo) * 4u;
          mbar_expect_tx(&full[s], vbytes + cbytes);
          if (vbytes) bulk_g2s(buf, values + t.v_lo, vbytes, &full[s]);
          if (cbytes) bulk_g2s(buf + sizeof(double) * val_cap, node_colidx + t.c_lo, cbytes, &full[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int group = warp / kTmaGroupWarps;
    const int row = (warp % kTmaGroupWarps) * 32 + lane;  // row of the tile owned by this lane
    const int node_in_tile = row / D, a = row - node_in_tile * D;
    int64_t tile = blockIdx.x + (int64_t)gridDim.x * group;
    // row pointers of this lane's node for the first tile; later tiles are prefetched one ahead
    int64_t r0 = 0, r1 = 0;
    int lo = 0, hi = 0;
    if (tile < n_tiles) {
      const int64_t n0 = tile * kTileNodes;
      const int64_t n1 = n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes;
      const int64_t node = n0 + node_in_tile;
      r0 = node_rowptr[n0];
      r1 = node_rowptr[n1];
      if (node < n1) {
        lo = node_rowptr[node];
        hi = node_rowptr[node + 1];
      }
    }
    for (int q = group; tile < n_tiles; q += kTmaGroups) {
      const int64_t n0 = tile * kTileNodes;
      const int64_t n1 = n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes;
      const int64_t node = n0 + node_in_tile;
      const bool active = node < n1;
      const int64_t next = tile + (int64_t)gridDim.x * kTmaGroups;
      int64_t nr0 = 0, nr1 = 0;
      int nlo = 0, nhi = 0;
      if (next < n_tiles) {
        const int64_t m0 = next * kTileNodes;
        const int64_t m1 = m0 + kTileNodes < n_nodes ? m0 + kTileNodes : n_nodes;
        nr0 = node_rowptr[m0];
        nr1 = node_rowptr[m1];
        if (m0 + node_in_tile < m1) {
          nlo = node_rowptr[m0 + node_in_tile];
          nhi = node_rowptr[m0 + node_in_tile + 1];
        }
      }
      const TileRange t = tile_range<D>(r0, r1, total_vals, total_cols);
      const int cnt = hi - lo;
      double out = 0.0;
      if (t.direct) {
        if (active) out = row_dot<D>(values + (int64_t)DD * lo + (int64_t)a * D * cnt, node_colidx + lo, cnt, x);
      } else {
        const int s = q % stages;
        const uint32_t ph = (uint32_t)(q / stages) & 1u;
        mbar_wait(&full[s], ph);
        if (active) {
          const unsigned char* buf = stage0 + (size_t)s * stage_bytes;
          const double* vs = reinterpret_cast<const double*>(buf) + ((int64_t)DD * lo - t.v_lo);
          const int32_t* cs = reinterpret_cast<const int32_t*>(buf + sizeof(double) * val_cap) + (lo - t.c_lo);
          out = row_dot<D>(vs + a * D * cnt, cs, cnt, x);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (active) {
        y[node * D + a] = out;
        if (DOT) dot = fma(out, x_own[node * D + a], dot);
      }
      tile = next;
      r0 = nr0;
      r1 = nr1;
      lo = nlo;
      hi = nhi;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kTmaThreads, 1)
spmv_tma_kernel(int64_t n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y, int stages,
                int val_cap, int col_cap) {
  extern __shared__ __align__(128) unsigned char s_tma[];
  double dot = 0.0;
  spmv_tma_body<D, false>(n_nodes, node_rowptr, node_colidx, values, x, y, nullptr, stages, val_cap, col_cap, s_tma,
                          dot);
}

// Can the bulk-copy path be used for this matrix on this device?  (16 B aligned arrays, at
// least two stages of shared memory for the widest tile.)
struct TmaPlan {
  bool ok;
  TmaLayout layout;
  int grid;
};

TmaPlan tma_plan(int d, int max_coupled, const void* values, const void* node_colidx, int64_t n_nodes);

}  // namespace fea
