// Node-block CSR SpMV with a TMA bulk-copy pipeline (sm_90+/sm_100a: cp.async.bulk + mbarrier).
//
// Why: the matrix stream (values + node-level column indices) of a run of consecutive nodes is
// ONE contiguous byte range of each array, so a single elected thread can stream it into shared
// memory with 1-D bulk copies (SASS: UBLKCP) through a ring of stages, completely decoupled
// from the warps that consume it.  That keeps whole tiles (~33 KB on a hex8 mesh) of HBM reads
// in flight per SM regardless of register pressure or instruction scheduling -- the plain-load
// kernel (spmv.cuh) only reaches ~2 KB per warp and ptxas serialises its loads when registers
// get tight (1.8 ms vs 0.8 ms per SpMV at 400x80x80).
//
// Once a tile sits in shared memory, random access is cheap, so the consumers do not work
// warp-per-row.  A tile of 16 nodes has 16*D rows; every row is split by column component b
// into D partial products, and one lane owns one (row, b) pair:
//     partial(row, b) = sum_k  V[row][D*k + b] * x[D*col_k + b]
//   * lanes are ordered b-major, so the 32 lanes of a warp read shared memory at a stride of one
//     row (81 doubles on a hex8 mesh: odd => conflict-free 64-bit accesses);
//   * all gathers of a lane (27 on a hex8 mesh) are issued back to back: a tile costs about ONE
//     L2 round trip -- under a saturated memory system that latency is what bounds a consumer,
//     so several groups of 5 warps keep > 10k gathers in flight per SM;
//   * neighbouring rows gather neighbouring x entries (few sectors per request, L1 hits);
//   * the D partials of a row meet in shared memory (double-buffered, one named barrier per
//     tile and group); y is written with unit stride.  No warp shuffles, no atomics.
//
//   producer warp         tile q = 0, 1, ... of this CTA: wait empty[q % S]; expect_tx;
//                         bulk-copy values[D*D*rp[n0] .. D*D*rp[n1]) and node_colidx[rp[n0] .. rp[n1])
//   consumer group g      tiles q = g, g+G, ...: wait full[q % S]; partials; arrive empty[q % S]
// S is a multiple of G, so a stage is always consumed by the same group (mbarrier phase
// discipline).  Several CTAs (= independent rings) share an SM: tools/membench.cu shows one deep
// ring per SM streams 6.2 TB/s, 3 independent rings 6.9 TB/s.  Persistent grid;
// tile = blockIdx + gridDim * q, so the chip sweeps one contiguous window of the matrix and of
// x at a time.
#pragma once
#include <type_traits>

#include "spmv.cuh"

namespace fea {

#ifndef FEA_TILE_NODES
#define FEA_TILE_NODES 16
#endif
constexpr int kTileNodes = FEA_TILE_NODES;
constexpr int kTmaMaxGroups = 4;
constexpr int kTmaMaxStages = 8;
constexpr int kTmaBarrierBytes = 128;  // full[kTmaMaxStages], empty[kTmaMaxStages]
constexpr int kTmaUnroll = 27;         // gathers in flight per lane and round

__host__ __device__ constexpr int tma_items(int d) { return d * d * kTileNodes; }  // (row, b) pairs per tile
__host__ __device__ constexpr int tma_group_warps(int d) { return (tma_items(d) + 31) / 32; }
__host__ __device__ constexpr int tma_threads(int d, int groups) { return (groups * tma_group_warps(d) + 1) * 32; }
// CTAs per SM the register allocation must leave room for (the pipeline wants 3 rings per SM).
__host__ __device__ constexpr int tma_min_blocks(int d, int groups) {
  return tma_threads(d, groups) <= 352 ? 3 : (tma_threads(d, groups) <= 512 ? 2 : 1);
}
__host__ __device__ constexpr size_t tma_fixed_bytes(int d, int groups) {
  return kTmaBarrierBytes + sizeof(double) * 2 * groups * tma_items(d);
}

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
// The matrix is read exactly once per SpMV: evict-first in L2, so that the 126 MB L2 keeps the
// gathered vector (63 MB at 7.9 M DOF) instead of 5 GB of dead matrix lines.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ uint64_t l2_evict_normal_policy() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
// The kernels' `stages` argument carries the L2 hint of the matrix stream in bits 16..17:
// 0 evict-first (matrix larger than L2: keep L2 for the vector), 1 normal, 2 evict-last (matrix fits in L2 and is
// re-read every iteration of a solve: BASELINE config 3, 85 MB).
constexpr int kTmaHintShift = 16;
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct TmaLayout {
  int stages;
  int val_cap;  // doubles per stage (even)
  int col_cap;  // int32 per stage (multiple of 4)
  size_t smem_bytes;
};

// Stage capacity for the worst tile: kTileNodes nodes of `maxc` coupled nodes each, plus the
// slack that rounding the byte range out to 16 B needs.
inline TmaLayout tma_layout(int d, int groups, int maxc, int max_stages, size_t smem_limit) {
  TmaLayout L;
  L.val_cap = kTileNodes * d * d * maxc + 2;
  L.val_cap += L.val_cap & 1;
  L.col_cap = (kTileNodes * maxc + 8 + 3) & ~3;
  const size_t per_stage = sizeof(double) * L.val_cap + sizeof(int32_t) * L.col_cap;
  const size_t fixed = tma_fixed_bytes(d, groups);
  int s = smem_limit > fixed ? (int)((smem_limit - fixed) / per_stage) : 0;
  if (s > max_stages) s = max_stages;
  if (s > kTmaMaxStages) s = kTmaMaxStages;
  s -= s % groups;  // a stage is always consumed by the same group
  L.stages = s;
  L.smem_bytes = fixed + (size_t)s * per_stage;
  return L;
}

// Byte ranges of one tile, rounded out to 16 B; `direct` when the rounding would run past the end
// of the arrays (only the last tile(s) of the matrix): those are read with plain loads instead.
struct TileRange {
  // element indices: value offsets are 64-bit (d*d*nnz_blocks passes 2^31 at ~26 M DOF on a hex8 mesh,
  // 17 GB of values -- small for a 180 GB part); block offsets fit int32 (node_rowptr is int32)
  int64_t v_lo, v_hi;
  int c_lo, c_hi;
  bool direct;
};
template <int D>
__device__ __forceinline__ TileRange tile_range(int r0, int r1, int total_cols) {
  TileRange t;
  t.v_lo = ((int64_t)(D * D) * r0) & ~(int64_t)1;
  t.v_hi = ((int64_t)(D * D) * r1 + 1) & ~(int64_t)1;
  t.c_lo = r0 & ~3;
  t.c_hi = (int)(((int64_t)r1 + 3) & ~(int64_t)3);
  t.direct = t.v_hi > (int64_t)(D * D) * total_cols || t.c_hi > total_cols || t.c_hi < 0;
  return t;
}

// Multi-GPU halo gate (fea_pcg_solve_p2p).  The rows of x that the z-neighbours own are stored into
// this rank's x by the neighbours' vector kernels over NVLink, followed by a tag.  Only the tiles
// next to a slab face gather such rows, so the sweep starts `lower_tiles` tiles in (interior first;
// the lower-face tiles wrap around to the very end) and a consumer group only looks at the tag right
// before its first boundary tile: the halo exchange hides behind the interior of the SpMV.  Face tiles
// gather x through L2, not through the read-only path (see row_part_dot).
struct HaloGate {  // a kernel argument (by value): every field is known to the host
  const long long* tag_lower;  // this rank's halo_tag[0] / [1]; nullptr without that neighbour
  const long long* tag_upper;
  const int* iter;             // PcgState::iter: the tag to wait for is (epoch << 32 | *iter + 1)
  long long epoch;
  int lower_tiles;             // tiles [0, lower_tiles) gather lower-halo rows
  int upper_tiles;             // tiles [n_tiles - upper_tiles, n_tiles) gather upper-halo rows
  int* error;                  // set to 1 when a neighbour never delivers
  int aligned;                 // owned / halo borders of x fall on 128-byte boundaries (see below)
};

// A consumer group is about to start its face tiles: its leader polls the neighbours' tags, then the
// group's named barrier orders every thread of the group after the acquire.
__device__ __forceinline__ void halo_gate_wait(const HaloGate& gate, bool leader, int barrier_id, int barrier_threads) {
  if (leader) {
    const long long want = (gate.epoch << 32) | ((long long)*gate.iter + 1);
    bool ok = true;
    if (gate.tag_lower != nullptr) ok = spin_until_backoff(gate.tag_lower, want) && ok;
    if (gate.tag_upper != nullptr) ok = spin_until_backoff(gate.tag_upper, want) && ok;
    if (!ok) *gate.error = 1;
  }
  group_barrier(barrier_id, barrier_threads);
}

// Component b of one DOF row against x: sum_k vrow[D*k + b] * x[D*cols[k] + b].
// `vrow` / `cols` may point to shared or global memory.  Every gather of a round is issued
// before the first FMA.
// COHERENT: gather through L2 (ld.global.cg) instead of the read-only path.  The FACE tiles of a
// multi-GPU slab need it: halo rows of x are written by the neighbour GPUs while the kernel runs, which
// ld.global.nc must never see, and a 128-byte line straddling the owned / halo border may already sit
// in L1 from an interior tile.  Interior tiles keep the read-only path (measured on one GPU: ordinary
// or L2 loads for EVERY tile cost the SpMV 7-10 %).
template <int D, bool COHERENT = false>
__device__ __forceinline__ double row_part_dot(const double* vrow, const int32_t* cols, int cnt, int b,
                                               const double* x) {
  double acc = 0.0;
  const double* xb = x + b;
  const double* vb = vrow + b;
  for (int k0 = 0; k0 < cnt; k0 += kTmaUnroll) {
    double xv[kTmaUnroll];
#pragma unroll
    for (int u = 0; u < kTmaUnroll; ++u) {
      const int k = k0 + u;
      if (COHERENT)
        xv[u] = k < cnt ? __ldcg(xb + (int64_t)D * cols[k]) : 0.0;
      else
        xv[u] = k < cnt ? __ldg(xb + (int64_t)D * cols[k]) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kTmaUnroll; ++u) {
      const int k = k0 + u;
      if (k < cnt) acc = fma(vb[D * k], xv[u], acc);
    }
  }
  return acc;
}

// DOT: also accumulate sum_owned x_own[row] * y[row] (the PCG p.Ap); the caller reduces `dot`.
template <int D, int G, bool DOT, bool GATED = false>
__device__ __forceinline__ void spmv_tma_body(int n_nodes, const int32_t* __restrict__ node_rowptr,
                                              const int32_t* __restrict__ node_colidx,
                                              const double* __restrict__ values, const double* x,
                                              double* __restrict__ y, const double* x_own, int stages_arg,
                                              int val_cap, int col_cap, unsigned char* smem, double& dot,
                                              const HaloGate& gate = HaloGate{}) {
  constexpr int DD = D * D;
  const int stages = stages_arg & ((1 << kTmaHintShift) - 1), l2_hint = stages_arg >> kTmaHintShift;
  constexpr int ROWS = D * kTileNodes;  // rows per tile
  constexpr int ITEMS = tma_items(D);   // (row, b) pairs per tile
  constexpr int GW = tma_group_warps(D);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kTmaMaxStages;
  double* parts = reinterpret_cast<double*>(smem + kTmaBarrierBytes);  // [2][G][ITEMS]
  unsigned char* stage0 = smem + tma_fixed_bytes(D, G);
  const int stage_bytes = (int)(sizeof(double) * val_cap + sizeof(int32_t) * col_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int n_tiles = (n_nodes + kTileNodes - 1) / kTileNodes;
  const int total_cols = node_rowptr[n_nodes];
  const int stride = (int)gridDim.x;
  // sweep position t -> tile: rotated by the lower-face tiles when gated (interior first)
  const int lower_tiles = GATED ? gate.lower_tiles : 0, upper_tiles = GATED ? gate.upper_tiles : 0;
  const int rot = lower_tiles;
  auto tile_of = [&](int64_t t) -> int {
    const int64_t p = t + rot;
    return (int)(p < n_tiles ? p : p - n_tiles);
  };

  if (warp == G * GW) {
    // ------------------------------------------------------------------ producer
    // The whole warp fetches the row pointers of the next 32 tiles at once (one memory latency
    // per 32 tiles instead of one per tile); lane 0 then issues the copies.
    auto fetch = [&](int q, int& a0, int& a1) {
      const int64_t t = blockIdx.x + (int64_t)stride * q;
      a0 = a1 = 0;
      if (t < n_tiles) {
        const int n0 = tile_of(t) * kTileNodes;
        a0 = node_rowptr[n0];
        a1 = node_rowptr[n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes];
      }
    };
    const uint64_t policy = l2_hint == 0 ? l2_evict_first_policy()
                                         : (l2_hint == 1 ? l2_evict_normal_policy() : l2_evict_last_policy());
    int nxt_r0, nxt_r1;
    fetch(lane, nxt_r0, nxt_r1);
    for (int q0 = 0; blockIdx.x + (int64_t)stride * q0 < n_tiles; q0 += 32) {
      const int my_r0 = nxt_r0, my_r1 = nxt_r1;
      fetch(q0 + 32 + lane, nxt_r0, nxt_r1);  // the batch after this one, in flight while this one is issued
      for (int j = 0; j < 32; ++j) {
        const int r0 = __shfl_sync(kFull, my_r0, j), r1 = __shfl_sync(kFull, my_r1, j);
        const int q = q0 + j;
        if (blockIdx.x + (int64_t)stride * q >= n_tiles) break;  // warp-uniform
        const TileRange t = tile_range<D>(r0, r1, total_cols);
        if (lane == 0 && !t.direct) {
          const int s = q % stages;
          const uint32_t ph = (uint32_t)(q / stages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          unsigned char* buf = stage0 + (size_t)s * stage_bytes;
          const uint32_t vbytes = (uint32_t)(t.v_hi - t.v_lo) * 8u, cbytes = (uint32_t)(t.c_hi - t.c_lo) * 4u;
          FEA_ASSERT(t.v_lo >= 0 && t.v_hi >= t.v_lo && t.v_hi - t.v_lo <= val_cap && t.c_hi - t.c_lo <= col_cap);
          FEA_ASSERT(t.v_hi <= (int64_t)DD * total_cols && t.c_hi <= total_cols && t.c_lo >= 0);
          FEA_ASSERT(vbytes % 16 == 0 && cbytes % 16 == 0);
          mbar_expect_tx(&full[s], vbytes + cbytes);
          if (vbytes) bulk_g2s(buf, values + t.v_lo, vbytes, &full[s], policy);
          if (cbytes) bulk_g2s(buf + sizeof(double) * val_cap, node_colidx + t.c_lo, cbytes, &full[s], policy);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int group = warp / GW;
    const int item = (warp - group * GW) * 32 + lane;  // b-major: item = b * ROWS + row
    const bool has_item = item < ITEMS;
    const int b = has_item ? item / ROWS : 0;
    const int row = has_item ? item - b * ROWS : 0;
    const int node_in_tile = row / D, a = row - node_in_tile * D;
    double* parts_g = parts + group * ITEMS;
    int64_t tile64 = blockIdx.x + (int64_t)stride * group;
    // row pointers of this lane's node for the first tile; later tiles are prefetched one ahead
    int r0 = 0, r1 = 0, lo = 0, hi = 0;
    if (tile64 < n_tiles) {
      const int n0 = tile_of(tile64) * kTileNodes;
      const int n1 = n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes;
      r0 = node_rowptr[n0];
      r1 = node_rowptr[n1];
      if (n0 + node_in_tile < n1) {
        lo = node_rowptr[n0 + node_in_tile];
        hi = node_rowptr[n0 + node_in_tile + 1];
      }
    }
    int buf_sel = 0;
    int q = group;
    // The tile loop as a function of where it stops: the gated kernel runs it twice, over the interior
    // positions and then over the face tiles, with the halo wait BETWEEN the two runs.  (With the wait --
    // a spin loop and a named barrier, inlined or not -- inside the loop, ptxas took the kernel from 50 to
    // 80 registers plus spills: 2 resident CTAs per SM instead of 3.)
    auto sweep = [&](int64_t limit, auto coherent_tag) {
    constexpr bool COHERENT = decltype(coherent_tag)::value;
    for (; tile64 < limit; q += G) {
      const int tile = tile_of(tile64);
      const int n0 = tile * kTileNodes;
      const int n1 = n0 + kTileNodes < n_nodes ? n0 + kTileNodes : n_nodes;
      const int node = n0 + node_in_tile;
      const bool active = has_item && node < n1;
      tile64 += (int64_t)stride * G;
      int nr0 = 0, nr1 = 0, nlo = 0, nhi = 0;
      if (tile64 < n_tiles) {
        const int m0 = tile_of(tile64) * kTileNodes;
        const int m1 = m0 + kTileNodes < n_nodes ? m0 + kTileNodes : n_nodes;
        nr0 = node_rowptr[m0];
        nr1 = node_rowptr[m1];
        if (m0 + node_in_tile < m1) {
          nlo = node_rowptr[m0 + node_in_tile];
          nhi = node_rowptr[m0 + node_in_tile + 1];
        }
      }
      const TileRange t = tile_range<D>(r0, r1, total_cols);
      const int cnt = hi - lo;
      double part = 0.0;
      if (t.direct) {
        if (active) {
          const double* vg = values + (int64_t)DD * lo + a * D * cnt;
          part = row_part_dot<D, COHERENT>(vg, node_colidx + lo, cnt, b, x);
        }
      } else {
        const int s = q % stages;
        const uint32_t ph = (uint32_t)(q / stages) & 1u;
        FEA_ASSERT(s % G == group % G || stages % G != 0);  // a stage is always consumed by the same group
        mbar_wait(&full[s], ph);
        if (active) {
          const unsigned char* buf = stage0 + (size_t)s * stage_bytes;
          const double* vs = reinterpret_cast<const double*>(buf) + (int)((int64_t)DD * lo - t.v_lo);
          const int32_t* cs = reinterpret_cast<const int32_t*>(buf + sizeof(double) * val_cap) + (lo - t.c_lo);
          FEA_ASSERT((int64_t)DD * lo >= t.v_lo && (int64_t)DD * hi <= t.v_hi && lo >= t.c_lo && hi <= t.c_hi);
          FEA_ASSERT(cnt >= 0 && (int)((int64_t)DD * lo - t.v_lo) + DD * cnt <= val_cap && (lo - t.c_lo) + cnt <= col_cap);
          part = row_part_dot<D, COHERENT>(vs + a * D * cnt, cs, cnt, b, x);
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
        y[(int64_t)node * D + a] = out;
        if (DOT) dot = fma(out, x_own[(int64_t)node * D + a], dot);
      }
      buf_sel ^= 1;
      r0 = nr0;
      r1 = nr1;
      lo = nlo;
      hi = nhi;
    }
    };
    if (GATED) {
      // sweep positions [0, n_tiles - lower - upper) are interior tiles, then the upper face, then the
      // lower face (rotation above); a group that owns face tiles waits for both neighbours once
      sweep(n_tiles - lower_tiles - upper_tiles, std::false_type{});
      if (tile64 < n_tiles) {  // group-uniform
        halo_gate_wait(gate, warp == group * GW && lane == 0, 1 + group, GW * 32);
        // With 128-byte aligned slab borders no line holds both owned and halo rows: halo lines are
        // first touched here, after the acquire, and may come through L1 like everything else (measured
        // at N = 8: a rank with two faces lost ~7 us per SpMV to gathering its face tiles through L2).
        if (gate.aligned)
          sweep(n_tiles, std::false_type{});
        else
          sweep(n_tiles, std::true_type{});
      }
    } else {
      sweep(n_tiles, std::false_type{});
    }
  }
}

template <int D, int G>
__global__ void __launch_bounds__(tma_threads(D, G))
spmv_tma_kernel(int n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y, int stages,
                int val_cap, int col_cap) {
  extern __shared__ __align__(128) unsigned char s_tma[];
  double dot = 0.0;
  spmv_tma_body<D, G, false>(n_nodes, node_rowptr, node_colidx, values, x, y, nullptr, stages, val_cap, col_cap,
                             s_tma, dot);
}

// Can the bulk-copy path be used for this matrix on this device?  (16 B aligned arrays, enough
// shared memory for the widest tile, int32 node count.)
struct TmaPlan {
  bool ok;
  TmaLayout layout;
  int max_grid;  // SMs x target CTAs per SM, capped by the number of tiles
  int groups;    // consumer groups per CTA: 1 .. 4
  int l2_hint = 0;   // see kTmaHintShift
  bool pdl = false;  // launch with programmatic stream serialisation (fea_pcg_solve's private stream / graph)
};

// `small_problem`: the caller will sweep this matrix many times and it is small (fea_pcg_solve below the
// single-reduction threshold): 3 consumer groups x 3 stages, 2 CTAs per SM instead of 2 x 2, 3 per SM.
TmaPlan tma_plan(int d, int max_coupled, const void* values, const void* node_colidx, int64_t n_nodes,
                 bool small_problem = false);

}  // namespace fea
