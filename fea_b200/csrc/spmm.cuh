// Node-block CSR times a block of right-hand sides: Y = K X with X, Y (n_dof, R) row-major
// (BASELINE config 5: 64 load cases).  Shared by fea_spmm (misc.cu) and the multi-RHS PCG (multi.cu).
//
// Mapping.  One warp per node; lane <-> CPL adjacent columns of a 32*CPL-wide column tile, so a
// row of X is one coalesced 256*CPL-byte segment (LDG.128 per lane for CPL = 2) and all D*D block
// entries act on it.  The first version of this kernel was latency-bound (ncu: 11 % L2, 12 % FP64
// pipe, 28 % warps active -- profiles/kernels_r01_ncu.txt): one dependent colidx -> X-row chain
// per trip and D*D warp-uniform global loads for the block.  Now, per round of up to 32 coupled
// nodes:
//   * lane l fetches column id l of the round (one coalesced load), ids are passed by shuffle;
//   * the round's D*D*32 matrix values are copied once, coalesced and streaming (evict-first),
//     into a warp-private shared-memory slab laid out [k][b][a] (padded to 16 B), so the FMA loop
//     reads them as broadcast LDS.128;
//   * the X rows of U coupled nodes (U*D independent 16-byte gathers per lane) are issued before
//     the first FMA of the group.
// X rows of neighbouring nodes overlap heavily, and the 8 warps of a CTA work on 8 consecutive
// nodes: most gathers hit L1.  Bound: FP64 pipe ~ HBM (18 flop per matrix byte at R = 64).
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace fea {

constexpr int kSpmmChunk = 32;  // coupled nodes staged per round
constexpr int kSpmmWarps = 8;   // warps (= nodes in flight) per CTA

__host__ __device__ constexpr int spmm_val_stride(int d) { return d * d + ((d * d) & 1); }  // doubles, 16 B multiple
__host__ __device__ constexpr int spmm_slab_doubles(int d) { return kSpmmChunk * spmm_val_stride(d); }

// 16-byte shared-memory load by 32-bit shared address (keeps one register for the slab base
// instead of a rematerialised generic pointer).
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t shared_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Address of X row (D*col + b), column col0: one 32x32->64 multiply-add per coupled node.
__device__ __forceinline__ const double* x_row_ptr(const double* Xc, unsigned col, unsigned node_bytes) {
  return reinterpret_cast<const double*>(reinterpret_cast<const char*>(Xc) + (unsigned long long)col * node_bytes);
}
__device__ __forceinline__ const double* x_next_row(const double* row, unsigned row_bytes) {
  return reinterpret_cast<const double*>(reinterpret_cast<const char*>(row) + row_bytes);
}

template <int CPL>
struct ColVec;
template <>
struct ColVec<1> {
  static __device__ __forceinline__ void load(const double* p, double (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void load_plain(const double* p, double (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void store(double* p, const double (&v)[1]) { *p = v[0]; }
};
template <>
struct ColVec<2> {  // requires an even R and 16-byte aligned arrays (checked by the launcher)
  static __device__ __forceinline__ void load(const double* p, double (&v)[2]) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = t.x;
    v[1] = t.y;
  }
  static __device__ __forceinline__ void load_plain(const double* p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x;
    v[1] = t.y;
  }
  static __device__ __forceinline__ void store(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  }
};

// acc[a][j] = sum_k sum_b K[(node,a), (col_k,b)] * X[D*col_k + b][col0 + j]   for one node.
// `cols` = node_colidx + lo, `v` = values + D*D*lo, `active` = this lane's columns exist (col0 < R).
// `slab` is this warp's private shared-memory staging area (spmm_slab_doubles(D) doubles, 16 B aligned).
template <int D, int CPL, int U>
__device__ __forceinline__ void spmm_node(const int32_t* __restrict__ cols, const double* __restrict__ v, int cnt,
                                          const double* __restrict__ X, int R, int col0, bool active, int lane,
                                          double* slab, double (&acc)[D][CPL]) {
  constexpr int DD = D * D;
  constexpr int VS = spmm_val_stride(D);
  constexpr int NV2 = VS / 2;
  const int row_len = D * cnt;
  const double* Xc = X + (active ? col0 : 0);
  const unsigned row_bytes = 8u * (unsigned)R, node_bytes = D * row_bytes;
  const uint32_t slab_s = shared_addr(slab);
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[a][j] = 0.0;
  for (int k0 = 0; k0 < cnt; k0 += kSpmmChunk) {
    const int kc = min(kSpmmChunk, cnt - k0);
    const int mycol = lane < kc ? ld_stream(cols + k0 + lane) : 0;
    __syncwarp();  // the previous round's FMAs are done with the slab
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double* src = v + a * row_len + D * k0;
      for (int c = lane; c < D * kc; c += 32) {
        const int k = c / D, b = c - k * D;
        slab[k * VS + b * D + a] = ld_stream(src + c);
      }
    }
    __syncwarp();
    // X rows of U coupled nodes are gathered (U*D independent loads per lane) before the first FMA of
    // the group.  No predication in the loops: lanes without columns (col0 >= R) read column 0 and
    // simply never store; the tail of the round runs one coupled node at a time.
    auto fma_block = [&](int k, const double (&x)[D][CPL]) {
      const uint32_t sv = slab_s + (uint32_t)k * (VS * 8);
      double m[2 * NV2];
#pragma unroll
      for (int q = 0; q < NV2; ++q) {
        const double2 t = lds_f64x2(sv + 16 * q);
        m[2 * q] = t.x;
        m[2 * q + 1] = t.y;
      }
#pragma unroll
      for (int b = 0; b < D; ++b)
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[a][j] = fma(m[b * D + a], x[b][j], acc[a][j]);
    };
    int k = 0;
    for (; k + U <= kc; k += U) {
      double xv[U][D][CPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double* xrow = x_row_ptr(Xc, (unsigned)__shfl_sync(kFull, mycol, k + u), node_bytes);
#pragma unroll
        for (int b = 0; b < D; ++b) {
          ColVec<CPL>::load(xrow, xv[u][b]);
          xrow = x_next_row(xrow, row_bytes);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) fma_block(k + u, xv[u]);
    }
    for (; k < kc; ++k) {
      double xv[D][CPL];
      const double* xrow = x_row_ptr(Xc, (unsigned)__shfl_sync(kFull, mycol, k), node_bytes);
#pragma unroll
      for (int b = 0; b < D; ++b) {
        ColVec<CPL>::load(xrow, xv[b]);
        xrow = x_next_row(xrow, row_bytes);
      }
      fma_block(k, xv);
    }
  }
  (void)DD;
}

// ---------------------------------------------------------------------------------------------
// G consecutive nodes per warp with register-level reuse of the gathered X rows.
//
// ncu on the one-node-per-warp kernel (lattice truss n = 93, 64 RHS): L1TEX 61 % busy, FP64 pipe
// 22 %: every coupled node costs D 512-byte gathers that feed only D*D*CPL FMAs per lane, and a
// multi-line LDG replays at ~2 cycles per 128-byte line.  Neighbouring nodes of a mesh share most of
// their coupled nodes (27-point stencils: 2/3 between i and i+1), so a warp that owns G consecutive
// nodes gathers the UNION of their column lists once and applies each X row to every node of the
// group that couples to it: 54 gathers instead of 108 for G = 4 on a lattice / hex mesh.
//
//   stage    column lists (sorted, padded with INT_MAX) and the [k][b][a] value slabs of the G
//            nodes -> warp-private shared memory (G*D independent coalesced loads per lane)
//   phase A  union table: list 0 in order, then the entries of list 1 not yet covered, ...; an entry
//            is (column id, position in each node's list or 0xFF), positions found by binary
//            search in the sorted lists; lanes work on one list entry each
//   phase B  U entries per trip: D gathers each, then for every node of the group that has the
//            entry 18*CPL FMAs against its block (broadcast LDS.128 from the slab)
// Groups containing a node with more than 32 coupled nodes take the one-node path.
template <int D, int G>
struct SpmmGroupSmem {
  double slab[G][kSpmmChunk * spmm_val_stride(D)];
  unsigned long long entry[G * kSpmmChunk];
  int cols[G][kSpmmChunk];
  unsigned char flag[G][kSpmmChunk];
};

template <int D, int CPL, int G, int U>
__device__ __forceinline__ void spmm_group(const int32_t* __restrict__ node_colidx,
                                           const double* __restrict__ values, const int (&lo)[G],
                                           const int (&cnt)[G], const double* __restrict__ X, int R, int col0,
                                           bool active, int lane, SpmmGroupSmem<D, G>& sm,
                                           double (&acc)[G][D][CPL]) {
  static_assert(G <= 4, "entry encoding holds 4 positions");
  constexpr int VS = spmm_val_stride(D);
  constexpr int NV2 = VS / 2;
  __syncwarp();  // previous group is done with the shared arrays
#pragma unroll
  for (int g = 0; g < G; ++g) {
    sm.cols[g][lane] = lane < cnt[g] ? ld_stream(node_colidx + lo[g] + lane) : INT32_MAX;
    sm.flag[g][lane] = 0;
    const int row_len = D * cnt[g];
    const double* v = values + (int64_t)(D * D) * lo[g];
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
      for (int i = 0; i < D; ++i) {  // D*cnt <= D*32 entries per row: D trips
        const int c = lane + 32 * i;
        if (c < row_len) {
          const int k = c / D, b = c - k * D;
          sm.slab[g][k * VS + b * D + a] = ld_stream(v + a * row_len + c);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[g][a][j] = 0.0;
  }
  __syncwarp();
  // ---- phase A: union of the G sorted lists
  int n_union = 0;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    const bool mine = lane < cnt[s];
    const bool fresh = mine && sm.flag[s][lane] == 0;
    const int col = sm.cols[s][lane];
    unsigned long long ent = (unsigned long long)(unsigned)col | 0xffffffff00000000ull;
    ent &= ~(0xffull << (32 + 8 * s));
    ent |= (unsigned long long)lane << (32 + 8 * s);
#pragma unroll
    for (int g = s + 1; g < G; ++g) {
      int pos = 0;  // lower bound in the padded 32-entry list
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (sm.cols[g][pos + step - 1] < col) pos += step;
      if (fresh && pos < cnt[g] && sm.cols[g][pos] == col) {
        sm.flag[g][pos] = 1;
        ent &= ~(0xffull << (32 + 8 * g));
        ent |= (unsigned long long)pos << (32 + 8 * g);
      }
    }
    const unsigned ballot = __ballot_sync(kFull, fresh);
    if (fresh) sm.entry[n_union + __popc(ballot & ((1u << lane) - 1u))] = ent;
    n_union += __popc(ballot);
    __syncwarp();
  }
  // ---- phase B (no predication: see spmm_node)
  const double* Xc = X + (active ? col0 : 0);
  const unsigned row_bytes = 8u * (unsigned)R, node_bytes = D * row_bytes;
  const uint32_t slab_s = shared_addr(&sm.slab[0][0]);
  auto apply = [&](unsigned long long ent, const double (&x)[D][CPL]) {
    const unsigned posw = (unsigned)(ent >> 32);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const unsigned pos = (posw >> (8 * g)) & 0xffu;
      if (pos != 0xffu) {  // warp-uniform
        const uint32_t sv = slab_s + (uint32_t)(g * kSpmmChunk + pos) * (VS * 8);
        double m[2 * NV2];
#pragma unroll
        for (int q = 0; q < NV2; ++q) {
          const double2 t = lds_f64x2(sv + 16 * q);
          m[2 * q] = t.x;
          m[2 * q + 1] = t.y;
        }
#pragma unroll
        for (int b = 0; b < D; ++b)
#pragma unroll
          for (int a = 0; a < D; ++a)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[g][a][j] = fma(m[b * D + a], x[b][j], acc[g][a][j]);
      }
    }
  };
  int e = 0;
  for (; e + U <= n_union; e += U) {
    unsigned long long ent[U];
    double xv[U][D][CPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ent[u] = sm.entry[e + u];
      const double* xrow = x_row_ptr(Xc, (unsigned)ent[u], node_bytes);
#pragma unroll
      for (int b = 0; b < D; ++b) {
        ColVec<CPL>::load(xrow, xv[u][b]);
        xrow = x_next_row(xrow, row_bytes);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) apply(ent[u], xv[u]);
  }
  for (; e < n_union; ++e) {
    const unsigned long long ent = sm.entry[e];
    double xv[D][CPL];
    const double* xrow = x_row_ptr(Xc, (unsigned)ent, node_bytes);
#pragma unroll
    for (int b = 0; b < D; ++b) {
      ColVec<CPL>::load(xrow, xv[b]);
      xrow = x_next_row(xrow, row_bytes);
    }
    apply(ent, xv);
  }
}

// Row pointers of the G nodes starting at node0 (cnt = 0 past the end); true if the group can take
// the union path (every node couples to at most 32 nodes).
template <int G>
__device__ __forceinline__ bool spmm_group_rows(const int32_t* __restrict__ node_rowptr, int64_t node0,
                                                int64_t n_nodes, int lane, int (&lo)[G], int (&cnt)[G]) {
  const int64_t idx = node0 + lane;
  const int rp = lane <= G && idx <= n_nodes ? node_rowptr[idx] : 0;
  bool ok = true;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    lo[g] = __shfl_sync(kFull, rp, g);
    const int hi = __shfl_sync(kFull, rp, g + 1);
    cnt[g] = node0 + g < n_nodes ? hi - lo[g] : 0;
    ok = ok && cnt[g] <= kSpmmChunk;
  }
  return ok;
}

// Sweep of one CTA (kSpmmWarps warps, persistent, group-strided) over the matrix for one column tile:
// Y rows are stored, and with DOT the per-column partial sum_rows X[row] * Y[row] is accumulated.
template <int D, int CPL, int G, int U, bool DOT>
__device__ __forceinline__ void spmm_sweep(int64_t n_nodes, const int32_t* __restrict__ node_rowptr,
                                           const int32_t* __restrict__ node_colidx,
                                           const double* __restrict__ values, const double* __restrict__ X,
                                           double* __restrict__ Y, int R, int col0, bool active, int lane, int warp,
                                           SpmmGroupSmem<D, G>& sm, double (&dot)[1][CPL],
                                           const double* __restrict__ Xown = nullptr) {
  // Xown: the rows of X that belong to the output rows (X + row offset of a slab's owned rows); X itself
  // when null.  Only the fused dot reads it.
  if (Xown == nullptr) Xown = X;
  auto finish = [&](int64_t node, const double (&acc)[D][CPL]) {
    if (!active) return;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const int64_t idx = (node * D + a) * R + col0;
      ColVec<CPL>::store(Y + idx, acc[a]);
      if (DOT) {
        double own[CPL];
        ColVec<CPL>::load(Xown + idx, own);
#pragma unroll
        for (int j = 0; j < CPL; ++j) dot[0][j] = fma(acc[a][j], own[j], dot[0][j]);
      }
    }
  };
  for (int64_t grp = (int64_t)blockIdx.x * kSpmmWarps + warp; grp * G < n_nodes;
       grp += (int64_t)gridDim.x * kSpmmWarps) {
    const int64_t node0 = grp * G;
    int lo[G], cnt[G];
    if (spmm_group_rows<G>(node_rowptr, node0, n_nodes, lane, lo, cnt)) {
      double acc[G][D][CPL];
      spmm_group<D, CPL, G, U>(node_colidx, values, lo, cnt, X, R, col0, active, lane, sm, acc);
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (node0 + g < n_nodes) finish(node0 + g, acc[g]);
    } else {
      for (int g = 0; g < G; ++g) {
        if (node0 + g >= n_nodes) break;
        double acc1[D][CPL];
        spmm_node<D, CPL, U>(node_colidx + lo[g], values + (int64_t)(D * D) * lo[g], cnt[g], X, R, col0, active, lane,
                             sm.slab[0], acc1);
        finish(node0 + g, acc1);
      }
    }
  }
}

constexpr int kSpmmGroup = 4;  // nodes per warp

// CPL = 2 needs 16-byte vector access: even R and 16-byte aligned X / Y.
inline bool spmm_can_vectorise(int R, const void* X, const void* Y) {
  return R % 2 == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0;
}

}  // namespace fea
