// Node-block CSR SpMV (the dominant kernel of the solve: cubebeam.py:98 np.linalg.solve is
// replaced by Jacobi-PCG, whose cost is one K p per iteration; cubebeam.py:106 K @ u reuses it).
//
// Layout in HBM.  `values` is in ordinary DOF-level CSR order, but because the D rows of a node
// share one column set the kernel reads the node-level column list (one int32 per DxD block,
// 4/D^2 bytes per non-zero instead of 4) -- 8.44 B per non-zero for D = 3 instead of CSR's 12.
// The D rows of node i are one contiguous chunk values[D*D*lo .. D*D*hi): row a starts at
// a*D*cnt, entry (k, b) at D*k + b.
//
// Mapping.  One warp per node; lane <-> column position c = D*k + b of the node's rows, so each
// row is read with unit-stride 8-byte loads (fully coalesced, streamed with evict-first so the
// 126 MB L2 stays available for the gathered vector), the gathered x[D*col + b] is loaded once
// and reused for the D rows, and D warp-shuffle reductions finish the node.  HBM-bound.
#pragma once
#include "common.cuh"

namespace fea {

constexpr int kSpmvThreads = 256;
constexpr int kSpmvWarps = kSpmvThreads / 32;
constexpr int kMaxPartials = 2048;  // upper bound on grid size of every reducing kernel

// Computes y[D*node + a] for one node; all lanes return the D row sums.
template <int D>
__device__ __forceinline__ void spmv_node(const int32_t* __restrict__ node_colidx, const double* __restrict__ values,
                                          const double* __restrict__ x, int lo, int cnt, int lane, double out[D]) {
  const int row_len = D * cnt;
  const double* v = values + (int64_t)(D * D) * lo;
  const int32_t* cols = node_colidx + lo;
  double acc[D];
#pragma unroll
  for (int a = 0; a < D; ++a) acc[a] = 0.0;
  for (int base = 0; base < row_len; base += 96) {
    // three column positions per lane per trip: issue every load before the first FMA
    double xv[3];
    double vv[3][D];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int c = base + 32 * u + lane;
      xv[u] = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) vv[u][a] = 0.0;
      if (c < row_len) {
        const int k = c / D, b = c - k * D;
        const int col = ld_stream(cols + k);
#pragma unroll
        for (int a = 0; a < D; ++a) vv[u][a] = ld_stream(v + a * row_len + c);
        xv[u] = __ldg(x + (int64_t)D * col + b);
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u)
#pragma unroll
      for (int a = 0; a < D; ++a) acc[a] = fma(vv[u][a], xv[u], acc[a]);
  }
#pragma unroll
  for (int a = 0; a < D; ++a) out[a] = warp_sum(acc[a]);
}

}  // namespace fea
