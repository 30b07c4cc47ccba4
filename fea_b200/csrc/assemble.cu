// Numeric assembly fused with element stiffness evaluation and Dirichlet handling.
//
// Replaces the reference's sequential dense scatter `K[np.ix_(d, d)] += Ke` (cubebeam.py:82-90,
// fea.py:89-97, euler_bernoulli.py:42-49) and the constraint reduction (cubebeam.py:92-96).
//
// Owner-computes gather: one warp owns the d rows of one node.  It walks the node's incident
// elements in ascending element order -- the reference's own summation order -- evaluates the
// block row of each Ke on chip and writes every CSR value of its rows exactly once, coalesced.
// No atomics, no zero-fill pass, no Ke in HBM; results are bit-reproducible and independent of
// how rows are partitioned over GPUs.  HBM traffic is the algorithmic minimum
// (8 B per non-zero written + coordinates + connectivity + incidence lists).
#include <algorithm>

#include "hex8.cuh"

namespace fea {

__device__ __forceinline__ int find_slot(const int32_t* cols, int cnt, int key) {
  int lo = 0, hi = cnt - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cols[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Dirichlet transform of one assembled value (mode FEA_ASSEMBLE_ELIMINATED): constrained rows and
// columns become identity rows/columns, everything else is untouched.
__device__ __forceinline__ double eliminate(double v, const uint8_t* __restrict__ fixed, int64_t row, int64_t col) {
  if (fixed[row] | fixed[col]) return row == col ? 1.0 : 0.0;
  return v;
}

// ------------------------------------------------------------------------------------------
// hex8: warp per node.  Rounds of up to 4 incident elements:
//   phase A  lane = (t, gp): geometry of element t at Gauss point gp -> shared staging
//   phase B  lane = (t, b) : 3x3 block K_{a_t b} of element t (a_t = local index of the owner node)
//   phase C  for t = 0..3 in order: lanes of element t add their block into the row accumulator
// then the accumulator (3 x 3cnt doubles, exactly the node's slice of `values`) is stored.
// ------------------------------------------------------------------------------------------
constexpr int kAsmWarps = 4;

__global__ void __launch_bounds__(kAsmWarps * 32)
assemble_hex8_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ elements, int64_t n_nodes,
                     Hex8Material mat, const int32_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e,
                     const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx, int maxc,
                     const uint8_t* __restrict__ fixed, int mode, double* __restrict__ values,
                     double* __restrict__ dinv, int32_t* status) {
  extern __shared__ double s_dyn[];
  // layout: shape table | per warp: grad, detj, acc[9*maxc], cols[maxc] (ints, padded to doubles)
  double* s_tab = s_dyn;
  const int cols_doubles = (maxc + 1) / 2;
  const int per_warp = kGradDoubles + 32 + 9 * maxc + cols_doubles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* grad = s_dyn + kShapeTable + (size_t)warp * per_warp;
  double* detj = grad + kGradDoubles;
  double* acc = detj + 32;
  int32_t* cols = reinterpret_cast<int32_t*>(acc + 9 * maxc);
  hex8_fill_shape_table(s_tab);
  __syncthreads();

  for (int64_t node = (int64_t)blockIdx.x * kAsmWarps + warp; node < n_nodes;
       node += (int64_t)gridDim.x * kAsmWarps) {
    const int lo = node_rowptr[node];
    const int cnt = node_rowptr[node + 1] - lo;
    const int row_len = 3 * cnt;
    for (int q = lane; q < cnt; q += 32) cols[q] = node_colidx[lo + q];
    for (int q = lane; q < 9 * cnt; q += 32) acc[q] = 0.0;
    const int inc_lo = n2e_ptr[node];
    const int deg = n2e_ptr[node + 1] - inc_lo;
    __syncwarp();

    for (int round = 0; round < deg; round += 4) {
      const int t = lane >> 3;
      const bool active = round + t < deg;
      int e = 0, a_own = 0;
      if (active) {
        const int inc = n2e[inc_lo + round + t];
        e = inc >> 3;
        a_own = inc & 7;
      }
      if (active) {  // phase A
        const int gp = lane & 7;
        const double det = hex8_geometry(nodes, elements + (int64_t)e * 8, s_tab, gp, t, grad);
        detj[gp * 4 + t] = det;
        if (!(det > 0.0)) raise_status(status, FEA_ERR_JACOBIAN, e);
      }
      __syncwarp();
      double blk[3][3];
      int slot = 0;
      if (active) {  // phase B
        const int b = lane & 7;
        hex8_block(grad, detj, t, a_own, b, mat, blk);
        slot = find_slot(cols, cnt, elements[(int64_t)e * 8 + b]);
        FEA_ASSERT(slot >= 0 && slot < cnt && cols[slot] == elements[(int64_t)e * 8 + b]);
      }
      // phase C: element order; inside one element the 8 column nodes are distinct
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        if (active && t == tt) {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[r * row_len + 3 * slot + c] += blk[r][c];
        }
        __syncwarp();
      }
    }

    const int64_t base = 9 * (int64_t)lo;
    if (mode == FEA_ASSEMBLE_ELIMINATED && fixed != nullptr) {
      for (int q = lane; q < 9 * cnt; q += 32) {
        const int r = q / row_len, within = q - r * row_len;
        const int k = within / 3, c = within - 3 * k;
        values[base + q] = eliminate(acc[q], fixed, 3 * node + r, 3 * (int64_t)cols[k] + c);
      }
    } else {
      for (int q = lane; q < 9 * cnt; q += 32) values[base + q] = acc[q];
    }
    if (dinv != nullptr && lane < 3) {
      double di = 0.0;  // a node no element references keeps u = 0
      if (cnt > 0) {
        const int kd = find_slot(cols, cnt, (int)node);
        const double diag = acc[lane * row_len + 3 * kd + lane];
        const bool is_fixed = fixed != nullptr && fixed[3 * node + lane];
        di = is_fixed ? 0.0 : 1.0 / diag;
      }
      dinv[3 * node + lane] = di;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Generic slot-owner gather for cheap elements (beam, truss): warp per node, lane k owns the
// block column k of the node's rows; every lane scans the incident elements in order and adds
// the blocks whose column node is its own.
// ------------------------------------------------------------------------------------------
struct BeamOp {
  static constexpr int D = 2, NPE = 2;
  const double* EI;
  const double* length;
  __device__ void block(int64_t e, int a, int b, double out[2][2], int32_t*) const {
    const double L = length[e];
    const double c = EI[e] / cube_rn(L);  // euler_bernoulli.py:22-39
    const double s = 6.0 * L, f = 4.0 * (L * L), h = 2.0 * (L * L);
    const double m[4][4] = {{12.0, s, -12.0, s}, {s, f, -s, h}, {-12.0, -s, 12.0, -s}, {s, h, -s, f}};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
        out[r][cc] = __dmul_rn(c, m[2 * a + r][2 * b + cc]);  // rounded like the reference's Ke entry: the caller's
                                                              // accumulation must not fuse it into an FMA (cond ~ n^4)
  }
};

struct TrussOp {
  static constexpr int D = 3, NPE = 2;
  const double* nodes;
  const int32_t* members;
  const double* k;
  __device__ void block(int64_t e, int a, int b, double out[3][3], int32_t* status) const {
    const double* xa = nodes + 3 * (int64_t)members[2 * e];
    const double* xb = nodes + 3 * (int64_t)members[2 * e + 1];
    double c[3] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2]};
    const double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    if (!(len > 0.0)) raise_status(status, FEA_ERR_DEGENERATE, (int)e);
    c[0] /= len;
    c[1] /= len;
    c[2] /= len;
    const double ke = k[e];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const double v = ke * (c[r] * c[s]);
        out[r][s] = (a == b) ? v : -v;
      }
  }
};

template <class Op>
__global__ void __launch_bounds__(256)
assemble_slot_owner_kernel(Op op, const int32_t* __restrict__ elements, int64_t n_nodes,
                           const int32_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e,
                           const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                           const uint8_t* __restrict__ fixed, int mode, double* __restrict__ values,
                           double* __restrict__ dinv, int32_t* status) {
  constexpr int D = Op::D, NPE = Op::NPE;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t node = warp_global; node < n_nodes; node += n_warps) {
    const int lo = node_rowptr[node];
    const int cnt = node_rowptr[node + 1] - lo;
    const int row_len = D * cnt;
    const int inc_lo = n2e_ptr[node], inc_hi = n2e_ptr[node + 1];
    const int64_t base = (int64_t)D * D * lo;
    if (cnt == 0 && dinv != nullptr && lane < D) dinv[D * node + lane] = 0.0;
    for (int k = lane; k < cnt; k += 32) {
      const int col = node_colidx[lo + k];
      double acc[D][D];
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) acc[r][c] = 0.0;
      for (int i = inc_lo; i < inc_hi; ++i) {
        const int inc = n2e[i];
        const int e = inc / NPE, a = inc - e * NPE;
#pragma unroll
        for (int b = 0; b < NPE; ++b) {
          if (elements[(int64_t)e * NPE + b] == col) {
            double blk[D][D];
            op.block(e, a, b, blk, status);
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
              for (int c = 0; c < D; ++c) acc[r][c] += blk[r][c];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double v = acc[r][c];
          if (mode == FEA_ASSEMBLE_ELIMINATED && fixed != nullptr)
            v = eliminate(v, fixed, D * node + r, (int64_t)D * col + c);
          values[base + (int64_t)r * row_len + D * k + c] = v;
        }
      if (dinv != nullptr && col == node) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
          const bool is_fixed = fixed != nullptr && fixed[D * node + r];
          dinv[D * node + r] = is_fixed ? 0.0 : 1.0 / acc[r][r];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
jacobi_dinv_kernel(int64_t n_nodes, int d, const int32_t* __restrict__ node_rowptr,
                   const int32_t* __restrict__ node_colidx, const double* __restrict__ values,
                   const uint8_t* __restrict__ fixed, double* __restrict__ dinv) {
  const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  const int lo = node_rowptr[node];
  const int cnt = node_rowptr[node + 1] - lo;
  if (cnt == 0) {
    for (int r = 0; r < d; ++r) dinv[d * node + r] = 0.0;
    return;
  }
  const int kd = find_slot(node_colidx + lo, cnt, (int)node);
  const bool has_diag = node_colidx[lo + kd] == (int)node;
  for (int r = 0; r < d; ++r) {
    const bool is_fixed = fixed != nullptr && fixed[d * node + r];
    double v = 0.0;
    if (has_diag && !is_fixed) v = 1.0 / values[(int64_t)d * d * lo + (int64_t)r * d * cnt + d * kd + r];
    dinv[d * node + r] = v;
  }
}

}  // namespace fea

using namespace fea;

extern "C" int fea_assemble_hex8(const double* nodes, const int32_t* elements, int64_t n_elem, int64_t n_nodes,
                                 double E, double nu, const int32_t* n2e_ptr, const int32_t* n2e,
                                 const int32_t* node_rowptr, const int32_t* node_colidx, int32_t max_coupled,
                                 const uint8_t* fixed, int32_t mode, double* values, double* dinv,
                                 int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!nodes || !elements || !n2e_ptr || !n2e || !node_rowptr || !node_colidx || !values) return FEA_ERR_INVALID;
  if (n_nodes <= 0 || n_elem < 0 || max_coupled < 1) return FEA_ERR_INVALID;
  const int maxc = max_coupled;
  const int per_warp = kGradDoubles + 32 + 9 * maxc + (maxc + 1) / 2;
  const size_t smem = sizeof(double) * (kShapeTable + (size_t)kAsmWarps * per_warp);
  if (smem > 200 * 1024) return FEA_ERR_INVALID;  // valence too high for the on-chip accumulator
  if (smem > 48 * 1024) {
    FEA_TRY(check(cudaFuncSetAttribute(assemble_hex8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  }
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n_nodes, kAsmWarps), 148LL * 64);
  assemble_hex8_kernel<<<blocks, kAsmWarps * 32, smem, stream>>>(nodes, elements, n_nodes, hex8_material(E, nu), n2e_ptr,
                                                                 n2e, node_rowptr, node_colidx, maxc, fixed, mode, values,
                                                                 dinv, status);
  return check_launch();
}

template <class Op>
static int launch_slot_owner(Op op, const int32_t* elements, int64_t n_nodes, const int32_t* n2e_ptr,
                             const int32_t* n2e, const int32_t* node_rowptr, const int32_t* node_colidx,
                             const uint8_t* fixed, int mode, double* values, double* dinv, int32_t* status,
                             cudaStream_t stream) {
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n_nodes, 8), 148LL * 32);
  assemble_slot_owner_kernel<Op><<<blocks, 256, 0, stream>>>(op, elements, n_nodes, n2e_ptr, n2e, node_rowptr,
                                                             node_colidx, fixed, mode, values, dinv, status);
  return check_launch();
}

extern "C" int fea_assemble_beam(const double* EI, const double* length, const int32_t* elements, int64_t n_elem,
                                 int64_t n_nodes, const int32_t* n2e_ptr, const int32_t* n2e,
                                 const int32_t* node_rowptr, const int32_t* node_colidx, const uint8_t* fixed,
                                 int32_t mode, double* values, double* dinv, void* stream_) {
  (void)n_elem;
  if (!EI || !length || !elements || !n2e_ptr || !n2e || !node_rowptr || !node_colidx || !values || n_nodes <= 0)
    return FEA_ERR_INVALID;
  BeamOp op{EI, length};
  return launch_slot_owner(op, elements, n_nodes, n2e_ptr, n2e, node_rowptr, node_colidx, fixed, mode, values, dinv,
                           nullptr, static_cast<cudaStream_t>(stream_));
}

extern "C" int fea_assemble_truss(const double* nodes, const int32_t* members, const double* k, int64_t n_elem,
                                  int64_t n_nodes, const int32_t* n2e_ptr, const int32_t* n2e,
                                  const int32_t* node_rowptr, const int32_t* node_colidx, const uint8_t* fixed,
                                  int32_t mode, double* values, double* dinv, int32_t* status, void* stream_) {
  (void)n_elem;
  if (!nodes || !members || !k || !n2e_ptr || !n2e || !node_rowptr || !node_colidx || !values || n_nodes <= 0)
    return FEA_ERR_INVALID;
  TrussOp op{nodes, members, k};
  return launch_slot_owner(op, members, n_nodes, n2e_ptr, n2e, node_rowptr, node_colidx, fixed, mode, values, dinv,
                           status, static_cast<cudaStream_t>(stream_));
}

extern "C" int fea_jacobi_dinv(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                               const int32_t* node_colidx, const double* values, const uint8_t* fixed, double* dinv,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !values || !dinv || n_nodes <= 0 || dof_per_node < 1) return FEA_ERR_INVALID;
  jacobi_dinv_kernel<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, stream>>>(n_nodes, dof_per_node, node_rowptr,
                                                                          node_colidx, values, fixed, dinv);
  return check_launch();
}
