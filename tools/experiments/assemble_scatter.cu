// EXPERIMENT, not part of libfea_b200.so: the textbook GPU assembly (element-parallel, FP64 atomics into
// zero-filled CSR values) that north_star sketches, kept as the comparison point for the owner-computes
// gather of fea_b200/csrc/assemble.cu.  Measured at 400x80x80 on B200 (round 1): 9.8-12.2 ms incl. the
// 5 GB zero-fill (up to 100 ms when L2 atomics contend) against 11.6 ms for the gather, which is
// deterministic and needs no zero-fill -- so the gather is the product path.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I fea_b200/csrc -c tools/experiments/assemble_scatter.cu
#include <algorithm>

#include "hex8.cuh"

namespace fea {
constexpr int kAsmWarps = 4;
__device__ __forceinline__ int find_slot(const int32_t* cols, int cnt, int key) {
  int lo = 0, hi = cnt - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cols[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}


// ------------------------------------------------------------------------------------------
// Scatter-add alternative for hex8 (the textbook GPU assembly): warp per 4 elements, every lane
// (t, b) pushes the 8 blocks K_ab of its element with FP64 atomics.  Kept for comparison.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAsmWarps * 32)
assemble_hex8_scatter_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ elements, int64_t n_elem,
                             Hex8Material mat, const int32_t* __restrict__ node_rowptr,
                             const int32_t* __restrict__ node_colidx, double* __restrict__ values,
                             int32_t* status) {
  __shared__ double s_tab[kShapeTable];
  __shared__ double s_grad[kAsmWarps][kGradDoubles];
  __shared__ double s_detj[kAsmWarps][32];
  hex8_fill_shape_table(s_tab);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* grad = s_grad[warp];
  double* detj = s_detj[warp];
  const int64_t n_groups = (n_elem + 3) / 4;
  for (int64_t g = (int64_t)blockIdx.x * kAsmWarps + warp; g < n_groups; g += (int64_t)gridDim.x * kAsmWarps) {
    const int t = lane >> 3;
    const int64_t e = g * 4 + t;
    if (e < n_elem) {
      const int gp = lane & 7;
      const double det = hex8_geometry(nodes, elements + e * 8, s_tab, gp, t, grad);
      detj[gp * 4 + t] = det;
      if (!(det > 0.0)) raise_status(status, FEA_ERR_JACOBIAN, (int)e);
    }
    __syncwarp();
    if (e < n_elem) {
      const int b = lane & 7;
      const int col = elements[e * 8 + b];
#pragma unroll 1
      for (int a = 0; a < 8; ++a) {
        const int row_node = elements[e * 8 + a];
        const int lo = node_rowptr[row_node];
        const int cnt = node_rowptr[row_node + 1] - lo;
        const int slot = find_slot(node_colidx + lo, cnt, col);
        double blk[3][3];
        hex8_block(grad, detj, t, a, b, mat, blk);
        double* dst = values + 9 * (int64_t)lo + 3 * slot;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) atomicAdd(dst + (int64_t)r * 3 * cnt + c, blk[r][c]);
      }
    }
    __syncwarp();
  }
}
}  // namespace fea

using namespace fea;

extern "C" int fea_assemble_hex8_scatter(const double* nodes, const int32_t* elements, int64_t n_elem, double E,
                                         double nu, const int32_t* node_rowptr, const int32_t* node_colidx,
                                         double* values, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!nodes || !elements || !node_rowptr || !node_colidx || !values || n_elem < 0) return FEA_ERR_INVALID;
  if (n_elem == 0) return FEA_OK;
  const int64_t groups = ceil_div(n_elem, 4);
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(groups, kAsmWarps), 148LL * 16);
  assemble_hex8_scatter_kernel<<<blocks, kAsmWarps * 32, 0, stream>>>(nodes, elements, n_elem, hex8_material(E, nu),
                                                                      node_rowptr, node_colidx, values, status);
  return check_launch();
}

