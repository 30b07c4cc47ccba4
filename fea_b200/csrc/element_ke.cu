// Batched FP64 element stiffness matrices, materialised in HBM.
// hex8: utils.py:127-239; beam: euler_bernoulli.py:22-39; truss: tangent of truss.py:78-92.
// These back the element-level API (`utils.hexahedral_stiffness_matrix`) and the Ke microbench;
// the assembly path (assemble.cu) evaluates the same device functions without storing Ke.
#include <algorithm>

#include "hex8.cuh"

namespace fea {

// One warp processes 4 elements at a time.
//   phase A: lane = (t, gp): Jacobian, detJ, global gradients of element t at Gauss point gp
//            -> per-warp shared staging (shape-function table staged once per CTA);
//   phase B: lane = (t, b): the eight 3x3 blocks K_ab, a = 0..7, written to ke[e][3a+r][3b+c].
constexpr int kKeWarps = 4;

__global__ void __launch_bounds__(kKeWarps * 32) ke_hex8_kernel(const double* __restrict__ nodes,
                                                                const int32_t* __restrict__ elements,
                                                                int64_t n_elem, Hex8Material mat,
                                                                double* __restrict__ ke, int32_t* status) {
  __shared__ double s_tab[kShapeTable];
  __shared__ double s_grad[kKeWarps][kGradDoubles];
  __shared__ double s_detj[kKeWarps][32];
  __shared__ double s_stage[kKeWarps][4 * 72];
  hex8_fill_shape_table(s_tab);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* grad = s_grad[warp];
  double* detj = s_detj[warp];
  double* stage = s_stage[warp];
  const int64_t n_groups = (n_elem + 3) / 4;
  for (int64_t g = (int64_t)blockIdx.x * kKeWarps + warp; g < n_groups; g += (int64_t)gridDim.x * kKeWarps) {
    {  // phase A
      const int t = lane >> 3, gp = lane & 7;
      const int64_t e = g * 4 + t;
      if (e < n_elem) {
        const double det = hex8_geometry(nodes, elements + e * 8, s_tab, gp, t, grad);
        detj[gp * 4 + t] = det;
        if (!(det > 0.0)) raise_status(status, FEA_ERR_JACOBIAN, (int)e);
      }
    }
    __syncwarp();
    {  // phase B: block row a of the 4 elements = 3 full rows of each Ke = 72 contiguous doubles per
       // element; staged through shared memory so that every store instruction of the warp writes
       // 256 contiguous bytes (direct stores touched 24 sectors per 256 B of payload).
      const int t = lane >> 3, b = lane & 7;
      const int64_t e = g * 4 + t;
      const int64_t e0 = g * 4;
      const int n_here = (int)min((int64_t)4, n_elem - e0);
#pragma unroll 1
      for (int a = 0; a < 8; ++a) {
        if (e < n_elem) {
          double blk[3][3];
          hex8_block(grad, detj, t, a, b, mat, blk);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) stage[t * 72 + r * 24 + 3 * b + c] = blk[r][c];
        }
        __syncwarp();
        for (int q = lane; q < 72 * n_here; q += 32) {
          const int tt = q / 72, w = q - tt * 72;
          __stcs(ke + (e0 + tt) * 576 + a * 72 + w, stage[q]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }
}

__global__ void ke_beam_kernel(const double* __restrict__ EI, const double* __restrict__ length, int64_t n_elem,
                               double* __restrict__ ke) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const double L = length[e];
  const double c = EI[e] / cube_rn(L);  // euler_bernoulli.py:22
  const double s = 6.0 * L, f = 4.0 * (L * L), h = 2.0 * (L * L);
  const double m[16] = {12.0, s, -12.0, s, s, f, -s, h, -12.0, -s, 12.0, -s, s, h, -s, f};
  double* out = ke + e * 16;
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = c * m[i];
}

__global__ void ke_truss_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ members,
                                const double* __restrict__ k, int64_t n_elem, double* __restrict__ ke,
                                int32_t* status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const double* xa = nodes + 3 * (int64_t)members[2 * e];
  const double* xb = nodes + 3 * (int64_t)members[2 * e + 1];
  double c[3] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2]};
  const double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  if (!(len > 0.0)) raise_status(status, FEA_ERR_DEGENERATE, (int)e);
  c[0] /= len;
  c[1] /= len;
  c[2] /= len;
  const double ke_ = k[e];
  double* out = ke + e * 36;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const double v = ke_ * (c[r] * c[s]);
      out[r * 6 + s] = v;
      out[(r + 3) * 6 + s + 3] = v;
      out[r * 6 + s + 3] = -v;
      out[(r + 3) * 6 + s] = -v;
    }
}

}  // namespace fea

using namespace fea;

extern "C" int fea_ke_hex8(const double* nodes, const int32_t* elements, int64_t n_elem, double E, double nu,
                           double* ke, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!nodes || !elements || !ke || n_elem < 0) return FEA_ERR_INVALID;
  if (n_elem == 0) return FEA_OK;
  const int64_t groups = ceil_div(n_elem, 4);
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(groups, kKeWarps), 148LL * 16);
  ke_hex8_kernel<<<blocks, kKeWarps * 32, 0, stream>>>(nodes, elements, n_elem, hex8_material(E, nu), ke, status);
  return check_launch();
}

extern "C" int fea_ke_beam(const double* EI, const double* length, int64_t n_elem, double* ke, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!EI || !length || !ke || n_elem < 0) return FEA_ERR_INVALID;
  if (n_elem == 0) return FEA_OK;
  ke_beam_kernel<<<(unsigned)ceil_div(n_elem, 256), 256, 0, stream>>>(EI, length, n_elem, ke);
  return check_launch();
}

extern "C" int fea_ke_truss(const double* nodes, const int32_t* members, const double* k, int64_t n_elem,
                            double* ke, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!nodes || !members || !k || !ke || n_elem < 0) return FEA_ERR_INVALID;
  if (n_elem == 0) return FEA_OK;
  ke_truss_kernel<<<(unsigned)ceil_div(n_elem, 256), 256, 0, stream>>>(nodes, members, k, n_elem, ke, status);
  return check_launch();
}
