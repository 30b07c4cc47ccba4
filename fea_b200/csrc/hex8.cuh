// hex8 element mathematics shared by the Ke and assembly kernels.
// Follows utils.py:127-239 (hexahedral_stiffness_matrix): 2x2x2 Gauss rule with the points in
// the reference's loop order (xi outer, eta, zeta inner; utils.py:200-204), trilinear shape
// derivatives divided by 8 (utils.py:159-197), J = dN X, dN_dx = J^-1 dN (utils.py:210-221),
// Ke += w (B^T C B) detJ (utils.py:224-237) with the B/C sparsity multiplied out per Gauss point:
//   K_ab[r][r] = C11 ga_r gb_r + C44 (ga_s gb_s + ga_t gb_t)       s,t = the other two axes
//   K_ab[r][c] = C12 ga_r gb_c + C44 ga_c gb_r                     r != c
// where g_a = grad N_a and C11, C12, C44 are the entries of the 6x6 C of utils.py:144-153
// (hex8_block sums detJ ga_r gb_c over the Gauss points first and applies C once).
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace fea {

struct Hex8Material {
  double c11, c12, c44;
};

__host__ __device__ inline Hex8Material hex8_material(double E, double nu) {
  const double c = E / ((1.0 + nu) * (1.0 - 2.0 * nu));  // utils.py:144
  Hex8Material m;
  m.c11 = c * (1.0 - nu);
  m.c12 = c * nu;
  m.c44 = c * ((1.0 - 2.0 * nu) / 2.0);
  return m;
}

// 1/sqrt(3) as numpy computes it (utils.py:140): 1 / np.sqrt(3) == 0.5773502691896258
#define FEA_GAUSS 0.5773502691896258

// Local node signs, bottom face CCW then top face CCW (utils.py:159-197, 351-353), bit-packed:
// bit a of kSignX is set when node a has xi = +1, etc.
constexpr unsigned kSignX = 0x66;  // nodes 1,2,5,6
constexpr unsigned kSignY = 0xCC;  // nodes 2,3,6,7
constexpr unsigned kSignZ = 0xF0;  // nodes 4..7

__host__ __device__ inline double sign_of(unsigned mask, int a) { return ((mask >> a) & 1u) ? 1.0 : -1.0; }

// dN_a/d(xi,eta,zeta) at Gauss point gp (0..7; bit 2 = xi, bit 1 = eta, bit 0 = zeta index).
__host__ __device__ inline void hex8_shape_derivative(int gp, int a, double out[3]) {
  const double xi = (gp & 4) ? FEA_GAUSS : -FEA_GAUSS;
  const double eta = (gp & 2) ? FEA_GAUSS : -FEA_GAUSS;
  const double zeta = (gp & 1) ? FEA_GAUSS : -FEA_GAUSS;
  const double sx = sign_of(kSignX, a), sy = sign_of(kSignY, a), sz = sign_of(kSignZ, a);
  const double fx = 1.0 + sx * xi, fy = 1.0 + sy * eta, fz = 1.0 + sz * zeta;
  out[0] = sx * fy * fz / 8.0;
  out[1] = sy * fx * fz / 8.0;
  out[2] = sz * fx * fy / 8.0;
}

// Shared-memory shape table: tab[(gp*3 + r)*8 + a] = dN_a/dxi_r at Gauss point gp (192 doubles).
constexpr int kShapeTable = 8 * 3 * 8;

__device__ inline void hex8_fill_shape_table(double* tab) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    const int gp = i >> 3, a = i & 7;
    double d[3];
    hex8_shape_derivative(gp, a, d);
    tab[(gp * 3 + 0) * 8 + a] = d[0];
    tab[(gp * 3 + 1) * 8 + a] = d[1];
    tab[(gp * 3 + 2) * 8 + a] = d[2];
  }
}

// Per-warp staging of global shape-function gradients for up to 4 elements x 8 Gauss points:
//   grad[(r*8 + gp)*kGradStride + t*8 + a]   (padded stride: conflict-free writes and reads)
//   detj[gp*4 + t]
constexpr int kGradStride = 33;
constexpr int kGradDoubles = 24 * kGradStride;

// Geometry of one (element, Gauss point) in two halves, so that the first can run once per element
// (hex8_gauss_geometry_kernel, assemble.cu) and the second once per incident node:
//   hex8_inverse_jacobian   J = dN X (utils.py:210), detJ (:211), J^-1 (:218)          -> returns detJ
//   hex8_gradients          dN_dx = J^-1 dN for the 8 nodes (:221) -> staging area, slot t
__device__ __forceinline__ double hex8_inverse_jacobian(const double* __restrict__ nodes,
                                                        const int32_t* __restrict__ conn,
                                                        const double* __restrict__ tab, int gp, double I[3][3]) {
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  const double* d0 = tab + (gp * 3 + 0) * 8;
  const double* d1 = tab + (gp * 3 + 1) * 8;
  const double* d2 = tab + (gp * 3 + 2) * 8;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const double* x = nodes + 3 * (int64_t)conn[a];
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    const double a0 = d0[a], a1 = d1[a], a2 = d2[a];
    J[0][0] += a0 * x0; J[0][1] += a0 * x1; J[0][2] += a0 * x2;
    J[1][0] += a1 * x0; J[1][1] += a1 * x1; J[1][2] += a1 * x2;
    J[2][0] += a2 * x0; J[2][1] += a2 * x1; J[2][2] += a2 * x2;
  }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double inv_det = 1.0 / det;
  I[0][0] = c00 * inv_det;
  I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv_det;
  I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv_det;
  I[1][0] = c01 * inv_det;
  I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv_det;
  I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv_det;
  I[2][0] = c02 * inv_det;
  I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv_det;
  I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv_det;
  return det;
}

__device__ __forceinline__ void hex8_gradients(const double I[3][3], const double* __restrict__ tab, int gp, int t,
                                               double* __restrict__ grad) {
  const double* d0 = tab + (gp * 3 + 0) * 8;
  const double* d1 = tab + (gp * 3 + 1) * 8;
  const double* d2 = tab + (gp * 3 + 2) * 8;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const double a0 = d0[a], a1 = d1[a], a2 = d2[a];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      grad[(r * 8 + gp) * kGradStride + t * 8 + a] = I[r][0] * a0 + I[r][1] * a1 + I[r][2] * a2;
  }
}

// Both halves: returns detJ; writes the gradients to the staging area for slot t.
__device__ __forceinline__ double hex8_geometry(const double* __restrict__ nodes, const int32_t* __restrict__ conn,
                                                const double* __restrict__ tab, int gp, int t,
                                                double* __restrict__ grad) {
  double I[3][3];
  const double det = hex8_inverse_jacobian(nodes, conn, tab, gp, I);
  hex8_gradients(I, tab, gp, t, grad);
  return det;
}

// 3x3 block K_ab of one element (slot t of the staging area).  With S[r][c] = sum_gp detJ ga_r gb_c
// (Gauss points in the reference's order, weights 1) the material constants factor out of the
// quadrature:  K_ab[r][r] = C11 S_rr + C44 (S_ss + S_tt),  K_ab[r][c] = C12 S_rc + C44 S_cr,
// 12 flops per Gauss point instead of 45 for the multiplied-out form (round 1 measured both: 5 % apart,
// the kernel is not flop-bound; the multiplied-out variant is gone).
__device__ __forceinline__ void hex8_block(const double* __restrict__ grad, const double* __restrict__ detj, int t,
                                           int a, int b, const Hex8Material& m, double blk[3][3]) {
  double S[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) S[r][c] = 0.0;
#pragma unroll
  for (int gp = 0; gp < 8; ++gp) {
    const double w = detj[gp * 4 + t];
    double wa[3], gb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      wa[r] = w * grad[(r * 8 + gp) * kGradStride + t * 8 + a];
      gb[r] = grad[(r * 8 + gp) * kGradStride + t * 8 + b];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) S[r][c] = fma(wa[r], gb[c], S[r][c]);
  }
  blk[0][0] = m.c11 * S[0][0] + m.c44 * (S[1][1] + S[2][2]);
  blk[1][1] = m.c11 * S[1][1] + m.c44 * (S[0][0] + S[2][2]);
  blk[2][2] = m.c11 * S[2][2] + m.c44 * (S[0][0] + S[1][1]);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (r != c) blk[r][c] = m.c12 * S[r][c] + m.c44 * S[c][r];
}

}  // namespace fea
