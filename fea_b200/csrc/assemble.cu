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
constexpr int kGaussDoubles = 10;  // J^-1 (row-major) and detJ of one (element, Gauss point)

// Thread per (element, Gauss point): the half of the geometry that does not depend on the node a block row belongs
// to.  Round 1 recomputed it in each of the 8 nodes of an element (a third of the Gauss-point kernel's FP64 work).
__global__ void __launch_bounds__(256)
hex8_gauss_geometry_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ elements, int64_t n_elem,
                           double* __restrict__ gauss, int32_t* status) {
  __shared__ double s_tab[kShapeTable];
  hex8_fill_shape_table(s_tab);
  __syncthreads();
  const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = id >> 3;
  const int gp = (int)(id & 7);
  if (e >= n_elem) return;
  double I[3][3];
  const double det = hex8_inverse_jacobian(nodes, elements + e * 8, s_tab, gp, I);
  if (!(det > 0.0)) raise_status(status, FEA_ERR_JACOBIAN, (int)e);
  double2* out = reinterpret_cast<double2*>(gauss + id * kGaussDoubles);
  out[0] = make_double2(I[0][0], I[0][1]);
  out[1] = make_double2(I[0][2], I[1][0]);
  out[2] = make_double2(I[1][1], I[1][2]);
  out[3] = make_double2(I[2][0], I[2][1]);
  out[4] = make_double2(I[2][2], det);
}

__global__ void __launch_bounds__(kAsmWarps * 32, 5)
assemble_hex8_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ elements, int64_t n_nodes,
                     Hex8Material mat, const int32_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e,
                     const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx, int maxc,
                     const uint8_t* __restrict__ fixed, int mode, double* __restrict__ values,
                     double* __restrict__ dinv, const uint8_t* __restrict__ todo, const double* __restrict__ gauss,
                     int32_t* status) {
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

  // a warp takes 8 consecutive nodes per trip and works through those still to do (all of them without
  // `todo`; with it, the nodes assemble_hex8_affine_kernel left: one coalesced flag load per 8 nodes)
  for (int64_t base = ((int64_t)blockIdx.x * kAsmWarps + warp) * 8; base < n_nodes;
       base += (int64_t)gridDim.x * kAsmWarps * 8) {
    const bool mine = lane < 8 && base + lane < n_nodes && (todo == nullptr || todo[base + lane] != 0);
    unsigned pending = __ballot_sync(kFull, mine);
  while (pending != 0) {
    const int64_t node = base + (__ffs(pending) - 1);
    pending &= pending - 1;
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
      if (active) {  // phase A: J^-1 and detJ come from hex8_gauss_geometry_kernel (once per element), the gradients here
        const int gp = lane & 7;
        const double2* rec = reinterpret_cast<const double2*>(gauss + ((int64_t)e * 8 + gp) * kGaussDoubles);
        const double2 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3), q4 = __ldg(rec + 4);
        const double I[3][3] = {{q0.x, q0.y, q1.x}, {q1.y, q2.x, q2.y}, {q3.x, q3.y, q4.x}};
        hex8_gradients(I, s_tab, gp, t, grad);
        detj[gp * 4 + t] = q4.y;
        if (!(q4.y > 0.0)) raise_status(status, FEA_ERR_JACOBIAN, e);
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
}

// ------------------------------------------------------------------------------------------
// hex8, exactly affine elements (parallelepipeds: every element of an extruded grid).  The trilinear
// map x(xi,eta,zeta) = sum_m c_m * monomial_m has its 8 coefficient vectors c_m = the Walsh-Hadamard
// transform of the corner coordinates over the three sign bits of the local node.  When the four
// higher coefficients (xi.eta, xi.zeta, eta.zeta, xi.eta.zeta) are EXACTLY zero the Jacobian is the
// same at all Gauss points and the quadrature of utils.py:200-237 collapses to
//     S_ab = sum_gp detJ g_a g_b^T = 1/(8 detJc) * adj(Jc) M_ab adj(Jc)^T,      Jc = 8 J = [c_xi; c_eta; c_zeta]
// with the constant table M_ab = sum_gp dN_a dN_b^T: 54 FMA per 3x3 block and no per-Gauss-point geometry
// at all (the general kernel spends 2/3 of its FP64 instructions there, once per incident NODE).
//
// Pass 0 (hex8_affine_geometry_kernel, thread per element): adj(Jc) and 1/(8 detJc) -- 80 bytes per element in
// a stream-ordered scratch array; scale = -1 marks an element that is not exactly affine.  Geometry is
// evaluated ONCE per element here; the owner-computes gather below only reads it (L2-resident reuse).
// Pass 1 (assemble_hex8_affine_kernel): a warp owns 4 consecutive nodes; lane = (node q, column node j).
// Round k handles the k-th incident element of each of the 4 nodes; lane j evaluates the block K_{a_own, j}.
// Rounds are ascending element order per node -- the reference's summation order (cubebeam.py:82-90) -- and
// inside a round the 8 column nodes of an element are distinct, so accumulation needs no serialisation and no
// atomics.  A node with any non-affine incident element is left to assemble_hex8_kernel (flag in `todo`): the
// choice depends on the node's own elements only, so rows stay independent of the GPU partition.
// ------------------------------------------------------------------------------------------
constexpr int kAffWarps = 4;
constexpr int kAffNodes = 4;   // nodes per warp
constexpr int kMtab = 64 * 9;  // M_ab[k][l], a-major
constexpr int kGeomDoubles = 10;

__global__ void __launch_bounds__(256)
hex8_affine_geometry_kernel(const double* __restrict__ nodes, const int32_t* __restrict__ elements, int64_t n_elem,
                            double* __restrict__ geom, unsigned* __restrict__ n_general, int32_t* status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  // c[m] = corner whose (x, y, z) sign bits are the bits of m (hex8.cuh: kSignX/Y/Z)
  double c[8][3];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int corner = (m & 4) | ((m & 3) ^ ((m >> 1) & 1));
    const double* x = nodes + 3 * (int64_t)elements[e * 8 + corner];
    c[m][0] = x[0];
    c[m][1] = x[1];
    c[m][2] = x[2];
  }
#pragma unroll
  for (int bit = 1; bit < 8; bit <<= 1)
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (!(m & bit)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double lo = c[m][k], hi = c[m | bit][k];
          c[m][k] = lo + hi;
          c[m | bit][k] = hi - lo;
        }
      }
  bool affine = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) affine = affine && c[3][k] == 0.0 && c[5][k] == 0.0 && c[6][k] == 0.0 && c[7][k] == 0.0;
  const double(&J0)[3] = c[1];
  const double(&J1)[3] = c[2];
  const double(&J2)[3] = c[4];
  double A[9];  // adjugate of Jc, row-major
  A[0] = J1[1] * J2[2] - J1[2] * J2[1];
  A[3] = J1[2] * J2[0] - J1[0] * J2[2];
  A[6] = J1[0] * J2[1] - J1[1] * J2[0];
  A[1] = J0[2] * J2[1] - J0[1] * J2[2];
  A[4] = J0[0] * J2[2] - J0[2] * J2[0];
  A[7] = J0[1] * J2[0] - J0[0] * J2[1];
  A[2] = J0[1] * J1[2] - J0[2] * J1[1];
  A[5] = J0[2] * J1[0] - J0[0] * J1[2];
  A[8] = J0[0] * J1[1] - J0[1] * J1[0];
  const double det = J0[0] * A[0] + J0[1] * A[3] + J0[2] * A[6];
  double scale = -1.0;
  if (affine) {
    if (det > 0.0) scale = 0.125 / det;
    else raise_status(status, FEA_ERR_JACOBIAN, (int)e);
  }
  double* out = geom + e * kGeomDoubles;
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = A[k];
  out[9] = scale;
  // elements the closed form cannot take (not affine, or inverted): one atomic per warp
  const unsigned active = __activemask();  // the threads of the warp that have an element (and are here together)
  const unsigned votes = __ballot_sync(active, !(scale > 0.0));
  if (votes != 0 && (int)(threadIdx.x & 31) == __ffs(active) - 1) atomicAdd(n_general, (unsigned)__popc(votes));
}

// Position of `key` in a sorted list padded to 32 entries with INT32_MAX: five loads, no branches.
__device__ __forceinline__ int find_slot32(const int32_t* cols, int key) {
  int pos = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1)
    if (cols[pos + step - 1] < key) pos += step;
  return pos;
}

// ints per node for the column list: >= 32 (find_slot32 reads the padding), odd stride in banks so that the four
// groups of a warp, which probe the same positions of their own lists, land on different banks; the region sits at
// the end of the shared-memory block, so it needs no 8-byte alignment per node
__host__ __device__ constexpr int aff_cols_ints(int maxc) { return (maxc < 32 ? 32 : maxc) | 1; }

struct AffGeom {
  double2 a01, a23, a45, a67, a8s;  // adj(Jc) row-major, then 1/(8 detJc)
};
__device__ __forceinline__ AffGeom load_geom(const double* __restrict__ geom, int e) {
  const double2* g = reinterpret_cast<const double2*>(geom + (int64_t)e * kGeomDoubles);
  AffGeom r;
  r.a01 = __ldg(g);
  r.a23 = __ldg(g + 1);
  r.a45 = __ldg(g + 2);
  r.a67 = __ldg(g + 3);
  r.a8s = __ldg(g + 4);
  return r;
}

__global__ void __launch_bounds__(kAffWarps * 32, 5)
assemble_hex8_affine_kernel(const double* __restrict__ geom, const int32_t* __restrict__ elements, int64_t n_nodes,
                            Hex8Material mat, const int32_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e,
                            const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                            int maxc, const uint8_t* __restrict__ fixed, int mode, double* __restrict__ values,
                            double* __restrict__ dinv, uint8_t* __restrict__ todo) {
  extern __shared__ __align__(16) double s_dyn[];
  // layout: M table | per (warp, node): acc[9*maxc] ... | cols[aff_cols_ints] ...
  double* s_m = s_dyn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane >> 3, j = lane & 7;
  const int slot_id = warp * kAffNodes + q;
  double* acc = s_dyn + kMtab + (size_t)slot_id * 9 * maxc;
  int32_t* cols = reinterpret_cast<int32_t*>(s_dyn + kMtab + (size_t)kAffWarps * kAffNodes * 9 * maxc) +
                  (size_t)slot_id * aff_cols_ints(maxc);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    const int a = i >> 3, b = i & 7;
    double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int gp = 0; gp < 8; ++gp) {
      double da[3], db[3];
      hex8_shape_derivative(gp, a, da);
      hex8_shape_derivative(gp, b, db);
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) m[3 * k + l] = fma(da[k], db[l], m[3 * k + l]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) s_m[i * 9 + k] = m[k];
  }
  __syncthreads();
  const bool small_rows = maxc <= 32;

  const int64_t n_groups = (n_nodes + kAffNodes - 1) / kAffNodes;
  for (int64_t g = (int64_t)blockIdx.x * kAffWarps + warp; g < n_groups; g += (int64_t)gridDim.x * kAffWarps) {
    const int64_t node = g * kAffNodes + q;
    const bool valid = node < n_nodes;
    int lo = 0, cnt = 0, inc_lo = 0, deg = 0;
    if (valid) {
      lo = node_rowptr[node];
      cnt = node_rowptr[node + 1] - lo;
      inc_lo = n2e_ptr[node];
      deg = n2e_ptr[node + 1] - inc_lo;
    }
    const int row_len = 3 * cnt;
    for (int i = j; i < cnt; i += 8) cols[i] = node_colidx[lo + i];
    if (small_rows)
      for (int i = cnt + j; i < 32; i += 8) cols[i] = INT32_MAX;
    for (int i = j; i < 9 * cnt; i += 8) acc[i] = 0.0;
    const int max_deg = warp_max_i(deg);
    bool dead = false;  // uniform inside a group
    __syncwarp();

    // Rounds in chunks of 8 (one incidence entry per lane of the group); the geometry record and the column node
    // of round k+1 are in flight during the arithmetic of round k.  Lanes without work read element 0.
    for (int k0 = 0; k0 < max_deg; k0 += 8) {
      const int my_inc = k0 + j < deg ? n2e[inc_lo + k0 + j] : -1;
      int inc_n = __shfl_sync(kFull, my_inc, 8 * q);
      AffGeom g_n = load_geom(geom, max(inc_n, 0) >> 3);
      int col_n = elements[(int64_t)(max(inc_n, 0) >> 3) * 8 + j];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (k0 + kk >= max_deg) break;  // warp-uniform
        const int inc = inc_n, col = col_n;
        const AffGeom G = g_n;
        if (kk + 1 < 8) {
          inc_n = __shfl_sync(kFull, my_inc, 8 * q + kk + 1);
          g_n = load_geom(geom, max(inc_n, 0) >> 3);
          col_n = elements[(int64_t)(max(inc_n, 0) >> 3) * 8 + j];
        }
        const double scale = G.a8s.y;
        if (inc >= 0 && !(scale > 0.0)) dead = true;  // same record for the whole group
        if (inc >= 0 && !dead) {
          const int slot = small_rows ? find_slot32(cols, col) : find_slot(cols, cnt, col);
          FEA_ASSERT(slot >= 0 && slot < cnt && cols[slot] == col);
          const double A[3][3] = {{G.a01.x, G.a01.y, G.a23.x}, {G.a23.y, G.a45.x, G.a45.y}, {G.a67.x, G.a67.y, G.a8s.x}};
          const double* M = s_m + ((inc & 7) * 8 + j) * 9;
          double W[3][3], S[3][3];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int l = 0; l < 3; ++l) W[r][l] = A[r][0] * M[l] + A[r][1] * M[3 + l] + A[r][2] * M[6 + l];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) S[r][c] = W[r][0] * A[c][0] + W[r][1] * A[c][1] + W[r][2] * A[c][2];
          const double c11 = mat.c11 * scale, c12 = mat.c12 * scale, c44 = mat.c44 * scale;
          double* dst = acc + 3 * slot;
          dst[0] += c11 * S[0][0] + c44 * (S[1][1] + S[2][2]);
          dst[row_len + 1] += c11 * S[1][1] + c44 * (S[0][0] + S[2][2]);
          dst[2 * row_len + 2] += c11 * S[2][2] + c44 * (S[0][0] + S[1][1]);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
              if (r != c) dst[r * row_len + c] += c12 * S[r][c] + c44 * S[c][r];
        }
        __syncwarp();
      }
    }

    if (valid && j == 0) todo[node] = dead ? 1 : 0;
    if (valid && !dead) {
      const int64_t base = 9 * (int64_t)lo;
      if (mode == FEA_ASSEMBLE_ELIMINATED && fixed != nullptr) {
        for (int i = j; i < 9 * cnt; i += 8) {
          const int r = i / row_len, within = i - r * row_len;
          const int kk = within / 3, c = within - 3 * kk;
          values[base + i] = eliminate(acc[i], fixed, 3 * node + r, 3 * (int64_t)cols[kk] + c);
        }
      } else {
        for (int i = j; i < 9 * cnt; i += 8) values[base + i] = acc[i];
      }
      if (dinv != nullptr && j < 3) {
        double di = 0.0;  // a node no element references keeps u = 0
        if (cnt > 0) {
          const int kd = find_slot(cols, cnt, (int)node);
          const double diag = acc[j * row_len + 3 * kd + j];
          const bool is_fixed = fixed != nullptr && fixed[3 * node + j];
          di = is_fixed ? 0.0 : 1.0 / diag;
        }
        dinv[3 * node + j] = di;
      }
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

// FEA_ASSEMBLE_AFFINE=0: every node through the general kernel (A/B switch).
static bool affine_pass_disabled() {
  const char* v = std::getenv("FEA_ASSEMBLE_AFFINE");
  return v != nullptr && v[0] == '0';
}

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
  // Pass 0 + 1: nodes whose incident elements are all exactly affine (hex8_affine_geometry_kernel,
  // assemble_hex8_affine_kernel); a per-node flag is left for the others and the elements the closed form cannot
  // take are counted.  If there are any: pass 0b (J^-1 and detJ of every element at its 8 Gauss points, once) and
  // pass 2 (the Gauss-point kernel on the flagged nodes).  Scratch is stream-ordered memory of the device's default
  // pool, which keeps what it has handed out once (release threshold): no OS allocation per assembly.
  int dev = 0;
  cudaMemPool_t pool = nullptr;
  FEA_TRY(check(cudaGetDevice(&dev)));
  FEA_TRY(check(cudaDeviceGetDefaultMemPool(&pool, dev)));
  uint64_t keep = 1ull << 31;
  FEA_TRY(check(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)));
  uint8_t* todo = nullptr;
  double* geom = nullptr;
  const size_t smem_aff =
      sizeof(double) * (kMtab + (size_t)kAffWarps * kAffNodes * (9 * (size_t)maxc + (aff_cols_ints(maxc) + 1) / 2));
  int rc = FEA_OK;
  unsigned n_general = 1;  // without the affine pass every node is "general"
  if (smem_aff <= 100 * 1024 && n_elem > 0 && !affine_pass_disabled()) {
    const size_t geom_bytes = sizeof(double) * kGeomDoubles * (size_t)n_elem;
    const size_t todo_bytes = ((size_t)n_nodes + 15) & ~(size_t)15;
    FEA_TRY(check(cudaMallocAsync(&geom, geom_bytes + todo_bytes + 16, stream)));
    todo = reinterpret_cast<uint8_t*>(geom) + geom_bytes;
    unsigned* counter = reinterpret_cast<unsigned*>(todo + todo_bytes);
    rc = check(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
    if (rc == FEA_OK) {
      hex8_affine_geometry_kernel<<<(unsigned)ceil_div(n_elem, 256), 256, 0, stream>>>(nodes, elements, n_elem, geom,
                                                                                      counter, status);
      if (smem_aff > 48 * 1024)
        rc = check(cudaFuncSetAttribute(assemble_hex8_affine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_aff));
    }
    if (rc == FEA_OK) {
      const int64_t n_groups = ceil_div(n_nodes, kAffNodes);
      const unsigned blocks_aff = (unsigned)std::min<int64_t>(ceil_div(n_groups, kAffWarps), 148LL * 64);
      assemble_hex8_affine_kernel<<<blocks_aff, kAffWarps * 32, smem_aff, stream>>>(
          geom, elements, n_nodes, hex8_material(E, nu), n2e_ptr, n2e, node_rowptr, node_colidx, maxc, fixed, mode,
          values, dinv, todo);
      rc = check_launch(2);
    }
    if (rc == FEA_OK) rc = check(cudaMemcpyAsync(&n_general, counter, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    if (rc == FEA_OK) rc = check(cudaStreamSynchronize(stream));
  }
  double* gauss = nullptr;
  if (rc == FEA_OK && n_general != 0 && n_elem > 0) {
    rc = check(cudaMallocAsync(&gauss, sizeof(double) * kGaussDoubles * 8 * (size_t)n_elem, stream));
    if (rc == FEA_OK) {
      hex8_gauss_geometry_kernel<<<(unsigned)ceil_div(8 * n_elem, 256), 256, 0, stream>>>(nodes, elements, n_elem, gauss,
                                                                                         status);
      const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n_nodes, kAsmWarps * 8), 148LL * 64);
      assemble_hex8_kernel<<<blocks, kAsmWarps * 32, smem, stream>>>(nodes, elements, n_nodes, hex8_material(E, nu),
                                                                     n2e_ptr, n2e, node_rowptr, node_colidx, maxc, fixed,
                                                                     mode, values, dinv, todo, gauss, status);
      rc = check_launch(2);
    }
  } else if (rc == FEA_OK && n_elem == 0) {
    // no elements at all: every node is isolated; the Gauss-point kernel writes dinv = 0 (and no values)
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n_nodes, kAsmWarps * 8), 148LL * 64);
    assemble_hex8_kernel<<<blocks, kAsmWarps * 32, smem, stream>>>(nodes, elements, n_nodes, hex8_material(E, nu), n2e_ptr,
                                                                   n2e, node_rowptr, node_colidx, maxc, fixed, mode,
                                                                   values, dinv, nullptr, nullptr, status);
    rc = check_launch();
  }
  if (gauss != nullptr) cudaFreeAsync(gauss, stream);
  if (geom != nullptr) cudaFreeAsync(geom, stream);
  return rc;
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
