// Direct solver for chain meshes (node i couples to i-1, i, i+1 only): the Euler-Bernoulli beam of
// euler_bernoulli.py:42-73, whose `np.linalg.solve` (euler_bernoulli.py:69) Jacobi-PCG cannot replace --
// cond(K) of the Hermite beam grows like n^4 (5e8 at n = 100, 5e20 at the 100 k elements of BASELINE
// config 2), and CG needs ~sqrt(cond) iterations.
//
// K is block tridiagonal with D x D blocks (D = 2: deflection and slope per node).  Parallel cyclic
// reduction: in step s every block row i eliminates its couplings to rows i -+ 2^s with those rows
// themselves,
//     alpha = -A_i B_{i-h}^-1,  gamma = -C_i B_{i+h}^-1        (h = 2^s)
//     A_i' = alpha A_{i-h},  C_i' = gamma C_{i+h},
//     B_i' = B_i + alpha C_{i-h} + gamma A_{i+h},  d_i' = d_i + alpha d_{i-h} + gamma d_{i+h},
// which couples it to rows i -+ 2^(s+1); after ceil(log2 n) steps every row stands alone and
// u_i = B_i^-1 d_i.  All rows of a step are independent: n threads, log2 n launches, O(n log n) work
// (17 steps for 100,001 nodes), ping-pong buffers.  K is SPD on the free DOF, so every B_i met on the
// way is a Schur complement of an SPD matrix: no pivoting needed.  Constrained DOF (homogeneous,
// euler_bernoulli.py:61-66) become identity rows / columns when the blocks are extracted, which keeps
// the chain structure and returns exactly 0 there (H6).
//
// Precision.  The elimination runs in double-double arithmetic (two FP64 words, ~106 significant bits;
// `extended` = 0 selects plain FP64).  The matrix handed in is the FP64 matrix the reference would have
// assembled, and its entries carry all the physics there is -- but FP64 *elimination* of a matrix with
// cond ~ 5e20 returns noise (LAPACK's LU and scipy's sparse LU are 99.9 % off at n = 100 k, SURVEY.md
// H3), while the exact solution of that same FP64 matrix is within 1.3e-5 of the analytic deflection
// P x^2 (3L - x) / 6EI (probed with 80-digit decimals).  Double-double costs ~20x the flops of a solve
// that takes microseconds, and turns BASELINE config 2 from "not reportable" into a 1e-5 answer.
#include <algorithm>

#include "common.cuh"

namespace fea {

constexpr int FEA_ERR_NOT_CHAIN = FEA_ERR_INVALID;

// ---- double-double (Dekker / Knuth error-free transformations; explicit _rn intrinsics so that
// nvcc's fused-multiply-add contraction cannot rewrite them)
struct dd {
  double hi, lo;
};
__device__ __forceinline__ dd quick_two_sum(double a, double b) {
  const double s = __dadd_rn(a, b);
  return dd{s, __dsub_rn(b, __dsub_rn(s, a))};
}
__device__ __forceinline__ dd two_sum(double a, double b) {
  const double s = __dadd_rn(a, b), bb = __dsub_rn(s, a);
  return dd{s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
__device__ __forceinline__ dd two_prod(double a, double b) {
  const double p = __dmul_rn(a, b);
  return dd{p, __fma_rn(a, b, -p)};
}
__device__ __forceinline__ dd operator+(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s = quick_two_sum(s.hi, __dadd_rn(s.lo, t.hi));
  return quick_two_sum(s.hi, __dadd_rn(s.lo, t.lo));
}
__device__ __forceinline__ dd operator-(dd a) { return dd{-a.hi, -a.lo}; }
__device__ __forceinline__ dd operator-(dd a, dd b) { return a + (-b); }
__device__ __forceinline__ dd operator*(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo = __dadd_rn(p.lo, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi)));
  return quick_two_sum(p.hi, p.lo);
}
__device__ __forceinline__ dd operator/(dd a, dd b) {
  const double q1 = __ddiv_rn(a.hi, b.hi);
  dd r = a - b * dd{q1, 0.0};
  const double q2 = __ddiv_rn(r.hi, b.hi);
  r = r - b * dd{q2, 0.0};
  const double q3 = __ddiv_rn(r.hi, b.hi);
  const dd q = quick_two_sum(q1, q2);
  return q + dd{q3, 0.0};
}
template <typename Real>
__device__ __forceinline__ Real make_real(double v);
template <>
__device__ __forceinline__ double make_real<double>(double v) {
  return v;
}
template <>
__device__ __forceinline__ dd make_real<dd>(double v) {
  return dd{v, 0.0};
}
__device__ __forceinline__ double to_double(double v) { return v; }
__device__ __forceinline__ double to_double(dd v) { return v.hi; }
__device__ __forceinline__ bool is_zero_or_nan(double v) { return v == 0.0 || v != v; }
__device__ __forceinline__ bool is_zero_or_nan(dd v) { return v.hi == 0.0 || v.hi != v.hi; }

template <int D, typename Real>
struct ChainRow {
  Real a[D * D], b[D * D], c[D * D], d[D];
};

template <int D, typename Real>
__device__ __forceinline__ void mat_mul(const Real* x, const Real* y, Real* out) {  // out = x y
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      Real s = x[r * D] * y[c];
#pragma unroll
      for (int k = 1; k < D; ++k) s = s + x[r * D + k] * y[k * D + c];
      out[r * D + c] = s;
    }
}
template <int D, typename Real>
__device__ __forceinline__ void mat_vec(const Real* m, const Real* v, Real* out) {
#pragma unroll
  for (int r = 0; r < D; ++r) {
    Real s = m[r * D] * v[0];
#pragma unroll
    for (int k = 1; k < D; ++k) s = s + m[r * D + k] * v[k];
    out[r] = s;
  }
}

// inverse of a 1x1 / 2x2 block; false if singular
template <int D, typename Real>
__device__ __forceinline__ bool mat_inv(const Real* m, Real* inv) {
  const Real one = make_real<Real>(1.0);
  if (D == 1) {
    inv[0] = one / m[0];
    return !is_zero_or_nan(m[0]);
  }
  const Real det = m[0] * m[3] - m[1] * m[2];
  const Real r = one / det;
  inv[0] = m[3] * r;
  inv[1] = -(m[1] * r);
  inv[2] = -(m[2] * r);
  inv[3] = m[0] * r;
  return !is_zero_or_nan(det);
}

// CSR blocks of node i -> (A_i, B_i, C_i, d_i), with Dirichlet rows / columns replaced by identity.
// status: FEA_ERR_INVALID if a node couples to anything but i-1, i, i+1.
template <int D, typename Real>
__global__ void chain_extract_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                                     const double* __restrict__ values, const uint8_t* __restrict__ fixed,
                                     const double* __restrict__ rhs, ChainRow<D, Real>* __restrict__ rows,
                                     int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ChainRow<D, Real> row;
#pragma unroll
  for (int q = 0; q < D * D; ++q) row.a[q] = row.b[q] = row.c[q] = make_real<Real>(0.0);
  const int lo = rp[i], cnt = rp[i + 1] - lo;
  FEA_ASSERT(cnt >= 0 && cnt <= 3);
  bool has_diag = false;
  for (int k = 0; k < cnt; ++k) {
    const int64_t j = ci[lo + k];
    Real* dst = j == i - 1 ? row.a : (j == i ? row.b : (j == i + 1 ? row.c : nullptr));
    if (dst == nullptr) {
      raise_status(status, FEA_ERR_NOT_CHAIN, (int)i);
      continue;
    }
    has_diag = has_diag || j == i;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        // DOF-level CSR order: row r of node i starts at D*D*lo + r*D*cnt, entry (k, c) at D*k + c
        double v = values[(int64_t)D * D * lo + (int64_t)r * D * cnt + D * k + c];
        if (fixed != nullptr && (fixed[D * i + r] || fixed[D * j + c])) v = (j == i && r == c) ? 1.0 : 0.0;
        dst[r * D + c] = make_real<Real>(v);
      }
  }
  if (!has_diag) raise_status(status, FEA_ERR_NOT_CHAIN, (int)i);
#pragma unroll
  for (int r = 0; r < D; ++r) row.d[r] = make_real<Real>((fixed != nullptr && fixed[D * i + r]) ? 0.0 : rhs[D * i + r]);
  rows[i] = row;
}

template <int D, typename Real>
__global__ void chain_pcr_step_kernel(int64_t n, int64_t h, const ChainRow<D, Real>* __restrict__ in,
                                      ChainRow<D, Real>* __restrict__ out, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ChainRow<D, Real> me = in[i];
  ChainRow<D, Real> res = me;
#pragma unroll
  for (int q = 0; q < D * D; ++q) res.a[q] = res.c[q] = make_real<Real>(0.0);
  if (i - h >= 0) {
    const ChainRow<D, Real> lo = in[i - h];
    Real inv[D * D], alpha[D * D], t[D * D], tv[D];
    if (!mat_inv<D, Real>(lo.b, inv)) raise_status(status, FEA_ERR_BREAKDOWN, (int)(i - h));
    mat_mul<D, Real>(me.a, inv, alpha);  // alpha = A_i B_{i-h}^-1 (sign applied below)
    mat_mul<D, Real>(alpha, lo.a, t);
#pragma unroll
    for (int q = 0; q < D * D; ++q) res.a[q] = -t[q];
    mat_mul<D, Real>(alpha, lo.c, t);
#pragma unroll
    for (int q = 0; q < D * D; ++q) res.b[q] = res.b[q] - t[q];
    mat_vec<D, Real>(alpha, lo.d, tv);
#pragma unroll
    for (int r = 0; r < D; ++r) res.d[r] = res.d[r] - tv[r];
  }
  if (i + h < n) {
    const ChainRow<D, Real> hi = in[i + h];
    Real inv[D * D], gamma[D * D], t[D * D], tv[D];
    if (!mat_inv<D, Real>(hi.b, inv)) raise_status(status, FEA_ERR_BREAKDOWN, (int)(i + h));
    mat_mul<D, Real>(me.c, inv, gamma);
    mat_mul<D, Real>(gamma, hi.c, t);
#pragma unroll
    for (int q = 0; q < D * D; ++q) res.c[q] = -t[q];
    mat_mul<D, Real>(gamma, hi.a, t);
#pragma unroll
    for (int q = 0; q < D * D; ++q) res.b[q] = res.b[q] - t[q];
    mat_vec<D, Real>(gamma, hi.d, tv);
#pragma unroll
    for (int r = 0; r < D; ++r) res.d[r] = res.d[r] - tv[r];
  }
  out[i] = res;
}

template <int D, typename Real>
__global__ void chain_finish_kernel(int64_t n, const ChainRow<D, Real>* __restrict__ rows,
                                    const uint8_t* __restrict__ fixed, double* __restrict__ x, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ChainRow<D, Real> row = rows[i];
  Real inv[D * D], sol[D];
  if (!mat_inv<D, Real>(row.b, inv)) raise_status(status, FEA_ERR_BREAKDOWN, (int)i);
  mat_vec<D, Real>(inv, row.d, sol);
#pragma unroll
  for (int r = 0; r < D; ++r)
    x[D * i + r] = (fixed != nullptr && fixed[D * i + r]) ? 0.0 : to_double(sol[r]);  // exactly 0 on constrained DOF (H6)
}

template <int D, typename Real>
static int chain_solve(int64_t n, const int32_t* rp, const int32_t* ci, const double* values, const uint8_t* fixed,
                       const double* rhs, double* x, void* work, int32_t* status, cudaStream_t stream) {
  ChainRow<D, Real>* buf0 = static_cast<ChainRow<D, Real>*>(work);
  ChainRow<D, Real>* buf1 = buf0 + n;
  const unsigned blocks = (unsigned)ceil_div(n, 128);
  chain_extract_kernel<D, Real><<<blocks, 128, 0, stream>>>(n, rp, ci, values, fixed, rhs, buf0, status);
  int launches = 1;
  for (int64_t h = 1; h < n; h *= 2) {
    chain_pcr_step_kernel<D, Real><<<blocks, 128, 0, stream>>>(n, h, buf0, buf1, status);
    std::swap(buf0, buf1);
    ++launches;
  }
  chain_finish_kernel<D, Real><<<blocks, 128, 0, stream>>>(n, buf0, fixed, x, status);
  return check_launch(launches + 1);
}

}  // namespace fea

using namespace fea;

extern "C" size_t fea_chain_solve_workspace(int64_t n_nodes, int32_t d, int32_t extended) {
  const size_t words = (size_t)(3 * d * d + d) * (extended ? 2 : 1);
  return 2 * sizeof(double) * words * (size_t)std::max<int64_t>(n_nodes, 1);
}

extern "C" int fea_chain_solve(int64_t n_nodes, int32_t d, const int32_t* node_rowptr, const int32_t* node_colidx,
                               const double* values, const uint8_t* fixed, const double* b, double* x,
                               int32_t extended, void* work, size_t work_bytes, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !values || !b || !x || !work || n_nodes <= 0) return FEA_ERR_INVALID;
  if (d != 1 && d != 2) return FEA_ERR_INVALID;
  if (work_bytes < fea_chain_solve_workspace(n_nodes, d, extended)) return FEA_ERR_WORKSPACE;
  static_assert(sizeof(ChainRow<2, dd>) == 2 * sizeof(ChainRow<2, double>), "workspace formula");
  if (extended) {
    if (d == 1) return chain_solve<1, dd>(n_nodes, node_rowptr, node_colidx, values, fixed, b, x, work, status, stream);
    return chain_solve<2, dd>(n_nodes, node_rowptr, node_colidx, values, fixed, b, x, work, status, stream);
  }
  if (d == 1) return chain_solve<1, double>(n_nodes, node_rowptr, node_colidx, values, fixed, b, x, work, status, stream);
  return chain_solve<2, double>(n_nodes, node_rowptr, node_colidx, values, fixed, b, x, work, status, stream);
}
