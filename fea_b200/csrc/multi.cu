// Batched multi-right-hand-side Jacobi-PCG (BASELINE config 5: 10 M-member space truss, 64 load
// cases).  Same recurrence as pcg.cu, independently per column: every column has its own
// alpha / beta and stops updating once its own recurrence residual meets the tolerance
// (alpha = beta = 0 from then on).  Vectors are (n_dof, R) row-major so that one matrix pass
// serves all R systems: K is read once per iteration instead of R times.
//
//   step 1  AP = K P  (SpMM)        + per-column partial p.ap
//   step 2  X += a P; Rv -= a AP    + per-column partial r.dinv.r and r.r (free DOF)
//   step 3  P = dinv Rv + b P       + per-column convergence bookkeeping in the last block
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "spmm.cuh"

namespace fea {

constexpr int kMaxRhs = 256;
constexpr int kMultiBlocks = 148 * 4;  // upper bound on grid size of the reducing kernels

struct MultiState {
  int32_t iter, done, status, max_iter, n_active, n_rhs;
  uint32_t counter[4];
  double tol2;
  double pad[2];
};
static_assert(sizeof(MultiState) == 64, "MultiState layout");

struct MultiWork {
  MultiState* st;
  double *rz, *bnorm2, *rz_new, *rr, *pap;  // R doubles each: the sums the kernels READ
  // ... and where the reducing kernels WRITE their column sums.  One GPU: the same arrays.  Multi-GPU
  // driver: a second set; the driver copies it over the first and all-reduces that.  A kernel that
  // returns early (solve finished, iterations still enqueued) then leaves its local sums untouched and
  // the world sums come out the same again -- all-reducing the shared arrays in place multiplied the
  // frozen residuals by the world size once per left-over iteration.
  double *l_rz, *l_bnorm2, *l_rz_new, *l_rr, *l_pap;
  int32_t *active, *iters;                  // R ints each
  double* partials;                         // kMultiBlocks * 2 * R
  double *Rv, *P, *AP;                      // n * R each
};

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static size_t multi_bytes(int64_t n, int R) {
  return 256 + align256(sizeof(double) * 10 * R) + align256(sizeof(int32_t) * 2 * R) +
         align256(sizeof(double) * (size_t)kMultiBlocks * 2 * R) + 3 * align256(sizeof(double) * (size_t)n * R);
}

static MultiWork carve_multi(void* work, int64_t n, int R, bool split = false) {
  MultiWork w;
  char* c = static_cast<char*>(work);
  w.st = reinterpret_cast<MultiState*>(c);
  c += 256;
  double* d = reinterpret_cast<double*>(c);
  w.rz = d; w.bnorm2 = d + R; w.rz_new = d + 2 * R; w.rr = d + 3 * R; w.pap = d + 4 * R;
  double* l = split ? d + 5 * R : d;
  w.l_rz = l; w.l_bnorm2 = l + R; w.l_rz_new = l + 2 * R; w.l_rr = l + 3 * R; w.l_pap = l + 4 * R;
  c += align256(sizeof(double) * 10 * R);
  w.active = reinterpret_cast<int32_t*>(c);
  w.iters = w.active + R;
  c += align256(sizeof(int32_t) * 2 * R);
  w.partials = reinterpret_cast<double*>(c);
  c += align256(sizeof(double) * (size_t)kMultiBlocks * 2 * R);
  const size_t vec = align256(sizeof(double) * (size_t)n * R);
  w.Rv = reinterpret_cast<double*>(c); c += vec;
  w.P = reinterpret_cast<double*>(c); c += vec;
  w.AP = reinterpret_cast<double*>(c);
  return w;
}

// Per-block column sums -> partials[block][s][col]; returns true (all threads) in the last block.
// s_cols: n_scalars * R doubles of shared memory already holding this block's sums.
__device__ __forceinline__ bool publish_columns(double* partials, const double* s_cols, int n_scalars, int R,
                                                uint32_t* counter) {
  __shared__ bool s_last;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  for (int i = tid; i < n_scalars * R; i += nthreads)
    partials[((size_t)blockIdx.x * 2) * R + i] = s_cols[i];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  return s_last;
}

// Fixed-order sum over blocks of partials[.][s][col]; every thread handles columns tid, tid+T, ...
__device__ __forceinline__ void reduce_columns(const double* partials, int s, int R, double* out) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  for (int col = tid; col < R; col += nthreads) {
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partials + ((size_t)b * 2 + s) * R + col);
    out[col] = t;
  }
}

// block = (32, 8): x <-> CPL adjacent columns inside a 32*CPL-wide tile, y <-> row lane.
// CPL = 2 (even R, 16-byte aligned arrays): every access is a 16-byte vector, a warp covers one
// 512-byte row segment per load and two rows are in flight per trip -- the vector kernels are pure
// HBM streams (update: 4 reads + 2 writes of n*R doubles, direction: 2 reads + 1 write).

// Sum over the 8 row lanes of per-thread column sums -> s_cols[s * R + col].
template <int CPL, int NS>
__device__ __forceinline__ void fold_columns(const double (&v)[NS][CPL], int col0, int R, double* s_cols,
                                             double (*s_red)[8][32 * CPL + 1]) {
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int j = 0; j < CPL; ++j) s_red[s][ty][tx * CPL + j] = v[s][j];
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (col0 + j < R) {
          double t = 0.0;
          for (int y = 0; y < 8; ++y) t += s_red[s][y][tx * CPL + j];
          s_cols[s * R + col0 + j] = t;
        }
  }
  __syncthreads();
}

template <int CPL>
__global__ void __launch_bounds__(256)
multi_init_kernel(int64_t n, int R, const double* __restrict__ B, const double* __restrict__ dinv,
                  double* __restrict__ X, double* __restrict__ Rv, double* __restrict__ P, double tol, int max_iter,
                  MultiWork w) {
  __shared__ double s_red[2][8][32 * CPL + 1];
  __shared__ double s_cols[2 * kMaxRhs];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tiles = (R + 32 * CPL - 1) / (32 * CPL);
  for (int tile = 0; tile < tiles; ++tile) {
    const int col0 = tile * 32 * CPL + tx * CPL;
    double acc[2][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[0][j] = acc[1][j] = 0.0;
    if (col0 < R) {
      for (int64_t i = (int64_t)blockIdx.x * 8 + ty; i < n; i += (int64_t)gridDim.x * 8) {
        const double di = dinv[i];
        double bi[CPL], ri[CPL], zi[CPL], zero[CPL];
        ColVec<CPL>::load_plain(B + i * R + col0, bi);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          ri[j] = di != 0.0 ? bi[j] : 0.0;
          zi[j] = di * ri[j];
          zero[j] = 0.0;
          acc[0][j] = fma(ri[j], zi[j], acc[0][j]);
          acc[1][j] = fma(ri[j], ri[j], acc[1][j]);
        }
        ColVec<CPL>::store(X + i * R + col0, zero);
        ColVec<CPL>::store(Rv + i * R + col0, ri);
        ColVec<CPL>::store(P + i * R + col0, zi);
      }
    }
    fold_columns<CPL, 2>(acc, col0, R, s_cols, s_red);
  }
  if (publish_columns(w.partials, s_cols, 2, R, &w.st->counter[3])) {
    reduce_columns(w.partials, 0, R, w.l_rz);
    reduce_columns(w.partials, 1, R, w.l_bnorm2);
    __syncthreads();
    const int tid = ty * 32 + tx;
    int act = 0;
    for (int col = tid; col < R; col += 256) {
      const int a = w.l_bnorm2[col] > 0.0 ? 1 : 0;
      w.active[col] = a;
      w.iters[col] = 0;
      w.rr[col] = w.l_bnorm2[col];
      w.rz_new[col] = 0.0;
      w.pap[col] = 0.0;
      act += a;
    }
    act = __syncthreads_count(act);  // number of threads with any active column; > 0 is all we need
    if (tid == 0) {
      MultiState* st = w.st;
      st->iter = 0;
      st->status = FEA_OK;
      st->max_iter = max_iter;
      st->n_rhs = R;
      st->n_active = act;
      st->done = act == 0;
      st->tol2 = tol * tol;
      st->counter[0] = st->counter[1] = st->counter[2] = st->counter[3] = 0;
    }
  }
}

// step 1: AP = K P, G consecutive nodes per warp (spmm.cuh), fused with the per-column partial p.ap.
template <int D, int CPL, int G, int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
multi_spmm_kernel(int64_t n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                  const double* __restrict__ values, const double* __restrict__ P, double* __restrict__ AP, int R,
                  MultiWork w, const double* __restrict__ Pown) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ double s_cols[kMaxRhs];
  __shared__ double s_red[1][8][32 * CPL + 1];
  if (w.st->done) return;
  const int lane = threadIdx.x, warp = threadIdx.y;
  SpmmGroupSmem<D, G>& sm = reinterpret_cast<SpmmGroupSmem<D, G>*>(s_dyn)[warp];
  const int tiles = (R + 32 * CPL - 1) / (32 * CPL);
  for (int tile = 0; tile < tiles; ++tile) {
    const int col0 = tile * 32 * CPL + lane * CPL;
    double dot[1][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) dot[0][j] = 0.0;
    spmm_sweep<D, CPL, G, U, true>(n_nodes, node_rowptr, node_colidx, values, P, AP, R, col0, col0 < R, lane, warp, sm,
                                   dot, Pown);
    fold_columns<CPL, 1>(dot, col0, R, s_cols, s_red);
  }
  if (publish_columns(w.partials, s_cols, 1, R, &w.st->counter[0])) {
    reduce_columns(w.partials, 0, R, w.l_pap);
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) w.st->counter[0] = 0;
  }
}

// step 2
template <int CPL>
__global__ void __launch_bounds__(256)
multi_update_kernel(int64_t n, int R, const double* __restrict__ dinv, const double* __restrict__ P,
                    const double* __restrict__ AP, double* __restrict__ X, double* __restrict__ Rv, MultiWork w) {
  __shared__ double s_red[2][8][32 * CPL + 1];
  __shared__ double s_cols[2 * kMaxRhs];
  __shared__ int s_bad;
  MultiState* st = w.st;
  if (st->done) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (tx == 0 && ty == 0) s_bad = 0;
  __syncthreads();
  const int tiles = (R + 32 * CPL - 1) / (32 * CPL);
  const int64_t step = (int64_t)gridDim.x * 8;
  for (int tile = 0; tile < tiles; ++tile) {
    const int col0 = tile * 32 * CPL + tx * CPL;
    double acc[2][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[0][j] = acc[1][j] = 0.0;
    if (col0 < R) {
      double alpha[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        alpha[j] = 0.0;
        if (col0 + j < R && w.active[col0 + j]) {
          const double pap = w.pap[col0 + j];
          if (pap > 0.0) alpha[j] = w.rz[col0 + j] / pap; else s_bad = 1;
        }
      }
      auto finish = [&](int64_t idx, double di, const double (&ap)[CPL], double (&rv)[CPL], const double (&pv)[CPL],
                        double (&xv)[CPL]) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          rv[j] = fma(-alpha[j], ap[j], rv[j]);
          xv[j] = fma(alpha[j], pv[j], xv[j]);
          if (di != 0.0) {
            acc[0][j] = fma(rv[j] * di, rv[j], acc[0][j]);
            acc[1][j] = fma(rv[j], rv[j], acc[1][j]);
          }
        }
        ColVec<CPL>::store(X + idx, xv);
        ColVec<CPL>::store(Rv + idx, rv);
      };
      int64_t i = (int64_t)blockIdx.x * 8 + ty;
      for (; i + step < n; i += 2 * step) {  // two rows in flight: 8 independent vector loads per thread
        const int64_t i0 = i * R + col0, i1 = (i + step) * R + col0;
        double ap0[CPL], rv0[CPL], pv0[CPL], xv0[CPL], ap1[CPL], rv1[CPL], pv1[CPL], xv1[CPL];
        const double d0 = dinv[i], d1 = dinv[i + step];
        ColVec<CPL>::load_plain(AP + i0, ap0);
        ColVec<CPL>::load_plain(Rv + i0, rv0);
        ColVec<CPL>::load_plain(P + i0, pv0);
        ColVec<CPL>::load_plain(X + i0, xv0);
        ColVec<CPL>::load_plain(AP + i1, ap1);
        ColVec<CPL>::load_plain(Rv + i1, rv1);
        ColVec<CPL>::load_plain(P + i1, pv1);
        ColVec<CPL>::load_plain(X + i1, xv1);
        finish(i0, d0, ap0, rv0, pv0, xv0);
        finish(i1, d1, ap1, rv1, pv1, xv1);
      }
      for (; i < n; i += step) {
        const int64_t i0 = i * R + col0;
        double ap0[CPL], rv0[CPL], pv0[CPL], xv0[CPL];
        const double d0 = dinv[i];
        ColVec<CPL>::load_plain(AP + i0, ap0);
        ColVec<CPL>::load_plain(Rv + i0, rv0);
        ColVec<CPL>::load_plain(P + i0, pv0);
        ColVec<CPL>::load_plain(X + i0, xv0);
        finish(i0, d0, ap0, rv0, pv0, xv0);
      }
    }
    fold_columns<CPL, 2>(acc, col0, R, s_cols, s_red);
  }
  const bool bad = s_bad != 0;
  if (publish_columns(w.partials, s_cols, 2, R, &st->counter[1])) {
    reduce_columns(w.partials, 0, R, w.l_rz_new);
    reduce_columns(w.partials, 1, R, w.l_rr);
    __syncthreads();
    if (tx == 0 && ty == 0) {
      st->iter += 1;
      st->counter[1] = 0;
      if (bad) {  // p.Ap <= 0 on an active column: K_ff not positive definite
        st->done = 1;
        st->status = FEA_ERR_BREAKDOWN;
      }
    }
  }
}

// step 3
template <int CPL>
__global__ void __launch_bounds__(256)
multi_direction_kernel(int64_t n, int R, const double* __restrict__ dinv, const double* __restrict__ Rv,
                       double* __restrict__ P, MultiWork w) {
  __shared__ bool s_last;
  MultiState* st = w.st;
  if (st->done) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tiles = (R + 32 * CPL - 1) / (32 * CPL);
  const double tol2 = st->tol2;
  const int64_t step = (int64_t)gridDim.x * 8;
  for (int tile = 0; tile < tiles; ++tile) {
    const int col0 = tile * 32 * CPL + tx * CPL;
    if (col0 < R) {
      double beta[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int col = col0 + j;
        beta[j] = 0.0;
        if (col < R && w.active[col] && !(w.rr[col] <= tol2 * w.bnorm2[col])) beta[j] = w.rz_new[col] / w.rz[col];
      }
      int64_t i = (int64_t)blockIdx.x * 8 + ty;
      for (; i + 3 * step < n; i += 4 * step) {  // four rows in flight
        double rv[4][CPL], pv[4][CPL], di[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t idx = (i + u * step) * R + col0;
          di[u] = dinv[i + u * step];
          ColVec<CPL>::load_plain(Rv + idx, rv[u]);
          ColVec<CPL>::load_plain(P + idx, pv[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) pv[u][j] = fma(beta[j], pv[u][j], di[u] * rv[u][j]);
          ColVec<CPL>::store(P + (i + u * step) * R + col0, pv[u]);
        }
      }
      for (; i < n; i += step) {
        const int64_t idx = i * R + col0;
        const double di = dinv[i];
        double rv[CPL], pv[CPL];
        ColVec<CPL>::load_plain(Rv + idx, rv);
        ColVec<CPL>::load_plain(P + idx, pv);
#pragma unroll
        for (int j = 0; j < CPL; ++j) pv[j] = fma(beta[j], pv[j], di * rv[j]);
        ColVec<CPL>::store(P + idx, pv);
      }
    }
  }
  __threadfence();
  __syncthreads();
  const int tid = ty * 32 + tx;
  if (tid == 0) s_last = atomicAdd(&st->counter[2], 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    int act = 0;
    for (int col = tid; col < R; col += 256) {
      if (w.active[col]) {
        if (w.rr[col] <= tol2 * w.bnorm2[col]) {
          w.active[col] = 0;
          w.iters[col] = st->iter;
        } else {
          act = 1;
          w.iters[col] = st->iter;
        }
      }
      w.rz[col] = w.rz_new[col];
    }
    act = __syncthreads_count(act);
    if (tid == 0) {
      st->n_active = act;
      st->counter[2] = 0;
      if (act == 0) {
        st->done = 1;
      } else if (st->iter >= st->max_iter) {
        st->done = 1;
        st->status = FEA_ERR_MAXITER;
      }
    }
  }
}

template <int D, int CPL, int G, int U, int MINB>
static int launch_multi_spmm_variant(int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                                     const double* P, double* AP, int R, const MultiWork& w, cudaStream_t stream,
                                     const double* Pown = nullptr) {
  constexpr size_t smem = sizeof(SpmmGroupSmem<D, G>) * kSpmmWarps;
  // per call: the attribute is per device, and a process may drive several (a few hundred ns on the host)
  FEA_TRY(check(cudaFuncSetAttribute(multi_spmm_kernel<D, CPL, G, U, MINB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  const unsigned blocks = (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(ceil_div(n_nodes, G), kSpmmWarps), std::min(148 * 2 * MINB, kMultiBlocks)));
  multi_spmm_kernel<D, CPL, G, U, MINB><<<blocks, dim3(32, 8), smem, stream>>>(n_nodes, rp, ci, values, P, AP, R, w,
                                                                               Pown);
  return FEA_OK;
}

template <int D>
static int launch_multi_spmm(int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                             const double* P, double* AP, int R, bool vec, const MultiWork& w, cudaStream_t stream,
                             const double* Pown = nullptr) {
  if (!vec) return launch_multi_spmm_variant<D, 1, kSpmmGroup, 2, 2>(n_nodes, rp, ci, values, P, AP, R, w, stream, Pown);
  return launch_multi_spmm_variant<D, 2, 2, 3, 2>(n_nodes, rp, ci, values, P, AP, R, w, stream, Pown);
}

// Multi-GPU driver: after the world sums of (r.z, ||b||^2) have replaced the local ones, derive the
// per-column activity flags and the done flag from the WORLD norms (the init kernel did it from local
// ones: a rank whose slab carries no load of some column would otherwise switch that column off).
__global__ void multi_activate_kernel(MultiWork w) {
  const int R = w.st->n_rhs;
  int act = 0;
  for (int col = threadIdx.x; col < R; col += blockDim.x) {
    const int a = w.bnorm2[col] > 0.0 ? 1 : 0;
    w.active[col] = a;
    w.rr[col] = w.bnorm2[col];
    w.rz_new[col] = 0.0;
    act += a;
  }
  act = __syncthreads_count(act);
  if (threadIdx.x == 0) {
    w.st->n_active = act;
    w.st->done = act == 0;
  }
}

}  // namespace fea

using namespace fea;

extern "C" size_t fea_pcg_multi_workspace(int64_t n_dof, int32_t n_rhs) { return multi_bytes(n_dof, n_rhs); }

extern "C" int fea_pcg_solve_multi(int64_t n_nodes, int32_t d, const int32_t* node_rowptr,
                                   const int32_t* node_colidx, const double* values, const double* dinv,
                                   const double* B, double* X, int32_t n_rhs, double tol, int32_t max_iter, void* work,
                                   size_t work_bytes, int32_t* iterations_host, fea_pcg_result* result_host,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !values || !dinv || !B || !X || !work || !result_host) return FEA_ERR_INVALID;
  if (n_nodes <= 0 || d < 1 || d > 3 || max_iter < 1 || n_rhs < 1 || n_rhs > kMaxRhs) return FEA_ERR_INVALID;
  const int64_t n = n_nodes * d;
  const int R = n_rhs;
  if (work_bytes < multi_bytes(n, R)) return FEA_ERR_WORKSPACE;
  MultiWork w = carve_multi(work, n, R);

  MultiState* snap = static_cast<MultiState*>(pinned_scratch(0, 2 * sizeof(MultiState)));
  if (snap == nullptr) return FEA_ERR_CUDA;
  cudaEvent_t ev[2];
  int rc = check(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  if (rc != FEA_OK) return rc;
  rc = check(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  if (rc != FEA_OK) {
    cudaEventDestroy(ev[0]);
    return rc;
  }
  const unsigned vb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 8 * 8), kMultiBlocks));
  const dim3 block(32, 8);
  // 16-byte vector path: even R and aligned caller arrays (the workspace vectors are 256 B aligned)
  const bool vec = spmm_can_vectorise(R, B, X);
  rc = check(cudaMemsetAsync(w.st, 0, 256, stream));
  if (rc == FEA_OK) {
    if (vec)
      multi_init_kernel<2><<<vb, block, 0, stream>>>(n, R, B, dinv, X, w.Rv, w.P, tol, max_iter, w);
    else
      multi_init_kernel<1><<<vb, block, 0, stream>>>(n, R, B, dinv, X, w.Rv, w.P, tol, max_iter, w);
    rc = check_launch();
  }
  const int chunk = 16;
  int enqueued = 0, slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  while (rc == FEA_OK && !finished) {
    const int todo = std::min(chunk, max_iter - enqueued);
    for (int it = 0; it < todo; ++it) {
      switch (d) {
        case 1: rc = launch_multi_spmm<1>(n_nodes, node_rowptr, node_colidx, values, w.P, w.AP, R, vec, w, stream); break;
        case 2: rc = launch_multi_spmm<2>(n_nodes, node_rowptr, node_colidx, values, w.P, w.AP, R, vec, w, stream); break;
        default: rc = launch_multi_spmm<3>(n_nodes, node_rowptr, node_colidx, values, w.P, w.AP, R, vec, w, stream); break;
      }
      if (rc != FEA_OK) break;
      if (vec) {
        multi_update_kernel<2><<<vb, block, 0, stream>>>(n, R, dinv, w.P, w.AP, X, w.Rv, w);
        multi_direction_kernel<2><<<vb, block, 0, stream>>>(n, R, dinv, w.Rv, w.P, w);
      } else {
        multi_update_kernel<1><<<vb, block, 0, stream>>>(n, R, dinv, w.P, w.AP, X, w.Rv, w);
        multi_direction_kernel<1><<<vb, block, 0, stream>>>(n, R, dinv, w.Rv, w.P, w);
      }
    }
    if (rc == FEA_OK) rc = check_launch(3 * todo);
    if (rc != FEA_OK) break;
    enqueued += todo;
    rc = check(cudaMemcpyAsync(&snap[slot], w.st, sizeof(MultiState), cudaMemcpyDeviceToHost, stream));
    if (rc != FEA_OK) break;
    rc = check(cudaEventRecord(ev[slot], stream));
    if (rc != FEA_OK) break;
    pending[slot] = true;
    const int prev = slot ^ 1;
    if (pending[prev]) {
      rc = check(cudaEventSynchronize(ev[prev]));
      pending[prev] = false;
      if (rc == FEA_OK && snap[prev].done) finished = true;
    }
    if (!finished && enqueued >= max_iter) finished = true;
    slot ^= 1;
  }
  double worst = 0.0, bn = 0.0;
  if (rc == FEA_OK) {
    rc = check(cudaMemcpyAsync(&snap[0], w.st, sizeof(MultiState), cudaMemcpyDeviceToHost, stream));
    if (rc == FEA_OK && iterations_host != nullptr)
      rc = check(cudaMemcpyAsync(iterations_host, w.iters, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, stream));
    double* cols = static_cast<double*>(pinned_scratch(1, sizeof(double) * 2 * R));
    if (rc == FEA_OK && cols == nullptr) rc = FEA_ERR_CUDA;
    if (rc == FEA_OK) {
      rc = check(cudaMemcpyAsync(cols, w.bnorm2, sizeof(double) * R, cudaMemcpyDeviceToHost, stream));
      if (rc == FEA_OK) rc = check(cudaMemcpyAsync(cols + R, w.rr, sizeof(double) * R, cudaMemcpyDeviceToHost, stream));
      if (rc == FEA_OK) rc = check(cudaStreamSynchronize(stream));
      if (rc == FEA_OK) {
        for (int j = 0; j < R; ++j) {
          if (cols[j] > 0.0) worst = std::max(worst, std::sqrt(cols[R + j] / cols[j]));
          bn = std::max(bn, std::sqrt(cols[j]));
        }
      }
    }
  }
  if (rc == FEA_OK) {
    result_host->iterations = snap[0].iter;
    result_host->status = snap[0].status;
    if (!snap[0].done && snap[0].status == FEA_OK) result_host->status = FEA_ERR_MAXITER;
    result_host->rel_residual = worst;
    result_host->bnorm = bn;
  } else {
    cudaStreamSynchronize(stream);
  }
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  return rc;
}


// ---------------------------------------------------------------------------------------------
// Step-level entry points of the batched solver, for the multi-GPU driver (fea_b200/dist.py:
// distributed_pcg_multi): the same kernels as fea_pcg_solve_multi on one rank's slab, with the caller
// all-reducing the per-column scalars in between -- exactly the places where the single-GPU solver's
// last blocks finish their column sums:
//   init                 -> all-reduce scalars[0 .. 2R)  (r.z, ||b||^2)        -> activate
//   step_spmm            -> all-reduce scalars[4R .. 5R) (p.Ap)
//   step_update          -> all-reduce scalars[2R .. 4R) (r.z new, r.r)
//   step_direction          (convergence bookkeeping on the world sums: identical on every rank)
// P is the caller's halo-extended (n_local_dof, R) array; the solver's other vectors live in `work`.
static bool multi_vec_ok(int R, const void* a, const void* b, const void* c) {
  return spmm_can_vectorise(R, a, b) && (reinterpret_cast<uintptr_t>(c) & 15u) == 0;
}

extern "C" int fea_pcg_multi_layout(int32_t n_rhs, int64_t* offsets_host) {
  // byte offsets inside the workspace: [0] state (64 B: iter, done, status, max_iter, n_active, n_rhs),
  // [1] the 5R WORLD sums rz | bnorm2 | rz_new | rr | pap that the kernels read, [2] active (R int32),
  // [3] iters (R int32), [4] the 5R LOCAL sums (same order) that the reducing kernels write: the driver
  // copies [4] over [1] and all-reduces [1]
  if (!offsets_host || n_rhs < 1) return FEA_ERR_INVALID;
  offsets_host[0] = 0;
  offsets_host[1] = 256;
  offsets_host[2] = 256 + (int64_t)align256(sizeof(double) * 10 * n_rhs);
  offsets_host[3] = offsets_host[2] + (int64_t)sizeof(int32_t) * n_rhs;
  offsets_host[4] = 256 + (int64_t)sizeof(double) * 5 * n_rhs;
  return FEA_OK;
}

extern "C" int fea_pcg_multi_init(int64_t n_dof, int32_t n_rhs, const double* B, const double* dinv, double* X,
                                  double* P_own, double tol, int32_t max_iter, void* work, size_t work_bytes,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!B || !dinv || !X || !P_own || !work || n_dof <= 0 || n_rhs < 1 || n_rhs > kMaxRhs || max_iter < 1)
    return FEA_ERR_INVALID;
  if (work_bytes < multi_bytes(n_dof, n_rhs)) return FEA_ERR_WORKSPACE;
  MultiWork w = carve_multi(work, n_dof, n_rhs, true);
  const unsigned vb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_dof, 8 * 8), kMultiBlocks));
  FEA_TRY(check(cudaMemsetAsync(w.st, 0, 256, stream)));
  if (multi_vec_ok(n_rhs, B, X, P_own))
    multi_init_kernel<2><<<vb, dim3(32, 8), 0, stream>>>(n_dof, n_rhs, B, dinv, X, w.Rv, P_own, tol, max_iter, w);
  else
    multi_init_kernel<1><<<vb, dim3(32, 8), 0, stream>>>(n_dof, n_rhs, B, dinv, X, w.Rv, P_own, tol, max_iter, w);
  return check_launch();
}

extern "C" int fea_pcg_multi_activate(int64_t n_dof, int32_t n_rhs, void* work, void* stream_) {
  if (!work || n_dof <= 0 || n_rhs < 1 || n_rhs > kMaxRhs) return FEA_ERR_INVALID;
  multi_activate_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(carve_multi(work, n_dof, n_rhs, true));
  return check_launch();
}

extern "C" int fea_pcg_multi_step_spmm(int64_t n_owned_nodes, int32_t d, const int32_t* node_rowptr_owned,
                                       const int32_t* node_colidx, const double* values, const double* P_ext,
                                       int64_t p_row_offset, int32_t n_rhs, void* work, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr_owned || !node_colidx || !values || !P_ext || !work || n_owned_nodes <= 0 || n_rhs < 1 ||
      n_rhs > kMaxRhs || p_row_offset < 0)
    return FEA_ERR_INVALID;
  const int64_t n = n_owned_nodes * d;
  MultiWork w = carve_multi(work, n, n_rhs, true);
  const double* Pown = P_ext + (size_t)p_row_offset * d * n_rhs;
  const bool vec = multi_vec_ok(n_rhs, P_ext, w.AP, Pown);
  int rc;
  switch (d) {
    case 1: rc = launch_multi_spmm<1>(n_owned_nodes, node_rowptr_owned, node_colidx, values, P_ext, w.AP, n_rhs, vec, w, stream, Pown); break;
    case 2: rc = launch_multi_spmm<2>(n_owned_nodes, node_rowptr_owned, node_colidx, values, P_ext, w.AP, n_rhs, vec, w, stream, Pown); break;
    case 3: rc = launch_multi_spmm<3>(n_owned_nodes, node_rowptr_owned, node_colidx, values, P_ext, w.AP, n_rhs, vec, w, stream, Pown); break;
    default: return FEA_ERR_INVALID;
  }
  FEA_TRY(rc);
  return check_launch();
}

extern "C" int fea_pcg_multi_step_update(int64_t n_dof, int32_t n_rhs, const double* dinv, const double* P_own,
                                         double* X, void* work, void* stream_) {
  if (!dinv || !P_own || !X || !work || n_dof <= 0 || n_rhs < 1 || n_rhs > kMaxRhs) return FEA_ERR_INVALID;
  MultiWork w = carve_multi(work, n_dof, n_rhs, true);
  const unsigned vb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_dof, 8 * 8), kMultiBlocks));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (multi_vec_ok(n_rhs, P_own, X, w.AP))
    multi_update_kernel<2><<<vb, dim3(32, 8), 0, stream>>>(n_dof, n_rhs, dinv, P_own, w.AP, X, w.Rv, w);
  else
    multi_update_kernel<1><<<vb, dim3(32, 8), 0, stream>>>(n_dof, n_rhs, dinv, P_own, w.AP, X, w.Rv, w);
  return check_launch();
}

extern "C" int fea_pcg_multi_step_direction(int64_t n_dof, int32_t n_rhs, const double* dinv, double* P_own,
                                            void* work, void* stream_) {
  if (!dinv || !P_own || !work || n_dof <= 0 || n_rhs < 1 || n_rhs > kMaxRhs) return FEA_ERR_INVALID;
  MultiWork w = carve_multi(work, n_dof, n_rhs, true);
  const unsigned vb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_dof, 8 * 8), kMultiBlocks));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (multi_vec_ok(n_rhs, P_own, w.Rv, w.AP))
    multi_direction_kernel<2><<<vb, dim3(32, 8), 0, stream>>>(n_dof, n_rhs, dinv, w.Rv, P_own, w);
  else
    multi_direction_kernel<1><<<vb, dim3(32, 8), 0, stream>>>(n_dof, n_rhs, dinv, w.Rv, P_own, w);
  return check_launch();
}
