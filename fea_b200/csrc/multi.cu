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
  double *rz, *bnorm2, *rz_new, *rr, *pap;  // R doubles each
  int32_t *active, *iters;                  // R ints each
  double* partials;                         // kMultiBlocks * 2 * R
  double *Rv, *P, *AP;                      // n * R each
};

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static size_t multi_bytes(int64_t n, int R) {
  return 256 + align256(sizeof(double) * 5 * R) + align256(sizeof(int32_t) * 2 * R) +
         align256(sizeof(double) * (size_t)kMultiBlocks * 2 * R) + 3 * align256(sizeof(double) * (size_t)n * R);
}

static MultiWork carve_multi(void* work, int64_t n, int R) {
  MultiWork w;
  char* c = static_cast<char*>(work);
  w.st = reinterpret_cast<MultiState*>(c);
  c += 256;
  double* d = reinterpret_cast<double*>(c);
  w.rz = d; w.bnorm2 = d + R; w.rz_new = d + 2 * R; w.rr = d + 3 * R; w.pap = d + 4 * R;
  c += align256(sizeof(double) * 5 * R);
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

// block = (32, 8): x <-> column inside a 32-wide tile, y <-> row lane.
constexpr int kTilesMax = kMaxRhs / 32;

__global__ void __launch_bounds__(256)
multi_init_kernel(int64_t n, int R, const double* __restrict__ B, const double* __restrict__ dinv,
                  double* __restrict__ X, double* __restrict__ Rv, double* __restrict__ P, double tol, int max_iter,
                  MultiWork w) {
  __shared__ double s_red[2][8][33];
  __shared__ double s_cols[2 * kMaxRhs];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tiles = (R + 31) / 32;
  for (int tile = 0; tile < tiles; ++tile) {
    const int col = tile * 32 + tx;
    double a = 0.0, c = 0.0;
    if (col < R) {
      for (int64_t i = (int64_t)blockIdx.x * 8 + ty; i < n; i += (int64_t)gridDim.x * 8) {
        const double di = dinv[i];
        const double ri = di != 0.0 ? B[i * R + col] : 0.0;
        const double zi = di * ri;
        X[i * R + col] = 0.0;
        Rv[i * R + col] = ri;
        P[i * R + col] = zi;
        a = fma(ri, zi, a);
        c = fma(ri, ri, c);
      }
    }
    s_red[0][ty][tx] = a;
    s_red[1][ty][tx] = c;
    __syncthreads();
    if (ty == 0 && col < R) {
      double sa = 0.0, sc = 0.0;
      for (int y = 0; y < 8; ++y) {
        sa += s_red[0][y][tx];
        sc += s_red[1][y][tx];
      }
      s_cols[col] = sa;
      s_cols[R + col] = sc;
    }
    __syncthreads();
  }
  if (publish_columns(w.partials, s_cols, 2, R, &w.st->counter[3])) {
    reduce_columns(w.partials, 0, R, w.rz);
    reduce_columns(w.partials, 1, R, w.bnorm2);
    __syncthreads();
    const int tid = ty * 32 + tx;
    int act = 0;
    for (int col = tid; col < R; col += 256) {
      const int a = w.bnorm2[col] > 0.0 ? 1 : 0;
      w.active[col] = a;
      w.iters[col] = 0;
      w.rr[col] = w.bnorm2[col];
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

// step 1: warp per node, lane <-> CPL adjacent columns of each 32*CPL-wide tile.
template <int D, int CPL>
__global__ void __launch_bounds__(256)
multi_spmm_kernel(int64_t n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
                  const double* __restrict__ values, const double* __restrict__ P, double* __restrict__ AP, int R,
                  MultiWork w) {
  constexpr int TILES = kMaxRhs / (32 * CPL);
  __shared__ double s_cols[kMaxRhs];
  __shared__ double s_warp[8][32 * CPL + 1];
  if (w.st->done) return;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int tiles = (R + 32 * CPL - 1) / (32 * CPL);
  double dot[TILES][CPL];
#pragma unroll
  for (int t = 0; t < TILES; ++t)
#pragma unroll
    for (int j = 0; j < CPL; ++j) dot[t][j] = 0.0;
  for (int64_t node = (int64_t)blockIdx.x * 8 + warp; node < n_nodes; node += (int64_t)gridDim.x * 8) {
    const int lo = node_rowptr[node];
    const int cnt = node_rowptr[node + 1] - lo;
    const int row_len = D * cnt;
    const double* v = values + (int64_t)(D * D) * lo;
#pragma unroll
    for (int tile = 0; tile < TILES; ++tile) {
      if (tile < tiles) {
        const int col0 = tile * 32 * CPL + lane * CPL;
        double acc[D][CPL];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[a][j] = 0.0;
        for (int k = 0; k < cnt; ++k) {
          const int64_t xrow = (int64_t)D * node_colidx[lo + k];
#pragma unroll
          for (int b = 0; b < D; ++b) {
            double xv[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) xv[j] = col0 + j < R ? P[(xrow + b) * R + col0 + j] : 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
              const double m = v[a * row_len + D * k + b];
#pragma unroll
              for (int j = 0; j < CPL; ++j) acc[a][j] = fma(m, xv[j], acc[a][j]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int j = 0; j < CPL; ++j)
            if (col0 + j < R) {
              const int64_t idx = (node * D + a) * R + col0 + j;
              AP[idx] = acc[a][j];
              dot[tile][j] = fma(acc[a][j], P[idx], dot[tile][j]);
            }
      }
    }
  }
  // block reduction of the per-column dots, tile by tile
#pragma unroll
  for (int tile = 0; tile < TILES; ++tile) {
    if (tile < tiles) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) s_warp[warp][lane * CPL + j] = dot[tile][j];
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int col = tile * 32 * CPL + lane * CPL + j;
          double t = 0.0;
          for (int y = 0; y < 8; ++y) t += s_warp[y][lane * CPL + j];
          if (col < R) s_cols[col] = t;
        }
      }
      __syncthreads();
    }
  }
  if (publish_columns(w.partials, s_cols, 1, R, &w.st->counter[0])) {
    reduce_columns(w.partials, 0, R, w.pap);
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) w.st->counter[0] = 0;
  }
}

// step 2
__global__ void __launch_bounds__(256)
multi_update_kernel(int64_t n, int R, const double* __restrict__ dinv, const double* __restrict__ P,
                    const double* __restrict__ AP, double* __restrict__ X, double* __restrict__ Rv, MultiWork w) {
  __shared__ double s_red[2][8][33];
  __shared__ double s_cols[2 * kMaxRhs];
  __shared__ int s_bad;
  MultiState* st = w.st;
  if (st->done) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (tx == 0 && ty == 0) s_bad = 0;
  __syncthreads();
  const int tiles = (R + 31) / 32;
  for (int tile = 0; tile < tiles; ++tile) {
    const int col = tile * 32 + tx;
    double a = 0.0, c = 0.0;
    if (col < R) {
      double alpha = 0.0;
      if (w.active[col]) {
        const double pap = w.pap[col];
        if (pap > 0.0) alpha = w.rz[col] / pap; else s_bad = 1;
      }
      for (int64_t i = (int64_t)blockIdx.x * 8 + ty; i < n; i += (int64_t)gridDim.x * 8) {
        const int64_t idx = i * R + col;
        const double di = dinv[i];
        const double ri = fma(-alpha, AP[idx], Rv[idx]);
        X[idx] = fma(alpha, P[idx], X[idx]);
        Rv[idx] = ri;
        if (di != 0.0) {
          a = fma(ri * di, ri, a);
          c = fma(ri, ri, c);
        }
      }
    }
    s_red[0][ty][tx] = a;
    s_red[1][ty][tx] = c;
    __syncthreads();
    if (ty == 0 && col < R) {
      double sa = 0.0, sc = 0.0;
      for (int y = 0; y < 8; ++y) {
        sa += s_red[0][y][tx];
        sc += s_red[1][y][tx];
      }
      s_cols[col] = sa;
      s_cols[R + col] = sc;
    }
    __syncthreads();
  }
  const bool bad = s_bad != 0;
  if (publish_columns(w.partials, s_cols, 2, R, &st->counter[1])) {
    reduce_columns(w.partials, 0, R, w.rz_new);
    reduce_columns(w.partials, 1, R, w.rr);
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
__global__ void __launch_bounds__(256)
multi_direction_kernel(int64_t n, int R, const double* __restrict__ dinv, const double* __restrict__ Rv,
                       double* __restrict__ P, MultiWork w) {
  __shared__ bool s_last;
  MultiState* st = w.st;
  if (st->done) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tiles = (R + 31) / 32;
  const double tol2 = st->tol2;
  for (int tile = 0; tile < tiles; ++tile) {
    const int col = tile * 32 + tx;
    if (col < R) {
      double beta = 0.0;
      if (w.active[col] && !(w.rr[col] <= tol2 * w.bnorm2[col])) beta = w.rz_new[col] / w.rz[col];
      for (int64_t i = (int64_t)blockIdx.x * 8 + ty; i < n; i += (int64_t)gridDim.x * 8) {
        const int64_t idx = i * R + col;
        P[idx] = fma(beta, P[idx], dinv[i] * Rv[idx]);
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

template <int D>
static void launch_multi_spmm(int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                              const double* P, double* AP, int R, const MultiWork& w, unsigned blocks,
                              cudaStream_t stream) {
  dim3 block(32, 8);
  if (R > 32)
    multi_spmm_kernel<D, 2><<<blocks, block, 0, stream>>>(n_nodes, rp, ci, values, P, AP, R, w);
  else
    multi_spmm_kernel<D, 1><<<blocks, block, 0, stream>>>(n_nodes, rp, ci, values, P, AP, R, w);
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

  MultiState* snap = nullptr;
  FEA_TRY(check(cudaMallocHost(&snap, 2 * sizeof(MultiState))));
  cudaEvent_t ev[2];
  int rc = check(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  if (rc != FEA_OK) {
    cudaFreeHost(snap);
    return rc;
  }
  rc = check(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  if (rc != FEA_OK) {
    cudaEventDestroy(ev[0]);
    cudaFreeHost(snap);
    return rc;
  }
  const unsigned vb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 8 * 8), kMultiBlocks));
  const unsigned sb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_nodes, 8), kMultiBlocks));
  const dim3 block(32, 8);
  rc = check(cudaMemsetAsync(w.st, 0, 256, stream));
  if (rc == FEA_OK) {
    multi_init_kernel<<<vb, block, 0, stream>>>(n, R, B, dinv, X, w.Rv, w.P, tol, max_iter, w);
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
        case 1: launch_multi_spmm<1>(n_nodes, node_rowptr, node_colidx, values, w.P, w.AP, R, w, sb, stream); break;
        case 2: launch_multi_spmm<2>(n_nodes, node_rowptr, node_colidx, values, w.P, w.AP, R, w, sb, stream); break;
        default: launch_multi_spmm<3>(n_nodes, node_rowptr, node_colidx, values, w.P, w.AP, R, w, sb, stream); break;
      }
      multi_update_kernel<<<vb, block, 0, stream>>>(n, R, dinv, w.P, w.AP, X, w.Rv, w);
      multi_direction_kernel<<<vb, block, 0, stream>>>(n, R, dinv, w.Rv, w.P, w);
    }
    rc = check_launch(3 * todo);
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
    double* cols = nullptr;
    if (rc == FEA_OK) rc = check(cudaMallocHost(&cols, sizeof(double) * 2 * R));
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
      cudaFreeHost(cols);
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
  cudaFreeHost(snap);
  return rc;
}
