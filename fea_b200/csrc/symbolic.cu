// Symbolic assembly: element connectivity -> deterministic node-block CSR pattern.
//
// Replaces the bookkeeping half of the reference's dense scatter (cubebeam.py:80-90,
// fea.py:87-97, euler_bernoulli.py:42-49): which (row, col) pairs of K exist.  The result is
// the STRUCTURAL pattern scipy's coo->csr produces from all element index pairs,
// with sorted column indices, independent of thread scheduling.
//
// Pipeline (all int32, HBM-bound, a few passes over the connectivity):
//   1. incidence histogram (atomicAdd int; sums are order-independent)
//   2. exclusive scan -> n2e_ptr
//   3. bucket fill with atomic cursors, then a per-node sort => ascending (element, local) order
//   4. per node (one warp): gather the nodes of all incident elements, rank-sort, unique-count
//   5. exclusive scan -> node_rowptr
//   6. per node (one warp): same gather/sort, write the unique list -> node_colidx
//   7. expansion to DOF-level rowptr / colidx (d x d blocks)
#include "common.cuh"

namespace fea {

// ------------------------------------------------------------------------------------------
// exclusive scan (3 kernels; in-place safe)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                    int32_t* __restrict__ block_sums,
                                                                    int32_t* __restrict__ max_out) {
  __shared__ int s_sum[kScanThreads / 32];
  __shared__ int s_max[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int sum = 0, mx = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    const int v = i < n ? in[i] : 0;
    sum += v;
    mx = max(mx, v);
  }
  sum = warp_sum_i(sum);
  mx = warp_max_i(mx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_sum[warp] = sum;
    s_max[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    sum = lane < kScanThreads / 32 ? s_sum[lane] : 0;
    mx = lane < kScanThreads / 32 ? s_max[lane] : 0;
    sum = warp_sum_i(sum);
    mx = warp_max_i(mx);
    if (lane == 0) {
      block_sums[blockIdx.x] = sum;
      if (max_out != nullptr) atomicMax(max_out, mx);
    }
  }
}

// Single block: exclusive scan of block_sums[0..nb) in place; block_sums[nb] = total.
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ block_sums, int nb) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    int inc = v;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;  // exclusive offset of each warp
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + s_warp[warp] + inc - v;
    if (i < nb) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[nb] = s_carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* in, int32_t* out, int64_t n,
                                                                   const int32_t* __restrict__ block_sums,
                                                                   int nb) {
  __shared__ int s_warp[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < kScanThreads / 32 ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kScanThreads / 32) s_warp[lane] = winc - w;
  }
  __syncthreads();
  int run = block_sums[blockIdx.x] + s_warp[warp] + inc - sum;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nb];
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* block_sums, int32_t* max_out,
                       cudaStream_t stream) {
  if (n <= 0) {
    FEA_TRY(check(cudaMemsetAsync(out, 0, sizeof(int32_t), stream)));
    return FEA_OK;
  }
  const int nb = (int)ceil_div(n, kScanChunk);
  scan_reduce_kernel<<<nb, kScanThreads, 0, stream>>>(in, n, block_sums, max_out);
  scan_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, nb);
  scan_apply_kernel<<<nb, kScanThreads, 0, stream>>>(in, out, n, block_sums, nb);
  return check_launch(3);
}

// ------------------------------------------------------------------------------------------
// node -> (element, local node) incidence
// ------------------------------------------------------------------------------------------
__global__ void incidence_count_kernel(const int32_t* __restrict__ elements, int64_t total, int64_t n_nodes,
                                       int32_t* __restrict__ deg, int32_t* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int node = elements[i];
    if (node < 0 || node >= n_nodes) {
      atomicMax(bad, 1);
      continue;
    }
    atomicAdd(&deg[node], 1);
  }
}

__global__ void incidence_fill_kernel(const int32_t* __restrict__ elements, int64_t total,
                                      int32_t* __restrict__ cursor, int32_t* __restrict__ n2e) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int pos = atomicAdd(&cursor[elements[i]], 1);
    n2e[pos] = (int32_t)i;
  }
}

// Ascending order inside every node's segment (insertion sort; segments are short: <= 8 on a
// structured hex mesh, 26 on the truss lattice).
__global__ void incidence_sort_kernel(const int32_t* __restrict__ n2e_ptr, int32_t* __restrict__ n2e,
                                      int64_t n_nodes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int lo = n2e_ptr[i], hi = n2e_ptr[i + 1];
  for (int a = lo + 1; a < hi; ++a) {
    const int key = n2e[a];
    int b = a - 1;
    while (b >= lo && n2e[b] > key) {
      n2e[b + 1] = n2e[b];
      --b;
    }
    n2e[b + 1] = key;
  }
}

// ------------------------------------------------------------------------------------------
// coupled-node lists: one warp per node; gather, rank-sort in shared memory, unique
// ------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void __launch_bounds__(128) node_neighbors_kernel(const int32_t* __restrict__ elements, int npe,
                                                             int64_t n_nodes, const int32_t* __restrict__ n2e_ptr,
                                                             const int32_t* __restrict__ n2e, int cap,
                                                             const int32_t* __restrict__ node_rowptr,
                                                             int32_t* __restrict__ out) {
  extern __shared__ int32_t s_buf[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  int32_t* cand = s_buf + (size_t)warp * 2 * cap;
  int32_t* sorted = cand + cap;
  for (int64_t node = (int64_t)blockIdx.x * warps_per_block + warp; node < n_nodes;
       node += (int64_t)gridDim.x * warps_per_block) {
    const int lo = n2e_ptr[node];
    const int deg = n2e_ptr[node + 1] - lo;
    const int n = deg * npe;
    if (n <= 64) {
      // Register path (every node of a hex8 grid: 8 elements x 8 nodes): bitonic network over 64 keys, two per
      // lane (key i lives in lane i & 31, register i >> 5), padded with INT32_MAX; then heads of runs are
      // counted / written in order.  Same sorted-unique list as the rank sort below, at a third of its cost.
      int v[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int q = lane + 32 * r;
        v[r] = INT32_MAX;
        if (q < n) {
          const int t = q / npe, b = q - t * npe;
          v[r] = elements[(int64_t)(n2e[lo + t] / npe) * npe + b];
        }
      }
#pragma unroll
      for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
          if (jj == 32) {
            const int a = min(v[0], v[1]), b = max(v[0], v[1]);
            v[0] = a;
            v[1] = b;
          } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {  // key index i = lane + 32 r: ascending run iff (i & k) == 0
              const bool up = k == 64 || (k == 32 ? r == 0 : (lane & k) == 0);
              const bool keep_min = ((lane & jj) == 0) == up;
              const int o = __shfl_xor_sync(kFull, v[r], jj);
              v[r] = keep_min ? min(v[r], o) : max(v[r], o);
            }
          }
        }
      }
      const int p0 = __shfl_up_sync(kFull, v[0], 1), p1 = __shfl_up_sync(kFull, v[1], 1);
      const int last0 = __shfl_sync(kFull, v[0], 31);
      const bool h0 = v[0] != INT32_MAX && (lane == 0 || v[0] != p0);
      const bool h1 = v[1] != INT32_MAX && v[1] != (lane == 0 ? last0 : p1);
      const unsigned m0 = __ballot_sync(kFull, h0), m1 = __ballot_sync(kFull, h1);
      if (WRITE) {
        const int base_out = node_rowptr[node];
        const unsigned below = (1u << lane) - 1u;
        if (h0) out[base_out + __popc(m0 & below)] = v[0];
        if (h1) out[base_out + __popc(m0) + __popc(m1 & below)] = v[1];
      } else if (lane == 0) {
        out[node] = __popc(m0) + __popc(m1);
      }
      continue;
    }
    for (int q = lane; q < n; q += 32) {
      const int t = q / npe, b = q - t * npe;
      const int e = n2e[lo + t] / npe;
      cand[q] = elements[(int64_t)e * npe + b];
    }
    __syncwarp();
    for (int q = lane; q < n; q += 32) {
      const int c = cand[q];
      int r = 0;
      for (int j = 0; j < n; ++j) {
        const int cj = cand[j];
        r += (cj < c) || (cj == c && j < q);
      }
      sorted[r] = c;
    }
    __syncwarp();
    int count = 0;
    const int base_out = WRITE ? node_rowptr[node] : 0;
    for (int base = 0; base < n; base += 32) {
      const int q = base + lane;
      const bool head = q < n && (q == 0 || sorted[q] != sorted[q - 1]);
      const unsigned mask = __ballot_sync(kFull, head);
      if (WRITE && head) out[base_out + count + __popc(mask & ((1u << lane) - 1u))] = sorted[q];
      count += __popc(mask);
    }
    if (!WRITE && lane == 0) out[node] = count;
    __syncwarp();
  }
}

// DOF-level CSR arrays from the node-block pattern.
__global__ void __launch_bounds__(256) expand_pattern_kernel(int64_t n_nodes, int d,
                                                             const int32_t* __restrict__ node_rowptr,
                                                             const int32_t* __restrict__ node_colidx,
                                                             int32_t* __restrict__ rowptr,
                                                             int32_t* __restrict__ colidx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t node = warp_global; node < n_nodes; node += n_warps) {
    const int lo = node_rowptr[node];
    const int cnt = node_rowptr[node + 1] - lo;
    const int64_t base = (int64_t)d * d * lo;
    const int row_len = d * cnt;
    if (lane < d) rowptr[node * d + lane] = (int32_t)(base + (int64_t)lane * row_len);
    for (int q = lane; q < d * row_len; q += 32) {
      const int within = q % row_len;
      const int k = within / d, b = within - k * d;
      colidx[base + q] = node_colidx[lo + k] * d + b;
    }
    if (node == n_nodes - 1 && lane == 0) rowptr[n_nodes * d] = (int32_t)(base + (int64_t)d * row_len);
  }
}

static int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

struct SymWorkspace {
  int32_t* cursor;      // n_nodes + 1
  int32_t* block_sums;  // scan scratch
  int32_t* flags;       // [0] max incident, [1] max coupled, [2] bad index, [3] unused
};

static SymWorkspace carve(void* ws, int64_t n_nodes) {
  SymWorkspace w;
  int32_t* p = static_cast<int32_t*>(ws);
  w.flags = p;
  p += 8;
  w.cursor = p;
  p += n_nodes + 1;
  w.block_sums = p;
  return w;
}

}  // namespace fea

using namespace fea;

extern "C" size_t fea_csr_symbolic_workspace(int64_t n_nodes, int64_t n_elem, int32_t nodes_per_elem) {
  (void)n_elem;
  (void)nodes_per_elem;
  return sizeof(int32_t) * (8 + (size_t)(n_nodes + 1) + scan_workspace_ints(n_nodes + 1));
}

static int launch_neighbors(bool write, const int32_t* elements, int npe, int64_t n_nodes, const int32_t* n2e_ptr,
                            const int32_t* n2e, int max_incident, const int32_t* node_rowptr, int32_t* out,
                            cudaStream_t stream) {
  const int cap = next_pow2(max_incident * npe);
  int warps = 4;
  size_t smem = (size_t)warps * 2 * cap * sizeof(int32_t);
  while (smem > 160 * 1024 && warps > 1) {
    warps >>= 1;
    smem = (size_t)warps * 2 * cap * sizeof(int32_t);
  }
  if (smem > 200 * 1024) return FEA_ERR_INVALID;  // node valence beyond what one warp can sort on chip
  const int threads = warps * 32;
  const int64_t blocks = std::min<int64_t>(ceil_div(n_nodes, warps), 148LL * 64);
  if (write) {
    if (smem > 48 * 1024)
      FEA_TRY(check(cudaFuncSetAttribute(node_neighbors_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem)));
    node_neighbors_kernel<true><<<(unsigned)blocks, threads, smem, stream>>>(elements, npe, n_nodes, n2e_ptr, n2e, cap,
                                                                           node_rowptr, out);
  } else {
    if (smem > 48 * 1024)
      FEA_TRY(check(cudaFuncSetAttribute(node_neighbors_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem)));
    node_neighbors_kernel<false><<<(unsigned)blocks, threads, smem, stream>>>(elements, npe, n_nodes, n2e_ptr, n2e,
                                                                            cap, node_rowptr, out);
  }
  return check_launch();
}

extern "C" int fea_csr_symbolic_count(const int32_t* elements, int64_t n_elem, int32_t nodes_per_elem,
                                      int64_t n_nodes, int32_t* n2e_ptr, int32_t* n2e, int32_t* node_rowptr,
                                      int64_t* sizes_host, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!elements || !n2e_ptr || !n2e || !node_rowptr || !sizes_host || !workspace) return FEA_ERR_INVALID;
  if (n_nodes <= 0 || n_elem < 0 || nodes_per_elem <= 0) return FEA_ERR_INVALID;
  const int64_t total = n_elem * nodes_per_elem;
  if (total >= (int64_t)INT32_MAX || n_nodes >= (int64_t)INT32_MAX) return FEA_ERR_INVALID;
  if (workspace_bytes < fea_csr_symbolic_workspace(n_nodes, n_elem, nodes_per_elem)) return FEA_ERR_WORKSPACE;
  SymWorkspace w = carve(workspace, n_nodes);

  FEA_TRY(check(cudaMemsetAsync(w.flags, 0, 8 * sizeof(int32_t), stream)));
  FEA_TRY(check(cudaMemsetAsync(n2e_ptr, 0, (size_t)(n_nodes + 1) * sizeof(int32_t), stream)));
  const int threads = 256;
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), 148LL * 32));
  if (total > 0) {
    incidence_count_kernel<<<blocks, threads, 0, stream>>>(elements, total, n_nodes, n2e_ptr, w.flags + 2);
    FEA_TRY(check_launch());
  }
  FEA_TRY(exclusive_scan_i32(n2e_ptr, n2e_ptr, n_nodes, w.block_sums, w.flags + 0, stream));
  int32_t flags_host[8];
  FEA_TRY(check(cudaMemcpyAsync(flags_host, w.flags, sizeof(flags_host), cudaMemcpyDeviceToHost, stream)));
  FEA_TRY(check(cudaStreamSynchronize(stream)));
  if (flags_host[2] != 0) return FEA_ERR_INVALID;  // connectivity references a node outside [0, n_nodes)
  const int max_incident = flags_host[0];

  FEA_TRY(check(cudaMemcpyAsync(w.cursor, n2e_ptr, (size_t)(n_nodes + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                stream)));
  if (total > 0) {
    incidence_fill_kernel<<<blocks, threads, 0, stream>>>(elements, total, w.cursor, n2e);
    incidence_sort_kernel<<<(unsigned)ceil_div(n_nodes, 128), 128, 0, stream>>>(n2e_ptr, n2e, n_nodes);
    FEA_TRY(check_launch(2));
  }
  // unique coupled-node count per node, then scan in place
  FEA_TRY(launch_neighbors(false, elements, nodes_per_elem, n_nodes, n2e_ptr, n2e, std::max(max_incident, 1), nullptr,
                           node_rowptr, stream));
  FEA_TRY(exclusive_scan_i32(node_rowptr, node_rowptr, n_nodes, w.block_sums, w.flags + 1, stream));
  int32_t nnzb = 0;
  FEA_TRY(check(cudaMemcpyAsync(flags_host, w.flags, sizeof(flags_host), cudaMemcpyDeviceToHost, stream)));
  FEA_TRY(check(cudaMemcpyAsync(&nnzb, node_rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, stream)));
  FEA_TRY(check(cudaStreamSynchronize(stream)));
  sizes_host[0] = nnzb;
  sizes_host[1] = flags_host[1];
  sizes_host[2] = max_incident;
  sizes_host[3] = 0;
  return FEA_OK;
}

extern "C" int fea_csr_symbolic_fill(const int32_t* elements, int64_t n_elem, int32_t nodes_per_elem, int64_t n_nodes,
                                     const int32_t* n2e_ptr, const int32_t* n2e, const int32_t* node_rowptr,
                                     int32_t* node_colidx, int32_t max_incident, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)n_elem;
  if (!elements || !n2e_ptr || !n2e || !node_rowptr || !node_colidx) return FEA_ERR_INVALID;
  return launch_neighbors(true, elements, nodes_per_elem, n_nodes, n2e_ptr, n2e, std::max(max_incident, 1),
                          node_rowptr, node_colidx, stream);
}

extern "C" int fea_csr_expand(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                              const int32_t* node_colidx, int32_t* rowptr, int32_t* colidx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !rowptr || !colidx || dof_per_node < 1 || n_nodes <= 0) return FEA_ERR_INVALID;
  const int64_t blocks = std::min<int64_t>(ceil_div(n_nodes, 8), 148LL * 32);
  expand_pattern_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n_nodes, dof_per_node, node_rowptr, node_colidx, rowptr,
                                                             colidx);
  return check_launch();
}
