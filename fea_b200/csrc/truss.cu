// Pin-jointed members as axial springs: the reference's truss.py (§8(f) N4).
//
//   compute_forces (truss.py:78-92)   dl = |X_b - X_a| - |x_b - x_a|, F = -k dl, f = F (x_b - x_a)/|x_b - x_a|,
//                                     forces[a] += f, forces[b] -= f, members in list order
//   relaxation loop (truss.py:95-119) forces from the current positions; residual at the first loaded
//                                     node (what the script prints); x_i += (load_i + f_i) / stiffness
//                                     for loaded nodes only; forever
//
// Both run node-parallel over the node -> member incidence lists of the symbolic pass (ascending
// member order): every node sums its members' contributions in the order the reference's sequential
// loop adds them, so the result is deterministic and run-to-run bit-identical (no FP atomics), and a
// relaxation step only touches the members of loaded nodes.  The whole loop stays on the device: two
// tiny kernels per step (residuals from the OLD positions, then the update -- the script's two
// phases, truss.py:112-119), no host round trip; T = float reproduces the script's float32 arithmetic
// (truss.py:9-10), T = double is the FP64 evaluator.
#include <algorithm>

#include "common.cuh"

namespace fea {

template <typename T>
__device__ __forceinline__ T norm3(T x, T y, T z) {
  return sqrt(x * x + y * y + z * z);
}

// Force on `node` from its incident members, summed in ascending member order.
// n2m entries are member * 2 + end (end 0 = start node, 1 = end node), as the symbolic pass writes them.
template <typename T>
__device__ __forceinline__ void node_force(const T* __restrict__ rest, const T* __restrict__ cur,
                                           const int32_t* __restrict__ members, const T* __restrict__ k,
                                           const int32_t* __restrict__ n2m_ptr, const int32_t* __restrict__ n2m,
                                           int64_t node, T f[3]) {
  f[0] = f[1] = f[2] = T(0);
  for (int32_t q = n2m_ptr[node]; q < n2m_ptr[node + 1]; ++q) {
    const int32_t entry = n2m[q], m = entry >> 1, end = entry & 1;
    const int64_t a = members[2 * m], b = members[2 * m + 1];
    const T r0 = rest[3 * b] - rest[3 * a], r1 = rest[3 * b + 1] - rest[3 * a + 1], r2 = rest[3 * b + 2] - rest[3 * a + 2];
    const T d0 = cur[3 * b] - cur[3 * a], d1 = cur[3 * b + 1] - cur[3 * a + 1], d2 = cur[3 * b + 2] - cur[3 * a + 2];
    const T len = norm3(d0, d1, d2);
    const T dl = norm3(r0, r1, r2) - len;  // truss.py:83-85
    const T force = -k[m] * dl;            // truss.py:87
    const T sgn = end == 0 ? T(1) : T(-1); // truss.py:91-92
    f[0] += sgn * (force * d0 / len);      // truss.py:88-90
    f[1] += sgn * (force * d1 / len);
    f[2] += sgn * (force * d2 / len);
  }
}

template <typename T>
__global__ void truss_member_forces_kernel(const T* __restrict__ rest, const T* __restrict__ cur,
                                           const int32_t* __restrict__ members, const T* __restrict__ k,
                                           const int32_t* __restrict__ n2m_ptr, const int32_t* __restrict__ n2m,
                                           int64_t n_nodes, T* __restrict__ forces) {
  const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  T f[3];
  node_force(rest, cur, members, k, n2m_ptr, n2m, node, f);
#pragma unroll
  for (int c = 0; c < 3; ++c) forces[3 * node + c] += f[c];  // accumulates, like the reference
}

// Phase 1 of a relaxation step: residual_i = load_i + f_i from the OLD positions, for every loaded node;
// history[step] = |residual of the first load| (truss.py:101-103).
template <typename T>
__global__ void truss_relax_residual_kernel(const T* __restrict__ rest, const T* __restrict__ cur,
                                            const int32_t* __restrict__ members, const T* __restrict__ k,
                                            const int32_t* __restrict__ n2m_ptr, const int32_t* __restrict__ n2m,
                                            const int32_t* __restrict__ load_nodes, const T* __restrict__ load_vecs,
                                            int64_t n_loads, T* __restrict__ residual, double* __restrict__ history,
                                            int step) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loads) return;
  T f[3];
  node_force(rest, cur, members, k, n2m_ptr, n2m, load_nodes[i], f);
  const T r0 = load_vecs[3 * i] + f[0], r1 = load_vecs[3 * i + 1] + f[1], r2 = load_vecs[3 * i + 2] + f[2];
  residual[3 * i] = r0;
  residual[3 * i + 1] = r1;
  residual[3 * i + 2] = r2;
  if (i == 0 && history != nullptr) history[step] = (double)norm3(r0, r1, r2);
}

// Phase 2: x_i += residual_i / stiffness (truss.py:112-119); loaded nodes are distinct.
template <typename T>
__global__ void truss_relax_update_kernel(const int32_t* __restrict__ load_nodes, const T* __restrict__ residual,
                                          int64_t n_loads, T stiffness, T* __restrict__ cur) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loads) return;
  const int64_t node = load_nodes[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) cur[3 * node + c] += residual[3 * i + c] / stiffness;
}

template <typename T>
static int relax(const void* rest, const int32_t* members, const void* k, const int32_t* n2m_ptr, const int32_t* n2m,
                 const int32_t* load_nodes, const void* load_vecs, int64_t n_loads, double stiffness, int32_t n_steps,
                 void* displaced, void* residual, double* history, cudaStream_t stream) {
  const unsigned blocks = (unsigned)std::max<int64_t>(1, ceil_div(n_loads, 128));
  for (int s = 0; s < n_steps; ++s) {
    truss_relax_residual_kernel<T><<<blocks, 128, 0, stream>>>(
        static_cast<const T*>(rest), static_cast<const T*>(displaced), members, static_cast<const T*>(k), n2m_ptr, n2m,
        load_nodes, static_cast<const T*>(load_vecs), n_loads, static_cast<T*>(residual), history, s);
    truss_relax_update_kernel<T><<<blocks, 128, 0, stream>>>(load_nodes, static_cast<const T*>(residual), n_loads,
                                                             (T)stiffness, static_cast<T*>(displaced));
  }
  return check_launch(2 * n_steps);
}

}  // namespace fea

using namespace fea;

extern "C" int fea_truss_member_forces(const void* nodes, const int32_t* members, const void* k, int64_t n_members,
                                       int64_t n_nodes, const int32_t* n2m_ptr, const int32_t* n2m,
                                       const void* displaced, void* forces, int32_t fp32, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!nodes || !members || !k || !n2m_ptr || !n2m || !displaced || !forces || n_members < 0 || n_nodes <= 0)
    return FEA_ERR_INVALID;
  const unsigned blocks = (unsigned)ceil_div(n_nodes, 128);
  if (fp32)
    truss_member_forces_kernel<float><<<blocks, 128, 0, stream>>>(
        static_cast<const float*>(nodes), static_cast<const float*>(displaced), members, static_cast<const float*>(k),
        n2m_ptr, n2m, n_nodes, static_cast<float*>(forces));
  else
    truss_member_forces_kernel<double><<<blocks, 128, 0, stream>>>(
        static_cast<const double*>(nodes), static_cast<const double*>(displaced), members,
        static_cast<const double*>(k), n2m_ptr, n2m, n_nodes, static_cast<double*>(forces));
  return check_launch();
}

extern "C" int fea_truss_relax(const void* nodes, const int32_t* members, const void* k, int64_t n_members,
                               int64_t n_nodes, const int32_t* n2m_ptr, const int32_t* n2m, const int32_t* load_nodes,
                               const void* load_vecs, int64_t n_loads, double stiffness, int32_t n_steps,
                               void* displaced, void* residual, double* history, int32_t fp32, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!nodes || !members || !k || !n2m_ptr || !n2m || !load_nodes || !load_vecs || !displaced || !residual)
    return FEA_ERR_INVALID;
  if (n_members < 0 || n_nodes <= 0 || n_loads <= 0 || n_steps < 0 || !(stiffness != 0.0)) return FEA_ERR_INVALID;
  if (fp32)
    return relax<float>(nodes, members, k, n2m_ptr, n2m, load_nodes, load_vecs, n_loads, stiffness, n_steps, displaced,
                        residual, history, stream);
  return relax<double>(nodes, members, k, n2m_ptr, n2m, load_nodes, load_vecs, n_loads, stiffness, n_steps, displaced,
                       residual, history, stream);
}
