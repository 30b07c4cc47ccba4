// Library identification, multi-RHS SpMM and the adjacent steps of the reference scripts
// (SURVEY.md §8(f)): member forces (truss.py:78-92), beam moment/shear (euler_bernoulli.py:76-102),
// mesh extrusion (utils.py:356-376).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "spmm.cuh"

namespace fea {

// The last CUDA error is kept per host thread (the thread that saw the failure asks for the text);
// the profile counters are atomics.
static thread_local char g_err[256] = "";
static Profile g_profile;

Profile& profile() { return g_profile; }

void* pinned_scratch(int slot, size_t bytes) {
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  thread_local Buf bufs[4];
  if (slot < 0 || slot >= 4) return nullptr;
  Buf& b = bufs[slot];
  if (b.cap < bytes) {
    if (b.p != nullptr) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes < 4096 ? 4096 : bytes;
    if (cudaMallocHost(&b.p, want) != cudaSuccess) {
      cudaGetLastError();
      b.p = nullptr;
      return nullptr;
    }
    b.cap = want;
  }
  return b.p;
}

void set_last_error(cudaError_t e) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
}

// Y = K X, X/Y (n_dof, R) row-major: G consecutive nodes per warp, see spmm.cuh.
template <int D, int CPL, int G, int U, int MINB>
__global__ void __launch_bounds__(32 * kSpmmWarps, MINB)
spmm_kernel(int64_t n_nodes, const int32_t* __restrict__ node_rowptr, const int32_t* __restrict__ node_colidx,
            const double* __restrict__ values, const double* __restrict__ X, double* __restrict__ Y, int R) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SpmmGroupSmem<D, G>& sm = reinterpret_cast<SpmmGroupSmem<D, G>*>(s_dyn)[warp];
  const int tiles = (R + 32 * CPL - 1) / (32 * CPL);
  double dot[1][CPL];
  for (int tile = 0; tile < tiles; ++tile) {
    const int col0 = tile * 32 * CPL + lane * CPL;
    spmm_sweep<D, CPL, G, U, false>(n_nodes, node_rowptr, node_colidx, values, X, Y, R, col0, col0 < R, lane, warp, sm,
                                    dot);
  }
}

__global__ void beam_moment_shear_kernel(const double* __restrict__ u, const double* __restrict__ EI,
                                         const double* __restrict__ length, int64_t n_elem,
                                         double* __restrict__ moment, double* __restrict__ shear) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_elem) return;
  if (i == n_elem) {  // the reference leaves the last node's entries at 0 (quirk Q7)
    moment[i] = 0.0;
    shear[i] = 0.0;
    return;
  }
  const double L = length[i], ei = EI[i];
  const double u0 = u[2 * i], u1 = u[2 * i + 1], u2 = u[2 * i + 2], u3 = u[2 * i + 3];
  moment[i] = ei / (L * L) * (12 * u0 - 6 * L * u1 - 12 * u2 + 6 * L * u3);                          // :81-91
  shear[i] = ei / (L * L * L) * (6 * L * u0 + 2 * (L * L) * u1 - 6 * L * u2 + 4 * (L * L) * u3);    // :92-102
}

__global__ void mesh_extrude_kernel(const double* __restrict__ nodes2d, int64_t n2d, const int32_t* __restrict__ faces2d,
                                    int64_t n_faces, const double* __restrict__ z, int64_t n_layers,
                                    double* __restrict__ nodes3d, int32_t* __restrict__ elements) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < n2d * n_layers; i += stride) {  // utils.py:363-365
    const int64_t layer = i / n2d, j = i - layer * n2d;
    nodes3d[3 * i] = nodes2d[2 * j];
    nodes3d[3 * i + 1] = nodes2d[2 * j + 1];
    nodes3d[3 * i + 2] = z[layer];
  }
  for (int64_t e = tid; e < n_faces * (n_layers - 1); e += stride) {  // utils.py:368-374
    const int64_t layer = e / n_faces, f = e - layer * n_faces;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int32_t v = faces2d[4 * f + c];
      elements[8 * e + c] = (int32_t)(v + layer * n2d);
      elements[8 * e + 4 + c] = (int32_t)(v + (layer + 1) * n2d);
    }
  }
}

}  // namespace fea

using namespace fea;

extern "C" const char* fea_version(void) { return "fea_b200 0.1.0 sm_100a"; }
extern "C" const char* fea_last_cuda_error(void) { return g_err; }

extern "C" void fea_profile_enable(int32_t enable) {
  g_profile.enabled = enable;
  g_profile.launches = 0;
  g_profile.spmv_samples = 0;
  g_profile.spmv_ns = 0;
  g_profile.pcg_iterations = 0;
}

extern "C" void fea_profile_read(double* out_host) {
  out_host[0] = (double)g_profile.launches;
  out_host[1] = (double)g_profile.spmv_samples;
  out_host[2] = (double)g_profile.spmv_ns.load() * 1e-6;
  out_host[3] = (double)g_profile.pcg_iterations;
}

template <int D, int CPL, int G, int U, int MINB>
static int launch_spmm_variant(int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                               const double* X, double* Y, int R, cudaStream_t stream) {
  constexpr size_t smem = sizeof(SpmmGroupSmem<D, G>) * kSpmmWarps;
  FEA_TRY(check(cudaFuncSetAttribute(spmm_kernel<D, CPL, G, U, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem)));  // per call: per-device attribute
  const unsigned blocks = (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(ceil_div(n_nodes, G), kSpmmWarps), 148LL * 2 * MINB));
  spmm_kernel<D, CPL, G, U, MINB><<<blocks, 32 * kSpmmWarps, smem, stream>>>(n_nodes, rp, ci, values, X, Y, R);
  return check_launch();
}

template <int D>
static int launch_spmm(int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values, const double* X,
                       double* Y, int R, cudaStream_t stream) {
  if (!spmm_can_vectorise(R, X, Y))
    return launch_spmm_variant<D, 1, kSpmmGroup, 2, 2>(n_nodes, rp, ci, values, X, Y, R, stream);
  // node pairs per warp, 3 coupled nodes in flight (round 1 also measured groups of 4 and single nodes:
  // 10-25 % slower, profiles/kernels_r01b_ncu.txt)
  return launch_spmm_variant<D, 2, 2, 3, 2>(n_nodes, rp, ci, values, X, Y, R, stream);
}

extern "C" int fea_spmm(int64_t n_nodes, int32_t d, const int32_t* node_rowptr, const int32_t* node_colidx,
                        const double* values, const double* X, double* Y, int32_t n_rhs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr || !node_colidx || !values || !X || !Y || n_nodes <= 0 || n_rhs < 1) return FEA_ERR_INVALID;
  switch (d) {
    case 1: return launch_spmm<1>(n_nodes, node_rowptr, node_colidx, values, X, Y, n_rhs, stream);
    case 2: return launch_spmm<2>(n_nodes, node_rowptr, node_colidx, values, X, Y, n_rhs, stream);
    case 3: return launch_spmm<3>(n_nodes, node_rowptr, node_colidx, values, X, Y, n_rhs, stream);
    default: return FEA_ERR_INVALID;
  }
}

extern "C" int fea_beam_moment_shear(const double* u, const double* EI, const double* length, int64_t n_elem,
                                     double* moment, double* shear, void* stream_) {
  if (!u || !EI || !length || !moment || !shear || n_elem < 0) return FEA_ERR_INVALID;
  beam_moment_shear_kernel<<<(unsigned)ceil_div(n_elem + 1, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      u, EI, length, n_elem, moment, shear);
  return check_launch();
}

extern "C" int fea_mesh_extrude(const double* nodes2d, int64_t n2d, const int32_t* faces2d, int64_t n_faces,
                                const double* z_heights, int64_t n_layers, double* nodes3d, int32_t* elements,
                                void* stream_) {
  if (!nodes2d || !faces2d || !z_heights || !nodes3d || !elements || n2d <= 0 || n_faces < 0 || n_layers < 1)
    return FEA_ERR_INVALID;
  if (n2d * n_layers >= (int64_t)INT32_MAX) return FEA_ERR_INVALID;
  const int64_t work = std::max(n2d * n_layers, n_faces * (n_layers - 1));
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, 256), 148LL * 32));
  mesh_extrude_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(nodes2d, n2d, faces2d, n_faces,
                                                                            z_heights, n_layers, nodes3d, elements);
  return check_launch();
}
