// Device-resident input generators: the meshes the reference builds with Python
// list comprehensions (cubebeam.py:28-57, utils.py:356-376, fea.py:28-48) and the frozen config-5
// lattice, written straight into HBM.  At 8 M DOF the host builders cost seconds
// before the hot path starts; these cost microseconds, and their outputs equal the host builders'
// arrays element for element (tests/test_gpu_parity.py::test_device_mesh_builders).
#include <algorithm>

#include "common.cuh"

namespace fea {

// np.linspace(start, stop, num): arange(num) * step + start with step = (stop - start) / (num - 1), last
// point set to `stop` -- as two separately rounded operations (no FMA), like numpy's.
__device__ __forceinline__ double linspace_at(double start, double stop, int64_t num, int64_t i) {
  if (num <= 1) return start;
  if (i == num - 1) return stop;
  const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(num - 1));
  return __dadd_rn(__dmul_rn((double)i, step), start);
}

// generate_quad_grid(nx, ny, width, height), cubebeam.py:28-57: nodes x-fastest, quads [n1, n2, n4, n3].
__global__ void quad_grid_kernel(int64_t nx, int64_t ny, double width, double height, double* __restrict__ nodes2d,
                                 int32_t* __restrict__ quads) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < (nx + 1) * (ny + 1); i += stride) {  // cubebeam.py:41-45
    const int64_t iy = i / (nx + 1), ix = i - iy * (nx + 1);
    nodes2d[2 * i] = linspace_at(0.0, width, nx + 1, ix);
    nodes2d[2 * i + 1] = linspace_at(0.0, height, ny + 1, iy);
  }
  for (int64_t e = tid; e < nx * ny; e += stride) {  // cubebeam.py:48-55
    const int64_t j = e / nx, i = e - j * nx;
    const int64_t n1 = j * (nx + 1) + i, n3 = n1 + (nx + 1);
    quads[4 * e] = (int32_t)n1;
    quads[4 * e + 1] = (int32_t)(n1 + 1);
    quads[4 * e + 2] = (int32_t)(n3 + 1);
    quads[4 * e + 3] = (int32_t)n3;
  }
}

// Tube cross-section of fea.py:28-48: inner ring then outer ring of n_seg points at
// theta_i = i * (2 pi / n_seg) (np.linspace(..., endpoint=False)), periodic quads
// [i, i+n, (i+1)%n + n, (i+1)%n].
__global__ void tube_section_kernel(int64_t n_seg, double r_in, double r_out, double* __restrict__ nodes2d,
                                    int32_t* __restrict__ quads) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg) return;
  const double step = __ddiv_rn(2.0 * 3.141592653589793, (double)n_seg);
  const double th = __dmul_rn((double)i, step);
  const double c = cos(th), s = sin(th);
  nodes2d[2 * i] = c * r_in;
  nodes2d[2 * i + 1] = s * r_in;
  nodes2d[2 * (i + n_seg)] = c * r_out;
  nodes2d[2 * (i + n_seg) + 1] = s * r_out;
  const int64_t nxt = (i + 1) % n_seg;
  quads[4 * i] = (int32_t)i;
  quads[4 * i + 1] = (int32_t)(i + n_seg);
  quads[4 * i + 2] = (int32_t)(nxt + n_seg);
  quads[4 * i + 3] = (int32_t)nxt;
}

// ---- numpy's PCG64 (the bit generator of np.random.default_rng): 128-bit LCG, XSL-RR output.
// Element i of a stream is reachable in O(log i) (LCG jump-ahead), so a thread can start anywhere.
typedef unsigned __int128 u128;
struct Pcg64 {
  u128 state, inc;
};
__device__ __forceinline__ u128 pcg_mult() {
  return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
}
__device__ __forceinline__ void pcg_advance(Pcg64& g, uint64_t delta) {
  u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = g.inc;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  g.state = acc_mult * g.state + acc_plus;
}
__device__ __forceinline__ uint64_t pcg_next64(Pcg64& g) {
  g.state = g.state * pcg_mult() + g.inc;
  const uint64_t hi = (uint64_t)(g.state >> 64), lo = (uint64_t)g.state;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(g.state >> 122);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
// Generator.uniform(low, high): low + (high - low) * next_double, next_double = (u64 >> 11) / 2^53.
__device__ __forceinline__ double pcg_uniform(Pcg64& g, double low, double range) {
  const double u = (double)(pcg_next64(g) >> 11) * (1.0 / 9007199254740992.0);
  return __dadd_rn(low, __dmul_rn(range, u));
}

constexpr int kRngChunk = 32;  // consecutive draws per thread (one jump-ahead each)

// nodes = grid * h + uniform(-0.1 h, 0.1 h, (n^3, 3)) [draws 0 .. 3 n^3),
// k = uniform(500, 1500, M) [draws 3 n^3 .. 3 n^3 + M), both from default_rng(0).
__global__ void lattice_random_kernel(int64_t n, double h, uint64_t s_hi, uint64_t s_lo, uint64_t i_hi, uint64_t i_lo,
                                      int64_t n_members, double* __restrict__ nodes, double* __restrict__ k) {
  const int64_t n3 = n * n * n, total = 3 * n3 + n_members;
  const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kRngChunk;
  if (first >= total) return;
  Pcg64 g{((u128)s_hi << 64) | s_lo, ((u128)i_hi << 64) | i_lo};
  pcg_advance(g, (uint64_t)first);
  const double lo = -0.1 * h, range = __dsub_rn(0.1 * h, -0.1 * h);
  for (int64_t j = first; j < min(first + kRngChunk, total); ++j) {
    if (j < 3 * n3) {
      const int64_t node = j / 3, c = j - 3 * node;
      const int64_t iz = node / (n * n), iy = (node / n) % n, ix = node % n;  // id = (iz n + iy) n + ix
      const double g0 = (double)(c == 0 ? ix : (c == 1 ? iy : iz));
      nodes[j] = __dadd_rn(__dmul_rn(g0, h), pcg_uniform(g, lo, range));
    } else {
      k[j - 3 * n3] = pcg_uniform(g, 500.0, 1000.0);
    }
  }
}

__constant__ int c_lattice_dirs[13][3] = {{1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},  {1, -1, 0}, {1, 0, 1}, {1, 0, -1},
                                          {0, 1, 1},  {0, 1, -1}, {1, 1, 1},  {1, 1, -1}, {1, -1, 1}, {1, -1, -1}};

// Members direction by direction (13 half-space neighbour directions), start nodes in id order.  For a
// direction (dx, dy, dz) the valid start nodes form the box ix in [x0, x1) etc.; in id order they are
// enumerated z-major, so member q of the direction decodes directly.
__global__ void lattice_members_kernel(int64_t n, const int64_t* __restrict__ dir_offset, int32_t* __restrict__ members) {
  const int d = blockIdx.y;
  const int dx = c_lattice_dirs[d][0], dy = c_lattice_dirs[d][1], dz = c_lattice_dirs[d][2];
  const int64_t x0 = dx < 0 ? 1 : 0, x1 = dx > 0 ? n - 1 : n, y0 = dy < 0 ? 1 : 0, y1 = dy > 0 ? n - 1 : n;
  const int64_t z0 = dz < 0 ? 1 : 0, z1 = dz > 0 ? n - 1 : n;
  const int64_t wx = x1 - x0, wy = y1 - y0, wz = z1 - z0;
  const int64_t count = wx * wy * wz;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) {
    const int64_t iz = z0 + q / (wx * wy), iy = y0 + (q / wx) % wy, ix = x0 + q % wx;
    const int64_t a = (iz * n + iy) * n + ix, b = ((iz + dz) * n + (iy + dy)) * n + (ix + dx);
    int32_t* m = members + 2 * (dir_offset[d] + q);
    m[0] = (int32_t)a;
    m[1] = (int32_t)b;
  }
}

}  // namespace fea

using namespace fea;

static unsigned grid_for(int64_t n, int threads = 256) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, threads), 148LL * 16));
}

extern "C" int fea_mesh_quad_grid(int64_t nx, int64_t ny, double width, double height, double* nodes2d,
                                  int32_t* quads, void* stream) {
  if (!nodes2d || !quads || nx < 1 || ny < 1 || (nx + 1) * (ny + 1) >= INT32_MAX) return FEA_ERR_INVALID;
  quad_grid_kernel<<<grid_for((nx + 1) * (ny + 1)), 256, 0, static_cast<cudaStream_t>(stream)>>>(nx, ny, width, height,
                                                                                                 nodes2d, quads);
  return check_launch();
}

extern "C" int fea_mesh_tube_section(int64_t n_seg, double r_in, double r_out, double* nodes2d, int32_t* quads,
                                     void* stream) {
  if (!nodes2d || !quads || n_seg < 3 || 2 * n_seg >= INT32_MAX) return FEA_ERR_INVALID;
  tube_section_kernel<<<grid_for(n_seg), 256, 0, static_cast<cudaStream_t>(stream)>>>(n_seg, r_in, r_out, nodes2d,
                                                                                      quads);
  return check_launch();
}

extern "C" int64_t fea_mesh_lattice_members(int64_t n) {
  if (n < 1) return 0;
  return 3 * n * n * (n - 1) + 6 * n * (n - 1) * (n - 1) + 4 * (n - 1) * (n - 1) * (n - 1);
}

extern "C" int fea_mesh_lattice(int64_t n, double h, const uint64_t* pcg64_state_host, double* nodes,
                                int32_t* members, double* k, int64_t* dir_offset_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pcg64_state_host || !nodes || !members || !k || !dir_offset_dev || n < 2 || n * n * n >= INT32_MAX)
    return FEA_ERR_INVALID;
  const int64_t n_members = fea_mesh_lattice_members(n);
  int64_t offsets[14];
  offsets[0] = 0;
  const int dirs[13][3] = {{1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},  {1, -1, 0}, {1, 0, 1}, {1, 0, -1},
                           {0, 1, 1},  {0, 1, -1}, {1, 1, 1},  {1, 1, -1}, {1, -1, 1}, {1, -1, -1}};
  int64_t largest = 0;
  for (int d = 0; d < 13; ++d) {
    int64_t c = 1;
    for (int a = 0; a < 3; ++a) c *= dirs[d][a] != 0 ? n - 1 : n;
    offsets[d + 1] = offsets[d] + c;
    largest = std::max(largest, c);
  }
  if (offsets[13] != n_members) return FEA_ERR_INVALID;
  FEA_TRY(check(cudaMemcpyAsync(dir_offset_dev, offsets, sizeof(int64_t) * 14, cudaMemcpyHostToDevice, stream)));
  const int64_t draws = 3 * n * n * n + n_members;
  lattice_random_kernel<<<(unsigned)ceil_div(ceil_div(draws, kRngChunk), 128), 128, 0, stream>>>(n, h, pcg64_state_host[0], pcg64_state_host[1], pcg64_state_host[2],
                                            pcg64_state_host[3], n_members, nodes, k);
  lattice_members_kernel<<<dim3(grid_for(largest), 13), 256, 0, stream>>>(n, dir_offset_dev, members);
  return check_launch(2);
}
