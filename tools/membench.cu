// Microbenchmarks that give the roofline denominators this path needs and that
// MEASURED_PEAKS.json does not hold: read-only HBM streaming (plain 128-bit loads and
// cp.async.bulk rings of various depth), and the FP64 DFMA peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/membench.bin tools/membench.cu
//   tools/membench.bin [GiB]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void read_sum(const int4* __restrict__ p, size_t n, int* out) {
  int acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    int4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += stride) acc += __ldcs(p + i).x;
  if (acc == 0x7fffffff) out[0] = acc;
}

__global__ void copy_k(const int4* __restrict__ p, int4* __restrict__ q, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) q[i] = p[i];
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// persistent CTAs; one producer thread streams `chunk`-byte pieces through `stages` smem buffers;
// one consumer warp acknowledges them (touching `touch` doubles per lane) -- the TMA-side ceiling.
__global__ void __launch_bounds__(64, 1) tma_stream(const unsigned char* __restrict__ src, size_t bytes, int chunk,
                                                    int stages, int touch, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  unsigned char* buf0 = sm + 256;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t n_chunks = bytes / chunk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc = 0;
  int q = 0;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++q) {
    const int s = q % stages;
    const uint32_t ph = (q / stages) & 1;
    if (warp == 1) {
      if (lane == 0) {
        asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}" ::"r"(s32(&empty[s])), "r"(ph ^ 1u) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf0 + (size_t)s * chunk)), "l"(src + c * (size_t)chunk), "r"(chunk), "r"(s32(&full[s])) : "memory");
      }
    } else {
      asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}" ::"r"(s32(&full[s])), "r"(ph) : "memory");
      const double* d = reinterpret_cast<const double*>(buf0 + (size_t)s * chunk);
      for (int t = 0; t < touch; ++t) acc += d[lane + 32 * t];
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

// same ring, but P producer lanes (lane j issues the chunks q = j mod P): is the per-copy cost on the
// issuing thread or in the copy engine?
__global__ void __launch_bounds__(64, 1) tma_stream_mp(const unsigned char* __restrict__ src, size_t bytes, int chunk,
                                                       int stages, int P, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  unsigned char* buf0 = sm + 256;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t n_chunks = bytes / chunk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int q = 0;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++q) {
    const int s = q % stages;
    const uint32_t ph = (q / stages) & 1;
    if (warp == 1) {
      if (lane == q % P) {
        asm volatile("{\n.reg .pred p;\nW3:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D3;\nbra W3;\nD3:\n}" ::"r"(s32(&empty[s])), "r"(ph ^ 1u) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf0 + (size_t)s * chunk)), "l"(src + c * (size_t)chunk), "r"(chunk), "r"(s32(&full[s])) : "memory");
      }
    } else {
      if (lane == 0) {
        asm volatile("{\n.reg .pred p;\nW4:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D4;\nbra W4;\nD4:\n}" ::"r"(s32(&full[s])), "r"(ph) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
      }
    }
  }
  if (q == -1) out[0] = 1;
}

__global__ void dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 5.0;
  const size_t bytes = ((size_t)(gib * (1ull << 30)) / (1 << 20)) * (1 << 20);
  unsigned char *src, *dst;
  double* out;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&dst, bytes));
  CK(cudaMalloc(&out, 1 << 24));
  CK(cudaMemset(src, 1, bytes));
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("{\"sms\": %d, \"bytes\": %zu", sms, bytes);
  for (int bps : {4, 8, 16}) {
    float ms = time_ms([&] { read_sum<<<sms * bps, 256>>>((const int4*)src, bytes / 16, (int*)out); }, 10);
    printf(", \"read_ldg128_%dcta_gbs\": %.0f", bps, bytes / ms / 1e6);
  }
  {
    float ms = time_ms([&] { copy_k<<<sms * 16, 256>>>((const int4*)src, (int4*)dst, bytes / 16); }, 10);
    printf(", \"copy_rw_gbs\": %.0f", 2.0 * bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int cfgs[][3] = {{32768, 2, 0}, {32768, 4, 0}, {32768, 6, 0}, {16384, 8, 0}, {16384, 12, 0}, {65536, 3, 0}, {32768, 5, 8}, {32768, 5, 128}};
  for (auto& c : cfgs) {
    const size_t smem = 256 + (size_t)c[0] * c[1];
    float ms = time_ms([&] { tma_stream<<<sms, 64, smem>>>(src, bytes, c[0], c[1], c[2], out); }, 10);
    printf(", \"tma_%dk_x%d_touch%d_gbs\": %.0f", c[0] / 1024, c[1], c[2], bytes / ms / 1e6);
  }
  // several CTAs (= several independent rings / producer threads) per SM
  const int multi[][3] = {{32768, 3, 2}, {32768, 2, 3}, {16384, 3, 4}, {32768, 1, 6}, {2048, 8, 8}, {34816, 3, 2}};
  for (auto& c : multi) {
    const size_t smem = 256 + (size_t)c[0] * c[1];
    float ms = time_ms([&] { tma_stream<<<sms * c[2], 64, smem>>>(src, bytes / c[0] * c[0], c[0], c[1], 0, out); }, 10);
    printf(", \"tma_%dB_x%d_%dcta_gbs\": %.0f", c[0], c[1], c[2], bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(tma_stream_mp, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int mp[][3] = {{16384, 8, 1}, {16384, 8, 2}, {16384, 8, 4}, {8192, 16, 1}, {8192, 16, 4}, {32768, 4, 2}, {32768, 6, 4}, {2048, 16, 1}, {2048, 16, 8}};
  for (auto& c : mp) {
    const size_t smem = 256 + (size_t)c[0] * c[1];
    float ms = time_ms([&] { tma_stream_mp<<<sms, 64, smem>>>(src, bytes, c[0], c[1], c[2], out); }, 10);
    printf(", \"tma_mp_%dk_x%d_P%d_gbs\": %.0f", c[0] / 1024, c[1], c[2], bytes / ms / 1e6);
  }
  {
    const int iters = 1 << 16;
    float ms = time_ms([&] { dfma_peak<<<sms * 8, 256>>>(out, iters); }, 3);
    printf(", \"fp64_dfma_tflops\": %.2f", 2.0 * 8 * iters * (double)sms * 8 * 256 / ms / 1e9);
  }
  printf("}\n");
  return 0;
}
