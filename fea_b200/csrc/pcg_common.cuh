// Pieces shared by the single-GPU solver (pcg.cu) and the peer-to-peer multi-GPU driver (p2p.cu).
#pragma once
#include "spmv_tma.cuh"

namespace fea {

// Device-resident solver state (FEA_PCG_STATE_BYTES = 256 bytes).
struct PcgState {
  double rz;        // FEA_PCG_RZ      r.z of the current iterate
  double bnorm2;    // FEA_PCG_BNORM2  ||b||^2 over free DOF
  double rz_new;    // FEA_PCG_RZ_NEW
  double rr;        // FEA_PCG_RR      ||r||^2 over free DOF
  double pap;       // FEA_PCG_PAP
  double tol2;      // FEA_PCG_TOL2
  double rr_final;  // FEA_PCG_RR_FINAL  ||r||^2 frozen when `done` is set (later no-op steps of a
                    //                   multi-rank driver keep all-reducing the live scalars)
  double spare[9];
  int32_t iter;      // int32 index 32
  int32_t done;      // 33
  int32_t status;    // 34
  int32_t max_iter;  // 35
  uint32_t counter[4];
  int32_t pad[24];
};
static_assert(sizeof(PcgState) == FEA_PCG_STATE_BYTES, "PcgState layout");

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// NVLink peer-memory exchange primitives of the multi-GPU solver (p2p.cu).  Every rank owns a
// communication block (header + halo-extended p) that all peers map through CUDA IPC.
//   * scalars: rank s stores its partial sum(s) and an iteration tag into slot [parity][kind][s] of
//     EVERY rank's header (peer_publish, one warp, lane r -> rank r); a consumer polls the slots of
//     its OWN header (local L2) and adds the world values in rank order (peer_collect): every rank
//     and every CTA gets the bitwise identical sum, no atomics;
//   * halo: boundary rows of p are stored straight into the neighbour's p vector, then a tag.
// Tags are (epoch << 32 | iteration + 1); slots are double-buffered by iteration parity.  A rank can
// never be two exchanges ahead of a peer, because every exchange needs every rank's contribution.
constexpr int kMaxPeers = FEA_MAX_PEERS;
constexpr size_t kCommHeaderBytes = 4096;
constexpr size_t kCommViewOffset = 3584;  // device copy of this rank's PeerView inside its own header

struct PeerSlot {
  double v[2];
  long long tag;
  long long pad;
};
// Per-iteration scalar exchange, NCCL-"LL" style: every 8-byte word carries 32 bits of payload and a
// 32-bit flag, so a word is either old or complete -- no fence on the writer's side, no acquire on the
// reader's (measured on one GPU with the protocol forced on: the release / acquire version cost the SpMV
// 9 us and the vector kernel 12 us per iteration, more than the NVLink latency it was hiding).
// Two doubles = four words = one 32-byte sector.
struct LlSlot {
  unsigned long long w[4];  // (flag << 32) | {v0.lo, v0.hi, v1.lo, v1.hi}
};
struct CommHeader {
  PeerSlot slots[2][3][kMaxPeers];  // [parity][kind][source rank]: kind 0 (first exchange of a solve) only
  LlSlot ll[2][2][kMaxPeers];       // [parity][kind - 1][source rank]: the per-iteration exchanges
  long long halo_tag[2];            // [0] written by the lower neighbour, [1] by the upper one
  unsigned int counter;             // last-block ticket of the halo kernel
  int error;
  // FEA_P2P_DEBUG=1: %globaltimer stamps of iterations [kDbgFirst, kDbgFirst + kDbgIters):
  // [it][0] SpMV start (block 0), [1] face wait begins, [2] face wait ends (block 0, group 0),
  // [3] local halo push starts, [4] local halo push has released its tags, [5] SpMV last block done
  unsigned long long dbg[8][6];
  unsigned long long dbg_cta_first, dbg_cta_last;  // earliest / latest CTA start of the SpMV of iteration kDbgFirst
  int dbg_on, dbg_pad;
};
constexpr int kDbgFirst = 200, kDbgIters = 8;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void dbg_stamp(CommHeader* own, int iter, int what) {
  if (own->dbg_on && iter >= kDbgFirst && iter < kDbgFirst + kDbgIters) own->dbg[iter - kDbgFirst][what] = global_ns();
}
static_assert(sizeof(CommHeader) <= kCommViewOffset, "comm header");

struct PeerView {
  int world, rank, lower, upper;
  CommHeader* hdr[kMaxPeers];
  double* lower_dst;        // neighbour's halo rows that mirror my lowest / highest owned rows
  double* upper_dst;
  const double* lower_src;  // = p_own + lower_off
  const double* upper_src;
  long long lower_cnt, upper_cnt;  // doubles
  long long lower_off, upper_off;  // offsets of the mirrored ranges inside p_own (doubles)
  long long epoch;
  int lower_tiles, upper_tiles;    // SpMV tiles next to the lower / upper slab face (HaloGate, spmv_tma.cuh)
  int algo, pad;
};
static_assert(sizeof(PeerView) <= kCommHeaderBytes - kCommViewOffset, "PeerView fits the header page");

__device__ __forceinline__ long long peer_tag(const PeerView& pv, long long k) { return (pv.epoch << 32) | (k + 1); }

// What a kernel needs at its very start, passed BY VALUE as a kernel argument: polling the own slot
// array through the PeerView in device memory costs three dependent loads (view -> header pointer ->
// slot) before the first poll, on the critical path of every iteration.
struct PeerKey {
  CommHeader* own;   // this rank's header (local memory)
  long long epoch;
  int world, rank;
  int lower_tiles, upper_tiles;  // SpMV tiles next to the lower / upper slab face, 0 without that neighbour
  int has_lower, has_upper;
  int aligned;  // the owned rows of x start and end on 128-byte boundaries
};
__device__ __forceinline__ long long peer_tag(const PeerKey& key, long long k) { return (key.epoch << 32) | (k + 1); }
// Host-side bundle handed down the launch chain: the view in device memory (targets of the publishes
// and of the halo stores) and the key.  nullptr = single GPU.
struct PeerLaunch {
  const PeerView* view;
  PeerKey key;
};
// A peer never delivered: stop the solve with FEA_ERR_PEER instead of hanging the GPU.
__device__ __forceinline__ void peer_failure(const PeerKey& pv, PcgState* st) {
  pv.own->error = 1;
  st->status = FEA_ERR_PEER;
  st->done = 1;
  st->rr_final = st->rr;
}

__device__ __forceinline__ unsigned int ll_flag(long long epoch, long long k) {
  // never 0 (fresh blocks are zeroed); a slot is rewritten every second iteration, so a stale word can
  // only match after 2^21 iterations or 2^10 solves without a single write
  return 0x80000000u | ((unsigned int)(epoch & 0x3FF) << 21) | (unsigned int)((k + 1) & 0x1FFFFF);
}
__device__ __forceinline__ void st_volatile_v2(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// One full warp: lane r stores (v0, v1) + flag into rank r's slot for this rank (NVLink stores).
// (The peers' header pointers come from the PeerView in device memory, indexed by lane: as a by-value
// kernel argument the indexed array went to local memory and took the SpMV from 50 to 98 registers.)
__device__ __forceinline__ void peer_publish(const PeerView& pv, int kind, long long k, double v0, double v1) {
  const int lane = threadIdx.x & 31;
  FEA_ASSERT(kind >= 1 && kind <= 2 && pv.world <= kMaxPeers && pv.rank >= 0 && pv.rank < pv.world);
  if (lane < pv.world) {
    LlSlot* dst = &pv.hdr[lane]->ll[(int)(k & 1)][kind - 1][pv.rank];
    const unsigned long long f = (unsigned long long)ll_flag(pv.epoch, k) << 32;
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0);
    const unsigned long long b1 = (unsigned long long)__double_as_longlong(v1);
    st_volatile_v2(&dst->w[0], f | (b0 & 0xFFFFFFFFull), f | (b0 >> 32));
    st_volatile_v2(&dst->w[2], f | (b1 & 0xFFFFFFFFull), f | (b1 >> 32));
  }
}
// Polls one slot until all four words carry `flag`; false after ~2 s.
__device__ __forceinline__ bool ll_read(const LlSlot* src, unsigned int flag, double& v0, double& v1) {
  for (int i = 0; i < (1 << 24); ++i) {
    unsigned long long w0, w1, w2, w3;
    ld_volatile_v2(&src->w[0], w0, w1);
    ld_volatile_v2(&src->w[2], w2, w3);
    if ((unsigned int)(w0 >> 32) == flag && (unsigned int)(w1 >> 32) == flag && (unsigned int)(w2 >> 32) == flag &&
        (unsigned int)(w3 >> 32) == flag) {
      v0 = __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32)));
      v1 = __longlong_as_double((long long)((w2 & 0xFFFFFFFFull) | (w3 << 32)));
      return true;
    }
    if (i > 64) __nanosleep(100);
  }
  return false;
}
// One full warp: waits for every rank's slot of exchange (kind, k) in the own header and returns the
// rank-ordered sums in all lanes; false if a peer never arrived.
__device__ __forceinline__ bool peer_collect(const PeerKey& pv, int kind, long long k, double& s0, double& s1) {
  const int lane = threadIdx.x & 31;
  double v0 = 0.0, v1 = 0.0;
  bool ok = true;
  if (lane < pv.world) ok = ll_read(&pv.own->ll[(int)(k & 1)][kind - 1][lane], ll_flag(pv.epoch, k), v0, v1);
  ok = __all_sync(kFull, ok);
  s0 = s1 = 0.0;
  for (int r = 0; r < pv.world; ++r) {
    s0 += __shfl_sync(kFull, v0, r);
    s1 += __shfl_sync(kFull, v1, r);
  }
  return ok;
}

// Grid of the single-reduction vector kernel: ONE resident wave (3 CTAs per SM), every CTA polls the
// peer slots exactly once.
inline unsigned cgcg_blocks(int64_t n) {
  const int64_t b = (n + 256 * 2 - 1) / (256 * 2);
  return (unsigned)(b < 1 ? 1 : (b > 148LL * 3 ? 148LL * 3 : b));
}

inline unsigned vec_blocks(int64_t n) {
  // 8 resident CTAs of 256 threads per SM, one full wave (<= kMaxPartials blocks)
  const int64_t b = (n + 256 * 4 - 1) / (256 * 4);
  return (unsigned)(b < 1 ? 1 : (b > 148LL * 8 ? 148LL * 8 : b));
}

// One full warp, two exchanges polled at once: lanes [0, world) wait for exchange (kind_a, ka), lanes
// [kMaxPeers, kMaxPeers + world) for (kind_b, kb) (skipped when kb < 0).  a0 = sum of v[0] of the first,
// b0 / b1 = sums of v[0] / v[1] of the second, all in rank order, in all lanes.
__device__ __forceinline__ bool peer_collect2(const PeerKey& pv, int kind_a, long long ka, int kind_b, long long kb,
                                              double& a0, double& b0, double& b1) {
  static_assert(2 * kMaxPeers <= 32, "two exchanges fit one warp");
  const int lane = threadIdx.x & 31;
  const bool second = lane >= kMaxPeers;
  const int src_rank = second ? lane - kMaxPeers : lane;
  const bool mine = src_rank < pv.world && lane < 2 * kMaxPeers && (!second || kb >= 0);
  double v0 = 0.0, v1 = 0.0;
  bool ok = true;
  if (mine) {
    const long long k = second ? kb : ka;
    ok = ll_read(&pv.own->ll[(int)(k & 1)][(second ? kind_b : kind_a) - 1][src_rank], ll_flag(pv.epoch, k), v0, v1);
  }
  ok = __all_sync(kFull, ok);
  a0 = b0 = b1 = 0.0;
  for (int r = 0; r < pv.world; ++r) {
    a0 += __shfl_sync(kFull, v0, r);
    b0 += __shfl_sync(kFull, v0, kMaxPeers + r);
    b1 += __shfl_sync(kFull, v1, kMaxPeers + r);
  }
  return ok;
}

// Solver kernels defined in pcg.cu, launched by both drivers.
// `pv` (device pointer, may be null): with a PeerView the dot products are exchanged with the
// peer ranks inside these kernels and the direction kernel pushes the halo rows (see p2p.cu).
__global__ void __launch_bounds__(256) pcg_update_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ p,
                                  const double* __restrict__ ap, double* __restrict__ x, double* __restrict__ r,
                                  PcgState* st, double* partials, const PeerView* pv, PeerKey key);
__global__ void __launch_bounds__(256) pcg_direction_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                     double* __restrict__ p, PcgState* st, double* history, const PeerView* pv,
                                     PeerKey key);
__global__ void __launch_bounds__(256, 3) pcg_cgcg_kernel(int64_t n, const double* __restrict__ dinv, double* __restrict__ u,
                                const double* __restrict__ w, double* __restrict__ p, double* __restrict__ s,
                                double* __restrict__ x, double* __restrict__ r, PcgState* st, double* partials,
                                double* history, const PeerView* pv, PeerKey key);
__global__ void __launch_bounds__(256) pcg_init_kernel(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
                                double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, double tol,
                                int max_iter, PcgState* st, double* partials);

// step 1 (ap = K p over `n_nodes` rows, p.ap into state) on the best available kernel.
int pcg_step_spmv(int d, int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                  const double* p, double* ap, int64_t p_row_offset, PcgState* st, double* partials,
                  cudaStream_t stream, const TmaPlan* plan, const PeerLaunch* peer = nullptr);
void pcg_match_carveout();
// 0: classical PCG (3 kernels, 2 reductions per iteration); 1: single-reduction variant (2 kernels).
// FEA_PCG_ALGO=0|1 overrides the automatic choice (see pcg.cu).
constexpr int64_t kSingleReductionBelowDof = 3000000;
int pcg_algorithm(int64_t n_dof, bool multi_gpu);

}  // namespace fea
