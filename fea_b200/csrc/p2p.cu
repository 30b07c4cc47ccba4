// Multi-GPU Jacobi-PCG over NVLink peer memory, without NCCL in the iteration.
//
// The slab partition (fea_b200/dist.py) needs, per PCG iteration, one halo exchange of
// p with the two z-neighbours (157 KB each at 400x80x80) and two tiny all-reduces.  Through NCCL
// that is three collectives of ~15-20 us each per iteration, issued by the host; at 8 ranks a rank's
// kernels only take ~0.13 ms per iteration, so latency dominates.  Here every rank maps its peers'
// communication blocks (CUDA IPC, NVSwitch gives every pair full bandwidth) and the exchange
// happens INSIDE the three solver kernels of an iteration (pcg.cu, primitives in pcg_common.cuh),
// which stay the only launches of the CUDA graph:
//
//   SpMV       interior tiles first; a consumer group looks at the neighbours' halo tags only right
//              before its first face tile (HaloGate, spmv_tma.cuh).  Last block: publish this rank's
//              p.Ap to every rank's slot array (one warp, NVLink stores, flag-in-word: no fences)
//   update     every CTA: collect the world's p.Ap from the own slot array (local L2 polls), add
//              in rank order -> alpha; last block: publish (r.z, r.r)
//   direction  every CTA: collect (r.z, r.r) -> beta / convergence; new p
//   halo push  on a SIDE stream, concurrent with the next SpMV's interior tiles: the boundary rows of
//              the new p go into the neighbours' halo rows (NVLink stores), then ONE system fence and
//              the iteration tag.  Everything that needs system-scope ordering lives here, off the
//              critical path: measured on one GPU, a __threadfence_system costs ~5 us, and with the
//              stores and the tag inside the vector kernel it sat in every iteration's dependency chain.
// (single-reduction recurrence: update + direction are one kernel, one reduction per iteration.)
//
// Every rank forms bitwise identical sums (rank order, no atomics), hence identical convergence
// decisions.  Spins are bounded: a peer that never arrives raises FEA_ERR_PEER instead of hanging.
// Halo rows are reused safely without a second handshake: a neighbour overwrites them in its vector
// kernel of iteration k, which starts only after it has collected every rank's p.Ap of iteration k,
// i.e. after this rank's SpMV k (the reader of the old rows) has finished.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pcg_common.cuh"

namespace fea {

// First exchange of a solve (after the init kernel): world sums of (r.z, ||b||^2), and a check that
// every rank runs the same recurrence (`algo` rides in the slot's spare word).  It is the only point
// where ranks meet with host-side skew (each rank sliced, copied and assembled its slab on its own),
// so its spin bound is ~1 min instead of the ~2 s of the per-iteration exchanges.
__global__ void __launch_bounds__(32) p2p_init_exchange_kernel(PeerView pv, PcgState* st) {
  const int lane = threadIdx.x;
  constexpr int kind = 0;
  const long long k = 0, tag = peer_tag(pv, k);
  if (lane < pv.world) {
    PeerSlot* dst = &pv.hdr[lane]->slots[0][kind][pv.rank];
    dst->v[0] = st->rz;
    dst->v[1] = st->bnorm2;
    dst->pad = pv.algo;
    __threadfence_system();
    st_release_sys(&dst->tag, tag);
  }
  double v0 = 0.0, v1 = 0.0;
  bool ok = true;
  if (lane < pv.world) {
    const PeerSlot* src = &pv.hdr[pv.rank]->slots[0][kind][lane];
    ok = spin_until(&src->tag, tag, false, 1 << 27);
    v0 = ld_volatile_f64(&src->v[0]);
    v1 = ld_volatile_f64(&src->v[1]);
    ok = ok && ld_acquire_sys(&src->pad) == pv.algo;
  }
  ok = __all_sync(kFull, ok);
  double s0 = 0.0, s1 = 0.0;
  for (int r = 0; r < pv.world; ++r) {
    s0 += __shfl_sync(kFull, v0, r);
    s1 += __shfl_sync(kFull, v1, r);
  }
  if (lane == 0) {
    if (!ok) {
      pv.hdr[pv.rank]->error = 1;
      st->status = FEA_ERR_PEER;
      st->done = 1;
      st->rr_final = st->rr;
    } else {
      st->rz = s0;
      st->bnorm2 = s1;
      st->rr = s1;
    }
  }
}

// Halo rows of the initial p (later iterations store them from inside the vector kernels); the tag
// is released by the last block, nobody waits: the first SpMV gates its face tiles on it.
// Small on purpose (blocks of 128 threads, <= 48 registers): during the iteration it runs NEXT TO the
// persistent SpMV, which leaves 6400 registers and ~990 thread slots free per SM (3 CTAs x 11 warps x 56
// registers).  A 256-thread block with 32 registers does not fit beside it and would only start once
// SpMV CTAs exit -- which they do after their face tiles, which wait for this very kernel's tag on the
// neighbour.
constexpr int kHaloThreads = 128;
__global__ void __launch_bounds__(kHaloThreads, 10) p2p_halo_kernel(PeerView pv, PcgState* st) {
  __shared__ bool s_last;
  if (st->done) return;
  const long long tag = peer_tag(pv, st->iter);
  if (blockIdx.x == 0 && threadIdx.x == 0) dbg_stamp(pv.hdr[pv.rank], st->iter, 3);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pv.lower >= 0)
    for (long long i = tid; i < pv.lower_cnt; i += stride) pv.lower_dst[i] = pv.lower_src[i];
  if (pv.upper >= 0)
    for (long long i = tid; i < pv.upper_cnt; i += stride) pv.upper_dst[i] = pv.upper_src[i];
  __threadfence_system();
  __syncthreads();
  CommHeader* own = pv.hdr[pv.rank];
  if (threadIdx.x == 0) s_last = atomicAdd(&own->counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  own->counter = 0;
  __threadfence_system();
  if (pv.lower >= 0) st_release_sys(&pv.hdr[pv.lower]->halo_tag[1], tag);  // I am its upper neighbour
  if (pv.upper >= 0) st_release_sys(&pv.hdr[pv.upper]->halo_tag[0], tag);
  dbg_stamp(own, st->iter, 4);
}

// ---------------------------------------------------------------------------------------------
// Stand-alone halo exchange (fea_peer_push / fea_peer_wait): the many-load-case solver on slabs moves
// (rows x 64) search directions, 13 MB per neighbour and iteration at BASELINE config 5.
// ---------------------------------------------------------------------------------------------
constexpr int kPushThreads = 256;
__global__ void __launch_bounds__(kPushThreads)
peer_push_kernel(CommHeader* own, CommHeader* lower, const double2* __restrict__ src_lower, double2* dst_lower,
                 long long n2_lower, CommHeader* upper, const double2* __restrict__ src_upper, double2* dst_upper,
                 long long n2_upper, const int32_t* iter, long long epoch) {
  __shared__ bool s_last;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lower != nullptr)
    for (long long i = tid; i < n2_lower; i += stride) dst_lower[i] = src_lower[i];
  if (upper != nullptr)
    for (long long i = tid; i < n2_upper; i += stride) dst_upper[i] = src_upper[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&own->counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  own->counter = 0;
  __threadfence_system();
  const long long tag = (epoch << 32) | ((long long)*iter + 1);
  if (lower != nullptr) st_release_sys(&lower->halo_tag[1], tag);  // I am its upper neighbour
  if (upper != nullptr) st_release_sys(&upper->halo_tag[0], tag);
}

__global__ void peer_wait_kernel(CommHeader* own, int has_lower, int has_upper, const int32_t* iter, long long epoch) {
  const long long want = (epoch << 32) | ((long long)*iter + 1);
  bool ok = true;
  if (has_lower) ok = spin_until(&own->halo_tag[0], want, true) && ok;
  if (has_upper) ok = spin_until(&own->halo_tag[1], want, true) && ok;
  if (!ok) own->error = 1;
}

}  // namespace fea

using namespace fea;

extern "C" int fea_peer_push(void* own_block, void* lower_block, const double* src_lower, double* dst_lower,
                             int64_t count_lower, void* upper_block, const double* src_upper, double* dst_upper,
                             int64_t count_upper, const int32_t* iter_dev, int64_t epoch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!own_block || !iter_dev || count_lower < 0 || count_upper < 0) return FEA_ERR_INVALID;
  const bool lo = lower_block != nullptr && count_lower > 0, hi = upper_block != nullptr && count_upper > 0;
  if (lo && (!src_lower || !dst_lower || (count_lower & 1) || ((uintptr_t)src_lower & 15) || ((uintptr_t)dst_lower & 15)))
    return FEA_ERR_INVALID;
  if (hi && (!src_upper || !dst_upper || (count_upper & 1) || ((uintptr_t)src_upper & 15) || ((uintptr_t)dst_upper & 15)))
    return FEA_ERR_INVALID;
  if (!lo && !hi) return FEA_OK;
  const int64_t n2 = std::max<int64_t>(lo ? count_lower / 2 : 0, hi ? count_upper / 2 : 0);
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n2, kPushThreads * 4), 296));
  peer_push_kernel<<<blocks, kPushThreads, 0, stream>>>(
      static_cast<CommHeader*>(own_block), lo ? static_cast<CommHeader*>(lower_block) : nullptr,
      reinterpret_cast<const double2*>(src_lower), reinterpret_cast<double2*>(dst_lower), count_lower / 2,
      hi ? static_cast<CommHeader*>(upper_block) : nullptr, reinterpret_cast<const double2*>(src_upper),
      reinterpret_cast<double2*>(dst_upper), count_upper / 2, iter_dev, (long long)epoch);
  return check_launch();
}

extern "C" int fea_peer_wait(void* own_block, int32_t has_lower, int32_t has_upper, const int32_t* iter_dev,
                             int64_t epoch, void* stream_) {
  if (!own_block || !iter_dev) return FEA_ERR_INVALID;
  if (!has_lower && !has_upper) return FEA_OK;
  peer_wait_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<CommHeader*>(own_block), has_lower,
                                                                   has_upper, iter_dev, (long long)epoch);
  return check_launch();
}

extern "C" int fea_comm_error(void* own_block, int32_t* error_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!own_block || !error_host) return FEA_ERR_INVALID;
  int err = 0;
  FEA_TRY(check(cudaMemcpyAsync(&err, &static_cast<CommHeader*>(own_block)->error, sizeof(int), cudaMemcpyDeviceToHost,
                                stream)));
  FEA_TRY(check(cudaStreamSynchronize(stream)));
  *error_host = err;
  return FEA_OK;
}

extern "C" size_t fea_comm_bytes(int64_t n_local_dof) {
  return kCommHeaderBytes + align_up(sizeof(double) * (size_t)n_local_dof, 256);
}

extern "C" int fea_comm_alloc(size_t bytes, void** out) {
  if (!out || bytes < kCommHeaderBytes) return FEA_ERR_INVALID;
  FEA_TRY(check(cudaMalloc(out, bytes)));
  return check(cudaMemset(*out, 0, bytes));
}

extern "C" int fea_comm_free(void* ptr) { return check(cudaFree(ptr)); }

extern "C" int fea_comm_ipc_export(void* ptr, unsigned char* handle64) {
  if (!ptr || !handle64) return FEA_ERR_INVALID;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  FEA_TRY(check(cudaIpcGetMemHandle(&h, ptr)));
  std::memcpy(handle64, &h, 64);
  return FEA_OK;
}

extern "C" int fea_comm_ipc_open(const unsigned char* handle64, void** out) {
  if (!handle64 || !out) return FEA_ERR_INVALID;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  return check(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
}

extern "C" int fea_comm_ipc_close(void* ptr) { return check(cudaIpcCloseMemHandle(ptr)); }

extern "C" int fea_pcg_solve_p2p(int64_t n_owned_nodes, int32_t d, const int32_t* node_rowptr_owned,
                                 const int32_t* node_colidx, const double* values, int32_t max_coupled,
                                 const double* dinv, const double* b, double* x, double tol, int32_t max_iter,
                                 void* work, size_t work_bytes, double* history, const fea_peer_comm* comm,
                                 fea_pcg_result* result_host, void* stream_) {
  cudaStream_t caller = static_cast<cudaStream_t>(stream_);
  if (!node_rowptr_owned || !node_colidx || !values || !dinv || !b || !x || !work || !comm || !result_host)
    return FEA_ERR_INVALID;
  if (n_owned_nodes <= 0 || d < 1 || d > 3 || max_iter < 1) return FEA_ERR_INVALID;
  if (comm->world < 1 || comm->world > kMaxPeers || comm->rank < 0 || comm->rank >= comm->world) return FEA_ERR_INVALID;
  const int64_t n = n_owned_nodes * d;
  if (work_bytes < fea_pcg_workspace(n)) return FEA_ERR_WORKSPACE;

  // workspace: state | partials | r | (p lives in the comm block) | ap
  char* c = static_cast<char*>(work);
  PcgState* state = reinterpret_cast<PcgState*>(c);
  c += FEA_PCG_STATE_BYTES;
  double* partials = reinterpret_cast<double*>(c);
  c += sizeof(double) * 2 * kMaxPartials;
  const size_t vec = align_up(sizeof(double) * (size_t)n, 256);
  double* r = reinterpret_cast<double*>(c);
  c += vec;
  double* ap = reinterpret_cast<double*>(c);
  c += vec;
  double* p2 = reinterpret_cast<double*>(c);  // single-reduction variant: search direction and s = K p
  c += vec;
  double* s_vec = reinterpret_cast<double*>(c);

  PeerView pv;
  std::memset(&pv, 0, sizeof(pv));
  pv.world = comm->world;
  pv.rank = comm->rank;
  pv.lower = comm->lower_peer;
  pv.upper = comm->upper_peer;
  pv.epoch = comm->epoch;
  // The recurrence must be the same on every rank: chosen from a rank-invariant size (slabs are uneven,
  // a per-rank choice could straddle the threshold), or given by the caller; verified in the first exchange.
  const bool multi = comm->world > 1;
  const int algo = comm->algo == 0 || comm->algo == 1
                       ? comm->algo
                       : pcg_algorithm(comm->max_rank_dof > 0 ? comm->max_rank_dof : n, multi);
  pv.algo = algo;
  const int64_t n_tiles = ceil_div(n_owned_nodes, kTileNodes);
  pv.lower_tiles = (int)std::min<int64_t>(n_tiles, ceil_div(std::max<int64_t>(comm->boundary_lower_nodes, 0), kTileNodes));
  {
    // tiles that contain any of the last boundary_upper_nodes owned nodes
    const int64_t first_node = n_owned_nodes - std::min<int64_t>(std::max<int64_t>(comm->boundary_upper_nodes, 0), n_owned_nodes);
    pv.upper_tiles = comm->boundary_upper_nodes > 0 ? (int)(n_tiles - first_node / kTileNodes) : 0;
  }
  if (std::getenv("FEA_P2P_FAKE_TILES") == nullptr) {  // (experiment: keep face tiles without neighbours)
    if (pv.lower < 0) pv.lower_tiles = 0;
    if (pv.upper < 0) pv.upper_tiles = 0;
  }
  // experiment switches: FEA_HALO_GATE=0 waits for the halo before the FIRST tile (no
  // overlap, the round-1 behaviour moved into the SpMV); FEA_P2P_FORCE_GATED=1 runs the gated kernels on
  // a single rank too (A/B of the gated against the plain SpMV on one GPU)
  if (const char* env = std::getenv("FEA_HALO_GATE")) {
    if (env[0] == '0') pv.lower_tiles = pv.upper_tiles = (int)n_tiles;
  }
  const bool debug = std::getenv("FEA_P2P_DEBUG") != nullptr;
  {
    const int on = debug ? 1 : 0;
    CommHeader* hdr0 = static_cast<CommHeader*>(comm->comm[comm->rank]);
    cudaMemcpy(&hdr0->dbg_on, &on, sizeof(int), cudaMemcpyHostToDevice);
    const unsigned long long lo = ~0ULL, hi = 0ULL;
    cudaMemcpy(&hdr0->dbg_cta_first, &lo, sizeof(lo), cudaMemcpyHostToDevice);
    cudaMemcpy(&hdr0->dbg_cta_last, &hi, sizeof(hi), cudaMemcpyHostToDevice);
  }
  const char* force_env = std::getenv("FEA_P2P_FORCE_GATED");
  const bool force_gated = force_env != nullptr && force_env[0] == '1';
  for (int i = 0; i < comm->world; ++i) {
    if (!comm->comm[i]) return FEA_ERR_INVALID;
    pv.hdr[i] = static_cast<CommHeader*>(comm->comm[i]);
  }
  auto p_ext_of = [&](int rank) {
    return reinterpret_cast<double*>(static_cast<char*>(comm->comm[rank]) + kCommHeaderBytes);
  };
  double* p_ext = p_ext_of(comm->rank);
  double* p_own = p_ext + comm->own_offset_nodes * d;
  if (pv.lower >= 0) {
    pv.lower_off = comm->send_lower_first * d;
    pv.lower_src = p_own + pv.lower_off;
    pv.lower_cnt = comm->send_lower_count * d;
    pv.lower_dst = p_ext_of(pv.lower) + comm->send_lower_dst * d;
  }
  if (pv.upper >= 0) {
    pv.upper_off = comm->send_upper_first * d;
    pv.upper_src = p_own + pv.upper_off;
    pv.upper_cnt = comm->send_upper_count * d;
    pv.upper_dst = p_ext_of(pv.upper) + comm->send_upper_dst * d;
  }
  // device copy of the view inside this rank's own header page (peers only write the slot arrays)
  const PeerView* pv_dev =
      reinterpret_cast<const PeerView*>(static_cast<char*>(comm->comm[comm->rank]) + kCommViewOffset);
  const TmaPlan plan = tma_plan(d, max_coupled, values, node_colidx, n_owned_nodes);

  PcgState* snap = static_cast<PcgState*>(pinned_scratch(0, 2 * sizeof(PcgState)));
  if (snap == nullptr) return FEA_ERR_CUDA;
  cudaEvent_t ev[2] = {nullptr, nullptr}, ev_order = nullptr;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int rc = check(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  if (rc == FEA_OK) rc = check(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  if (rc == FEA_OK) rc = check(cudaEventCreateWithFlags(&ev_order, cudaEventDisableTiming));
  if (rc == FEA_OK) rc = check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  if (rc == FEA_OK) rc = check(cudaEventRecord(ev_order, caller));
  if (rc == FEA_OK) rc = check(cudaStreamWaitEvent(stream, ev_order, 0));

  const unsigned vb = vec_blocks(n);
  const unsigned halo_blocks = 16;
  auto exchange = [&]() {
    if (multi) p2p_init_exchange_kernel<<<1, 32, 0, stream>>>(pv, state);
  };
  auto halo = [&]() {
    if (multi) p2p_halo_kernel<<<halo_blocks, kHaloThreads, 0, stream>>>(pv, state);
  };
  PeerLaunch peer;
  peer.view = pv_dev;
  peer.key.own = pv.hdr[pv.rank];
  peer.key.epoch = pv.epoch;
  peer.key.world = pv.world;
  peer.key.rank = pv.rank;
  peer.key.lower_tiles = pv.lower_tiles;
  peer.key.upper_tiles = pv.upper_tiles;
  peer.key.has_lower = pv.lower >= 0;
  peer.key.has_upper = pv.upper >= 0;
  {
    const size_t lo_bytes = sizeof(double) * (size_t)d * (size_t)comm->own_offset_nodes;
    const size_t hi_bytes = lo_bytes + sizeof(double) * (size_t)n;
    peer.key.aligned = (reinterpret_cast<uintptr_t>(p_ext) % 128 == 0) && (pv.lower < 0 || lo_bytes % 128 == 0) &&
                       (pv.upper < 0 || hi_bytes % 128 == 0) && std::getenv("FEA_HALO_COHERENT") == nullptr;
  }
  const PeerLaunch* peer_it = multi || force_gated ? &peer : nullptr;
  const PeerView* pv_it = peer_it ? pv_dev : nullptr;
  // measurement hook (as in fea_pcg_solve): CUDA-event pairs around one SpMV launch per chunk
  constexpr int kMaxSamples = 256;
  cudaEvent_t* sample_ev = nullptr;
  int n_samples = 0;
  int sample_iter[kMaxSamples];
  if (profile().enabled) {
    sample_ev = new cudaEvent_t[2 * kMaxSamples];
    for (int i = 0; i < 2 * kMaxSamples; ++i) cudaEventCreate(&sample_ev[i]);
  }
  int64_t enqueued = 0;
  // side stream of the halo push: forks after the vector kernel(s), joins before the next ones
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool join_pending = false;
  if (rc == FEA_OK && multi) {
    // The push must run WHILE the persistent SpMV owns every SM (3 CTAs each, nearly all shared memory):
    // same shared-memory carve-out as the SpMV (an SM cannot hold two carve-outs at once: with the default
    // preference the push only started once the SpMV had drained, and both ranks' face tiles sat waiting
    // for each other's tag -- 705 us per SpMV instead of 440 at N = 2), and the highest stream priority so
    // that its 8 small blocks are dispatched first.
    cudaFuncSetAttribute(p2p_halo_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    rc = check(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
    if (rc == FEA_OK) rc = check(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    if (rc == FEA_OK) rc = check(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  auto join_side = [&]() {  // the vector kernels overwrite the rows the push is still reading
    if (join_pending) cudaStreamWaitEvent(stream, ev_join, 0);
    join_pending = false;
  };
  const char* side_env = std::getenv("FEA_P2P_SIDE");  // experiment switch: 0 = push on the solver's own stream
  const bool use_side = side_env == nullptr || side_env[0] != '0';
  auto fork_halo = [&]() {
    if (!multi) return;
    if (!use_side) {
      p2p_halo_kernel<<<halo_blocks, kHaloThreads, 0, stream>>>(pv, state);
      return;
    }
    cudaEventRecord(ev_fork, stream);
    cudaStreamWaitEvent(side, ev_fork, 0);
    p2p_halo_kernel<<<halo_blocks, kHaloThreads, 0, side>>>(pv, state);
    cudaEventRecord(ev_join, side);
    join_pending = true;
  };
  auto iteration = [&](bool sample) -> int {
    if (sample) sample_iter[n_samples] = (int)enqueued;
    if (sample) cudaEventRecord(sample_ev[2 * n_samples], stream);
    const int r1 = pcg_step_spmv(d, n_owned_nodes, node_rowptr_owned, node_colidx, values, p_ext, ap,
                                 comm->own_offset_nodes, state, partials, stream, &plan, peer_it);
    if (sample) cudaEventRecord(sample_ev[2 * n_samples++ + 1], stream);
    join_side();
    if (algo == 1) {  // p_own holds u = dinv r here (the SpMV input)
      pcg_cgcg_kernel<<<cgcg_blocks(n), 256, 0, stream>>>(n, dinv, p_own, ap, p2, s_vec, x, r, state, partials, history,
                                                          pv_it, peer.key);
    } else {
      pcg_update_kernel<<<vb, 256, 0, stream>>>(n, dinv, p_own, ap, x, r, state, partials, pv_it, peer.key);
      pcg_direction_kernel<<<vb, 256, 0, stream>>>(n, dinv, r, p_own, state, history, pv_it, peer.key);
    }
    fork_halo();
    return r1;
  };
  const int launches_per_iteration = (algo == 1 ? 2 : 3) + (multi ? 1 : 0);
  const int64_t enqueue_limit = (int64_t)max_iter + (algo == 1 ? 1 : 0);

  if (rc == FEA_OK) {
    pcg_match_carveout();
    rc = check(cudaMemsetAsync(state, 0, FEA_PCG_STATE_BYTES, stream));
  }
  if (rc == FEA_OK)
    rc = check(cudaMemcpyAsync(const_cast<PeerView*>(pv_dev), &pv, sizeof(PeerView), cudaMemcpyHostToDevice, stream));
  if (rc == FEA_OK) {
    pcg_init_kernel<<<vb, 256, 0, stream>>>(n, b, dinv, x, r, p_own, tol, max_iter, state, partials);
    exchange();
    halo();
    rc = check_launch(multi ? 3 : 1);
  }
  const int chunk = 32;
  int slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  if (rc == FEA_OK && max_iter >= chunk && std::getenv("FEA_PCG_NO_GRAPH") == nullptr) {
    rc = iteration(false);  // warm-up outside capture
    join_side();
    if (rc == FEA_OK) rc = check_launch(launches_per_iteration);
    enqueued += 1;
    cudaGraph_t graph = nullptr;
    if (rc == FEA_OK && cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      for (int it = 0; it < chunk - 1; ++it) iteration(false);
      join_side();  // the side stream must rejoin the capturing stream before the capture ends
      if (cudaStreamEndCapture(stream, &graph) != cudaSuccess || graph == nullptr ||
          cudaGraphInstantiate(&graph_exec, graph, 0) != cudaSuccess)
        graph_exec = nullptr;
      if (graph != nullptr) cudaGraphDestroy(graph);
    }
    cudaGetLastError();
  }
  while (rc == FEA_OK && !finished) {
    const int todo = (int)std::min<int64_t>(chunk, enqueue_limit - enqueued);
    if (graph_exec != nullptr && todo == chunk) {  // one plain iteration (carries the timing sample) + the graph
      rc = iteration(sample_ev != nullptr && n_samples < kMaxSamples);
      join_side();
      if (rc == FEA_OK) rc = check(cudaGraphLaunch(graph_exec, stream));
    } else {
      for (int it = 0; it < todo && rc == FEA_OK; ++it)
        rc = iteration(sample_ev != nullptr && it == 0 && n_samples < kMaxSamples);
    }
    if (rc == FEA_OK) rc = check_launch(launches_per_iteration * todo);
    if (rc != FEA_OK) break;
    enqueued += todo;
    rc = check(cudaMemcpyAsync(&snap[slot], state, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
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
    if (!finished && enqueued >= enqueue_limit) finished = true;
    slot ^= 1;
  }
  join_side();
  if (rc == FEA_OK) {
    rc = check(cudaMemcpyAsync(&snap[0], state, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
    if (rc == FEA_OK) rc = check(cudaStreamSynchronize(stream));
  }
  if (side != nullptr) cudaStreamSynchronize(side);
  if (debug) {
    unsigned long long h[8][6];
    cudaMemcpy(h, pv.hdr[pv.rank]->dbg, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < kDbgIters; ++i) {
      const double t0 = (double)h[i][0];
      std::fprintf(stderr,
                   "[p2p dbg rank %d it %d] spmv_start 0  push_start %+.1f  push_done %+.1f  "
                   "spmv_done %+.1f  next_spmv_start %+.1f us\n",
                   pv.rank, kDbgFirst + i, ((double)h[i][3] - t0) * 1e-3, ((double)h[i][4] - t0) * 1e-3, ((double)h[i][5] - t0) * 1e-3,
                   i + 1 < kDbgIters ? ((double)h[i + 1][0] - t0) * 1e-3 : 0.0);
    }
  }
  if (rc == FEA_OK) {
    const PcgState& s = snap[0];
    result_host->iterations = s.iter;
    result_host->status = s.status;
    if (!s.done && s.status == FEA_OK) result_host->status = FEA_ERR_MAXITER;
    result_host->bnorm = std::sqrt(s.bnorm2);
    const double rr = s.done ? s.rr_final : s.rr;
    result_host->rel_residual = s.bnorm2 > 0.0 ? std::sqrt(rr / s.bnorm2) : 0.0;
    profile().pcg_iterations += s.iter;
    for (int i = 0; i < n_samples; ++i) {
      if (sample_iter[i] >= s.iter) break;  // launches after convergence are no-ops
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sample_ev[2 * i], sample_ev[2 * i + 1]) == cudaSuccess) {
        profile().add_spmv_sample(ms);
      }
    }
  } else if (stream != nullptr) {
    cudaStreamSynchronize(stream);
  }
  if (sample_ev != nullptr) {
    for (int i = 0; i < 2 * kMaxSamples; ++i) cudaEventDestroy(sample_ev[i]);
    delete[] sample_ev;
  }
  if (graph_exec != nullptr) cudaGraphExecDestroy(graph_exec);
  if (side != nullptr) cudaStreamDestroy(side);
  for (cudaEvent_t e : {ev_fork, ev_join})
    if (e != nullptr) cudaEventDestroy(e);
  if (stream != nullptr) {
    if (ev_order != nullptr && cudaEventRecord(ev_order, stream) == cudaSuccess) cudaStreamWaitEvent(caller, ev_order, 0);
    cudaStreamDestroy(stream);
  }
  for (cudaEvent_t e : {ev[0], ev[1], ev_order})
    if (e != nullptr) cudaEventDestroy(e);
  return rc;
}
