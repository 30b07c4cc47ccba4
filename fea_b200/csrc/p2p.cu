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
//   SpMV       last block: publish this rank's p.Ap to every rank's slot array (one warp, NVLink stores)
//   update     every CTA: collect the world's p.Ap from the own slot array (local L2 polls), add
//              in rank order -> alpha; last block: publish (r.z, r.r)
//   direction  every CTA: collect (r.z, r.r) -> beta / convergence; the boundary rows of the new p
//              are stored straight into the neighbours' halo rows while p is written; last block:
//              release the iteration tag to the neighbours and wait for theirs
//
// The first version used three extra single-purpose kernels per iteration (halo push, two
// all-reduces: 6 launches instead of 3); they remain for the one exchange after the init kernel.
// Every rank forms bitwise identical sums (rank order, no atomics), hence identical convergence
// decisions.  Spins are bounded: a peer that never arrives raises FEA_ERR_PEER instead of hanging.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pcg_common.cuh"

namespace fea {

// kind 0: (rz, bnorm2) after init; kind 1: pap after step 1; kind 2: (rz_new, rr) after step 2.
__global__ void __launch_bounds__(32) p2p_allreduce_kernel(PeerView pv, PcgState* st, int kind) {
  if (st->done) return;
  const int lane = threadIdx.x;
  // iteration this exchange belongs to: step 2 has already incremented st->iter
  const long long k = kind == 2 ? st->iter - 1 : st->iter;
  double mine0, mine1;
  if (kind == 0) {
    mine0 = st->rz;
    mine1 = st->bnorm2;
  } else if (kind == 1) {
    mine0 = st->pap;
    mine1 = 0.0;
  } else {
    mine0 = st->rz_new;
    mine1 = st->rr;
  }
  peer_publish(pv, kind, k, mine0, mine1);
  double s0, s1;
  const bool all_ok = peer_collect(pv, kind, k, s0, s1);
  if (lane == 0) {
    if (!all_ok) {
      peer_failure(pv, st);
    } else if (kind == 0) {
      st->rz = s0;
      st->bnorm2 = s1;
      st->rr = s1;
    } else if (kind == 1) {
      st->pap = s0;
    } else {
      st->rz_new = s0;
      st->rr = s1;
    }
  }
}

__global__ void __launch_bounds__(256) p2p_halo_kernel(PeerView pv, PcgState* st) {
  __shared__ bool s_last;
  if (st->done) return;
  const long long tag = peer_tag(pv, st->iter);
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
  bool ok = true;
  if (pv.lower >= 0) ok = spin_until(&own->halo_tag[0], tag, true) && ok;
  if (pv.upper >= 0) ok = spin_until(&own->halo_tag[1], tag, true) && ok;
  if (!ok) peer_failure(pv, st);
}

}  // namespace fea

using namespace fea;

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
                                 void* work, size_t work_bytes, const fea_peer_comm* comm,
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
  const bool multi = comm->world > 1;
  const unsigned halo_blocks = 8;
  auto exchange = [&](int kind) {
    if (multi) p2p_allreduce_kernel<<<1, 32, 0, stream>>>(pv, state, kind);
  };
  auto halo = [&]() {
    if (multi) p2p_halo_kernel<<<halo_blocks, 256, 0, stream>>>(pv, state);
  };
  const PeerView* pv_it = multi ? pv_dev : nullptr;
  const int algo = pcg_algorithm(n, multi);
  auto iteration = [&]() -> int {
    const int r1 = pcg_step_spmv(d, n_owned_nodes, node_rowptr_owned, node_colidx, values, p_ext, ap,
                                 comm->own_offset_nodes, state, partials, stream, &plan, pv_it);
    if (algo == 1) {  // p_own holds u = dinv r here (the SpMV input)
      pcg_cgcg_kernel<<<cgcg_blocks(n), 256, 0, stream>>>(n, dinv, p_own, ap, p2, s_vec, x, r, state, partials, nullptr,
                                                          pv_it);
    } else {
      pcg_update_kernel<<<vb, 256, 0, stream>>>(n, dinv, p_own, ap, x, r, state, partials, pv_it);
      pcg_direction_kernel<<<vb, 256, 0, stream>>>(n, dinv, r, p_own, state, nullptr, pv_it);
    }
    return r1;
  };
  const int launches_per_iteration = algo == 1 ? 2 : 3;
  const int64_t enqueue_limit = (int64_t)max_iter + (algo == 1 ? 1 : 0);

  if (rc == FEA_OK) {
    pcg_match_carveout();
    rc = check(cudaMemsetAsync(state, 0, FEA_PCG_STATE_BYTES, stream));
  }
  if (rc == FEA_OK)
    rc = check(cudaMemcpyAsync(const_cast<PeerView*>(pv_dev), &pv, sizeof(PeerView), cudaMemcpyHostToDevice, stream));
  if (rc == FEA_OK) {
    pcg_init_kernel<<<vb, 256, 0, stream>>>(n, b, dinv, x, r, p_own, tol, max_iter, state, partials);
    exchange(0);
    halo();
    rc = check_launch(multi ? 3 : 1);
  }
  const int chunk = 32;
  int64_t enqueued = 0;
  int slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  if (rc == FEA_OK && max_iter >= chunk && std::getenv("FEA_PCG_NO_GRAPH") == nullptr) {
    rc = iteration();  // warm-up outside capture
    if (rc == FEA_OK) rc = check_launch(launches_per_iteration);
    enqueued += 1;
    cudaGraph_t graph = nullptr;
    if (rc == FEA_OK && cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      for (int it = 0; it < chunk; ++it) iteration();
      if (cudaStreamEndCapture(stream, &graph) != cudaSuccess || graph == nullptr ||
          cudaGraphInstantiate(&graph_exec, graph, 0) != cudaSuccess)
        graph_exec = nullptr;
      if (graph != nullptr) cudaGraphDestroy(graph);
    }
    cudaGetLastError();
  }
  while (rc == FEA_OK && !finished) {
    const int todo = (int)std::min<int64_t>(chunk, enqueue_limit - enqueued);
    if (graph_exec != nullptr && todo == chunk) {
      rc = check(cudaGraphLaunch(graph_exec, stream));
    } else {
      for (int it = 0; it < todo && rc == FEA_OK; ++it) rc = iteration();
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
  if (rc == FEA_OK) {
    rc = check(cudaMemcpyAsync(&snap[0], state, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
    if (rc == FEA_OK) rc = check(cudaStreamSynchronize(stream));
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
  } else if (stream != nullptr) {
    cudaStreamSynchronize(stream);
  }
  if (graph_exec != nullptr) cudaGraphExecDestroy(graph_exec);
  if (stream != nullptr) {
    if (ev_order != nullptr && cudaEventRecord(ev_order, stream) == cudaSuccess) cudaStreamWaitEvent(caller, ev_order, 0);
    cudaStreamDestroy(stream);
  }
  for (cudaEvent_t e : {ev[0], ev[1], ev_order})
    if (e != nullptr) cudaEventDestroy(e);
  return rc;
}
