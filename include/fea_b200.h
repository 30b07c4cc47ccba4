/*
 * fea_b200.h -- C ABI of libfea_b200.so: the B200 (sm_100a) hot path of jjrreett/fea
 *               element stiffness -> global assembly -> solve of K u = f.
 *
 * The reference has no FFI of its own (it is six numpy scripts); the boundary
 * it offers is its Python callables.  Each entry point below replaces the body of one of them
 * and cites it (file:line in /root/reference).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends
 *     in `_host`; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - the caller owns every buffer; the library allocates nothing persistent;
 *   - all calls are asynchronous on `stream` unless stated otherwise;
 *   - return value: FEA_OK or an FEA_ERR_* code for argument / launch errors.  Data-dependent
 *     errors (detJ <= 0, PCG breakdown) are written to a caller-supplied device status slot
 *     `int32_t status[2]`, zeroed by the caller: status[0] = code, status[1] = 0x7fffffff minus
 *     the smallest offending element index (so that a zeroed slot needs no other initialisation).
 *   - node coordinates: double (n_nodes, 3) row-major (cubebeam.py:45, utils.py:359);
 *     connectivity: int32 (n_elem, nodes_per_elem) row-major, local order as the reference
 *     (utils.py:351-353); global DOF = dof_per_node * node + component (cubebeam.py:86).
 */
#ifndef FEA_B200_H
#define FEA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  FEA_OK = 0,
  FEA_ERR_INVALID = 1,   /* bad argument (null pointer, size overflow, unsupported valence) */
  FEA_ERR_CUDA = 2,      /* a CUDA runtime call or launch failed */
  FEA_ERR_JACOBIAN = 3,  /* detJ <= 0 in a hex8 element (utils.py:212-215 -> ValueError) */
  FEA_ERR_BREAKDOWN = 4, /* PCG p.Ap <= 0: K_ff not positive definite (cubebeam.py:98 -> LinAlgError) */
  FEA_ERR_MAXITER = 5,   /* PCG hit max_iter before the tolerance */
  FEA_ERR_WORKSPACE = 6, /* caller workspace too small */
  FEA_ERR_DEGENERATE = 7, /* zero-length truss member */
  FEA_ERR_PEER = 8,       /* multi-GPU: a peer rank never delivered its halo / partial sum */
  FEA_ERR_STAGNATION = 9  /* PCG residual stopped improving: singular (under-constrained) reduced system
                             with an inconsistent load (cubebeam.py:98 -> LinAlgError) */
};

/* Library / build identification ("fea_b200 <version> sm_100a"). */
const char* fea_version(void);
/* cudaGetLastError/cudaGetErrorString of the last failing CUDA call made by this library. */
const char* fea_last_cuda_error(void);

/* Measurement hooks (bench.py).  fea_profile_enable(1) zeroes the counters and makes
 * fea_pcg_solve time a sample of its SpMV launches with CUDA events on the solve's own stream
 * (one launch per 32-iteration chunk, at most 256 per solve).  fea_profile_read fills
 * out_host[4] (HOST) = {kernels launched by the library, SpMV launches timed, sum of their
 * durations in ms, PCG iterations executed}. */
void fea_profile_enable(int32_t enable);
void fea_profile_read(double* out_host);

/* ------------------------------------------------------------------------------------------
 * (1) Batched FP64 element stiffness, materialised (used by the element-level API and tests;
 *     the assembly path below never writes Ke to HBM).
 * ---------------------------------------------------------------------------------------- */

/* hexahedral_stiffness_matrix(nodes, E, nu) -> (24,24), utils.py:127-239, batched over elements.
 * ke: (n_elem, 24, 24) row-major.  detJ <= 0 sets status = {FEA_ERR_JACOBIAN, first element}. */
int fea_ke_hex8(const double* nodes, const int32_t* elements, int64_t n_elem, double E, double nu,
                double* ke, int32_t* status, void* stream);

/* element_stiffness_matrix of euler_bernoulli.py:22-39 with per-element EI and length.
 * ke: (n_elem, 4, 4). */
int fea_ke_beam(const double* EI, const double* length, int64_t n_elem, double* ke, void* stream);

/* Linearised pin-jointed member (SURVEY.md T1', tangent of truss.py:78-92 at rest):
 * Ke = k [[cc^T, -cc^T], [-cc^T, cc^T]].  ke: (n_elem, 6, 6). */
int fea_ke_truss(const double* nodes, const int32_t* members, const double* k, int64_t n_elem,
                 double* ke, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2a) Symbolic assembly: connectivity -> deterministic CSR pattern.
 *      Replaces the dense `np.zeros((ndof, ndof))` + `np.ix_` bookkeeping of cubebeam.py:80-90.
 *      The pattern is STRUCTURAL (every (row, col) pair of every element) and equals
 *      scipy's coo->csr + sum_duplicates + sort_indices bit for bit.
 *
 *      Outputs are kept at node-block level (one entry per coupled node pair) because every
 *      DOF row of a node shares the node's column set:
 *        row d*i+a of the DOF-level CSR = { d*j+b : j in node_colidx[node_rowptr[i]..), b<d }
 *      and its values start at  d*d*node_rowptr[i] + a*d*cnt_i.  `fea_csr_expand` writes the
 *      ordinary DOF-level rowptr / colidx from that.
 * ---------------------------------------------------------------------------------------- */

/* Bytes of scratch `fea_csr_symbolic_count` / `_fill` need. */
size_t fea_csr_symbolic_workspace(int64_t n_nodes, int64_t n_elem, int32_t nodes_per_elem);

/* Phase 1 (synchronises `stream` once to return sizes).
 *   n2e_ptr  [n_nodes+1]            out: node -> incident (element, local node) list offsets
 *   n2e      [n_elem*nodes_per_elem] out: entries e*nodes_per_elem + a, ascending per node
 *   node_rowptr [n_nodes+1]         out: offsets of each node's coupled-node list
 *   sizes_host[4]                   out (HOST): {nnz_blocks, max coupled nodes per node,
 *                                                max incident elements per node, 0} */
int fea_csr_symbolic_count(const int32_t* elements, int64_t n_elem, int32_t nodes_per_elem,
                           int64_t n_nodes, int32_t* n2e_ptr, int32_t* n2e, int32_t* node_rowptr,
                           int64_t* sizes_host, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Phase 2: node_colidx [nnz_blocks] out, ascending within each node row. */
int fea_csr_symbolic_fill(const int32_t* elements, int64_t n_elem, int32_t nodes_per_elem,
                          int64_t n_nodes, const int32_t* n2e_ptr, const int32_t* n2e,
                          const int32_t* node_rowptr, int32_t* node_colidx, int32_t max_incident,
                          void* stream);

/* DOF-level CSR arrays: rowptr [n_nodes*d + 1], colidx [d*d*nnz_blocks]. */
int fea_csr_expand(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                   const int32_t* node_colidx, int32_t* rowptr, int32_t* colidx, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2b) Numeric assembly, fused with element stiffness evaluation (Ke never reaches HBM) and
 *      with Dirichlet handling.  Owner-computes: one warp owns one node's d rows, walks the
 *      node's incident elements in ascending element order (the reference's summation order,
 *      cubebeam.py:82-90) and writes each CSR value exactly once -- no atomics, bit-reproducible.
 *
 *      values  [d*d*nnz_blocks]  out, DOF-level CSR order (see 2a)
 *      dinv    [n_nodes*d]       out, may be NULL: Jacobi preconditioner 1/K_ii, and 0 on
 *                                constrained DOF (this is how the solver eliminates them)
 *      fixed   [n_nodes*d] uint8 in, may be NULL: 1 = constrained (cubebeam.py:92, homogeneous)
 *      mode    FEA_ASSEMBLE_FULL: K as assembled (needed for reactions, cubebeam.py:106)
 *              FEA_ASSEMBLE_ELIMINATED: constrained rows/columns replaced by identity
 *
 *      fea_assemble_hex8 evaluates the geometry of an element ONCE, in stream-ordered scratch of the
 *      device's default memory pool (80 B per element; 640 B more per element if the mesh has elements
 *      that are not exactly affine, i.e. need the 2x2x2 quadrature of utils.py:200-237 point by point;
 *      exactly affine elements take its closed form).  It synchronises `stream` once (a counter tells
 *      the host whether the Gauss-point pass is needed).
 * ---------------------------------------------------------------------------------------- */
enum { FEA_ASSEMBLE_FULL = 0, FEA_ASSEMBLE_ELIMINATED = 1 };

int fea_assemble_hex8(const double* nodes, const int32_t* elements, int64_t n_elem, int64_t n_nodes,
                      double E, double nu, const int32_t* n2e_ptr, const int32_t* n2e,
                      const int32_t* node_rowptr, const int32_t* node_colidx, int32_t max_coupled,
                      const uint8_t* fixed, int32_t mode, double* values, double* dinv,
                      int32_t* status, void* stream);

int fea_assemble_beam(const double* EI, const double* length, const int32_t* elements,
                      int64_t n_elem, int64_t n_nodes, const int32_t* n2e_ptr, const int32_t* n2e,
                      const int32_t* node_rowptr, const int32_t* node_colidx, const uint8_t* fixed,
                      int32_t mode, double* values, double* dinv, void* stream);

int fea_assemble_truss(const double* nodes, const int32_t* members, const double* k, int64_t n_elem,
                       int64_t n_nodes, const int32_t* n2e_ptr, const int32_t* n2e,
                       const int32_t* node_rowptr, const int32_t* node_colidx, const uint8_t* fixed,
                       int32_t mode, double* values, double* dinv, int32_t* status, void* stream);

/* dinv[i] = fixed[i] ? 0 : 1 / K_ii from an assembled block-pattern matrix. */
int fea_jacobi_dinv(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                    const int32_t* node_colidx, const double* values, const uint8_t* fixed,
                    double* dinv, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3) Sparse matrix-vector products and the Jacobi-PCG solver (replaces np.linalg.solve,
 *     cubebeam.py:98, fea.py:105, euler_bernoulli.py:69; and K @ u, cubebeam.py:106).
 * ---------------------------------------------------------------------------------------- */

/* y = K x on the node-block pattern (dof_per_node in {1,2,3}; 1 = ordinary CSR with
 * node_rowptr/node_colidx = rowptr/colidx). */
int fea_spmv(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
             const int32_t* node_colidx, const double* values, int32_t max_coupled, const double* x,
             double* y, void* stream);
/* `max_coupled` (here and below): the largest number of coupled nodes of any node row, as
 * returned by fea_csr_symbolic_count.  It sizes the shared-memory stages of the TMA bulk-copy
 * SpMV pipeline (cp.async.bulk + mbarrier); pass 0 if unknown to use the plain-load kernel. */

/* Y = K X for n_rhs right-hand sides, X and Y (n_dof, n_rhs) row-major. */
int fea_spmm(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
             const int32_t* node_colidx, const double* values, const double* X, double* Y,
             int32_t n_rhs, void* stream);

/* Device-resident solver state: FEA_PCG_STATE_BYTES bytes (layout below). */
#define FEA_PCG_STATE_BYTES 256
size_t fea_pcg_workspace(int64_t n_dof);

typedef struct {
  int32_t iterations; /* PCG iterations performed */
  int32_t status;     /* FEA_OK, FEA_ERR_BREAKDOWN or FEA_ERR_MAXITER */
  double rel_residual;/* recurrence ||r|| / ||b|| over free DOF at exit */
  double bnorm;       /* ||b|| over free DOF */
} fea_pcg_result;

/* Solve K_ff u_f = f_f by Jacobi-PCG; constrained DOF are the ones with dinv == 0: their u stays 0
 * and their rows/columns never enter (equivalent to the reference's reduced system,
 * cubebeam.py:92-104).  x0 = 0.  Stops when ||r||_2 <= tol * ||b||_2 (recurrence residual).
 *   b [n_dof] in; x [n_dof] out; work: fea_pcg_workspace(n_dof) bytes of device scratch;
 *   result_host: HOST pointer (pinned preferred), filled before return (the call synchronises).
 *   history, may be NULL: device array [max_iter] receiving ||r||/||b|| per iteration. */
int fea_pcg_solve(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                  const int32_t* node_colidx, const double* values, int32_t max_coupled,
                  const double* dinv, const double* b, double* x, double tol, int32_t max_iter,
                  void* work, size_t work_bytes, double* history, fea_pcg_result* result_host,
                  void* stream);

/* The three kernels of one PCG iteration, exposed so that a multi-GPU driver can interleave its
 * halo exchange and all-reduces.  `state` is FEA_PCG_STATE_BYTES of device memory; viewed as
 * double[] the scalars sit at the FEA_PCG_* indices below, viewed as int32[] the counters sit at
 * the FEA_PCG_*_I32 indices.  `partials`: 2*FEA_PCG_PARTIALS doubles of reduction scratch.
 *   init  : x = 0, r = b on free DOF, p = dinv r;  state.rz, state.bnorm2 (local sums)
 *   step 1: ap = K p over the owned nodes;          state.pap    = sum_owned p.ap
 *   step 2: x += a p; r -= a ap (a = rz/pap);       state.rz_new = sum r.dinv.r; state.rr = sum_free r.r;
 *           iteration += 1
 *   step 3: if rr <= tol^2 bnorm2: done = 1; else p = dinv r + (rz_new/rz) p, rz = rz_new
 * A multi-rank driver all-reduces [RZ, BNORM2] after init, [PAP] after step 1 and [RZ_NEW, RR]
 * after step 2 (adjacent doubles).  Every step is a no-op once state.done != 0.
 * In step 1 `p` may carry halo entries: columns index p directly, the owned rows start at node
 * `p_row_offset` of p. */
enum {
  FEA_PCG_RZ = 0, FEA_PCG_BNORM2 = 1, FEA_PCG_RZ_NEW = 2, FEA_PCG_RR = 3, FEA_PCG_PAP = 4,
  FEA_PCG_TOL2 = 5, FEA_PCG_RR_FINAL = 6 /* ||r||^2 frozen at the moment `done` was set */,
  FEA_PCG_ITER_I32 = 32, FEA_PCG_DONE_I32 = 33, FEA_PCG_STATUS_I32 = 34, FEA_PCG_MAXITER_I32 = 35
};
#define FEA_PCG_PARTIALS 2048
int fea_pcg_init(int64_t n_dof, const double* b, const double* dinv, double* x, double* r, double* p,
                 double tol, int32_t max_iter, void* state, void* partials, void* stream);
int fea_pcg_step_spmv(int64_t n_owned_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                      const int32_t* node_colidx, const double* values, int32_t max_coupled,
                      const double* p, double* ap, int64_t p_row_offset, void* state, void* partials,
                      void* stream);
int fea_pcg_step_update(int64_t n_dof, const double* dinv, const double* p, const double* ap,
                        double* x, double* r, void* state, void* partials, void* stream);
/* history, may be NULL: device array [max_iter], entry it-1 receives ||r||/||b|| of iteration it. */
int fea_pcg_step_direction(int64_t n_dof, const double* dinv, const double* r, double* p,
                           void* state, double* history, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3b) Multi-GPU Jacobi-PCG over NVLink peer memory (one process per GPU, 1-D slab partition).
 *      Every rank allocates one communication block (header + its halo-extended p vector) with
 *      fea_comm_alloc, exports it with fea_comm_ipc_export, and maps its peers' blocks with
 *      fea_comm_ipc_open (CUDA IPC; the 64-byte handles travel through any host channel, e.g.
 *      torch.distributed.all_gather_object).  fea_pcg_solve_p2p then runs the iteration as one CUDA
 *      graph per 32 iterations, without NCCL: the dot products are exchanged through the peers' slot
 *      arrays and the halo rows of p are stored into the neighbours' memory from inside the three
 *      solver kernels (3 launches per iteration; csrc/p2p.cu, csrc/pcg_common.cuh).  The last 512
 *      bytes of the 4 KiB header of a block are reserved for the library (device copy of the view).
 * ---------------------------------------------------------------------------------------- */
#define FEA_MAX_PEERS 8
typedef struct {
  int32_t world, rank;
  int32_t lower_peer, upper_peer;  /* neighbour ranks, -1 if none */
  void* comm[FEA_MAX_PEERS];       /* comm[r]: rank r's block as mapped in this process (comm[rank] = own) */
  int64_t own_offset_nodes;        /* owned rows start at this node of the local (halo-extended) range */
  /* my owned nodes [first, first+count) (relative to the first owned node) go to node `dst` of the
   * neighbour's local range */
  int64_t send_lower_first, send_lower_count, send_lower_dst;
  int64_t send_upper_first, send_upper_count, send_upper_dst;
  int64_t epoch;                   /* distinct for every solve that reuses the blocks (same on all ranks) */
  /* Owned nodes [0, boundary_lower_nodes) couple to lower-halo nodes and the last boundary_upper_nodes
   * owned nodes to upper-halo nodes: the SpMV sweeps the other rows first and only then looks at the
   * neighbours' halo tags (the exchange hides behind the interior of the slab). */
  int64_t boundary_lower_nodes, boundary_upper_nodes;
  /* PCG recurrence, identical on every rank (checked in the first exchange): 0 classical, 1 single
   * reduction (Chronopoulos-Gear), -1 = choose from max_rank_dof (the LARGEST owned DOF count of any
   * rank -- a rank-invariant quantity; slabs are uneven). */
  int32_t algo, reserved;
  int64_t max_rank_dof;
} fea_peer_comm;

/* Host-side partition helper: one multi-threaded pass over the GLOBAL connectivity in HOST memory
 * (int64 or int32 node ids, index_bytes = 8 | 4).  For each node range r = [ranges_host[2r],
 * ranges_host[2r+1]) it reports, in out_host[5r..5r+5): the number of elements with at least one
 * node in the range, the first and last such element id (-1 if none), and the smallest / largest
 * node id those elements reference.  A rank's slab = its owned range; its halo = [min, max] beyond
 * it; what it sends = the neighbours' ranges scanned the same way (fea_b200/dist.py:plan_slab). */
int fea_slab_scan(const void* elements_host, int32_t index_bytes, int64_t n_elem, int32_t nodes_per_elem,
                  const int64_t* ranges_host, int32_t n_ranges, int64_t* out_host);

size_t fea_comm_bytes(int64_t n_local_dof);
int fea_comm_alloc(size_t bytes, void** out);  /* cudaMalloc + zero */
int fea_comm_free(void* ptr);
int fea_comm_ipc_export(void* ptr, unsigned char* handle64_host);
int fea_comm_ipc_open(const unsigned char* handle64_host, void** out);
int fea_comm_ipc_close(void* ptr);

/* Stand-alone halo exchange over the same blocks, for solvers driven from the host side (the many-load-case
 * PCG on slabs, BASELINE config 5: fea_b200/dist.py distributed_pcg_multi).  Replaces a pair of NCCL
 * send / recv per neighbour and iteration.
 * fea_peer_push: copy `count_*` doubles (multiples of 2, 16-byte aligned) from this rank's vector into the
 *   neighbours' vectors (peer pointers from fea_comm_ipc_open), one system-scope fence, then the tag
 *   (epoch << 32 | *iter_dev + 1) into the neighbours' headers.  A null destination skips that side.
 * fea_peer_wait: hold the stream until both neighbours' tags for (epoch, *iter_dev + 1) have arrived in this
 *   rank's header (bounded, ~2 s: then the header's error word is set and fea_comm_error returns it).
 * `own_block` / `lower_block` / `upper_block` are block base pointers (header first); `iter_dev` is the
 * solver's device-resident iteration counter, so both calls can sit in a replayed CUDA graph.
 * Reuse of the halo rows needs no second handshake when a world-wide reduction separates the reader's SpMM
 * from the next push, as in PCG. */
int fea_peer_push(void* own_block, void* lower_block, const double* src_lower, double* dst_lower, int64_t count_lower,
                  void* upper_block, const double* src_upper, double* dst_upper, int64_t count_upper,
                  const int32_t* iter_dev, int64_t epoch, void* stream);
int fea_peer_wait(void* own_block, int32_t has_lower, int32_t has_upper, const int32_t* iter_dev, int64_t epoch,
                  void* stream);
/* error word of a block's header (device -> host copy, synchronises `stream`): 0, or 1 after a timed-out wait */
int fea_comm_error(void* own_block, int32_t* error_host, void* stream);

/* Arguments as fea_pcg_solve, restricted to this rank's owned rows: node_rowptr_owned points at
 * the first owned node's entry of the slab's node_rowptr (entries stay absolute), dinv / b / x have
 * n_owned_nodes * dof_per_node entries.  history as in fea_pcg_solve (every rank records the same
 * world residuals).  Synchronises; identical result on every rank.
 * result_host->status may also be FEA_ERR_PEER (a peer never arrived, or the ranks disagree on the
 * recurrence). */
int fea_pcg_solve_p2p(int64_t n_owned_nodes, int32_t dof_per_node, const int32_t* node_rowptr_owned,
                      const int32_t* node_colidx, const double* values, int32_t max_coupled,
                      const double* dinv, const double* b, double* x, double tol, int32_t max_iter,
                      void* work, size_t work_bytes, double* history, const fea_peer_comm* comm,
                      fea_pcg_result* result_host, void* stream);

/* Step-level entry points of the batched solver, for a multi-GPU driver (fea_b200/dist.py:
 * distributed_pcg_multi): the kernels of fea_pcg_solve_multi on one rank's slab of rows; the caller sums
 * the per-column scalars over the ranks in between (any all-reduce; NCCL in dist.py):
 *   fea_pcg_multi_init             -> sum scalars[0 .. 2R) (r.z, ||b||^2), then fea_pcg_multi_activate
 *   fea_pcg_multi_step_spmm        -> sum scalars[4R .. 5R) (p.Ap); P_ext is the halo-extended
 *                                     (n_local_dof, R) array, owned rows start at node p_row_offset
 *   fea_pcg_multi_step_update      -> sum scalars[2R .. 4R) (new r.z, r.r)
 *   fea_pcg_multi_step_direction      convergence bookkeeping on the world sums (identical everywhere)
 * The reducing kernels write their column sums to a LOCAL set of the 5R scalars; the caller copies the
 * local set over the WORLD set and all-reduces that (idempotent when a finished solve returns early).
 * `work`: fea_pcg_multi_workspace(n_owned_dof, R) bytes.  fea_pcg_multi_layout fills offsets_host[5], the
 * byte offsets in it of {state (int32: iter, done, status, max_iter, n_active, n_rhs), the 5R world doubles
 * rz | bnorm2 | rz_new | rr | pap, active (R int32), iterations per column (R int32), the 5R local doubles}. */
int fea_pcg_multi_layout(int32_t n_rhs, int64_t* offsets_host);
int fea_pcg_multi_init(int64_t n_dof, int32_t n_rhs, const double* B, const double* dinv, double* X,
                       double* P_own, double tol, int32_t max_iter, void* work, size_t work_bytes,
                       void* stream);
int fea_pcg_multi_activate(int64_t n_dof, int32_t n_rhs, void* work, void* stream);
int fea_pcg_multi_step_spmm(int64_t n_owned_nodes, int32_t dof_per_node, const int32_t* node_rowptr_owned,
                            const int32_t* node_colidx, const double* values, const double* P_ext,
                            int64_t p_row_offset, int32_t n_rhs, void* work, void* stream);
int fea_pcg_multi_step_update(int64_t n_dof, int32_t n_rhs, const double* dinv, const double* P_own,
                              double* X, void* work, void* stream);
int fea_pcg_multi_step_direction(int64_t n_dof, int32_t n_rhs, const double* dinv, double* P_own,
                                 void* work, void* stream);

/* Direct solve of a CHAIN mesh (every node couples to its two neighbours only: the Euler-Bernoulli beam
 * of euler_bernoulli.py:42-73, dof_per_node = 2; or 1): block-tridiagonal parallel cyclic reduction,
 * ceil(log2 n) steps, replaces `np.linalg.solve` (euler_bernoulli.py:69) where Jacobi-PCG cannot
 * (cond(K) ~ n^4).  `fixed` (uint8 per DOF, may be NULL) marks the homogeneous constraints
 * (euler_bernoulli.py:61-66): x is exactly 0 there.  status = {FEA_ERR_INVALID, node} if the pattern is
 * not a chain, {FEA_ERR_BREAKDOWN, node} for a singular pivot block (under-constrained beam).
 * extended != 0: the elimination runs in double-double arithmetic (~106 bits) on the same FP64 matrix
 * and right-hand side, x is rounded back to FP64 -- cond(K) ~ 5 n^4 reaches 5e20 at the 100 k elements
 * of BASELINE config 2, where FP64 elimination returns noise. */
size_t fea_chain_solve_workspace(int64_t n_nodes, int32_t dof_per_node, int32_t extended);
int fea_chain_solve(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                    const int32_t* node_colidx, const double* values, const uint8_t* fixed,
                    const double* b, double* x, int32_t extended, void* work, size_t work_bytes,
                    int32_t* status, void* stream);

/* Batched multi-RHS Jacobi-PCG (BASELINE config 5): n_rhs <= 256 independent systems sharing K,
 * each column with its own alpha/beta and stopping rule.  B, X: (n_dof, n_rhs) row-major.
 * iterations_host [n_rhs] (HOST, may be NULL): iterations each column needed.
 * result_host: iterations = total performed, rel_residual = worst column, bnorm = largest ||b_j||. */
size_t fea_pcg_multi_workspace(int64_t n_dof, int32_t n_rhs);
int fea_pcg_solve_multi(int64_t n_nodes, int32_t dof_per_node, const int32_t* node_rowptr,
                        const int32_t* node_colidx, const double* values, const double* dinv,
                        const double* B, double* X, int32_t n_rhs, double tol, int32_t max_iter,
                        void* work, size_t work_bytes, int32_t* iterations_host,
                        fea_pcg_result* result_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * (4) Adjacent steps of the reference scripts ("next" rows of SURVEY.md §8(f)).
 * ---------------------------------------------------------------------------------------- */

/* compute_forces(nodes, members, displaced_nodes, forces), truss.py:78-92: accumulates member
 * forces into `forces` (n_nodes,3) IN PLACE.  k: per-member spring rate.  Node-parallel over the
 * node -> member incidence lists of the symbolic pass (n2m_ptr / n2m = n2e_ptr / n2e of
 * fea_csr_symbolic_count on the (n_members, 2) connectivity): every node adds its members'
 * contributions in ascending member order, the order of the reference's loop -- deterministic,
 * no floating-point atomics.  fp32 != 0: nodes / k / displaced / forces are float (the script's own
 * dtype, truss.py:9-10), else double. */
int fea_truss_member_forces(const void* nodes, const int32_t* members, const void* k,
                            int64_t n_members, int64_t n_nodes, const int32_t* n2m_ptr,
                            const int32_t* n2m, const void* displaced, void* forces, int32_t fp32,
                            void* stream);

/* The relaxation loop of truss.py:95-119, n_steps passes, entirely on the device (two small kernels
 * per step, no host round trip): forces from the current positions, residual_i = load_i + f_i for the
 * n_loads loaded nodes (distinct), history[step] = |residual of the FIRST load| (what the script
 * prints, truss.py:101-103), then displaced_i += residual_i / stiffness for loaded nodes only.
 * load_nodes int32 [n_loads], load_vecs [n_loads,3], displaced [n_nodes,3] in/out (start: a copy of
 * nodes, truss.py:95), residual [n_loads,3] scratch/out (the last step's residuals), history double
 * [n_steps] (may be NULL).  fp32 as above. */
int fea_truss_relax(const void* nodes, const int32_t* members, const void* k, int64_t n_members,
                    int64_t n_nodes, const int32_t* n2m_ptr, const int32_t* n2m,
                    const int32_t* load_nodes, const void* load_vecs, int64_t n_loads,
                    double stiffness, int32_t n_steps, void* displaced, void* residual,
                    double* history, int32_t fp32, void* stream);

/* moment_vector / shear_vector of euler_bernoulli.py:76-102 (the reference's own formulas). */
int fea_beam_moment_shear(const double* u, const double* EI, const double* length, int64_t n_elem,
                          double* moment, double* shear, void* stream);

/* stack_faces_2d on device (utils.py:356-376): nodes3d (n2d*n_layers,3), elements
 * (n_faces*(n_layers-1), 8) int32. */
int fea_mesh_extrude(const double* nodes2d, int64_t n2d, const int32_t* faces2d, int64_t n_faces,
                     const double* z_heights, int64_t n_layers, double* nodes3d, int32_t* elements,
                     void* stream);

/* generate_quad_grid(nx, ny, width, height), cubebeam.py:28-57, in device memory: nodes2d
 * [(nx+1)(ny+1), 2] x-fastest with np.linspace's values, quads int32 [nx*ny, 4] = [n1, n2, n4, n3]. */
int fea_mesh_quad_grid(int64_t nx, int64_t ny, double width, double height, double* nodes2d,
                       int32_t* quads, void* stream);

/* Tube cross-section of fea.py:28-48: nodes2d [2*n_seg, 2] (inner ring, then outer ring, angles
 * i * 2 pi / n_seg), periodic quads int32 [n_seg, 4] = [i, i+n, (i+1)%n + n, (i+1)%n]. */
int fea_mesh_tube_section(int64_t n_seg, double r_in, double r_out, double* nodes2d, int32_t* quads,
                          void* stream);

/* BASELINE config 5's frozen generator: jittered cubic lattice of n^3 nodes, members
 * along the 13 half-space neighbour directions (fea_mesh_lattice_members(n) of them, direction by
 * direction, start nodes in id order), per-member spring rates.  The random draws are numpy's PCG64
 * stream: pcg64_state_host[4] = {state hi, state lo, inc hi, inc lo} of np.random.PCG64(seed)
 * (default_rng(0) for the frozen case); nodes = grid*h + uniform(-0.1h, 0.1h) consumes draws
 * [0, 3n^3), k = uniform(500, 1500) the next M, each thread jumping ahead to its chunk (LCG
 * skip-ahead).  nodes [n^3,3], members int32 [M,2], k [M], dir_offset_dev int64 [14] scratch. */
int64_t fea_mesh_lattice_members(int64_t n);
int fea_mesh_lattice(int64_t n, double h, const uint64_t* pcg64_state_host, double* nodes,
                     int32_t* members, double* k, int64_t* dir_offset_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEA_B200_H */
