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

inline unsigned vec_blocks(int64_t n) {
  // 8 resident CTAs of 256 threads per SM, one full wave (<= kMaxPartials blocks)
  const int64_t b = (n + 256 * 4 - 1) / (256 * 4);
  return (unsigned)(b < 1 ? 1 : (b > 148LL * 8 ? 148LL * 8 : b));
}

// Solver kernels defined in pcg.cu, launched by both drivers.
__global__ void __launch_bounds__(256) pcg_update_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ p,
                                  const double* __restrict__ ap, double* __restrict__ x, double* __restrict__ r,
                                  PcgState* st, double* partials);
__global__ void __launch_bounds__(256) pcg_direction_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                     double* __restrict__ p, PcgState* st, double* history);
__global__ void __launch_bounds__(256) pcg_init_kernel(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
                                double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, double tol,
                                int max_iter, PcgState* st, double* partials);

// step 1 (ap = K p over `n_nodes` rows, p.ap into state) on the best available kernel.
int pcg_step_spmv(int d, int64_t n_nodes, const int32_t* rp, const int32_t* ci, const double* values,
                  const double* p, double* ap, int64_t p_row_offset, PcgState* st, double* partials,
                  cudaStream_t stream, const TmaPlan* plan);
void pcg_match_carveout();

}  // namespace fea
