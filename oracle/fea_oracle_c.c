/* CPU oracle, C restatement of jjrreett/fea's hot path, multi-threaded with OpenMP.
 *
 * TEST INFRASTRUCTURE ONLY -- this is the checker and the CPU baseline, never the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it (through oracle/c_oracle.py).  Nothing under fea_b200/ links or loads it.
 *
 * Why it exists next to oracle/fea_oracle.py (numpy/scipy): the reference is single-threaded
 * Python; a fair CPU baseline on a many-core host needs the same algorithm on every core.  Each
 * function follows the cited reference lines in the reference's own (dense) order of
 * operations; it is validated against the numpy oracle and, through it, against the
 * reference-generated golden fixtures (tests/test_oracle_golden.py::test_c_oracle_*).
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle/fea_oracle_c.c -o oracle/libfea_oracle_c.so -lm
 * (done by oracle/c_oracle.py on first use and by __graft_entry__.build()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fea_c_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Pin the OpenMP team size (bench.py: torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm
 * uses the host's cores whatever the launcher set). */
void fea_c_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Local node signs, bottom face CCW then top face CCW (utils.py:159-197, 351-353). */
static const double SX[8] = {-1, +1, +1, -1, -1, +1, +1, -1};
static const double SY[8] = {-1, -1, +1, +1, -1, -1, +1, +1};
static const double SZ[8] = {-1, -1, -1, -1, +1, +1, +1, +1};

/* Ke (24x24, row-major) of one hex8, utils.py:127-239 in the reference's order:
 * Gauss points xi outer / eta / zeta inner (utils.py:200-204), dN (utils.py:159-197), J = dN X,
 * det, inverse (utils.py:210-218), dN_dx = J^-1 dN (utils.py:221), B (utils.py:224-234),
 * Ke += w (B^T C B) detJ (utils.py:237).  Returns 1 if some detJ <= 0 (utils.py:212-215). */
int fea_c_hex8_ke(const double* x /* 8x3 */, double E, double nu, double* ke /* 576 */) {
  double C[6][6];
  memset(C, 0, sizeof(C));
  const double c = E / ((1.0 + nu) * (1.0 - 2.0 * nu)); /* utils.py:144 */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i][j] = c * (i == j ? 1.0 - nu : nu);
  for (int i = 3; i < 6; ++i) C[i][i] = c * ((1.0 - 2.0 * nu) / 2.0);
  memset(ke, 0, 576 * sizeof(double));
  const double g[2] = {-1.0 / sqrt(3.0), 1.0 / sqrt(3.0)}; /* utils.py:140 */
  int bad = 0;
  for (int ix = 0; ix < 2; ++ix)
    for (int iy = 0; iy < 2; ++iy)
      for (int iz = 0; iz < 2; ++iz) {
        const double xi = g[ix], eta = g[iy], zeta = g[iz];
        double dN[3][8];
        for (int a = 0; a < 8; ++a) {
          const double fx = 1.0 + SX[a] * xi, fy = 1.0 + SY[a] * eta, fz = 1.0 + SZ[a] * zeta;
          dN[0][a] = SX[a] * fy * fz / 8.0;
          dN[1][a] = SY[a] * fx * fz / 8.0;
          dN[2][a] = SZ[a] * fx * fy / 8.0;
        }
        double J[3][3];
        for (int r = 0; r < 3; ++r)
          for (int cc = 0; cc < 3; ++cc) {
            double s = 0.0;
            for (int a = 0; a < 8; ++a) s += dN[r][a] * x[3 * a + cc];
            J[r][cc] = s;
          }
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        if (!(det > 0.0)) bad = 1;
        double I[3][3];
        I[0][0] = c00 / det;
        I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        I[1][0] = c01 / det;
        I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
        I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        I[2][0] = c02 / det;
        I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        double B[6][24];
        memset(B, 0, sizeof(B));
        for (int a = 0; a < 8; ++a) {
          double d[3];
          for (int r = 0; r < 3; ++r) d[r] = I[r][0] * dN[0][a] + I[r][1] * dN[1][a] + I[r][2] * dN[2][a];
          B[0][3 * a + 0] = d[0]; /* utils.py:224-234 */
          B[1][3 * a + 1] = d[1];
          B[2][3 * a + 2] = d[2];
          B[3][3 * a + 0] = d[1];
          B[3][3 * a + 1] = d[0];
          B[4][3 * a + 1] = d[2];
          B[4][3 * a + 2] = d[1];
          B[5][3 * a + 0] = d[2];
          B[5][3 * a + 2] = d[0];
        }
        double CB[6][24];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 24; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += C[i][k] * B[k][j];
            CB[i][j] = s;
          }
        for (int i = 0; i < 24; ++i)
          for (int j = 0; j < 24; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += B[k][i] * CB[k][j];
            ke[24 * i + j] += 1.0 * s * det; /* utils.py:237, weight 1 */
          }
      }
  return bad;
}

/* Batched Ke: elements (M,8) int64 node ids, nodes (N,3).  Returns the number of bad elements. */
int64_t fea_c_hex8_ke_batch(const double* nodes, const int64_t* elements, int64_t M, double E, double nu,
                            double* ke_out /* M*576 */) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t e = 0; e < M; ++e) {
    double x[24];
    for (int a = 0; a < 8; ++a)
      for (int cc = 0; cc < 3; ++cc) x[3 * a + cc] = nodes[3 * elements[8 * e + a] + cc];
    bad += fea_c_hex8_ke(x, E, nu, ke_out + 576 * e);
  }
  return bad;
}

static int64_t find_col(const int32_t* idx, int64_t lo, int64_t hi, int32_t key) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (idx[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* Ke evaluation + `K[np.ix_(d, d)] += Ke` (cubebeam.py:82-90) into an existing CSR pattern
 * (indptr, sorted indices; data zeroed by the caller).  DOF map 3*node + c (cubebeam.py:86).
 * Elements are processed in parallel, the adds are atomic: the summation order differs from the
 * reference's sequential loop by rounding only. */
int64_t fea_c_assemble_hex8(const double* nodes, const int64_t* elements, int64_t M, double E, double nu,
                            const int32_t* indptr, const int32_t* indices, double* data) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t e = 0; e < M; ++e) {
    double x[24], ke[576];
    int32_t dof[24];
    for (int a = 0; a < 8; ++a)
      for (int cc = 0; cc < 3; ++cc) {
        x[3 * a + cc] = nodes[3 * elements[8 * e + a] + cc];
        dof[3 * a + cc] = (int32_t)(3 * elements[8 * e + a] + cc);
      }
    bad += fea_c_hex8_ke(x, E, nu, ke);
    for (int i = 0; i < 24; ++i) {
      const int64_t lo = indptr[dof[i]], hi = indptr[dof[i] + 1];
      for (int j = 0; j < 24; ++j) {
        const int64_t p = find_col(indices, lo, hi, dof[j]);
#pragma omp atomic
        data[p] += ke[24 * i + j];
      }
    }
  }
  return bad;
}

/* y = A x, CSR (cubebeam.py:106 `K @ u`). */
void fea_c_spmv(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, const double* x,
                double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) s += data[p] * x[indices[p]];
    y[i] = s;
  }
}

/* Jacobi-preconditioned CG on the REDUCED system (cubebeam.py:92-98 with np.linalg.solve replaced
 * by the iterative solver BASELINE.json names): x0 = 0, stop when ||r|| <= tol ||b|| (recurrence
 * residual) or after maxiter iterations.  Same recurrence as oracle/fea_oracle.py:jacobi_pcg.
 * Returns the iteration count; *relres = ||r|| / ||b||. */
int64_t fea_c_jacobi_pcg(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                         const double* b, double* x, double tol, int64_t maxiter, double* relres) {
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  double* z = (double*)malloc(sizeof(double) * (size_t)n);
  double* p = (double*)malloc(sizeof(double) * (size_t)n);
  double* ap = (double*)malloc(sizeof(double) * (size_t)n);
  double* dinv = (double*)malloc(sizeof(double) * (size_t)n);
  double bb = 0.0, rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bb, rz)
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    for (int64_t q = indptr[i]; q < indptr[i + 1]; ++q)
      if (indices[q] == i) d = data[q];
    dinv[i] = 1.0 / d;
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i];
    rz += r[i] * z[i];
  }
  double rr = bb;
  int64_t it = 0;
  while (it < maxiter && bb > 0.0 && !(rr <= tol * tol * bb)) {
    fea_c_spmv(n, indptr, indices, data, p, ap);
    double pap = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pap)
    for (int64_t i = 0; i < n; ++i) pap += p[i] * ap[i];
    const double alpha = rz / pap;
    double rz_new = 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * ap[i];
      z[i] = dinv[i] * r[i];
      rz_new += r[i] * z[i];
      rr += r[i] * r[i];
    }
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(r);
  free(z);
  free(p);
  free(ap);
  free(dinv);
  return it;
}

/* Constraint reduction `K[np.ix_(free, free)]` (cubebeam.py:92-96) on CSR, two passes.
 * map[i] = rank of DOF i within `free` (ascending), or -1 for a constrained DOF.
 * Pass 1 (out_indices == NULL): out_indptr[r + 1] = length of reduced row r; the caller turns the
 * counts into offsets.  Pass 2: fills indices / data at those offsets. */
void fea_c_reduce_csr(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                      const int32_t* map, int64_t* out_indptr, int32_t* out_indices, double* out_data) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int32_t r = map[i];
    if (r < 0) continue;
    if (out_indices == NULL) {
      int64_t c = 0;
      for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) c += map[indices[p]] >= 0;
      out_indptr[r + 1] = c;
    } else {
      int64_t q = out_indptr[r];
      for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
        const int32_t c = map[indices[p]];
        if (c >= 0) {
          out_indices[q] = c;
          out_data[q] = data[p];
          ++q;
        }
      }
    }
  }
}

/* Expansion of a node-level CSR pattern (indptr_n, sorted indices_n) to d DOF per node: row d*i+a
 * holds columns d*j+b, j over the node's list, b < d (DOF map d*node + component, cubebeam.py:86).
 * out_indptr has d*n_nodes+1 entries, out_indices d*d*nnz_nodes. */
void fea_c_expand_pattern(int64_t n_nodes, int32_t d, const int32_t* indptr_n, const int32_t* indices_n,
                          int32_t* out_indptr, int32_t* out_indices) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_nodes; ++i) {
    const int64_t lo = indptr_n[i], cnt = indptr_n[i + 1] - lo;
    for (int32_t a = 0; a < d; ++a) {
      const int64_t base = (int64_t)d * d * lo + (int64_t)a * d * cnt;
      out_indptr[d * i + a] = (int32_t)base;
      for (int64_t k = 0; k < cnt; ++k)
        for (int32_t b = 0; b < d; ++b) out_indices[base + d * k + b] = d * indices_n[lo + k] + b;
    }
  }
  out_indptr[d * n_nodes] = (int32_t)((int64_t)d * d * indptr_n[n_nodes]);
}
