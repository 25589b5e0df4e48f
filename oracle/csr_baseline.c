/* CPU baseline for bench.py: what the reference executes on its hot path, restated in C.
 *
 * TEST/BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into the product.
 *
 * The reference assembles a sparse matrix (pattern: GDM::System::create_sparsity_pattern,
 * include/gdm/system.h:586-599 -- every pair of DoFs sharing a cell, (2p+1)^dim per interior
 * row; values: the cell loops of include/gdm/matrix_creator.h:21-61 and
 * tests/poisson_02_gdm.cc:160-206) and calls SparseMatrix::vmult inside deal.II's SolverCG
 * (tests/poisson_02_gdm.cc:213-215), one MPI rank per core with the rows split into slabs along
 * the last coordinate (system.h:720-757).  Here: CSR with int32 columns and FP64 values built from
 * the Kronecker structure (identical values, proven in tests/test_oracle_golden.py), OpenMP with
 * static contiguous row chunks (= the rank slabs), and the same CG loop.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline uses all host cores explicitly */
void gdm_oracle_set_num_threads(int n)
{
  if (n > 0)
    omp_set_num_threads(n);
}

int gdm_oracle_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Count / fill the CSR of  sum_d B_d (x) prod_{e != d} A_e  (has_b) or prod A_d on an
 * n[0] x n[1] x n[2] node grid.  Band tables: t[d][row*(2p+1)+tap], column = row+tap-p.
 * Structural pattern: taps where the (unconstrained) mass band is non-zero in every direction,
 * given as pat[d] with the same shape (non-zero = coupled).
 * Pass col == NULL to only count; returns nnz. */
int64_t gdm_oracle_kron_csr(const int *n, int p, int has_b, const double *const *A, const double *const *B,
                            const double *const *pat, int64_t *rowptr, int32_t *col, double *val)
{
  const int     W = 2 * p + 1;
  const int64_t n_rows = (int64_t)n[0] * n[1] * n[2];
  int64_t       nnz = 0;
  /* first pass: row lengths */
  for (int k = 0; k < n[2]; ++k)
    for (int j = 0; j < n[1]; ++j)
      for (int i = 0; i < n[0]; ++i)
        {
          int cx = 0, cy = 0, cz = 0;
          for (int t = 0; t < W; ++t)
            {
              const int ci = i + t - p, cj = j + t - p, ck = k + t - p;
              if (ci >= 0 && ci < n[0] && pat[0][(int64_t)i * W + t] != 0.0) ++cx;
              if (cj >= 0 && cj < n[1] && pat[1][(int64_t)j * W + t] != 0.0) ++cy;
              if (ck >= 0 && ck < n[2] && pat[2][(int64_t)k * W + t] != 0.0) ++cz;
            }
          if (n[1] == 1) cy = 1;
          if (n[2] == 1) cz = 1;
          const int64_t row = i + (int64_t)n[0] * (j + (int64_t)n[1] * k);
          rowptr[row] = nnz;
          nnz += (int64_t)cx * cy * cz;
        }
  rowptr[n_rows] = nnz;
  if (!col)
    return nnz;
#pragma omp parallel for schedule(static)
  for (int64_t row = 0; row < n_rows; ++row)
    {
      const int i = (int)(row % n[0]), j = (int)((row / n[0]) % n[1]), k = (int)(row / ((int64_t)n[0] * n[1]));
      int64_t   o = rowptr[row];
      for (int tz = 0; tz < W; ++tz)
        {
          const int ck = k + tz - p;
          if (n[2] == 1) { if (tz != p) continue; }
          else if (ck < 0 || ck >= n[2] || pat[2][(int64_t)k * W + tz] == 0.0) continue;
          for (int ty = 0; ty < W; ++ty)
            {
              const int cj = j + ty - p;
              if (n[1] == 1) { if (ty != p) continue; }
              else if (cj < 0 || cj >= n[1] || pat[1][(int64_t)j * W + ty] == 0.0) continue;
              for (int tx = 0; tx < W; ++tx)
                {
                  const int ci = i + tx - p;
                  if (ci < 0 || ci >= n[0] || pat[0][(int64_t)i * W + tx] == 0.0) continue;
                  const double ax = A[0][(int64_t)i * W + tx];
                  const double ay = n[1] == 1 ? 1.0 : A[1][(int64_t)j * W + ty];
                  const double az = n[2] == 1 ? 1.0 : A[2][(int64_t)k * W + tz];
                  double       v;
                  if (has_b)
                    {
                      const double bx = B[0][(int64_t)i * W + tx];
                      const double by = n[1] == 1 ? 0.0 : B[1][(int64_t)j * W + ty];
                      const double bz = n[2] == 1 ? 0.0 : B[2][(int64_t)k * W + tz];
                      v = bx * ay * az + ax * by * az + ax * ay * bz;
                    }
                  else
                    v = ax * ay * az;
                  col[o] = (int32_t)(ci + (int64_t)n[0] * (cj + (int64_t)n[1] * ck));
                  val[o] = v;
                  ++o;
                }
            }
        }
    }
  return nnz;
}

/* y = A x  (the reference's SparseMatrix::vmult) */
void gdm_oracle_spmv(int64_t n_rows, const int64_t *rowptr, const int32_t *col, const double *val, const double *x,
                     double *y)
{
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; ++r)
    {
      double acc = 0.0;
      for (int64_t o = rowptr[r]; o < rowptr[r + 1]; ++o)
        acc += val[o] * x[col[o]];
      y[r] = acc;
    }
}

static double dot(int64_t n, const double *a, const double *b)
{
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int64_t i = 0; i < n; ++i)
    s += a[i] * b[i];
  return s;
}

/* deal.II SolverCG + ReductionControl(max_steps, tol, reduce); dinv == NULL: PreconditionIdentity,
 * else Jacobi.  Returns the iteration count, or -1 - count on failure.  x must hold the start vector
 * (zeros in all reference call sites). */
int gdm_oracle_cg(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, double *x, const double *b,
                  const double *dinv, int max_steps, double tol, double reduce, double *last_value, double *work)
{
  double *g = work, *d = work + n, *h = work + 2 * n, *Ad = work + 3 * n;
  gdm_oracle_spmv(n, rowptr, col, val, x, g);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    g[i] -= b[i];
  double       res = sqrt(dot(n, g, g));
  const double reduced = res * reduce;
  *last_value = res;
  if (res < reduced || res <= tol)
    return 0;
  double gh = 0.0;
  for (int it = 1;; ++it)
    {
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i)
        h[i] = dinv ? dinv[i] * g[i] : g[i];
      if (it > 1)
        {
          const double old = gh;
          gh               = dot(n, g, h);
          const double beta = gh / old;
#pragma omp parallel for schedule(static)
          for (int64_t i = 0; i < n; ++i)
            d[i] = -h[i] + beta * d[i];
        }
      else
        {
#pragma omp parallel for schedule(static)
          for (int64_t i = 0; i < n; ++i)
            d[i] = -h[i];
          gh = dot(n, g, h);
        }
      gdm_oracle_spmv(n, rowptr, col, val, d, Ad);
      const double alpha = gh / dot(n, d, Ad);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i)
        {
          x[i] += alpha * d[i];
          g[i] += alpha * Ad[i];
        }
      res         = sqrt(dot(n, g, g));
      *last_value = res;
      if (res < reduced || res <= tol)
        return it;
      if (it >= max_steps || res != res)
        return -1 - it;
    }
}
