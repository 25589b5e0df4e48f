// GDM 1D basis and 1D band matrices (host side, setup only).
//
// Stands in for GDM::generate_polynomials_1D (reference include/gdm/fe.h:55-336): instead of
// storing the tables, the Lagrange basis through the integer nodes  k - v  (k = 0..p, variant v,
// cell = [0,1]; generator spec scripts/create_coefficients.py:15-39) is evaluated in closed form.
// The assembled 1D matrices follow the reference's cell loop restricted to one direction
// (include/gdm/matrix_creator.h:21-61, tests/poisson_02_gdm.cc:160-206) with QGauss(p+1), which
// is exact for these integrands; the window/variant rule is system.h:209-216 / 415-420.
#include <algorithm>
#include <cmath>

#include "gdm_internal.h"

namespace gdm
{
  void gauss_legendre_01(int n, std::vector<long double> &x, std::vector<long double> &w)
  {
    x.assign(n, 0);
    w.assign(n, 0);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < (n + 1) / 2; ++i)
      {
        long double z = std::cos(pi * (i + 0.75L) / (n + 0.5L));
        long double pp = 0;
        for (int it = 0; it < 100; ++it)
          {
            long double p1 = 1, p2 = 0;
            for (int j = 0; j < n; ++j)
              {
                const long double p3 = p2;
                p2                   = p1;
                p1                   = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
              }
            pp                  = n * (z * p1 - p2) / (z * z - 1);
            const long double dz = p1 / pp;
            z -= dz;
            if (std::fabs((double)dz) < 1e-19)
              break;
          }
        // recompute derivative at the converged root
        {
          long double p1 = 1, p2 = 0;
          for (int j = 0; j < n; ++j)
            {
              const long double p3 = p2;
              p2                   = p1;
              p1                   = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
            }
          pp = n * (z * p1 - p2) / (z * z - 1);
        }
        const long double wi = 2 / ((1 - z * z) * pp * pp);
        // map [-1,1] -> [0,1], ascending
        x[i]         = 0.5L * (1 - z);
        x[n - 1 - i] = 0.5L * (1 + z);
        w[i] = w[n - 1 - i] = 0.5L * wi;
      }
  }

  void lagrange_eval(int p, int v, long double x, long double *values, long double *derivs)
  {
    for (int k = 0; k <= p; ++k)
      {
        const long double xk = k - v;
        long double       denom = 1;
        for (int j = 0; j <= p; ++j)
          if (j != k)
            denom *= (xk - (j - v));
        long double val = 1;
        for (int j = 0; j <= p; ++j)
          if (j != k)
            val *= (x - (j - v));
        long double der = 0;
        for (int m = 0; m <= p; ++m)
          {
            if (m == k)
              continue;
            long double t = 1;
            for (int j = 0; j <= p; ++j)
              if (j != k && j != m)
                t *= (x - (j - v));
            der += t;
          }
        values[k] = val / denom;
        if (derivs)
          derivs[k] = der / denom;
      }
  }

  void lagrange_monomials(int p, int v, double *coeffs)
  {
    // numerator polynomial prod_{j != k} (x - xi_j) has integer coefficients
    for (int k = 0; k <= p; ++k)
      {
        std::vector<long long> poly(1, 1); // lowest power first
        long long              denom = 1;
        for (int j = 0; j <= p; ++j)
          {
            if (j == k)
              continue;
            const long long xj = j - v;
            denom *= ((k - v) - xj);
            std::vector<long long> next(poly.size() + 1, 0);
            for (size_t i = 0; i < poly.size(); ++i)
              {
                next[i + 1] += poly[i];
                next[i] -= xj * poly[i];
              }
            poly.swap(next);
          }
        for (int i = 0; i <= p; ++i)
          coeffs[k * (p + 1) + i] = (double)((long double)poly[i] / (long double)denom);
      }
  }

  CellMatrices1D cell_matrices_1d(int p, int v)
  {
    const int                n = p + 1;
    std::vector<long double> xq, wq;
    gauss_legendre_01(n, xq, wq);
    std::vector<long double> M(n * n, 0), K(n * n, 0), C(n * n, 0), f(n, 0);
    std::vector<long double> val(n), der(n);
    for (int q = 0; q < n; ++q)
      {
        lagrange_eval(p, v, xq[q], val.data(), der.data());
        for (int a = 0; a < n; ++a)
          {
            f[a] += wq[q] * val[a];
            for (int b = 0; b < n; ++b)
              {
                M[a * n + b] += wq[q] * val[a] * val[b];
                K[a * n + b] += wq[q] * der[a] * der[b];
                C[a * n + b] += wq[q] * val[a] * der[b];
              }
          }
      }
    CellMatrices1D out;
    out.M.assign(M.begin(), M.end());
    out.K.assign(K.begin(), K.end());
    out.C.assign(C.begin(), C.end());
    out.f.assign(f.begin(), f.end());
    return out;
  }

  namespace
  {
    const CellMatrices1D &cached_cell(int p, int v)
    {
      static std::vector<std::vector<CellMatrices1D>> cache(MAX_DEGREE + 1);
      if (cache[p].empty())
        for (int i = 0; i < p; ++i)
          cache[p].push_back(cell_matrices_1d(p, i));
      return cache[p][v];
    }
  } // namespace

  void band_matrix_1d(int p, int N, int kind, std::vector<double> &band)
  {
    const int W = 2 * p + 1;
    band.assign((size_t)(N + 1) * W, 0.0);
    for (int c = 0; c < N; ++c)
      {
        const int                  off = window_offset(p, N, c);
        const int                  v   = cell_variant(p, N, c);
        const CellMatrices1D      &cm  = cached_cell(p, v);
        const std::vector<double> &m   = (kind == 0) ? cm.M : (kind == 1) ? cm.K : cm.C;
        for (int a = 0; a <= p; ++a)
          for (int b = 0; b <= p; ++b)
            {
              const int row = off + a, col = off + b;
              band[(size_t)row * W + (col - row + p)] += m[a * (p + 1) + b];
            }
      }
  }

  void load_vector_1d(int p, int N, std::vector<double> &f)
  {
    f.assign(N + 1, 0.0);
    for (int c = 0; c < N; ++c)
      {
        const int             off = window_offset(p, N, c);
        const CellMatrices1D &cm  = cached_cell(p, cell_variant(p, N, c));
        for (int a = 0; a <= p; ++a)
          f[off + a] += cm.f[a];
      }
  }
} // namespace gdm
