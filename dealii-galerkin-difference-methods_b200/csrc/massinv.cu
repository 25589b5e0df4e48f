// Kronecker-direct inverse of the GDM mass operator.
//
// On a Cartesian grid the constrained mass operator is  M = scale * (A_z (x) A_y (x) A_x)  on the free nodes plus
// deal.II's diagonal on the constrained rows (Dirichlet faces, duplicate nodes of periodic directions), so
//     M^-1 b  =  1/scale * (A_z^-1 (x) A_y^-1 (x) A_x^-1) b   on the free nodes,     b_i / M_ii  on constrained rows,
// with A_d the 1D mass matrix of direction d restricted to its free nodes: SPD, banded with half-bandwidth p, or
// ring-banded if the direction is periodic (node N folded into node 0, include/gdm/system.h:427-463).
// This replaces the CG / ILU / AMG mass solves the reference runs in every Runge-Kutta stage
// (applications/advection/include/gdm/advection/problem.h:236-267, applications/wave/include/gdm/wave/problem.h:471-502,
// prototypes/advection_01_gdm.cc:208-216) by three sweeps of banded line solves: 32 B/DoF per direction.
//
// Factorisation (host, once per operator): band Cholesky A = L L^T; a periodic direction is split into the leading
// banded block B (the first m - p free nodes) and a border of p nodes,  A = [B C; C^T D]:
//     y = B^-1 b1,   z = S^-1 (b2 - C^T y),   x1 = y - W z,   x2 = z,      S = D - C^T B^-1 C,   W = B^-1 C.
// Device: one thread per grid line marches along it (forward, backward, border correction); lanes run over the fastest
// other index, so the y and z sweeps are coalesced; the x sweep strides by the row pitch and relies on L1 for the
// sectors it revisits.  One rank only (a slab-partitioned direction needs a distributed banded solve).
#include <algorithm>
#include <cmath>
#include <vector>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    constexpr int MAXP = MAX_DEGREE;

    struct DirTables
    {
      int     f0 = 0, m = 0, m1 = 0, nb = 0; // first free node, free nodes, banded block, border
      double *d_L = nullptr;                 // [m1][p+1]: L(i, i-k), k = 0 holds 1 / L(i,i)
      double *d_W = nullptr;                 // [m1][nb]
      double *d_C = nullptr;                 // [2p][nb]: rows of C that are not zero (first p and last p rows of the block)
      double *d_Sinv = nullptr;              // [nb][nb]
    };

    struct MassInvPlan
    {
      DirTables dir[3];
      double   *d_dinv = nullptr; // 1 / diagonal of the operator (used on constrained rows)
      ~MassInvPlan()
      {
        for (auto &t : dir)
          {
            cudaFree(t.d_L);
            cudaFree(t.d_W);
            cudaFree(t.d_C);
            cudaFree(t.d_Sinv);
          }
        cudaFree(d_dinv);
      }
    };

    // dense lower Cholesky in place (row major n x n); returns false if the matrix is not positive definite
    bool cholesky(std::vector<double> &a, int n)
    {
      for (int j = 0; j < n; ++j)
        {
          double d = a[(size_t)j * n + j];
          for (int k = 0; k < j; ++k)
            d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
          if (!(d > 0.0))
            return false;
          d                     = std::sqrt(d);
          a[(size_t)j * n + j] = d;
          for (int i = j + 1; i < n; ++i)
            {
              double s = a[(size_t)i * n + j];
              for (int k = 0; k < j; ++k)
                s -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
              a[(size_t)i * n + j] = s / d;
            }
          for (int i = 0; i < j; ++i)
            a[(size_t)i * n + j] = 0.0;
        }
      return true;
    }
    void chol_solve(const std::vector<double> &l, int n, double *x) // x := (L L^T)^-1 x
    {
      for (int i = 0; i < n; ++i)
        {
          double s = x[i];
          for (int k = 0; k < i; ++k)
            s -= l[(size_t)i * n + k] * x[k];
          x[i] = s / l[(size_t)i * n + i];
        }
      for (int i = n - 1; i >= 0; --i)
        {
          double s = x[i];
          for (int k = i + 1; k < n; ++k)
            s -= l[(size_t)k * n + i] * x[k];
          x[i] = s / l[(size_t)i * n + i];
        }
    }

    template <class T>
    T *upload(const std::vector<T> &h)
    {
      T *d = nullptr;
      GDM_CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(1, h.size()) * sizeof(T)));
      if (!h.empty())
        GDM_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
      return d;
    }

    // 1D factorisation of direction d
    void factor_direction(const Operator &op, int d, DirTables &t)
    {
      const Layout &L = op.sys->L;
      const int     p = L.p, W = 2 * p + 1, N = L.N[d], n = N + 1;
      std::vector<double> band;
      band_matrix_1d(p, N, 0, band);
      // dense unconstrained matrix h * M1
      std::vector<double> A((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < W; ++k)
          {
            const int j = i + k - p;
            if (j >= 0 && j < n)
              A[(size_t)i * n + j] = L.h[d] * band[(size_t)i * W + k];
          }
      int f0 = 0, f1 = n; // free nodes [f0, f1)
      if (op.periodic[d])
        {
          // C^T A C: fold row and column N into 0
          for (int j = 0; j < n; ++j)
            A[j] += A[(size_t)N * n + j];
          for (int i = 0; i < n; ++i)
            A[(size_t)i * n] += A[(size_t)i * n + N];
          f1 = N;
        }
      if (op.dirichlet[d][0])
        f0 = 1;
      if (op.dirichlet[d][1])
        f1 = std::min(f1, N);
      const int m = f1 - f0;
      GDM_REQUIRE(m > 2 * p, GDM_ERR_NOT_IMPLEMENTED, "mass inverse: too few free nodes in a direction");
      const int nb = op.periodic[d] ? p : 0, m1 = m - nb;
      t.f0 = f0, t.m = m, t.m1 = m1, t.nb = nb;
      // leading block
      std::vector<double> B((size_t)m1 * m1);
      for (int i = 0; i < m1; ++i)
        for (int j = 0; j < m1; ++j)
          B[(size_t)i * m1 + j] = A[(size_t)(f0 + i) * n + (f0 + j)];
      GDM_REQUIRE(cholesky(B, m1), GDM_ERR_INTERNAL, "mass inverse: 1D mass matrix is not positive definite");
      std::vector<double> hL((size_t)m1 * (p + 1), 0.0);
      for (int i = 0; i < m1; ++i)
        {
          hL[(size_t)i * (p + 1)] = 1.0 / B[(size_t)i * m1 + i];
          for (int k = 1; k <= p && k <= i; ++k)
            hL[(size_t)i * (p + 1) + k] = B[(size_t)i * m1 + (i - k)];
          // (the factor of a band matrix has the same band: entries further left are zero up to rounding)
        }
      t.d_L = upload(hL);
      if (nb > 0)
        {
          std::vector<double> C((size_t)m1 * nb), Wm((size_t)m1 * nb), S((size_t)nb * nb), col(m1);
          for (int i = 0; i < m1; ++i)
            for (int j = 0; j < nb; ++j)
              C[(size_t)i * nb + j] = A[(size_t)(f0 + i) * n + (f0 + m1 + j)];
          for (int j = 0; j < nb; ++j)
            {
              for (int i = 0; i < m1; ++i)
                col[i] = C[(size_t)i * nb + j];
              chol_solve(B, m1, col.data());
              for (int i = 0; i < m1; ++i)
                Wm[(size_t)i * nb + j] = col[i];
            }
          for (int a = 0; a < nb; ++a)
            for (int b = 0; b < nb; ++b)
              {
                double s = A[(size_t)(f0 + m1 + a) * n + (f0 + m1 + b)];
                for (int i = 0; i < m1; ++i)
                  s -= C[(size_t)i * nb + a] * Wm[(size_t)i * nb + b];
                S[(size_t)a * nb + b] = s;
              }
          // S^-1 by Cholesky solves of the unit vectors
          std::vector<double> Sl = S, Sinv((size_t)nb * nb);
          GDM_REQUIRE(cholesky(Sl, nb), GDM_ERR_INTERNAL, "mass inverse: Schur complement is not positive definite");
          for (int j = 0; j < nb; ++j)
            {
              std::vector<double> e(nb, 0.0);
              e[j] = 1.0;
              chol_solve(Sl, nb, e.data());
              for (int i = 0; i < nb; ++i)
                Sinv[(size_t)i * nb + j] = e[i];
            }
          // rows of C that can be nonzero: the first p and the last p rows of the block (all rows of a block shorter
          // than 2p)
          const int           nce = std::min(2 * p, m1);
          std::vector<double> Ce((size_t)2 * p * nb, 0.0);
          for (int r = 0; r < nce; ++r)
            {
              const int i = (r < p || m1 < 2 * p) ? r : m1 - 2 * p + r;
              for (int j = 0; j < nb; ++j)
                Ce[(size_t)r * nb + j] = C[(size_t)i * nb + j];
            }
          for (int i = p; i < m1 - p; ++i)
            for (int j = 0; j < nb; ++j)
              GDM_REQUIRE(C[(size_t)i * nb + j] == 0.0, GDM_ERR_INTERNAL, "mass inverse: unexpected coupling to the border");
          t.d_W    = upload(Wm);
          t.d_C    = upload(Ce);
          t.d_Sinv = upload(Sinv);
        }
    }

    struct LineK
    {
      double       *v;
      const double *Lb, *Wt, *Ce, *Sinv;
      int           m1, nb;
      int64_t       step;        // element stride of one node step along the line
      int64_t       off0;        // offset of the first free node of a line relative to the line origin
      // the other two (line) indices.  Index 0 carries the components: i0 = node0 * epn0 + component, offset =
      // node0 * sA + component; index 1: offset = i1 * s1.  Free ranges [lo, hi) in nodes.
      int64_t       n0, n1, sA, s1;
      int           lo0, hi0, lo1, hi1, epn0;
    };

    // one thread per line: forward and backward substitution with the band factor (rows of the factor are read by all
    // lanes at once: broadcast loads), then the border correction of a periodic direction
    template <int P>
    __global__ void line_solve_kernel(const LineK a)
    {
      const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (line >= a.n0 * a.n1)
        return;
      const int64_t i0 = line % a.n0, i1 = line / a.n0;
      // lines through constrained nodes of the other directions hold constrained rows only: not part of the product
      const int node0 = (int)(i0 / a.epn0), node1 = (int)i1;
      if (node0 < a.lo0 || node0 >= a.hi0 || node1 < a.lo1 || node1 >= a.hi1)
        return;
      double *v = a.v + (int64_t)node0 * a.sA + (i0 % a.epn0) + i1 * a.s1 + a.off0;
      double  w[P];
#pragma unroll
      for (int k = 0; k < P; ++k)
        w[k] = 0.0;
      for (int i = 0; i < a.m1; ++i)
        {
          const double *l = a.Lb + (size_t)i * (P + 1);
          double        s = v[(int64_t)i * a.step];
#pragma unroll
          for (int k = 1; k <= P; ++k)
            s = fma(-__ldg(l + k), w[k - 1], s);
          s *= __ldg(l);
#pragma unroll
          for (int k = P - 1; k > 0; --k)
            w[k] = w[k - 1];
          w[0]                     = s;
          v[(int64_t)i * a.step] = s;
        }
#pragma unroll
      for (int k = 0; k < P; ++k)
        w[k] = 0.0;
      for (int i = a.m1 - 1; i >= 0; --i)
        {
          double s = v[(int64_t)i * a.step];
#pragma unroll
          for (int k = 1; k <= P; ++k)
            if (i + k < a.m1)
              s = fma(-__ldg(a.Lb + (size_t)(i + k) * (P + 1) + k), w[k - 1], s);
          s *= __ldg(a.Lb + (size_t)i * (P + 1));
#pragma unroll
          for (int k = P - 1; k > 0; --k)
            w[k] = w[k - 1];
          w[0]                     = s;
          v[(int64_t)i * a.step] = s;
        }
      if (a.nb > 0) // (nb == P)
        {
          double t[P], z[P];
#pragma unroll
          for (int j = 0; j < P; ++j)
            t[j] = v[(int64_t)(a.m1 + j) * a.step];
          const int nce = min(2 * P, a.m1);
          for (int r = 0; r < nce; ++r)
            {
              const int    i = (r < P || a.m1 < 2 * P) ? r : a.m1 - 2 * P + r;
              const double y = v[(int64_t)i * a.step];
#pragma unroll
              for (int j = 0; j < P; ++j)
                t[j] = fma(-__ldg(a.Ce + r * P + j), y, t[j]);
            }
#pragma unroll
          for (int i = 0; i < P; ++i)
            {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < P; ++j)
                s = fma(__ldg(a.Sinv + i * P + j), t[j], s);
              z[i] = s;
            }
          for (int i = 0; i < a.m1; ++i)
            {
              double s = v[(int64_t)i * a.step];
#pragma unroll
              for (int j = 0; j < P; ++j)
                s = fma(-__ldg(a.Wt + (size_t)i * P + j), z[j], s);
              v[(int64_t)i * a.step] = s;
            }
#pragma unroll
          for (int j = 0; j < P; ++j)
            v[(int64_t)(a.m1 + j) * a.step] = z[j];
        }
    }

    struct PrepK
    {
      double       *dst;
      const double *src, *dinv;
      int           ln[3], lo[3], hi[3], nc;
      int64_t       stride[3];
      int64_t       n;
      double        inv_scale;
    };
    // dst = src / scale on the free nodes (input of the line solves), src / M_ii on the constrained rows
    __global__ void massinv_prepare_kernel(const PrepK a)
    {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (tid >= a.n)
        return;
      const int64_t row_len = (int64_t)a.ln[0] * a.nc;
      const int64_t e0 = tid % row_len, r = tid / row_len;
      const int     x = (int)(e0 / a.nc), y = (int)(r % a.ln[1]), z = (int)(r / a.ln[1]);
      const bool    free = x >= a.lo[0] && x < a.hi[0] && y >= a.lo[1] && y < a.hi[1] && z >= a.lo[2] && z < a.hi[2];
      const int64_t off  = e0 + y * a.stride[1] + z * a.stride[2];
      const double  s    = a.src[off];
      double        dv   = a.dinv[off];
      if (!isfinite(dv))
        dv = 0.0; // constrained rows without a diagonal (GDM_DIAG_ZERO): the solution stays zero there
      a.dst[off] = free ? s * a.inv_scale : s * dv;
    }

    template <int P>
    void launch_line(Context &ctx, const LineK &a)
    {
      const int64_t lines = a.n0 * a.n1;
      const int     th    = 128;
      line_solve_kernel<P><<<(unsigned)((lines + th - 1) / th), th, 0, ctx.stream>>>(a);
      ctx.launches++;
      GDM_CUDA_CHECK(cudaGetLastError());
    }
  } // namespace

  bool massinv_supported(const Operator &op)
  {
    const Layout &L = op.sys->L;
    return op.desc.kind == GDM_OP_MASS && L.n_ranks == 1 && !op.csr && L.p <= MAXP;
  }

  void massinv_destroy(Operator &op)
  {
    delete static_cast<MassInvPlan *>(op.massinv);
    op.massinv = nullptr;
  }

  void massinv_apply(Operator &op, double *dst, const double *src)
  {
    Context      &ctx = *op.sys->ctx;
    const Layout &L   = op.sys->L;
    GDM_REQUIRE(massinv_supported(op), GDM_ERR_NOT_IMPLEMENTED,
                "the direct mass inverse needs a MASS operator without irregular rows on one rank");
    if (!op.massinv)
      {
        std::unique_ptr<MassInvPlan> plan(new MassInvPlan);
        for (int d = 0; d < L.dim; ++d)
          factor_direction(op, d, plan->dir[d]);
        // 1 / diagonal of the operator (as for the Jacobi preconditioner, cg.cu)
        GDM_CUDA_CHECK(cudaMalloc(&plan->d_dinv, (size_t)L.size * sizeof(double)));
        GDM_CUDA_CHECK(cudaMemsetAsync(plan->d_dinv, 0, (size_t)L.size * sizeof(double), ctx.stream));
        launch_diagonal(ctx, L, op, plan->d_dinv);
        if (op.desc.constrained_diagonal == GDM_DIAG_ASSEMBLED)
          {
            ctx.ensure_scratch((size_t)L.size);
            blas_set(ctx, ctx.scratch[0], L.size, 1.0);
            launch_constrained_rows(ctx, L, op, plan->d_dinv, ctx.scratch[0], true);
          }
        blas_invert(ctx, plan->d_dinv + L.own_off, L.own_len);
        op.massinv = plan.release();
      }
    MassInvPlan &plan = *static_cast<MassInvPlan *>(op.massinv);
    // free ranges per direction (nodes); unused directions: one free node
    int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    for (int d = 0; d < L.dim; ++d)
      {
        lo[d] = plan.dir[d].f0;
        hi[d] = plan.dir[d].f0 + plan.dir[d].m;
      }
    {
      PrepK a;
      a.dst  = dst;
      a.src  = src;
      a.dinv = plan.d_dinv;
      a.nc   = L.nc;
      for (int d = 0; d < 3; ++d)
        {
          a.ln[d]     = L.ln[d];
          a.lo[d]     = lo[d];
          a.hi[d]     = hi[d];
          a.stride[d] = L.stride[d];
        }
      a.n         = (int64_t)L.ln[0] * L.nc * L.ln[1] * L.ln[2];
      a.inv_scale = 1.0 / op.desc.scale;
      const int th = 256;
      massinv_prepare_kernel<<<(unsigned)((a.n + th - 1) / th), th, 0, ctx.stream>>>(a);
      ctx.launches++;
      GDM_CUDA_CHECK(cudaGetLastError());
    }
    for (int d = 0; d < L.dim; ++d)
      {
        const DirTables &t = plan.dir[d];
        LineK            a;
        a.v    = dst;
        a.Lb   = t.d_L;
        a.Wt   = t.d_W;
        a.Ce   = t.d_C;
        a.Sinv = t.d_Sinv;
        a.m1   = t.m1;
        a.nb   = t.nb;
        a.step = L.stride[d];
        a.off0 = (int64_t)t.f0 * L.stride[d];
        // the two other directions; the components ride on the lowest one (x for the y / z sweeps, y for the x sweep)
        int e[2], k = 0;
        for (int q = 0; q < 3; ++q)
          if (q != d)
            e[k++] = q;
        a.epn0 = L.nc;
        a.n0   = (int64_t)L.ln[e[0]] * L.nc;
        a.sA   = L.stride[e[0]];
        a.lo0 = lo[e[0]], a.hi0 = hi[e[0]];
        a.n1 = L.ln[e[1]], a.s1 = L.stride[e[1]], a.lo1 = lo[e[1]], a.hi1 = hi[e[1]];
        switch (L.p)
          {
            case 1:
              launch_line<1>(ctx, a);
              break;
            case 3:
              launch_line<3>(ctx, a);
              break;
            case 5:
              launch_line<5>(ctx, a);
              break;
            case 7:
              launch_line<7>(ctx, a);
              break;
            case 9:
              launch_line<9>(ctx, a);
              break;
            default:
              throw Error(GDM_ERR_NOT_IMPLEMENTED, "direct mass inverse: fe_degree");
          }
      }
  }
} // namespace gdm
