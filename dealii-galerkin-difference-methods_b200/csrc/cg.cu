// Conjugate gradients with deal.II semantics, device resident.
//
// Stands in for dealii::SolverCG<VectorType>::solve(A, x, b, P) + ReductionControl as used at every
// solver.solve call site of the reference (tests/poisson_01_gdm.cc:164-170, tests/mass_01_gdm.cc:124-131,
// prototypes/advection_01_gdm.cc:208-216, applications/wave/include/gdm/wave/problem.h:471-502):
//   g = A x - b ; check(0, |g|) ; repeat { h = P g ; d = -h + beta d ; alpha = (g.h)/(d.Ad) ;
//   x += alpha d ; g += alpha Ad ; check(it, |g|) }   (here r = -g, p = d).
// The host never sees alpha/beta/|r|: kernels pass raw sums through ctx.d_sums, the convergence
// decision is taken on the device and freezes the state; the host only polls a status record
// every few iterations, so the stream never drains inside a batch.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    struct Work
    {
      Context &ctx;
      double  *r = nullptr, *p = nullptr, *q = nullptr, *dinv = nullptr;
      explicit Work(Context &c)
        : ctx(c)
      {}
      ~Work()
      {
        ctx.release(r);
        ctx.release(p);
        ctx.release(q);
        ctx.release(dinv);
      }
    };

    void apply(Operator &A, double *dst, const double *src)
    {
      Context &ctx = *A.sys->ctx;
      if (A.kernel_used == GDM_KERNEL_FUSED)
        fused_apply(A, dst, src, false, true); // ghost import overlapped with the interior planes
      else
        {
          if (A.sys->L.n_ranks > 1)
            comm_halo_exchange(ctx, A.sys->L, const_cast<double *>(src));
          generic_apply(A, dst, src, false);
        }
      if (A.csr)
        launch_csr_overlay(ctx, *A.csr, dst, src, false);
    }
  } // namespace

  int cg_solve(Operator &A, Vector &x, Vector &b, int precondition, Vector *pvec, gdm_reduction_control &ctl)
  {
    System       &sys = *A.sys;
    Context      &ctx = *sys.ctx;
    const Layout &L   = sys.L;
    const int64_t n     = L.own_len;
    const int64_t off   = L.own_off;

    Work w(ctx);
    w.r = ctx.acquire((size_t)L.size);
    w.p = ctx.acquire((size_t)L.size);
    w.q = ctx.acquire((size_t)L.size);
    const double *dinv = nullptr;
    if (precondition == GDM_PRECONDITION_JACOBI)
      {
        w.dinv = ctx.acquire((size_t)L.size);
        launch_diagonal(ctx, L, A, w.dinv);
        if (A.desc.constrained_diagonal == GDM_DIAG_ASSEMBLED)
          {
            ctx.ensure_scratch((size_t)L.size);
            blas_set(ctx, ctx.scratch[0], L.size, 1.0);
            launch_constrained_rows(ctx, L, A, w.dinv, ctx.scratch[0], true);
          }
        if (A.csr)
          launch_csr_diagonal(ctx, *A.csr, w.dinv);
        blas_invert(ctx, w.dinv + off, n);
        dinv = w.dinv + off;
      }
    else if (precondition == GDM_PRECONDITION_DIAGONAL)
      dinv = pvec->d + off;

    struct StatusGuard
    {
      void *p;
      ~StatusGuard()
      {
        cg_status_free(p);
      }
    } status{cg_status_alloc(ctx)};

    // r = b - A x  (deal.II skips the product when x == 0; same result)
    {
      blas_dot(ctx, x.d + off, x.d + off, n, SUM_TMP);
      const double xx = read_sum(ctx, SUM_TMP, true);
      blas_copy(ctx, w.r + off, b.d + off, n);
      if (xx != 0.0)
        {
          apply(A, w.q, x.d);
          blas_sadd(ctx, w.r + off, 1.0, -1.0, w.q + off, n);
        }
    }
    auto allreduce = [&](int slot, int count) {
      if (ctx.n_ranks > 1)
        comm_allreduce_sum(ctx, ctx.d_sums + slot, count);
    };

    cg_launch_init(ctx, w.r + off, w.p + off, dinv, n, status.p, cg_rr_slot(0), cg_rz_slot(0));
    allreduce(cg_rr_slot(0), 2);
    cg_launch_check0(ctx, status.p, cg_rr_slot(0), ctl.tolerance, ctl.reduce, ctl.max_steps);

    int      done = 0;
    unsigned last_step = 0;
    double   last_value = 0, initial = 0;
    cg_status_read(ctx, status.p, done, last_step, last_value, initial);

    Vector pv;
    pv.sys  = &sys;
    pv.d    = w.p;
    pv.owns = false;

    // p . A p inside the store epilogue of the tile kernel (no separate 16 B/DoF pass); GDM_CG_FUSED_DOT=0 disables
    const char *env_fd    = std::getenv("GDM_CG_FUSED_DOT");
    const bool  fused_dot = A.kernel_used == GDM_KERNEL_FUSED && !A.csr && fused_supports_dot(A) && !(env_fd && env_fd[0] == '0');
    unsigned it    = 0;
    unsigned batch = 4;
    while (done == 0)
      {
        const unsigned end = std::min<uint64_t>((uint64_t)it + batch, (uint64_t)ctl.max_steps);
        if (end == it)
          break;
        while (it < end)
          {
            ++it;
            if (fused_dot)
              fused_apply(A, w.q, w.p, false, true, SUM_PQ); // q = A p and p.q in one pass over p and q
            else
              {
                apply(A, w.q, w.p);
                blas_dot(ctx, w.p + off, w.q + off, n, SUM_PQ);
              }
            allreduce(SUM_PQ, 1);
            cg_launch_update(ctx, x.d + off, w.r + off, w.p + off, w.q + off, dinv, n, status.p, cg_rz_slot(it - 1),
                             cg_rr_slot(it), cg_rz_slot(it));
            allreduce(cg_rr_slot(it), 2);
            cg_launch_direction(ctx, w.r + off, w.p + off, dinv, n, status.p, cg_rz_slot(it - 1), cg_rr_slot(it),
                                cg_rz_slot(it), it, ctl.max_steps, ctl.tolerance);
          }
        cg_status_read(ctx, status.p, done, last_step, last_value, initial);
        batch = std::min(batch * 2, 16u); // at most 15 applies are enqueued behind a converged state (they are no-ops)
      }
    ctl.last_step     = last_step;
    ctl.last_value    = last_value;
    ctl.initial_value = initial;
    if (done == 1)
      return GDM_OK;
    return GDM_ERR_NO_CONVERGENCE;
  }
} // namespace gdm
