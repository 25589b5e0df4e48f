// Explicit Runge-Kutta driver with fused stage combinations.
//
// Stands in for dealii::TimeStepping::ExplicitRungeKutta<VectorType>::evolve_one_time_step(f, t, dt, y)
// (call sites: prototypes/advection_01_gdm.cc:268-281, applications/wave/include/gdm/wave/problem.h:330-345,
// applications/advection/include/gdm/advection/problem.h:91,169).  deal.II semantics: stages
// k_i = f(t + c_i dt, y + dt sum_j a_ij k_j), then y.sadd(1, dt b_i, k_i) in stage order.  Each stage
// input and the final update are ONE kernel each (blas_lincomb); stage vectors are allocated once.
#include "gdm_internal.h"

namespace gdm
{
  void blas_lincomb(Context &ctx, double *out, const double *y, int64_t n, int nt, const double *c,
                    const double *const *k);

  struct Rk
  {
    System                   *sys;
    int                       n_stages = 0, n_blocks = 1;
    std::vector<double>       a, b, c; // a[i*n_stages + j]
    std::vector<gdm_vector_t> Y;       // [block]
    std::vector<gdm_vector_t> K;       // [stage*n_blocks + block]
  };
} // namespace gdm

struct gdm_rk_s
{
  gdm::Rk impl;
};

using namespace gdm;

extern "C" {

int gdm_rk_create(gdm_system_t sys, int method, int n_blocks, gdm_rk_t *out)
{
  try
    {
      GDM_REQUIRE(sys && out, GDM_ERR_INVALID, "null argument");
      GDM_REQUIRE(n_blocks >= 1, GDM_ERR_INVALID, "n_blocks >= 1");
      std::unique_ptr<gdm_rk_s> r(new gdm_rk_s);
      Rk                       &rk = r->impl;
      rk.sys                       = &sys->impl;
      rk.n_blocks                  = n_blocks;
      switch (method)
        {
          case GDM_RK_FORWARD_EULER:
            rk.n_stages = 1;
            rk.a        = {0.0};
            rk.b        = {1.0};
            rk.c        = {0.0};
            break;
          case GDM_RK_THIRD_ORDER:
            rk.n_stages = 3;
            rk.a        = {0, 0, 0, 0.5, 0, 0, -1.0, 2.0, 0};
            rk.b        = {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0};
            rk.c        = {0.0, 0.5, 1.0};
            break;
          case GDM_RK_CLASSIC_FOURTH_ORDER:
            rk.n_stages = 4;
            rk.a        = {0, 0, 0, 0, 0.5, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 1.0, 0};
            rk.b        = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
            rk.c        = {0.0, 0.5, 0.5, 1.0};
            break;
          default:
            throw Error(GDM_ERR_NOT_IMPLEMENTED, "unknown Runge-Kutta method");
        }
      for (int i = 0; i < n_blocks * (1 + rk.n_stages); ++i)
        {
          gdm_vector_t v  = nullptr;
          const int    rc = gdm_vector_create(sys, &v);
          GDM_REQUIRE(rc == GDM_OK, rc, gdm_last_error());
          (i < n_blocks ? rk.Y : rk.K).push_back(v);
        }
      *out = r.release();
    }
  catch (const gdm::Error &e)
    {
      gdm::set_last_error(e.what());
      return e.code;
    }
  return GDM_OK;
}

int gdm_rk_destroy(gdm_rk_t rk)
{
  if (rk)
    {
      for (auto v : rk->impl.Y)
        gdm_vector_destroy(v);
      for (auto v : rk->impl.K)
        gdm_vector_destroy(v);
      delete rk;
    }
  return GDM_OK;
}

int gdm_rk_evolve_one_time_step(gdm_rk_t rkh, gdm_rk_rhs_fn f, void *user, double t, double dt, gdm_vector_t *y,
                                double *t_new)
{
  try
    {
      GDM_REQUIRE(rkh && f && y, GDM_ERR_INVALID, "null argument");
      Rk           &rk  = rkh->impl;
      Context      &ctx = *rk.sys->ctx;
      const Layout &L   = rk.sys->L;
      const int     S = rk.n_stages, NB = rk.n_blocks;
      for (int b = 0; b < NB; ++b)
        GDM_REQUIRE(y[b] && y[b]->impl.sys == rk.sys, GDM_ERR_INVALID, "vector/system mismatch");
      for (int i = 0; i < S; ++i)
        {
          for (int b = 0; b < NB; ++b)
            {
              double        c[4];
              const double *k[4];
              int           nt = 0;
              for (int j = 0; j < i; ++j)
                if (rk.a[i * S + j] != 0.0)
                  {
                    c[nt] = dt * rk.a[i * S + j];
                    k[nt] = rk.K[j * NB + b]->impl.d + L.own_off;
                    ++nt;
                  }
              blas_lincomb(ctx, rk.Y[b]->impl.d + L.own_off, y[b]->impl.d + L.own_off, L.own_len, nt, c, k);
            }
          const int rc = f(t + rk.c[i] * dt, rk.Y.data(), rk.K.data() + (size_t)i * NB, user);
          GDM_REQUIRE(rc == GDM_OK, rc, std::string("right-hand side callback failed: ") + gdm_last_error());
        }
      for (int b = 0; b < NB; ++b)
        {
          double        c[4];
          const double *k[4];
          for (int i = 0; i < S; ++i)
            {
              c[i] = dt * rk.b[i];
              k[i] = rk.K[i * NB + b]->impl.d + L.own_off;
            }
          blas_lincomb(ctx, y[b]->impl.d + L.own_off, y[b]->impl.d + L.own_off, L.own_len, S, c, k);
        }
      if (t_new)
        *t_new = t + dt;
    }
  catch (const gdm::Error &e)
    {
      gdm::set_last_error(e.what());
      return e.code;
    }
  return GDM_OK;
}

} // extern "C"
