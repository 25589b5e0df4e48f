// extern "C" entry points of libgdm_b200 (see include/gdm/cuda/gdm_c_api.h for the reference
// interfaces each one stands in for).  Host-side logic: layout/partition (system.h:703-761),
// constraints folded into the per-direction band tables, operator set-up, vector transfers.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "gdm_internal.h"

namespace gdm
{
  static thread_local std::string g_last_error;
  void set_last_error(const std::string &msg)
  {
    g_last_error = msg;
  }

  // ----------------------------------------------------------------- context
  void Context::ensure_scratch(size_t n)
  {
    if (n <= scratch_size)
      return;
    for (int i = 0; i < 4; ++i)
      {
        if (scratch[i])
          cudaFree(scratch[i]);
        scratch[i] = nullptr;
      }
    for (int i = 0; i < 4; ++i)
      {
        GDM_CUDA_CHECK(cudaMalloc(&scratch[i], n * sizeof(double)));
        GDM_CUDA_CHECK(cudaMemsetAsync(scratch[i], 0, n * sizeof(double), stream));
      }
    scratch_size = n;
  }

  double *Context::acquire(size_t n)
  {
    double *p = nullptr;
    for (size_t i = 0; i < pool_free.size(); ++i)
      if (pool_free[i].first == n)
        {
          p = pool_free[i].second;
          pool_free.erase(pool_free.begin() + i);
          break;
        }
    if (!p)
      GDM_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double)));
    GDM_CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(double), stream));
    pool_used.emplace_back(n, p);
    return p;
  }

  void Context::release(double *p)
  {
    if (!p)
      return;
    for (size_t i = 0; i < pool_used.size(); ++i)
      if (pool_used[i].second == p)
        {
          pool_free.push_back(pool_used[i]);
          pool_used.erase(pool_used.begin() + i);
          return;
        }
  }

  Context::~Context()
  {
    comm_destroy(*this);
    for (auto &e : pool_free)
      cudaFree(e.second);
    for (auto &e : pool_used)
      cudaFree(e.second);
    for (int i = 0; i < 4; ++i)
      if (scratch[i])
        cudaFree(scratch[i]);
    if (d_partials)
      cudaFree(d_partials);
    if (d_sums)
      cudaFree(d_sums);
    if (d_cg_status)
      cudaFree(d_cg_status);
    if (d_counters)
      cudaFree(d_counters);
    if (h_pinned)
      cudaFreeHost(h_pinned);
    if (comm_stream)
      cudaStreamDestroy(comm_stream);
    if (face_stream)
      cudaStreamDestroy(face_stream);
    if (h2d_stream)
      {
        cudaStreamDestroy(h2d_stream);
        cudaStreamDestroy(d2h_stream);
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 32; ++j)
            cudaEventDestroy(ev_pipe[i][j]);
      }
    if (ev_a)
      cudaEventDestroy(ev_a);
    if (ev_b)
      cudaEventDestroy(ev_b);
    if (ev_c)
      cudaEventDestroy(ev_c);
  }

  Vector::~Vector()
  {
    if (d && owns)
      cudaFree(d);
  }

  CsrOverlay::~CsrOverlay()
  {
    cudaFree(d_row_off);
    cudaFree(d_rowptr);
    cudaFree(d_col_rel);
    cudaFree(d_val);
    cudaFree(d_diag);
  }

  Constraints::~Constraints()
  {
    cudaFree(d_inhom);
  }

  Operator::~Operator()
  {
    for (int d = 0; d < 3; ++d)
      {
        cudaFree(dA[d]);
        cudaFree(dB[d]);
        cudaFree(ddiagA[d]);
        cudaFree(ddiagB[d]);
      }
    cudaFree(tmp);
    cudaFree(host_src);
    cudaFree(host_dst);
    cudaFree(stage_src);
    cudaFree(stage_dst);
    if (fused)
      fused_plan_destroy(*this);
    if (massinv)
      massinv_destroy(*this);
  }

  // ------------------------------------------------------------------ layout
  static void make_layout(const gdm_system_desc &desc, Layout &L)
  {
    GDM_REQUIRE(desc.dim >= 1 && desc.dim <= 3, GDM_ERR_INVALID, "dim must be 1, 2 or 3");
    GDM_REQUIRE(desc.fe_degree >= 1 && desc.fe_degree <= MAX_DEGREE && desc.fe_degree % 2 == 1,
                GDM_ERR_NOT_IMPLEMENTED, "fe_degree must be odd and <= 9 (fe.h:322)");
    GDM_REQUIRE(desc.n_components >= 1, GDM_ERR_INVALID, "n_components >= 1");
    GDM_REQUIRE(desc.n_ranks >= 1 && desc.rank >= 0 && desc.rank < desc.n_ranks, GDM_ERR_INVALID, "bad rank");
    L.dim = desc.dim;
    L.p   = desc.fe_degree;
    L.nc  = desc.n_components;
    for (int d = 0; d < 3; ++d)
      {
        if (d < L.dim)
          {
            GDM_REQUIRE((int)desc.n_subdivisions[d] >= L.p, GDM_ERR_INVALID,
                        "n_subdivisions must be >= fe_degree in every direction");
            GDM_REQUIRE(desc.hi[d] > desc.lo[d], GDM_ERR_INVALID, "empty domain");
            L.N[d]  = (int)desc.n_subdivisions[d];
            L.nn[d] = L.N[d] + 1;
            L.lo[d] = desc.lo[d];
            L.hi[d] = desc.hi[d];
            L.h[d]  = (desc.hi[d] - desc.lo[d]) / L.N[d];
          }
        else
          {
            L.N[d]  = 0;
            L.nn[d] = 1;
            L.lo[d] = 0;
            L.hi[d] = 1;
            L.h[d]  = 1;
          }
      }
    L.rank    = desc.rank;
    L.n_ranks = desc.n_ranks;
    L.pdim    = L.dim - 1;
    L.ghost   = L.p + (desc.add_ghost_layer ? 1 : 0);
    // system.h:729-737
    const int n_last = L.nn[L.pdim];
    const int stride = (L.N[L.pdim] + L.n_ranks - 1) / L.n_ranks;
    const int start  = (L.rank == 0) ? 0 : (stride * L.rank + 1);
    const int end    = stride * (L.rank + 1) + 1;
    L.own0           = std::min(start, n_last);
    L.own1           = std::min(end, n_last);
    L.loc0           = std::max(0, L.own0 - L.ghost);
    L.loc1           = std::min(n_last, L.own1 + L.ghost);
    if (L.own1 <= L.own0)
      L.loc0 = L.loc1 = L.own0 = L.own1; // empty rank
    for (int d = 0; d < 3; ++d)
      L.ln[d] = L.nn[d];
    L.ln[L.pdim] = L.loc1 - L.loc0;
    // rows start on 32-byte sector boundaries (and satisfy TMA's 16-byte global stride rule)
    const int64_t row = (int64_t)L.ln[0] * L.nc;
    L.pitch           = (L.dim == 1) ? std::max<int64_t>(round_up(row, 4), 4) : round_up(row, 4);
    L.plane           = L.pitch * L.ln[1];
    L.size            = std::max<int64_t>(L.plane * L.ln[2], 4);
    L.stride[0]       = L.nc;
    L.stride[1]       = L.pitch;
    L.stride[2]       = L.plane;
    L.own_off         = (int64_t)(L.own0 - L.loc0) * L.stride[L.pdim];
    L.own_len         = (int64_t)(L.own1 - L.own0) * L.stride[L.pdim];
    int64_t face      = L.nc;
    for (int d = 0; d < L.pdim; ++d)
      face *= L.nn[d];
    L.n_owned       = face * (L.own1 - L.own0);
    L.n_dofs_global = face * L.nn[L.pdim];
  }

  // --------------------------------------------------------- operator tables
  static void build_tables(Operator &op, const bool upload_to_device = true)
  {
    const Layout &L = op.sys->L;
    const int     p = L.p, W = 2 * p + 1;
    for (int d = 0; d < L.dim; ++d)
      {
        const int           N = L.N[d];
        std::vector<double> M, B;
        band_matrix_1d(p, N, 0, M);
        for (auto &v : M)
          v *= L.h[d];
        switch (op.desc.kind)
          {
            case GDM_OP_MASS:
              break;
            case GDM_OP_STIFFNESS:
              band_matrix_1d(p, N, 1, B);
              for (auto &v : B)
                v /= L.h[d];
              break;
            case GDM_OP_ADVECTION:
              band_matrix_1d(p, N, 2, B);
              for (auto &v : B)
                v *= op.desc.b[d];
              break;
            case GDM_OP_ADVECTION_T:
              {
                std::vector<double> C;
                band_matrix_1d(p, N, 2, C);
                B.assign(C.size(), 0.0);
                for (int i = 0; i <= N; ++i)
                  for (int t = 0; t < W; ++t)
                    {
                      const int j = i + t - p;
                      if (j < 0 || j > N)
                        continue;
                      B[(size_t)j * W + (i - j + p)] = op.desc.b[d] * C[(size_t)i * W + t];
                    }
                break;
              }
            default:
              throw Error(GDM_ERR_INVALID, "unknown operator kind");
          }
        // unconstrained diagonals (value of constrained rows: sum_cells |cell_matrix(i,i)|)
        std::vector<double> diagA(N + 1), diagB(N + 1, 0.0);
        for (int i = 0; i <= N; ++i)
          {
            diagA[i] = M[(size_t)i * W + p];
            if (!B.empty())
              diagB[i] = B[(size_t)i * W + p];
          }
        auto constrain = [&](std::vector<double> &T) {
          if (T.empty())
            return;
          if (op.periodic[d])
            {
              // fold row/column N into row/column 0 (C^T A C); reads wrap modulo N in the kernels
              for (int t = 0; t < W; ++t)
                {
                  T[t] += T[(size_t)N * W + t];
                  T[(size_t)N * W + t] = 0.0;
                }
            }
          for (int s = 0; s < 2; ++s)
            if (op.dirichlet[d][s])
              {
                const int r = (s == 0) ? 0 : N;
                for (int t = 0; t < W; ++t)
                  T[(size_t)r * W + t] = 0.0;
                for (int i = 0; i <= N; ++i)
                  for (int t = 0; t < W; ++t)
                    {
                      int c = i + t - p;
                      if (op.periodic[d] && i < N)
                        {
                          if (c < 0)
                            c += N;
                          else if (c >= N)
                            c -= N;
                        }
                      if (c == r)
                        T[(size_t)i * W + t] = 0.0;
                    }
              }
        };
        GDM_REQUIRE(!(op.periodic[d] && N <= 2 * p), GDM_ERR_NOT_IMPLEMENTED,
                    "periodic direction needs more than 2*fe_degree cells");
        // slice the local rows of the partitioned direction
        const int r0 = (d == L.pdim) ? L.loc0 : 0;
        const int nr = L.ln[d];
        if (op.periodic[d])
          {
            op.hAu[d].assign(M.begin() + (size_t)r0 * W, M.begin() + (size_t)(r0 + nr) * W);
            if (!B.empty())
              op.hBu[d].assign(B.begin() + (size_t)r0 * W, B.begin() + (size_t)(r0 + nr) * W);
          }
        constrain(M);
        constrain(B);
        op.hA[d].assign(M.begin() + (size_t)r0 * W, M.begin() + (size_t)(r0 + nr) * W);
        op.hdiagA[d].assign(diagA.begin() + r0, diagA.begin() + r0 + nr);
        if (!B.empty())
          {
            op.hB[d].assign(B.begin() + (size_t)r0 * W, B.begin() + (size_t)(r0 + nr) * W);
            op.hdiagB[d].assign(diagB.begin() + r0, diagB.begin() + r0 + nr);
          }
        auto upload = [&](const std::vector<double> &h, double *&dptr) {
          if (h.empty() || !upload_to_device)
            return;
          GDM_CUDA_CHECK(cudaMalloc(&dptr, h.size() * sizeof(double)));
          GDM_CUDA_CHECK(cudaMemcpy(dptr, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        };
        upload(op.hA[d], op.dA[d]);
        upload(op.hB[d], op.dB[d]);
        upload(op.hdiagA[d], op.ddiagA[d]);
        upload(op.hdiagB[d], op.ddiagB[d]);
      }
  }

  static void operator_apply(Operator &op, Vector &dst, Vector &src, bool accumulate)
  {
    GDM_REQUIRE(dst.sys == op.sys && src.sys == op.sys, GDM_ERR_INVALID, "vector/operator system mismatch");
    GDM_REQUIRE(dst.d != src.d, GDM_ERR_INVALID, "vmult: dst and src must not alias");
    Context &ctx = *op.sys->ctx;
    const bool overlap = (op.kernel_used == GDM_KERNEL_FUSED); // the fused path imports the ghosts itself, overlapped
    if (!overlap)
      vector_update_ghosts(src);
    // y += A x through a temporary: CSR overlay rows replace rows, and the fused periodic path folds rows of its output
    const bool fused_periodic = op.kernel_used == GDM_KERNEL_FUSED && (op.periodic[0] || op.periodic[1] || op.periodic[2]);
    const bool via_tmp = accumulate && (op.csr || fused_periodic);
    double    *out     = dst.d;
    if (via_tmp)
      {
        if (!op.tmp)
          {
            GDM_CUDA_CHECK(cudaMalloc(&op.tmp, (size_t)op.sys->L.size * sizeof(double)));
            GDM_CUDA_CHECK(cudaMemsetAsync(op.tmp, 0, (size_t)op.sys->L.size * sizeof(double), ctx.stream));
          }
        out = op.tmp;
      }
    const bool acc = accumulate && !via_tmp;
    if (op.kernel_used == GDM_KERNEL_FUSED)
      fused_apply(op, out, src.d, acc, true);
    else
      generic_apply(op, out, src.d, acc);
    if (op.csr)
      launch_csr_overlay(ctx, *op.csr, out, src.d, false);
    if (via_tmp)
      blas_sadd(ctx, dst.d + op.sys->L.own_off, 1.0, 1.0, out + op.sys->L.own_off, op.sys->L.own_len);
  }

  void vector_update_ghosts(Vector &v)
  {
    const Layout &L = v.sys->L;
    if (L.n_ranks == 1)
      return;
    comm_halo_exchange(*v.sys->ctx, L, v.d);
  }

  // storage offset of a global DoF index (must be stored locally)
  static int64_t storage_offset(const Layout &L, uint64_t dof)
  {
    const uint64_t node = dof / L.nc;
    const int      c    = (int)(dof % L.nc);
    int            idx[3];
    idx[0] = (int)(node % L.nn[0]);
    idx[1] = (int)((node / L.nn[0]) % L.nn[1]);
    idx[2] = (int)(node / ((uint64_t)L.nn[0] * L.nn[1]));
    idx[L.pdim] -= L.loc0;
    GDM_REQUIRE(idx[L.pdim] >= 0 && idx[L.pdim] < L.ln[L.pdim], GDM_ERR_INVALID,
                "DoF outside the locally stored range");
    return (int64_t)idx[2] * L.plane + (int64_t)idx[1] * L.pitch + (int64_t)idx[0] * L.nc + c;
  }
} // namespace gdm

using namespace gdm;

#define GDM_TRY try {
#define GDM_CATCH                                 \
  }                                               \
  catch (const gdm::Error &e)                     \
  {                                               \
    gdm::set_last_error(e.what());                \
    return e.code;                                \
  }                                               \
  catch (const std::exception &e)                 \
  {                                               \
    gdm::set_last_error(e.what());                \
    return GDM_ERR_INTERNAL;                      \
  }                                               \
  return GDM_OK;

#define GDM_ARG(x) GDM_REQUIRE((x) != nullptr, GDM_ERR_INVALID, "null argument " #x)

extern "C" {

const char *gdm_last_error(void)
{
  return gdm::g_last_error.c_str();
}

int gdm_api_version(void)
{
  return GDM_API_VERSION;
}

// ------------------------------------------------------------------ context
int gdm_context_create(int device, void *stream, gdm_context_t *out)
{
  GDM_TRY
  GDM_ARG(out);
  if (device < 0)
    {
      // description-only context: grid / DoF / partition queries work, every compute entry point
      // fails with GDM_ERR_CUDA (used by the CPU-side tests of the host logic)
      std::unique_ptr<gdm_context_s> c(new gdm_context_s);
      c->impl.device = -1;
      *out           = c.release();
      return GDM_OK;
    }
  int n_dev = 0;
  GDM_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
  GDM_REQUIRE(n_dev > 0 && device >= 0 && device < n_dev, GDM_ERR_CUDA,
              "no usable CUDA device (this library has no CPU fallback)");
  GDM_CUDA_CHECK(cudaSetDevice(device));
  std::unique_ptr<gdm_context_s> c(new gdm_context_s);
  Context                       &ctx = c->impl;
  ctx.device                         = device;
  ctx.stream                         = (cudaStream_t)stream;
  cudaDeviceProp prop;
  GDM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  ctx.sm_count         = prop.multiProcessorCount;
  ctx.partial_capacity = (size_t)ctx.sm_count * 8 * 4;
  GDM_CUDA_CHECK(cudaMalloc(&ctx.d_partials, ctx.partial_capacity * sizeof(double)));
  GDM_CUDA_CHECK(cudaMalloc(&ctx.d_sums, N_SUM_SLOTS * sizeof(double)));
  GDM_CUDA_CHECK(cudaMalloc(&ctx.d_counters, 16 * sizeof(unsigned)));
  GDM_CUDA_CHECK(cudaMemset(ctx.d_sums, 0, N_SUM_SLOTS * sizeof(double)));
  GDM_CUDA_CHECK(cudaMemset(ctx.d_counters, 0, 16 * sizeof(unsigned)));
  GDM_CUDA_CHECK(cudaMallocHost(&ctx.h_pinned, 64 * sizeof(double)));
  {
    // highest priority: ghost imports and the slab-face launches behind them overtake queued interior CTAs
    int lo_prio = 0, hi_prio = 0;
    GDM_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    GDM_CUDA_CHECK(cudaStreamCreateWithPriority(&ctx.comm_stream, cudaStreamNonBlocking, hi_prio));
    // lowest priority: the constrained-row (face) kernel of an apply must not take CTA slots from the persistent tile
    // kernel at its start; its blocks run as the first tile CTAs retire (kron3d.cu)
    GDM_CUDA_CHECK(cudaStreamCreateWithPriority(&ctx.face_stream, cudaStreamNonBlocking, lo_prio));
  }
  GDM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx.ev_a, cudaEventDisableTiming));
  GDM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx.ev_b, cudaEventDisableTiming));
  GDM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx.ev_c, cudaEventDisableTiming));
  *out = c.release();
  GDM_CATCH
}

int gdm_context_destroy(gdm_context_t ctx)
{
  GDM_TRY
  if (ctx)
    {
      if (ctx->impl.device >= 0)
        {
          cudaSetDevice(ctx->impl.device);
          cudaStreamSynchronize(ctx->impl.stream);
        }
      delete ctx;
    }
  GDM_CATCH
}

int gdm_context_set_stream(gdm_context_t ctx, void *stream)
{
  GDM_TRY
  GDM_ARG(ctx);
  ctx->impl.stream = (cudaStream_t)stream;
  GDM_CATCH
}

int gdm_context_synchronize(gdm_context_t ctx)
{
  GDM_TRY
  GDM_ARG(ctx);
  if (ctx->impl.device >= 0)
    GDM_CUDA_CHECK(cudaStreamSynchronize(ctx->impl.stream));
  GDM_CATCH
}

int gdm_context_launch_count(gdm_context_t ctx, uint64_t *count)
{
  GDM_TRY
  GDM_ARG(ctx);
  GDM_ARG(count);
  *count = ctx->impl.launches;
  GDM_CATCH
}

int gdm_comm_unique_id(void *id128)
{
  GDM_TRY
  GDM_ARG(id128);
  comm_unique_id(id128);
  GDM_CATCH
}

int gdm_context_comm_init(gdm_context_t ctx, const void *id128, int rank, int n_ranks)
{
  GDM_TRY
  GDM_ARG(ctx);
  GDM_ARG(id128);
  comm_init(ctx->impl, id128, rank, n_ranks);
  GDM_CATCH
}

// -------------------------------------------------------------------- basis
int gdm_polynomials_1d(int p, double *coeffs)
{
  GDM_TRY
  GDM_ARG(coeffs);
  GDM_REQUIRE(p >= 1 && p <= MAX_DEGREE && p % 2 == 1, GDM_ERR_NOT_IMPLEMENTED, "fe_degree must be odd and <= 9");
  for (int v = 0; v < p; ++v)
    lagrange_monomials(p, v, coeffs + (size_t)v * (p + 1) * (p + 1));
  GDM_CATCH
}

// ------------------------------------------------------------------- system
int gdm_system_create(gdm_context_t ctx, const gdm_system_desc *desc, gdm_system_t *out)
{
  GDM_TRY
  GDM_ARG(ctx);
  GDM_ARG(desc);
  GDM_ARG(out);
  std::unique_ptr<gdm_system_s> s(new gdm_system_s);
  s->impl.ctx  = &ctx->impl;
  s->impl.desc = *desc;
  make_layout(*desc, s->impl.L);
  GDM_REQUIRE(desc->n_ranks == 1 || ctx->impl.device < 0 || ctx->impl.n_ranks == desc->n_ranks, GDM_ERR_INVALID,
              "system n_ranks does not match the context communicator (call gdm_context_comm_init first)");
  GDM_REQUIRE(desc->n_ranks == 1 || ctx->impl.device < 0 || ctx->impl.rank == desc->rank, GDM_ERR_INVALID,
              "system rank does not match the rank of the context communicator");
  if (desc->n_ranks > 1 && ctx->impl.device >= 0) // (description-only contexts never exchange)
    {
      // every rank checks EVERY slab: a slab thinner than the ghost zone must fail identically on all ranks, not only
      // on the rank that owns it (its neighbours would otherwise enter the exchange and wait for it forever)
      for (int r = 0; r < desc->n_ranks; ++r)
        {
          gdm_system_desc dr = *desc;
          dr.rank            = r;
          Layout Lr;
          make_layout(dr, Lr);
          const int own = Lr.own1 - Lr.own0;
          GDM_REQUIRE(own <= 0 || own >= Lr.ghost, GDM_ERR_NOT_IMPLEMENTED,
                      "slab of rank " + std::to_string(r) + " is thinner than the ghost zone: use fewer ranks or a larger grid");
        }
    }
  *out = s.release();
  GDM_CATCH
}

int gdm_system_destroy(gdm_system_t sys)
{
  delete sys;
  return GDM_OK;
}

uint64_t gdm_system_n_dofs(gdm_system_t sys)
{
  return sys ? (uint64_t)sys->impl.L.n_dofs_global : 0;
}

uint64_t gdm_system_n_cells(gdm_system_t sys)
{
  if (!sys)
    return 0;
  uint64_t n = 1;
  for (int d = 0; d < sys->impl.L.dim; ++d)
    n *= sys->impl.L.N[d];
  return n;
}

int gdm_system_locally_owned_range(gdm_system_t sys, uint64_t *begin, uint64_t *end)
{
  GDM_TRY
  GDM_ARG(sys);
  const Layout &L    = sys->impl.L;
  uint64_t      face = L.nc;
  for (int d = 0; d < L.pdim; ++d)
    face *= L.nn[d];
  if (begin)
    *begin = face * L.own0;
  if (end)
    *end = face * L.own1;
  GDM_CATCH
}

int gdm_system_dofs_per_cell(gdm_system_t sys)
{
  if (!sys)
    return 0;
  int n = sys->impl.L.nc;
  for (int d = 0; d < sys->impl.L.dim; ++d)
    n *= sys->impl.L.p + 1;
  return n;
}

int gdm_system_get_dof_indices(gdm_system_t sys, uint64_t cell, uint64_t *out)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(out);
  const Layout &L = sys->impl.L;
  GDM_REQUIRE(cell < gdm_system_n_cells(sys), GDM_ERR_INVALID, "cell index out of range");
  int      ci[3] = {0, 0, 0}, off[3] = {0, 0, 0};
  uint64_t c = cell;
  for (int d = 0; d < L.dim; ++d)
    {
      ci[d] = (int)(c % L.N[d]);
      c /= L.N[d];
      off[d] = window_offset(L.p, L.N[d], ci[d]);
    }
  const int n1  = L.p + 1;
  const int npc = gdm_system_dofs_per_cell(sys) / L.nc;
  int       cc  = 0;
  for (int k = 0; k < (L.dim >= 3 ? n1 : 1); ++k)
    for (int j = 0; j < (L.dim >= 2 ? n1 : 1); ++j)
      for (int i = 0; i < n1; ++i, ++cc)
        {
          const uint64_t node = (uint64_t)(off[0] + i) + (uint64_t)L.nn[0] * ((off[1] + j) + (uint64_t)L.nn[1] * (off[2] + k));
          for (int comp = 0; comp < L.nc; ++comp)
            out[comp * npc + cc] = node * L.nc + comp;
        }
  GDM_CATCH
}

int gdm_system_active_fe_index(gdm_system_t sys, uint64_t cell, uint32_t *fe_index)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(fe_index);
  const Layout &L = sys->impl.L;
  GDM_REQUIRE(cell < gdm_system_n_cells(sys), GDM_ERR_INVALID, "cell index out of range");
  uint64_t c = cell;
  uint32_t idx = 0, mult = 1;
  for (int d = 0; d < L.dim; ++d)
    {
      const int ci = (int)(c % L.N[d]);
      c /= L.N[d];
      idx += mult * (uint32_t)cell_variant(L.p, L.N[d], ci);
      mult *= (uint32_t)L.p;
    }
  *fe_index = idx;
  GDM_CATCH
}

// Sparsity patterns (include/gdm/system.h:586-630) without materialising them.  On the Cartesian grid the pattern of a
// node is a union of boxes: with W(c) = [off(c), off(c)+p] the window of cell c in one direction and C(i) the cells whose
// window holds node i,  I(i) = U_{c in C(i)} W(c)  (cell coupling)  and  J(i) = I(i) U U_{c in C(i)} W(c-1) U W(c+1)
// (face neighbours, flux pattern); the row of node (i0,i1,i2) is  prod_d I_d  (create_sparsity_pattern) or
// U_d ( J_d x prod_{e != d} I_e )  (create_flux_sparsity_pattern).  Constrained rows and columns are kept, as
// AffineConstraints::add_entries_local_to_global does by default; periodic constraints (redirected entries) are not
// covered here.
namespace
{
  void coupling_intervals(int p, int N, int i, bool flux, int &lo, int &hi)
  {
    lo = N + 1;
    hi = -1;
    for (int c = 0; c < N; ++c)
      {
        const int o = gdm::window_offset(p, N, c);
        if (i < o || i > o + p)
          continue;
        for (int cc = (flux ? c - 1 : c); cc <= (flux ? c + 1 : c); ++cc)
          {
            if (cc < 0 || cc >= N)
              continue;
            const int oo = gdm::window_offset(p, N, cc);
            lo           = std::min(lo, oo);
            hi           = std::max(hi, oo + p);
          }
      }
  }
} // namespace

int gdm_system_sparsity_row(gdm_system_t sys, int flux, uint64_t row, uint64_t *cols, uint64_t cap, uint64_t *n_cols)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(n_cols);
  const Layout &L = sys->impl.L;
  GDM_REQUIRE(row < (uint64_t)L.n_dofs_global, GDM_ERR_INVALID, "row out of range");
  const uint64_t node = row / L.nc;
  int            idx[3] = {0, 0, 0};
  idx[0] = (int)(node % L.nn[0]);
  idx[1] = (int)((node / L.nn[0]) % L.nn[1]);
  idx[2] = (int)(node / ((uint64_t)L.nn[0] * L.nn[1]));
  int Il[3] = {0, 0, 0}, Ih[3] = {0, 0, 0}, Jl[3] = {0, 0, 0}, Jh[3] = {0, 0, 0};
  for (int d = 0; d < L.dim; ++d)
    {
      coupling_intervals(L.p, L.N[d], idx[d], false, Il[d], Ih[d]);
      coupling_intervals(L.p, L.N[d], idx[d], flux != 0, Jl[d], Jh[d]);
    }
  uint64_t n = 0;
  // enumerate the bounding box prod_d J_d in DoF order and keep the points that lie in at least one of the boxes
  for (int k = Jl[2]; k <= Jh[2]; ++k)
    for (int j = Jl[1]; j <= Jh[1]; ++j)
      for (int i = Jl[0]; i <= Jh[0]; ++i)
        {
          const int  q[3]  = {i, j, k};
          int        n_out = 0; // directions in which the point leaves the cell-coupling interval
          for (int d = 0; d < L.dim; ++d)
            n_out += (q[d] < Il[d] || q[d] > Ih[d]) ? 1 : 0;
          if (n_out > 1)
            continue;
          const uint64_t nd = (uint64_t)i + (uint64_t)L.nn[0] * ((uint64_t)j + (uint64_t)L.nn[1] * (uint64_t)k);
          for (int c = 0; c < L.nc; ++c, ++n)
            if (cols != nullptr && n < cap)
              cols[n] = nd * L.nc + c;
        }
  *n_cols = n;
  GDM_REQUIRE(cols == nullptr || n <= cap, GDM_ERR_INVALID, "column buffer too small");
  GDM_CATCH
}

int gdm_system_matrix_1d(gdm_system_t sys, int d, int kind, double *band)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(band);
  const Layout &L = sys->impl.L;
  GDM_REQUIRE(d >= 0 && d < L.dim && kind >= 0 && kind <= 2, GDM_ERR_INVALID, "bad direction or kind");
  std::vector<double> t;
  band_matrix_1d(L.p, L.N[d], kind, t);
  const double s = (kind == 0) ? L.h[d] : (kind == 1 ? 1.0 / L.h[d] : 1.0);
  for (size_t i = 0; i < t.size(); ++i)
    band[i] = t[i] * s;
  GDM_CATCH
}

int gdm_system_layout(gdm_system_t sys, gdm_layout_info *info)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(info);
  const Layout &L    = sys->impl.L;
  info->pitch        = L.pitch;
  info->plane        = L.plane;
  info->size         = L.size;
  info->owned_offset = L.own_off;
  info->owned_size   = L.own_len;
  for (int d = 0; d < 3; ++d)
    info->local_nodes[d] = L.ln[d];
  info->owned_begin  = L.own0;
  info->owned_end    = L.own1;
  info->stored_begin = L.loc0;
  info->stored_end   = L.loc1;
  GDM_CATCH
}

int gdm_system_halo_plan(gdm_system_t sys, int32_t *plan10)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(plan10);
  const HaloPlan h = halo_plan(sys->impl.L);
  const int      v[10] = {h.prev, h.next, h.send_lo_plane, h.send_lo_count, h.recv_lo_plane, h.recv_lo_count,
                          h.send_hi_plane, h.send_hi_count, h.recv_hi_plane, h.recv_hi_count};
  for (int i = 0; i < 10; ++i)
    plan10[i] = v[i];
  GDM_CATCH
}

int gdm_pers_partition(int tiles_x, int tiles_y, int k0, int k1, int slots, int min_len, int aligned, const int32_t *weights, int32_t *job_ptr,
                       int32_t cap_ptr, int32_t *jobs6, int32_t cap_jobs, int32_t *n_shares, int32_t *n_jobs)
{
  GDM_TRY
  GDM_ARG(n_shares);
  GDM_ARG(n_jobs);
  std::vector<int> ptr, jobs;
  pers_partition_host(tiles_x, tiles_y, k0, k1, slots, min_len, aligned, weights, ptr, jobs);
  *n_shares = (int32_t)ptr.size() - 1;
  *n_jobs   = (int32_t)(jobs.size() / 6);
  GDM_REQUIRE(job_ptr != nullptr && jobs6 != nullptr && cap_ptr >= (int32_t)ptr.size() && cap_jobs >= *n_jobs, GDM_ERR_INVALID,
              "partition buffers too small");
  for (size_t i = 0; i < ptr.size(); ++i)
    job_ptr[i] = ptr[i];
  for (size_t i = 0; i < jobs.size(); ++i)
    jobs6[i] = jobs[i];
  GDM_CATCH
}

// -------------------------------------------------------------- constraints
int gdm_constraints_create(gdm_system_t sys, gdm_constraints_t *out)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(out);
  *out            = new gdm_constraints_s;
  (*out)->impl.sys = &sys->impl;
  GDM_CATCH
}

int gdm_constraints_destroy(gdm_constraints_t c)
{
  delete c;
  return GDM_OK;
}

int gdm_constraints_make_zero_boundary(gdm_constraints_t c, int surface)
{
  GDM_TRY
  GDM_ARG(c);
  const int dim = c->impl.sys->L.dim;
  GDM_REQUIRE(surface >= -1 && surface < 2 * dim, GDM_ERR_INVALID, "surface out of range");
  GDM_REQUIRE(!c->impl.closed, GDM_ERR_INVALID, "constraints already closed");
  for (int s = 0; s < 2 * dim; ++s)
    if (surface < 0 || surface == s)
      c->impl.dirichlet[s / 2][s % 2] = true;
  GDM_CATCH
}

int gdm_constraints_make_periodicity(gdm_constraints_t c, int d)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_REQUIRE(d >= 0 && d < c->impl.sys->L.dim, GDM_ERR_INVALID, "direction out of range");
  GDM_REQUIRE(!c->impl.closed, GDM_ERR_INVALID, "constraints already closed");
  // system.h:454: DoFs that are already constrained keep their constraint
  if (!c->impl.dirichlet[d][1])
    c->impl.periodic[d] = true;
  GDM_CATCH
}

int gdm_constraints_close(gdm_constraints_t c)
{
  GDM_TRY
  GDM_ARG(c);
  // a periodic slave whose master is zero-constrained resolves to zero as well
  for (int d = 0; d < 3; ++d)
    if (c->impl.periodic[d] && c->impl.dirichlet[d][0])
      {
        c->impl.periodic[d]     = false;
        c->impl.dirichlet[d][1] = true;
      }
  c->impl.closed = true;
  GDM_CATCH
}

static bool node_constrained(const Constraints &c, const int idx[3])
{
  const Layout &L = c.sys->L;
  for (int d = 0; d < L.dim; ++d)
    {
      if (idx[d] == 0 && c.dirichlet[d][0])
        return true;
      if (idx[d] == L.N[d] && (c.dirichlet[d][1] || c.periodic[d]))
        return true;
    }
  return false;
}

uint64_t gdm_constraints_n_constraints(gdm_constraints_t c)
{
  if (!c)
    return 0;
  const Layout &L = c->impl.sys->L;
  // inclusion-exclusion over directions: count unconstrained nodes
  uint64_t free_nodes = 1, all = 1;
  for (int d = 0; d < L.dim; ++d)
    {
      int f = L.nn[d];
      if (c->impl.dirichlet[d][0])
        --f;
      if (c->impl.dirichlet[d][1] || c->impl.periodic[d])
        --f;
      free_nodes *= (uint64_t)std::max(f, 0);
      all *= (uint64_t)L.nn[d];
    }
  return (all - free_nodes) * L.nc;
}

int gdm_constraints_is_constrained(gdm_constraints_t c, uint64_t dof)
{
  if (!c)
    return 0;
  const Layout &L    = c->impl.sys->L;
  const uint64_t node = dof / L.nc;
  int            idx[3];
  idx[0] = (int)(node % L.nn[0]);
  idx[1] = (int)((node / L.nn[0]) % L.nn[1]);
  idx[2] = (int)(node / ((uint64_t)L.nn[0] * L.nn[1]));
  return node_constrained(c->impl, idx) ? 1 : 0;
}

int gdm_constraints_distribute(gdm_constraints_t c, gdm_vector_t v)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(v);
  GDM_REQUIRE(v->impl.sys == c->impl.sys, GDM_ERR_INVALID, "vector/constraints system mismatch");
  Context   &ctx = *c->impl.sys->ctx;
  const bool no_periodic[3] = {false, false, false};
  launch_set_constrained(ctx, c->impl.sys->L, c->impl.dirichlet, no_periodic, v->impl.d, 0.0);
  if (c->impl.d_inhom) // the boundary nodes are zero now, the inhomogeneity vector is zero everywhere else
    blas_sadd(ctx, v->impl.d + c->impl.sys->L.own_off, 1.0, 1.0, c->impl.d_inhom + c->impl.sys->L.own_off, c->impl.sys->L.own_len);
  launch_periodic_copy(ctx, c->impl.sys->L, c->impl.periodic, v->impl.d);
  GDM_CATCH
}

int gdm_constraints_set_zero(gdm_constraints_t c, gdm_vector_t v)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(v);
  GDM_REQUIRE(v->impl.sys == c->impl.sys, GDM_ERR_INVALID, "vector/constraints system mismatch");
  launch_set_constrained(*c->impl.sys->ctx, c->impl.sys->L, c->impl.dirichlet, c->impl.periodic, v->impl.d, 0.0);
  GDM_CATCH
}

// ------------------------------------------------------------------ vectors
int gdm_vector_create(gdm_system_t sys, gdm_vector_t *out)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(out);
  GDM_REQUIRE(sys->impl.ctx->device >= 0, GDM_ERR_CUDA, "description-only context: no CUDA device (no CPU fallback)");
  std::unique_ptr<gdm_vector_s> v(new gdm_vector_s);
  v->impl.sys = &sys->impl;
  GDM_CUDA_CHECK(cudaSetDevice(sys->impl.ctx->device));
  GDM_CUDA_CHECK(cudaMalloc(&v->impl.d, (size_t)sys->impl.L.size * sizeof(double)));
  GDM_CUDA_CHECK(cudaMemsetAsync(v->impl.d, 0, (size_t)sys->impl.L.size * sizeof(double), sys->impl.ctx->stream));
  *out = v.release();
  GDM_CATCH
}

int gdm_vector_destroy(gdm_vector_t v)
{
  delete v;
  return GDM_OK;
}

void *gdm_vector_device_ptr(gdm_vector_t v)
{
  return v ? v->impl.d : nullptr;
}

static void vector_transfer(Vector &v, double *host, bool upload)
{
  const Layout &L   = v.sys->L;
  Context      &ctx = *v.sys->ctx;
  double       *dev = v.d + L.own_off;
  if (L.own1 <= L.own0)
    return;
  if (L.dim == 1)
    {
      const size_t bytes = (size_t)(L.own1 - L.own0) * L.nc * sizeof(double);
      GDM_CUDA_CHECK(upload ? cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx.stream) :
                              cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx.stream));
    }
  else
    {
      const size_t width = (size_t)L.ln[0] * L.nc * sizeof(double);
      const size_t rows  = (L.dim == 2) ? (size_t)(L.own1 - L.own0) : (size_t)(L.own1 - L.own0) * L.ln[1];
      const size_t dpitch = (size_t)L.pitch * sizeof(double);
      GDM_CUDA_CHECK(upload ? cudaMemcpy2DAsync(dev, dpitch, host, width, width, rows, cudaMemcpyHostToDevice, ctx.stream) :
                              cudaMemcpy2DAsync(host, width, dev, dpitch, width, rows, cudaMemcpyDeviceToHost, ctx.stream));
    }
}

int gdm_vector_upload(gdm_vector_t v, const double *host)
{
  GDM_TRY
  GDM_ARG(v);
  GDM_ARG(host);
  vector_transfer(v->impl, const_cast<double *>(host), true);
  GDM_CUDA_CHECK(cudaStreamSynchronize(v->impl.sys->ctx->stream));
  GDM_CATCH
}

int gdm_vector_download(gdm_vector_t v, double *host)
{
  GDM_TRY
  GDM_ARG(v);
  GDM_ARG(host);
  vector_transfer(v->impl, host, false);
  GDM_CUDA_CHECK(cudaStreamSynchronize(v->impl.sys->ctx->stream));
  GDM_CATCH
}

int gdm_vector_set(gdm_vector_t v, double value)
{
  GDM_TRY
  GDM_ARG(v);
  const Layout &L = v->impl.sys->L;
  if (value == 0.0)
    GDM_CUDA_CHECK(cudaMemsetAsync(v->impl.d, 0, (size_t)L.size * sizeof(double), v->impl.sys->ctx->stream));
  else
    blas_set_strided(*v->impl.sys->ctx, L, v->impl.d, value);
  GDM_CATCH
}

#define GDM_SAME_SYS(a, b) GDM_REQUIRE((a)->impl.sys == (b)->impl.sys, GDM_ERR_INVALID, "vectors belong to different systems")

int gdm_vector_copy(gdm_vector_t dst, gdm_vector_t src)
{
  GDM_TRY
  GDM_ARG(dst);
  GDM_ARG(src);
  GDM_SAME_SYS(dst, src);
  const Layout &L = dst->impl.sys->L;
  blas_copy(*dst->impl.sys->ctx, dst->impl.d + L.own_off, src->impl.d + L.own_off, L.own_len);
  GDM_CATCH
}

int gdm_vector_scale(gdm_vector_t v, double a)
{
  GDM_TRY
  GDM_ARG(v);
  const Layout &L = v->impl.sys->L;
  blas_scale(*v->impl.sys->ctx, v->impl.d + L.own_off, L.own_len, a);
  GDM_CATCH
}

int gdm_vector_add(gdm_vector_t v, double a, gdm_vector_t x)
{
  GDM_TRY
  GDM_ARG(v);
  GDM_ARG(x);
  GDM_SAME_SYS(v, x);
  const Layout &L = v->impl.sys->L;
  blas_sadd(*v->impl.sys->ctx, v->impl.d + L.own_off, 1.0, a, x->impl.d + L.own_off, L.own_len);
  GDM_CATCH
}

int gdm_vector_sadd(gdm_vector_t v, double s, double a, gdm_vector_t x)
{
  GDM_TRY
  GDM_ARG(v);
  GDM_ARG(x);
  GDM_SAME_SYS(v, x);
  const Layout &L = v->impl.sys->L;
  blas_sadd(*v->impl.sys->ctx, v->impl.d + L.own_off, s, a, x->impl.d + L.own_off, L.own_len);
  GDM_CATCH
}

int gdm_vector_scale_by(gdm_vector_t v, gdm_vector_t d)
{
  GDM_TRY
  GDM_ARG(v);
  GDM_ARG(d);
  GDM_SAME_SYS(v, d);
  const Layout &L = v->impl.sys->L;
  blas_mul(*v->impl.sys->ctx, v->impl.d + L.own_off, d->impl.d + L.own_off, L.own_len);
  GDM_CATCH
}

int gdm_vector_dot(gdm_vector_t a, gdm_vector_t b, double *result)
{
  GDM_TRY
  GDM_ARG(a);
  GDM_ARG(b);
  GDM_ARG(result);
  GDM_SAME_SYS(a, b);
  const Layout &L   = a->impl.sys->L;
  Context      &ctx = *a->impl.sys->ctx;
  blas_dot(ctx, a->impl.d + L.own_off, b->impl.d + L.own_off, L.own_len, SUM_DOT);
  *result = read_sum(ctx, SUM_DOT, true);
  GDM_CATCH
}

int gdm_vector_l2_norm(gdm_vector_t a, double *result)
{
  double    d  = 0;
  const int rc = gdm_vector_dot(a, a, &d);
  if (rc == GDM_OK && result)
    *result = std::sqrt(d);
  return rc;
}

int gdm_vector_linfty_norm(gdm_vector_t a, double *result)
{
  GDM_TRY
  GDM_ARG(a);
  GDM_ARG(result);
  const Layout &L   = a->impl.sys->L;
  Context      &ctx = *a->impl.sys->ctx;
  blas_absmax(ctx, a->impl.d + L.own_off, L.own_len, SUM_DOT);
  *result = read_sum(ctx, SUM_DOT, true, true);
  GDM_CATCH
}

int gdm_vector_update_ghost_values(gdm_vector_t v)
{
  GDM_TRY
  GDM_ARG(v);
  vector_update_ghosts(v->impl);
  GDM_CATCH
}

// ---------------------------------------------------------------- operators
int gdm_operator_create(gdm_system_t sys, gdm_constraints_t c, const gdm_operator_desc *desc, gdm_operator_t *out)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(desc);
  GDM_ARG(out);
  GDM_REQUIRE(c == nullptr || c->impl.sys == &sys->impl, GDM_ERR_INVALID, "constraints belong to another system");
  GDM_REQUIRE(c == nullptr || c->impl.closed, GDM_ERR_INVALID, "constraints must be closed (AffineConstraints::close)");
  GDM_REQUIRE(sys->impl.ctx->device >= 0, GDM_ERR_CUDA, "description-only context: no CUDA device (no CPU fallback)");
  std::unique_ptr<gdm_operator_s> o(new gdm_operator_s);
  Operator                       &op = o->impl;
  op.sys                             = &sys->impl;
  op.desc                            = *desc;
  for (int d = 0; d < 3; ++d)
    {
      op.periodic[d]     = c ? c->impl.periodic[d] : false;
      op.dirichlet[d][0] = c ? c->impl.dirichlet[d][0] : false;
      op.dirichlet[d][1] = c ? c->impl.dirichlet[d][1] : false;
    }
  op.has_B      = desc->kind != GDM_OP_MASS;
  op.b_symmetry = (desc->kind == GDM_OP_STIFFNESS) ? +1 : -1;
  GDM_REQUIRE(!(desc->constrained_diagonal == GDM_DIAG_ASSEMBLED &&
                (desc->kind == GDM_OP_ADVECTION || desc->kind == GDM_OP_ADVECTION_T)),
              GDM_ERR_NOT_IMPLEMENTED, "advection operators use residual (vector assembly) semantics: GDM_DIAG_ZERO");
  GDM_CUDA_CHECK(cudaSetDevice(sys->impl.ctx->device));
  build_tables(op);
  op.kernel_used = GDM_KERNEL_GENERIC;
  if (desc->kernel != GDM_KERNEL_GENERIC && fused_supported(op))
    {
      fused_plan_create(op);
      op.kernel_used = GDM_KERNEL_FUSED;
    }
  GDM_REQUIRE(!(desc->kernel == GDM_KERNEL_FUSED && op.kernel_used != GDM_KERNEL_FUSED), GDM_ERR_NOT_IMPLEMENTED,
              "the fused kernel does not cover this configuration");
  GDM_REQUIRE(!(op.periodic[sys->impl.L.pdim] && sys->impl.L.n_ranks > 1 && op.kernel_used != GDM_KERNEL_FUSED),
              GDM_ERR_NOT_IMPLEMENTED,
              "periodicity along the partitioned direction with more than one rank needs the fused kernel (dim 3, scalar)");
  *out = o.release();
  GDM_CATCH
}

// The assembled operator in the triplet format of the reference's eigenvalue tool (write_matrix_to_file,
// applications/wave/wave-ev.cc:93-127): the entries of the SparseMatrix in its iteration order -- row by row, the
// diagonal entry first (deal.II stores it first in the rows of a square pattern), then ascending columns -- as text
// "row column value" lines or as binary (uint32, uint32, double) records.  The pattern is the reference's
// create_sparsity_pattern (system.h:586-599) with constrained rows and columns kept; values come from the 1D tables.
// Host only (works on a description-only context); one rank; no periodic directions.
int gdm_system_write_matrix(gdm_system_t sys, gdm_constraints_t c, const gdm_operator_desc *desc, const char *file_name,
                            int write_binary_file, uint64_t *n_entries)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(desc);
  GDM_ARG(file_name);
  GDM_REQUIRE(c == nullptr || (c->impl.sys == &sys->impl && c->impl.closed), GDM_ERR_INVALID, "constraints: other system or not closed");
  const Layout &L = sys->impl.L;
  GDM_REQUIRE(L.n_ranks == 1, GDM_ERR_NOT_IMPLEMENTED, "matrix dump: one rank");
  Operator op;
  op.sys  = &sys->impl;
  op.desc = *desc;
  for (int d = 0; d < 3; ++d)
    {
      op.periodic[d]     = c ? c->impl.periodic[d] : false;
      op.dirichlet[d][0] = c ? c->impl.dirichlet[d][0] : false;
      op.dirichlet[d][1] = c ? c->impl.dirichlet[d][1] : false;
      GDM_REQUIRE(!op.periodic[d], GDM_ERR_NOT_IMPLEMENTED, "matrix dump: periodic directions");
    }
  op.has_B = desc->kind != GDM_OP_MASS;
  build_tables(op, false);
  const int p = L.p, W = 2 * p + 1;
  FILE     *f = fopen(file_name, write_binary_file ? "wb" : "w");
  GDM_REQUIRE(f != nullptr, GDM_ERR_INVALID, std::string("cannot open ") + file_name);
  auto tap = [&](const std::vector<double> &t, int d, int i, int j) -> double {
    const int k = j - i + p;
    return (k < 0 || k >= W) ? 0.0 : t[(size_t)i * W + k];
  };
  auto split = [&](uint64_t dof, int (&idx)[3], int &comp) {
    const uint64_t node = dof / L.nc;
    comp                = (int)(dof % L.nc);
    idx[0]              = (int)(node % L.nn[0]);
    idx[1]              = (int)((node / L.nn[0]) % L.nn[1]);
    idx[2]              = (int)(node / ((uint64_t)L.nn[0] * L.nn[1]));
  };
  std::vector<uint64_t> cols(4096);
  uint64_t              total = 0;
  for (uint64_t row = 0; row < (uint64_t)L.n_dofs_global; ++row)
    {
      uint64_t n = 0;
      int      rc = gdm_system_sparsity_row(sys, 0, row, nullptr, 0, &n);
      if (n > cols.size())
        cols.resize(n);
      rc = gdm_system_sparsity_row(sys, 0, row, cols.data(), cols.size(), &n);
      if (rc != GDM_OK)
        {
          fclose(f);
          return rc;
        }
      int ri[3], rcomp;
      split(row, ri, rcomp);
      bool constrained = false;
      for (int d = 0; d < L.dim; ++d)
        constrained |= (op.dirichlet[d][0] && ri[d] == 0) || (op.dirichlet[d][1] && ri[d] == L.N[d]);
      auto value = [&](uint64_t col) -> double {
        int ci[3], ccomp;
        split(col, ci, ccomp);
        if (constrained)
          {
            if (col != row || desc->constrained_diagonal != GDM_DIAG_ASSEMBLED)
              return 0.0;
            // sum over cells of |cell_matrix(i, i)| = |scale * diagonal of the unconstrained operator|
            double v = 0.0;
            if (!op.has_B)
              {
                v = 1.0;
                for (int d = 0; d < L.dim; ++d)
                  v *= op.hdiagA[d][ri[d]];
              }
            else
              for (int d = 0; d < L.dim; ++d)
                {
                  double t = op.hdiagB[d][ri[d]];
                  for (int e = 0; e < L.dim; ++e)
                    if (e != d)
                      t *= op.hdiagA[e][ri[e]];
                  v += t;
                }
            return std::fabs(desc->scale * v);
          }
        if (ccomp != rcomp)
          return 0.0;
        double v = 0.0;
        if (!op.has_B)
          {
            v = 1.0;
            for (int d = 0; d < L.dim; ++d)
              v *= tap(op.hA[d], d, ri[d], ci[d]);
          }
        else
          for (int d = 0; d < L.dim; ++d)
            {
              double t = tap(op.hB[d], d, ri[d], ci[d]);
              for (int e = 0; e < L.dim; ++e)
                if (e != d)
                  t *= tap(op.hA[e], e, ri[e], ci[e]);
              v += t;
            }
        return desc->scale * v;
      };
      auto emit = [&](uint64_t col) {
        const double v = value(col);
        if (write_binary_file)
          {
            const unsigned int r32 = (unsigned int)row, c32 = (unsigned int)col;
            fwrite(&r32, sizeof(unsigned int), 1, f);
            fwrite(&c32, sizeof(unsigned int), 1, f);
            fwrite(&v, sizeof(double), 1, f);
          }
        else
          fprintf(f, "%llu %llu %g\n", (unsigned long long)row, (unsigned long long)col, v);
        ++total;
      };
      emit(row); // the diagonal entry is stored first
      for (uint64_t k = 0; k < n; ++k)
        if (cols[k] != row)
          emit(cols[k]);
    }
  fclose(f);
  if (n_entries)
    *n_entries = total;
  GDM_CATCH
}

// VTU file (ASCII UnstructuredGrid) of a nodal field for visual checks: the stand-in for GDM::DataOut (include/gdm/data_out.h,
// used at prototypes/advection_01_gdm.cc:248-255).  GDM DoFs are nodal values, so the grid cells (lines, quadrilaterals,
// hexahedra between neighbouring nodes) carry the DoFs as point data, one array per component.  Host only: `values` holds
// all DoFs of the system in the global numbering (gdm_vector_download on one rank).
int gdm_system_write_vtu(gdm_system_t sys, const double *values, const char *label, const char *file_name)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(values);
  GDM_ARG(label);
  GDM_ARG(file_name);
  const Layout &L = sys->impl.L;
  FILE         *f = fopen(file_name, "w");
  GDM_REQUIRE(f != nullptr, GDM_ERR_INVALID, std::string("cannot open ") + file_name);
  const int64_t n_nodes = (int64_t)L.nn[0] * L.nn[1] * L.nn[2];
  int64_t       n_cells = 1;
  for (int d = 0; d < L.dim; ++d)
    n_cells *= L.N[d];
  const int npc = 1 << L.dim; // points per cell
  fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n");
  fprintf(f, "<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n", (long long)n_nodes, (long long)n_cells);
  fprintf(f, "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n");
  for (int k = 0; k < L.nn[2]; ++k)
    for (int j = 0; j < L.nn[1]; ++j)
      for (int i = 0; i < L.nn[0]; ++i)
        fprintf(f, "%.17g %.17g %.17g\n", L.lo[0] + i * L.h[0], L.dim > 1 ? L.lo[1] + j * L.h[1] : 0.0,
                L.dim > 2 ? L.lo[2] + k * L.h[2] : 0.0);
  fprintf(f, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n");
  auto node = [&](int i, int j, int k) { return (long long)i + (long long)L.nn[0] * ((long long)j + (long long)L.nn[1] * k); };
  for (int k = 0; k < std::max(1, L.N[2]); ++k)
    for (int j = 0; j < std::max(1, L.N[1]); ++j)
      for (int i = 0; i < L.N[0]; ++i)
        {
          if (L.dim == 1)
            fprintf(f, "%lld %lld\n", node(i, 0, 0), node(i + 1, 0, 0));
          else if (L.dim == 2) // VTK_QUAD: counter-clockwise
            fprintf(f, "%lld %lld %lld %lld\n", node(i, j, 0), node(i + 1, j, 0), node(i + 1, j + 1, 0), node(i, j + 1, 0));
          else // VTK_HEXAHEDRON
            fprintf(f, "%lld %lld %lld %lld %lld %lld %lld %lld\n", node(i, j, k), node(i + 1, j, k), node(i + 1, j + 1, k),
                    node(i, j + 1, k), node(i, j, k + 1), node(i + 1, j, k + 1), node(i + 1, j + 1, k + 1), node(i, j + 1, k + 1));
        }
  fprintf(f, "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n");
  for (int64_t c = 1; c <= n_cells; ++c)
    fprintf(f, "%lld\n", (long long)(c * npc));
  fprintf(f, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
  const int vtk_type = (L.dim == 1) ? 3 : (L.dim == 2 ? 9 : 12);
  for (int64_t c = 0; c < n_cells; ++c)
    fprintf(f, "%d\n", vtk_type);
  fprintf(f, "</DataArray>\n</Cells>\n<PointData>\n");
  for (int c = 0; c < L.nc; ++c)
    {
      if (L.nc == 1)
        fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" format=\"ascii\">\n", label);
      else
        fprintf(f, "<DataArray type=\"Float64\" Name=\"%s_%d\" format=\"ascii\">\n", label, c);
      for (int64_t nd = 0; nd < n_nodes; ++nd)
        fprintf(f, "%.17g\n", values[nd * L.nc + c]);
      fprintf(f, "</DataArray>\n");
    }
  fprintf(f, "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
  fclose(f);
  GDM_CATCH
}

int gdm_operator_destroy(gdm_operator_t op)
{
  if (op && op->impl.transposed)
    {
      delete static_cast<gdm_operator_s *>(op->impl.transposed);
      op->impl.transposed = nullptr;
    }
  if (op && op->impl.unconstrained)
    {
      delete static_cast<gdm_operator_s *>(op->impl.unconstrained);
      op->impl.unconstrained = nullptr;
    }
  delete op;
  return GDM_OK;
}

// SparseMatrix::Tvmult: mass and stiffness are symmetric (Tvmult == vmult); the advection operators are not: the
// transposed operator (kind ADVECTION <-> ADVECTION_T, same velocity, scale and constraints) is created on first use.
int gdm_operator_tvmult(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(dst);
  GDM_ARG(src);
  Operator &o = op->impl;
  if (o.desc.kind == GDM_OP_MASS || o.desc.kind == GDM_OP_STIFFNESS)
    {
      operator_apply(o, dst->impl, src->impl, false);
      return GDM_OK;
    }
  GDM_REQUIRE(!o.csr, GDM_ERR_NOT_IMPLEMENTED, "Tvmult of an operator with CSR overlay rows");
  if (!o.transposed)
    {
      gdm_constraints_s cs;
      cs.impl.sys    = o.sys;
      cs.impl.closed = true;
      for (int d = 0; d < 3; ++d)
        {
          cs.impl.periodic[d]     = o.periodic[d];
          cs.impl.dirichlet[d][0] = o.dirichlet[d][0];
          cs.impl.dirichlet[d][1] = o.dirichlet[d][1];
        }
      gdm_operator_desc desc = o.desc;
      desc.kind              = (o.desc.kind == GDM_OP_ADVECTION) ? GDM_OP_ADVECTION_T : GDM_OP_ADVECTION;
      gdm_system_s *sys_h    = reinterpret_cast<gdm_system_s *>(o.sys); // gdm_system_s has the System as its only member
      gdm_operator_t t       = nullptr;
      const int      rc      = gdm_operator_create(sys_h, &cs, &desc, &t);
      if (rc != GDM_OK)
        return rc;
      o.transposed = t;
    }
  operator_apply(static_cast<gdm_operator_s *>(o.transposed)->impl, dst->impl, src->impl, false);
  GDM_CATCH
}

int gdm_operator_attach_csr(gdm_operator_t op, uint64_t n_rows, const uint64_t *row_ids, const uint64_t *rowptr,
                            const uint64_t *col, const double *val)
{
  GDM_TRY
  GDM_ARG(op);
  const Layout &L = op->impl.sys->L;
  std::unique_ptr<CsrOverlay> csr(new CsrOverlay);
  csr->n_rows = (int64_t)n_rows;
  if (n_rows > 0)
    {
      GDM_ARG(row_ids);
      GDM_ARG(rowptr);
      GDM_REQUIRE(rowptr[0] == 0, GDM_ERR_INVALID, "CSR rowptr must start at 0");
      for (uint64_t i = 0; i < n_rows; ++i)
        GDM_REQUIRE(rowptr[i + 1] >= rowptr[i], GDM_ERR_INVALID, "CSR rowptr must be non-decreasing");
      csr->nnz = (int64_t)rowptr[n_rows];
      if (csr->nnz > 0)
        {
          GDM_ARG(col);
          GDM_ARG(val);
        }
      std::vector<int64_t> row_off(n_rows), rp(n_rows + 1);
      std::vector<int32_t> col_rel(csr->nnz);
      std::vector<double>  diag(n_rows, 0.0);
      uint64_t own_b, own_e;
      gdm_system_locally_owned_range(reinterpret_cast<gdm_system_t>(op->impl.sys), &own_b, &own_e);
      for (uint64_t i = 0; i < n_rows; ++i)
        {
          GDM_REQUIRE(row_ids[i] >= own_b && row_ids[i] < own_e, GDM_ERR_INVALID, "CSR row is not locally owned");
          row_off[i] = storage_offset(L, row_ids[i]);
          rp[i]      = (int64_t)rowptr[i];
          for (uint64_t j = rowptr[i]; j < rowptr[i + 1]; ++j)
            {
              // columns must be stored on this rank (owned or ghost planes); the offset relative to the row fits 32 bits
              const int64_t rel = storage_offset(L, col[j]) - row_off[i];
              GDM_REQUIRE(rel >= INT32_MIN && rel <= INT32_MAX, GDM_ERR_INVALID, "CSR column too far from its row");
              col_rel[j] = (int32_t)rel;
              if (col[j] == row_ids[i])
                diag[i] += val[j];
            }
        }
      rp[n_rows] = csr->nnz;
      auto up = [&](auto *&dptr, const void *h, size_t bytes) {
        GDM_CUDA_CHECK(cudaMalloc(&dptr, std::max<size_t>(bytes, 8)));
        GDM_CUDA_CHECK(cudaMemcpy(dptr, h, bytes, cudaMemcpyHostToDevice));
      };
      up(csr->d_row_off, row_off.data(), row_off.size() * 8);
      up(csr->d_rowptr, rp.data(), rp.size() * 8);
      up(csr->d_col_rel, col_rel.data(), col_rel.size() * 4);
      up(csr->d_val, val, (size_t)csr->nnz * 8);
      up(csr->d_diag, diag.data(), diag.size() * 8);
    }
  op->impl.csr = std::move(csr);
  GDM_CATCH
}

int gdm_operator_vmult(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(dst);
  GDM_ARG(src);
  operator_apply(op->impl, dst->impl, src->impl, false);
  GDM_CATCH
}

int gdm_operator_vmult_dot(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src, double *src_dot_dst)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(dst);
  GDM_ARG(src);
  GDM_ARG(src_dot_dst);
  Operator &o   = op->impl;
  Context  &ctx = *o.sys->ctx;
  GDM_REQUIRE(dst->impl.sys == o.sys && src->impl.sys == o.sys, GDM_ERR_INVALID, "vector/operator system mismatch");
  GDM_REQUIRE(dst->impl.d != src->impl.d, GDM_ERR_INVALID, "vmult: dst and src must not alias");
  if (o.kernel_used == GDM_KERNEL_FUSED && !o.csr && fused_supports_dot(o))
    fused_apply(o, dst->impl.d, src->impl.d, false, true, SUM_DOT); // dot product in the store epilogue
  else
    {
      operator_apply(o, dst->impl, src->impl, false);
      blas_dot(ctx, src->impl.d + o.sys->L.own_off, dst->impl.d + o.sys->L.own_off, o.sys->L.own_len, SUM_DOT);
    }
  *src_dot_dst = read_sum(ctx, SUM_DOT, true);
  GDM_CATCH
}

int gdm_operator_vmult_add(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(dst);
  GDM_ARG(src);
  operator_apply(op->impl, dst->impl, src->impl, true);
  GDM_CATCH
}

int gdm_operator_vmult_host(gdm_operator_t op, double *dst_host, const double *src_host)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(dst_host);
  GDM_ARG(src_host);
  System  &sys = *op->impl.sys;
  Context &ctx = *sys.ctx;
  Operator &o = op->impl;
  if (!o.host_src)
    {
      GDM_CUDA_CHECK(cudaMalloc(&o.host_src, (size_t)sys.L.size * sizeof(double)));
      GDM_CUDA_CHECK(cudaMalloc(&o.host_dst, (size_t)sys.L.size * sizeof(double)));
      GDM_CUDA_CHECK(cudaMemsetAsync(o.host_src, 0, (size_t)sys.L.size * sizeof(double), ctx.stream));
      GDM_CUDA_CHECK(cudaMemsetAsync(o.host_dst, 0, (size_t)sys.L.size * sizeof(double), ctx.stream));
    }
  Vector vs, vd;
  vs.sys = vd.sys = &sys;
  vs.d    = o.host_src;
  vd.d    = o.host_dst;
  vs.owns = vd.owns = false;
  const Layout &L = sys.L;
  const int own_lo = L.own0 - L.loc0, own_hi = L.own1 - L.loc0; // owned planes (local indices); the host buffers hold these
  if (o.kernel_used == GDM_KERNEL_FUSED && !o.csr && own_hi - own_lo >= 16 * L.p &&
      !(o.periodic[0] || o.periodic[1] || o.periodic[2]))
    {
      // Pipelined over z chunks: H2D of chunk c+1, apply of the planes whose inputs have arrived and D2H
      // of finished planes overlap on three streams (PCIe is full duplex; the apply hides behind it).
      // Several ranks: the P output planes next to a neighbouring slab need its ghost planes; they are applied after the
      // last chunk has arrived and the ghost planes have been exchanged (one short tail instead of a serial path).
      if (!ctx.h2d_stream)
        {
          GDM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx.h2d_stream, cudaStreamNonBlocking));
          GDM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx.d2h_stream, cudaStreamNonBlocking));
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 32; ++j)
              GDM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx.ev_pipe[i][j], cudaEventDisableTiming));
        }
      const int    nz = L.ln[2], P = L.p, n_own = own_hi - own_lo;
      const bool   nb_lo = own_lo > 0, nb_hi = own_hi < nz; // neighbouring slabs (ghost planes below / above)
      // chunks of about 8p planes (4p planes measured no faster: 4.46 vs 4.58 GDoF/s; the PCIe rate of the box with both
      // directions busy bounds the path at 6.2 GDoF/s, tools/pcie_probe.py); event slots 28 and 29: slab-face windows
      int          n_chunks = std::min(28, std::max(2, n_own / (8 * P)));
      if (const char *env = std::getenv("GDM_HOST_CHUNKS")) // diagnostic
        n_chunks = std::min(28, std::max(2, atoi(env)));
      const int    cz = (n_own + n_chunks - 1) / n_chunks;
      const size_t plane_host = (size_t)L.ln[0] * L.nc * L.ln[1];
      if (!o.stage_src)
        {
          GDM_CUDA_CHECK(cudaMalloc(&o.stage_src, plane_host * nz * sizeof(double)));
          GDM_CUDA_CHECK(cudaMalloc(&o.stage_dst, plane_host * nz * sizeof(double)));
        }
      // the staging buffers were produced on ctx.stream (memset): order the side streams behind it
      GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_a, ctx.stream));
      GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.h2d_stream, ctx.ev_a, 0));
      GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.d2h_stream, ctx.ev_a, 0));
      // output planes [z0, z1) (local indices): apply, repack, copy to the host behind the apply
      auto window = [&](int z0, int z1, int ev) {
        if (z1 <= z0)
          return;
        fused_apply_window(o, o.host_dst, o.host_src, z0, z1);
        launch_repack(ctx, L, o.host_dst, o.stage_dst, z0, z1, false);
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_pipe[1][ev], ctx.stream));
        GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.d2h_stream, ctx.ev_pipe[1][ev], 0));
        GDM_CUDA_CHECK(cudaMemcpyAsync(dst_host + (size_t)(z0 - own_lo) * plane_host, o.stage_dst + (size_t)z0 * plane_host,
                                       (size_t)(z1 - z0) * plane_host * sizeof(double), cudaMemcpyDeviceToHost, ctx.d2h_stream));
      };
      int done = nb_lo ? own_lo + P : own_lo; // output planes [.., done) have been launched
      for (int c = 0; c < n_chunks; ++c)
        {
          const int c0 = own_lo + c * cz, c1 = std::min(own_hi, c0 + cz);
          if (c1 <= c0)
            break;
          // contiguous 1D copy (row-wise 2D copies of 2 KB rows reach only a fraction of the PCIe rate)
          GDM_CUDA_CHECK(cudaMemcpyAsync(o.stage_src + (size_t)c0 * plane_host, src_host + (size_t)(c0 - own_lo) * plane_host,
                                         (size_t)(c1 - c0) * plane_host * sizeof(double), cudaMemcpyHostToDevice, ctx.h2d_stream));
          GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_pipe[0][c], ctx.h2d_stream));
          GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_pipe[0][c], 0));
          launch_repack(ctx, L, o.host_src, o.stage_src, c0, c1, true);
          // outputs whose inputs (up to +P planes) are on the device
          const int upto = (c1 == own_hi) ? (nb_hi ? own_hi - P : own_hi) : c1 - P;
          if (upto > done)
            {
              window(done, upto, c);
              done = upto;
            }
        }
      if (nb_lo || nb_hi)
        {
          comm_halo_exchange(ctx, L, o.host_src);
          if (nb_lo)
            window(own_lo, own_lo + P, 28);
          if (nb_hi)
            window(own_hi - P, own_hi, 29);
        }
      GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.d2h_stream));
      GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    }
  else
    {
      vector_transfer(vs, const_cast<double *>(src_host), true);
      operator_apply(o, vd, vs, false);
      vector_transfer(vd, dst_host, false);
      GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    }
  GDM_CATCH
}

int gdm_operator_mass_inverse(gdm_operator_t op, gdm_vector_t dst, gdm_vector_t src)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(dst);
  GDM_ARG(src);
  GDM_REQUIRE(dst->impl.sys == op->impl.sys && src->impl.sys == op->impl.sys, GDM_ERR_INVALID, "vector/operator system mismatch");
  GDM_CUDA_CHECK(cudaSetDevice(op->impl.sys->ctx->device));
  massinv_apply(op->impl, dst->impl.d, src->impl.d);
  GDM_CATCH
}

int gdm_operator_diagonal(gdm_operator_t op, gdm_vector_t diag)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(diag);
  GDM_REQUIRE(diag->impl.sys == op->impl.sys, GDM_ERR_INVALID, "vector/operator system mismatch");
  System  &sys = *op->impl.sys;
  Context &ctx = *sys.ctx;
  launch_diagonal(ctx, sys.L, op->impl, diag->impl.d);
  if (op->impl.desc.constrained_diagonal == GDM_DIAG_ASSEMBLED)
    {
      ctx.ensure_scratch((size_t)sys.L.size);
      blas_set(ctx, ctx.scratch[0], sys.L.size, 1.0);
      launch_constrained_rows(ctx, sys.L, op->impl, diag->impl.d, ctx.scratch[0], true);
    }
  if (op->impl.csr) // irregular rows replace the tensor-product rows: so do their diagonal entries
    launch_csr_diagonal(ctx, *op->impl.csr, diag->impl.d);
  GDM_CATCH
}

int gdm_operator_lumped_mass_inverse(gdm_operator_t op, gdm_vector_t inv)
{
  GDM_TRY
  GDM_ARG(op);
  GDM_ARG(inv);
  GDM_REQUIRE(op->impl.desc.kind == GDM_OP_MASS, GDM_ERR_INVALID, "lumped mass needs a MASS operator");
  GDM_REQUIRE(inv->impl.sys == op->impl.sys, GDM_ERR_INVALID, "vector/operator system mismatch");
  // cell_vector(i) = sum_j (phi_i, phi_j) over ALL local j, then distribute_local_to_global on a
  // vector (matrix_creator.h:99-112): periodic slave rows fold into their masters, Dirichlet rows are
  // dropped, but the COLUMNS of constrained DoFs still count.  => apply a mass operator that keeps the
  // periodic folding and has no Dirichlet masking to the ones vector, then clear the constrained rows.
  // Constrained entries return 0 here where the reference produces 1/0 = inf (never used downstream).
  System  &sys = *op->impl.sys;
  Context &ctx = *sys.ctx;
  Operator tmp;
  tmp.sys                       = &sys;
  tmp.desc                      = op->impl.desc;
  tmp.desc.constrained_diagonal = GDM_DIAG_ZERO;
  tmp.desc.scale                = 1.0;
  for (int d = 0; d < 3; ++d)
    {
      tmp.periodic[d]     = op->impl.periodic[d];
      tmp.dirichlet[d][0] = tmp.dirichlet[d][1] = false;
    }
  build_tables(tmp);
  ctx.ensure_scratch((size_t)sys.L.size);
  std::unique_ptr<gdm_vector_s> ones(new gdm_vector_s);
  ones->impl.sys = &sys;
  GDM_CUDA_CHECK(cudaMalloc(&ones->impl.d, (size_t)sys.L.size * sizeof(double)));
  GDM_CUDA_CHECK(cudaMemsetAsync(ones->impl.d, 0, (size_t)sys.L.size * sizeof(double), ctx.stream));
  // all stored planes (ghosts included) hold ones
  {
    Layout all = sys.L;
    all.own0   = all.loc0;
    all.own1   = all.loc1;
    all.own_off = 0;
    blas_set_strided(ctx, all, ones->impl.d, 1.0);
  }
  generic_apply(tmp, inv->impl.d, ones->impl.d, false);
  launch_set_constrained(ctx, sys.L, op->impl.dirichlet, op->impl.periodic, inv->impl.d, 0.0);
  blas_invert(ctx, inv->impl.d + sys.L.own_off, sys.L.own_len);
  GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
  GDM_CATCH
}

int gdm_operator_kernel_used(gdm_operator_t op)
{
  return op ? op->impl.kernel_used : 0;
}

uint64_t gdm_operator_m(gdm_operator_t op)
{
  return op ? (uint64_t)op->impl.sys->L.n_dofs_global : 0;
}

// ------------------------------------------------------------------- solver
int gdm_solver_cg(gdm_operator_t A, gdm_vector_t x, gdm_vector_t b, int precondition, gdm_vector_t pvec,
                  gdm_reduction_control *control)
{
  int status = GDM_OK;
  GDM_TRY
  GDM_ARG(A);
  GDM_ARG(x);
  GDM_ARG(b);
  GDM_ARG(control);
  GDM_REQUIRE(x->impl.sys == A->impl.sys && b->impl.sys == A->impl.sys, GDM_ERR_INVALID, "vector/operator system mismatch");
  GDM_REQUIRE(precondition != GDM_PRECONDITION_DIAGONAL || pvec != nullptr, GDM_ERR_INVALID, "DIAGONAL needs a vector");
  status = cg_solve(A->impl, x->impl, b->impl, precondition, pvec ? &pvec->impl : nullptr, *control);
  if (status == GDM_ERR_NO_CONVERGENCE)
    gdm::set_last_error("Iterative method reported convergence failure in step " + std::to_string(control->last_step) +
                        ". The residual in the last step was " + std::to_string(control->last_value) + ".");
  if (status != GDM_OK)
    return status;
  GDM_CATCH
}

// ------------------------------------------------------------- vector tools
int gdm_interpolate(gdm_system_t sys, gdm_function_fn f, void *user, gdm_vector_t v)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(f);
  GDM_ARG(v);
  GDM_REQUIRE(v->impl.sys == &sys->impl, GDM_ERR_INVALID, "vector/system mismatch");
  const Layout       &L = sys->impl.L;
  std::vector<double> host((size_t)L.n_owned);
  // owned nodes in lexicographic order
  int lo[3] = {0, 0, 0}, hi[3] = {L.nn[0], L.nn[1], L.nn[2]};
  lo[L.pdim] = L.own0;
  hi[L.pdim] = L.own1;
  size_t o   = 0;
  for (int k = lo[2]; k < hi[2]; ++k)
    for (int j = lo[1]; j < hi[1]; ++j)
      for (int i = lo[0]; i < hi[0]; ++i)
        {
          const double pt[3] = {L.lo[0] + i * L.h[0], L.lo[1] + j * L.h[1], L.lo[2] + k * L.h[2]};
          for (int c = 0; c < L.nc; ++c)
            host[o++] = f(pt, c, user);
        }
  vector_transfer(v->impl, host.data(), true);
  GDM_CUDA_CHECK(cudaStreamSynchronize(sys->impl.ctx->stream));
  GDM_CATCH
}

// System::interpolate_boundary_values (include/gdm/system.h:511-547): every boundary node (boundary id 0: the uncoloured
// hyper rectangle has one id for all faces) is constrained to the value of the function there.  DoFs that are already
// constrained keep their constraint (system.h:533).
int gdm_constraints_interpolate_boundary_values(gdm_constraints_t c, int boundary_id, gdm_function_fn f, void *user)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(f);
  Constraints  &cs = c->impl;
  const Layout &L  = cs.sys->L;
  GDM_REQUIRE(!cs.closed, GDM_ERR_INVALID, "constraints already closed");
  GDM_REQUIRE(boundary_id == 0, GDM_ERR_INVALID, "the hyper rectangle has boundary id 0 on every face");
  GDM_REQUIRE(cs.sys->ctx->device >= 0, GDM_ERR_CUDA, "description-only context: no CUDA device (no CPU fallback)");
  for (int d = 0; d < L.dim; ++d)
    GDM_REQUIRE(!cs.periodic[d], GDM_ERR_NOT_IMPLEMENTED, "boundary values together with periodicity");
  bool was[3][2];
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s)
      was[d][s] = cs.dirichlet[d][s];
  std::vector<double> host((size_t)L.n_owned, 0.0);
  int lo[3] = {0, 0, 0}, hi[3] = {L.nn[0], L.nn[1], L.nn[2]};
  lo[L.pdim] = L.own0;
  hi[L.pdim] = L.own1;
  size_t o   = 0;
  for (int k = lo[2]; k < hi[2]; ++k)
    for (int j = lo[1]; j < hi[1]; ++j)
      for (int i = lo[0]; i < hi[0]; ++i)
        {
          const int idx[3] = {i, j, k};
          bool      on_new = false, on_old = false;
          for (int d = 0; d < L.dim; ++d)
            for (int s = 0; s < 2; ++s)
              if (idx[d] == (s == 0 ? 0 : L.N[d]))
                (was[d][s] ? on_old : on_new) = true;
          const double pt[3] = {L.lo[0] + i * L.h[0], L.lo[1] + j * L.h[1], L.lo[2] + k * L.h[2]};
          for (int cc = 0; cc < L.nc; ++cc, ++o)
            if (on_new && !on_old)
              host[o] = f(pt, cc, user);
        }
  if (!cs.d_inhom)
    {
      GDM_CUDA_CHECK(cudaMalloc(&cs.d_inhom, (size_t)L.size * sizeof(double)));
      GDM_CUDA_CHECK(cudaMemset(cs.d_inhom, 0, (size_t)L.size * sizeof(double)));
    }
  else
    {
      // keep the values of earlier calls on the faces they constrained
      Vector old;
      old.sys  = cs.sys;
      old.d    = cs.d_inhom;
      old.owns = false;
      std::vector<double> prev((size_t)L.n_owned);
      vector_transfer(old, prev.data(), false);
      GDM_CUDA_CHECK(cudaStreamSynchronize(cs.sys->ctx->stream));
      for (size_t q = 0; q < host.size(); ++q)
        if (prev[q] != 0.0)
          host[q] = prev[q];
    }
  Vector tmp;
  tmp.sys  = cs.sys;
  tmp.d    = cs.d_inhom;
  tmp.owns = false;
  vector_transfer(tmp, host.data(), true);
  GDM_CUDA_CHECK(cudaStreamSynchronize(cs.sys->ctx->stream));
  for (int d = 0; d < L.dim; ++d)
    cs.dirichlet[d][0] = cs.dirichlet[d][1] = true;
  GDM_CATCH
}

// The right-hand side part of AffineConstraints::distribute_local_to_global(cell_matrix, cell_rhs, dofs, A, rhs) with
// inhomogeneous constraints (tests/poisson_02_gdm.cc:201): free rows  b_i -= sum_j A_ij g_j  over the constrained columns j,
// constrained rows  b_j = (diagonal the matrix keeps on row j) g_j,  so that the solve returns u_j = g_j.
// `op` is the constrained operator; the unconstrained twin that provides A_ij g_j is created on first use.
int gdm_constraints_condense_rhs(gdm_constraints_t c, gdm_operator_t op, gdm_vector_t rhs)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(op);
  GDM_ARG(rhs);
  Constraints &cs = c->impl;
  Operator    &o  = op->impl;
  GDM_REQUIRE(cs.closed, GDM_ERR_INVALID, "constraints must be closed");
  GDM_REQUIRE(o.sys == cs.sys && rhs->impl.sys == cs.sys, GDM_ERR_INVALID, "constraints/operator/vector system mismatch");
  if (!cs.d_inhom)
    return GDM_OK; // homogeneous: nothing to lift
  GDM_REQUIRE(!o.csr, GDM_ERR_NOT_IMPLEMENTED, "boundary values with CSR overlay rows");
  Context      &ctx = *cs.sys->ctx;
  const Layout &L   = cs.sys->L;
  if (!o.unconstrained)
    {
      gdm_operator_desc desc   = o.desc;
      desc.constrained_diagonal = GDM_DIAG_ZERO;
      gdm_operator_t t          = nullptr;
      const int      rc         = gdm_operator_create(reinterpret_cast<gdm_system_s *>(o.sys), nullptr, &desc, &t);
      if (rc != GDM_OK)
        return rc;
      o.unconstrained = t;
    }
  Vector g;
  g.sys  = cs.sys;
  g.d    = cs.d_inhom;
  g.owns = false;
  Vector t;
  t.sys  = cs.sys;
  t.d    = ctx.acquire((size_t)L.size);
  t.owns = false;
  operator_apply(static_cast<gdm_operator_s *>(o.unconstrained)->impl, t, g, false);
  blas_sadd(ctx, rhs->impl.d + L.own_off, 1.0, -1.0, t.d + L.own_off, L.own_len);
  ctx.release(t.d);
  // constrained rows: the diagonal deal.II keeps there times the boundary value
  launch_constrained_rows(ctx, L, o, rhs->impl.d, cs.d_inhom, false);
  GDM_CATCH
}

int gdm_integrate_difference(gdm_system_t sys, gdm_vector_t v, gdm_function_fn exact, void *user, double *cellwise,
                             double *global_l2)
{
  GDM_TRY
  GDM_ARG(sys);
  GDM_ARG(v);
  GDM_ARG(exact);
  GDM_REQUIRE(v->impl.sys == &sys->impl, GDM_ERR_INVALID, "vector/system mismatch");
  const Layout &L = sys->impl.L;
  Context      &ctx = *sys->impl.ctx;
  // postprocessing, not on the hot path: evaluate on the host from the stored (owned + ghost) block
  vector_update_ghosts(v->impl);
  std::vector<double> u((size_t)L.size);
  GDM_CUDA_CHECK(cudaMemcpyAsync(u.data(), v->impl.d, u.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
  const int                p = L.p, n1 = p + 1;
  std::vector<long double> xq, wq;
  gauss_legendre_01(n1, xq, wq);
  // shape values per variant: val[v][q][k]
  std::vector<double> val((size_t)p * n1 * n1);
  for (int vv = 0; vv < p; ++vv)
    for (int q = 0; q < n1; ++q)
      {
        long double t[MAX_DEGREE + 1];
        lagrange_eval(p, vv, xq[q], t, nullptr);
        for (int k = 0; k < n1; ++k)
          val[((size_t)vv * n1 + q) * n1 + k] = (double)t[k];
      }
  const uint64_t n_cells = gdm_system_n_cells(sys);
  if (cellwise)
    std::fill(cellwise, cellwise + n_cells, 0.0);
  // owned cells: last index in [stride*rank, stride*(rank+1))  (system.h:755)
  const int stride  = (L.N[L.pdim] + L.n_ranks - 1) / L.n_ranks;
  double    sum_sq  = 0.0;
  const int nq[3]   = {n1, L.dim >= 2 ? n1 : 1, L.dim >= 3 ? n1 : 1};
  double    jac     = 1.0;
  for (int d = 0; d < L.dim; ++d)
    jac *= L.h[d];
  for (uint64_t cell = 0; cell < n_cells; ++cell)
    {
      int      ci[3] = {0, 0, 0}, off[3] = {0, 0, 0}, var[3] = {0, 0, 0};
      uint64_t c = cell;
      for (int d = 0; d < L.dim; ++d)
        {
          ci[d] = (int)(c % L.N[d]);
          c /= L.N[d];
          off[d] = window_offset(p, L.N[d], ci[d]);
          var[d] = cell_variant(p, L.N[d], ci[d]);
        }
      if (ci[L.pdim] / stride != L.rank)
        continue;
      double diff = 0.0;
      for (int qz = 0; qz < nq[2]; ++qz)
        for (int qy = 0; qy < nq[1]; ++qy)
          for (int qx = 0; qx < nq[0]; ++qx)
            {
              const double pt[3] = {L.lo[0] + (ci[0] + (double)xq[qx]) * L.h[0],
                                    L.lo[1] + (ci[1] + (L.dim >= 2 ? (double)xq[qy] : 0.0)) * L.h[1],
                                    L.lo[2] + (ci[2] + (L.dim >= 3 ? (double)xq[qz] : 0.0)) * L.h[2]};
              double       w     = (double)wq[qx] * (L.dim >= 2 ? (double)wq[qy] : 1.0) * (L.dim >= 3 ? (double)wq[qz] : 1.0) * jac;
              for (int comp = 0; comp < L.nc; ++comp)
                {
                  double uh = 0.0;
                  for (int kz = 0; kz < nq[2]; ++kz)
                    {
                      const double sz = L.dim >= 3 ? val[((size_t)var[2] * n1 + qz) * n1 + kz] : 1.0;
                      for (int ky = 0; ky < nq[1]; ++ky)
                        {
                          const double sy = L.dim >= 2 ? val[((size_t)var[1] * n1 + qy) * n1 + ky] : 1.0;
                          for (int kx = 0; kx < n1; ++kx)
                            {
                              int idx[3] = {off[0] + kx, off[1] + ky, off[2] + kz};
                              idx[L.pdim] -= L.loc0;
                              const int64_t o = (int64_t)idx[2] * L.plane + (int64_t)idx[1] * L.pitch + (int64_t)idx[0] * L.nc + comp;
                              uh += val[((size_t)var[0] * n1 + qx) * n1 + kx] * sy * sz * u[(size_t)o];
                            }
                        }
                    }
                  const double e = uh - exact(pt, comp, user);
                  diff += e * e * w;
                }
            }
      if (cellwise)
        cellwise[cell] = std::sqrt(diff);
      sum_sq += diff;
    }
  if (global_l2)
    {
      if (L.n_ranks > 1)
        {
          GDM_CUDA_CHECK(cudaMemcpyAsync(ctx.d_sums + SUM_TMP, &sum_sq, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
          sum_sq = read_sum(ctx, SUM_TMP, true);
        }
      *global_l2 = std::sqrt(sum_sq);
    }
  GDM_CATCH
}

} // extern "C"
