// BLAS-1 on padded device vectors and the fused CG vector updates.
//
// Stand in for LinearAlgebra::distributed::Vector<double>::{add, sadd, scale, operator*, l2_norm}
// and for the vector part of deal.II's SolverCG (call sites: every solver.solve in the reference,
// e.g. tests/poisson_01_gdm.cc:164-170).  All kernels are HBM-bound streaming kernels:
// 128-bit loads/stores where the block is 16-byte aligned, grid sized to the SM count,
// warp-shuffle + shared-memory block reduction, deterministic final reduction by the last
// block to finish (fixed summation order => bitwise reproducible dots).
// Scalars (alpha, beta, residual) never visit the host: each kernel derives them from the raw
// sums left in ctx.d_sums by its predecessor.
#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    constexpr int RED_THREADS = 256;

    __device__ __forceinline__ double warp_sum(double v)
    {
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
      return v;
    }
    __device__ __forceinline__ double warp_max(double v)
    {
      for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      return v;
    }

    // Block-reduce NV values; thread 0 of the last block to arrive reduces the per-block partials
    // in index order and writes sums[slot[i]].
    template <int NV, bool MAX>
    __device__ void grid_reduce(double (&v)[NV], double *partials, unsigned *counter, double *sums,
                                const int (&slot)[NV])
    {
      __shared__ double sm[NV][RED_THREADS / 32];
      __shared__ bool   is_last;
      const int         lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        {
          const double w = MAX ? warp_max(v[i]) : warp_sum(v[i]);
          if (lane == 0)
            sm[i][warp] = w;
        }
      __syncthreads();
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int i = 0; i < NV; ++i)
            {
              double t = sm[i][0];
              for (int w = 1; w < RED_THREADS / 32; ++w)
                t = MAX ? fmax(t, sm[i][w]) : t + sm[i][w];
              partials[(size_t)i * gridDim.x + blockIdx.x] = t;
            }
          __threadfence();
          const unsigned ticket = atomicAdd(counter, 1u);
          is_last               = (ticket == gridDim.x - 1);
        }
      __syncthreads();
      if (is_last)
        {
          __threadfence();
          // parallel fixed-order reduction of the partials by the last block
#pragma unroll
          for (int i = 0; i < NV; ++i)
            {
              double t = MAX ? 0.0 : 0.0;
              for (unsigned b = threadIdx.x; b < gridDim.x; b += RED_THREADS)
                {
                  const double x = __ldcg(partials + (size_t)i * gridDim.x + b);
                  t              = MAX ? fmax(t, x) : t + x;
                }
              const double w = MAX ? warp_max(t) : warp_sum(t);
              __syncthreads();
              if (lane == 0)
                sm[i][warp] = w;
              __syncthreads();
              if (threadIdx.x == 0)
                {
                  double r = sm[i][0];
                  for (int w2 = 1; w2 < RED_THREADS / 32; ++w2)
                    r = MAX ? fmax(r, sm[i][w2]) : r + sm[i][w2];
                  sums[slot[i]] = r;
                }
            }
          if (threadIdx.x == 0)
            *counter = 0u;
        }
    }

    // ---------------------------------------------------------------- simple kernels
    template <typename F>
    __global__ void map_kernel(int64_t n, F f)
    {
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        f(i);
    }

    __global__ void set_rows_kernel(double *v, double value, int64_t row_len, int64_t pitch, int64_t n_rows)
    {
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_len * n_rows; i += stride)
        v[(i / row_len) * pitch + (i % row_len)] = value;
    }

    __global__ void __launch_bounds__(RED_THREADS)
      dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t n, double *partials,
                 unsigned *counter, double *sums, int slot)
    {
      double        acc[1] = {0.0};
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      const int64_t tid    = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if ((((uintptr_t)a | (uintptr_t)b) & 15) == 0)
        {
          const int64_t  n2 = n >> 1;
          const double2 *a2 = reinterpret_cast<const double2 *>(a);
          const double2 *b2 = reinterpret_cast<const double2 *>(b);
          for (int64_t i = tid; i < n2; i += stride)
            {
              const double2 x = a2[i], y = b2[i];
              acc[0] = fma(x.x, y.x, acc[0]);
              acc[0] = fma(x.y, y.y, acc[0]);
            }
          if (tid == 0 && (n & 1))
            acc[0] = fma(a[n - 1], b[n - 1], acc[0]);
        }
      else
        for (int64_t i = tid; i < n; i += stride)
          acc[0] = fma(a[i], b[i], acc[0]);
      const int slots[1] = {slot};
      grid_reduce<1, false>(acc, partials, counter, sums, slots);
    }

    __global__ void __launch_bounds__(RED_THREADS)
      absmax_kernel(const double *__restrict__ a, int64_t n, double *partials, unsigned *counter,
                    double *sums, int slot)
    {
      double        acc[1] = {0.0};
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc[0] = fmax(acc[0], fabs(a[i]));
      const int slots[1] = {slot};
      grid_reduce<1, true>(acc, partials, counter, sums, slots);
    }

    int reduce_grid(const Context &ctx, int64_t n)
    {
      const int64_t want = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
      const int64_t cap  = (int64_t)ctx.sm_count * 8;
      return (int)std::max<int64_t>(1, std::min(want, cap));
    }
  } // namespace

  // ------------------------------------------------------------------ host wrappers
  void blas_set(Context &ctx, double *v, int64_t n, double value)
  {
    if (n <= 0)
      return;
    map_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(n, [=] __device__(int64_t i) { v[i] = value; });
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  // set the true entries of the owned block (pads stay untouched)
  void blas_set_strided(Context &ctx, const Layout &L, double *v, double value)
  {
    int64_t row_len, pitch, n_rows;
    double *base = v + L.own_off;
    if (L.dim == 1)
      {
        row_len = (int64_t)(L.own1 - L.own0) * L.nc;
        pitch   = row_len;
        n_rows  = 1;
      }
    else
      {
        row_len = (int64_t)L.ln[0] * L.nc;
        pitch   = L.pitch;
        n_rows  = (L.dim == 2) ? (L.own1 - L.own0) : (int64_t)(L.own1 - L.own0) * L.ln[1];
      }
    if (row_len * n_rows <= 0)
      return;
    set_rows_kernel<<<reduce_grid(ctx, row_len * n_rows), RED_THREADS, 0, ctx.stream>>>(base, value, row_len, pitch, n_rows);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void blas_copy(Context &ctx, double *dst, const double *src, int64_t n)
  {
    if (n <= 0 || dst == src)
      return;
    GDM_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  }

  void blas_scale(Context &ctx, double *v, int64_t n, double a)
  {
    if (n <= 0)
      return;
    map_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(n, [=] __device__(int64_t i) { v[i] *= a; });
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  // v = s v + a x   (deal.II Vector::sadd; s == 1 is Vector::add)
  void blas_sadd(Context &ctx, double *v, double s, double a, const double *x, int64_t n)
  {
    if (n <= 0)
      return;
    if (s == 1.0)
      map_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(n, [=] __device__(int64_t i) { v[i] = fma(a, x[i], v[i]); });
    else
      map_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(n, [=] __device__(int64_t i) { v[i] = fma(s, v[i], a * x[i]); });
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void blas_mul(Context &ctx, double *v, const double *d, int64_t n)
  {
    if (n <= 0)
      return;
    map_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(n, [=] __device__(int64_t i) { v[i] *= d[i]; });
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  // v = 1/v on non-zero entries (pads are zero and stay zero)
  void blas_invert(Context &ctx, double *v, int64_t n)
  {
    if (n <= 0)
      return;
    map_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(n, [=] __device__(int64_t i) {
      const double x = v[i];
      v[i]           = (x != 0.0) ? 1.0 / x : 0.0;
    });
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void blas_dot(Context &ctx, const double *a, const double *b, int64_t n, int slot)
  {
    const int grid = reduce_grid(ctx, std::max<int64_t>(n, 1));
    dot_kernel<<<grid, RED_THREADS, 0, ctx.stream>>>(a, b, n, ctx.d_partials, ctx.d_counters, ctx.d_sums, slot);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  namespace
  {
    // fixed-order sum of n partials: thread t adds elements t, t+1024, ...; then a fixed binary tree
    __global__ void sum_partials_kernel(const double *__restrict__ partials, int n, double *__restrict__ out)
    {
      __shared__ double sh[1024];
      double            t = 0.0;
      for (int i = threadIdx.x; i < n; i += 1024)
        t += partials[i];
      sh[threadIdx.x] = t;
      __syncthreads();
      for (int o = 512; o > 0; o >>= 1)
        {
          if ((int)threadIdx.x < o)
            sh[threadIdx.x] += sh[threadIdx.x + o];
          __syncthreads();
        }
      if (threadIdx.x == 0)
        *out = sh[0];
    }
  } // namespace

  void blas_sum_partials(Context &ctx, const double *partials, int n, int slot)
  {
    sum_partials_kernel<<<1, 1024, 0, ctx.stream>>>(partials, n, ctx.d_sums + slot);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void blas_absmax(Context &ctx, const double *a, int64_t n, int slot)
  {
    const int grid = reduce_grid(ctx, std::max<int64_t>(n, 1));
    absmax_kernel<<<grid, RED_THREADS, 0, ctx.stream>>>(a, n, ctx.d_partials, ctx.d_counters, ctx.d_sums, slot);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  double read_sum(Context &ctx, int slot, bool allreduce, bool is_max)
  {
    if (allreduce && ctx.n_ranks > 1)
      comm_allreduce_sum(ctx, ctx.d_sums + slot, 1, is_max);
    GDM_CUDA_CHECK(cudaMemcpyAsync(ctx.h_pinned, ctx.d_sums + slot, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    return ctx.h_pinned[0];
  }

  // ================================================================== CG
  namespace
  {
    struct CgStatus // mirrored on the host (pinned)
    {
      int      done; // 0 iterate, 1 success, 2 failure
      unsigned last_step;
      double   last_value;
      double   initial_value;
      double   reduced_tol;
    };

    struct CgK
    {
      double       *x, *r, *p;
      const double *q, *dinv; // dinv == nullptr: identity
      int64_t       n;
      double       *sums, *partials;
      unsigned     *counter;
      CgStatus     *status;
      int           rz_cur, rz_new, rr_new; // slots
      unsigned      it;
      unsigned      max_steps;
      double        tol;
    };

    // step 0: z = P r, p = z, sums[RR] = r.r, sums[rz_new] = r.z
    __global__ void __launch_bounds__(RED_THREADS) cg_init_kernel(const CgK a)
    {
      double        acc[2] = {0.0, 0.0};
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        {
          const double r = a.r[i];
          const double z = a.dinv ? a.dinv[i] * r : r;
          a.p[i]         = z;
          acc[0]         = fma(r, r, acc[0]);
          acc[1]         = fma(r, z, acc[1]);
        }
      const int slots[2] = {a.rr_new, a.rz_new};
      grid_reduce<2, false>(acc, a.partials, a.counter, a.sums, slots);
    }

    // ReductionControl::check(0, |r0|)
    __global__ void cg_check0_kernel(CgStatus *st, const double *sums, int rr_slot, double tol, double reduce, unsigned max_steps)
    {
      const double res  = sqrt(sums[rr_slot]);
      st->initial_value = res;
      st->reduced_tol   = res * reduce;
      st->last_step     = 0;
      st->last_value    = res;
      int done          = 0;
      if (res < st->reduced_tol || res <= tol)
        done = 1;
      else if (0 >= max_steps || res != res)
        done = 2;
      st->done = done;
    }

    // x += alpha p ; r -= alpha q ; sums[RR] = r.r ; sums[rz_new] = r.(P r)    alpha = rz_cur / pq
    __global__ void __launch_bounds__(RED_THREADS) cg_update_kernel(const CgK a)
    {
      if (a.status->done != 0)
        return;
      const double  alpha  = a.sums[a.rz_cur] / a.sums[SUM_PQ];
      double        acc[2] = {0.0, 0.0};
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      const int64_t tid    = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      const bool    vec    = ((((uintptr_t)a.x | (uintptr_t)a.r | (uintptr_t)a.p | (uintptr_t)a.q | (uintptr_t)a.dinv) & 15) == 0);
      if (vec)
        {
          const int64_t  n2 = a.n >> 1;
          double2       *x2 = reinterpret_cast<double2 *>(a.x);
          double2       *r2 = reinterpret_cast<double2 *>(a.r);
          const double2 *p2 = reinterpret_cast<const double2 *>(a.p);
          const double2 *q2 = reinterpret_cast<const double2 *>(a.q);
          const double2 *d2 = reinterpret_cast<const double2 *>(a.dinv);
          for (int64_t i = tid; i < n2; i += stride)
            {
              double2       x = x2[i], r = r2[i];
              const double2 p = p2[i], q = q2[i];
              x.x = fma(alpha, p.x, x.x);
              x.y = fma(alpha, p.y, x.y);
              r.x = fma(-alpha, q.x, r.x);
              r.y = fma(-alpha, q.y, r.y);
              x2[i] = x;
              r2[i] = r;
              double2 z = r;
              if (d2)
                {
                  const double2 d = d2[i];
                  z.x *= d.x;
                  z.y *= d.y;
                }
              acc[0] = fma(r.x, r.x, acc[0]);
              acc[0] = fma(r.y, r.y, acc[0]);
              acc[1] = fma(r.x, z.x, acc[1]);
              acc[1] = fma(r.y, z.y, acc[1]);
            }
        }
      for (int64_t i = (vec ? ((a.n >> 1) << 1) : 0) + tid; i < a.n; i += stride)
        {
          const double x = fma(alpha, a.p[i], a.x[i]);
          const double r = fma(-alpha, a.q[i], a.r[i]);
          a.x[i]         = x;
          a.r[i]         = r;
          const double z = a.dinv ? a.dinv[i] * r : r;
          acc[0]         = fma(r, r, acc[0]);
          acc[1]         = fma(r, z, acc[1]);
        }
      const int slots[2] = {a.rr_new, a.rz_new};
      grid_reduce<2, false>(acc, a.partials, a.counter, a.sums, slots);
    }

    // ReductionControl::check(it, |r|); if iterate: p = P r + beta p, beta = rz_new / rz_cur
    __global__ void __launch_bounds__(RED_THREADS) cg_direction_kernel(const CgK a)
    {
      if (a.status->done != 0)
        return;
      const double res = sqrt(a.sums[a.rr_new]);
      int          done = 0;
      if (res < a.status->reduced_tol || res <= a.tol)
        done = 1;
      else if (a.it >= a.max_steps || res != res)
        done = 2;
      if (blockIdx.x == 0 && threadIdx.x == 0)
        {
          // `done` itself is written last and only when it changes: other blocks of this launch
          // reach the same decision from the sums, later launches see the flag.
          a.status->last_step  = a.it;
          a.status->last_value = res;
          if (done)
            {
              __threadfence();
              a.status->done = done;
            }
        }
      if (done)
        return;
      const double  beta   = a.sums[a.rz_new] / a.sums[a.rz_cur];
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        {
          const double r = a.r[i];
          const double z = a.dinv ? a.dinv[i] * r : r;
          a.p[i]         = fma(beta, a.p[i], z);
        }
    }
  } // namespace

  // Exposed to cg.cu through plain functions (keeps CgK private to this TU)
  struct CgLaunch
  {
    Context *ctx;
    CgK      k;
  };

  // one status record per context, reused by every solve (no cudaMalloc / cudaFree -- which synchronises the device --
  // per solve: the RK loops do four mass solves per step)
  void *cg_status_alloc(Context &ctx)
  {
    if (!ctx.d_cg_status)
      GDM_CUDA_CHECK(cudaMalloc(&ctx.d_cg_status, sizeof(CgStatus)));
    GDM_CUDA_CHECK(cudaMemsetAsync(ctx.d_cg_status, 0, sizeof(CgStatus), ctx.stream));
    return ctx.d_cg_status;
  }
  void cg_status_free(void *)
  {}
  void cg_status_read(Context &ctx, void *d_status, int &done, unsigned &last_step, double &last_value, double &initial)
  {
    CgStatus *h = reinterpret_cast<CgStatus *>(ctx.h_pinned + 8);
    GDM_CUDA_CHECK(cudaMemcpyAsync(h, d_status, sizeof(CgStatus), cudaMemcpyDeviceToHost, ctx.stream));
    GDM_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    done       = h->done;
    last_step  = h->last_step;
    last_value = h->last_value;
    initial    = h->initial_value;
  }
  const int *cg_status_done_flag(void *d_status)
  {
    return &reinterpret_cast<CgStatus *>(d_status)->done;
  }

  void cg_launch_init(Context &ctx, double *r, double *p, const double *dinv, int64_t n, void *status, int rr_new, int rz_new)
  {
    CgK k{};
    k.r = r;
    k.p = p;
    k.dinv = dinv;
    k.n = n;
    k.sums = ctx.d_sums;
    k.partials = ctx.d_partials;
    k.counter = ctx.d_counters;
    k.status = reinterpret_cast<CgStatus *>(status);
    k.rz_new = rz_new;
    k.rr_new = rr_new;
    cg_init_kernel<<<reduce_grid(ctx, std::max<int64_t>(n, 1)), RED_THREADS, 0, ctx.stream>>>(k);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }
  void cg_launch_check0(Context &ctx, void *status, int rr_slot, double tol, double reduce, unsigned max_steps)
  {
    cg_check0_kernel<<<1, 1, 0, ctx.stream>>>(reinterpret_cast<CgStatus *>(status), ctx.d_sums, rr_slot, tol, reduce, max_steps);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }
  void cg_launch_update(Context &ctx, double *x, double *r, const double *p, const double *q, const double *dinv,
                        int64_t n, void *status, int rz_cur, int rr_new, int rz_new)
  {
    CgK k{};
    k.x = x;
    k.r = r;
    k.p = const_cast<double *>(p);
    k.q = q;
    k.dinv = dinv;
    k.n = n;
    k.sums = ctx.d_sums;
    k.partials = ctx.d_partials;
    k.counter = ctx.d_counters;
    k.status = reinterpret_cast<CgStatus *>(status);
    k.rz_cur = rz_cur;
    k.rz_new = rz_new;
    k.rr_new = rr_new;
    cg_update_kernel<<<reduce_grid(ctx, std::max<int64_t>(n, 1)), RED_THREADS, 0, ctx.stream>>>(k);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }
  void cg_launch_direction(Context &ctx, const double *r, double *p, const double *dinv, int64_t n, void *status,
                           int rz_cur, int rr_new, int rz_new, unsigned it, unsigned max_steps, double tol)
  {
    CgK k{};
    k.r = const_cast<double *>(r);
    k.p = p;
    k.dinv = dinv;
    k.n = n;
    k.sums = ctx.d_sums;
    k.status = reinterpret_cast<CgStatus *>(status);
    k.rz_cur = rz_cur;
    k.rz_new = rz_new;
    k.rr_new = rr_new;
    k.it = it;
    k.max_steps = max_steps;
    k.tol = tol;
    cg_direction_kernel<<<reduce_grid(ctx, std::max<int64_t>(n, 1)), RED_THREADS, 0, ctx.stream>>>(k);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gdm

// ================================================================== RK stage combinations
namespace gdm
{
  namespace
  {
    struct LinComb
    {
      double       *out;
      const double *y;
      const double *k[4];
      double        c[4];
      int           nt;
      int64_t       n;
    };

    // out = y + c0 k0 + c1 k1 + ... accumulated left to right (Vector::sadd(1, c_j, k_j) order)
    __global__ void __launch_bounds__(RED_THREADS) lincomb_kernel(const LinComb a)
    {
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        {
          double v = a.y[i];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < a.nt)
              v = fma(a.c[j], a.k[j][i], v);
          a.out[i] = v;
        }
    }
  } // namespace

  void blas_lincomb(Context &ctx, double *out, const double *y, int64_t n, int nt, const double *c,
                    const double *const *k)
  {
    if (n <= 0)
      return;
    LinComb a{};
    a.out = out;
    a.y   = y;
    a.nt  = nt;
    a.n   = n;
    for (int j = 0; j < nt && j < 4; ++j)
      {
        a.c[j] = c[j];
        a.k[j] = k[j];
      }
    lincomb_kernel<<<reduce_grid(ctx, n), RED_THREADS, 0, ctx.stream>>>(a);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gdm
