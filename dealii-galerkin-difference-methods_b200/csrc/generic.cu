// Generic (any dim <= 3, any odd degree, any n_components, periodic) tensor-product apply:
// one banded 1D pass per launch, sum-factorised
//     P_d = A_d P_{d-1},  S_d = A_d S_{d-1} + B_d P_{d-1},   y = scale * S_{dim-1}  (or P_{dim-1}).
// It is the coverage path (and the cross-check for the fused sm_100a kernel in kron3d.cu);
// traffic is ~5x the fused kernel's because the intermediate fields round-trip through HBM.
// Also here: constrained-row diagonal, CSR overlay for irregular rows, constraint helpers.
//
// Replaces SparseMatrix::vmult on matrices assembled by the reference's cell loops
// (include/gdm/matrix_creator.h:21-61, tests/poisson_02_gdm.cc:160-206).
#include <algorithm>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    struct BandK
    {
      const double *src1, *tab1, *src2, *tab2;
      double       *dst;
      int           dir, p, nc;
      int           n_dir; // stored nodes along dir
      int           wrap;  // periodic modulus (cells) or 0
      int64_t       pitch, plane, stride;
      int           lo[3], hi[3]; // compute window (local node indices)
      int           accumulate;
      double        scale;
    };

    __global__ void band_pass_kernel(const BandK a)
    {
      const int64_t ex = (int64_t)a.lo[0] * a.nc + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (ex >= (int64_t)a.hi[0] * a.nc)
        return;
      const int     j    = a.lo[1] + blockIdx.y;
      const int     k    = a.lo[2] + blockIdx.z;
      const int     r    = (a.dir == 0) ? (int)(ex / a.nc) : (a.dir == 1 ? j : k);
      const int64_t base = (int64_t)k * a.plane + (int64_t)j * a.pitch + ex;
      const int     W    = 2 * a.p + 1;
      const double *t1   = a.tab1 + (int64_t)r * W;
      const double *t2   = a.tab2 ? a.tab2 + (int64_t)r * W : nullptr;
      double        acc  = 0.0;
      for (int t = 0; t < W; ++t)
        {
          int c = r + t - a.p;
          if (a.wrap > 0)
            {
              if (c < 0)
                c += a.wrap;
              else if (c >= a.wrap && r < a.wrap)
                c -= a.wrap;
            }
          if (c < 0 || c >= a.n_dir)
            continue;
          const int64_t off = base + (int64_t)(c - r) * a.stride;
          acc               = fma(__ldg(t1 + t), __ldg(a.src1 + off), acc);
          if (t2)
            acc = fma(__ldg(t2 + t), __ldg(a.src2 + off), acc);
        }
      acc *= a.scale;
      if (a.accumulate)
        acc += a.dst[base];
      a.dst[base] = acc;
    }

    struct FaceK
    {
      double       *dst;
      const double *src;
      const double *diagA[3], *diagB[3];
      int           dim, nc, has_B, accumulate;
      int           n_faces;           // all constrained faces are handled by ONE launch
      int           face_d[6], face_node[6]; // face f: local index face_node[f] in direction face_d[f]
      int64_t       face_off[7];       // prefix sums of the face sizes
      int           ln[3];
      int           con_lo[3], con_hi[3]; // constrained end nodes per direction (local index or -1)
      int           own_lo, own_hi, pdim; // owned window (local indices) in pdim
      int64_t       stride[3];
      double        scale;
      double       *dot_partials; // optional: per-block partial sums of src * dst
    };

    __device__ __forceinline__ void face_index(int dim, int d, const int *ln, int64_t tid, int node,
                                               int *idx, bool &valid)
    {
      int e0 = -1, e1 = -1;
      for (int e = 0; e < dim; ++e)
        if (e != d)
          {
            if (e0 < 0)
              e0 = e;
            else
              e1 = e;
          }
      const int64_t n0 = e0 >= 0 ? ln[e0] : 1;
      const int64_t n1 = e1 >= 0 ? ln[e1] : 1;
      valid            = tid < n0 * n1;
      idx[0] = idx[1] = idx[2] = 0;
      idx[d]                   = node;
      if (e0 >= 0)
        idx[e0] = (int)(tid % n0);
      if (e1 >= 0)
        idx[e1] = (int)(tid / n0);
    }

    // value written on a constrained (Dirichlet face) row: deal.II's diagonal convention, |scale * diag|
    __device__ __forceinline__ bool constrained_row(const FaceK &a, int64_t tid, int64_t &off, double &val)
    {
      if (tid >= a.face_off[a.n_faces])
        return false;
      int f = 0;
      while (tid >= a.face_off[f + 1])
        ++f;
      tid -= a.face_off[f];
      const int fd = a.face_d[f];
      int       idx[3];
      bool      valid;
      face_index(a.dim, fd, a.ln, tid, a.face_node[f], idx, valid);
      if (!valid)
        return false;
      if (idx[a.pdim] < a.own_lo || idx[a.pdim] >= a.own_hi)
        return false;
      for (int e = 0; e < fd; ++e) // the lowest constrained direction handles the node
        if (idx[e] == a.con_lo[e] || idx[e] == a.con_hi[e])
          return false;
      if (a.has_B)
        {
          val = 0.0;
          for (int d = 0; d < a.dim; ++d)
            {
              double t = a.diagB[d][idx[d]];
              for (int e = 0; e < a.dim; ++e)
                if (e != d)
                  t *= a.diagA[e][idx[e]];
              val += t;
            }
        }
      else
        {
          val = 1.0;
          for (int d = 0; d < a.dim; ++d)
            val *= a.diagA[d][idx[d]];
        }
      val = fabs(val * a.scale);
      off = 0;
      for (int d = 0; d < a.dim; ++d)
        off += idx[d] * a.stride[d];
      return true;
    }

    __global__ void constrained_rows_kernel(const FaceK a)
    {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      int64_t       off = 0;
      double        val = 0.0, dsum = 0.0;
      if (constrained_row(a, tid, off, val))
        for (int c = 0; c < a.nc; ++c)
          {
            const double x = a.src[off + c];
            double       r = val * x;
            dsum           = fma(x, r, dsum); // fused dot <src, A src> over the constrained rows
            if (a.accumulate)
              r += a.dst[off + c];
            a.dst[off + c] = r;
          }
      if (a.dot_partials) // block sum in a fixed order (deterministic)
        {
          __shared__ double wsum[32];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            dsum += __shfl_down_sync(0xffffffffu, dsum, o);
          if ((threadIdx.x & 31) == 0)
            wsum[threadIdx.x >> 5] = dsum;
          __syncthreads();
          if (threadIdx.x == 0)
            {
              double t = 0.0;
              for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                t += wsum[w];
              a.dot_partials[blockIdx.x] = t;
            }
        }
    }

    __global__ void csr_overlay_kernel(int64_t n_rows, const int64_t *row_off, const int64_t *rowptr,
                                       const int64_t *col_off, const double *val, double *dst,
                                       const double *src, int accumulate)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t row  = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      if (row >= n_rows)
        return;
      double        acc = 0.0;
      const int64_t b = rowptr[row], e = rowptr[row + 1];
      for (int64_t i = b + lane; i < e; i += 32)
        acc = fma(val[i], src[col_off[i]], acc);
      for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0)
        {
          const int64_t off = row_off[row];
          dst[off]          = accumulate ? dst[off] + acc : acc;
        }
    }

    struct SetFaceK
    {
      double *v;
      int     dim, nc, d, node, src_node; // src_node >= 0: copy from that node (periodic distribute)
      int     ln[3];
      int64_t stride[3];
      double  value;
    };

    __global__ void set_face_kernel(const SetFaceK a)
    {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      int           idx[3];
      bool          valid;
      face_index(a.dim, a.d, a.ln, tid, a.node, idx, valid);
      if (!valid)
        return;
      int64_t off = 0;
      for (int d = 0; d < a.dim; ++d)
        off += idx[d] * a.stride[d];
      const int64_t soff = off + (int64_t)(a.src_node - a.node) * a.stride[a.d];
      for (int c = 0; c < a.nc; ++c)
        a.v[off + c] = (a.src_node >= 0) ? a.v[soff + c] : a.value;
    }

    struct DiagK
    {
      double       *diag;
      const double *tA[3], *tB[3];
      int           dim, p, nc, has_B;
      int           lo[3], hi[3];
      int64_t       pitch, plane;
      double        scale;
    };

    __global__ void diagonal_kernel(const DiagK a)
    {
      const int64_t ex = (int64_t)a.lo[0] * a.nc + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (ex >= (int64_t)a.hi[0] * a.nc)
        return;
      const int j = a.lo[1] + blockIdx.y, k = a.lo[2] + blockIdx.z;
      int       idx[3] = {(int)(ex / a.nc), j, k};
      const int W      = 2 * a.p + 1;
      double    val;
      if (a.has_B)
        {
          val = 0.0;
          for (int d = 0; d < a.dim; ++d)
            {
              double t = a.tB[d][(int64_t)idx[d] * W + a.p];
              for (int e = 0; e < a.dim; ++e)
                if (e != d)
                  t *= a.tA[e][(int64_t)idx[e] * W + a.p];
              val += t;
            }
        }
      else
        {
          val = 1.0;
          for (int d = 0; d < a.dim; ++d)
            val *= a.tA[d][(int64_t)idx[d] * W + a.p];
        }
      a.diag[(int64_t)k * a.plane + (int64_t)j * a.pitch + ex] = val * a.scale;
    }

    void window(const Layout &L, bool owned_only, int lo[3], int hi[3])
    {
      for (int d = 0; d < 3; ++d)
        {
          lo[d] = 0;
          hi[d] = L.ln[d];
        }
      if (owned_only)
        {
          lo[L.pdim] = L.own0 - L.loc0;
          hi[L.pdim] = L.own1 - L.loc0;
        }
    }
  } // namespace

  void launch_band_pass(Context &ctx, const Layout &L, const bool periodic[3], const BandPassArgs &in)
  {
    BandK a;
    a.src1       = in.src1;
    a.tab1       = in.tab1;
    a.src2       = in.src2;
    a.tab2       = in.tab2;
    a.dst        = in.dst;
    a.dir        = in.dir;
    a.p          = L.p;
    a.nc         = L.nc;
    a.n_dir      = L.ln[in.dir];
    a.wrap       = periodic[in.dir] ? L.N[in.dir] : 0;
    a.pitch      = L.pitch;
    a.plane      = L.plane;
    a.stride     = L.stride[in.dir];
    a.accumulate = in.accumulate ? 1 : 0;
    a.scale      = in.scale;
    window(L, in.owned_only, a.lo, a.hi);
    const int64_t x_elems = (int64_t)(a.hi[0] - a.lo[0]) * L.nc;
    if (x_elems <= 0 || a.hi[1] <= a.lo[1] || a.hi[2] <= a.lo[2])
      return;
    const int  threads = 128;
    const dim3 grid((unsigned)((x_elems + threads - 1) / threads), (unsigned)(a.hi[1] - a.lo[1]),
                    (unsigned)(a.hi[2] - a.lo[2]));
    band_pass_kernel<<<grid, threads, 0, ctx.stream>>>(a);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  namespace
  {
    // local index of the constrained end nodes of direction d (or -1)
    void constrained_ends(const Layout &L, const bool dirichlet[3][2], const bool periodic[3], int d,
                          int &lo, int &hi)
    {
      const int shift = (d == L.pdim) ? L.loc0 : 0;
      lo = hi = -1;
      if (d >= L.dim)
        return;
      if (dirichlet[d][0])
        {
          const int g = 0 - shift;
          if (g >= 0 && g < L.ln[d])
            lo = g;
        }
      if (dirichlet[d][1] || periodic[d])
        {
          const int g = L.N[d] - shift;
          if (g >= 0 && g < L.ln[d])
            hi = g;
        }
    }

    int64_t face_points(const Layout &L, int d)
    {
      int64_t n = 1;
      for (int e = 0; e < L.dim; ++e)
        if (e != d)
          n *= L.ln[e];
      return n;
    }
  } // namespace

  int launch_constrained_rows(Context &ctx, const Layout &L, const Operator &op, double *dst,
                              const double *src, bool accumulate, int plane_lo, int plane_hi, double *dot_partials)
  {
    FaceK a;
    a.dst = dst;
    a.src = src;
    a.dim = L.dim;
    a.nc  = L.nc;
    a.has_B      = op.has_B ? 1 : 0;
    a.accumulate = accumulate ? 1 : 0;
    a.pdim       = L.pdim;
    a.own_lo     = (plane_lo >= 0) ? plane_lo : L.own0 - L.loc0; // optional window in the partitioned direction
    a.own_hi     = (plane_lo >= 0) ? plane_hi : L.own1 - L.loc0;
    // GDM_DIAG_ZERO: constrained rows are written as exact zeros (residual semantics)
    a.scale      = (op.desc.constrained_diagonal == GDM_DIAG_ASSEMBLED) ? op.desc.scale : 0.0;
    for (int d = 0; d < 3; ++d)
      {
        a.diagA[d]  = op.ddiagA[d];
        a.diagB[d]  = op.ddiagB[d];
        a.ln[d]     = L.ln[d];
        a.stride[d] = L.stride[d];
        constrained_ends(L, op.dirichlet, op.periodic, d, a.con_lo[d], a.con_hi[d]);
      }
    a.n_faces     = 0;
    a.face_off[0] = 0;
    for (int d = 0; d < L.dim; ++d)
      for (int s = 0; s < 2; ++s)
        {
          const int node = s == 0 ? a.con_lo[d] : a.con_hi[d];
          if (node < 0 || (s == 1 && node == a.con_lo[d]))
            continue;
          a.face_d[a.n_faces]       = d;
          a.face_node[a.n_faces]    = node;
          a.face_off[a.n_faces + 1] = a.face_off[a.n_faces] + face_points(L, d);
          ++a.n_faces;
        }
    if (a.n_faces == 0)
      return 0;
    const int64_t n  = a.face_off[a.n_faces];
    const int     th = 128;
    a.dot_partials = dot_partials;
    const unsigned blocks = (unsigned)((n + th - 1) / th);
    constrained_rows_kernel<<<blocks, th, 0, ctx.stream>>>(a);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
    return (int)blocks;
  }

  int constrained_rows_max_blocks(const Layout &L)
  {
    int64_t n = 0;
    for (int d = 0; d < L.dim; ++d)
      n += 2 * face_points(L, d);
    return (int)((n + 127) / 128) + 1;
  }

  void launch_csr_overlay(Context &ctx, const CsrOverlay &csr, double *dst, const double *src,
                          bool accumulate)
  {
    if (csr.n_rows == 0)
      return;
    const int     th     = 128;
    const int64_t blocks = (csr.n_rows * 32 + th - 1) / th;
    csr_overlay_kernel<<<(unsigned)blocks, th, 0, ctx.stream>>>(csr.n_rows, csr.d_row_off, csr.d_rowptr,
                                                                csr.d_col_off, csr.d_val, dst, src,
                                                                accumulate ? 1 : 0);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  // AffineConstraints::distribute for periodic constraints: v[N_d] = v[0] (after Dirichlet zeroing)
  void launch_periodic_copy(Context &ctx, const Layout &L, const bool periodic[3], double *v)
  {
    for (int d = 0; d < L.dim; ++d)
      {
        if (!periodic[d])
          continue;
        GDM_REQUIRE(!(d == L.pdim && L.n_ranks > 1), GDM_ERR_NOT_IMPLEMENTED,
                    "periodicity along the partitioned direction with more than one rank");
        SetFaceK a;
        a.v        = v;
        a.dim      = L.dim;
        a.nc       = L.nc;
        a.d        = d;
        a.node     = L.N[d];
        a.src_node = 0;
        a.value    = 0.0;
        for (int e = 0; e < 3; ++e)
          {
            a.ln[e]     = L.ln[e];
            a.stride[e] = L.stride[e];
          }
        const int64_t n  = face_points(L, d);
        const int     th = 128;
        set_face_kernel<<<(unsigned)((n + th - 1) / th), th, 0, ctx.stream>>>(a);
        ctx.launches++;
      }
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void launch_set_constrained(Context &ctx, const Layout &L, const bool dirichlet[3][2],
                              const bool periodic[3], double *v, double value)
  {
    for (int d = 0; d < L.dim; ++d)
      {
        int lo, hi;
        constrained_ends(L, dirichlet, periodic, d, lo, hi);
        for (int s = 0; s < 2; ++s)
          {
            const int node = s == 0 ? lo : hi;
            if (node < 0)
              continue;
            SetFaceK a;
            a.v        = v;
            a.dim      = L.dim;
            a.nc       = L.nc;
            a.d        = d;
            a.node     = node;
            a.src_node = -1;
            a.value    = value;
            for (int e = 0; e < 3; ++e)
              {
                a.ln[e]     = L.ln[e];
                a.stride[e] = L.stride[e];
              }
            const int64_t n  = face_points(L, d);
            const int     th = 128;
            set_face_kernel<<<(unsigned)((n + th - 1) / th), th, 0, ctx.stream>>>(a);
            ctx.launches++;
          }
      }
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void launch_diagonal(Context &ctx, const Layout &L, const Operator &op, double *diag)
  {
    DiagK a;
    a.diag  = diag;
    a.dim   = L.dim;
    a.p     = L.p;
    a.nc    = L.nc;
    a.has_B = op.has_B ? 1 : 0;
    a.pitch = L.pitch;
    a.plane = L.plane;
    a.scale = op.desc.scale;
    for (int d = 0; d < 3; ++d)
      {
        a.tA[d] = op.dA[d];
        a.tB[d] = op.dB[d];
      }
    window(L, true, a.lo, a.hi);
    const int64_t x_elems = (int64_t)(a.hi[0] - a.lo[0]) * L.nc;
    if (x_elems <= 0 || a.hi[1] <= a.lo[1] || a.hi[2] <= a.lo[2])
      return;
    const int  th = 128;
    const dim3 grid((unsigned)((x_elems + th - 1) / th), (unsigned)(a.hi[1] - a.lo[1]),
                    (unsigned)(a.hi[2] - a.lo[2]));
    diagonal_kernel<<<grid, th, 0, ctx.stream>>>(a);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void generic_apply(Operator &op, double *dst, const double *src, bool accumulate)
  {
    Context      &ctx = *op.sys->ctx;
    const Layout &L   = op.sys->L;
    const int     dim = L.dim;
    ctx.ensure_scratch((size_t)L.size);
    const double *P = src, *S = nullptr;
    int           next = 0;
    for (int d = 0; d < dim; ++d)
      {
        const bool last       = (d == dim - 1);
        const bool owned_only = (d == L.pdim);
        if (op.has_B)
          {
            double      *Sn = last ? dst : ctx.scratch[next++ & 3];
            BandPassArgs a;
            a.dir        = d;
            a.owned_only = owned_only;
            a.dst        = Sn;
            a.scale      = last ? op.desc.scale : 1.0;
            a.accumulate = last && accumulate;
            if (d == 0)
              {
                a.src1 = P;
                a.tab1 = op.dB[0];
              }
            else
              {
                a.src1 = S;
                a.tab1 = op.dA[d];
                a.src2 = P;
                a.tab2 = op.dB[d];
              }
            launch_band_pass(ctx, L, op.periodic, a);
            if (!last)
              {
                double      *Pn = ctx.scratch[next++ & 3];
                BandPassArgs b;
                b.dir  = d;
                b.dst  = Pn;
                b.src1 = P;
                b.tab1 = op.dA[d];
                launch_band_pass(ctx, L, op.periodic, b);
                P = Pn;
              }
            S = Sn;
          }
        else
          {
            double      *Pn = last ? dst : ctx.scratch[next++ & 3];
            BandPassArgs a;
            a.dir        = d;
            a.owned_only = owned_only;
            a.dst        = Pn;
            a.src1       = P;
            a.tab1       = op.dA[d];
            a.scale      = last ? op.desc.scale : 1.0;
            a.accumulate = last && accumulate;
            launch_band_pass(ctx, L, op.periodic, a);
            P = Pn;
          }
      }
    if (op.desc.constrained_diagonal == GDM_DIAG_ASSEMBLED)
      launch_constrained_rows(ctx, L, op, dst, src, true);
  }
  namespace
  {
    // compact (host order) <-> padded device layout for a range of rows
    __global__ void repack_kernel(double *padded, double *compact, int64_t row_len, int64_t pitch, int64_t n_rows, int to_padded)
    {
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_len * n_rows; i += stride)
        {
          const int64_t r = i / row_len, c = i - r * row_len;
          if (to_padded)
            padded[r * pitch + c] = compact[i];
          else
            compact[i] = padded[r * pitch + c];
        }
    }
  } // namespace

  // planes [p0, p1) of a 3D vector: padded storage <-> contiguous staging buffer (host order)
  void launch_repack(Context &ctx, const Layout &L, double *padded, double *compact, int p0, int p1, bool to_padded)
  {
    const int64_t row_len = (int64_t)L.ln[0] * L.nc, n_rows = (int64_t)(p1 - p0) * L.ln[1];
    if (n_rows <= 0)
      return;
    const int64_t n      = row_len * n_rows;
    const int     th     = 256;
    const int     blocks = (int)std::min<int64_t>((n + th - 1) / th, (int64_t)ctx.sm_count * 16);
    repack_kernel<<<blocks, th, 0, ctx.stream>>>(padded + (int64_t)p0 * L.plane, compact + (int64_t)p0 * row_len * L.ln[1], row_len,
                                                L.pitch, n_rows, to_padded ? 1 : 0);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gdm
