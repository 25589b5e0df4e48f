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
#include <cstdlib>
#include <type_traits>

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

    // ---- fused band pass: both chains of the sum factorisation in one launch
    //     oS = (S ? A S : 0) + B P   (scaled / accumulated on the last direction),   oP = A P   (optional)
    // dir == 0: one thread per output, neighbours along x come from the same cache lines;
    // dir >= 1: a thread marches along `dir` over RJ outputs with the 2p+1 window of P (and S) in registers, so
    // every input value is read once per (RJ + 2p)/RJ instead of 2p+1 times through L2.
    struct BandK2
    {
      const double *S, *P, *tabA, *tabB;
      double       *oS, *oP;
      int           dir, nc;
      int           n_dir, wrap;
      int64_t       pitch, plane, stride;
      int           lo[3], hi[3];
      int           accumulate;
      double        scale;
    };
    constexpr int BAND_RJ = 32;

    // column index of tap t of row r (periodic fold: rows below `wrap` wrap around, the duplicate last row does not)
    __device__ __forceinline__ int band_col(int r, int t, int p, int wrap)
    {
      int c = r + t - p;
      if (wrap > 0)
        {
          if (c < 0)
            c += wrap;
          else if (c >= wrap && r < wrap)
            c -= wrap;
        }
      return c;
    }

    // Thread indices are flattened over (x element, row): a row of 257 nodes would otherwise leave the third 128-thread
    // block of every row with one active lane.
    template <int P>
    __global__ void band_x_kernel(const BandK2 a)
    {
      constexpr int W      = 2 * P + 1;
      const int64_t row_el = (int64_t)(a.hi[0] - a.lo[0]) * a.nc; // computed elements of a row
      const int64_t idx    = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= row_el * (a.hi[1] - a.lo[1]))
        return;
      const int     j    = a.lo[1] + (int)(idx / row_el);
      const int64_t ex   = (int64_t)a.lo[0] * a.nc + idx % row_el;
      const int     k    = a.lo[2] + blockIdx.z;
      const int     r    = (int)(ex / a.nc);
      const int64_t base = (int64_t)k * a.plane + (int64_t)j * a.pitch + ex;
      const double *tA = a.tabA + (int64_t)r * W, *tB = a.tabB + (int64_t)r * W;
      double        s = 0.0, q = 0.0;
#pragma unroll
      for (int t = 0; t < W; ++t)
        {
          const int c = band_col(r, t, P, a.wrap);
          if (c < 0 || c >= a.n_dir)
            continue;
          const int64_t off = base + (int64_t)(c - r) * a.stride;
          const double  pv  = __ldg(a.P + off);
          const double  ca = __ldg(tA + t), cb = __ldg(tB + t);
          s                 = fma(cb, pv, s);
          if (a.S)
            s = fma(ca, __ldg(a.S + off), s);
          q = fma(ca, pv, q);
        }
      s *= a.scale;
      if (a.accumulate)
        s += a.oS[base];
      a.oS[base] = s;
      if (a.oP)
        a.oP[base] = q;
    }

    template <int P>
    __global__ void band_march_kernel(const BandK2 a)
    {
      constexpr int W      = 2 * P + 1;
      const int64_t row_el = (int64_t)(a.hi[0] - a.lo[0]) * a.nc; // computed elements of a row
      // independent columns (x element, the index that is not marched) flattened over blockIdx.x; blockIdx.y: march chunk
      const int     n_other = (a.dir == 1) ? (a.hi[2] - a.lo[2]) : (a.hi[1] - a.lo[1]);
      const int64_t col     = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (col >= row_el * n_other)
        return;
      const int64_t ex    = (int64_t)a.lo[0] * a.nc + col % row_el;
      const int     other = (int)(col / row_el);
      int           r0, r1;
      int64_t       base0; // offset of marching index 0
      if (a.dir == 1)
        {
          r0    = a.lo[1] + blockIdx.y * BAND_RJ;
          r1    = min(r0 + BAND_RJ, a.hi[1]);
          base0 = (int64_t)(a.lo[2] + other) * a.plane + ex;
        }
      else
        {
          r0    = a.lo[2] + blockIdx.y * BAND_RJ;
          r1    = min(r0 + BAND_RJ, a.hi[2]);
          base0 = (int64_t)(a.lo[1] + other) * a.pitch + ex;
        }
      const bool has_S = a.S != nullptr, has_oP = a.oP != nullptr;
      // value at logical column c of the sliding window (rows below the periodic seam)
      auto ld = [&](const double *f, int c) -> double {
        if (a.wrap > 0)
          {
            if (c < 0)
              c += a.wrap;
            else if (c >= a.wrap)
              c -= a.wrap;
          }
        return (c < 0 || c >= a.n_dir) ? 0.0 : __ldg(f + base0 + (int64_t)c * a.stride);
      };
      double wP[W], wS[W];
#pragma unroll
      for (int t = 1; t < W; ++t)
        {
          wP[t] = ld(a.P, r0 - P + t - 1);
          wS[t] = has_S ? ld(a.S, r0 - P + t - 1) : 0.0;
        }
      constexpr int U = 4; // rows per trip: their new window values are requested together
      for (int rb = r0; rb < r1; rb += U)
        {
          double nP[U], nS[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            {
              nP[u] = ld(a.P, rb + u + P);
              nS[u] = has_S ? ld(a.S, rb + u + P) : 0.0;
            }
#pragma unroll
          for (int u = 0; u < U; ++u)
            {
              const int r = rb + u;
#pragma unroll
              for (int t = 0; t < W - 1; ++t)
                {
                  wP[t] = wP[t + 1];
                  wS[t] = wS[t + 1];
                }
              wP[W - 1] = nP[u];
              wS[W - 1] = nS[u];
              if (r >= r1)
                continue;
              const double *tA = a.tabA + (int64_t)r * W, *tB = a.tabB + (int64_t)r * W;
              double        s = 0.0, q = 0.0;
              if (a.wrap > 0 && r >= a.wrap)
                {
                  // the duplicate last row of a periodic direction does not wrap: gather it directly
#pragma unroll
                  for (int t = 0; t < W; ++t)
                    {
                      const int c = r + t - P;
                      if (c < 0 || c >= a.n_dir)
                        continue;
                      const double pv = __ldg(a.P + base0 + (int64_t)c * a.stride);
                      s               = fma(__ldg(tB + t), pv, s);
                      if (has_S)
                        s = fma(__ldg(tA + t), __ldg(a.S + base0 + (int64_t)c * a.stride), s);
                      q = fma(__ldg(tA + t), pv, q);
                    }
                }
              else
                {
#pragma unroll
                  for (int t = 0; t < W; ++t)
                    {
                      const double ca = __ldg(tA + t), cb = __ldg(tB + t);
                      s               = fma(cb, wP[t], s);
                      if (has_S)
                        s = fma(ca, wS[t], s);
                      q = fma(ca, wP[t], q);
                    }
                }
              const int64_t o = base0 + (int64_t)r * a.stride;
              s *= a.scale;
              if (a.accumulate)
                s += a.oS[o];
              a.oS[o] = s;
              if (has_oP)
                a.oP[o] = q;
            }
        }
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

    // irregular rows (cut cells, ghost penalty, Nitsche): one warp per row, 12 B per nonzero (value + 32-bit storage
    // offset of the column relative to the row) + 16 B per row
    __global__ void csr_overlay_kernel(int64_t n_rows, const int64_t *__restrict__ row_off, const int64_t *__restrict__ rowptr,
                                       const int32_t *__restrict__ col_rel, const double *__restrict__ val, double *dst,
                                       const double *__restrict__ src, int accumulate)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t row  = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      if (row >= n_rows)
        return;
      const int64_t off = row_off[row];
      const double *s   = src + off;
      double        acc = 0.0, acc2 = 0.0;
      const int64_t b = rowptr[row], e = rowptr[row + 1];
      int64_t       i = b + lane;
      for (; i + 32 < e; i += 64)
        {
          acc  = fma(val[i], s[col_rel[i]], acc);
          acc2 = fma(val[i + 32], s[col_rel[i + 32]], acc2);
        }
      if (i < e)
        acc = fma(val[i], s[col_rel[i]], acc);
      acc += acc2;
      for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0)
        dst[off] = accumulate ? dst[off] + acc : acc;
    }

    __global__ void csr_diagonal_kernel(int64_t n_rows, const int64_t *row_off, const double *d, double *diag)
    {
      const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (row < n_rows)
        diag[row_off[row]] = d[row];
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

  // fused pass of direction `dir`: oS = (S ? A S : 0) + B P (scaled / accumulated), oP = A P (optional).
  // Returns false when the degree has no instantiation (the caller falls back to the one-output passes).
  bool launch_band_pass2(Context &ctx, const Layout &L, const bool periodic[3], int dir, bool owned_only, const double *S,
                         const double *P, const double *tabA, const double *tabB, double *oS, double *oP, double scale,
                         bool accumulate)
  {
    BandK2 a;
    a.S = S, a.P = P, a.tabA = tabA, a.tabB = tabB, a.oS = oS, a.oP = oP;
    a.dir        = dir;
    a.nc         = L.nc;
    a.n_dir      = L.ln[dir];
    a.wrap       = periodic[dir] ? L.N[dir] : 0;
    a.pitch      = L.pitch;
    a.plane      = L.plane;
    a.stride     = L.stride[dir];
    a.accumulate = accumulate ? 1 : 0;
    a.scale      = scale;
    window(L, owned_only, a.lo, a.hi);
    const int64_t x_elems = (int64_t)(a.hi[0] - a.lo[0]) * L.nc;
    if (x_elems <= 0 || a.hi[1] <= a.lo[1] || a.hi[2] <= a.lo[2])
      return true;
    const int     threads = 128;
    const int     ny = a.hi[1] - a.lo[1], nz = a.hi[2] - a.lo[2];
    dim3          grid;
    if (dir == 0) // all outputs of a plane flattened; one plane per blockIdx.z
      grid = dim3((unsigned)((x_elems * ny + threads - 1) / threads), 1, (unsigned)nz);
    else if (dir == 1) // columns (x element, plane) flattened; chunks of BAND_RJ rows
      grid = dim3((unsigned)((x_elems * nz + threads - 1) / threads), (unsigned)((ny + BAND_RJ - 1) / BAND_RJ), 1);
    else // columns (x element, row) flattened; chunks of BAND_RJ planes
      grid = dim3((unsigned)((x_elems * ny + threads - 1) / threads), (unsigned)((nz + BAND_RJ - 1) / BAND_RJ), 1);
    auto launch = [&](auto pc) {
      constexpr int PP = decltype(pc)::value;
      if (dir == 0)
        band_x_kernel<PP><<<grid, threads, 0, ctx.stream>>>(a);
      else
        band_march_kernel<PP><<<grid, threads, 0, ctx.stream>>>(a);
    };
    switch (L.p)
      {
        case 1: launch(std::integral_constant<int, 1>{}); break;
        case 3: launch(std::integral_constant<int, 3>{}); break;
        case 5: launch(std::integral_constant<int, 5>{}); break;
        case 7: launch(std::integral_constant<int, 7>{}); break;
        case 9: launch(std::integral_constant<int, 9>{}); break;
        default: return false;
      }
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
    return true;
  }

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
                              const double *src, bool accumulate, int plane_lo, int plane_hi, double *dot_partials,
                              cudaStream_t stream)
  {
    if (stream == nullptr)
      stream = ctx.stream;
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
    constrained_rows_kernel<<<blocks, th, 0, stream>>>(a);
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
                                                                csr.d_col_rel, csr.d_val, dst, src,
                                                                accumulate ? 1 : 0);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void launch_csr_diagonal(Context &ctx, const CsrOverlay &csr, double *diag)
  {
    if (csr.n_rows == 0)
      return;
    const int th = 128;
    csr_diagonal_kernel<<<(unsigned)((csr.n_rows + th - 1) / th), th, 0, ctx.stream>>>(csr.n_rows, csr.d_row_off, csr.d_diag, diag);
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
        if (d == L.pdim && L.n_ranks > 1)
          {
            // the duplicate plane lives on another rank: one plane from the owner of plane 0 to the owner of plane N
            const int first = comm_plane_owner(L, 0), last = comm_plane_owner(L, L.N[d]);
            if (first != last)
              {
                if (L.rank == first)
                  comm_send(ctx, v + (int64_t)(0 - L.loc0) * L.stride[d], L.stride[d], last, ctx.stream);
                if (L.rank == last)
                  comm_recv(ctx, v + (int64_t)(L.N[d] - L.loc0) * L.stride[d], L.stride[d], first, ctx.stream);
                continue;
              }
          }
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
    const char   *env_f = std::getenv("GDM_GENERIC_FUSED");
    const bool    fused_passes = !(env_f && env_f[0] == '0') && (L.p == 1 || L.p == 3 || L.p == 5 || L.p == 7 || L.p == 9);
    for (int d = 0; d < dim; ++d)
      {
        const bool last       = (d == dim - 1);
        const bool owned_only = (d == L.pdim);
        if (fused_passes)
          {
            // one launch per direction: S_d = A_d S_{d-1} + B_d P_{d-1} and P_d = A_d P_{d-1} (mass: P_d only)
            double *oS = last ? dst : ctx.scratch[next++ & 3];
            double *oP = (op.has_B && !last) ? ctx.scratch[next++ & 3] : nullptr;
            launch_band_pass2(ctx, L, op.periodic, d, owned_only, op.has_B ? S : nullptr, P, op.dA[d], op.has_B ? op.dB[d] : op.dA[d], oS,
                              oP, last ? op.desc.scale : 1.0, last && accumulate);
            if (op.has_B)
              {
                S = oS;
                if (oP)
                  P = oP;
              }
            else
              P = oS;
            continue;
          }
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
