// Fused matrix-free GDM operator apply for dim == 3 on sm_100a:
//
//     y = scale * ( B_x (x) A_y (x) A_z  +  A_x (x) B_y (x) A_z  +  A_x (x) A_y (x) B_z ) x      (HASB)
//     y = scale * ( A_x (x) A_y (x) A_z ) x                                                      (mass)
//
// A_d, B_d are the assembled 1D GDM band matrices (half bandwidth p, Toeplitz except for the p+1
// one-sided rows at each end; Dirichlet masks folded in).  This replaces the assembled
// SparseMatrix::vmult of the reference (343 nnz/row at p=3: tests/poisson_02_gdm.cc:215,
// applications/wave/include/gdm/wave/problem.h:486-488) by sum factorisation with every input
// element read from HBM once and every output written once (16 B/DoF).
//
// Structure (2.5D blocking): a CTA owns a TX x TY tile of the xy plane and streams a chunk of z.
//   * TMA (cp.async.bulk.tensor.3d, mbarrier complete_tx) stages the (TX+2p) x (TY+2p) input tile of
//     plane k+STAGES into shared memory while plane k is processed; out-of-range coordinates are
//     zero filled by the TMA unit, which provides the domain-edge padding for free.
//   * x pass: tasks of RX consecutive outputs per row, lanes mapped to rows (conflict-free 128-bit
//     LDS/STS through odd 16-byte pitches); symmetric taps share the pair sums between A and B.
//   * y pass: lanes mapped to x (conflict-free 64-bit LDS), RY consecutive rows per thread.
//   * z pass: in registers, scatter form: 2p running accumulators per point, shifted by the FMA
//     itself (acc[j-1] = acc[j] + c_j u); per-plane coefficient rows come from a table, so the
//     one-sided z rows and the slab offset of a multi-GPU partition need no special code.
//   * one __syncthreads per plane; a,b fields are double buffered.
// Non-Toeplitz rows in x / y are recomputed from the row tables by the few threads that own them.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <type_traits>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    // Diagnostic switches of the kernels (GDM_FUSED_DBG: ablations used for profiles/r1) exist only in the experimental
    // build; the production kernels carry no debug branches.
#ifdef GDM_FUSED_EXPERIMENTAL
#define GDM_DBG(g, bit) (((g).dbg & (bit)) != 0)
#else
#define GDM_DBG(g, bit) false
#endif

    // ------------------------------------------------------------------ PTX helpers
    __device__ __forceinline__ uint32_t smem_u32(const void *p)
    {
      return (uint32_t)__cvta_generic_to_shared(p);
    }
    __device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void mbar_fence_init()
    {
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
    {
      asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
    }
    __device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
    {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst)),
                   "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                   : "memory");
    }

    // ------------------------------------------------------------------ configuration
    template <int P_, int TX_, int RY_, int NRB_, int RX_, int STAGES_, int MINB_, int NXW_ = 0>
    struct Cfg
    {
      static constexpr bool V4 = false, V5 = false, V6 = false, V7 = false;
      static constexpr int P = P_, TX = TX_, RY = RY_, NRB = NRB_, RX = RX_, STAGES = STAGES_, MINB = MINB_;
      static constexpr int NXW = NXW_; // > 0: warp-specialised kernel with NXW dedicated x-pass warps
      static constexpr int W       = 2 * P + 1;
      static constexpr int TY      = RY * NRB;
      static constexpr int NR      = TY + 2 * P;                         // rows of the staged tile
      static constexpr int PIN     = TX + 2 * P;                         // pitch of the staged tile (dense TMA box)
      static constexpr int PY      = ((NR / 2) & 1) ? NR : NR + 2;       // column pitch of the transposed a/b fields
      static constexpr int THREADS = TX * NRB + 32 * NXW_;
      static constexpr int NWARPS  = THREADS / 32;
      static constexpr int NXB     = TX / RX;
      static constexpr int NTASK   = NXB * NR;
      static constexpr int NAB     = 3;                                  // a/b buffers in flight (split barrier)
      // x pass work split: block tasks (RX outputs) on rows [0, NRM), single outputs on rows [NRM, NR)
      static constexpr int NRM     = (NTASK <= TX * NRB) ? NR : ((TX * NRB) / NXB);
      static constexpr int NREM    = (NR - NRM) * TX;                    // single-output tasks
      static constexpr int NBT     = 2 * (P + 1);                        // boundary (non-Toeplitz) rows per direction
      static constexpr int TB_DOUBLES = 2 * 2 * NBT * 8 * ((2 * P + 1 + 7) / 8); // [dir][field][row][taps padded]
      static_assert(TX % 32 == 0 && TX % RX == 0 && RX % 2 == 0 && NR % 2 == 0, "tile shape");
      static_assert(((PIN / 2) & 1) == 1 && ((PY / 2) & 1) == 1, "16-byte pitches must be odd for conflict-free LDS.128");
      static constexpr int STAGE_DOUBLES = (NR * PIN + 15) / 16 * 16;    // stage stride, 128-byte aligned for TMA
    };

    template <int P>
    struct KArgs
    {
      double       *dst;
      int64_t       pitch, plane;
      int           cx0, cx1, cy0, cy1, cz0, cz1; // output window (local node indices)
      int           xorg;                         // x origin of the tile grid: xorg - P is even (TMA needs 16-byte aligned box starts)
      int           nx, ny;                       // cells per direction (boundary rows: <= P or >= N-P)
      int           tiles_x, tiles_y, lz;
      int           nz_local;
      int           kz_lo, kz_hi;                 // input planes [kz_lo, kz_hi) only touch Toeplitz z rows
      int           dbg;                          // profiling ablations (GDM_FUSED_DBG): 1 no stores, 4 no x pass, 8 no y/z pass, 16 no phase barrier, 32 TMA wait by warp 0 only
      const double *tabAx, *tabBx, *tabAy, *tabBy; // row tables [node][2P+1]
      const double *zsA, *zsB;                     // scatter rows [plane][2P+1], scale folded in
      double        Ax[P + 1], Bx[P + 1], Ay[P + 1], By[P + 1]; // interior taps by distance
      double        Az[2 * P + 1], Bz[2 * P + 1];               // interior scatter row, scale folded in
      double        sigma;                                      // v4 tap split: sum_d alpha_d (0 without the split)
      const double *zt;                                         // v4: scatter rows of the non-Toeplitz plane classes [class][field][W+1]
      const int4   *segs;                                       // v5: work segments {tile x, tile y, z0, z1}
      const int    *seg_ptr;                                    // v5: segments of CTA b are [seg_ptr[b], seg_ptr[b+1])
      // fused dot product <src, A src> (CG: p . A p): every CTA writes the sum over the points it stored to
      // dot_partials[blockIdx.x]; null = disabled.  dot_src has the layout of dst.
      const double *dot_src;
      double       *dot_partials;
    };

    // sum of `v` over the CTA in a fixed order (warp shuffles, then the warp sums in index order): deterministic
    template <int NWARPS>
    __device__ __forceinline__ void block_dot_store(double v, double *scratch, double *dst_partial)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
      __syncthreads(); // the scratch area (a TMA stage) is no longer read by anyone
      if ((threadIdx.x & 31) == 0)
        scratch[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0)
        {
          double t = 0.0;
          for (int w = 0; w < NWARPS; ++w)
            t += scratch[w];
          *dst_partial = t;
        }
    }

    template <class C, bool HASB>
    constexpr size_t smem_bytes()
    {
      return (size_t)(C::STAGES * C::STAGE_DOUBLES + C::NAB * (HASB ? 2 : 1) * C::TX * C::PY + C::NAB * 2 * 16 + C::TB_DOUBLES) * sizeof(double) +
             (2 * C::STAGES + 2 * C::NAB + 2) * sizeof(uint64_t) + 128;
    }

    __device__ __forceinline__ void mbar_arrive(uint64_t *bar)
    {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }

    // z pass (scatter form): y_r += A_z[r][k] u2 + B_z[r][k] u1 for the 2P+1 rows r around input plane k;
    // the accumulators shift by one plane through the FMAs themselves.
    template <int P, int RY, bool HASB>
    __device__ __forceinline__ void z_pass(const double (&zA)[2 * P + 1], const double (&zB)[2 * P + 1], const double (&u1)[RY],
                                           const double (&u2)[RY], double (&acc)[RY][2 * P], double (&res)[RY])
    {
#pragma unroll
      for (int i = 0; i < RY; ++i)
        {
          const double u = HASB ? u2[i] : u1[i];
          double       t = fma(zA[0], u, acc[i][0]);
          if (HASB)
            t = fma(zB[0], u1[i], t);
          res[i] = t;
#pragma unroll
          for (int j = 1; j < 2 * P; ++j)
            {
              double s = fma(zA[j], u, acc[i][j]);
              if (HASB)
                s = fma(zB[j], u1[i], s);
              acc[i][j - 1] = s;
            }
          double s = zA[2 * P] * u;
          if (HASB)
            s = fma(zB[2 * P], u1[i], s);
          acc[i][2 * P - 1] = s;
        }
    }

    // ------------------------------------------------------------------ the kernel
    template <class C, bool HASB, int BSYM, bool ACCUM, bool DOT = false>
    __global__ void __launch_bounds__(C::THREADS, C::MINB) kron3d_kernel(const __grid_constant__ CUtensorMap tmap, const KArgs<C::P> g)
    {
      constexpr int P = C::P, W = C::W, TX = C::TX, RY = C::RY, RX = C::RX, NR = C::NR, PIN = C::PIN, PY = C::PY;
      constexpr int NF = HASB ? 2 : 1, NAB = C::NAB;
      // dynamic shared memory, addressed through the typed array so that the compiler emits LDS/STS
      extern __shared__ __align__(128) double smem[];
      constexpr int AB_BUF  = NF * TX * PY;                 // one a/b buffer: [field][x][PY] (y contiguous)
      constexpr int OFF_AB  = C::STAGES * C::STAGE_DOUBLES;
      constexpr int OFF_ZT  = OFF_AB + NAB * AB_BUF;        // [NAB][2][16]
      constexpr int OFF_TB  = OFF_ZT + NAB * 2 * 16;        // boundary-row coefficient tables [dir][field][row class][WP]
      constexpr int WP      = 8 * ((W + 7) / 8);            // padded taps per row
      constexpr int NBT     = C::NBT;
      constexpr int OFF_BAR = OFF_TB + C::TB_DOUBLES;       // [STAGES] TMA barriers + 1 phase barrier
      uint64_t     *bars    = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
      uint64_t     *xbar    = bars + C::STAGES;             // "x pass of plane k done by every warp"

      const int tid = threadIdx.x;
      int       b   = blockIdx.x;
      const int tx  = b % g.tiles_x;
      b /= g.tiles_x;
      const int ty    = b % g.tiles_y;
      const int chunk = b / g.tiles_y;
      const int x0    = g.xorg + tx * TX;
      const int y0    = g.cy0 + ty * C::TY;
      const int zc0   = g.cz0 + chunk * g.lz;
      const int zc1   = min(zc0 + g.lz, g.cz1);
      const int kbeg = zc0 - P, kend = zc1 + P;
      const int ncols   = min(TX, g.cx1 - x0);    // output columns of this tile (columns below cx0 are masked)
      const int nrows   = min(C::TY, g.cy1 - y0); // valid output rows
      const int nr_need = nrows + 2 * P;          // staged rows that feed valid outputs

      constexpr unsigned STAGE_BYTES = NR * PIN * sizeof(double);
      if (tid == 0)
        {
          for (int s = 0; s < C::STAGES; ++s)
            mbar_init(&bars[s], 1);
          mbar_init(xbar, C::NWARPS);
          mbar_fence_init();
        }
      // the 2(P+1) one-sided rows of A/B in x and y (row class c: node c for c <= P, node N-P+(c-P-1) above)
      for (int e = tid; e < 2 * 2 * NBT * W; e += C::THREADS)
        {
          const int t = e % W, c = (e / W) % NBT, f = (e / (W * NBT)) % 2, d = e / (W * NBT * 2);
          const int n    = d ? g.ny : g.nx;
          const int node = (c <= P) ? c : n - P + (c - P - 1);
          const double *tab = d ? (f ? g.tabBy : g.tabAy) : (f ? g.tabBx : g.tabAx);
          smem[OFF_TB + ((d * 2 + f) * NBT + c) * WP + t] = (HASB || f == 0) ? __ldg(tab + node * W + t) : 0.0;
        }
      __syncthreads();
      if (tid == 0)
        for (int s = 0; s < C::STAGES; ++s)
          if (kbeg + s < kend)
            {
              mbar_expect_tx(&bars[s], STAGE_BYTES);
              tma_load_3d(smem + s * C::STAGE_DOUBLES, &tmap, &bars[s], x0 - P, y0 - P, kbeg + s);
            }

      // y/z pass ownership: lane -> x, rb -> RY consecutive rows
      const int  lx        = tid % TX;
      const int  rb        = tid / TX;
      const int  gy_first  = y0 + rb * RY;
      const bool yz_active = (lx < ncols) && (x0 + lx >= g.cx0) && (rb * RY < nrows);
      const bool y_bnd     = (gy_first <= P) || (gy_first + RY - 1 >= g.ny - P);
      // output plane pointer; advanced by one plane at the top of every iteration (first target: kbeg - P)
      double    *out       = g.dst + (int64_t)(kbeg - P - 1) * g.plane + (int64_t)gy_first * g.pitch + (x0 + lx);

      double acc[RY][2 * P];
#pragma unroll
      for (int i = 0; i < RY; ++i)
#pragma unroll
        for (int j = 0; j < 2 * P; ++j)
          acc[i][j] = 0.0;

      // per-plane z coefficients for the non-Toeplitz planes, fetched one plane ahead: ZT[it % NAB][field][j]
      const int     zj = tid % W, zf = (tid / W) & 1;
      const double *zsrc = (zf ? g.zsB : g.zsA) + zj;
      double        znext = 0.0;
      if (tid < 2 * W && kbeg >= 0 && kbeg < g.nz_local)
        znext = __ldg(zsrc + (int64_t)kbeg * W);

      // x pass of one plane: staged tile `stage` -> a/b buffer `ab`
      auto x_pass = [&](int stage, int ab) {
        const int in_off = stage * C::STAGE_DOUBLES;
        const int a_off  = OFF_AB + ab * AB_BUF;
        const int b_off  = a_off + (NF - 1) * TX * PY;
        for (int task = tid; task < C::NXB * C::NRM; task += C::THREADS)
          {
            const int r  = task % C::NRM;
            const int xb = task / C::NRM;
            if (r >= nr_need || xb * RX >= ncols)
              continue;
            double v[RX + 2 * P];
            {
              const double2 *src = reinterpret_cast<const double2 *>(smem + in_off + r * PIN + xb * RX);
#pragma unroll
              for (int q = 0; q < (RX + 2 * P) / 2; ++q)
                {
                  const double2 t = src[q];
                  v[2 * q]        = t.x;
                  v[2 * q + 1]    = t.y;
                }
            }
            double a[RX], bb[RX];
#pragma unroll
            for (int j = 0; j < RX; ++j)
              {
                const int c   = j + P;
                double    ra  = g.Ax[0] * v[c];
                double    rbv = (HASB && BSYM > 0) ? g.Bx[0] * v[c] : 0.0;
#pragma unroll
                for (int d = 1; d <= P; ++d)
                  {
                    const double s = v[c - d] + v[c + d];
                    ra             = fma(g.Ax[d], s, ra);
                    if (HASB)
                      {
                        if (BSYM > 0)
                          rbv = fma(g.Bx[d], s, rbv);
                        else
                          rbv = fma(g.Bx[d], v[c + d] - v[c - d], rbv);
                      }
                  }
                a[j]  = ra;
                bb[j] = rbv;
              }
            const int gx_first = x0 + xb * RX;
            if (gx_first <= P || gx_first + RX - 1 >= g.nx - P)
              {
#pragma unroll
                for (int j = 0; j < RX; ++j)
                  {
                    const int gx = gx_first + j;
                    if ((gx <= P || gx >= g.nx - P) && gx >= 0 && gx <= g.nx)
                      {
                        const int     rc = (gx <= P) ? gx : gx - (g.nx - P) + P + 1;
                        const double *ta = smem + OFF_TB + (0 * NBT + rc) * WP;
                        const double *tb = smem + OFF_TB + (1 * NBT + rc) * WP;
                        double        ra = 0.0, rbv = 0.0;
#pragma unroll
                        for (int t = 0; t < W; ++t)
                          {
                            ra = fma(ta[t], v[j + t], ra);
                            if (HASB)
                              rbv = fma(tb[t], v[j + t], rbv);
                          }
                        a[j]  = ra;
                        bb[j] = rbv;
                      }
                  }
              }
            // transposed store: [x][row], consecutive lanes -> consecutive rows
#pragma unroll
            for (int j = 0; j < RX; ++j)
              {
                smem[a_off + (xb * RX + j) * PY + r] = a[j];
                if (HASB)
                  smem[b_off + (xb * RX + j) * PY + r] = bb[j];
              }
          }
        // remainder rows: one output per task, spread evenly over all warps (keeps the warps in step)
        if constexpr (C::NREM > 0)
          {
            constexpr int PER_WARP = (C::NREM + C::NWARPS - 1) / C::NWARPS;
            for (int l = (tid & 31); l < PER_WARP; l += 32)
              {
                const int o = (tid >> 5) * PER_WARP + l;
                const int r = C::NRM + o / TX, x = o % TX;
                if (o >= C::NREM || r >= nr_need || x >= ncols)
                  continue;
                double v[W];
#pragma unroll
                for (int t = 0; t < W; ++t)
                  v[t] = smem[in_off + r * PIN + x + t];
                const int gx = x0 + x;
                double    ra, rbv = 0.0;
                if ((gx <= P || gx >= g.nx - P) && gx >= 0 && gx <= g.nx)
                  {
                    const int     rc = (gx <= P) ? gx : gx - (g.nx - P) + P + 1;
                    const double *ta = smem + OFF_TB + (0 * NBT + rc) * WP;
                    const double *tb = smem + OFF_TB + (1 * NBT + rc) * WP;
                    ra               = 0.0;
#pragma unroll
                    for (int t = 0; t < W; ++t)
                      {
                        ra = fma(ta[t], v[t], ra);
                        if (HASB)
                          rbv = fma(tb[t], v[t], rbv);
                      }
                  }
                else
                  {
                    ra = g.Ax[0] * v[P];
                    if (HASB && BSYM > 0)
                      rbv = g.Bx[0] * v[P];
#pragma unroll
                    for (int d = 1; d <= P; ++d)
                      {
                        const double s = v[P - d] + v[P + d];
                        ra             = fma(g.Ax[d], s, ra);
                        if (HASB)
                          {
                            if (BSYM > 0)
                              rbv = fma(g.Bx[d], s, rbv);
                            else
                              rbv = fma(g.Bx[d], v[P + d] - v[P - d], rbv);
                          }
                      }
                  }
                smem[a_off + x * PY + r] = ra;
                if (HASB)
                  smem[b_off + x * PY + r] = rbv;
              }
          }
      };

      // prologue: x pass of the first plane
      int stage = 0, parity = 0; // TMA ring position of the plane whose x pass runs next
      int xphase = 0;            // phase parity of xbar
      int abk    = 0;            // a/b buffer of the plane consumed by the current y/z pass
      mbar_wait(&bars[0], 0);
      x_pass(0, 0);
      if (tid < 2 * W)
        {
          smem[OFF_ZT + (0 * 2 + zf) * 16 + zj] = znext;
          znext = (kbeg + 1 >= 0 && kbeg + 1 < g.nz_local) ? __ldg(zsrc + (int64_t)(kbeg + 1) * W) : 0.0;
        }
      __syncwarp();
      if ((tid & 31) == 0)
        mbar_arrive(xbar);
      mbar_wait(xbar, xphase);
      xphase ^= 1;
      if (tid == 0 && kbeg + C::STAGES < kend)
        {
          mbar_expect_tx(&bars[0], STAGE_BYTES);
          tma_load_3d(smem, &tmap, &bars[0], x0 - P, y0 - P, kbeg + C::STAGES);
        }
      stage = 1 % C::STAGES;
      if (stage == 0)
        parity ^= 1;

      [[maybe_unused]] double dsum = 0.0; // fused dot product (DOT): sum of src * (A src) over the points this thread stores
      for (int k = kbeg; k < kend; ++k)
        {
          const int abn = (abk + 1 == NAB) ? 0 : abk + 1;
          // ---- x pass of plane k+1 (its buffer was last read by the y/z pass of plane k-2)
          if (k + 1 < kend)
            {
              if (!GDM_DBG(g, 32) || tid < 32)
                mbar_wait(&bars[stage], parity);
              if (!GDM_DBG(g, 4))
                x_pass(stage, abn);
              if (tid < 2 * W)
                {
                  smem[OFF_ZT + (abn * 2 + zf) * 16 + zj] = znext;
                  znext = (k + 2 >= 0 && k + 2 < g.nz_local) ? __ldg(zsrc + (int64_t)(k + 2) * W) : 0.0;
                }
            }
          __syncwarp();
          if ((tid & 31) == 0 && !GDM_DBG(g, 16))
            mbar_arrive(xbar);

          // ---- y pass + z pass of plane k (overlaps the other warps' x pass of plane k+1)
          out += g.plane;
          if (yz_active && !GDM_DBG(g, 8))
            {
              const int a_off = OFF_AB + abk * AB_BUF + lx * PY + rb * RY;
              const int b_off = a_off + (NF - 1) * TX * PY;
              double    u1[RY], u2[RY];
              {
                double aw[RY + 2 * P], bw[RY + 2 * P];
                if constexpr (RY % 2 == 0)
                  {
                    const double2 *pa = reinterpret_cast<const double2 *>(smem + a_off);
                    const double2 *pb = reinterpret_cast<const double2 *>(smem + b_off);
#pragma unroll
                    for (int q = 0; q < (RY + 2 * P) / 2; ++q)
                      {
                        const double2 t = pa[q];
                        aw[2 * q]       = t.x;
                        aw[2 * q + 1]   = t.y;
                        if (HASB)
                          {
                            const double2 s = pb[q];
                            bw[2 * q]       = s.x;
                            bw[2 * q + 1]   = s.y;
                          }
                      }
                  }
                else
                  {
#pragma unroll
                    for (int j = 0; j < RY + 2 * P; ++j)
                      {
                        aw[j] = smem[a_off + j];
                        if (HASB)
                          bw[j] = smem[b_off + j];
                      }
                  }
                if (!y_bnd)
                  {
#pragma unroll
                    for (int i = 0; i < RY; ++i)
                      {
                        const int c  = i + P;
                        double    t1 = g.Ay[0] * aw[c], t2 = 0.0;
                        if (HASB)
                          {
                            t2 = g.Ay[0] * bw[c];
                            if (BSYM > 0)
                              t2 = fma(g.By[0], aw[c], t2);
                          }
#pragma unroll
                        for (int d = 1; d <= P; ++d)
                          {
                            const double sa = aw[c - d] + aw[c + d];
                            t1              = fma(g.Ay[d], sa, t1);
                            if (HASB)
                              {
                                const double sb = bw[c - d] + bw[c + d];
                                t2              = fma(g.Ay[d], sb, t2);
                                if (BSYM > 0)
                                  t2 = fma(g.By[d], sa, t2);
                                else
                                  t2 = fma(g.By[d], aw[c + d] - aw[c - d], t2);
                              }
                          }
                        u1[i] = t1;
                        u2[i] = t2;
                      }
                  }
                else
                  {
#pragma unroll
                    for (int i = 0; i < RY; ++i)
                      {
                        const int gy = min(gy_first + i, g.ny); // rows past the domain are never stored
                        double    t1, t2 = 0.0;
                        if (gy <= P || gy >= g.ny - P)
                          {
                            const int     rc = (gy <= P) ? gy : gy - (g.ny - P) + P + 1;
                            const double *ta = smem + OFF_TB + (2 * NBT + rc) * WP;
                            const double *tb = smem + OFF_TB + (3 * NBT + rc) * WP;
                            t1               = 0.0;
#pragma unroll
                            for (int t = 0; t < W; ++t)
                              {
                                const double ca = ta[t];
                                t1              = fma(ca, aw[i + t], t1);
                                if (HASB)
                                  {
                                    t2 = fma(ca, bw[i + t], t2);
                                    t2 = fma(tb[t], aw[i + t], t2);
                                  }
                              }
                          }
                        else
                          {
                            const int c = i + P;
                            t1          = g.Ay[0] * aw[c];
                            if (HASB)
                              {
                                t2 = g.Ay[0] * bw[c];
                                if (BSYM > 0)
                                  t2 = fma(g.By[0], aw[c], t2);
                              }
#pragma unroll
                            for (int d = 1; d <= P; ++d)
                              {
                                const double sa = aw[c - d] + aw[c + d];
                                t1              = fma(g.Ay[d], sa, t1);
                                if (HASB)
                                  {
                                    const double sb = bw[c - d] + bw[c + d];
                                    t2              = fma(g.Ay[d], sb, t2);
                                    if (BSYM > 0)
                                      t2 = fma(g.By[d], sa, t2);
                                    else
                                      t2 = fma(g.By[d], aw[c + d] - aw[c - d], t2);
                                  }
                              }
                          }
                        u1[i] = t1;
                        u2[i] = t2;
                      }
                  }
              }
              double res[RY];
              if (k >= g.kz_lo && k < g.kz_hi)
                z_pass<P, RY, HASB>(g.Az, g.Bz, u1, u2, acc, res); // Toeplitz rows: coefficients from the constant bank
              else
                {
                  double zA[W], zB[W];
#pragma unroll
                  for (int j = 0; j < W; ++j)
                    {
                      zA[j] = smem[OFF_ZT + (abk * 2 + 0) * 16 + j];
                      zB[j] = HASB ? smem[OFF_ZT + (abk * 2 + 1) * 16 + j] : 0.0;
                    }
                  z_pass<P, RY, HASB>(zA, zB, u1, u2, acc, res);
                }
              const int r_out = k - P; // output plane completed by this input plane
              if (r_out >= zc0 && r_out < zc1 && !GDM_DBG(g, 1))
                {
#pragma unroll
                  for (int i = 0; i < RY; ++i)
                    if (gy_first + i < g.cy1)
                      {
                        double *o = out + (int64_t)i * g.pitch;
                        double  t = res[i];
                        if constexpr (DOT)
                          dsum = fma(__ldg(g.dot_src + (o - g.dst)), t, dsum);
                        if (ACCUM)
                          t += *o;
                        *o = t;
                      }
                }
            }
          // ---- every warp has finished the x pass of plane k+1: its stage can be refilled
          if (!GDM_DBG(g, 16))
            mbar_wait(xbar, xphase);
          xphase ^= 1;
          if (tid == 0 && k + 1 + C::STAGES < kend)
            {
              mbar_expect_tx(&bars[stage], STAGE_BYTES);
              tma_load_3d(smem + stage * C::STAGE_DOUBLES, &tmap, &bars[stage], x0 - P, y0 - P, k + 1 + C::STAGES);
            }
          if (++stage == C::STAGES)
            {
              stage = 0;
              parity ^= 1;
            }
          abk = abn;
        }
      if constexpr (DOT)
        block_dot_store<C::NWARPS>(dsum, smem, g.dot_partials + blockIdx.x);
    }

    // ------------------------------------------------------------------ warp-specialised variant
    // NXW dedicated x-pass warps (producers) and NRB y/z warps (consumers) run concurrently: the x pass is
    // shared-memory heavy, the y/z pass FP64 heavy, so the two pipes are busy at the same time instead of
    // alternating.  No CTA-wide barrier: mbarrier rings full_ab/empty_ab (a/b buffers) and full_in/empty_in
    // (TMA stages) carry the dependencies; the producers run up to NAB-1 planes ahead.
    template <class C, bool HASB, int BSYM, bool ACCUM>
    __global__ void __launch_bounds__(C::THREADS, 1) kron3d_ws_kernel(const __grid_constant__ CUtensorMap tmap, const KArgs<C::P> g)
    {
      constexpr int P = C::P, W = C::W, TX = C::TX, RY = C::RY, RX = C::RX, NR = C::NR, PIN = C::PIN, PY = C::PY;
      constexpr int NF = HASB ? 2 : 1, NAB = C::NAB, S = C::STAGES;
      constexpr int YZ_THREADS = TX * C::NRB, X_THREADS = 32 * C::NXW;
      extern __shared__ __align__(128) double smem[];
      constexpr int AB_BUF  = NF * TX * PY;
      constexpr int OFF_AB  = S * C::STAGE_DOUBLES;
      constexpr int OFF_ZT  = OFF_AB + NAB * AB_BUF;
      constexpr int OFF_TB  = OFF_ZT + NAB * 2 * 16;
      constexpr int WP      = 8 * ((W + 7) / 8);
      constexpr int NBT     = C::NBT;
      constexpr int OFF_BAR = OFF_TB + C::TB_DOUBLES;
      uint64_t     *full_in  = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
      uint64_t     *empty_in = full_in + S;
      uint64_t     *full_ab  = empty_in + S;
      uint64_t     *empty_ab = full_ab + NAB;

      const int tid = threadIdx.x;
      int       b   = blockIdx.x;
      const int tx  = b % g.tiles_x;
      b /= g.tiles_x;
      const int ty    = b % g.tiles_y;
      const int chunk = b / g.tiles_y;
      const int x0    = g.xorg + tx * TX;
      const int y0    = g.cy0 + ty * C::TY;
      const int zc0   = g.cz0 + chunk * g.lz;
      const int zc1   = min(zc0 + g.lz, g.cz1);
      const int kbeg = zc0 - P, kend = zc1 + P;
      const int ncols   = min(TX, g.cx1 - x0);
      const int nrows   = min(C::TY, g.cy1 - y0);
      const int nr_need = nrows + 2 * P;
      constexpr unsigned STAGE_BYTES = NR * PIN * sizeof(double);

      if (tid == 0)
        {
          for (int s = 0; s < S; ++s)
            {
              mbar_init(&full_in[s], 1);
              mbar_init(&empty_in[s], C::NXW);
            }
          for (int i = 0; i < NAB; ++i)
            {
              mbar_init(&full_ab[i], C::NXW);
              mbar_init(&empty_ab[i], C::NRB * (TX / 32));
            }
          mbar_fence_init();
        }
      for (int e = tid; e < 2 * 2 * NBT * W; e += C::THREADS)
        {
          const int t = e % W, c = (e / W) % NBT, f = (e / (W * NBT)) % 2, d = e / (W * NBT * 2);
          const int n    = d ? g.ny : g.nx;
          const int node = (c <= P) ? c : n - P + (c - P - 1);
          const double *tab = d ? (f ? g.tabBy : g.tabAy) : (f ? g.tabBx : g.tabAx);
          smem[OFF_TB + ((d * 2 + f) * NBT + c) * WP + t] = (HASB || f == 0) ? __ldg(tab + node * W + t) : 0.0;
        }
      __syncthreads();

      if (tid >= YZ_THREADS)
        {
          // =========================================================== producers: TMA + x pass
          const int xt = tid - YZ_THREADS;
          if (xt == 0)
            for (int s = 0; s < S; ++s)
              if (kbeg + s < kend)
                {
                  mbar_expect_tx(&full_in[s], STAGE_BYTES);
                  tma_load_3d(smem + s * C::STAGE_DOUBLES, &tmap, &full_in[s], x0 - P, y0 - P, kbeg + s);
                }
          const int     zj = xt % W, zf = (xt / W) & 1;
          const double *zsrc = (zf ? g.zsB : g.zsA) + zj;
          double        znext = 0.0;
          if (xt < 2 * W && kbeg >= 0 && kbeg < g.nz_local)
            znext = __ldg(zsrc + (int64_t)kbeg * W);
          int it = 0;
          for (int k = kbeg; k < kend; ++k, ++it)
            {
              const int s = it % S, ab = it % NAB;
              // refill the stage of the previous plane once every producer warp has left it
              if (xt == 0 && it > 0)
                {
                  const int sp = (it - 1) % S, kp = k - 1 + S;
                  if (kp < kend)
                    {
                      mbar_wait(&empty_in[sp], ((it - 1) / S) & 1);
                      mbar_expect_tx(&full_in[sp], STAGE_BYTES);
                      tma_load_3d(smem + sp * C::STAGE_DOUBLES, &tmap, &full_in[sp], x0 - P, y0 - P, kp);
                    }
                }
              mbar_wait(&full_in[s], (it / S) & 1);
              mbar_wait(&empty_ab[ab], ((it / NAB) & 1) ^ 1);
              const int in_off = s * C::STAGE_DOUBLES;
              const int a_off  = OFF_AB + ab * AB_BUF;
              const int b_off  = a_off + (NF - 1) * TX * PY;
              if (xt < 2 * W)
                {
                  smem[OFF_ZT + (ab * 2 + zf) * 16 + zj] = znext;
                  znext = (k + 1 >= 0 && k + 1 < g.nz_local) ? __ldg(zsrc + (int64_t)(k + 1) * W) : 0.0;
                }
              for (int task = xt; task < C::NTASK; task += X_THREADS)
                {
                  const int r  = task % NR;
                  const int xb = task / NR;
                  if (r >= nr_need || xb * RX >= ncols)
                    continue;
                  double v[RX + 2 * P];
                  {
                    const double2 *src = reinterpret_cast<const double2 *>(smem + in_off + r * PIN + xb * RX);
#pragma unroll
                    for (int q = 0; q < (RX + 2 * P) / 2; ++q)
                      {
                        const double2 t = src[q];
                        v[2 * q]        = t.x;
                        v[2 * q + 1]    = t.y;
                      }
                  }
                  double a[RX], bb[RX];
#pragma unroll
                  for (int j = 0; j < RX; ++j)
                    {
                      const int c   = j + P;
                      double    ra  = g.Ax[0] * v[c];
                      double    rbv = (HASB && BSYM > 0) ? g.Bx[0] * v[c] : 0.0;
#pragma unroll
                      for (int d = 1; d <= P; ++d)
                        {
                          const double sm = v[c - d] + v[c + d];
                          ra              = fma(g.Ax[d], sm, ra);
                          if (HASB)
                            {
                              if (BSYM > 0)
                                rbv = fma(g.Bx[d], sm, rbv);
                              else
                                rbv = fma(g.Bx[d], v[c + d] - v[c - d], rbv);
                            }
                        }
                      a[j]  = ra;
                      bb[j] = rbv;
                    }
                  const int gx_first = x0 + xb * RX;
                  if (gx_first <= P || gx_first + RX - 1 >= g.nx - P)
                    {
#pragma unroll
                      for (int j = 0; j < RX; ++j)
                        {
                          const int gx = gx_first + j;
                          if ((gx <= P || gx >= g.nx - P) && gx >= 0 && gx <= g.nx)
                            {
                              const int     rc = (gx <= P) ? gx : gx - (g.nx - P) + P + 1;
                              const double *ta = smem + OFF_TB + (0 * NBT + rc) * WP;
                              const double *tb = smem + OFF_TB + (1 * NBT + rc) * WP;
                              double        ra = 0.0, rbv = 0.0;
#pragma unroll
                              for (int t = 0; t < W; ++t)
                                {
                                  ra = fma(ta[t], v[j + t], ra);
                                  if (HASB)
                                    rbv = fma(tb[t], v[j + t], rbv);
                                }
                              a[j]  = ra;
                              bb[j] = rbv;
                            }
                        }
                    }
#pragma unroll
                  for (int j = 0; j < RX; ++j)
                    {
                      smem[a_off + (xb * RX + j) * PY + r] = a[j];
                      if (HASB)
                        smem[b_off + (xb * RX + j) * PY + r] = bb[j];
                    }
                }
              __syncwarp();
              if ((xt & 31) == 0)
                {
                  mbar_arrive(&full_ab[ab]);
                  mbar_arrive(&empty_in[s]);
                }
            }
        }
      else
        {
          // =========================================================== consumers: y pass + z pass
          const int  lx        = tid % TX;
          const int  rb        = tid / TX;
          const int  gy_first  = y0 + rb * RY;
          const bool yz_active = (lx < ncols) && (x0 + lx >= g.cx0) && (rb * RY < nrows);
          const bool y_bnd     = (gy_first <= P) || (gy_first + RY - 1 >= g.ny - P);
          double    *out       = g.dst + (int64_t)(kbeg - P - 1) * g.plane + (int64_t)gy_first * g.pitch + (x0 + lx);
          double     acc[RY][2 * P];
#pragma unroll
          for (int i = 0; i < RY; ++i)
#pragma unroll
            for (int j = 0; j < 2 * P; ++j)
              acc[i][j] = 0.0;
          int it = 0;
          for (int k = kbeg; k < kend; ++k, ++it)
            {
              const int ab = it % NAB;
              mbar_wait(&full_ab[ab], (it / NAB) & 1);
              out += g.plane;
              if (yz_active)
                {
                  const int a_off = OFF_AB + ab * AB_BUF + lx * PY + rb * RY;
                  const int b_off = a_off + (NF - 1) * TX * PY;
                  double    u1[RY], u2[RY];
                  {
                    double aw[RY + 2 * P], bw[RY + 2 * P];
                    if constexpr (RY % 2 == 0)
                      {
                        const double2 *pa = reinterpret_cast<const double2 *>(smem + a_off);
                        const double2 *pb = reinterpret_cast<const double2 *>(smem + b_off);
#pragma unroll
                        for (int q = 0; q < (RY + 2 * P) / 2; ++q)
                          {
                            const double2 t = pa[q];
                            aw[2 * q]       = t.x;
                            aw[2 * q + 1]   = t.y;
                            if (HASB)
                              {
                                const double2 s2 = pb[q];
                                bw[2 * q]        = s2.x;
                                bw[2 * q + 1]    = s2.y;
                              }
                          }
                      }
                    else
                      {
#pragma unroll
                        for (int j = 0; j < RY + 2 * P; ++j)
                          {
                            aw[j] = smem[a_off + j];
                            if (HASB)
                              bw[j] = smem[b_off + j];
                          }
                      }
#pragma unroll
                    for (int i = 0; i < RY; ++i)
                      {
                        const int gy = min(gy_first + i, g.ny);
                        double    t1, t2 = 0.0;
                        if (y_bnd && (gy <= P || gy >= g.ny - P))
                          {
                            const int     rc = (gy <= P) ? gy : gy - (g.ny - P) + P + 1;
                            const double *ta = smem + OFF_TB + (2 * NBT + rc) * WP;
                            const double *tb = smem + OFF_TB + (3 * NBT + rc) * WP;
                            t1               = 0.0;
#pragma unroll
                            for (int t = 0; t < W; ++t)
                              {
                                const double ca = ta[t];
                                t1              = fma(ca, aw[i + t], t1);
                                if (HASB)
                                  {
                                    t2 = fma(ca, bw[i + t], t2);
                                    t2 = fma(tb[t], aw[i + t], t2);
                                  }
                              }
                          }
                        else
                          {
                            const int c = i + P;
                            t1          = g.Ay[0] * aw[c];
                            if (HASB)
                              {
                                t2 = g.Ay[0] * bw[c];
                                if (BSYM > 0)
                                  t2 = fma(g.By[0], aw[c], t2);
                              }
#pragma unroll
                            for (int d = 1; d <= P; ++d)
                              {
                                const double sa = aw[c - d] + aw[c + d];
                                t1              = fma(g.Ay[d], sa, t1);
                                if (HASB)
                                  {
                                    const double sb = bw[c - d] + bw[c + d];
                                    t2              = fma(g.Ay[d], sb, t2);
                                    if (BSYM > 0)
                                      t2 = fma(g.By[d], sa, t2);
                                    else
                                      t2 = fma(g.By[d], aw[c + d] - aw[c - d], t2);
                                  }
                              }
                          }
                        u1[i] = t1;
                        u2[i] = t2;
                      }
                  }
                  double res[RY];
                  if (k >= g.kz_lo && k < g.kz_hi)
                    z_pass<P, RY, HASB>(g.Az, g.Bz, u1, u2, acc, res);
                  else
                    {
                      double zA[W], zB[W];
#pragma unroll
                      for (int j = 0; j < W; ++j)
                        {
                          zA[j] = smem[OFF_ZT + (ab * 2 + 0) * 16 + j];
                          zB[j] = HASB ? smem[OFF_ZT + (ab * 2 + 1) * 16 + j] : 0.0;
                        }
                      z_pass<P, RY, HASB>(zA, zB, u1, u2, acc, res);
                    }
                  const int r_out = k - P;
                  if (r_out >= zc0 && r_out < zc1)
                    {
#pragma unroll
                      for (int i = 0; i < RY; ++i)
                        if (gy_first + i < g.cy1)
                          {
                            double *o = out + (int64_t)i * g.pitch;
                            double  t = res[i];
                            if (ACCUM)
                              t += *o;
                            *o = t;
                          }
                    }
                }
              // this warp no longer needs the a/b buffer (and its z-table slot)
              __syncwarp();
              if ((tid & 31) == 0)
                mbar_arrive(&empty_ab[ab]);
            }
        }
    }

#include "kron3d_v4.cuh"
#include "kron3d_v5.cuh"
#include "kron3d_v6.cuh"
#include "kron3d_v7.cuh"

    // ------------------------------------------------------------------ host side
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

    EncodeTiledFn encode_fn()
    {
      static EncodeTiledFn fn = nullptr;
      if (!fn)
        {
          void                           *p = nullptr;
          cudaDriverEntryPointQueryResult qres;
          GDM_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
          GDM_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, GDM_ERR_CUDA, "cuTensorMapEncodeTiled not available");
          fn = reinterpret_cast<EncodeTiledFn>(p);
        }
      return fn;
    }

    struct FusedPlan
    {
      void    *pers = nullptr; // persistent ramp-free kernel (kron3d_pers.cu); used for every launch when set
      int      slots_limit = 0; // persistent kernel: CTAs of the next launch (0: all slots)
      int      cfg = 0;     // index into GDM_FUSED_CONFIGS
      int      tiles_x = 0, tiles_y = 0, n_chunks = 0, lz = 0;
      int      cx0, cx1, cy0, cy1, cz0, cz1, xorg;
      bool     tune = false; // pick the z-chunk length by timing candidates on the first apply
      int      wz0 = -1, wz1 = -1, wlz = 0; // output-plane sub-window of the next launch (-1: whole slab)
      bool     use_comm_stream = false;     // launch on the communication stream (slab faces, behind the ghost import)
      double  *d_zsA = nullptr, *d_zsB = nullptr;
      // v4: effective B tables (B, or R = B - alpha A with the tap split), plane-class scatter table
      bool                rsplit = false;
      double              sigma  = 0.0;
      int                 kz_lo = 0, kz_hi = 0;
      std::vector<double> hBe[3];
      double             *d_Be[2] = {nullptr, nullptr};
      double             *d_zt    = nullptr;
      // v5: balanced partitions per output-plane window (cz0, cz1) -> device segment lists
      struct Partition
      {
        int   grid = 0;
        int4 *d_segs = nullptr;
        int  *d_ptr  = nullptr;
      };
      std::map<std::pair<int, int>, Partition> parts;
      // fused dot product: per-CTA partial sums of the launches of one apply (tile kernels, then the face kernel)
      double *d_dot      = nullptr;
      size_t  dot_cap    = 0;
      int     dot_cursor = -1; // >= 0 while an apply with a fused dot is being enqueued
      const double *dot_src = nullptr;
      std::map<const void *, CUtensorMap> maps;
      ~FusedPlan()
      {
        if (pers)
          pers_plan_destroy(pers);
        cudaFree(d_zsA);
        cudaFree(d_zsB);
        cudaFree(d_Be[0]);
        cudaFree(d_Be[1]);
        cudaFree(d_zt);
        cudaFree(d_dot);
        for (auto &kv : parts)
          {
            cudaFree(kv.second.d_segs);
            cudaFree(kv.second.d_ptr);
          }
      }
    };

    // available tile configurations (selected per degree; GDM_FUSED_CFG=<id> overrides for tuning)
    //                      id        P  TX RY NRB RX ST MINB
    // Production configurations (defaults per degree) and, with -DGDM_FUSED_EXPERIMENTAL (build.py:
    // GDM_BUILD_EXPERIMENTAL=1), the control-structure experiments of round 1 (v5/v6/v7).
#define GDM_FUSED_CONFIGS_CORE(X)       \
  /* v3 (split phase barrier):  P  TX RY NRB RX ST MINB [NXW] */ \
  X(0, Cfg<1, 32, 8, 4, 8, 3, 2>)       \
  X(2, Cfg<5, 32, 4, 6, 8, 3, 2>)       \
  X(14, Cfg<3, 32, 4, 8, 4, 3, 2>)      \
  X(15, Cfg<5, 32, 4, 8, 8, 3, 1>)      \
  /* v4 (lean, one CTA barrier per plane, tap split): P RY NRB RX ST MINB */ \
  X(100, Cfg4<3, 4, 8, 4, 3, 2>)        \
  X(101, Cfg4<1, 4, 8, 4, 3, 2>)        \
  X(102, Cfg4<5, 4, 8, 4, 3, 2>)
#ifdef GDM_FUSED_EXPERIMENTAL
#define GDM_FUSED_CONFIGS_EXP(X)        \
  X(6, Cfg<3, 32, 4, 8, 8, 3, 2>)       \
  X(40, Cfg<3, 32, 4, 8, 4, 3, 1, 4>)   \
  X(103, Cfg4<3, 4, 8, 8, 3, 2>)        \
  X(109, Cfg4<1, 8, 4, 4, 3, 2>)        \
  X(130, Cfg4<3, 4, 4, 4, 3, 4>)        \
  X(133, Cfg4<3, 4, 6, 4, 3, 2>)        \
  X(121, Cfg4<1, 8, 4, 4, 3, 4>)        \
  X(124, Cfg4<1, 4, 8, 4, 3, 3>)        \
  /* v5 (mbarrier rings, tile-major balanced partition) */ \
  X(200, Cfg5<3, 4, 8, 4, 3, 2>)        \
  /* v6 (register resident, warps independent):  P RY NW ST MINB */ \
  X(300, Cfg6<3, 8, 4, 4, 3>)           \
  X(301, Cfg6<1, 8, 4, 4, 4>)           \
  X(304, Cfg6<3, 8, 4, 4, 2>)           \
  X(312, Cfg6<3, 8, 4, 8, 2>)           \
  /* v7 (v4 + aligned balanced partition): P RY NRB RX ST MINB */ \
  X(400, Cfg7<3, 4, 8, 4, 3, 2>)        \
  X(401, Cfg7<1, 4, 8, 4, 3, 2>)        \
  X(402, Cfg7<5, 4, 8, 4, 3, 2>)
#else
#define GDM_FUSED_CONFIGS_EXP(X)
#endif
#define GDM_FUSED_CONFIGS(X) GDM_FUSED_CONFIGS_CORE(X) GDM_FUSED_CONFIGS_EXP(X)

    template <class F>
    void with_config(int id, F &&f)
    {
      switch (id)
        {
#define GDM_CASE(ID, ...) \
  case ID:                \
    f(__VA_ARGS__{});     \
    break;
          GDM_FUSED_CONFIGS(GDM_CASE)
#undef GDM_CASE
          default:
            throw Error(GDM_ERR_INVALID, "fused kernel configuration " + std::to_string(id) +
                        " is not in this build (experimental families need GDM_BUILD_EXPERIMENTAL=1 at build time)");
        }
    }

    int default_config(int p)
    {
      // defaults by measurement on B200 (profiles/r1/ops_families.log): p=1 -> v4, p=3 and p=5 -> v3.
      // GDM_FUSED_FAMILY=3|4|6|7 selects a kernel family for every degree, GDM_FUSED_CFG=<id> one configuration.
      int id = (p == 1) ? 101 : (p == 3 ? 14 : 15);
      if (const char *fam = std::getenv("GDM_FUSED_FAMILY"))
        {
          if (fam[0] == '3')
            id = (p == 1) ? 0 : (p == 3 ? 14 : 15);
          else if (fam[0] == '4')
            id = (p == 1) ? 101 : (p == 3 ? 100 : 102);
          else if (fam[0] == '6' && p != 5)
            id = (p == 1) ? 301 : 304;
          else if (fam[0] == '7')
            id = (p == 1) ? 401 : (p == 3 ? 400 : 402);
        }
      if (const char *env = std::getenv("GDM_FUSED_CFG"))
        {
          const int e = atoi(env);
          int       ep = -1;
          try
            {
              with_config(e, [&](auto c) { ep = decltype(c)::P; });
            }
          catch (...)
            {}
          if (ep == p)
            id = e;
        }
      return id;
    }

    template <class C>
    void fill_interior(const Operator &op, const FusedPlan &plan, KArgs<C::P> &a)
    {
      constexpr int P = C::P, W = C::W;
      const Layout &L = op.sys->L;
      // B tables seen by the kernel: the v4 kernels may run on R = B - alpha A (tap split)
      const std::vector<double> *hB = C::V4 ? plan.hBe : op.hB;
      // any interior (Toeplitz) row: P+1 is interior because N >= 2P+2 is required
      const int ix = P + 1, iy = P + 1;
      for (int d = 0; d <= P; ++d)
        {
          a.Ax[d] = op.hA[0][(size_t)ix * W + P + d];
          a.Ay[d] = op.hA[1][(size_t)iy * W + P + d];
          a.Bx[d] = op.has_B ? hB[0][(size_t)ix * W + P + d] : 0.0;
          a.By[d] = op.has_B ? hB[1][(size_t)iy * W + P + d] : 0.0;
        }
      // interior scatter row of z: zs[k][j] = scale * T_z[k - P + j][2P - j] with all rows Toeplitz,
      // valid for input planes whose 2P+1 target rows are interior: P < k - P and k + P < N_z - P (global)
      const int iz = P + 1; // global interior row; local index below
      a.kz_lo      = 0;
      a.kz_hi      = 0;
      if (L.N[2] >= 4 * P + 2)
        {
          // the table of direction 2 holds local rows (loc0 .. loc1): use the global Toeplitz row through any local interior row
          int local_interior = -1;
          for (int r = 0; r < L.ln[2]; ++r)
            if (r + L.loc0 > P && r + L.loc0 < L.N[2] - P)
              {
                local_interior = r;
                break;
              }
          if (local_interior >= 0)
            {
              for (int j = 0; j < W; ++j)
                {
                  a.Az[j] = op.desc.scale * op.hA[2][(size_t)local_interior * W + (2 * P - j)];
                  a.Bz[j] = op.has_B ? op.desc.scale * hB[2][(size_t)local_interior * W + (2 * P - j)] : 0.0;
                }
              a.kz_lo = std::max(0, 2 * P + 1 - L.loc0);
              a.kz_hi = std::min(L.ln[2], L.N[2] - 2 * P - L.loc0);
              if (a.kz_hi < a.kz_lo)
                a.kz_hi = a.kz_lo;
            }
        }
      (void)iz;
    }

    // Static partition of the (tile, plane) work of the output window [cz0, cz1) into at most `slots` CTAs.
    // A CTA costs its planes plus 2P ramp planes per segment.
    //  * linear sweep (v5, and the rest of v7): contiguous shares of the tile-major work list; a share that would end
    //    within min_seg planes of a tile column's end is snapped to it;
    //  * aligned part (v7): m = slots / tiles full segments per tile column, every column cut at the same planes so
    //    that neighbouring tiles stream through the same planes at the same time (their halos meet in L2); the planes
    //    above m L are shared among the spare slots by the sweep.  L is chosen by direct search: minimise the cost of
    //    the most expensive CTA.
    struct PartitionPlan
    {
      std::vector<int4> segs;
      std::vector<int>  ptr;
      int64_t           max_cost = 0;
      int               zl       = 0;
    };

    // sweep planes [zl, cz1) of all tiles over at most `slots` CTAs, appending to pp; returns false if it does not fit
    inline void sweep_partition(PartitionPlan &pp, int tiles, int tiles_x, int zl, int cz1, int slots, int P)
    {
      const int     min_seg = 2 * P;
      const int     nz      = cz1 - zl;
      if (nz <= 0)
        return;
      const int64_t work = (int64_t)tiles * nz;
      const int     G    = (int)std::max<int64_t>(1, std::min<int64_t>(slots, work / (2 * P)));
      int64_t       T    = std::max<int64_t>((work + (int64_t)2 * P * (G + tiles) + G - 1) / G, 4 * P);
      const size_t  segs_fixed = pp.segs.size(), ptr_fixed = pp.ptr.size();
      for (int attempt = 0;; ++attempt)
        {
          pp.segs.resize(segs_fixed);
          pp.ptr.resize(ptr_fixed);
          int64_t c = 0, cmax = 0;
          for (int t = 0; t < tiles; ++t)
            {
              int z = zl;
              while (z < cz1)
                {
                  const int64_t room = T - c - 2 * P;
                  if (room < std::min(min_seg, cz1 - z) && c > 0)
                    {
                      pp.ptr.push_back((int)pp.segs.size());
                      cmax = std::max(cmax, c);
                      c    = 0;
                      continue;
                    }
                  int       take = (int)std::min<int64_t>(cz1 - z, std::max<int64_t>(room, 1));
                  const int rem  = cz1 - z - take;
                  if (rem > 0 && rem < min_seg)
                    take = (take - (min_seg - rem) >= min_seg) ? take - (min_seg - rem) : cz1 - z;
                  pp.segs.push_back(make_int4(t % tiles_x, t / tiles_x, z, z + take));
                  c += take + 2 * P;
                  z += take;
                }
            }
          if (c > 0)
            {
              pp.ptr.push_back((int)pp.segs.size());
              cmax = std::max(cmax, c);
            }
          if ((int)(pp.ptr.size() - ptr_fixed) <= slots || attempt > 400)
            {
              pp.max_cost = std::max(pp.max_cost, cmax);
              return;
            }
          T += std::max<int64_t>(1, T / 64);
        }
    }

    inline PartitionPlan make_partition(bool aligned, int tiles, int tiles_x, int cz0, int cz1, int slots, int P, int forced_L)
    {
      const int min_seg = 2 * P;
      const int nzw     = cz1 - cz0;
      auto      build   = [&](int L) {
        PartitionPlan pp;
        pp.ptr.assign(1, 0);
        pp.zl     = cz0;
        int spare = slots;
        if (L > 0)
          {
            const int m = slots / tiles;
            for (int c = 0; c < m; ++c)
              {
                const int z0 = cz0 + c * L, z1 = std::min(cz1, z0 + L);
                if (z0 >= z1)
                  break;
                for (int t = 0; t < tiles; ++t)
                  {
                    pp.segs.push_back(make_int4(t % tiles_x, t / tiles_x, z0, z1));
                    pp.ptr.push_back((int)pp.segs.size());
                  }
                pp.max_cost = std::max<int64_t>(pp.max_cost, z1 - z0 + 2 * P);
                pp.zl       = z1;
              }
            spare = std::max(1, slots - (int)pp.ptr.size() + 1);
          }
        sweep_partition(pp, tiles, tiles_x, pp.zl, cz1, spare, P);
        return pp;
      };
      if (!aligned || slots < tiles || nzw < 4 * min_seg)
        return build(0);
      const int m    = slots / tiles;
      const int Lmax = (nzw + m - 1) / m;
      if (forced_L > 0)
        return build(std::min(std::max(forced_L, min_seg), Lmax));
      PartitionPlan best = build(Lmax);
      if (slots - m * tiles > 0)
        for (int L = Lmax - 1; L >= std::max(min_seg, Lmax / 2); --L)
          {
            if (nzw - m * L < min_seg)
              continue;
            PartitionPlan pp = build(L);
            if ((int)pp.ptr.size() - 1 <= slots && pp.max_cost < best.max_cost)
              best = std::move(pp);
          }
      return best;
    }

    template <class C>
    const FusedPlan::Partition &get_partition(Context &ctx, FusedPlan &plan, int cz0, int cz1)
    {
      const auto key = std::make_pair(cz0, cz1);
      auto       it  = plan.parts.find(key);
      if (it != plan.parts.end())
        return it->second;
      if (plan.parts.size() > 256)
        {
          GDM_CUDA_CHECK(cudaDeviceSynchronize());
          for (auto &kv : plan.parts)
            {
              cudaFree(kv.second.d_segs);
              cudaFree(kv.second.d_ptr);
            }
          plan.parts.clear();
        }
      const int tiles = std::max(1, plan.tiles_x * plan.tiles_y);
      int       slots = ctx.sm_count * C::MINB;
      if (const char *env = std::getenv("GDM_FUSED_SLOTS"))
        slots = std::max(1, atoi(env));
      const char         *envL = std::getenv("GDM_FUSED_L");
      const PartitionPlan pp   = make_partition(C::V7, tiles, plan.tiles_x, cz0, cz1, slots, C::P, envL ? atoi(envL) : 0);
      FusedPlan::Partition part;
      part.grid = (int)pp.ptr.size() - 1;
      GDM_CUDA_CHECK(cudaMalloc(&part.d_segs, std::max<size_t>(1, pp.segs.size()) * sizeof(int4)));
      GDM_CUDA_CHECK(cudaMalloc(&part.d_ptr, pp.ptr.size() * sizeof(int)));
      GDM_CUDA_CHECK(cudaMemcpy(part.d_segs, pp.segs.data(), pp.segs.size() * sizeof(int4), cudaMemcpyHostToDevice));
      GDM_CUDA_CHECK(cudaMemcpy(part.d_ptr, pp.ptr.data(), pp.ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
      GDM_CUDA_CHECK(cudaDeviceSynchronize());
      if (std::getenv("GDM_FUSED_VERBOSE"))
        fprintf(stderr, "[gdm] fused partition: planes [%d, %d) (aligned up to %d), %d tiles -> %d CTAs, %zu segments, longest CTA %lld planes\n",
                cz0, cz1, pp.zl, tiles, part.grid, pp.segs.size(), (long long)pp.max_cost);
      return plan.parts.emplace(key, part).first->second;
    }

    template <class C, bool HASB, int BSYM, bool ACCUM>
    void launch_variant(Operator &op, FusedPlan &plan, const CUtensorMap &map, double *dst)
    {
      Context      &ctx = *op.sys->ctx;
      const Layout &L   = op.sys->L;
      KArgs<C::P>   a;
      a.dst        = dst;
      a.pitch      = L.pitch;
      a.plane      = L.plane;
      a.cx0        = plan.cx0;
      a.cx1        = plan.cx1;
      a.cy0        = plan.cy0;
      a.cy1        = plan.cy1;
      // optional sub-window of output planes (multi-GPU: interior first, slab faces after the halo arrived)
      a.cz0        = (plan.wz0 >= 0) ? std::max(plan.cz0, plan.wz0) : plan.cz0;
      a.cz1        = (plan.wz0 >= 0) ? std::min(plan.cz1, plan.wz1) : plan.cz1;
      a.xorg       = plan.xorg;
      a.nx         = L.N[0];
      a.ny         = L.N[1];
      a.tiles_x    = plan.tiles_x;
      a.tiles_y    = plan.tiles_y;
      a.lz         = (plan.wz0 >= 0 && plan.wlz > 0) ? plan.wlz : plan.lz;
      a.nz_local   = L.ln[2];
      a.tabAx      = op.dA[0];
      a.tabBx      = (C::V4 && op.has_B) ? plan.d_Be[0] : op.dB[0];
      a.tabAy      = op.dA[1];
      a.tabBy      = (C::V4 && op.has_B) ? plan.d_Be[1] : op.dB[1];
      a.zsA        = plan.d_zsA;
      a.zsB        = plan.d_zsB;
      a.sigma      = plan.sigma;
      a.zt         = plan.d_zt;
      a.segs       = nullptr;
      a.seg_ptr    = nullptr;
      a.dot_src    = nullptr;
      a.dot_partials = nullptr;
      fill_interior<C>(op, plan, a);
      a.dbg = std::getenv("GDM_FUSED_DBG") ? atoi(std::getenv("GDM_FUSED_DBG")) : 0;
      void (*kern)(const CUtensorMap, const KArgs<C::P>) = nullptr;
      size_t smem = 0;
      int v5_grid = -1;
      if constexpr (C::V7)
        {
          constexpr int MODE = !HASB ? 0 : (BSYM > 0 ? 1 : 2);
          if constexpr (MODE == 1)
            kern = plan.rsplit ? kron3d_v7_kernel<C, 1, true, ACCUM> : kron3d_v7_kernel<C, 1, false, ACCUM>;
          else
            kern = kron3d_v7_kernel<C, MODE, false, ACCUM>;
          smem = smem_bytes_v7<C, HASB>();
          if (a.cz1 > a.cz0)
            {
              const FusedPlan::Partition &part = get_partition<C>(ctx, plan, a.cz0, a.cz1);
              a.segs                           = part.d_segs;
              a.seg_ptr                        = part.d_ptr;
              v5_grid                          = part.grid;
            }
        }
      else if constexpr (C::V6)
        {
          constexpr int MODE = !HASB ? 0 : (BSYM > 0 ? 1 : 2);
          if constexpr (MODE == 1)
            kern = plan.rsplit ? kron3d_v6_kernel<C, 1, true, ACCUM> : kron3d_v6_kernel<C, 1, false, ACCUM>;
          else
            kern = kron3d_v6_kernel<C, MODE, false, ACCUM>;
          smem = smem_bytes_v6<C>();
        }
      else if constexpr (C::V5)
        {
          constexpr int MODE = !HASB ? 0 : (BSYM > 0 ? 1 : 2);
          if constexpr (MODE == 1)
            kern = plan.rsplit ? kron3d_v5_kernel<C, 1, true, ACCUM> : kron3d_v5_kernel<C, 1, false, ACCUM>;
          else
            kern = kron3d_v5_kernel<C, MODE, false, ACCUM>;
          smem = smem_bytes_v5<C, HASB>();
          if (a.cz1 > a.cz0)
            {
              const FusedPlan::Partition &part = get_partition<C>(ctx, plan, a.cz0, a.cz1);
              a.segs                           = part.d_segs;
              a.seg_ptr                        = part.d_ptr;
              v5_grid                          = part.grid;
            }
        }
      else if constexpr (C::V4)
        {
          constexpr int MODE = !HASB ? 0 : (BSYM > 0 ? 1 : 2);
          if constexpr (MODE == 1)
            kern = plan.rsplit ? kron3d_v4_kernel<C, 1, true, ACCUM> : kron3d_v4_kernel<C, 1, false, ACCUM>;
          else
            kern = kron3d_v4_kernel<C, MODE, false, ACCUM>;
#ifdef GDM_FUSED_EXPERIMENTAL
          if constexpr (!ACCUM && MODE == 1)
            if (a.dbg & 32) // diagnostic: LSU + synchronisation skeleton without the FP64 arithmetic
              kern = kron3d_v4_kernel<C, 1, true, false, false, true>;
#endif
          if constexpr (!ACCUM)
            if (plan.dot_cursor >= 0) // store epilogue with the fused dot product
              {
                if constexpr (MODE == 1)
                  kern = plan.rsplit ? kron3d_v4_kernel<C, 1, true, false, true> : kron3d_v4_kernel<C, 1, false, false, true>;
                else
                  kern = kron3d_v4_kernel<C, MODE, false, false, true>;
              }
          smem = smem_bytes_v4<C, HASB>();
        }
      else
        {
          kern = (C::NXW > 0) ? kron3d_ws_kernel<C, HASB, BSYM, ACCUM> : kron3d_kernel<C, HASB, BSYM, ACCUM>;
          if constexpr (!ACCUM && C::NXW == 0)
            if (plan.dot_cursor >= 0)
              kern = kron3d_kernel<C, HASB, BSYM, false, true>;
          smem = smem_bytes<C, HASB>();
        }
      GDM_REQUIRE(smem <= 227 * 1024, GDM_ERR_INTERNAL, "fused kernel configuration exceeds the shared memory of an SM");
      static std::map<const void *, bool> attr_set;
      if (!attr_set[(const void *)kern])
        {
          GDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr_set[(const void *)kern] = true;
        }
      if (a.cz1 <= a.cz0)
        return;
      const int n_chunks = (a.cz1 - a.cz0 + a.lz - 1) / a.lz;
      const int grid     = (C::V5 || C::V7) ? v5_grid : plan.tiles_x * plan.tiles_y * n_chunks;
      if (grid <= 0)
        return;
      if (plan.dot_cursor >= 0)
        {
          GDM_REQUIRE(!ACCUM && (size_t)(plan.dot_cursor + grid) <= plan.dot_cap, GDM_ERR_INTERNAL, "fused dot: partial buffer too small");
          a.dot_src      = plan.dot_src;
          a.dot_partials = plan.d_dot + plan.dot_cursor;
          plan.dot_cursor += grid;
        }
      kern<<<grid, C::THREADS, smem, plan.use_comm_stream ? ctx.comm_stream : ctx.stream>>>(map, a);
      ctx.launches++;
      GDM_CUDA_CHECK(cudaGetLastError());
    }

    template <class C>
    const CUtensorMap &get_map(Operator &op, FusedPlan &plan, const double *src)
    {
      auto it = plan.maps.find(src);
      if (it != plan.maps.end())
        return it->second;
      if (plan.maps.size() > 64)
        plan.maps.clear();
      const Layout &L = op.sys->L;
      CUtensorMap   m;
      cuuint64_t    dims[3]    = {(cuuint64_t)L.ln[0], (cuuint64_t)L.ln[1], (cuuint64_t)L.ln[2]};
      cuuint64_t    strides[2] = {(cuuint64_t)L.pitch * 8, (cuuint64_t)L.plane * 8};
      cuuint32_t    box[3]     = {(cuuint32_t)C::PIN, (cuuint32_t)C::NR, 1};
      cuuint32_t    estr[3]    = {1, 1, 1};
      const CUresult rc = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(src), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      GDM_REQUIRE(rc == CUDA_SUCCESS, GDM_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
      return plan.maps.emplace(src, m).first->second;
    }
  } // namespace

  void fused_partition_host(bool aligned, int tiles_x, int tiles_y, int z0, int z1, int slots, int p, std::vector<int> &seg_ptr,
                            std::vector<int> &segs4)
  {
    GDM_REQUIRE(tiles_x > 0 && tiles_y > 0 && z1 >= z0 && slots > 0 && p > 0, GDM_ERR_INVALID, "invalid partition request");
    const PartitionPlan pp = make_partition(aligned, tiles_x * tiles_y, tiles_x, z0, z1, slots, p, 0);
    seg_ptr                = pp.ptr;
    segs4.clear();
    for (const int4 &s : pp.segs)
      {
        segs4.push_back(s.x);
        segs4.push_back(s.y);
        segs4.push_back(s.z);
        segs4.push_back(s.w);
      }
  }

  bool fused_supported(const Operator &op)
  {
    const Layout &L = op.sys->L;
    if (L.dim != 3 || L.nc != 1)
      return false;
    if (!(L.p == 1 || L.p == 3 || L.p == 5))
      return false;
    const char *env = std::getenv("GDM_DISABLE_FUSED");
    if (env && env[0] == '1')
      return false;
    bool any_periodic = false;
    for (int d = 0; d < 3; ++d)
      {
        if (L.N[d] < 2 * L.p + 2)
          return false;
        any_periodic |= op.periodic[d];
      }
    if (L.own1 <= L.own0)
      return false;
    if (any_periodic) // only the persistent kernel handles periodic directions (fold . A . duplicate)
      {
        const char *fam = std::getenv("GDM_FUSED_FAMILY");
        return (!fam || fam[0] == '8') && !std::getenv("GDM_FUSED_CFG") && pers_supported(op);
      }
    return true;
  }

  void fused_plan_create(Operator &op)
  {
    const Layout &L    = op.sys->L;
    Context      &ctx  = *op.sys->ctx;
    auto         *plan = new FusedPlan;
    op.fused           = plan;
    const int P = L.p, W = 2 * P + 1;
    // output window: Dirichlet faces are written by the constrained-row kernel
    plan->cx0 = op.dirichlet[0][0] ? 1 : 0;
    plan->cx1 = L.nn[0] - (op.dirichlet[0][1] ? 1 : 0);
    plan->cy0 = op.dirichlet[1][0] ? 1 : 0;
    plan->cy1 = L.nn[1] - (op.dirichlet[1][1] ? 1 : 0);
    int z0 = L.own0, z1 = L.own1; // global
    if (op.dirichlet[2][0])
      z0 = std::max(z0, 1);
    if (op.dirichlet[2][1])
      z1 = std::min(z1, L.nn[2] - 1);
    plan->cz0 = z0 - L.loc0;
    plan->cz1 = std::max(z1 - L.loc0, plan->cz0);
    {
      // default family: the persistent ramp-free kernel; GDM_FUSED_FAMILY=3|4 selects the round-1 tile kernels
      const char *fam = std::getenv("GDM_FUSED_FAMILY");
      if ((!fam || fam[0] == '8') && !std::getenv("GDM_FUSED_CFG") && pers_supported(op))
        plan->pers = pers_plan_create(op);
      GDM_REQUIRE(plan->pers || !(op.periodic[0] || op.periodic[1] || op.periodic[2]), GDM_ERR_NOT_IMPLEMENTED,
                  "fused kernel: periodic directions need the persistent kernel");
    }
    int tx = 32, ty = 32, min_blocks = 2;
    plan->cfg = default_config(P);
    with_config(plan->cfg, [&](auto c) {
      using C    = decltype(c);
      tx         = C::TX;
      ty         = C::TY;
      min_blocks = C::MINB;
    });
    // TMA box starts must be 16-byte aligned: keep (x0 - P) even by starting one column early if needed
    plan->xorg    = plan->cx0 - ((plan->cx0 - P) & 1);
    plan->tiles_x = (plan->cx1 - plan->xorg + tx - 1) / tx;
    plan->tiles_y = (plan->cy1 - plan->cy0 + ty - 1) / ty;
    const int nz    = plan->cz1 - plan->cz0;
    const int tiles = std::max(1, plan->tiles_x * plan->tiles_y);
    const int slots = ctx.sm_count * min_blocks;
    // z chunking: start from "one CTA per slot"; large problems are refined by a timed search below
    int       nch   = std::max(1, slots / tiles);
    int       lz    = (nz + nch - 1) / std::max(nch, 1);
    const int min_lz = 4 * P;
    if (lz < min_lz)
      lz = std::min(nz, min_lz);
    bool lz_forced = false;
    if (const char *env = std::getenv("GDM_FUSED_LZ"))
      {
        lz        = std::max(1, atoi(env));
        lz_forced = true;
      }
    lz             = std::max(lz, 1);
    plan->lz       = lz;
    plan->n_chunks = (nz + lz - 1) / lz;
    plan->tune     = !plan->pers && !lz_forced && (int64_t)nz * (plan->cx1 - plan->cx0) * (plan->cy1 - plan->cy0) > (int64_t)(1 << 21);
    // scatter rows: zs[k][j] = scale * T_z[k - P + j][2P - j]
    std::vector<double> zsA((size_t)L.ln[2] * W, 0.0), zsB((size_t)L.ln[2] * W, 0.0);
    for (int k = 0; k < L.ln[2]; ++k)
      for (int j = 0; j < W; ++j)
        {
          const int r = k - P + j;
          if (r < 0 || r >= L.ln[2])
            continue;
          zsA[(size_t)k * W + j] = op.desc.scale * op.hA[2][(size_t)r * W + (2 * P - j)];
          if (op.has_B)
            zsB[(size_t)k * W + j] = op.desc.scale * op.hB[2][(size_t)r * W + (2 * P - j)];
        }
    GDM_CUDA_CHECK(cudaMalloc(&plan->d_zsA, zsA.size() * sizeof(double)));
    GDM_CUDA_CHECK(cudaMalloc(&plan->d_zsB, zsB.size() * sizeof(double)));
    GDM_CUDA_CHECK(cudaMemcpy(plan->d_zsA, zsA.data(), zsA.size() * sizeof(double), cudaMemcpyHostToDevice));
    GDM_CUDA_CHECK(cudaMemcpy(plan->d_zsB, zsB.data(), zsB.size() * sizeof(double), cudaMemcpyHostToDevice));

    // ---- v4 kernels: effective B tables (tap split), plane-class scatter table
    bool is_v4 = false;
    with_config(plan->cfg, [&](auto c) { is_v4 = decltype(c)::V4; });
    if (!is_v4)
      return;
    // Toeplitz z planes [kz_lo, kz_hi) (same rule as fill_interior) and one local interior row of direction 2
    int local_interior = -1;
    for (int r = 0; r < L.ln[2]; ++r)
      if (r + L.loc0 > P && r + L.loc0 < L.N[2] - P)
        {
          local_interior = r;
          break;
        }
    plan->kz_lo = plan->kz_hi = 0;
    if (L.N[2] >= 4 * P + 2 && local_interior >= 0)
      {
        plan->kz_lo = std::max(0, 2 * P + 1 - L.loc0);
        plan->kz_hi = std::max(plan->kz_lo, std::min(L.ln[2], L.N[2] - 2 * P - L.loc0));
      }
    const int zrows = 2 * W;
    GDM_REQUIRE(plan->kz_lo + (L.ln[2] - plan->kz_hi) <= zrows, GDM_ERR_INTERNAL, "fused v4: too many non-Toeplitz planes");
    for (int d = 0; d < 3; ++d)
      plan->hBe[d] = op.hB[d];
    plan->rsplit = false;
    plan->sigma  = 0.0;
    const char *env_split = std::getenv("GDM_FUSED_RSPLIT");
    if (op.has_B && op.b_symmetry > 0 && !(env_split && env_split[0] == '0'))
      {
        // K_d = alpha_d M_d + R_d with alpha_d = (outer tap of K_d) / (outer tap of M_d): R_d has zero outer taps on
        // Toeplitz rows; the identity holds row by row, so the one-sided and masked rows need no special treatment
        double alpha[3] = {0, 0, 0};
        bool   ok       = true;
        for (int d = 0; d < 3 && ok; ++d)
          {
            const int row = (d == 2) ? local_interior : P + 1;
            if (row < 0 || L.N[d] < 2 * P + 2)
              {
                ok = false;
                break;
              }
            const double m = op.hA[d][(size_t)row * W + 2 * P], k = op.hB[d][(size_t)row * W + 2 * P];
            if (m == 0.0 || op.hA[d][(size_t)row * W] != m || op.hB[d][(size_t)row * W] != k)
              ok = false;
            alpha[d] = ok ? k / m : 0.0;
          }
        if (ok)
          {
            for (int d = 0; d < 3; ++d)
              {
                const int rows = (int)(op.hB[d].size() / W);
                for (int r = 0; r < rows; ++r)
                  {
                    for (int t = 0; t < W; ++t)
                      plan->hBe[d][(size_t)r * W + t] = op.hB[d][(size_t)r * W + t] - alpha[d] * op.hA[d][(size_t)r * W + t];
                    const int gr = r + (d == 2 ? L.loc0 : 0); // global row
                    if (gr > P && gr < L.N[d] - P)
                      plan->hBe[d][(size_t)r * W] = plan->hBe[d][(size_t)r * W + 2 * P] = 0.0;
                  }
              }
            plan->rsplit = true;
            plan->sigma  = alpha[0] + alpha[1] + alpha[2];
          }
      }
    if (op.has_B)
      for (int d = 0; d < 2; ++d)
        {
          GDM_CUDA_CHECK(cudaMalloc(&plan->d_Be[d], plan->hBe[d].size() * sizeof(double)));
          GDM_CUDA_CHECK(
            cudaMemcpy(plan->d_Be[d], plan->hBe[d].data(), plan->hBe[d].size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    const int           wz = W + 1;
    std::vector<double> zt((size_t)zrows * 2 * wz, 0.0);
    for (int c = 0; c < zrows; ++c)
      {
        const int k = (c < plan->kz_lo) ? c : plan->kz_hi + (c - plan->kz_lo);
        if (k < 0 || k >= L.ln[2])
          continue;
        for (int j = 0; j < W; ++j)
          {
            const int r = k - P + j;
            if (r < 0 || r >= L.ln[2])
              continue;
            zt[(size_t)(c * 2 + 0) * wz + j] = op.desc.scale * op.hA[2][(size_t)r * W + (2 * P - j)];
            if (op.has_B)
              zt[(size_t)(c * 2 + 1) * wz + j] = op.desc.scale * plan->hBe[2][(size_t)r * W + (2 * P - j)];
          }
      }
    GDM_CUDA_CHECK(cudaMalloc(&plan->d_zt, zt.size() * sizeof(double)));
    GDM_CUDA_CHECK(cudaMemcpy(plan->d_zt, zt.data(), zt.size() * sizeof(double), cudaMemcpyHostToDevice));
  }

  void fused_plan_destroy(Operator &op)
  {
    delete static_cast<FusedPlan *>(op.fused);
    op.fused = nullptr;
  }

  // one launch of the persistent kernel over the current output window of the plan
  static void dispatch_pers(Operator &op, FusedPlan &plan, double *dst, const double *src, bool accumulate)
  {
    Context  &ctx = *op.sys->ctx;
    const int z0  = (plan.wz0 >= 0) ? plan.wz0 : plan.cz0;
    const int z1  = (plan.wz0 >= 0) ? plan.wz1 : plan.cz1;
    double   *dp  = nullptr;
    if (plan.dot_cursor >= 0)
      {
        GDM_REQUIRE(!accumulate && (size_t)(plan.dot_cursor + pers_max_grid(op, plan.pers)) <= plan.dot_cap, GDM_ERR_INTERNAL,
                    "fused dot: partial buffer too small");
        dp = plan.d_dot + plan.dot_cursor;
      }
    const int grid = pers_launch(op, plan.pers, dst, src, accumulate, z0, z1, plan.use_comm_stream ? ctx.comm_stream : ctx.stream,
                                 dp ? plan.dot_src : nullptr, dp, plan.slots_limit);
    if (dp)
      plan.dot_cursor += grid;
  }

  template <class C>
  static void dispatch(Operator &op, FusedPlan &plan, double *dst, const double *src, bool accumulate)
  {
    if (plan.pers)
      {
        dispatch_pers(op, plan, dst, src, accumulate);
        return;
      }
    const CUtensorMap &map = get_map<C>(op, plan, src);
    if (!op.has_B)
      accumulate ? launch_variant<C, false, +1, true>(op, plan, map, dst) : launch_variant<C, false, +1, false>(op, plan, map, dst);
    else if (op.b_symmetry > 0)
      accumulate ? launch_variant<C, true, +1, true>(op, plan, map, dst) : launch_variant<C, true, +1, false>(op, plan, map, dst);
    else
      accumulate ? launch_variant<C, true, -1, true>(op, plan, map, dst) : launch_variant<C, true, -1, false>(op, plan, map, dst);
  }

  bool fused_supports_dot(const Operator &op)
  {
    if (!op.fused)
      return false;
    const FusedPlan &plan = *static_cast<const FusedPlan *>(op.fused);
    if (plan.pers)
      return true;
    bool             ok   = false;
    // the store epilogue with the dot product exists in the v3 tile kernel and in v4
    with_config(plan.cfg, [&](auto c) {
      using C = decltype(c);
      ok      = (!C::V4 && C::NXW == 0) || (C::V4 && !C::V5 && !C::V6 && !C::V7);
    });
    return ok;
  }

  void fused_apply(Operator &op, double *dst, const double *src, bool accumulate, bool exchange_ghosts, int dot_slot)
  {
    FusedPlan    &plan = *static_cast<FusedPlan *>(op.fused);
    Context      &ctx  = *op.sys->ctx;
    const Layout &L    = op.sys->L;
    bool is_v5 = false;
    with_config(plan.cfg, [&](auto c) { is_v5 = decltype(c)::V5 || decltype(c)::V7; });
    if (is_v5)
      plan.tune = false; // v5 partitions the work statically, there is no chunk length to tune
    if (plan.tune)
      {
        // FFTW-style plan refinement: time a few chunk counts around one/two/three CTAs per slot on
        // scratch output (same input), keep the fastest.  Runs once per operator.
        plan.tune = false;
        ctx.ensure_scratch((size_t)L.size);
        double     *tmp = ctx.scratch[0];
        const int   nz  = plan.cz1 - plan.cz0;
        const int   P   = L.p;
        int         mb  = 2;
        with_config(plan.cfg, [&](auto c) { mb = decltype(c)::MINB; });
        const int    tiles = std::max(1, plan.tiles_x * plan.tiles_y);
        const double base  = (double)ctx.sm_count * mb / tiles;
        cudaEvent_t  e0, e1;
        GDM_CUDA_CHECK(cudaEventCreate(&e0));
        GDM_CUDA_CHECK(cudaEventCreate(&e1));
        int   best_lz = plan.lz;
        float best_ms = 1e30f;
        int   last_nch = -1;
        for (double f : {1.0, 1.25, 1.5, 1.75, 2.0, 2.5, 3.0, 4.0})
          {
            const int nch = std::max(1, (int)(base * f));
            int       lz  = std::max((nz + nch - 1) / nch, std::min(nz, 2 * P + 2));
            const int n   = (nz + lz - 1) / lz;
            if (n == last_nch)
              continue;
            last_nch      = n;
            plan.lz       = lz;
            plan.n_chunks = n;
            float ms = 1e30f;
            for (int rep = 0; rep < 3; ++rep)
              {
                GDM_CUDA_CHECK(cudaEventRecord(e0, ctx.stream));
                with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, tmp, src, false); });
                GDM_CUDA_CHECK(cudaEventRecord(e1, ctx.stream));
                GDM_CUDA_CHECK(cudaEventSynchronize(e1));
                float t;
                GDM_CUDA_CHECK(cudaEventElapsedTime(&t, e0, e1));
                if (rep > 0)
                  ms = std::min(ms, t);
              }
            if (ms < best_ms)
              {
                best_ms = ms;
                best_lz = lz;
              }
          }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        plan.lz       = best_lz;
        plan.n_chunks = (nz + best_lz - 1) / best_lz;
        if (std::getenv("GDM_FUSED_VERBOSE"))
          fprintf(stderr, "[gdm] fused plan: cfg %d, tiles %d x %d, lz %d (%d chunks), %.3f ms\n", plan.cfg, plan.tiles_x,
                  plan.tiles_y, plan.lz, plan.n_chunks, best_ms);
      }
    const int P = L.p;
    const bool want_dot = dot_slot >= 0;
    // periodic directions (persistent kernel): C^T A C x = fold(A(dup x)); src is patched in place and restored
    const bool periodic = plan.pers && pers_has_periodic(plan.pers);
    if (periodic)
      {
        GDM_REQUIRE(!accumulate, GDM_ERR_INTERNAL, "fused periodic apply cannot accumulate (use a temporary)");
        pers_periodic_pre(op, plan.pers, const_cast<double *>(src), ctx.stream);
      }
    double    *face_partials = nullptr; // where the face kernel puts its partial sums (set below)
    if (want_dot)
      {
        GDM_REQUIRE(!accumulate && fused_supports_dot(op), GDM_ERR_INTERNAL, "fused dot product not available for this configuration");
        // partial sums: one per CTA of the (up to three) tile launches, then one per block of the face kernel
        const int    tiles   = std::max(1, plan.tiles_x * plan.tiles_y);
        const int    nzw     = std::max(1, plan.cz1 - plan.cz0);
        const int    lz_min  = std::max(1, std::min(plan.lz, P));
        size_t       need    = (size_t)tiles * (nzw / lz_min + 4) + (size_t)constrained_rows_max_blocks(L) + 64;
        if (plan.pers)
          need = (size_t)3 * pers_max_grid(op, plan.pers) + (size_t)constrained_rows_max_blocks(L) + 64;
        if (need > plan.dot_cap)
          {
            GDM_CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(plan.d_dot);
            plan.d_dot = nullptr;
            GDM_CUDA_CHECK(cudaMalloc(&plan.d_dot, need * sizeof(double)));
            plan.dot_cap = need;
          }
        plan.dot_src    = src;
        plan.dot_cursor = 0;
        // the face kernel may run concurrently with the tile kernels: give it the tail of the buffer
        face_partials = plan.d_dot + (plan.dot_cap - (size_t)constrained_rows_max_blocks(L));
      }
    int  face_blocks = 0;
    auto finish_dot  = [&]() {
      if (!want_dot)
        return;
      // compact: [tile partials | face partials] are summed separately in a fixed order into the slot
      const int n_tile = plan.dot_cursor;
      plan.dot_cursor  = -1;
      plan.dot_src     = nullptr;
      if (face_blocks > 0) // move the face partials behind the tile partials (device-side copy, same stream)
        GDM_CUDA_CHECK(cudaMemcpyAsync(plan.d_dot + n_tile, face_partials, (size_t)face_blocks * sizeof(double),
                                       cudaMemcpyDeviceToDevice, ctx.stream));
      blas_sum_partials(ctx, plan.d_dot, n_tile + face_blocks, dot_slot);
    };
    if (L.n_ranks > 1 && exchange_ghosts)
      {
        // overlap the ghost import (NCCL on the comm stream) with the planes that do not need it
        const int lo = L.own0 - L.loc0, hi = L.own1 - L.loc0; // owned planes (local indices)
        const bool thick = (hi - lo) > 4 * P;
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_a, ctx.stream));
        GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.comm_stream, ctx.ev_a, 0));
        comm_halo_exchange(ctx, L, const_cast<double *>(src), ctx.comm_stream);
        if (thick)
          {
            // slab faces: behind the ghost import on the comm stream, concurrent with the interior planes.  Persistent
            // kernel: the interior launch leaves a share of the CTA slots free (as many as the face planes' share of the
            // work, and they start late by the latency of the exchange), so the face launches find room the moment the
            // ghost planes arrive instead of queueing behind the interior CTAs.
            int face_slots = 0, all_slots = 0;
            if (plan.pers)
              {
                all_slots = pers_max_grid(op, plan.pers);
                // face work: 2 windows x 3P input planes of (hi - lo) + 2P; +15 % for the late start
                face_slots = std::max(4, (int)(1.15 * all_slots * (3.0 * P) / (double)((hi - lo) + 4 * P) + 0.5));
                if (const char *env = std::getenv("GDM_PERS_FACE_SLOTS"))
                  face_slots = std::max(1, atoi(env));
              }
            plan.use_comm_stream = true;
            plan.slots_limit     = face_slots;
            plan.wlz             = P;
            plan.wz0             = lo;
            plan.wz1             = lo + P;
            with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, accumulate); });
            plan.wz0 = hi - P;
            plan.wz1 = hi;
            with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, accumulate); });
            plan.use_comm_stream = false;
            GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_b, ctx.comm_stream));
            plan.wz0 = lo + P;
            plan.wz1 = hi - P;
            plan.wlz = 0;
            plan.slots_limit = plan.pers ? std::max(1, all_slots - 2 * face_slots) : 0;
            with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, accumulate); });
            plan.slots_limit = 0;
            plan.wz0 = plan.wz1 = -1;
            GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_b, 0));
          }
        else
          {
            GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_b, ctx.comm_stream));
            GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_b, 0));
            with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, accumulate); });
          }
      }
    else if (periodic)
      {
        with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, accumulate); });
      }
    else
      {
        // Dirichlet faces (skipped by the tiles) and deal.II's constrained diagonal: disjoint outputs, so the small
        // face kernel runs beside the tile kernel on the high-priority stream instead of after it (8 of 155 us)
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_a, ctx.stream));
        GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.comm_stream, ctx.ev_a, 0));
        std::swap(ctx.stream, ctx.comm_stream);
        face_blocks = launch_constrained_rows(ctx, L, op, dst, src, accumulate, -1, -1, face_partials);
        std::swap(ctx.stream, ctx.comm_stream);
        GDM_CUDA_CHECK(cudaEventRecord(ctx.ev_b, ctx.comm_stream));
        with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, accumulate); });
        GDM_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.ev_b, 0));
        finish_dot();
        return;
      }
    if (periodic)
      pers_periodic_post(op, plan.pers, dst, const_cast<double *>(src), ctx.stream);
    // Dirichlet faces (skipped by the tiles) and deal.II's constrained diagonal
    face_blocks = launch_constrained_rows(ctx, L, op, dst, src, accumulate, -1, -1, face_partials);
    finish_dot();
  }
  void fused_apply_window(Operator &op, double *dst, const double *src, int z0, int z1)
  {
    FusedPlan    &plan = *static_cast<FusedPlan *>(op.fused);
    Context      &ctx  = *op.sys->ctx;
    const Layout &L    = op.sys->L;
    plan.wz0           = z0;
    plan.wz1           = z1;
    plan.wlz           = std::max(1, std::min(plan.lz, z1 - z0));
    with_config(plan.cfg, [&](auto c) { dispatch<decltype(c)>(op, plan, dst, src, false); });
    plan.wz0 = plan.wz1 = -1;
    plan.wlz            = 0;
    launch_constrained_rows(ctx, L, op, dst, src, false, z0, z1);
  }
} // namespace gdm
