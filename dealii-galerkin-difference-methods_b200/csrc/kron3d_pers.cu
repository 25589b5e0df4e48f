// kron3d_pers -- persistent, ramp-free fused 3D tensor-product apply (sm_100a).
//
//     y = scale * ( B_x (x) A_y (x) A_z  +  A_x (x) B_y (x) A_z  +  A_x (x) A_y (x) B_z ) x      (stiffness / advection)
//     y = scale * ( A_x (x) A_y (x) A_z ) x                                                      (mass)
//
// Replaces SparseMatrix::vmult of the assembled operator (reference: tests/poisson_02_gdm.cc:215,
// prototypes/advection_01_gdm.cc:216, applications/wave/include/gdm/wave/problem.h:486-488).
//
// Same arithmetic as the round-1 tile kernels (TMA-staged (TX+2P) x (TY+2P) tile per plane, x pass -> transposed a/r
// fields in shared memory -> y pass -> z pass in registers in scatter form, tap split K_d = alpha_d M_d + R_d), but
// the work decomposition is new:
//   * one CTA per resident slot (grid = SMs x CTAs/SM), every CTA streams an equal share of the (tile, plane) work list;
//   * a CTA only processes the input planes of its own share: there is NO z ramp.  At a seam between two shares of one
//     tile column the upper share writes its first 2P emitted (partial) planes to a small scratch area and raises a
//     flag; the lower share adds them to its 2P running accumulators when it flushes, so every output plane is
//     written exactly once, by one CTA, in a fixed order (bitwise reproducible, no atomics on data);
//   * work is handed out by a ticket (atomicAdd) in descending share order, so a CTA only ever waits for a share that
//     was started before it: no deadlock even if the grid is not fully resident;
//   * the planes of consecutive jobs of a CTA form one sequence for the TMA pipeline (no refill bubble at a job switch);
//   * planes that cannot contribute (Dirichlet planes, zero fill outside the slab) are never staged.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <cstdio>
#include <map>
#include <tuple>
#include <type_traits>

#include "gdm_internal.h"

namespace gdm
{
  namespace
  {
    // ------------------------------------------------------------------ PTX helpers
    __device__ __forceinline__ uint32_t smem_u32(const void *p)
    {
      return (uint32_t)__cvta_generic_to_shared(p);
    }
    __device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    }
    __device__ __forceinline__ void mbar_fence_init()
    {
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
    {
      asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
    }
    __device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
    {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                   "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                   : "memory");
    }
    __device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
    {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
      return v;
    }
    __device__ __forceinline__ void st_release(unsigned *p, unsigned v)
    {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }

#ifdef GDM_PERS_WATCHDOG
    // Diagnostic build (GDM_BUILD_WATCHDOG=1, libgdm_b200_wd.so): every wait of the kernel is bounded; the first waits
    // that time out are recorded in mapped host memory, then all waits are abandoned so that the launch ends (with wrong
    // results) and the host can print what was being waited for.  Each CTA also publishes its progress, so a launch that
    // hangs somewhere else can still be inspected from the host while it hangs.
    struct WdHost
    {
      unsigned count, abort;
      int      rec[64][8];    // {code, cta, warp, q, share, seq, parity, aux}
      int      prog[2048][4]; // per CTA: {share, shares done, marker, seq}
    };
    __device__ __forceinline__ bool wd_try(uint32_t bar, unsigned parity)
    {
      unsigned ok;
      asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
      return ok != 0u;
    }
#endif

    // ------------------------------------------------------------------ configuration
    // SPLIT_ = 1: no CTA barrier in the plane loop; the x pass -> y/z pass hand-over runs through full/empty mbarriers
    // over three a/r buffers, so a warp only ever waits for what the others did one plane earlier (the warps that own an
    // extra x task or the TMA issue no longer hold everybody back at every plane)
    template <int P_, int RY_, int NW_, int RX_, int STAGES_, int MINB_, int SPLIT_ = 0>
    struct CfgP
    {
      static constexpr int  P = P_, TX = 32, RY = RY_, NW = NW_, RX = RX_, STAGES = STAGES_, MINB = MINB_;
      static constexpr bool SPLIT = SPLIT_ != 0;
      static constexpr int  NAB   = SPLIT ? 3 : 2; // a/r buffers
      static constexpr int  W       = 2 * P + 1;
      static constexpr int  TY      = RY * NW;
      static constexpr int  NR      = TY + 2 * P;                   // rows of the staged tile
      static constexpr int  PIN     = TX + 2 * P;                   // pitch of the staged tile (dense TMA box)
      static constexpr int  PY      = ((NR / 2) & 1) ? NR : NR + 2; // column pitch of the transposed a/r fields
      static constexpr int  THREADS = 32 * NW;
      static constexpr int  NXB     = TX / RX;
      // x pass: warp tasks (x block, group of 32 rows); the NR % 32 rows left over are block tasks packed densely
      // (lane <-> (x block, row)) into NLW more warp tasks; the task -> warp map rotates with the plane counter so that
      // no scheduler owns the extra tasks permanently
      static constexpr int  NG_FULL  = NR / 32;
      static constexpr int  REM      = NR % 32;
      static constexpr int  NWT_FULL = NXB * NG_FULL;
      static constexpr int  XPT      = (REM > 0) ? 32 / REM : 1;       // whole x blocks per left-over warp task
      static constexpr int  NLW      = (REM > 0) ? (NXB + XPT - 1) / XPT : 0;
      static constexpr int  XROT     = NW;                             // period of the task -> warp rotation (SPLIT)
      static constexpr int  NWT      = NWT_FULL + NLW;
      static constexpr int  NBT        = 2 * (P + 1);
      static constexpr int  WP         = 8 * ((W + 7) / 8);
      static constexpr int  TB_DOUBLES = 2 * 2 * NBT * WP;
      static constexpr int  ZROWS      = 2 * W; // non-Toeplitz plane classes: 2P+1 at either end
      static constexpr int  WZ         = W + 1;
      static constexpr int  ZT_DOUBLES = ZROWS * 2 * WZ;
      static constexpr int  STAGE_DOUBLES = (NR * PIN + 15) / 16 * 16;
      static_assert(TX % RX == 0 && RX % 2 == 0 && NR % 2 == 0 && RY % 2 == 0, "tile shape");
      static_assert(((PIN / 2) & 1) == 1 && ((PY / 2) & 1) == 1, "16-byte pitches must be odd for conflict-free LDS.128");
    };

    struct JobP // one contiguous run of input planes of one tile column
    {
      int x0, y0;   // first output column / row of the tile
      int k0, k1;   // input planes [k0, k1) (local plane indices)
      int seam_lo;  // >= 0: the first 2P emitted planes are partial sums owned by the job below: scratch slot
      int seam_hi;  // >= 0: the job above left its partial sums for the 2P planes this job flushes in this slot
      int pad0, pad1;
    };

    template <int P>
    struct ArgsP
    {
      double       *dst;
      int64_t       pitch, plane;
      int           cx0, cx1, cy0, cy1, cz0, cz1; // output window (local node indices)
      int           nx, ny;                       // cells per direction (boundary rows: <= P or >= N-P)
      int           nz_local;
      int           kz_lo, kz_hi;                 // input planes [kz_lo, kz_hi) only touch Toeplitz z rows
      int           n_shares;                     // number of shares (>= grid: CTAs take shares until none is left)
      unsigned      ticket_base, epoch;
      const double *tabAx, *tabBx, *tabAy, *tabBy; // row tables [node][2P+1]
      double        Ax[P + 1], Bx[P + 1], Ay[P + 1], By[P + 1]; // interior taps by distance
      double        Az[2 * P + 1], Bz[2 * P + 1];               // interior scatter row, scale folded in
      // one-sided rows of A/B in x and y (row class c: node c for c <= P, node N-P+(c-P-1) above): in the constant bank,
      // the row class is warp-uniform where they are used
      double        tbx[2][2 * (P + 1)][2 * P + 1], tby[2][2 * (P + 1)][2 * P + 1];
      double        sigma;                                      // tap split: sum_d alpha_d
      const double *zt;                                         // scatter rows of the non-Toeplitz plane classes [class][field][W+1]
      const JobP   *jobs;
      const int    *job_ptr; // jobs of share w: [job_ptr[w], job_ptr[w+1])
      unsigned     *ticket;  // work counter (monotone; ticket_base = its value before this launch)
      unsigned     *flags;   // [job]: == epoch once the job's first 2P partial planes are in scratch
      double       *scratch; // [job][2P][TY][TX]
      int          *error;   // set if a seam wait timed out
      long long    *trace;   // diagnostic (GDM_PERS_TRACE): per share {clock cycles, SM id}
      const unsigned short *xassign; // SPLIT: x tasks of warp w at plane q of tile t: xassign[(t * XROT + q % XROT) * NW + w] (bit mask)
      const double *dot_src; // fused dot product <src, A src>: the sum over the points of share w goes to dot_partials[w]
      double       *dot_partials;
#ifdef GDM_PERS_WATCHDOG
      WdHost       *wd;
#endif
    };

    template <class C, bool HASB>
    constexpr size_t smem_bytes_p()
    {
      return (size_t)(C::STAGES * C::STAGE_DOUBLES + C::NAB * (HASB ? 2 : 1) * C::TX * C::PY + C::ZT_DOUBLES + C::TB_DOUBLES) * sizeof(double) +
             (size_t)(C::STAGES + 2 * C::NAB) * sizeof(uint64_t) + 64 + 128;
    }

    // MODE 0: mass; MODE 1: B symmetric with the tap split (B tables hold R = B - alpha A); MODE 2: B antisymmetric.
    // The per-thread state of the kernel lives in one object whose members stay in registers (everything is inlined and
    // every array index is a compile-time constant after unrolling).
    template <class C, int MODE, bool ACCUM, bool DOT>
    struct PersWorker
    {
      static constexpr int  P = C::P, W = C::W, TX = C::TX, TY = C::TY, RY = C::RY, RX = C::RX, NR = C::NR, PIN = C::PIN, PY = C::PY;
      static constexpr bool HASB = MODE != 0, SYM = MODE == 1;
      static constexpr int  NF = HASB ? 2 : 1;
      static constexpr int  PB = SYM ? P - 1 : P; // outermost tap of the interior B rows in x and y
      static constexpr int  S  = C::STAGES;
      static constexpr int  AB_BUF   = NF * TX * PY; // [field][x][PY] (y contiguous)
      static constexpr int  OFF_AB   = S * C::STAGE_DOUBLES;
      static constexpr int  NAB      = C::NAB;
      static constexpr bool SPLIT    = C::SPLIT;
      static constexpr int  OFF_ZT   = OFF_AB + NAB * AB_BUF;
      static constexpr int  OFF_TB   = OFF_ZT + C::ZT_DOUBLES;
      static constexpr int  OFF_BAR  = OFF_TB + C::TB_DOUBLES;
      static constexpr int  OFF_MISC = OFF_BAR + S + 2 * NAB; // barriers: [S] TMA stages, [NAB] a/r full, [NAB] a/r empty
      static constexpr int  WP = C::WP, NBT = C::NBT, WZ = C::WZ;
      static constexpr unsigned STAGE_BYTES = NR * PIN * sizeof(double);
      static constexpr int  FULL_ROUNDS = C::NWT / C::NW, EXTRA = C::NWT % C::NW;
      static_assert(sizeof(JobP) == 2 * sizeof(int4), "job size");
      static_assert(EXTRA == 0 || C::NW > 1, "x pass task map");

      const ArgsP<P>    &g;
      const CUtensorMap *tmap;
      double            *smem;
      uint32_t           sb, bar0;
      int                tid, lane, warp, yz_off;
      // the thread that issues the TMA loads sits in a middle warp: the first and last warp own the one-sided y rows of
      // edge tiles, the x tasks of the one-sided columns are mapped away from all three (x_pass)
      static constexpr int ISSUE_WARP = C::NW / 2, ISSUE_TID = 32 * ISSUE_WARP;
      static constexpr int XSHIFT     = (C::NW >= 8) ? 3 : 0; // task t of a full round runs on warp (t + XSHIFT) % NW
      double             acc[RY][2 * P];
      double             dsum; // fused dot product (DOT): sum of src * (A src) over the points this thread stores
      // pipeline state: seq = number of planes whose x pass is done; everything else is derived from it
      int seq;
      int njobs, jb;
      // (the TMA issue cursor of thread 0 -- job index, job, next plane -- lives in shared memory: smisc[4..12))

      struct JobCtx // per-thread view of the current job
      {
        JobP    J;
        bool    more;      // another job follows in this share
        int     gy_first;  // first row of this thread
        int     nst;       // rows i < nst of this thread are stored
        double *out;       // output plane k - P at (gy_first, gx)
        double *sp;        // scratch slot position of this thread
        int     k_scr;     // planes k < k_scr emit into the scratch slot
        int     x0_next;   // x origin of the job that follows
        int     tile_next; // its tile index
      };

      // x tasks of this warp for plane sequence number q of tile `tile` (SPLIT; bit mask made by the host)
      __device__ __forceinline__ unsigned xm(const int tile, const int q) const
      {
        if constexpr (SPLIT)
          return (unsigned)__ldg(g.xassign + ((size_t)tile * C::XROT + q % C::XROT) * C::NW + warp);
        else
          return 0u;
      }

      __device__ __forceinline__ PersWorker(const ArgsP<P> &g_, const CUtensorMap *tmap_, double *smem_)
        : g(g_)
        , tmap(tmap_)
        , smem(smem_)
      {}

      __device__ __forceinline__ JobP load_job(const int j) const
      {
        const int4 *q = reinterpret_cast<const int4 *>(g.jobs + j);
        const int4  a = __ldg(q), b = __ldg(q + 1);
        JobP        J;
        J.x0 = a.x, J.y0 = a.y, J.k0 = a.z, J.k1 = a.w, J.seam_lo = b.x, J.seam_hi = b.y, J.pad0 = b.z, J.pad1 = 0; // pad0: tile index
        return J;
      }

      // thread 0: request the next plane of the share's plane sequence into stage st
      __device__ __forceinline__ void issue(const int st)
      {
        int *cur = reinterpret_cast<int *>(smem + OFF_MISC) + 4; // {ij, ik, x0 - P, y0 - P, k1}
        const int ij = cur[0];
        if (ij >= njobs)
          return;
        const int ik = cur[1];
        mbar_expect_tx(bar0 + 8 * st, STAGE_BYTES);
        tma_load_3d(sb + st * C::STAGE_DOUBLES * 8, tmap, bar0 + 8 * st, cur[2], cur[3], ik);
        if (ik + 1 >= cur[4])
          {
            cur[0] = ij + 1;
            if (ij + 1 < njobs)
              {
                const JobP Jn = load_job(jb + ij + 1);
                cur[1]        = Jn.k0;
                cur[2]        = Jn.x0 - P;
                cur[3]        = Jn.y0 - P;
                cur[4]        = Jn.k1;
              }
          }
        else
          cur[1] = ik + 1;
      }
      // every mbarrier wait of the kernel goes through here (code: 1 a/r full, 2 a/r empty, 3 TMA stage)
      __device__ __forceinline__ void wait(const uint32_t bar, const unsigned parity, [[maybe_unused]] const int code, [[maybe_unused]] const int q)
      {
#ifdef GDM_PERS_WATCHDOG
        const long long t0 = clock64();
        for (;;)
          {
            if (wd_try(bar, parity))
              return;
            // (the mapped host flag is only read once a wait is already very long: the timing of a healthy launch is
            // that of the product build)
            const long long dt = clock64() - t0;
            if (dt > (1ll << 22))
              {
                if (*(volatile unsigned *)&g.wd->abort != 0u)
                  return;
                if (dt > (1ll << 28))
                  {
                    if (lane == 0)
                      wd_record(code, q, (int)parity, 0);
                    return;
                  }
              }
          }
#else
        mbar_wait(bar, parity);
#endif
      }
#ifdef GDM_PERS_WATCHDOG
      int cur_share = -1, shares_done = 0;
      __device__ __forceinline__ void wd_record(const int code, const int q, const int parity, const int aux)
      {
        const unsigned slot = atomicAdd(&g.wd->count, 1u);
        if (slot < 64u)
          {
            int *r = g.wd->rec[slot];
            r[0] = code, r[1] = (int)blockIdx.x, r[2] = warp, r[3] = q, r[4] = cur_share, r[5] = seq, r[6] = parity, r[7] = aux;
          }
        __threadfence_system();
        *(volatile unsigned *)&g.wd->abort = 1u;
      }
      __device__ __forceinline__ void wd_progress(const int k1)
      {
        if (tid == 0 && blockIdx.x < 2048)
          {
            volatile int *pr = g.wd->prog[blockIdx.x];
            pr[0] = cur_share, pr[1] = shares_done, pr[2] = k1, pr[3] = seq;
          }
      }
#endif
      // stage / parity of the TMA ring and a/r buffers of plane sequence number q
      __device__ __forceinline__ int      stage_of(const int q) const { return q % S; }
      __device__ __forceinline__ unsigned parity_of(const int q) const { return (unsigned)(q / S) & 1u; }
      __device__ __forceinline__ int      ab_of(const int q) const { return OFF_AB + (q % NAB) * AB_BUF; }
      // split hand-over (SPLIT): full[b] = every warp finished its x tasks into buffer b (and its reads of the TMA stage),
      // empty[b] = every warp finished reading buffer b in its y/z pass
      __device__ __forceinline__ uint32_t bar_full(const int q) const { return bar0 + 8 * (S + q % NAB); }
      __device__ __forceinline__ uint32_t bar_empty(const int q) const { return bar0 + 8 * (S + NAB + q % NAB); }
      __device__ __forceinline__ void     warp_arrive(const uint32_t bar) const
      {
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      }
      // before the y/z pass of plane q: its a/r buffer is complete; thread 0 then refills the TMA stage plane q was read from
      __device__ __forceinline__ void yz_acquire(const int q)
      {
        if constexpr (SPLIT)
          {
            wait(bar_full(q), (unsigned)(q / NAB) & 1u, 1, q);
            if (tid == ISSUE_TID)
              issue(stage_of(q));
          }
      }
      __device__ __forceinline__ void yz_release(const int q)
      {
        if constexpr (SPLIT)
          warp_arrive(bar_empty(q));
      }

      // ---- x pass of one plane: staged tile at in_off -> a/r buffer at a_off (doubles); x0 = first output column.
      // FIX: the tile touches one-sided rows of A_x/B_x (recomputed from the row tables by the threads that own them).
      template <bool FIX>
      __device__ __forceinline__ void x_task(const int wt, const int in_off, const int a_off, const int x0)
      {
        const int      b_off = a_off + (NF - 1) * TX * PY;
        int            xb, r;
        if (wt < C::NWT_FULL)
          {
            xb = wt % C::NXB;
            r  = (wt / C::NXB) * 32 + lane;
          }
        else
          {
            const int sub = lane / C::REM; // (REM > 0 here)
            xb            = (wt - C::NWT_FULL) * C::XPT + sub;
            if (sub >= C::XPT || xb >= C::NXB)
              return;
            r = C::NG_FULL * 32 + lane % C::REM;
          }
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
            double    rbv = (HASB && SYM) ? g.Bx[0] * v[c] : 0.0;
#pragma unroll
            for (int d = 1; d <= P; ++d)
              {
                const double s = v[c - d] + v[c + d];
                ra             = fma(g.Ax[d], s, ra);
                if (HASB)
                  {
                    if (SYM)
                      {
                        if (d <= PB)
                          rbv = fma(g.Bx[d], s, rbv);
                      }
                    else
                      rbv = fma(g.Bx[d], v[c + d] - v[c - d], rbv);
                  }
              }
            a[j]  = ra;
            bb[j] = rbv;
          }
        if constexpr (FIX)
          {
            const int gx_first = x0 + xb * RX;
            if (gx_first <= P || gx_first + RX - 1 >= g.nx - P)
              {
#pragma unroll
                for (int j = 0; j < RX; ++j)
                  {
                    const int gxx = gx_first + j;
                    if ((gxx <= P || gxx >= g.nx - P) && gxx >= 0 && gxx <= g.nx)
                      {
                        const int     rc = (gxx <= P) ? gxx : gxx - (g.nx - P) + P + 1;
                        const double *ta = g.tbx[0][rc];
                        const double *tb = g.tbx[1][rc];
                        // two partial sums per row: the one-sided rows sit on the critical path of their warp
                        double ra = 0.0, rbv = 0.0, ra2 = 0.0, rbv2 = 0.0;
#pragma unroll
                        for (int t = 0; t < W; ++t)
                          {
                            if (t & 1)
                              {
                                ra2 = fma(ta[t], v[j + t], ra2);
                                if (HASB)
                                  rbv2 = fma(tb[t], v[j + t], rbv2);
                              }
                            else
                              {
                                ra = fma(ta[t], v[j + t], ra);
                                if (HASB)
                                  rbv = fma(tb[t], v[j + t], rbv);
                              }
                          }
                        a[j]  = ra + ra2;
                        bb[j] = rbv + rbv2;
                      }
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

      // Task -> warp map: full rounds warp + rd NW; the NWT % NW tasks left over rotate over warps 1..NW-1 with the plane
      // counter (warp 0 issues the TMA loads instead)
      // x pass of plane sequence number q (waits for its TMA stage)
      template <bool FIX>
      __device__ __forceinline__ void x_pass(const int q, const int x0, [[maybe_unused]] const unsigned xmask = 0u)
      {
        const int st = stage_of(q);
        if constexpr (SPLIT)
          if (q >= NAB)
            wait(bar_empty(q), (unsigned)(q / NAB - 1) & 1u, 2, q); // the y/z pass of plane q - NAB has released the buffer
        wait(bar0 + 8 * st, parity_of(q), 3, q);
        const int in_off = st * C::STAGE_DOUBLES, a_off = ab_of(q);
        if constexpr (SPLIT)
          {
            // the task -> warp map of this tile and plane comes from a table made by the host (x_assignment): the warps that
            // recompute one-sided rows in their y/z pass, or issue the TMA loads, get fewer x tasks, so that the warps of an
            // edge tile carry about the same load (without the plane-wide barrier only the average load of a warp counts)
            unsigned mask = xmask;
            while (mask != 0u)
              {
                const int t = __ffs(mask) - 1;
                mask &= mask - 1u;
                x_task<FIX>(t, in_off, a_off, x0);
              }
            warp_arrive(bar_full(q));
            return;
          }
        const int wrot = (warp >= XSHIFT) ? warp - XSHIFT : warp - XSHIFT + C::NW;
#pragma unroll
        for (int rd = 0; rd < FULL_ROUNDS; ++rd)
          x_task<FIX>(wrot + rd * C::NW, in_off, a_off, x0);
        if constexpr (EXTRA > 0)
          {
            // the tasks left over rotate over the warps other than the issuer's with the plane counter
            const int wi = (warp > ISSUE_WARP) ? warp - 1 : warp;
            const int e  = (wi + q) % (C::NW - 1);
            if (warp != ISSUE_WARP && e < EXTRA)
              x_task<FIX>(FULL_ROUNDS * C::NW + e, in_off, a_off, x0);
          }
        if constexpr (SPLIT)
          warp_arrive(bar_full(q));
      }

      // ---- y pass + z pass of input plane k from the a/r buffer at ab_off: res = emitted plane k - P.
      // FIX: the tile touches one-sided rows of A_y/B_y; TOEP: plane k only touches Toeplitz rows of A_z/B_z.
      // Row by row: the y result of a row goes straight into the z accumulators (few values live at a time).
      template <bool FIX, bool TOEP>
      __device__ __forceinline__ void yz_rows(const double (&zA)[W], const double (&zB)[W], const int ab_off, const int gy_first,
                                              double (&res)[RY])
      {
        constexpr bool OUTER = HASB && !(SYM && TOEP); // outer taps of the B chain present in z
        double         aw[RY + 2 * P], bw[RY + 2 * P];
        const double2 *pa = reinterpret_cast<const double2 *>(smem + ab_off + yz_off);
        const double2 *pb = reinterpret_cast<const double2 *>(smem + ab_off + (NF - 1) * TX * PY + yz_off);
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
        [[maybe_unused]] const bool y_edge = FIX && (gy_first <= P || gy_first + RY - 1 >= g.ny - P);
#pragma unroll
        for (int i = 0; i < RY; ++i)
          {
            const int c  = i + P;
            double    t1 = g.Ay[0] * aw[c], t2 = 0.0;
            if (HASB)
              {
                t2 = g.Ay[0] * bw[c];
                if (SYM)
                  t2 = fma(g.By[0], aw[c], t2);
              }
#pragma unroll
            for (int d = 1; d <= P; ++d)
              {
                const double sa = aw[c - d] + aw[c + d];
                t1              = fma(g.Ay[d], sa, t1);
                if (HASB)
                  {
                    const double sbv = bw[c - d] + bw[c + d];
                    t2               = fma(g.Ay[d], sbv, t2);
                    if (SYM)
                      {
                        if (d <= PB)
                          t2 = fma(g.By[d], sa, t2);
                      }
                    else
                      t2 = fma(g.By[d], aw[c + d] - aw[c - d], t2);
                  }
              }
            if constexpr (FIX)
              if (y_edge)
                {
                  const int gy = gy_first + i;
                  if ((gy <= P || gy >= g.ny - P) && gy <= g.ny)
                    {
                      const int     rc = (gy <= P) ? gy : gy - (g.ny - P) + P + 1;
                      const double *ta = g.tby[0][rc];
                      const double *tb = g.tby[1][rc];
                      t1 = 0.0;
                      t2 = 0.0;
                      double t1b = 0.0, t2b = 0.0, t2c = 0.0, t2d = 0.0; // independent partial sums (short chains)
#pragma unroll
                      for (int t = 0; t < W; ++t)
                        {
                          const double ca = ta[t];
                          if (t & 1)
                            t1b = fma(ca, aw[i + t], t1b);
                          else
                            t1 = fma(ca, aw[i + t], t1);
                          if (HASB)
                            {
                              if (t & 1)
                                {
                                  t2b = fma(ca, bw[i + t], t2b);
                                  t2d = fma(tb[t], aw[i + t], t2d);
                                }
                              else
                                {
                                  t2  = fma(ca, bw[i + t], t2);
                                  t2c = fma(tb[t], aw[i + t], t2c);
                                }
                            }
                        }
                      t1 += t1b;
                      if (HASB)
                        t2 = (t2 + t2b) + (t2c + t2d);
                    }
                }
            // z pass, scatter form: y_r += A_z[r][k] ua + R_z[r][k] t1 for the 2P+1 rows r around input plane k; the
            // accumulators shift by one plane through the FMAs themselves.  Tap split (SYM): the A chain runs on
            // ua = t2 + sigma t1 and the R chain has no outer taps on Toeplitz planes.
            double ua = t1;
            if (HASB)
              ua = SYM ? fma(g.sigma, t1, t2) : t2;
            double r0 = fma(zA[0], ua, acc[i][0]);
            if (OUTER)
              r0 = fma(zB[0], t1, r0);
            res[i] = r0;
#pragma unroll
            for (int j = 1; j < 2 * P; ++j)
              {
                double s = fma(zA[j], ua, acc[i][j]);
                if (HASB)
                  s = fma(zB[j], t1, s);
                acc[i][j - 1] = s;
              }
            double s = zA[2 * P] * ua;
            if (OUTER)
              s = fma(zB[2 * P], t1, s);
            acc[i][2 * P - 1] = s;
          }
      }

      template <bool FIX>
      __device__ __forceinline__ void yz_pass(const bool toep, const int k, const int ab_off, const int gy_first, double (&res)[RY])
      {
        if (toep)
          yz_rows<FIX, true>(g.Az, g.Bz, ab_off, gy_first, res);
        else
          {
            // plane class: planes below kz_lo by index, planes from kz_hi on after them
            const int kk = min(max(k, 0), g.nz_local - 1);
            const int c  = (kk < g.kz_lo) ? kk : g.kz_lo + (kk - g.kz_hi);
            double    zA[W], zB[W];
#pragma unroll
            for (int j = 0; j < W; ++j)
              {
                zA[j] = smem[OFF_ZT + (c * 2 + 0) * WZ + j];
                zB[j] = HASB ? smem[OFF_ZT + (c * 2 + 1) * WZ + j] : 0.0;
              }
            yz_rows<FIX, false>(zA, zB, ab_off, gy_first, res);
          }
      }

      // store RY values of an output plane (rows gy_first + i, column gx); FULL: every row and column of the tile is valid
      template <bool FULL>
      __device__ __forceinline__ void store_rows(double *o, const double (&val)[RY], const int nst)
      {
#pragma unroll
        for (int i = 0; i < RY; ++i)
          if (FULL || i < nst)
            {
              double *q = o + (int64_t)i * g.pitch;
              double  t = val[i];
              if constexpr (DOT)
                dsum = fma(__ldg(g.dot_src + (q - g.dst)), t, dsum);
              if (ACCUM)
                t += *q;
              *q = t;
            }
      }

      // general planes [ka, kb) of the job: every flag is evaluated per plane
      __device__ __forceinline__ void slow_planes(JobCtx &c, const int ka, const int kb)
      {
        for (int k = ka; k < kb; ++k)
          {
            const bool last = (k + 1 == c.J.k1);
            if (!last || c.more)
              x_pass<true>(seq, last ? c.x0_next : c.J.x0, xm(last ? c.tile_next : c.J.pad0, seq));
            double res[RY];
            yz_acquire(seq - 1);
            yz_pass<true>(k >= g.kz_lo && k < g.kz_hi, k, ab_of(seq - 1), c.gy_first, res);
            yz_release(seq - 1);
            if (k < c.k_scr)
              {
#pragma unroll
                for (int i = 0; i < RY; ++i)
                  __stcg(c.sp + i * TX, res[i]);
                c.sp += TY * TX;
              }
            else if (k - P >= g.cz0)
              store_rows<false>(c.out, res, c.nst);
            const bool signal = (k + 1 == c.k_scr && c.J.seam_lo >= 0); // the partial planes of the seam are complete
            if (!SPLIT || signal)
              __syncthreads();
            if (tid == ISSUE_TID)
              {
                if constexpr (!SPLIT)
                  issue(stage_of(seq));
                if (signal)
                  {
                    __threadfence();
                    st_release(g.flags + c.J.seam_lo, g.epoch);
                  }
              }
            ++seq;
            c.out += g.plane;
          }
      }

      // fast planes [ka, kb): Toeplitz in z, followed by another plane of this job; INNER tiles touch no one-sided row in
      // x or y and store every row and column: their plane body has no data-dependent branch.  SCR: the planes are the
      // first 2P of a job that hands its partial sums down (emitted into the scratch slot instead of dst).
      template <bool INNER, bool SCR>
      __device__ __forceinline__ void fast_planes(JobCtx &c, const int ka, const int kb)
      {
        for (int k = ka; k < kb; ++k)
          {
            x_pass<!INNER>(seq, c.J.x0, xm(c.J.pad0, seq));
            double res[RY];
            yz_acquire(seq - 1);
            yz_rows<!INNER, true>(g.Az, g.Bz, ab_of(seq - 1), c.gy_first, res);
            yz_release(seq - 1);
            if constexpr (SCR)
              {
#pragma unroll
                for (int i = 0; i < RY; ++i)
                  __stcg(c.sp + i * TX, res[i]);
                c.sp += TY * TX;
              }
            else
              store_rows<INNER>(c.out, res, c.nst);
            if constexpr (!SPLIT)
              {
                __syncthreads();
                if (tid == ISSUE_TID)
                  issue(stage_of(seq));
              }
            ++seq;
            c.out += g.plane;
          }
        if constexpr (SCR)
          {
            // the partial planes of the seam are complete: raise the flag of this job
            if constexpr (SPLIT)
              __syncthreads();
            if (tid == ISSUE_TID)
              {
                __threadfence();
                st_release(g.flags + c.J.seam_lo, g.epoch);
              }
          }
      }

      __device__ __forceinline__ void run()
      {
        sb   = smem_u32(smem);
        bar0 = sb + OFF_BAR * 8;
        tid  = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        const long long t_start = clock64();
        int            *smisc   = reinterpret_cast<int *>(smem + OFF_MISC);

        if (tid == 0)
          {
            for (int s = 0; s < S; ++s)
              mbar_init(bar0 + 8 * s, 1);
            for (int s = 0; s < 2 * NAB; ++s)
              mbar_init(bar0 + 8 * (S + s), C::NW);
            mbar_fence_init();
          }
        // one-sided rows of A/B in x and y (row class c: node c for c <= P, node N-P+(c-P-1) above)
        for (int e = tid; e < 2 * 2 * NBT * W; e += C::THREADS)
          {
            const int     t = e % W, c = (e / W) % NBT, f = (e / (W * NBT)) % 2, d = e / (W * NBT * 2);
            const int     n    = d ? g.ny : g.nx;
            const int     node = (c <= P) ? c : n - P + (c - P - 1);
            const double *tab  = d ? (f ? g.tabBy : g.tabAy) : (f ? g.tabBx : g.tabAx);
            smem[OFF_TB + ((d * 2 + f) * NBT + c) * WP + t] = (HASB || f == 0) ? __ldg(tab + node * W + t) : 0.0;
          }
        for (int e = tid; e < C::ZT_DOUBLES; e += C::THREADS)
          smem[OFF_ZT + e] = __ldg(g.zt + e);
        // y/z pass ownership: lane -> x, warp -> RY consecutive rows
        yz_off = lane * PY + warp * RY; // start of this thread's window in an a/r buffer
#pragma unroll
        for (int i = 0; i < RY; ++i)
#pragma unroll
          for (int j = 0; j < 2 * P; ++j)
            acc[i][j] = 0.0;
        dsum = 0.0;
        seq  = 0;
        // ---- shares are handed out by a ticket in DESCENDING order: a share only ever waits (at a seam) for a share with a
        // larger index, i.e. one that was started before it -- no deadlock even when there are more shares than resident
        // CTAs.  With the guided partition the shares that are handed out last are the short ones (no tail).
        // The plane sequence number seq runs on across shares (TMA stages and a/r buffers keep their mbarrier phases).
        for (;;)
          {
        __syncthreads(); // every warp is done with the previous share (and with the tables on the first pass)
        if (tid == 0)
          {
            const unsigned t = atomicAdd(g.ticket, 1u) - g.ticket_base;
            smisc[0]         = (t < (unsigned)g.n_shares) ? g.n_shares - 1 - (int)t : -1;
          }
        __syncthreads();
        const int share = smisc[0];
        if (share < 0)
          break;
        jb    = __ldg(g.job_ptr + share);
        njobs = __ldg(g.job_ptr + share + 1) - jb;
#ifdef GDM_PERS_WATCHDOG
        cur_share = share;
        wd_progress(-1);
#endif
        if (njobs <= 0)
          {
            if constexpr (DOT)
              if (tid == 0)
                g.dot_partials[share] = 0.0;
            continue;
          }
        JobP Jn = load_job(jb); // next job of the compute loop (every thread); the issuer keeps its own cursor
        if (tid == ISSUE_TID)
          {
            int *cur = smisc + 4;
            cur[0]   = 0;
            cur[1]   = Jn.k0;
            cur[2]   = Jn.x0 - P;
            cur[3]   = Jn.y0 - P;
            cur[4]   = Jn.k1;
            for (int s = 0; s < S; ++s)
              issue(stage_of(seq + s));
          }

        // prologue: x pass of the first plane of the first job
        x_pass<true>(seq, Jn.x0, xm(Jn.pad0, seq));
        if constexpr (!SPLIT)
          {
            __syncthreads();
            if (tid == ISSUE_TID)
              issue(stage_of(seq));
          }
        ++seq;

        for (int j = 0; j < njobs; ++j)
          {
            JobCtx c;
            c.J    = Jn;
            c.more = j + 1 < njobs;
            if (c.more)
              Jn = load_job(jb + j + 1);
            const int gx = c.J.x0 + lane;
            c.gy_first   = c.J.y0 + warp * RY;
            c.nst        = (gx >= g.cx0 && gx < g.cx1) ? (g.cy1 - c.gy_first) : 0;
            c.out        = g.dst + (int64_t)(c.J.k0 - P) * g.plane + (int64_t)c.gy_first * g.pitch + gx;
            c.sp         = g.scratch + ((int64_t)max(c.J.seam_lo, 0) * (2 * P)) * (TY * TX) + (warp * RY) * TX + lane;
            c.k_scr      = (c.J.seam_lo >= 0) ? c.J.k0 + 2 * P : c.J.k0;
            c.x0_next    = Jn.x0;
            c.tile_next  = Jn.pad0;
            const bool inner = c.J.x0 > P && c.J.x0 + TX - 1 < g.nx - P && c.J.y0 > P && c.J.y0 + TY - 1 < g.ny - P &&
                               c.J.x0 >= g.cx0 && c.J.x0 + TX <= g.cx1 && c.J.y0 >= g.cy0 && c.J.y0 + TY <= g.cy1;
            const int fa = max(max(c.J.k0, c.k_scr), max(g.cz0 + P, g.kz_lo));
            const int fb = min(c.J.k1 - 1, g.kz_hi);
            if (fb > fa)
              {
                if (c.J.seam_lo >= 0 && c.J.k0 >= g.kz_lo && fa == c.k_scr)
                  {
                    // the 2P partial planes of the seam run the fast body as well
                    if (inner)
                      fast_planes<true, true>(c, c.J.k0, c.k_scr);
                    else
                      fast_planes<false, true>(c, c.J.k0, c.k_scr);
                  }
                else
                  slow_planes(c, c.J.k0, fa);
                if (inner)
                  fast_planes<true, false>(c, fa, fb);
                else
                  fast_planes<false, false>(c, fa, fb);
                slow_planes(c, fb, c.J.k1);
              }
            else
              slow_planes(c, c.J.k0, c.J.k1);
            // ---- flush: output planes k1-P .. k1+P-1 hold the sums of this job's planes; the job above adds the rest
            const double *sq = nullptr;
            if (c.J.seam_hi >= 0)
              {
                if (tid == 0)
                  {
                    const unsigned *f  = g.flags + c.J.seam_hi;
                    const long long t0 = clock64();
                    while (ld_acquire(f) != g.epoch)
                      {
                        __nanosleep(64);
#ifdef GDM_PERS_WATCHDOG
                        if (clock64() - t0 > (1ll << 22))
                          {
                            if (*(volatile unsigned *)&g.wd->abort != 0u)
                              break;
                            if (clock64() - t0 > (1ll << 28))
                              {
                                wd_record(9, c.J.seam_hi, (int)ld_acquire(f), (int)g.epoch);
                                break;
                              }
                          }
#endif
                        if (clock64() - t0 > (1ll << 32))
                          {
                            *g.error = 1;
                            break;
                          }
                      }
                  }
                __syncthreads();
                sq = g.scratch + ((int64_t)c.J.seam_hi * (2 * P)) * (TY * TX) + (warp * RY) * TX + lane;
              }
#pragma unroll
            for (int jz = 0; jz < 2 * P; ++jz)
              {
                const int o = c.J.k1 - P + jz;
                double    val[RY];
#pragma unroll
                for (int i = 0; i < RY; ++i)
                  {
                    val[i]     = acc[i][jz];
                    acc[i][jz] = 0.0;
                  }
                if (sq != nullptr)
                  {
#pragma unroll
                    for (int i = 0; i < RY; ++i)
                      val[i] += __ldcg(sq + (int64_t)jz * (TY * TX) + i * TX);
                  }
                if (o >= g.cz0 && o < g.cz1)
                  store_rows<false>(c.out, val, c.nst);
                c.out += g.plane;
              }
          }
            // the last plane of the share had no x pass of a next plane: its sequence number stays free for the first
            // plane of the next share (every sequence number is used exactly once: the mbarrier phases depend on it)
            --seq;
#ifdef GDM_PERS_WATCHDOG
            ++shares_done;
            wd_progress(-2);
#endif
            // fused dot product: one partial sum per SHARE, reduced in a fixed order (which CTA runs a share depends on
            // the ticket order, the value of the share's sum does not: bitwise reproducible)
            if constexpr (DOT)
              {
                double *red = smem + OFF_MISC + 8;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                  dsum += __shfl_down_sync(0xffffffffu, dsum, o);
                if (lane == 0)
                  red[warp] = dsum;
                dsum = 0.0;
                __syncthreads(); // (the barrier at the top of the share loop separates this use of red[] from the next)
                if (tid == 0)
                  {
                    double t = 0.0;
                    for (int w = 0; w < C::NW; ++w)
                      t += red[w];
                    g.dot_partials[share] = t;
                  }
              }
          } // shares
        if (g.trace != nullptr && tid == 0)
          {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g.trace[2 * blockIdx.x]     = clock64() - t_start;
            g.trace[2 * blockIdx.x + 1] = (long long)smid;
          }
      }
    };

    template <class C, int MODE, bool ACCUM, bool DOT>
    __global__ void __launch_bounds__(C::THREADS, C::MINB) kron3d_pers_kernel(const __grid_constant__ CUtensorMap tmap, const ArgsP<C::P> g)
    {
      extern __shared__ __align__(128) double smem[];
      PersWorker<C, MODE, ACCUM, DOT> w(g, &tmap, smem);
      w.run();
    }

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

    struct PartitionP // work partition of one output-plane window
    {
      int       grid = 0, n_shares = 0, n_jobs = 0, n_seams = 0;
      JobP     *d_jobs = nullptr;
      int      *d_ptr  = nullptr;
      unsigned *d_sync = nullptr; // [0] ticket, [1] error, [2 ...] flags per job
      double   *d_scratch = nullptr;
      long long *d_trace  = nullptr;
      std::vector<JobP> h_jobs;
      std::vector<int>  h_ptr;
      unsigned  ticket_base = 0, epoch = 0;
      int64_t   max_planes = 0;
    };

    struct PersPlan
    {
      int      cfg = 0;
      int      tiles_x = 0, tiles_y = 0;
      int      cx0, cx1, cy0, cy1, cz0, cz1, xorg;
      int      in_lo = 0, in_hi = 0; // stored input planes that can contribute (local indices)
      int      mode = 0;
      double   sigma = 0.0;
      int      kz_lo = 0, kz_hi = 0;
      // effective tables: A (unfolded in periodic directions), B or R = B - alpha A (tap split)
      std::vector<double> hAe[3], hBe[3];
      double  *d_Ae[2] = {nullptr, nullptr};
      double  *d_Be[2] = {nullptr, nullptr};
      unsigned short *d_xassign = nullptr; // SPLIT configurations: x task masks [tile][rotation][warp]
      int      wx = 1300, wy = 1400, wxy = 1600; // per-mille plane cost of tiles with one-sided rows in x / y / both
      bool     tuned = false;
      bool     periodic[3] = {false, false, false};
      double  *d_save = nullptr; // periodic: saved values of the patched nodes
      int64_t  n_save = 0;
      double  *d_zt    = nullptr;
      std::map<std::tuple<int, int, int>, PartitionP> parts; // (first output plane, end, slots)
      std::map<const void *, CUtensorMap>        maps;
      std::map<std::pair<int, const void *>, bool> attr_set; // (device, kernel)
      ~PersPlan()
      {
        cudaFree(d_Ae[0]);
        cudaFree(d_Ae[1]);
        cudaFree(d_Be[0]);
        cudaFree(d_Be[1]);
        cudaFree(d_save);
        cudaFree(d_xassign);
        cudaFree(d_zt);
        for (auto &kv : parts)
          free_part(kv.second);
      }
      static void free_part(PartitionP &p)
      {
        cudaFree(p.d_jobs);
        cudaFree(p.d_ptr);
        cudaFree(p.d_sync);
        cudaFree(p.d_scratch);
        cudaFree(p.d_trace);
      }
    };

    //                    id   P  RY NW RX ST MINB
#define GDM_PERS_CONFIGS_CORE(X)      \
  X(801, CfgP<1, 4, 8, 4, 3, 2>)      \
  X(833, CfgP<3, 4, 6, 8, 4, 2, 1>)   \
  X(836, CfgP<5, 4, 8, 4, 4, 1, 1>)
#ifdef GDM_FUSED_EXPERIMENTAL
#define GDM_PERS_CONFIGS_EXP(X)       \
  X(800, CfgP<3, 4, 8, 4, 3, 2>)      \
  X(825, CfgP<3, 4, 6, 4, 4, 2>)      \
  X(831, CfgP<3, 4, 6, 4, 4, 2, 1>)
#else
#define GDM_PERS_CONFIGS_EXP(X)
#endif
#define GDM_PERS_CONFIGS(X) GDM_PERS_CONFIGS_CORE(X) GDM_PERS_CONFIGS_EXP(X)

    template <class F>
    void with_config(int id, F &&f)
    {
      switch (id)
        {
#define GDM_CASE(ID, ...) \
  case ID:                \
    f(__VA_ARGS__{});     \
    break;
          GDM_PERS_CONFIGS(GDM_CASE)
#undef GDM_CASE
          default:
            throw Error(GDM_ERR_INVALID, "persistent fused kernel configuration " + std::to_string(id) + " is not in this build");
        }
    }

    int default_config(int p)
    {
      // defaults by measurement on B200 (profiles/r2): p=1 barrier per plane, 8 warps; p=3 split hand-over, 6 warps x 168
      // registers, RX=8; p=5 split hand-over, one CTA of 8 warps x 255 registers per SM
      int id = (p == 1) ? 801 : (p == 3 ? 833 : 836);
      if (const char *env = std::getenv("GDM_PERS_CFG"))
        {
          const int e  = atoi(env);
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
  } // namespace

  // Work partition of the persistent kernel (host logic, also exported for the CPU tests).
  //   tiles_x x tiles_y tile columns, input planes [k0, k1), at most `slots` shares (one CTA each).  weights[t] is the
  //   cost of one plane of tile t (per mille of an interior tile; nullptr: all equal): tiles that touch one-sided rows
  //   are slower, so their columns are cut into shorter jobs.  No job is shorter than min_len planes unless it is a whole
  //   column.  aligned: every column is cut into equal-cost chunks from the bottom (neighbouring tiles of the same class
  //   stream the same planes at the same time: their halos meet in L2); the planes left above the last full chunk are
  //   swept tile-major over the spare shares.  Otherwise: one tile-major sweep.
  // Output: job_ptr (size grid+1) and jobs as 6 ints {tile x, tile y, k_begin, k_end, seam_lo, seam_hi}; seam ids are job
  // indices.  A job's upper neighbour always lies in a share with a larger index (the kernel's ticket order relies on it).
  // mode 2 (guided): every column is cut at the same planes into levels whose length shrinks from the top of the column to
  //   its bottom, length = planes left x tiles / (guide_k x slots), at least guide_min planes; one share per (tile, level),
  //   bottom level first.  The kernel hands shares out from the last to the first, so the CTAs start on the long top levels
  //   and finish on short ones: self-scheduling with a short tail, whatever the tiles cost.
  void pers_partition_host(int tiles_x, int tiles_y, int k0, int k1, int slots, int min_len, int mode, const int *weights,
                           std::vector<int> &job_ptr, std::vector<int> &jobs6, double guide_k, int guide_min)
  {
    const bool aligned = mode == 1;
    GDM_REQUIRE(tiles_x > 0 && tiles_y > 0 && k1 >= k0 && slots > 0 && min_len > 0, GDM_ERR_INVALID, "invalid partition request");
    struct Piece
    {
      int tile, a, b;
    };
    typedef std::vector<std::vector<Piece>> Shares;
    const int tiles = tiles_x * tiles_y;
    const int nz    = k1 - k0;
    job_ptr.assign(1, 0);
    jobs6.clear();
    if (nz == 0)
      return;
    auto wt = [&](int t) -> int64_t { return weights ? std::max(1, weights[t]) : 1000; };
    int64_t total = 0;
    for (int t = 0; t < tiles; ++t)
      total += wt(t) * nz;
    // tile-major sweep of planes [za[t], k1) of every tile over at most n shares of about equal cost; a cut is moved to
    // the nearest position that leaves no piece shorter than min_len
    auto sweep = [&](Shares &shares, const std::vector<int> &za, int n) {
      int64_t rem = 0, planes = 0;
      int     cols = 0, max_piece = 0;
      for (int t = 0; t < tiles; ++t)
        if (k1 > za[t])
          {
            rem += wt(t) * (k1 - za[t]);
            planes += k1 - za[t];
            max_piece = std::max(max_piece, k1 - za[t]);
            ++cols;
          }
      if (rem == 0 || n <= 0)
        return;
      // number of shares: at least 2 min_len planes each; columns shorter than 2 min_len are not cut
      int64_t G = std::max<int64_t>(1, std::min<int64_t>(n, planes / (2 * (int64_t)min_len)));
      if (max_piece < 2 * min_len)
        G = std::min<int64_t>(G, cols);
      const double target = (double)rem / (double)G;
      std::vector<Piece> cur;
      int64_t            closed = 0;
      double             done = 0.0; // cost swept so far
      for (int t = 0; t < tiles; ++t)
        {
          int z = za[t];
          while (z < k1)
            {
              const double w     = (double)wt(t);
              const double room  = target * (double)(closed + 1) - done; // cost left in the current share
              int          take  = (int)(room / w + 0.5);
              const int    avail = k1 - z;
              const bool   lastshare = closed + 1 >= G;
              if (lastshare || take >= avail)
                take = avail;
              else
                {
                  // valid cut positions inside this column piece: >= min_len from either end of [z0, k1) where z0 is the
                  // start of the piece being carved (z itself starts a new piece of this share)
                  if (take < min_len)
                    take = (2 * take < min_len && !cur.empty()) ? 0 : min_len;
                  if (avail - take < min_len)
                    take = (2 * (avail - take) < min_len || avail < 2 * min_len) ? avail : avail - min_len;
                  if (take > avail)
                    take = avail;
                  if (avail < 2 * min_len && take != 0)
                    take = avail;
                }
              if (take > 0)
                {
                  if (!cur.empty() && cur.back().tile == t && cur.back().b == z)
                    cur.back().b = z + take;
                  else
                    cur.push_back({t, z, z + take});
                  z += take;
                  done += w * take;
                }
              if (!lastshare && (take == 0 || done >= target * (double)(closed + 1) - 0.5 * w))
                {
                  if (!cur.empty())
                    {
                      shares.push_back(cur);
                      cur.clear();
                    }
                  ++closed;
                }
            }
        }
      if (!cur.empty())
        shares.push_back(cur);
    };
    // aligned chunks of cost T per column from the bottom, the rest swept over the spare shares
    auto build = [&](double T) {
      Shares           shares;
      std::vector<int> za(tiles, k0);
      if (T > 0.0)
        {
          std::vector<int> Lt(tiles), mt(tiles);
          int              mmax = 0;
          for (int t = 0; t < tiles; ++t)
            {
              Lt[t] = std::max(min_len, (int)(T / (double)wt(t) + 0.5));
              mt[t] = nz / Lt[t];
              if (mt[t] > 0 && nz - mt[t] * Lt[t] > 0 && nz - mt[t] * Lt[t] < min_len)
                --mt[t]; // keep the remainder a valid piece (it is swept)
              mmax = std::max(mmax, mt[t]);
            }
          for (int c = 0; c < mmax; ++c)
            for (int t = 0; t < tiles; ++t)
              if (c < mt[t])
                {
                  shares.push_back({{t, k0 + c * Lt[t], k0 + (c + 1) * Lt[t]}});
                  za[t] = k0 + (c + 1) * Lt[t];
                }
        }
      sweep(shares, za, std::max(1, slots - (int)shares.size()));
      return shares;
    };
    auto longest = [&](const Shares &shares) {
      int64_t mx = 0;
      for (auto &sh : shares)
        {
          int64_t n = 0;
          for (auto &p : sh)
            n += wt(p.tile) * (p.b - p.a);
          mx = std::max(mx, n);
        }
      return mx;
    };
    Shares shares;
    const bool cuttable = nz >= 4 * min_len && slots >= tiles;
    if (mode == 2)
      {
        std::vector<int> len; // level lengths from the top of the column down
        int              left = nz;
        const int        gmin = std::max(min_len, guide_min);
        while (left > 0)
          {
            int l = (int)((double)left * tiles / (std::max(0.1, guide_k) * slots) + 0.5);
            l     = std::max(l, gmin);
            if (l >= left)
              l = left;
            else if (left - l < gmin)
              {
                // a remainder too short for a level of its own goes to the top level (handed out first)
                const int rem = left - l;
                if (len.empty())
                  l += rem;
                else
                  len[0] += rem;
                left -= len.empty() ? 0 : rem;
              }
            len.push_back(l);
            left -= l;
          }
        int z = k0;
        for (int i = (int)len.size() - 1; i >= 0; --i) // bottom level first
          {
            for (int t = 0; t < tiles; ++t)
              shares.push_back({{t, z, z + len[i]}});
            z += len[i];
          }
      }
    else if (aligned && cuttable)
      {
        // search the chunk cost around the ideal share for the shortest longest share
        const double T0   = (double)total / (double)slots;
        int64_t      best = -1;
        for (int i = -6; i <= 12; ++i)
          {
            const double T = T0 * (1.0 + 0.01 * i);
            if (T / 1000.0 < (double)min_len)
              continue;
            Shares        cand = build(T);
            const int64_t mx   = longest(cand);
            if ((int)cand.size() <= slots && (best < 0 || mx < best))
              {
                best   = mx;
                shares = std::move(cand);
              }
          }
        if (best < 0)
          shares = build(0.0);
      }
    else
      shares = build(0.0);
    // jobs in share order; seams: the job that starts at plane b of tile t is the upper neighbour of the job ending at b
    std::map<std::pair<int, int>, int> starts; // (tile, first plane) -> job index
    int                                nj = 0;
    for (auto &sh : shares)
      for (auto &p : sh)
        starts[{p.tile, p.a}] = nj++;
    for (auto &sh : shares)
      {
        for (auto &p : sh)
          {
            const int self = starts[{p.tile, p.a}];
            int       hi   = -1;
            auto      it   = starts.find({p.tile, p.b});
            if (p.b < k1 && it != starts.end())
              hi = it->second;
            jobs6.push_back(p.tile % tiles_x);
            jobs6.push_back(p.tile / tiles_x);
            jobs6.push_back(p.a);
            jobs6.push_back(p.b);
            jobs6.push_back(p.a > k0 ? self : -1);
            jobs6.push_back(hi);
          }
        job_ptr.push_back((int)(jobs6.size() / 6));
      }
  }

  namespace
  {
    template <class C>
    PartitionP &get_partition(Operator &op, PersPlan &plan, int oz0, int oz1, int slots_limit)
    {
      const auto key = std::make_tuple(oz0, oz1, slots_limit);
      auto       it  = plan.parts.find(key);
      if (it != plan.parts.end())
        return it->second;
      if (plan.parts.size() > 256)
        {
          GDM_CUDA_CHECK(cudaDeviceSynchronize());
          for (auto &kv : plan.parts)
            PersPlan::free_part(kv.second);
          plan.parts.clear();
        }
      constexpr int P = C::P;
      Context      &ctx = *op.sys->ctx;
      const int     k0 = std::max(oz0 - P, plan.in_lo), k1 = std::min(oz1 + P, plan.in_hi);
      int           slots = ctx.sm_count * C::MINB;
      if (const char *env = std::getenv("GDM_PERS_SLOTS"))
        slots = std::max(1, atoi(env));
      if (slots_limit > 0)
        slots = std::min(slots, slots_limit);
      const int   min_len = 2 * P; // the shortest job that can hand its first 2P partial planes down
      const char *env_al  = std::getenv("GDM_PERS_ALIGNED");
      const bool  aligned = !(env_al && env_al[0] == '0');
      const char *env_L   = std::getenv("GDM_PERS_L");
      std::vector<int> ptr, j6;
      // cost of a plane per tile (per mille of an interior tile): tiles that touch one-sided rows in x / y or store
      // partial rows run the general plane body and wait for the warps that recompute the boundary rows
      // (measured on B200, profiles/r2/pers_trace_*.txt).  GDM_PERS_WEIGHTS="x,y,xy" overrides (per mille).
      // Defaults in PersPlan; pers_tune() replaces them at operator creation by the best of a few measured candidates.
      int wx = plan.wx, wy = plan.wy, wxy = plan.wxy;
      if (const char *env = std::getenv("GDM_PERS_WEIGHTS"))
        sscanf(env, "%d,%d,%d", &wx, &wy, &wxy);
      const Layout    &L = op.sys->L;
      std::vector<int> weights((size_t)plan.tiles_x * plan.tiles_y, 1000);
      for (int ty = 0; ty < plan.tiles_y; ++ty)
        for (int tx = 0; tx < plan.tiles_x; ++tx)
          {
            const int  x0 = plan.xorg + tx * C::TX, y0 = plan.cy0 + ty * C::TY;
            const bool xe = !(x0 > P && x0 + C::TX - 1 < L.N[0] - P && x0 >= plan.cx0 && x0 + C::TX <= plan.cx1);
            const bool ye = !(y0 > P && y0 + C::TY - 1 < L.N[1] - P && y0 >= plan.cy0 && y0 + C::TY <= plan.cy1);
            weights[(size_t)ty * plan.tiles_x + tx] = (xe && ye) ? wxy : (xe ? wx : (ye ? wy : 1000));
          }
      (void)env_L;
      // partition mode: one weighted share per CTA (static).  GDM_PERS_MODE=guided selects the self-scheduled level
      // partition (more shares than CTAs; the ticket order keeps it deadlock free).  Measured on B200 (profiles/r2/
      // sessions_k_to_t.md) it balances the CTAs but loses more at the share switches (pipeline refill, seam hand-over)
      // than it gains: 125-129 vs 133 GDoF/s; it stays a diagnostic switch.
      int         mode   = aligned ? 1 : 0;
      double      gk     = 1.0;
      int         gmin   = 8;
      if (const char *env = std::getenv("GDM_PERS_MODE"))
        mode = (env[0] == 'g') ? 2 : mode;
      if (const char *env = std::getenv("GDM_PERS_GUIDE"))
        sscanf(env, "%lf,%d", &gk, &gmin);
      pers_partition_host(plan.tiles_x, plan.tiles_y, k0, std::max(k0, k1), slots, min_len, mode, weights.data(), ptr, j6, gk, gmin);
      PartitionP part;
      part.n_shares = (int)ptr.size() - 1;
      part.grid     = std::min(part.n_shares, slots);
      part.n_jobs = (int)(j6.size() / 6);
      std::vector<JobP> jobs(part.n_jobs);
      for (int j = 0; j < part.n_jobs; ++j)
        {
          jobs[j].x0      = plan.xorg + j6[6 * j + 0] * C::TX;
          jobs[j].y0      = plan.cy0 + j6[6 * j + 1] * C::TY;
          jobs[j].k0      = j6[6 * j + 2];
          jobs[j].k1      = j6[6 * j + 3];
          jobs[j].seam_lo = j6[6 * j + 4];
          jobs[j].seam_hi = j6[6 * j + 5];
          jobs[j].pad0 = j6[6 * j + 1] * plan.tiles_x + j6[6 * j + 0]; // tile index (x task assignment table)
          jobs[j].pad1 = 0;
          if (jobs[j].seam_lo >= 0)
            {
              part.n_seams++;
              GDM_REQUIRE(jobs[j].k1 - jobs[j].k0 >= 2 * P, GDM_ERR_INTERNAL, "persistent partition: job shorter than a seam");
            }
        }
      for (int w = 0; w < part.n_shares; ++w)
        {
          int64_t n = 0;
          for (int j = ptr[w]; j < ptr[w + 1]; ++j)
            n += jobs[j].k1 - jobs[j].k0;
          part.max_planes = std::max(part.max_planes, n);
        }
      if (part.grid > 0)
        {
          GDM_CUDA_CHECK(cudaMalloc(&part.d_jobs, std::max<size_t>(1, jobs.size()) * sizeof(JobP)));
          GDM_CUDA_CHECK(cudaMalloc(&part.d_ptr, ptr.size() * sizeof(int)));
          GDM_CUDA_CHECK(cudaMemcpy(part.d_jobs, jobs.data(), jobs.size() * sizeof(JobP), cudaMemcpyHostToDevice));
          GDM_CUDA_CHECK(cudaMemcpy(part.d_ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
          GDM_CUDA_CHECK(cudaMalloc(&part.d_sync, (size_t)(2 + part.n_jobs) * sizeof(unsigned)));
          GDM_CUDA_CHECK(cudaMemset(part.d_sync, 0, (size_t)(2 + part.n_jobs) * sizeof(unsigned)));
          if (part.n_seams > 0) // slots are indexed by job: simple, 48 KB per job at p = 3
            GDM_CUDA_CHECK(cudaMalloc(&part.d_scratch, (size_t)part.n_jobs * 2 * P * C::TY * C::TX * sizeof(double)));
          if (std::getenv("GDM_PERS_TRACE"))
            {
              GDM_CUDA_CHECK(cudaMalloc(&part.d_trace, (size_t)2 * part.grid * sizeof(long long)));
              part.h_jobs = jobs;
              part.h_ptr  = ptr;
            }
          GDM_CUDA_CHECK(cudaDeviceSynchronize());
        }
      if (std::getenv("GDM_FUSED_VERBOSE"))
        fprintf(stderr, "[gdm] persistent partition: outputs [%d, %d), inputs [%d, %d), %d x %d tiles -> %d shares, %d jobs, %d seams, longest share %lld planes\n",
                oz0, oz1, k0, k1, plan.tiles_x, plan.tiles_y, part.n_shares, part.n_jobs, part.n_seams, (long long)part.max_planes);
      return plan.parts.emplace(key, part).first->second;
    }

    template <class C>
    const CUtensorMap &get_map(Operator &op, PersPlan &plan, const double *src)
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

    // x task -> warp assignment of every tile (SPLIT configurations): longest-processing-time greedy over a cost model in
    // units of one full x task.  Loads before the x tasks: the y/z pass of every warp, the one-sided y rows recomputed by
    // the warp that owns them, the TMA issue; an x task that contains one-sided columns costs more.  The interior pattern
    // rotates with the plane counter so that the tasks left over after a full round visit every warp.
    template <class C>
    void build_x_assignment(const Operator &op, PersPlan &plan)
    {
      if (!C::SPLIT || plan.d_xassign)
        return;
      constexpr int P = C::P, NW = C::NW, R = C::XROT;
      const Layout &L = op.sys->L;
      double c_yz = 4.0, c_yfix_row = 0.85, c_xfix = 1.5, c_issue = 0.3;
      if (const char *env = std::getenv("GDM_PERS_COSTS"))
        sscanf(env, "%lf,%lf,%lf,%lf", &c_yz, &c_yfix_row, &c_xfix, &c_issue);
      const int                   tiles = plan.tiles_x * plan.tiles_y;
      std::vector<unsigned short> tab((size_t)tiles * R * NW, 0);
      for (int ty = 0; ty < plan.tiles_y; ++ty)
        for (int tx = 0; tx < plan.tiles_x; ++tx)
          {
            const int x0 = plan.xorg + tx * C::TX, y0 = plan.cy0 + ty * C::TY;
            // task costs
            double tc[C::NWT];
            for (int t = 0; t < C::NWT; ++t)
              {
                int    xb0, xb1;
                double fill = 1.0;
                if (t < C::NWT_FULL)
                  xb0 = t % C::NXB, xb1 = xb0 + 1;
                else
                  {
                    xb0  = (t - C::NWT_FULL) * C::XPT;
                    xb1  = std::min(C::NXB, xb0 + C::XPT);
                    fill = std::max(0.25, (double)(xb1 - xb0) * C::REM / 32.0);
                  }
                bool fix = false;
                for (int xb = xb0; xb < xb1; ++xb)
                  {
                    const int gx = x0 + xb * C::RX;
                    fix |= (gx <= P || gx + C::RX - 1 >= L.N[0] - P);
                  }
                tc[t] = fill + (fix ? c_xfix : 0.0);
              }
            for (int r = 0; r < R; ++r)
              {
                double load[NW];
                for (int w = 0; w < NW; ++w)
                  {
                    const int gy = y0 + w * C::RY;
                    int       nfix = 0;
                    for (int i = 0; i < C::RY; ++i)
                      nfix += ((gy + i <= P || gy + i >= L.N[1] - P) && gy + i >= 0 && gy + i <= L.N[1]) ? 1 : 0;
                    load[w] = c_yz + c_yfix_row * nfix + (w == NW / 2 ? c_issue : 0.0);
                  }
                // tasks by decreasing cost (stable), ties among warps broken starting at warp r (rotation)
                int order[C::NWT];
                for (int t = 0; t < C::NWT; ++t)
                  order[t] = t;
                std::stable_sort(order, order + C::NWT, [&](int a, int b) { return tc[a] > tc[b]; });
                unsigned short mask[NW] = {};
                for (int k = 0; k < C::NWT; ++k)
                  {
                    int best = -1;
                    for (int i = 0; i < NW; ++i)
                      {
                        const int w = (r + i) % NW;
                        if (best < 0 || load[w] < load[best] - 1e-9)
                          best = w;
                      }
                    load[best] += tc[order[k]];
                    mask[best] |= (unsigned short)(1u << order[k]);
                  }
                for (int w = 0; w < NW; ++w)
                  tab[((size_t)(ty * plan.tiles_x + tx) * R + r) * NW + w] = mask[w];
              }
          }
      GDM_CUDA_CHECK(cudaMalloc(&plan.d_xassign, tab.size() * sizeof(unsigned short)));
      GDM_CUDA_CHECK(cudaMemcpy(plan.d_xassign, tab.data(), tab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    }

    template <class C, int MODE>
    int launch_p(Operator &op, PersPlan &plan, double *dst, const double *src, bool accumulate, int oz0, int oz1, cudaStream_t stream,
                 const double *dot_src, double *dot_partials, int slots_limit)
    {
      constexpr int P = C::P, W = C::W;
      Context      &ctx = *op.sys->ctx;
      const Layout &L   = op.sys->L;
      if (oz1 <= oz0)
        return 0;
      build_x_assignment<C>(op, plan);
      PartitionP &part = get_partition<C>(op, plan, oz0, oz1, slots_limit);
      if (part.grid <= 0)
        return 0;
      ArgsP<P> a;
      a.dst      = dst;
      a.pitch    = L.pitch;
      a.plane    = L.plane;
      a.cx0      = plan.cx0;
      a.cx1      = plan.cx1;
      a.cy0      = plan.cy0;
      a.cy1      = plan.cy1;
      a.cz0      = oz0;
      a.cz1      = oz1;
      a.nx       = L.N[0];
      a.ny       = L.N[1];
      a.nz_local = L.ln[2];
      a.kz_lo    = plan.kz_lo;
      a.kz_hi    = plan.kz_hi;
      a.n_shares = part.n_shares;
      a.tabAx    = plan.d_Ae[0];
      a.tabAy    = plan.d_Ae[1];
      a.tabBx    = plan.d_Be[0];
      a.tabBy    = plan.d_Be[1];
      a.sigma    = plan.sigma;
      for (int d = 0; d < 2; ++d)
        for (int f = 0; f < 2; ++f)
          for (int c = 0; c < 2 * (P + 1); ++c)
            {
              const int n    = L.N[d];
              const int node = (c <= P) ? c : n - P + (c - P - 1);
              for (int t = 0; t < W; ++t)
                {
                  const std::vector<double> &tab = f ? plan.hBe[d] : plan.hAe[d];
                  const double               v   = (f == 0 || op.has_B) ? tab[(size_t)node * W + t] : 0.0;
                  (d == 0 ? a.tbx : a.tby)[f][c][t] = v;
                }
            }
      a.zt       = plan.d_zt;
      a.jobs     = part.d_jobs;
      a.job_ptr  = part.d_ptr;
      a.ticket   = part.d_sync;
      a.error    = reinterpret_cast<int *>(part.d_sync + 1);
      a.flags    = part.d_sync + 2;
      a.scratch  = part.d_scratch;
      a.trace    = part.d_trace;
      a.xassign  = plan.d_xassign;
      a.dot_src  = dot_src;
      a.dot_partials = dot_partials;
      const std::vector<double> *hA = plan.hAe, *hB = plan.hBe;
      const int                  ir = P + 1; // any interior (Toeplitz) row
      for (int d = 0; d <= P; ++d)
        {
          a.Ax[d] = hA[0][(size_t)ir * W + P + d];
          a.Ay[d] = hA[1][(size_t)ir * W + P + d];
          a.Bx[d] = op.has_B ? hB[0][(size_t)ir * W + P + d] : 0.0;
          a.By[d] = op.has_B ? hB[1][(size_t)ir * W + P + d] : 0.0;
        }
      for (int j = 0; j < W; ++j)
        a.Az[j] = a.Bz[j] = 0.0;
      if (plan.kz_hi > plan.kz_lo)
        {
          // interior rows of direction 2 are Toeplitz: the scatter row of an interior input plane is any interior row reversed
          int r = -1;
          for (int q = 0; q < L.ln[2]; ++q)
            if (q + L.loc0 > P && q + L.loc0 < L.N[2] - P)
              {
                r = q;
                break;
              }
          GDM_REQUIRE(r >= 0, GDM_ERR_INTERNAL, "persistent fused kernel: no interior z row on this rank");
          for (int j = 0; j < W; ++j)
            {
              a.Az[j] = op.desc.scale * hA[2][(size_t)r * W + (2 * P - j)];
              a.Bz[j] = op.has_B ? op.desc.scale * hB[2][(size_t)r * W + (2 * P - j)] : 0.0;
            }
        }
      void (*kern)(const CUtensorMap, const ArgsP<P>) = nullptr;
      const bool dot = dot_partials != nullptr;
      GDM_REQUIRE(!(dot && accumulate), GDM_ERR_INTERNAL, "fused dot product with accumulation");
      GDM_REQUIRE(!dot || part.n_shares <= pers_max_partials(op, &plan), GDM_ERR_INTERNAL, "fused dot product: too many shares");
      if (dot)
        kern = kron3d_pers_kernel<C, MODE, false, true>;
      else if (accumulate)
        kern = kron3d_pers_kernel<C, MODE, true, false>;
      else
        kern = kron3d_pers_kernel<C, MODE, false, false>;
      const size_t smem = smem_bytes_p<C, MODE != 0>();
      GDM_REQUIRE(smem <= 227 * 1024, GDM_ERR_INTERNAL, "persistent fused kernel exceeds the shared memory of an SM");
      bool &attr = plan.attr_set[{ctx.device, (const void *)kern}];
      if (!attr)
        {
          GDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr = true;
        }
#ifdef GDM_PERS_WATCHDOG
      static WdHost *wd_host = nullptr;
      if (!wd_host)
        {
          GDM_CUDA_CHECK(cudaHostAlloc(&wd_host, sizeof(WdHost), cudaHostAllocMapped));
          memset(wd_host, 0, sizeof(WdHost));
        }
      GDM_CUDA_CHECK(cudaHostGetDevicePointer(&a.wd, wd_host, 0));
#endif
      part.epoch += 1;
      a.epoch       = part.epoch;
      a.ticket_base = part.ticket_base;
      part.ticket_base += (unsigned)(part.n_shares + part.grid); // every CTA draws one ticket past the last share
      const CUtensorMap &map = get_map<C>(op, plan, src);
      kern<<<part.grid, C::THREADS, smem, stream>>>(map, a);
      ctx.launches++;
      GDM_CUDA_CHECK(cudaGetLastError());
#ifdef GDM_PERS_WATCHDOG
      {
        // wait for the launch (up to 20 s); print what the kernel recorded and, if it still runs, where every CTA stands
        const double t0   = (double)clock() / CLOCKS_PER_SEC;
        bool         done = false;
        while (!(done = (cudaStreamQuery(stream) == cudaSuccess)) && (double)clock() / CLOCKS_PER_SEC - t0 < 20.0)
          {}
        volatile WdHost *w   = wd_host;
        const unsigned   cnt = w->count;
        if (!done || cnt > 0)
          {
            fprintf(stderr, "[gdm watchdog] launch %u of partition (%d shares, grid %d): %s, %u waits timed out\n", part.epoch, part.n_shares,
                    part.grid, done ? "finished" : "STILL RUNNING after 20 s", cnt);
            for (unsigned i = 0; i < std::min(cnt, 64u); ++i)
              fprintf(stderr, "  wait code %d (1 full, 2 empty, 3 tma, 9 seam flag) cta %d warp %d q %d share %d seq %d parity/flag %d aux %d\n",
                      w->rec[i][0], w->rec[i][1], w->rec[i][2], w->rec[i][3], w->rec[i][4], w->rec[i][5], w->rec[i][6], w->rec[i][7]);
            for (int b = 0; b < std::min(part.grid, 2048); ++b)
              fprintf(stderr, "  cta %d: share %d, %d shares done, marker %d, seq %d\n", b, w->prog[b][0], w->prog[b][1], w->prog[b][2],
                      w->prog[b][3]);
            fflush(stderr);
            if (!done)
              abort();
            memset(wd_host, 0, sizeof(WdHost));
            throw Error(GDM_ERR_INTERNAL, "persistent kernel watchdog: a wait timed out (see stderr)");
          }
      }
#endif
      if (part.d_trace != nullptr) // diagnostic: per-share cycles and SM of this launch -> file named by GDM_PERS_TRACE
        {
          GDM_CUDA_CHECK(cudaStreamSynchronize(stream));
          std::vector<long long> t((size_t)2 * part.grid);
          GDM_CUDA_CHECK(cudaMemcpy(t.data(), part.d_trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
          if (FILE *f = fopen(std::getenv("GDM_PERS_TRACE"), "w"))
            {
              fprintf(f, "# cta cycles smid\n");
              for (int w = 0; w < part.grid; ++w)
                fprintf(f, "%d %lld %lld\n", w, t[2 * w], t[2 * w + 1]);
              fclose(f);
            }
        }
      return part.n_shares;
    }
  } // namespace

  namespace
  {
    // ---- periodic directions: C^T A C x = fold( A ( dup x ) ) on the N+1 stored nodes per direction.
    // With several ranks the partitioned direction wraps between the first and the last slab: the last rank receives
    // plane 0 of rank 0 (zimg) before it patches its plane N, rank 0 receives row N of the last rank (zimg again) before
    // it folds its plane 0 (one extra plane between the last and the first rank).
    struct PeriodicK
    {
      double       *v;     // vector that is patched / folded / restored
      double       *save;  // saved values of the patched nodes
      const double *zimg;  // one plane received from the peer rank (nullptr: the wrap in z is local)
      int           ln[3]; // stored nodes per direction
      int           face[3]; // coordinate (local index) of the face this kernel enumerates per direction, -1: none
      int           hi[3], lo[3]; // local index of the duplicate node N / of node 0 per periodic direction (-1: not periodic;
                                  // for direction 2 the index may lie outside this rank's planes)
      int64_t       stride[3];
      int64_t       face_off[4]; // prefix sums of the face sizes (in direction order)
      int           z_lo, z_hi;  // planes of direction 2 this rank works on (local indices: its owned planes)
    };

    // thread -> node of the union of the enumerated faces; a node on several faces is handled through the lowest
    // direction.  Returns false if the thread has no node.
    __device__ __forceinline__ bool periodic_node(const PeriodicK &a, int64_t tid, int (&idx)[3])
    {
      if (tid >= a.face_off[3])
        return false;
      int d = 0;
      while (tid >= a.face_off[d + 1])
        ++d;
      tid -= a.face_off[d];
      const int e0 = (d == 0) ? 1 : 0, e1 = (d == 2) ? 1 : 2;
      idx[d]       = a.face[d];
      idx[e0]      = (int)(tid % a.ln[e0]);
      idx[e1]      = (int)(tid / a.ln[e0]);
      if (idx[e1] >= a.ln[e1])
        return false;
      for (int e = 0; e < d; ++e)
        if (a.face[e] >= 0 && idx[e] == a.face[e])
          return false;
      return idx[2] >= a.z_lo && idx[2] < a.z_hi;
    }

    // x~ = C x: every node with a periodic coordinate N takes the value of its image in the fundamental domain; the old
    // value is saved (the vector is restored after the apply)
    __global__ void periodic_patch_kernel(const PeriodicK a)
    {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      int           idx[3];
      if (!periodic_node(a, tid, idx))
        return;
      int64_t off = 0, img = 0;
      bool    remote = false;
      for (int d = 0; d < 3; ++d)
        {
          off += idx[d] * a.stride[d];
          const bool dup = a.hi[d] >= 0 && idx[d] == a.hi[d];
          if (d == 2 && dup && a.zimg != nullptr)
            remote = true; // the image plane (global plane 0) was received into zimg
          else
            img += (dup ? a.lo[d] : idx[d]) * a.stride[d];
        }
      a.save[tid] = a.v[off];
      a.v[off]    = remote ? a.zimg[img] : a.v[img];
    }
    __global__ void periodic_restore_kernel(const PeriodicK a)
    {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      int           idx[3];
      if (!periodic_node(a, tid, idx))
        return;
      int64_t off = 0;
      for (int d = 0; d < 3; ++d)
        off += idx[d] * a.stride[d];
      a.v[off] = a.save[tid];
    }
    // y = C^T y~: a node with periodic coordinates equal to 0 adds the rows of its images (coordinate N)
    __global__ void periodic_fold_kernel(const PeriodicK a)
    {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      int           idx[3];
      if (!periodic_node(a, tid, idx))
        return;
      for (int d = 0; d < 3; ++d)
        if (a.hi[d] >= 0 && idx[d] == a.hi[d])
          return; // itself a duplicate node: its row is overwritten by the constrained-row kernel
      int64_t off = 0;
      for (int d = 0; d < 3; ++d)
        off += idx[d] * a.stride[d];
      double sum = a.v[off];
      for (int m = 1; m < 8; ++m)
        {
          int64_t o  = 0;
          bool    ok = true, remote = false;
          for (int d = 0; d < 3; ++d)
            {
              const bool shift = (m >> d) & 1;
              if (shift && !(a.lo[d] >= 0 && idx[d] == a.lo[d] && a.hi[d] != -1))
                ok = false;
              if (shift && d == 2 && a.zimg != nullptr)
                remote = true; // row N of the last rank was received into zimg
              else
                o += (shift ? a.hi[d] : idx[d]) * a.stride[d];
            }
          if (ok)
            sum += remote ? a.zimg[o] : a.v[o];
        }
      a.v[off] = sum;
    }

    // which: 0 patch / restore (faces at the duplicate nodes), 1 fold (faces at node 0)
    PeriodicK periodic_args(const Operator &op, PersPlan &plan, double *v, int which)
    {
      const Layout &L = op.sys->L;
      PeriodicK     a;
      a.v    = v;
      a.zimg = nullptr;
      a.z_lo = L.own0 - L.loc0;
      a.z_hi = L.own1 - L.loc0;
      a.face_off[0] = 0;
      for (int d = 0; d < 3; ++d)
        {
          a.ln[d]     = L.ln[d];
          a.stride[d] = L.stride[d];
          const int shift = (d == 2) ? L.loc0 : 0;
          a.hi[d]     = plan.periodic[d] ? L.N[d] - shift : -1;
          a.lo[d]     = plan.periodic[d] ? 0 - shift : -1;
          if (plan.periodic[d] && a.hi[d] == -1)
            a.hi[d] = -2; // (periodic, but the duplicate plane is far below this rank's planes: keep it distinct from -1)
          const int f = plan.periodic[d] ? (which == 0 ? a.hi[d] : a.lo[d]) : -1;
          bool stored = f >= 0 && f < L.ln[d];
          if (d == 2)
            stored = stored && f >= a.z_lo && f < a.z_hi;
          a.face[d]   = stored ? f : -1;
          const int e0 = (d == 0) ? 1 : 0, e1 = (d == 2) ? 1 : 2;
          a.face_off[d + 1] = a.face_off[d] + (a.face[d] >= 0 ? (int64_t)L.ln[e0] * L.ln[e1] : 0);
        }
      const int64_t need = std::max<int64_t>(a.face_off[3], 1) + L.plane; // saved values + one received plane
      if (plan.n_save < need)
        {
          GDM_CUDA_CHECK(cudaDeviceSynchronize());
          cudaFree(plan.d_save);
          plan.d_save = nullptr;
          GDM_CUDA_CHECK(cudaMalloc(&plan.d_save, (size_t)need * sizeof(double)));
          plan.n_save = need;
        }
      a.save = plan.d_save + L.plane;
      return a;
    }
  } // namespace

  bool pers_has_periodic(const void *p)
  {
    const PersPlan &plan = *static_cast<const PersPlan *>(p);
    return plan.periodic[0] || plan.periodic[1] || plan.periodic[2];
  }

  void pers_periodic_pre(Operator &op, void *p, double *src, cudaStream_t stream)
  {
    PersPlan &plan = *static_cast<PersPlan *>(p);
    if (!pers_has_periodic(p))
      return;
    const Layout &L = op.sys->L;
    Context      &ctx = *op.sys->ctx;
    PeriodicK     a = periodic_args(op, plan, src, 0);
    if (plan.periodic[2] && L.n_ranks > 1)
      {
        // plane 0 of the first rank -> the rank that owns plane N
        const int first = comm_plane_owner(L, 0), last = comm_plane_owner(L, L.N[2]);
        if (first != last)
          {
            if (L.rank == first)
              comm_send(ctx, src + (int64_t)(0 - L.loc0) * L.plane, L.plane, last, stream);
            if (L.rank == last)
              {
                comm_recv(ctx, plan.d_save, L.plane, first, stream);
                a.zimg = plan.d_save;
              }
          }
      }
    if (a.face_off[3] == 0)
      return;
    const int th = 128;
    periodic_patch_kernel<<<(unsigned)((a.face_off[3] + th - 1) / th), th, 0, stream>>>(a);
    ctx.launches++;
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  void pers_periodic_post(Operator &op, void *p, double *dst, double *src, cudaStream_t stream)
  {
    PersPlan &plan = *static_cast<PersPlan *>(p);
    if (!pers_has_periodic(p))
      return;
    const Layout &L = op.sys->L;
    Context      &ctx = *op.sys->ctx;
    const int     th = 128;
    // restore the input first (the received plane buffer is reused below)
    PeriodicK r = periodic_args(op, plan, src, 0);
    if (r.face_off[3] > 0)
      {
        periodic_restore_kernel<<<(unsigned)((r.face_off[3] + th - 1) / th), th, 0, stream>>>(r);
        ctx.launches++;
      }
    PeriodicK a = periodic_args(op, plan, dst, 1);
    if (plan.periodic[2] && L.n_ranks > 1)
      {
        // row N of the last rank -> the first rank (added to its plane 0)
        const int first = comm_plane_owner(L, 0), last = comm_plane_owner(L, L.N[2]);
        if (first != last)
          {
            if (L.rank == last)
              comm_send(ctx, dst + (int64_t)(L.N[2] - L.loc0) * L.plane, L.plane, first, stream);
            if (L.rank == first)
              {
                comm_recv(ctx, plan.d_save, L.plane, last, stream);
                a.zimg = plan.d_save;
              }
          }
      }
    if (a.face_off[3] > 0)
      {
        periodic_fold_kernel<<<(unsigned)((a.face_off[3] + th - 1) / th), th, 0, stream>>>(a);
        ctx.launches++;
      }
    GDM_CUDA_CHECK(cudaGetLastError());
  }

  bool pers_supported(const Operator &op)
  {
    const Layout &L = op.sys->L;
    if (L.dim != 3 || L.nc != 1)
      return false;
    if (!(L.p == 1 || L.p == 3 || L.p == 5))
      return false;
    for (int d = 0; d < 3; ++d)
      {
        if (L.N[d] < 2 * L.p + 2)
          return false;
        // a periodic direction is applied as fold . A . duplicate with the plain one-sided tables
        if (op.periodic[d] && (op.dirichlet[d][0] || op.dirichlet[d][1] || op.hAu[d].empty()))
          return false;
      }
    if (L.own1 <= L.own0)
      return false;
    const char *env = std::getenv("GDM_PERS_PERIODIC");
    if (env && env[0] == '0' && (op.periodic[0] || op.periodic[1] || op.periodic[2]))
      return false;
    return true;
  }

  // returns nullptr if the operator cannot use the tap split it relies on (the caller falls back to the round-1 kernels)
  void *pers_plan_create(Operator &op)
  {
    const Layout &L = op.sys->L;
    const int     P = L.p, W = 2 * P + 1;
    std::unique_ptr<PersPlan> plan(new PersPlan);
    plan->cfg = default_config(P);
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
    // stored planes that can contribute: Dirichlet planes have zero columns
    plan->in_lo = (op.dirichlet[2][0] && L.loc0 == 0) ? 1 : 0;
    plan->in_hi = L.ln[2] - ((op.dirichlet[2][1] && L.loc1 == L.nn[2]) ? 1 : 0);
    int tx = 32, ty = 32;
    with_config(plan->cfg, [&](auto c) {
      tx = decltype(c)::TX;
      ty = decltype(c)::TY;
    });
    // TMA box starts must be 16-byte aligned: keep (x0 - P) even by starting one column early if needed
    plan->xorg    = plan->cx0 - ((plan->cx0 - P) & 1);
    plan->tiles_x = (plan->cx1 - plan->xorg + tx - 1) / tx;
    plan->tiles_y = (plan->cy1 - plan->cy0 + ty - 1) / ty;
    plan->mode    = !op.has_B ? 0 : (op.b_symmetry > 0 ? 1 : 2);
    // Toeplitz z planes [kz_lo, kz_hi): input planes whose 2P+1 target rows are all interior rows of A_z, B_z
    plan->kz_lo = plan->kz_hi = 0;
    if (L.N[2] >= 4 * P + 2)
      {
        plan->kz_lo = std::max(0, 2 * P + 1 - L.loc0);
        plan->kz_hi = std::max(plan->kz_lo, std::min(L.ln[2], L.N[2] - 2 * P - L.loc0));
      }
    const int zrows = 2 * W;
    if (plan->kz_lo + (L.ln[2] - plan->kz_hi) > zrows)
      return nullptr;
    for (int d = 0; d < 3; ++d)
      {
        plan->periodic[d] = op.periodic[d];
        plan->hAe[d]      = op.periodic[d] ? op.hAu[d] : op.hA[d];
        plan->hBe[d]      = op.periodic[d] ? op.hBu[d] : op.hB[d];
      }
    const std::vector<double> *hA = plan->hAe;
    const std::vector<double>  hB0[3] = {plan->hBe[0], plan->hBe[1], plan->hBe[2]};
    plan->sigma = 0.0;
    if (plan->mode == 1)
      {
        // K_d = alpha_d M_d + R_d with alpha_d = (outer tap of K_d) / (outer tap of M_d): R_d has zero outer taps on
        // Toeplitz rows; the identity holds row by row, so the one-sided and masked rows need no special treatment
        double alpha[3] = {0, 0, 0};
        for (int d = 0; d < 3; ++d)
          {
            int row = P + 1;
            if (d == 2)
              {
                row = -1;
                for (int r = 0; r < L.ln[2]; ++r)
                  if (r + L.loc0 > P && r + L.loc0 < L.N[2] - P)
                    {
                      row = r;
                      break;
                    }
              }
            if (row < 0)
              return nullptr;
            const double m = hA[d][(size_t)row * W + 2 * P], k = hB0[d][(size_t)row * W + 2 * P];
            if (m == 0.0 || hA[d][(size_t)row * W] != m || hB0[d][(size_t)row * W] != k)
              return nullptr;
            alpha[d] = k / m;
          }
        for (int d = 0; d < 3; ++d)
          {
            const int rows = (int)(hB0[d].size() / W);
            for (int r = 0; r < rows; ++r)
              {
                for (int t = 0; t < W; ++t)
                  plan->hBe[d][(size_t)r * W + t] = hB0[d][(size_t)r * W + t] - alpha[d] * hA[d][(size_t)r * W + t];
                const int gr = r + (d == 2 ? L.loc0 : 0); // global row
                if (gr > P && gr < L.N[d] - P)
                  plan->hBe[d][(size_t)r * W] = plan->hBe[d][(size_t)r * W + 2 * P] = 0.0;
              }
          }
        plan->sigma = alpha[0] + alpha[1] + alpha[2];
      }
    for (int d = 0; d < 2; ++d)
      {
        GDM_CUDA_CHECK(cudaMalloc(&plan->d_Ae[d], plan->hAe[d].size() * sizeof(double)));
        GDM_CUDA_CHECK(cudaMemcpy(plan->d_Ae[d], plan->hAe[d].data(), plan->hAe[d].size() * sizeof(double), cudaMemcpyHostToDevice));
        if (op.has_B)
          {
            GDM_CUDA_CHECK(cudaMalloc(&plan->d_Be[d], plan->hBe[d].size() * sizeof(double)));
            GDM_CUDA_CHECK(cudaMemcpy(plan->d_Be[d], plan->hBe[d].data(), plan->hBe[d].size() * sizeof(double), cudaMemcpyHostToDevice));
          }
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
            zt[(size_t)(c * 2 + 0) * wz + j] = op.desc.scale * hA[2][(size_t)r * W + (2 * P - j)];
            if (op.has_B)
              zt[(size_t)(c * 2 + 1) * wz + j] = op.desc.scale * plan->hBe[2][(size_t)r * W + (2 * P - j)];
          }
      }
    GDM_CUDA_CHECK(cudaMalloc(&plan->d_zt, zt.size() * sizeof(double)));
    GDM_CUDA_CHECK(cudaMemcpy(plan->d_zt, zt.data(), zt.size() * sizeof(double), cudaMemcpyHostToDevice));
    return plan.release();
  }

  void pers_plan_destroy(void *p)
  {
    delete static_cast<PersPlan *>(p);
  }

  void pers_window(const void *p, int &cz0, int &cz1)
  {
    const PersPlan &plan = *static_cast<const PersPlan *>(p);
    cz0                  = plan.cz0;
    cz1                  = plan.cz1;
  }

  int pers_max_grid(const Operator &op, const void *p)
  {
    const PersPlan &plan = *static_cast<const PersPlan *>(p);
    int             mb   = 2;
    with_config(plan.cfg, [&](auto c) { mb = decltype(c)::MINB; });
    int slots = op.sys->ctx->sm_count * mb;
    if (const char *env = std::getenv("GDM_PERS_SLOTS"))
      slots = std::max(slots, atoi(env));
    return slots;
  }

  // Plan search at operator creation (never inside an apply): the time of a launch is a jagged function of the plane
  // costs the static partition assumes for edge tiles (132 ... 145 GDoF/s over nine nearby weight triples at 257^3,
  // profiles/r2/session_z2_weights.txt), because the cost of a share depends on the CTA it shares an SM with.  So a few
  // candidate triples are measured on two scratch vectors (median of 5 launches each after 2 warm-up launches) and the
  // fastest is kept.  Only for windows large enough to be cut into shares; GDM_PERS_TUNE=0 or GDM_PERS_WEIGHTS disable it
  // (the partition, and with it the last bits of the result, then no longer depend on a measurement).
  void pers_tune(Operator &op, void *p)
  {
    PersPlan     &plan = *static_cast<PersPlan *>(p);
    Context      &ctx  = *op.sys->ctx;
    const Layout &L    = op.sys->L;
    const char   *env  = std::getenv("GDM_PERS_TUNE");
    if ((env && env[0] == '0') || std::getenv("GDM_PERS_WEIGHTS") || plan.tuned)
      return;
    plan.tuned           = true;
    const int64_t tiles  = (int64_t)plan.tiles_x * plan.tiles_y;
    const int64_t work   = tiles * std::max(0, plan.in_hi - plan.in_lo);
    const int     slots  = pers_max_grid(op, p);
    if (tiles < 8 || slots < tiles || work < (int64_t)24 * slots || plan.cz1 - plan.cz0 < 16 * L.p)
      return;
    static const int cand[][3] = {{1300, 1400, 1600}, {1250, 1250, 1400}, {1450, 1450, 1650}, {1400, 1400, 1550},
                                  {1250, 1350, 1450}, {1550, 1550, 1750}, {1300, 1300, 1450}, {1150, 1150, 1250}};
    auto clear_parts = [&]() {
      GDM_CUDA_CHECK(cudaDeviceSynchronize());
      for (auto &kv : plan.parts)
        PersPlan::free_part(kv.second);
      plan.parts.clear();
    };
    double *src = ctx.acquire((size_t)L.size), *dst = ctx.acquire((size_t)L.size);
    GDM_CUDA_CHECK(cudaMemsetAsync(src, 0, (size_t)L.size * sizeof(double), ctx.stream));
    cudaEvent_t e0, e1;
    GDM_CUDA_CHECK(cudaEventCreate(&e0));
    GDM_CUDA_CHECK(cudaEventCreate(&e1));
    int    best = 0;
    double best_ms = 1e30;
    const bool verbose = std::getenv("GDM_FUSED_VERBOSE") != nullptr;
    for (int c = 0; c < (int)(sizeof(cand) / sizeof(cand[0])); ++c)
      {
        plan.wx = cand[c][0], plan.wy = cand[c][1], plan.wxy = cand[c][2];
        clear_parts();
        float ms[5];
        for (int i = -2; i < 5; ++i)
          {
            GDM_CUDA_CHECK(cudaEventRecord(e0, ctx.stream));
            pers_launch(op, p, dst, src, false, plan.cz0, plan.cz1, ctx.stream, nullptr, nullptr, 0);
            GDM_CUDA_CHECK(cudaEventRecord(e1, ctx.stream));
            GDM_CUDA_CHECK(cudaEventSynchronize(e1));
            if (i >= 0)
              GDM_CUDA_CHECK(cudaEventElapsedTime(&ms[i], e0, e1));
          }
        std::sort(ms, ms + 5);
        if (verbose)
          fprintf(stderr, "[gdm] plan search: edge-tile plane costs %d,%d,%d -> %.4f ms (median of 5)\n", cand[c][0], cand[c][1],
                  cand[c][2], ms[2]);
        if (ms[2] < best_ms)
          {
            best_ms = ms[2];
            best    = c;
          }
      }
    plan.wx = cand[best][0], plan.wy = cand[best][1], plan.wxy = cand[best][2];
    clear_parts();
    ctx.launches -= 7 * (int)(sizeof(cand) / sizeof(cand[0])); // (the search is not part of any apply)
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx.release(src);
    ctx.release(dst);
  }

  // upper bound of the number of shares (= fused-dot partial sums) of one launch
  int pers_max_partials(const Operator &op, const void *p)
  {
    return std::max(pers_max_grid(op, p), 8192);
  }

  // output planes [oz0, oz1) (local indices, clipped to the plan's window); returns the number of shares of the launch
  // (= the number of fused-dot partial sums it wrote)
  int pers_launch(Operator &op, void *p, double *dst, const double *src, bool accumulate, int oz0, int oz1, cudaStream_t stream,
                  const double *dot_src, double *dot_partials, int slots_limit)
  {
    PersPlan &plan = *static_cast<PersPlan *>(p);
    oz0            = std::max(oz0, plan.cz0);
    oz1            = std::min(oz1, plan.cz1);
    int grid       = 0;
    with_config(plan.cfg, [&](auto c) {
      using C = decltype(c);
      if (plan.mode == 0)
        grid = launch_p<C, 0>(op, plan, dst, src, accumulate, oz0, oz1, stream, dot_src, dot_partials, slots_limit);
      else if (plan.mode == 1)
        grid = launch_p<C, 1>(op, plan, dst, src, accumulate, oz0, oz1, stream, dot_src, dot_partials, slots_limit);
      else
        grid = launch_p<C, 2>(op, plan, dst, src, accumulate, oz0, oz1, stream, dot_src, dot_partials, slots_limit);
    });
    return grid;
  }

  // 0 if no seam wait of any launch of this plan timed out (diagnostic; synchronises the device)
  int pers_error_flag(void *p)
  {
    PersPlan &plan = *static_cast<PersPlan *>(p);
    int       any  = 0;
    GDM_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto &kv : plan.parts)
      if (kv.second.d_sync)
        {
          unsigned e = 0;
          GDM_CUDA_CHECK(cudaMemcpy(&e, kv.second.d_sync + 1, sizeof(unsigned), cudaMemcpyDeviceToHost));
          any |= (int)e;
        }
    return any;
  }
} // namespace gdm
