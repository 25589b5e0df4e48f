// kron3d_v5 -- asynchronous, load-balanced variant of kron3d_v4 (included by kron3d.cu inside gdm::<anon>).
//
// ncu on v4 (profiles/r1/kron3d_v4a_*): 20 % of the warp time in the per-plane CTA barrier, 23 % of the SM time idle
// (576 CTAs on 296 slots = 1.95 waves), 20 % redundant x/y work in the z ramp of the short chunks.  v5 keeps the
// arithmetic of v4 (same x/y/z passes, same tap split) and changes the control structure:
//   * static balanced partition: one CTA per SM slot, each with a contiguous run of (tile, plane) work computed on the
//     host (segments {tile x, tile y, z0, z1}); a CTA crosses into the next tile column when its share does, so all
//     slots finish together and the number of z ramps is (slots + tiles) instead of (tiles x chunks);
//   * no CTA barrier in the plane loop: full/empty mbarrier rings (TMA stage full / stage empty, a/b buffer full / a/b
//     buffer empty, three a/b buffers).  A warp waits for "all warps finished the x pass of plane k" one iteration
//     after it arrived itself, so warps drift by up to one plane and hide each other's LDS / barrier latency.

template <int P_, int RY_, int NRB_, int RX_, int STAGES_, int MINB_>
struct Cfg5
{
  static constexpr bool V4 = true, V5 = true, V6 = false, V7 = false;
  static constexpr int  P = P_, TX = 32, RY = RY_, NRB = NRB_, RX = RX_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int  NXW     = 0;
  static constexpr int  W       = 2 * P + 1;
  static constexpr int  TY      = RY * NRB;
  static constexpr int  NR      = TY + 2 * P;
  static constexpr int  PIN     = TX + 2 * P;
  static constexpr int  PY      = ((NR / 2) & 1) ? NR : NR + 2;
  static constexpr int  THREADS = 32 * NRB;
  static constexpr int  NWARPS  = NRB;
  static constexpr int  NXB     = TX / RX;
  static constexpr int  NAB     = 3;
  // x pass decomposition: warp tasks (x block, group of 32 rows); a last partial group with fewer than 16 rows is
  // computed as single outputs spread over all warps
  static constexpr int  NG_FULL    = NR / 32;
  static constexpr int  REM        = NR % 32;
  static constexpr bool REM_BLOCK  = REM >= 16;
  static constexpr int  NWT        = NXB * (NG_FULL + (REM_BLOCK ? 1 : 0));
  static constexpr int  ROUNDS     = (NWT + NWARPS - 1) / NWARPS;
  static constexpr int  NSINGLE    = REM_BLOCK ? 0 : REM * TX;
  static constexpr int  S_PER_WARP = (NSINGLE + NWARPS - 1) / NWARPS;
  static constexpr int  S_ROUNDS   = (S_PER_WARP + 31) / 32;
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

template <class C, bool HASB>
constexpr size_t smem_bytes_v5()
{
  return (size_t)(C::STAGES * C::STAGE_DOUBLES + C::NAB * (HASB ? 2 : 1) * C::TX * C::PY + C::ZT_DOUBLES + C::TB_DOUBLES) * sizeof(double) +
         (size_t)(2 * C::STAGES + 2 * C::NAB) * sizeof(uint64_t) + 128;
}

__device__ __forceinline__ void mbar_arrive_a(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_test_a(uint32_t bar, unsigned parity)
{
  unsigned ok;
  asm volatile(
    "{\n"
    ".reg .pred P1;\n"
    "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
    "selp.u32 %0, 1, 0, P1;\n"
    "}\n"
    : "=r"(ok)
    : "r"(bar), "r"(parity)
    : "memory");
  return ok != 0;
}

template <class C, int MODE, bool RSPLIT, bool ACCUM>
__global__ void __launch_bounds__(C::THREADS, C::MINB) kron3d_v5_kernel(const __grid_constant__ CUtensorMap tmap, const KArgs<C::P> g)
{
  constexpr int  P = C::P, W = C::W, TX = C::TX, RY = C::RY, RX = C::RX, NR = C::NR, PIN = C::PIN, PY = C::PY;
  constexpr bool HASB = MODE != 0, SYM = MODE == 1;
  constexpr int  NF = HASB ? 2 : 1;
  constexpr int  PB = (RSPLIT && SYM) ? P - 1 : P; // outermost tap of the interior B rows in x and y
  extern __shared__ __align__(128) double smem[];
  constexpr int AB_BUF  = NF * TX * PY; // [field][x][PY] (y contiguous)
  constexpr int OFF_AB  = C::STAGES * C::STAGE_DOUBLES;
  constexpr int OFF_ZT  = OFF_AB + C::NAB * AB_BUF;
  constexpr int OFF_TB  = OFF_ZT + C::ZT_DOUBLES;
  constexpr int OFF_BAR = OFF_TB + C::TB_DOUBLES;
  constexpr int WP = C::WP, NBT = C::NBT, WZ = C::WZ;
  constexpr unsigned STAGE_BYTES = NR * PIN * sizeof(double);

  const uint32_t sb       = smem_u32(smem);
  const uint32_t bar0     = sb + OFF_BAR * 8;          // tma_full[STAGES]
  const uint32_t bar_te   = bar0 + 8 * C::STAGES;      // tma_empty[STAGES]
  const uint32_t bar_af   = bar_te + 8 * C::STAGES;    // ab_full[NAB]
  const uint32_t bar_ae   = bar_af + 8 * C::NAB;       // ab_empty[NAB]
  const int      tid      = threadIdx.x;
  const int      lane = tid & 31, warp = tid >> 5;

  if (tid == 0)
    {
      for (int s = 0; s < C::STAGES; ++s)
        {
          mbar_init_a(bar0 + 8 * s, 1);
          mbar_init_a(bar_te + 8 * s, C::NWARPS);
        }
      for (int s = 0; s < C::NAB; ++s)
        {
          mbar_init_a(bar_af + 8 * s, C::NWARPS);
          mbar_init_a(bar_ae + 8 * s, C::NWARPS);
        }
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
  __syncthreads();

  // ring state, running over all segments of this CTA: x pass (TMA stage + a/b buffer written), y/z pass (a/b buffer
  // read), TMA issue (thread 0)
  int      xs = 0, xb_ = 0, yb_ = 0, is = 0;
  unsigned xs_par = 0, xb_par = 0, yb_par = 0, is_par = 0;

  const int yz_off = lane * PY + warp * RY; // start of this thread's window in an a/b buffer
  // per-segment geometry (set at the top of the segment loop)
  int     x0 = 0, y0 = 0, gy_first = 0, nst = 0;
  bool    y_fix = false;
  double *out   = nullptr;

  double acc[RY][2 * P];

  // ---- x pass of one plane: staged tile at in_off -> a/b buffer at a_off (doubles)
  auto x_pass = [&](const int in_off, const int a_off) {
    const int b_off = a_off + (NF - 1) * TX * PY;
#pragma unroll
    for (int rd = 0; rd < C::ROUNDS; ++rd)
      {
        const int wt = warp + rd * C::NWARPS;
        if ((C::NWT % C::NWARPS) != 0 && wt >= C::NWT)
          break;
        const int xb = wt % C::NXB, gr = wt / C::NXB;
        const int r  = gr * 32 + lane;
        if (C::REM_BLOCK && r >= NR)
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
    if constexpr (C::NSINGLE > 0)
      {
#pragma unroll
        for (int sr = 0; sr < C::S_ROUNDS; ++sr)
          {
            const int l = lane + 32 * sr;
            const int o = warp * C::S_PER_WARP + l;
            if (l >= C::S_PER_WARP || o >= C::NSINGLE)
              continue;
            const int r = C::NG_FULL * 32 + o / TX, x = o % TX;
            double    v[W];
#pragma unroll
            for (int t = 0; t < W; ++t)
              v[t] = smem[in_off + r * PIN + x + t];
            const int gxx = x0 + x;
            double    ra, rbv = 0.0;
            if ((gxx <= P || gxx >= g.nx - P) && gxx >= 0 && gxx <= g.nx)
              {
                const int     rc = (gxx <= P) ? gxx : gxx - (g.nx - P) + P + 1;
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
                if (HASB && SYM)
                  rbv = g.Bx[0] * v[P];
#pragma unroll
                for (int d = 1; d <= P; ++d)
                  {
                    const double s = v[P - d] + v[P + d];
                    ra             = fma(g.Ax[d], s, ra);
                    if (HASB)
                      {
                        if (SYM)
                          {
                            if (d <= PB)
                              rbv = fma(g.Bx[d], s, rbv);
                          }
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

  // ---- y pass + z pass of input plane k from the a/b buffer at ab_off; stores output plane k - P
  auto yz_pass = [&](const int k, const int ab_off, const bool store) {
    double u1[RY], u2[RY];
    {
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
          u1[i] = t1;
          u2[i] = t2;
        }
      if (y_fix)
        {
#pragma unroll
          for (int i = 0; i < RY; ++i)
            {
              const int gy = gy_first + i;
              if ((gy <= P || gy >= g.ny - P) && gy <= g.ny)
                {
                  const int     rc = (gy <= P) ? gy : gy - (g.ny - P) + P + 1;
                  const double *ta = smem + OFF_TB + (2 * NBT + rc) * WP;
                  const double *tb = smem + OFF_TB + (3 * NBT + rc) * WP;
                  double        t1 = 0.0, t2 = 0.0;
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
                  u1[i] = t1;
                  u2[i] = t2;
                }
            }
        }
    }
    double res[RY];
    if (k >= g.kz_lo && k < g.kz_hi)
      z_pass_v4<P, RY, MODE, RSPLIT, true>(g.Az, g.Bz, g.sigma, u1, u2, acc, res);
    else
      {
        // plane class: planes below kz_lo by index, planes from kz_hi on after them; planes outside the slab carry zeros
        const int kk = min(max(k, 0), g.nz_local - 1);
        const int c  = (kk < g.kz_lo) ? kk : g.kz_lo + (kk - g.kz_hi);
        double    zA[W], zB[W];
#pragma unroll
        for (int j = 0; j < W; ++j)
          {
            zA[j] = smem[OFF_ZT + (c * 2 + 0) * WZ + j];
            zB[j] = HASB ? smem[OFF_ZT + (c * 2 + 1) * WZ + j] : 0.0;
          }
        z_pass_v4<P, RY, MODE, RSPLIT, false>(zA, zB, g.sigma, u1, u2, acc, res);
      }
    if (store)
      {
#pragma unroll
        for (int i = 0; i < RY; ++i)
          if (i < nst)
            {
              double *o = out + (int64_t)i * g.pitch;
              double  t = res[i];
              if (ACCUM)
                t += *o;
              *o = t;
            }
      }
  };

  // ---- one x pass with its ring bookkeeping
  auto x_step = [&]() {
    mbar_wait_a(bar0 + 8 * xs, xs_par);           // TMA stage landed
    mbar_wait_a(bar_ae + 8 * xb_, xb_par ^ 1u);   // a/b buffer released by the y pass that read it last
    if (!GDM_DBG(g, 4))
      x_pass(xs * C::STAGE_DOUBLES, OFF_AB + xb_ * AB_BUF);
    __syncwarp();
    if (lane == 0)
      {
        mbar_arrive_a(bar_af + 8 * xb_);
        mbar_arrive_a(bar_te + 8 * xs);
      }
    if (++xs == C::STAGES)
      {
        xs = 0;
        xs_par ^= 1u;
      }
    if (++xb_ == C::NAB)
      {
        xb_ = 0;
        xb_par ^= 1u;
      }
  };

  for (int si = g.seg_ptr[blockIdx.x]; si < g.seg_ptr[blockIdx.x + 1]; ++si)
    {
      const int4 sg  = g.segs[si];
      x0             = g.xorg + sg.x * TX;
      y0             = g.cy0 + sg.y * C::TY;
      const int zc0  = sg.z;
      const int kbeg = sg.z - P, kend = sg.w + P;
      const int gx   = x0 + lane;
      gy_first       = y0 + warp * RY;
      nst            = (gx >= g.cx0 && gx < g.cx1) ? (g.cy1 - gy_first) : 0; // rows i < nst are stored
      y_fix          = (gy_first <= P) || (gy_first + RY - 1 >= g.ny - P);
      out            = g.dst + (int64_t)(kbeg - P) * g.plane + (int64_t)gy_first * g.pitch + gx;
#pragma unroll
      for (int i = 0; i < RY; ++i)
#pragma unroll
        for (int j = 0; j < 2 * P; ++j)
          acc[i][j] = 0.0;

      // TMA issue (thread 0): plane kn goes to stage `is` once every warp has finished the x pass that read it last
      int  kn    = kbeg;
      auto issue = [&]() {
        mbar_wait_a(bar_te + 8 * is, is_par ^ 1u);
        mbar_expect_tx_a(bar0 + 8 * is, STAGE_BYTES);
        tma_load_3d_a(sb + is * C::STAGE_DOUBLES * 8, &tmap, bar0 + 8 * is, x0 - P, y0 - P, kn);
        ++kn;
        if (++is == C::STAGES)
          {
            is = 0;
            is_par ^= 1u;
          }
      };
      if (tid == 0)
        for (int s = 0; s < C::STAGES && kn < kend; ++s)
          issue();

      x_step(); // x pass of the first plane of the segment
      for (int k = kbeg; k < kend; ++k)
        {
          if (k + 1 < kend)
            x_step();
          mbar_wait_a(bar_af + 8 * yb_, yb_par); // every warp has written its part of plane k
          if (!GDM_DBG(g, 8))
            yz_pass(k, OFF_AB + yb_ * AB_BUF, (k - P >= zc0) && !GDM_DBG(g, 1));
          __syncwarp();
          if (lane == 0)
            mbar_arrive_a(bar_ae + 8 * yb_);
          if (++yb_ == C::NAB)
            {
              yb_ = 0;
              yb_par ^= 1u;
            }
          // keep STAGES planes in flight behind the x pass (plane k+1 has been consumed: refill up to k+1+STAGES)
          if (tid == 0)
            while (kn < kend && kn <= k + 1 + C::STAGES)
              issue();
          out += g.plane;
        }
    }
}
