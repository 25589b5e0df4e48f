// kron3d_v6 -- register-resident variant of the fused 3D tensor-product apply (included by kron3d.cu inside gdm::<anon>).
//
// ncu on v3/v4/v5 (profiles/r1): the x pass -> shared memory -> y pass hand-over makes every warp run a chain of four
// latency-bound stages per plane with a CTA-wide dependency in the middle; with 128 registers per thread only four warps
// per scheduler are resident and the FP64 pipe idles 58 % of the time no matter how the barrier is built.
// v6 removes the hand-over: a warp owns TX = 32 columns x RY rows of the tile, lane <-> x, and every thread
//   * reads the raw staged tile (TMA, zero filled outside the domain) directly: 2P+1 LDS.64 per input row, RY+2P rows;
//   * computes the x pass for its RY+2P rows itself (the 2P halo rows are recomputed instead of exchanged:
//     x pass cost (RY+2P)/RY instead of (TY+2P)/TY -- at p=3, RY=8: 47.5 instead of 42 FP64 operations per DoF);
//   * keeps a sliding window of 2P+1 rows of (a, r) in registers for the y pass (gather form, shared pair sums) and the
//     2P running z accumulators per point (scatter form) as before.
// Warps never exchange data: the only synchronisation is the recycling of the TMA stages (full / empty mbarriers), the
// whole plane body is one basic block with RY+2P independent x rows and RY independent output chains for the scheduler.
// MODE / RSPLIT as in v4.  Used for p = 1 and p = 3 (at p = 5 the halo recomputation costs more than the hand-over).

template <int P_, int RY_, int NW_, int STAGES_, int MINB_>
struct Cfg6
{
  static constexpr bool V4 = true, V5 = false, V6 = true, V7 = false; // V4: tap split tables and plane-class z table of the plan
  static constexpr int  P = P_, TX = 32, RY = RY_, NW = NW_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int  NXW     = 0;
  static constexpr int  W       = 2 * P + 1;
  static constexpr int  TY      = RY * NW;
  static constexpr int  NR      = TY + 2 * P;   // rows of the staged tile
  static constexpr int  NIN     = RY + 2 * P;   // input rows per warp
  static constexpr int  PIN     = TX + 2 * P;   // pitch of the staged tile (dense TMA box); even -> 16-byte rows
  static constexpr int  THREADS = 32 * NW;
  static constexpr int  NBT     = 2 * (P + 1);
  static constexpr int  WP      = 8 * ((W + 7) / 8);
  static constexpr int  TB_DOUBLES = 2 * 2 * NBT * WP;
  static constexpr int  ZROWS      = 2 * W;
  static constexpr int  WZ         = W + 1;
  static constexpr int  ZT_DOUBLES = ZROWS * 2 * WZ;
  static constexpr int  STAGE_DOUBLES = (NR * PIN + 15) / 16 * 16;
};

template <class C>
constexpr size_t smem_bytes_v6()
{
  return (size_t)(C::STAGES * C::STAGE_DOUBLES + C::ZT_DOUBLES + C::TB_DOUBLES) * sizeof(double) + (size_t)(2 * C::STAGES) * sizeof(uint64_t) + 128;
}

template <class C, int MODE, bool RSPLIT, bool ACCUM>
__global__ void __launch_bounds__(C::THREADS, C::MINB) kron3d_v6_kernel(const __grid_constant__ CUtensorMap tmap, const KArgs<C::P> g)
{
  constexpr int  P = C::P, W = C::W, TX = C::TX, RY = C::RY, NR = C::NR, NIN = C::NIN, PIN = C::PIN, S = C::STAGES;
  constexpr bool HASB = MODE != 0, SYM = MODE == 1;
  constexpr int  PB = (RSPLIT && SYM) ? P - 1 : P; // outermost tap of the interior B rows in x and y
  extern __shared__ __align__(128) double smem[];
  constexpr int OFF_ZT  = S * C::STAGE_DOUBLES;
  constexpr int OFF_TB  = OFF_ZT + C::ZT_DOUBLES;
  constexpr int OFF_BAR = OFF_TB + C::TB_DOUBLES;
  constexpr int WP = C::WP, NBT = C::NBT, WZ = C::WZ;
  constexpr unsigned STAGE_BYTES = NR * PIN * sizeof(double);

  const uint32_t sb     = smem_u32(smem);
  const uint32_t bar_f  = sb + OFF_BAR * 8; // full[S]
  const uint32_t bar_e  = bar_f + 8 * S;    // empty[S]
  const int      tid    = threadIdx.x;
  const int      lane = tid & 31, warp = tid >> 5;
  int            b    = blockIdx.x;
  const int      tx   = b % g.tiles_x;
  b /= g.tiles_x;
  const int ty    = b % g.tiles_y;
  const int chunk = b / g.tiles_y;
  const int x0    = g.xorg + tx * TX;
  const int y0    = g.cy0 + ty * C::TY;
  const int zc0   = g.cz0 + chunk * g.lz;
  const int zc1   = min(zc0 + g.lz, g.cz1);
  const int kbeg = zc0 - P, kend = zc1 + P;

  if (tid == 0)
    {
      for (int s = 0; s < S; ++s)
        {
          mbar_init_a(bar_f + 8 * s, 1);
          mbar_init_a(bar_e + 8 * s, C::NW);
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
  if (tid == 0)
    for (int s = 0; s < S; ++s)
      if (kbeg + s < kend)
        {
          mbar_expect_tx_a(bar_f + 8 * s, STAGE_BYTES);
          tma_load_3d_a(sb + s * C::STAGE_DOUBLES * 8, &tmap, bar_f + 8 * s, x0 - P, y0 - P, kbeg + s);
        }

  // ---- ownership: lane -> x, warp -> RY consecutive rows
  const int gx       = x0 + lane;
  const int gy_first = y0 + warp * RY;
  const int nst      = (gx >= g.cx0 && gx < g.cx1) ? (g.cy1 - gy_first) : 0; // rows i < nst are stored
  double   *out      = g.dst + (int64_t)(kbeg - P) * g.plane + (int64_t)gy_first * g.pitch + gx;
  const int in_off   = warp * RY * PIN + lane; // this thread's first input value inside a stage
  // non-Toeplitz rows: x (per lane), y (per output row, warp uniform)
  const bool x_bnd   = (gx <= P || gx >= g.nx - P) && gx >= 0 && gx <= g.nx && !GDM_DBG(g, 16);
  const int  xrc     = (gx <= P) ? gx : gx - (g.nx - P) + P + 1;
  const int  xta_off = OFF_TB + (0 * NBT + (x_bnd ? xrc : 0)) * WP;
  const bool any_fix = (x0 <= P) || (x0 + TX - 1 >= g.nx - P) || (gy_first <= P) || (gy_first + RY - 1 >= g.ny - P);

  double acc[RY][2 * P];
#pragma unroll
  for (int i = 0; i < RY; ++i)
#pragma unroll
    for (int j = 0; j < 2 * P; ++j)
      acc[i][j] = 0.0;

  // ---- one input plane: x pass of RY+2P rows, y pass (gather) and z pass (scatter) of RY rows, store of plane k-P
  auto plane = [&](auto fix_c, auto toep_c, const int st_off, const int zc, const int nstore) {
    constexpr bool FIX = decltype(fix_c)::value, TOEP = decltype(toep_c)::value;
    constexpr bool OUTER = HASB && !(RSPLIT && TOEP);
    double         zA[W], zB[W];
    if (TOEP)
      {
#pragma unroll
        for (int j = 0; j < W; ++j)
          {
            zA[j] = g.Az[j];
            zB[j] = g.Bz[j];
          }
      }
    else
      {
#pragma unroll
        for (int j = 0; j < W; ++j)
          {
            zA[j] = smem[OFF_ZT + (zc * 2 + 0) * WZ + j];
            zB[j] = HASB ? smem[OFF_ZT + (zc * 2 + 1) * WZ + j] : 0.0;
          }
      }
    const double *tile = smem + st_off + in_off;
    double        a_[NIN], r_[NIN];
#pragma unroll
    for (int j = 0; j < NIN; ++j)
      {
        double v[W];
#pragma unroll
        for (int t = 0; t < W; ++t)
          v[t] = tile[j * PIN + t];
        double ra = g.Ax[0] * v[P];
        double rr = (HASB && SYM) ? g.Bx[0] * v[P] : 0.0;
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
                      rr = fma(g.Bx[d], s, rr);
                  }
                else
                  rr = fma(g.Bx[d], v[P + d] - v[P - d], rr);
              }
          }
        if (FIX)
          if (x_bnd)
            {
              ra = 0.0;
              rr = 0.0;
#pragma unroll
              for (int t = 0; t < W; ++t)
                {
                  ra = fma(smem[xta_off + t], v[t], ra);
                  if (HASB)
                    rr = fma(smem[xta_off + NBT * WP + t], v[t], rr);
                }
            }
        a_[j] = ra;
        r_[j] = rr;
        if (j >= 2 * P)
          {
            const int i = j - 2 * P; // output row: window rows i .. i+2P, centre i+P
            const int c = i + P;
            double    p2 = g.Ay[0] * a_[c], s2 = 0.0;
            if (HASB)
              {
                s2 = g.Ay[0] * r_[c];
                if (SYM)
                  s2 = fma(g.By[0], a_[c], s2);
              }
#pragma unroll
            for (int d = 1; d <= P; ++d)
              {
                const double sa = a_[c - d] + a_[c + d];
                p2              = fma(g.Ay[d], sa, p2);
                if (HASB)
                  {
                    const double sr = r_[c - d] + r_[c + d];
                    s2              = fma(g.Ay[d], sr, s2);
                    if (SYM)
                      {
                        if (d <= PB)
                          s2 = fma(g.By[d], sa, s2);
                      }
                    else
                      s2 = fma(g.By[d], a_[c + d] - a_[c - d], s2);
                  }
              }
            if (FIX)
              {
                const int gy = gy_first + i;
                if ((gy <= P || gy >= g.ny - P) && gy <= g.ny)
                  {
                    const int     rc = (gy <= P) ? gy : gy - (g.ny - P) + P + 1;
                    const double *ta = smem + OFF_TB + (2 * NBT + rc) * WP;
                    const double *tb = smem + OFF_TB + (3 * NBT + rc) * WP;
                    p2               = 0.0;
                    s2               = 0.0;
#pragma unroll
                    for (int t = 0; t < W; ++t)
                      {
                        const double ca = ta[t];
                        p2              = fma(ca, a_[i + t], p2);
                        if (HASB)
                          {
                            s2 = fma(ca, r_[i + t], s2);
                            s2 = fma(tb[t], a_[i + t], s2);
                          }
                      }
                  }
              }
            // z pass (scatter form)
            double ua = p2;
            if (HASB)
              ua = RSPLIT ? fma(g.sigma, p2, s2) : s2;
            double res = fma(zA[0], ua, acc[i][0]);
            if (OUTER)
              res = fma(zB[0], p2, res);
#pragma unroll
            for (int jz = 1; jz < 2 * P; ++jz)
              {
                double s = fma(zA[jz], ua, acc[i][jz]);
                if (HASB)
                  s = fma(zB[jz], p2, s);
                acc[i][jz - 1] = s;
              }
            double s = zA[2 * P] * ua;
            if (OUTER)
              s = fma(zB[2 * P], p2, s);
            acc[i][2 * P - 1] = s;
            if (i < nstore)
              {
                double *o = out + (int64_t)i * g.pitch;
                if (ACCUM)
                  res += *o;
                *o = res;
              }
          }
      }
  };

  int      st = 0, st_off = 0;
  unsigned par = 0;
  // issuer (thread 0): next plane to request, the stage it goes to (the one that held plane kn - S) and that stage's
  // release parity; a stage is refilled as soon as every warp has released it (non-blocking probe, never a stall)
  int      kn = kbeg + S, rs = 0;
  unsigned rpar = 0;
  auto     try_refill = [&]() {
    while (kn < kend)
      {
        if (!mbar_test_a(bar_e + 8 * rs, rpar))
          break;
        mbar_expect_tx_a(bar_f + 8 * rs, STAGE_BYTES);
        tma_load_3d_a(sb + rs * C::STAGE_DOUBLES * 8, &tmap, bar_f + 8 * rs, x0 - P, y0 - P, kn);
        ++kn;
        if (++rs == S)
          {
            rs = 0;
            rpar ^= 1u;
          }
      }
  };
  for (int k = kbeg; k < kend; ++k)
    {
      // thread 0 is the only TMA issuer: it must never block while a stage can be refilled (it could be waiting for a
      // plane that only itself can request), so it polls its own full barrier and keeps refilling meanwhile
      if (tid == 0)
        {
          try_refill();
          while (!mbar_test_a(bar_f + 8 * st, par))
            try_refill();
        }
      mbar_wait_a(bar_f + 8 * st, par);
      const int  nstore = (k - P >= zc0 && !GDM_DBG(g, 1)) ? nst : 0;
      const bool toep   = (k >= g.kz_lo && k < g.kz_hi);
      int        zc     = 0;
      if (!toep)
        {
          // plane class: planes below kz_lo by index, planes from kz_hi on after them; planes outside the slab carry zeros
          const int kk = min(max(k, 0), g.nz_local - 1);
          zc           = (kk < g.kz_lo) ? kk : g.kz_lo + (kk - g.kz_hi);
        }
      if (!GDM_DBG(g, 8))
        {
          if (!any_fix)
            {
              if (toep)
                plane(std::false_type{}, std::true_type{}, st_off, 0, nstore);
              else
                plane(std::false_type{}, std::false_type{}, st_off, zc, nstore);
            }
          else
            {
              if (toep)
                plane(std::true_type{}, std::true_type{}, st_off, 0, nstore);
              else
                plane(std::true_type{}, std::false_type{}, st_off, zc, nstore);
            }
        }
      __syncwarp();
      if (lane == 0)
        mbar_arrive_a(bar_e + 8 * st);
      st_off += C::STAGE_DOUBLES;
      if (++st == S)
        {
          st     = 0;
          st_off = 0;
          par ^= 1u;
        }
      out += g.plane;
    }
}
