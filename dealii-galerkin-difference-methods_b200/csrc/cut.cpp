// Cut-cell set-up (host side, setup only): level-set classification, cut quadrature, CutFEM Poisson rows.
//
// The step BEFORE the hot path in BASELINE configuration 5: it produces the irregular CSR rows that
// gdm_operator_attach_csr lays over the tensor-product stiffness apply, the right-hand side and the error norm.
// Stands in for
//   * NonMatching::MeshClassifier for a Q1 level set (prototypes/cut_poisson_01_gdm.cc:105-121,
//     applications/wave/include/gdm/wave/discretization.h:79-97): inside / outside / intersected from the vertex signs;
//   * NonMatching::FEValues / QuadratureGenerator (prototypes/cut_poisson_01_gdm.cc:176-190): height-function recursion
//     on the unit cell, QGauss<1>(p+1) per direction.  The cell's level set is multilinear, so the bounds the general
//     algorithm estimates are exact here: a multilinear function and its partial derivatives take their extrema at the
//     vertices, and a coordinate line crosses the zero set at most once (the root is found in closed form);
//   * the assembly loop prototypes/cut_poisson_01_gdm.cc:196-329: volume term on the inside part, symmetric Nitsche
//     terms on the surface, ghost penalty on the faces between an intersected cell and a non-outside neighbour
//     (:123-146), zero diagonal -> 1 (:324-329); gp_h_power = 3 gives the scaling of the wave application's matrix
//     (applications/wave/include/gdm/wave/stiffness.h:760-765);
//   * the inside L2 error prototypes/cut_poisson_01_gdm.cc:349-398.
// Only the rows that differ from the plain stiffness operator are assembled: rows of DoFs in the window of a cell that
// is not inside or of a ghost-penalty face ("band rows", a dense (2p+3)^dim box of column offsets per row while
// assembling) and identity rows of DoFs no active cell touches.
#include <algorithm>
#include <cmath>
#include <array>
#include <cstring>
#include <chrono>
#include <cstdlib>
#include <map>
#include <thread>

#include "gdm_internal.h"

namespace gdm
{
  namespace cut
  {
    enum : uint8_t
    {
      INSIDE      = 0,
      OUTSIDE     = 1,
      INTERSECTED = 2
    };

    // multilinear function on a box of dimension d: corner values, bit e of the index = upper end in direction e
    struct MLF
    {
      int    d;
      double c[8];
    };

    struct Pt
    {
      double x[3];
      double w;
      double n[3];
    };

    static double mlf_min(const MLF &f)
    {
      double m = f.c[0];
      for (int i = 1; i < (1 << f.d); ++i)
        m = std::min(m, f.c[i]);
      return m;
    }
    static double mlf_max(const MLF &f)
    {
      double m = f.c[0];
      for (int i = 1; i < (1 << f.d); ++i)
        m = std::max(m, f.c[i]);
      return m;
    }
    static MLF face(const MLF &f, int k, int side)
    {
      MLF r;
      r.d = f.d - 1;
      for (int i = 0; i < (1 << r.d); ++i)
        {
          const int low = i & ((1 << k) - 1), high = i >> k;
          r.c[i] = f.c[low | (side << k) | (high << (k + 1))];
        }
      return r;
    }
    static double eval(const MLF &f, const double *t)
    {
      double tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < (1 << f.d); ++i)
        tmp[i] = f.c[i];
      for (int e = 0; e < f.d; ++e)
        for (int i = 0; i < (1 << (f.d - e - 1)); ++i)
          tmp[i] = tmp[2 * i] * (1.0 - t[e]) + tmp[2 * i + 1] * t[e];
      return tmp[0];
    }
    static void split(const MLF &f, int e, MLF &lo, MLF &hi)
    {
      lo = f;
      hi = f;
      for (int i = 0; i < (1 << f.d); ++i)
        if (!(i & (1 << e)))
          {
            const double mid = 0.5 * (f.c[i] + f.c[i | (1 << e)]);
            lo.c[i | (1 << e)] = mid;
            hi.c[i]            = mid;
          }
    }
    // d_k f on the box: difference of the two faces, min and max over the corners
    static void dk_range(const MLF &f, int k, double &mn, double &mx)
    {
      const MLF a = face(f, k, 0), b = face(f, k, 1);
      mn = mx = b.c[0] - a.c[0];
      for (int i = 1; i < (1 << a.d); ++i)
        {
          mn = std::min(mn, b.c[i] - a.c[i]);
          mx = std::max(mx, b.c[i] - a.c[i]);
        }
    }
    // direction in which every function is strictly monotone over the box; the one with the largest worst-case
    // |d_k f| relative to the gradient; -1 if there is none
    static int height_direction(const std::vector<MLF> &funcs)
    {
      const int d    = funcs[0].d;
      int       best = -1;
      double    best_score = 0.0;
      for (int k = 0; k < d; ++k)
        {
          double score = INFINITY;
          for (const MLF &f : funcs)
            {
              double mn, mx;
              dk_range(f, k, mn, mx);
              if (!(mn > 0 || mx < 0))
                {
                  score = 0.0;
                  break;
                }
              double tot = 0;
              for (int e = 0; e < d; ++e)
                {
                  double a, b;
                  dk_range(f, e, a, b);
                  tot += std::max(std::fabs(a), std::fabs(b));
                }
              score = std::min(score, std::min(std::fabs(mn), std::fabs(mx)) / tot);
            }
          if (score > best_score)
            {
              best       = k;
              best_score = score;
            }
        }
      return best;
    }

    struct Gauss
    {
      std::vector<double> x, w;
    };

    static void tensor_gauss(const double *lo, const double *hi, int d, const Gauss &g, std::vector<Pt> &out)
    {
      const int n     = (int)g.x.size();
      int       total = 1;
      for (int e = 0; e < d; ++e)
        total *= n;
      for (int q = 0; q < total; ++q)
        {
          Pt  p{};
          int r = q;
          p.w   = 1.0;
          for (int e = 0; e < d; ++e)
            {
              const int i = r % n;
              r /= n;
              p.x[e] = lo[e] + (hi[e] - lo[e]) * g.x[i];
              p.w *= (hi[e] - lo[e]) * g.w[i];
            }
          out.push_back(p);
        }
    }

    // 1D Gauss rules in direction k over every base point, on the sub-intervals between the roots where all sign
    // conditions hold
    static void line(const std::vector<MLF> &funcs, const std::vector<int> &signs, const std::vector<Pt> &base, int k,
                     const double *lo, const double *hi, int d, const Gauss &g, std::vector<Pt> &out)
    {
      int rest[3], nr = 0;
      for (int e = 0; e < d; ++e)
        if (e != k)
          rest[nr++] = e;
      const double        L = hi[k] - lo[k];
      std::vector<double> a(funcs.size()), b(funcs.size()), roots;
      for (const Pt &bp : base)
        {
          double tl[3] = {0, 0, 0};
          for (int j = 0; j < nr; ++j)
            tl[j] = (bp.x[j] - lo[rest[j]]) / (hi[rest[j]] - lo[rest[j]]);
          roots.assign({0.0, 1.0});
          for (size_t i = 0; i < funcs.size(); ++i)
            {
              a[i] = eval(face(funcs[i], k, 0), tl);
              b[i] = eval(face(funcs[i], k, 1), tl);
              if (a[i] * b[i] < 0)
                roots.push_back(a[i] / (a[i] - b[i]));
            }
          std::sort(roots.begin(), roots.end());
          for (size_t s = 0; s + 1 < roots.size(); ++s)
            {
              const double r0 = roots[s], r1 = roots[s + 1];
              if (r1 - r0 <= 1e-14)
                continue;
              const double m  = 0.5 * (r0 + r1);
              bool         ok = true;
              for (size_t i = 0; i < funcs.size() && ok; ++i)
                ok = signs[i] == 0 || signs[i] * (a[i] + (b[i] - a[i]) * m) > 0;
              if (!ok)
                continue;
              for (size_t q = 0; q < g.x.size(); ++q)
                {
                  Pt p{};
                  for (int j = 0; j < nr; ++j)
                    p.x[rest[j]] = bp.x[j];
                  p.x[k] = lo[k] + L * (r0 + (r1 - r0) * g.x[q]);
                  p.w    = bp.w * L * (r1 - r0) * g.w[q];
                  out.push_back(p);
                }
            }
        }
    }

    // quadrature of {x in box : s_i f_i(x) > 0 for all i with s_i != 0}, partitioned along the zero sets of all f_i
    static void volume(std::vector<MLF> funcs, std::vector<int> signs, const double *lo, const double *hi, int d,
                       const Gauss &g, int depth, std::vector<Pt> &out)
    {
      {
        std::vector<MLF> kf;
        std::vector<int> ks;
        for (size_t i = 0; i < funcs.size(); ++i)
          {
            if (mlf_min(funcs[i]) > 0)
              {
                if (signs[i] < 0)
                  return;
              }
            else if (mlf_max(funcs[i]) < 0)
              {
                if (signs[i] > 0)
                  return;
              }
            else
              {
                kf.push_back(funcs[i]);
                ks.push_back(signs[i]);
              }
          }
        funcs.swap(kf);
        signs.swap(ks);
      }
      if (funcs.empty())
        {
          tensor_gauss(lo, hi, d, g, out);
          return;
        }
      if (d == 1)
        {
          Pt b{};
          b.w = 1.0;
          line(funcs, signs, {b}, 0, lo, hi, 1, g, out);
          return;
        }
      const int k = height_direction(funcs);
      if (k < 0)
        {
          if (depth >= 16)
            { // give up: plain Gauss points, sign test per point
              std::vector<Pt> pts;
              tensor_gauss(lo, hi, d, g, pts);
              for (const Pt &p : pts)
                {
                  double t[3];
                  for (int e = 0; e < d; ++e)
                    t[e] = (p.x[e] - lo[e]) / (hi[e] - lo[e]);
                  bool ok = true;
                  for (size_t i = 0; i < funcs.size() && ok; ++i)
                    ok = signs[i] == 0 || signs[i] * eval(funcs[i], t) > 0;
                  if (ok)
                    out.push_back(p);
                }
              return;
            }
          int e = 0;
          for (int j = 1; j < d; ++j)
            if (hi[j] - lo[j] > hi[e] - lo[e])
              e = j;
          std::vector<MLF> f0(funcs.size()), f1(funcs.size());
          for (size_t i = 0; i < funcs.size(); ++i)
            split(funcs[i], e, f0[i], f1[i]);
          double hi0[3], lo1[3];
          for (int j = 0; j < d; ++j)
            {
              hi0[j] = hi[j];
              lo1[j] = lo[j];
            }
          hi0[e] = lo1[e] = 0.5 * (lo[e] + hi[e]);
          volume(f0, signs, lo, hi0, d, g, depth + 1, out);
          volume(f1, signs, lo1, hi, d, g, depth + 1, out);
          return;
        }
      std::vector<MLF> bf;
      std::vector<int> bs;
      for (size_t i = 0; i < funcs.size(); ++i)
        {
          double mn, mx;
          dk_range(funcs[i], k, mn, mx);
          const int gsign = mn > 0 ? 1 : -1, s = signs[i];
          // the column over a base point meets {s f > 0} iff s f > 0 on the face where s f is largest
          bf.push_back(face(funcs[i], k, 0));
          bs.push_back(s * gsign < 0 ? s : 0);
          bf.push_back(face(funcs[i], k, 1));
          bs.push_back(s * gsign > 0 ? s : 0);
        }
      double blo[3], bhi[3];
      int    nr = 0;
      for (int e = 0; e < d; ++e)
        if (e != k)
          {
            blo[nr] = lo[e];
            bhi[nr] = hi[e];
            ++nr;
          }
      std::vector<Pt> base;
      volume(bf, bs, blo, bhi, d - 1, g, depth, base);
      line(funcs, signs, base, k, lo, hi, d, g, out);
    }

    // quadrature of {f = 0} inside the box: points, weights, unit normals grad f / |grad f| (unit-cell coordinates)
    static void surface(const MLF &f, const double *lo, const double *hi, int d, const Gauss &g, int depth,
                        std::vector<Pt> &out)
    {
      if (mlf_min(f) > 0 || mlf_max(f) < 0)
        return;
      if (d == 1)
        {
          const double a = f.c[0], b = f.c[1];
          if (a * b >= 0)
            return;
          Pt p{};
          p.x[0] = lo[0] + (hi[0] - lo[0]) * a / (a - b);
          p.w    = 1.0;
          p.n[0] = b > a ? 1.0 : -1.0;
          out.push_back(p);
          return;
        }
      const int k = height_direction({f});
      if (k < 0)
        {
          if (depth >= 16)
            return;
          int e = 0;
          for (int j = 1; j < d; ++j)
            if (hi[j] - lo[j] > hi[e] - lo[e])
              e = j;
          MLF f0, f1;
          split(f, e, f0, f1);
          double hi0[3], lo1[3];
          for (int j = 0; j < d; ++j)
            {
              hi0[j] = hi[j];
              lo1[j] = lo[j];
            }
          hi0[e] = lo1[e] = 0.5 * (lo[e] + hi[e]);
          surface(f0, lo, hi0, d, g, depth + 1, out);
          surface(f1, lo1, hi, d, g, depth + 1, out);
          return;
        }
      double mn, mx;
      dk_range(f, k, mn, mx);
      const int gsign = mn > 0 ? 1 : -1;
      int       rest[3], nr = 0;
      double    blo[3], bhi[3];
      for (int e = 0; e < d; ++e)
        if (e != k)
          {
            rest[nr] = e;
            blo[nr]  = lo[e];
            bhi[nr]  = hi[e];
            ++nr;
          }
      std::vector<Pt> base;
      volume({face(f, k, 0), face(f, k, 1)}, {-gsign, gsign}, blo, bhi, d - 1, g, 0, base);
      for (const Pt &bp : base)
        {
          double tl[3] = {0, 0, 0};
          for (int j = 0; j < nr; ++j)
            tl[j] = (bp.x[j] - blo[j]) / (bhi[j] - blo[j]);
          const double a = eval(face(f, k, 0), tl), b = eval(face(f, k, 1), tl);
          if (a * b >= 0)
            continue;
          double t[3] = {0, 0, 0};
          for (int j = 0; j < nr; ++j)
            t[rest[j]] = tl[j];
          t[k] = a / (a - b);
          double grad[3] = {0, 0, 0}, gn = 0;
          for (int e = 0; e < d; ++e)
            {
              double te[3];
              int    m = 0;
              for (int j = 0; j < d; ++j)
                if (j != e)
                  te[m++] = t[j];
              grad[e] = (eval(face(f, e, 1), te) - eval(face(f, e, 0), te)) / (hi[e] - lo[e]);
              gn += grad[e] * grad[e];
            }
          gn = std::sqrt(gn);
          Pt p{};
          for (int e = 0; e < d; ++e)
            {
              p.x[e] = lo[e] + (hi[e] - lo[e]) * t[e];
              p.n[e] = grad[e] / gn;
            }
          p.w = bp.w * gn / std::fabs(grad[k]);
          out.push_back(p);
        }
    }

    static Gauss make_gauss(int n)
    {
      std::vector<long double> x, w;
      gauss_legendre_01(n, x, w);
      Gauss g;
      g.x.assign(x.begin(), x.end());
      g.w.assign(w.begin(), w.end());
      return g;
    }

    static void cut_quadrature(int dim, const double *vertex_values, const Gauss &g, std::vector<Pt> &inside,
                               std::vector<Pt> &surf)
    {
      MLF f;
      f.d = dim;
      for (int i = 0; i < (1 << dim); ++i)
        f.c[i] = vertex_values[i];
      const double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
      volume({f}, {-1}, lo, hi, dim, g, 0, inside);
      surface(f, lo, hi, dim, g, 0, surf);
    }

    // ------------------------------------------------------------------------------------------------ assembly
    struct Assembly
    {
      gdm_cut_desc desc;
      int          dim, p, npc;
      int          N[3], nn[3];
      double       h[3], lo[3];
      uint64_t     n_dofs, n_cells;
      Gauss        gauss;
      std::vector<double>   ls;
      std::vector<uint8_t>  location;
      std::vector<uint64_t> row_ids, rowptr, col;
      std::vector<double>   val, rhs;
      uint64_t              n_band_rows = 0, n_identity_rows = 0, counts[3] = {0, 0, 0};

      void cell_index(uint64_t cell, int *idx) const
      {
        for (int e = 0; e < 3; ++e)
          idx[e] = 0;
        idx[0] = (int)(cell % N[0]);
        if (dim >= 2)
          idx[1] = (int)((cell / N[0]) % N[1]);
        if (dim >= 3)
          idx[2] = (int)(cell / ((uint64_t)N[0] * N[1]));
      }
      uint64_t cell_of(const int *idx) const
      {
        return (uint64_t)idx[0] + (uint64_t)N[0] * ((dim >= 2 ? idx[1] : 0) + (uint64_t)(dim >= 2 ? N[1] : 1) * (dim >= 3 ? idx[2] : 0));
      }
      uint64_t node_of(const int *i) const
      {
        return (uint64_t)i[0] + (uint64_t)nn[0] * ((dim >= 2 ? i[1] : 0) + (uint64_t)(dim >= 2 ? nn[1] : 1) * (dim >= 3 ? i[2] : 0));
      }
      void vertex_values(const int *idx, double *v) const
      {
        for (int c = 0; c < (1 << dim); ++c)
          {
            int node[3] = {0, 0, 0};
            for (int e = 0; e < dim; ++e)
              node[e] = idx[e] + ((c >> e) & 1);
            v[c] = ls[node_of(node)];
          }
      }
      // window offsets per direction and the global DoFs of the cell, lexicographic with x fastest (system.h:195-246)
      void cell_dofs(const int *idx, int *off, std::vector<uint64_t> &dofs) const
      {
        for (int e = 0; e < 3; ++e)
          off[e] = e < dim ? window_offset(p, N[e], idx[e]) : 0;
        dofs.resize(npc);
        int m = 0;
        for (int k = 0; k < (dim >= 3 ? p + 1 : 1); ++k)
          for (int j = 0; j < (dim >= 2 ? p + 1 : 1); ++j)
            for (int i = 0; i <= p; ++i)
              {
                const int node[3] = {off[0] + i, off[1] + j, off[2] + k};
                dofs[m++]         = node_of(node);
              }
      }
      int category(const int *idx) const
      {
        int c = 0, s = 1;
        for (int e = 0; e < dim; ++e)
          {
            c += cell_variant(p, N[e], idx[e]) * s;
            s *= p;
          }
        return c;
      }
      // values[q*npc + i] and physical gradients grads[e][q*npc + i] of the cell's basis at unit-cell points
      void shape_at_points(const int *idx, const std::vector<Pt> &pts, std::vector<double> &value,
                           std::vector<double> grads[3]) const
      {
        const size_t nq = pts.size();
        value.assign(nq * npc, 0.0);
        for (int e = 0; e < dim; ++e)
          grads[e].assign(nq * npc, 0.0);
        long double v1[3][MAX_DEGREE + 1], d1[3][MAX_DEGREE + 1];
        for (size_t q = 0; q < nq; ++q)
          {
            for (int e = 0; e < dim; ++e)
              lagrange_eval(p, cell_variant(p, N[e], idx[e]), pts[q].x[e], v1[e], d1[e]);
            int m = 0;
            for (int k = 0; k < (dim >= 3 ? p + 1 : 1); ++k)
              for (int j = 0; j < (dim >= 2 ? p + 1 : 1); ++j)
                for (int i = 0; i <= p; ++i, ++m)
                  {
                    const int ii[3] = {i, j, k};
                    long double v   = 1;
                    for (int e = 0; e < dim; ++e)
                      v *= v1[e][ii[e]];
                    value[q * npc + m] = (double)v;
                    for (int dd = 0; dd < dim; ++dd)
                      {
                        long double gq = 1;
                        for (int e = 0; e < dim; ++e)
                          gq *= (e == dd) ? d1[e][ii[e]] / h[e] : v1[e][ii[e]];
                        grads[dd][q * npc + m] = (double)gq;
                      }
                  }
          }
      }
      double cell_volume() const
      {
        double v = 1;
        for (int e = 0; e < dim; ++e)
          v *= h[e];
        return v;
      }
      double h_min() const
      {
        double v = h[0];
        for (int e = 1; e < dim; ++e)
          v = std::min(v, h[e]);
        return v;
      }
      // neighbour across face (d, side) if the face carries the ghost penalty, else -1
      int64_t ghost_penalty_neighbor(const int *idx, int d, int side) const
      {
        int nb[3] = {idx[0], idx[1], idx[2]};
        nb[d] += side ? 1 : -1;
        if (nb[d] < 0 || nb[d] >= N[d])
          return -1;
        const uint64_t nc = cell_of(nb);
        const uint8_t  a = location[cell_of(idx)], b = location[nc];
        if ((a == INTERSECTED && b != OUTSIDE) || (b == INTERSECTED && a != OUTSIDE))
          return (int64_t)nc;
        return -1;
      }

      struct CutScratch
      {
        std::vector<Pt>     ipts, spts;
        std::vector<double> value, grads[3];
      };
      // ---- level set of degree q > 1 (2D): values on the Gauss-Lobatto refined grid, (q N_e + 1) points per direction
      int                 q = 1;
      std::vector<double> gll, to_bernstein; // support points of FE_Q(q) on [0,1]; (q+1)^2 Lagrange -> Bernstein
      void    setup_q();
      void    local_values_q(const int *idx, double *c) const; // c[i + (q+1) j] at (gll_i, gll_j) of the cell
      double  eval_q(const double *c, double x, double y, double *grad) const;
      uint8_t classify_cell(const int *idx) const;
      void    rules_q(const int *idx, std::vector<Pt> &inside, std::vector<Pt> &surf) const;
      // inside and surface rules of a cell for either kind of level set
      void cell_rules(const int *idx, std::vector<Pt> &inside, std::vector<Pt> &surf) const
      {
        if (q > 1)
          {
            rules_q(idx, inside, surf);
            return;
          }
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        vertex_values(idx, v);
        cut_quadrature(dim, v, gauss, inside, surf);
      }
      void cut_cell_matrix(const int *idx, CutScratch &sc, double *local, double *lrhs) const;
      // part of the face (d, side) of a cell on the box boundary where the level set is negative, as unit-cell points
      // with face weights (NonMatching::FEInterfaceValues::reinit(cell, f), wave/stiffness.h:268-283)
      void boundary_face_rule(const int *idx, int d, int side, std::vector<Pt> &pts) const;
      bool at_box_boundary(const int *idx) const
      {
        for (int e = 0; e < dim; ++e)
          if (idx[e] == 0 || idx[e] == N[e] - 1)
            return true;
        return false;
      }
      // Nitsche terms on the box boundary (wave/stiffness.h:262-340, term IV): adds the cell matrix to local
      // (npc x npc, may be null) and, for a boundary function g, the load  <gamma_D / h v - d_n v, g>  to lrhs (may be null)
      void boundary_terms(const int *idx, CutScratch &sc, double *local, gdm_function_fn g, void *user, double *lrhs) const;
      void coupling(int which, std::vector<uint64_t> &rows, std::vector<uint64_t> &rp, std::vector<uint64_t> &cols,
                    std::vector<double> &vals) const;
      void build();
      void load_vector(gdm_function_fn f, void *f_user, gdm_function_fn g, void *g_user, double *out);

      // data-independent part of the load functionals on the intersected cells (built on first use)
      struct CutCellLoad
      {
        uint64_t            cell;
        std::vector<double> vpts, vW, spts, sW; // points [q*dim + e], weights [q*npc + i]
      };
      std::vector<CutCellLoad> cut_loads;
      bool                     cut_loads_built = false;
    };


    // dense cell matrix and load vector of an intersected cell (prototypes/cut_poisson_01_gdm.cc:212-283); a pure
    // function of the cell, evaluated by several threads
    void Assembly::cut_cell_matrix(const int *idx, CutScratch &sc, double *local, double *lrhs) const
    {
      const bool   mass = desc.kind == 1;
      const double vol = cell_volume(), nitsche = desc.nitsche_parameter / h_min();
      sc.ipts.clear();
      sc.spts.clear();
      cell_rules(idx, sc.ipts, sc.spts);
      std::fill(local, local + (size_t)npc * npc, 0.0);
      std::fill(lrhs, lrhs + npc, 0.0);
      if (!sc.ipts.empty())
        {
          shape_at_points(idx, sc.ipts, sc.value, sc.grads);
          // upper triangle only (the products commute, so the mirrored entries carry the same bits), direction sums
          // written out so that the j loop vectorises
          for (size_t q = 0; q < sc.ipts.size(); ++q)
            {
              const double  jxw = sc.ipts[q].w * vol;
              const double *vq = sc.value.data() + q * npc, *a0 = sc.grads[0].data() + q * npc,
                           *a1 = dim > 1 ? sc.grads[1].data() + q * npc : nullptr,
                           *a2 = dim > 2 ? sc.grads[2].data() + q * npc : nullptr;
              for (int i = 0; i < npc; ++i)
                {
                  lrhs[i] += desc.rhs_value * vq[i] * jxw;
                  double *row = local + (size_t)i * npc;
                  if (mass)
                    for (int j = i; j < npc; ++j)
                      row[j] += (vq[i] * vq[j]) * jxw;
                  else if (dim == 1)
                    for (int j = i; j < npc; ++j)
                      row[j] += (a0[i] * a0[j]) * jxw;
                  else if (dim == 2)
                    for (int j = i; j < npc; ++j)
                      row[j] += (a0[i] * a0[j] + a1[i] * a1[j]) * jxw;
                  else
                    for (int j = i; j < npc; ++j)
                      row[j] += ((a0[i] * a0[j] + a1[i] * a1[j]) + a2[i] * a2[j]) * jxw;
                }
            }
          for (int i = 1; i < npc; ++i)
            for (int j = 0; j < i; ++j)
              local[(size_t)i * npc + j] = local[(size_t)j * npc + i];
        }
      if (!sc.spts.empty() && !mass && !desc.no_surface_terms)
        {
          shape_at_points(idx, sc.spts, sc.value, sc.grads);
          std::vector<double> ng(npc);
          for (size_t q = 0; q < sc.spts.size(); ++q)
            {
              // unit-cell normal and measure -> physical (anisotropic spacing allowed)
              double nph[3] = {0, 0, 0}, scale = 0;
              for (int e = 0; e < dim; ++e)
                {
                  nph[e] = sc.spts[q].n[e] / h[e];
                  scale += nph[e] * nph[e];
                }
              scale = std::sqrt(scale);
              for (int e = 0; e < dim; ++e)
                nph[e] /= scale;
              const double jxw = sc.spts[q].w * vol * scale;
              for (int i = 0; i < npc; ++i)
                {
                  double s = 0;
                  for (int e = 0; e < dim; ++e)
                    s += nph[e] * sc.grads[e][q * npc + i];
                  ng[i] = s;
                }
              for (int i = 0; i < npc; ++i)
                {
                  const double vi = sc.value[q * npc + i];
                  lrhs[i] += desc.boundary_value * (nitsche * vi - ng[i]) * jxw;
                  for (int j = 0; j < npc; ++j)
                    {
                      const double vj = sc.value[q * npc + j];
                      local[(size_t)i * npc + j] += (-ng[i] * vj - ng[j] * vi + nitsche * vi * vj) * jxw;
                    }
                }
            }
        }
    }



    // ------------------------------------------------------------- level set of degree q (2D presets of applications/wave)
    void Assembly::setup_q()
    {
      // Gauss-Lobatto points: +-1 and the roots of P_q' (Newton on the Legendre recurrence)
      gll.assign(q + 1, 0.0);
      gll[q] = 1.0;
      const double pi = 3.14159265358979323846;
      for (int k = 1; k < q; ++k)
        {
          double x = -std::cos(pi * k / q);
          for (int it = 0; it < 100; ++it)
            {
              // P_q, P_q' and P_q'' at x
              double p0 = 1, p1 = x;
              for (int j = 2; j <= q; ++j)
                {
                  const double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
                  p0 = p1;
                  p1 = p2;
                }
              const double dp  = q * (x * p1 - p0) / (x * x - 1);
              const double ddp = (2 * x * dp - q * (q + 1) * p1) / (1 - x * x);
              const double dx  = dp / ddp;
              x -= dx;
              if (std::fabs(dx) < 1e-16)
                break;
            }
          gll[k] = 0.5 * (x + 1.0);
        }
      // B[j][i] = C(q, i) t_j^i (1 - t_j)^(q - i); to_bernstein = B^-1 (Gauss-Jordan)
      const int           m = q + 1;
      std::vector<double> B(m * m), I(m * m, 0.0);
      for (int j = 0; j < m; ++j)
        {
          double binom = 1;
          for (int i = 0; i < m; ++i)
            {
              B[j * m + i] = binom * std::pow(gll[j], i) * std::pow(1 - gll[j], q - i);
              binom        = binom * (q - i) / (i + 1);
            }
          I[j * m + j] = 1.0;
        }
      for (int c = 0; c < m; ++c)
        {
          int piv = c;
          for (int r = c + 1; r < m; ++r)
            if (std::fabs(B[r * m + c]) > std::fabs(B[piv * m + c]))
              piv = r;
          for (int k = 0; k < m; ++k)
            {
              std::swap(B[c * m + k], B[piv * m + k]);
              std::swap(I[c * m + k], I[piv * m + k]);
            }
          const double d = B[c * m + c];
          for (int k = 0; k < m; ++k)
            {
              B[c * m + k] /= d;
              I[c * m + k] /= d;
            }
          for (int r = 0; r < m; ++r)
            if (r != c)
              {
                const double f = B[r * m + c];
                for (int k = 0; k < m; ++k)
                  {
                    B[r * m + k] -= f * B[c * m + k];
                    I[r * m + k] -= f * I[c * m + k];
                  }
              }
        }
      to_bernstein = I;
    }

    void Assembly::local_values_q(const int *idx, double *c) const
    {
      const uint64_t nx = (uint64_t)q * N[0] + 1;
      for (int j = 0; j <= q; ++j)
        for (int i = 0; i <= q; ++i)
          c[i + (q + 1) * j] = ls[(uint64_t)(q * idx[0] + i) + nx * (uint64_t)(q * idx[1] + j)];
    }

    // psi(x, y) = sum c_ij l_i(x) l_j(y) and its gradient (unit-cell coordinates)
    double Assembly::eval_q(const double *c, double x, double y, double *grad) const
    {
      double lx[MAX_DEGREE + 1], ly[MAX_DEGREE + 1], dx[MAX_DEGREE + 1], dy[MAX_DEGREE + 1];
      for (int pass = 0; pass < 2; ++pass)
        {
          const double t = pass ? y : x;
          double      *l = pass ? ly : lx, *d = pass ? dy : dx;
          for (int k = 0; k <= q; ++k)
            {
              double denom = 1, val = 1, der = 0;
              for (int j = 0; j <= q; ++j)
                if (j != k)
                  {
                    denom *= gll[k] - gll[j];
                    val *= t - gll[j];
                  }
              for (int m = 0; m <= q; ++m)
                if (m != k)
                  {
                    double pr = 1;
                    for (int j = 0; j <= q; ++j)
                      if (j != k && j != m)
                        pr *= t - gll[j];
                    der += pr;
                  }
              l[k] = val / denom;
              d[k] = der / denom;
            }
        }
      double v = 0, gx = 0, gy = 0;
      for (int j = 0; j <= q; ++j)
        for (int i = 0; i <= q; ++i)
          {
            const double cij = c[i + (q + 1) * j];
            v += cij * lx[i] * ly[j];
            gx += cij * dx[i] * ly[j];
            gy += cij * lx[i] * dy[j];
          }
      if (grad)
        {
          grad[0] = gx;
          grad[1] = gy;
        }
      return v;
    }

    // NonMatching::MeshClassifier: a face is inside / outside when all Bernstein coefficients of the level set on it are
    // negative / positive; a cell is inside / outside when all its faces are, intersected otherwise
    uint8_t Assembly::classify_cell(const int *idx) const
    {
      if (q <= 1)
        {
          double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          vertex_values(idx, v);
          double mn = v[0], mx = v[0];
          for (int c = 1; c < (1 << dim); ++c)
            {
              mn = std::min(mn, v[c]);
              mx = std::max(mx, v[c]);
            }
          return mx < 0 ? INSIDE : (mn > 0 ? OUTSIDE : INTERSECTED);
        }
      const int m = q + 1;
      double    c[(MAX_DEGREE + 1) * (MAX_DEGREE + 1)];
      local_values_q(idx, c);
      int n_in = 0, n_out = 0;
      for (int f = 0; f < 4; ++f)
        {
          double mn = INFINITY, mx = -INFINITY;
          for (int r = 0; r < m; ++r)
            {
              double b = 0;
              for (int k = 0; k < m; ++k)
                {
                  const double e = f == 0 ? c[0 + m * k] : f == 1 ? c[q + m * k] : f == 2 ? c[k + m * 0] : c[k + m * q];
                  b += to_bernstein[r * m + k] * e;
                }
              mn = std::min(mn, b);
              mx = std::max(mx, b);
            }
          n_in += mx < 0;
          n_out += mn > 0;
        }
      return n_in == 4 ? INSIDE : (n_out == 4 ? OUTSIDE : INTERSECTED);
    }

    // height-function quadrature (NonMatching::QuadratureGenerator) for the polynomial level set of a cell whose zero set
    // is a graph over one coordinate direction: the base interval is split at the roots of the level set on the two faces
    // across the height direction, QGauss<1>(p+1) on every piece and on the inside part of every column
    void Assembly::rules_q(const int *idx, std::vector<Pt> &inside, std::vector<Pt> &surf) const
    {
      double c[(MAX_DEGREE + 1) * (MAX_DEGREE + 1)], g[2];
      local_values_q(idx, c);
      eval_q(c, 0.5, 0.5, g);
      const int k = std::fabs(g[0]) * (1 + 1e-8) >= std::fabs(g[1]) ? 0 : 1; // height direction; x on (near-)ties
      auto at = [&](double s, double t, double *grad) { return k == 0 ? eval_q(c, t, s, grad) : eval_q(c, s, t, grad); };
      // monotone in the height direction?
      {
        double ref[2];
        at(0.5, 0.5, ref);
        for (int i = 0; i <= 8; ++i)
          for (int j = 0; j <= 8; ++j)
            {
              at(i / 8.0, j / 8.0, g);
              GDM_REQUIRE(g[k] * ref[k] > 0, GDM_ERR_NOT_IMPLEMENTED,
                          "level set of degree > 1: the zero set is not a graph over a coordinate direction in a cell");
            }
      }
      auto bisect = [&](auto &&fn, double a, double b) {
        double fa = fn(a);
        for (int it = 0; it < 200 && b - a > 1e-16; ++it)
          {
            const double m = 0.5 * (a + b), fm = fn(m);
            if ((fa < 0) == (fm < 0))
              {
                a  = m;
                fa = fm;
              }
            else
              b = m;
          }
        return 0.5 * (a + b);
      };
      std::vector<double> breaks = {0.0, 1.0};
      for (int face_t = 0; face_t < 2; ++face_t)
        {
          auto        fn   = [&](double s) { return at(s, (double)face_t, nullptr); };
          const int   n_s  = 64; // sign changes on a fine sampling: simple roots of a resolved geometry
          double      prev = fn(0.0);
          for (int i = 1; i <= n_s; ++i)
            {
              const double s1 = (double)i / n_s, cur = fn(s1);
              if ((prev < 0) != (cur < 0))
                breaks.push_back(bisect(fn, (double)(i - 1) / n_s, s1));
              prev = cur;
            }
        }
      std::sort(breaks.begin(), breaks.end());
      const size_t ng = gauss.x.size();
      for (size_t b = 0; b + 1 < breaks.size(); ++b)
        {
          const double s0 = breaks[b], s1 = breaks[b + 1];
          if (s1 - s0 <= 1e-14)
            continue;
          for (size_t qs = 0; qs < ng; ++qs)
            {
              const double sq = s0 + (s1 - s0) * gauss.x[qs], ws = (s1 - s0) * gauss.w[qs];
              const double a = at(sq, 0.0, nullptr), e = at(sq, 1.0, nullptr);
              double       t0, t1;
              if (a * e < 0)
                {
                  const double r = bisect([&](double t) { return at(sq, t, nullptr); }, 0.0, 1.0);
                  t0             = a < 0 ? 0.0 : r;
                  t1             = a < 0 ? r : 1.0;
                  Pt p{};
                  p.x[k]     = r;
                  p.x[1 - k] = sq;
                  eval_q(c, p.x[0], p.x[1], g);
                  const double gn = std::sqrt(g[0] * g[0] + g[1] * g[1]);
                  p.w    = ws * gn / std::fabs(g[k]);
                  p.n[0] = g[0] / gn;
                  p.n[1] = g[1] / gn;
                  surf.push_back(p);
                }
              else if (a < 0)
                {
                  t0 = 0.0;
                  t1 = 1.0;
                }
              else
                continue;
              for (size_t qt = 0; qt < ng; ++qt)
                {
                  Pt p{};
                  p.x[k]     = t0 + (t1 - t0) * gauss.x[qt];
                  p.x[1 - k] = sq;
                  p.w        = ws * (t1 - t0) * gauss.w[qt];
                  inside.push_back(p);
                }
            }
        }
    }

    void Assembly::boundary_face_rule(const int *idx, int d, int side, std::vector<Pt> &pts) const
    {
      pts.clear();
      double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      vertex_values(idx, v);
      MLF f;
      f.d = dim;
      for (int i = 0; i < (1 << dim); ++i)
        f.c[i] = v[i];
      const MLF        fc = face(f, d, side);
      std::vector<Pt>  base;
      const double     l0[3] = {0, 0, 0}, h1[3] = {1, 1, 1};
      if (dim == 1)
        {
          if (fc.c[0] < 0)
            {
              Pt b{};
              b.w = 1.0;
              base.push_back(b);
            }
        }
      else
        volume({fc}, {-1}, l0, h1, dim - 1, gauss, 0, base);
      for (const Pt &b : base)
        {
          Pt  p{};
          int m = 0;
          for (int e = 0; e < dim; ++e)
            p.x[e] = (e == d) ? (double)side : b.x[m++];
          p.w = b.w;
          pts.push_back(p);
        }
    }

    void Assembly::boundary_terms(const int *idx, CutScratch &sc, double *local, gdm_function_fn g, void *user, double *lrhs) const
    {
      const double nitsche = desc.nitsche_parameter / h_min();
      for (int d = 0; d < dim; ++d)
        for (int side = 0; side < 2; ++side)
          {
            if (idx[d] != (side == 0 ? 0 : N[d] - 1))
              continue;
            boundary_face_rule(idx, d, side, sc.spts);
            if (sc.spts.empty())
              continue;
            shape_at_points(idx, sc.spts, sc.value, sc.grads);
            double area = 1;
            for (int e = 0; e < dim; ++e)
              if (e != d)
                area *= h[e];
            const double sgn = side ? 1.0 : -1.0;
            for (size_t q = 0; q < sc.spts.size(); ++q)
              {
                const double  jxw = sc.spts[q].w * area;
                const double *vq = sc.value.data() + q * npc, *gq = sc.grads[d].data() + q * npc;
                double        gval = 0;
                if (lrhs && g)
                  {
                    double x[3] = {0, 0, 0};
                    for (int e = 0; e < dim; ++e)
                      x[e] = lo[e] + (idx[e] + sc.spts[q].x[e]) * h[e];
                    gval = g(x, 0, user);
                  }
                for (int i = 0; i < npc; ++i)
                  {
                    const double ngi = sgn * gq[i];
                    if (lrhs && g)
                      lrhs[i] += gval * (nitsche * vq[i] - ngi) * jxw;
                    if (local)
                      for (int j = 0; j < npc; ++j)
                        local[(size_t)i * npc + j] += (-ngi * vq[j] - sgn * gq[j] * vq[i] + nitsche * vq[i] * vq[j]) * jxw;
                  }
              }
          }
    }

    // which = 0: P_ij = sum_q (n . grad phi_i) phi_j JxW, 1: P^T, 2: Q_ij = sum_q phi_i phi_j JxW over the cut surface
    // (n = normal of the level set): the interface coupling of the two-domain residual, wave/stiffness.h:441-574
    void Assembly::coupling(int which, std::vector<uint64_t> &rows, std::vector<uint64_t> &rp, std::vector<uint64_t> &cols,
                            std::vector<double> &vals) const
    {
      struct Trip
      {
        uint64_t r, c;
        double   v;
      };
      std::vector<Trip>     trips;
      CutScratch            sc;
      std::vector<uint64_t> dofs;
      std::vector<double>   ng(npc);
      int                   idx[3], off[3];
      const double          vol = cell_volume();
      for (uint64_t cell = 0; cell < n_cells; ++cell)
        {
          if (location[cell] != INTERSECTED)
            continue;
          cell_index(cell, idx);
          sc.ipts.clear();
          sc.spts.clear();
          cell_rules(idx, sc.ipts, sc.spts);
          if (sc.spts.empty())
            continue;
          cell_dofs(idx, off, dofs);
          shape_at_points(idx, sc.spts, sc.value, sc.grads);
          std::vector<double> local((size_t)npc * npc, 0.0);
          for (size_t q = 0; q < sc.spts.size(); ++q)
            {
              double nph[3] = {0, 0, 0}, scale = 0;
              for (int e = 0; e < dim; ++e)
                {
                  nph[e] = sc.spts[q].n[e] / h[e];
                  scale += nph[e] * nph[e];
                }
              scale            = std::sqrt(scale);
              const double jxw = sc.spts[q].w * vol * scale;
              for (int i = 0; i < npc; ++i)
                {
                  double t = 0;
                  for (int e = 0; e < dim; ++e)
                    t += nph[e] / scale * sc.grads[e][q * npc + i];
                  ng[i] = t;
                }
              const double *vq = sc.value.data() + q * npc;
              for (int i = 0; i < npc; ++i)
                for (int j = 0; j < npc; ++j)
                  local[(size_t)i * npc + j] += (which == 0 ? ng[i] * vq[j] : which == 1 ? vq[i] * ng[j] : vq[i] * vq[j]) * jxw;
            }
          for (int i = 0; i < npc; ++i)
            for (int j = 0; j < npc; ++j)
              trips.push_back({dofs[i], dofs[j], local[(size_t)i * npc + j]});
        }
      std::stable_sort(trips.begin(), trips.end(), [](const Trip &a, const Trip &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
      rows.clear();
      cols.clear();
      vals.clear();
      rp.assign(1, 0);
      for (size_t k = 0; k < trips.size();)
        {
          const uint64_t r = trips[k].r;
          rows.push_back(r);
          while (k < trips.size() && trips[k].r == r)
            {
              const uint64_t c   = trips[k].c;
              double         sum = 0;
              for (; k < trips.size() && trips[k].r == r && trips[k].c == c; ++k)
                sum += trips[k].v;
              cols.push_back(c);
              vals.push_back(sum);
            }
          rp.push_back(cols.size());
        }
    }

    void Assembly::build()
    {
      const int B = 2 * p + 3, Bc = p + 1; // column offsets -Bc..Bc per direction
      uint64_t  box = 1;
      for (int e = 0; e < dim; ++e)
        box *= B;
      location.assign(n_cells, 0);
      std::vector<uint8_t> touched(n_dofs, 0), irregular(n_dofs, 0);
      std::vector<uint64_t> dofs, ndofs;
      int                   idx[3], off[3], noff[3];
      for (uint64_t cell = 0; cell < n_cells; ++cell)
        {
          cell_index(cell, idx);
          location[cell] = classify_cell(idx);
          ++counts[location[cell]];
        }
      for (uint64_t cell = 0; cell < n_cells; ++cell)
        {
          cell_index(cell, idx);
          cell_dofs(idx, off, dofs);
          if (location[cell] != OUTSIDE)
            for (uint64_t i : dofs)
              touched[i] = 1;
          if (location[cell] != INSIDE)
            for (uint64_t i : dofs)
              irregular[i] = 1;
          if (desc.ghost_penalty && location[cell] != OUTSIDE)
            for (int d = 0; d < dim; ++d)
              for (int side = 0; side < 2; ++side)
                if (ghost_penalty_neighbor(idx, d, side) >= 0)
                  for (uint64_t i : dofs) // the neighbour marks its own window when the loop gets to it
                    irregular[i] = 1;
          if (desc.domain_boundary_terms && desc.kind == 0 && location[cell] != OUTSIDE && at_box_boundary(idx))
            for (uint64_t i : dofs)
              irregular[i] = 1;
        }
      std::vector<int64_t> slot(n_dofs, -1);
      n_band_rows = 0;
      // rows of this rank only (slab partition of the DoFs, system.h:720-757): every rank runs the generator on its range
      const uint64_t row_b = desc.row_end > desc.row_begin ? desc.row_begin : 0;
      const uint64_t row_e = desc.row_end > desc.row_begin ? std::min<uint64_t>(desc.row_end, n_dofs) : n_dofs;
      for (uint64_t i = row_b; i < row_e; ++i)
        if (touched[i] && irregular[i])
          slot[i] = (int64_t)n_band_rows++;
      std::vector<double> acc(n_band_rows * box, 0.0);
      rhs.assign(n_dofs, 0.0);

      // local index -> node offsets inside the window
      std::vector<int> lidx(3 * npc, 0);
      {
        int m = 0;
        for (int k = 0; k < (dim >= 3 ? p + 1 : 1); ++k)
          for (int j = 0; j < (dim >= 2 ? p + 1 : 1); ++j)
            for (int i = 0; i <= p; ++i, ++m)
              {
                lidx[3 * m]     = i;
                lidx[3 * m + 1] = j;
                lidx[3 * m + 2] = k;
              }
      }
      // box offset of (row i, column j) = base(window offsets) - lpart[i] + lpart[j]
      std::vector<int64_t> lpart(npc);
      for (int m = 0; m < npc; ++m)
        {
          int64_t o = 0, st = 1;
          for (int e = 0; e < dim; ++e)
            {
              o += (int64_t)lidx[3 * m + e] * st;
              st *= B;
            }
          lpart[m] = o;
        }
      auto scatter = [&](const std::vector<uint64_t> &rdofs, const int *roff, const std::vector<uint64_t> &cdofs,
                         const int *coff, const double *mat, int ld, int r0, int c0) {
        (void)cdofs;
        int64_t base = 0, st = 1;
        for (int e = 0; e < dim; ++e)
          {
            base += (int64_t)(coff[e] - roff[e] + Bc) * st;
            st *= B;
          }
        for (int i = 0; i < npc; ++i)
          {
            const int64_t sl = slot[rdofs[i]];
            if (sl < 0)
              continue;
            double       *row = &acc[(uint64_t)sl * box] + (base - lpart[i]);
            const double *m   = mat + (size_t)(r0 + i) * ld + c0;
            for (int j = 0; j < npc; ++j)
              row[lpart[j]] += m[j];
          }
      };

      // full-cell tables per category
      const bool   mass = desc.kind == 1;
      const double vol = cell_volume(), hm = h_min();
      const double nitsche = desc.nitsche_parameter / hm;
      std::vector<Pt> full;
      {
        const double l0[3] = {0, 0, 0}, h1[3] = {1, 1, 1};
        tensor_gauss(l0, h1, dim, gauss, full);
      }
      std::map<int, std::pair<std::vector<double>, std::vector<double>>> inside_cache; // category -> (K, f)
      std::vector<double> value, grads[3], local((size_t)npc * npc), lrhs(npc);
      auto inside_tables = [&](const int *cidx) -> const std::pair<std::vector<double>, std::vector<double>> & {
        const int cat = category(cidx);
        auto      it  = inside_cache.find(cat);
        if (it != inside_cache.end())
          return it->second;
        shape_at_points(cidx, full, value, grads);
        std::vector<double> K((size_t)npc * npc, 0.0), f(npc, 0.0);
        for (size_t q = 0; q < full.size(); ++q)
          {
            const double jxw = full[q].w * vol;
            for (int i = 0; i < npc; ++i)
              {
                f[i] += desc.rhs_value * value[q * npc + i] * jxw;
                for (int j = 0; j < npc; ++j)
                  {
                    double s = 0;
                    if (mass)
                      s = value[q * npc + i] * value[q * npc + j];
                    else
                      for (int e = 0; e < dim; ++e)
                        s += grads[e][q * npc + i] * grads[e][q * npc + j];
                    K[(size_t)i * npc + j] += s * jxw;
                  }
              }
          }
        return inside_cache.emplace(cat, std::make_pair(std::move(K), std::move(f))).first->second;
      };

      // ghost-penalty face matrices by (direction, side, category here, category there)
      std::map<std::array<int, 4>, std::vector<double>> gp_cache;
      std::vector<Pt>                                   face_pts;
      {
        const double l0[3] = {0, 0, 0}, h1[3] = {1, 1, 1};
        tensor_gauss(l0, h1, dim - 1, gauss, face_pts);
      }
      std::vector<double> gh[3], gt[3], vtmp;
      auto gp_matrix = [&](const int *cidx, const int *nidx, int d, int side) -> const std::vector<double> & {
        const std::array<int, 4> key = {d, side, category(cidx), category(nidx)};
        auto                     it  = gp_cache.find(key);
        if (it != gp_cache.end())
          return it->second;
        std::vector<Pt> here(face_pts.size()), there(face_pts.size());
        double          area = 1;
        for (int e = 0; e < dim; ++e)
          if (e != d)
            area *= h[e];
        for (size_t q = 0; q < face_pts.size(); ++q)
          {
            int m = 0;
            for (int e = 0; e < dim; ++e)
              {
                const double x = (e == d) ? 0.0 : face_pts[q].x[m++];
                here[q].x[e]   = (e == d) ? (double)side : x;
                there[q].x[e]  = (e == d) ? (double)(1 - side) : x;
              }
          }
        shape_at_points(cidx, here, vtmp, gh);
        shape_at_points(nidx, there, vtmp, gt);
        const int           n2 = 2 * npc;
        std::vector<double> S((size_t)n2 * n2, 0.0), jump(n2);
        const double        coef = 0.5 * desc.ghost_parameter * std::pow(hm, desc.gp_h_power);
        for (size_t q = 0; q < face_pts.size(); ++q)
          {
            for (int i = 0; i < npc; ++i)
              {
                jump[i]       = gh[d][q * npc + i];
                jump[npc + i] = -gt[d][q * npc + i];
              }
            const double jxw = coef * face_pts[q].w * area;
            for (int i = 0; i < n2; ++i)
              for (int j = 0; j < n2; ++j)
                S[(size_t)i * n2 + j] += jxw * jump[i] * jump[j];
          }
        return gp_cache.emplace(key, std::move(S)).first->second;
      };

      // Blocks of consecutive cells: the intersected cells of a block are evaluated by all threads, then the block is
      // scattered in cell order (the sums do not depend on the number of threads).
      double     t_cut = 0, t_scatter = 0;
      const auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
      unsigned n_threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      if (const char *env = std::getenv("GDM_CUT_THREADS"))
        n_threads = (unsigned)std::max(1, std::atoi(env));
      const size_t          block_cut_cells = 64 * (size_t)n_threads;
      std::vector<double>   block_local(block_cut_cells * npc * npc), block_rhs(block_cut_cells * npc);
      std::vector<uint64_t> block_cells;
      std::vector<int32_t>  block_slot;
      std::vector<CutScratch> scratch(n_threads);
      auto cell_in_range = [&](const int *cidx) {
        if (row_b == 0 && row_e == n_dofs)
          return true;
        int                   o[3];
        std::vector<uint64_t> d;
        cell_dofs(cidx, o, d);
        for (uint64_t i : d)
          if (i >= row_b && i < row_e)
            return true;
        return false;
      };
      for (uint64_t b0 = 0; b0 < n_cells;)
        {
          uint64_t b1 = b0;
          block_cells.clear();
          block_slot.clear();
          for (; b1 < n_cells && block_cells.size() < block_cut_cells; ++b1)
            {
              block_slot.push_back(-1);
              if (location[b1] != INTERSECTED)
                continue;
              block_slot.back() = (int32_t)block_cells.size();
              block_cells.push_back(b1);
            }
          const double t0 = now();
          auto worker = [&](unsigned tid) {
            int cidx[3];
            for (size_t k = tid; k < block_cells.size(); k += n_threads)
              {
                cell_index(block_cells[k], cidx);
                if (!cell_in_range(cidx))
                  {
                    std::fill(&block_rhs[k * npc], &block_rhs[(k + 1) * npc], 0.0);
                    continue;
                  }
                cut_cell_matrix(cidx, scratch[tid], &block_local[k * npc * npc], &block_rhs[k * npc]);
              }
          };
          if (n_threads > 1 && block_cells.size() > 1)
            {
              std::vector<std::thread> pool;
              for (unsigned tid = 0; tid < n_threads; ++tid)
                pool.emplace_back(worker, tid);
              for (auto &th : pool)
                th.join();
            }
          else
            worker(0);
          const double t1 = now();
          t_cut += t1 - t0;
          for (uint64_t cell = b0; cell < b1; ++cell)
            {
              if (location[cell] == OUTSIDE)
                continue;
              cell_index(cell, idx);
              cell_dofs(idx, off, dofs);
              if (location[cell] == INSIDE)
                {
                  const auto &t = inside_tables(idx);
                  for (int i = 0; i < npc; ++i)
                    rhs[dofs[i]] += t.second[i];
                  bool any = false;
                  for (int i = 0; i < npc && !any; ++i)
                    any = slot[dofs[i]] >= 0;
                  if (any)
                    scatter(dofs, off, dofs, off, t.first.data(), npc, 0, 0);
                }
              else
                {
                  const double *local = &block_local[(size_t)block_slot[cell - b0] * npc * npc];
                  const double *lrhs  = &block_rhs[(size_t)block_slot[cell - b0] * npc];
                  for (int i = 0; i < npc; ++i)
                    rhs[dofs[i]] += lrhs[i];
                  if (cell_in_range(idx))
                    scatter(dofs, off, dofs, off, local, npc, 0, 0);
                }
              if (desc.domain_boundary_terms && !mass && at_box_boundary(idx) && cell_in_range(idx))
                {
                  std::fill(local.begin(), local.end(), 0.0);
                  boundary_terms(idx, scratch[0], local.data(), nullptr, nullptr, nullptr);
                  scatter(dofs, off, dofs, off, local.data(), npc, 0, 0);
                }
              if (desc.ghost_penalty)
                for (int d = 0; d < dim; ++d)
                  for (int side = 0; side < 2; ++side)
                    {
                      const int64_t nc = ghost_penalty_neighbor(idx, d, side);
                      if (nc < 0)
                        continue;
                      int nidx[3];
                      cell_index((uint64_t)nc, nidx);
                      cell_dofs(nidx, noff, ndofs);
                      const std::vector<double> &S = gp_matrix(idx, nidx, d, side);
                      const int                  n2 = 2 * npc;
                      scatter(dofs, off, dofs, off, S.data(), n2, 0, 0);
                      scatter(dofs, off, ndofs, noff, S.data(), n2, 0, npc);
                      scatter(ndofs, noff, dofs, off, S.data(), n2, npc, 0);
                      scatter(ndofs, noff, ndofs, noff, S.data(), n2, npc, npc);
                    }
            }
          b0 = b1;
          t_scatter += now() - t1;
        }
      if (std::getenv("GDM_CUT_VERBOSE"))
        std::fprintf(stderr, "gdm_cut: %u threads, cut cells %.3f s, scatter %.3f s\n", n_threads, t_cut, t_scatter);

      // rows in ascending DoF order; columns ascending inside a row
      row_ids.clear();
      rowptr.assign(1, 0);
      col.clear();
      val.clear();
      n_identity_rows = 0;
      for (uint64_t i = row_b; i < row_e; ++i)
        {
          if (!touched[i])
            {
              row_ids.push_back(i);
              col.push_back(i);
              val.push_back(desc.outside_diagonal);
              rowptr.push_back(col.size());
              ++n_identity_rows;
              continue;
            }
          if (slot[i] < 0)
            continue;
          int node[3] = {0, 0, 0};
          node[0]     = (int)(i % nn[0]);
          if (dim >= 2)
            node[1] = (int)((i / nn[0]) % nn[1]);
          if (dim >= 3)
            node[2] = (int)(i / ((uint64_t)nn[0] * nn[1]));
          const double *row = &acc[(uint64_t)slot[i] * box];
          row_ids.push_back(i);
          for (uint64_t o = 0; o < box; ++o)
            {
              uint64_t r = o;
              int      cn[3] = {0, 0, 0};
              bool     in = true, diag = true;
              for (int e = 0; e < dim; ++e)
                {
                  const int delta = (int)(r % B) - Bc;
                  r /= B;
                  cn[e] = node[e] + delta;
                  in    = in && cn[e] >= 0 && cn[e] < nn[e];
                  diag  = diag && delta == 0;
                }
              if (!in)
                continue;
              double v = row[o];
              if (diag && v == 0.0)
                v = 1.0; // prototypes/cut_poisson_01_gdm.cc:324-329
              if (v != 0.0 || diag)
                {
                  col.push_back(node_of(cn));
                  val.push_back(v);
                }
            }
          rowptr.push_back(col.size());
        }
    }
    // out_i = sum_q f(x_q) phi_i JxW over the inside part + sum_q g(x_q) (gamma_D / h phi_i - d_n phi_i) JxW on the surface:
    // the data-dependent part of the residual applications/wave/include/gdm/wave/stiffness.h:186-260
    void Assembly::load_vector(gdm_function_fn f, void *f_user, gdm_function_fn g, void *g_user, double *out)
    {
      std::fill(out, out + n_dofs, 0.0);
      const double          vol = cell_volume(), nitsche = desc.nitsche_parameter / h_min();
      std::vector<double>   value, grads[3];
      std::vector<uint64_t> dofs;
      int                   idx[3], off[3];
      if (!cut_loads_built)
        {
          std::vector<Pt> ipts, spts;
          for (uint64_t cell = 0; cell < n_cells; ++cell)
            if (location[cell] == INTERSECTED)
              {
                cell_index(cell, idx);
                ipts.clear();
                spts.clear();
                cell_rules(idx, ipts, spts);
                CutCellLoad L;
                L.cell = cell;
                auto physical = [&](const std::vector<Pt> &pts, std::vector<double> &x) {
                  x.resize(pts.size() * dim);
                  for (size_t q = 0; q < pts.size(); ++q)
                    for (int e = 0; e < dim; ++e)
                      x[q * dim + e] = lo[e] + (idx[e] + pts[q].x[e]) * h[e];
                };
                if (!ipts.empty())
                  {
                    shape_at_points(idx, ipts, value, grads);
                    physical(ipts, L.vpts);
                    L.vW.resize(ipts.size() * npc);
                    for (size_t q = 0; q < ipts.size(); ++q)
                      for (int i = 0; i < npc; ++i)
                        L.vW[q * npc + i] = value[q * npc + i] * (ipts[q].w * vol);
                  }
                if (!spts.empty())
                  {
                    shape_at_points(idx, spts, value, grads);
                    physical(spts, L.spts);
                    L.sW.resize(spts.size() * npc);
                    for (size_t q = 0; q < spts.size(); ++q)
                      {
                        double nph[3] = {0, 0, 0}, scale = 0;
                        for (int e = 0; e < dim; ++e)
                          {
                            nph[e] = spts[q].n[e] / h[e];
                            scale += nph[e] * nph[e];
                          }
                        scale = std::sqrt(scale);
                        const double jxw = spts[q].w * vol * scale;
                        for (int i = 0; i < npc; ++i)
                          {
                            double ng = 0;
                            for (int e = 0; e < dim; ++e)
                              ng += nph[e] / scale * grads[e][q * npc + i];
                            L.sW[q * npc + i] = (nitsche * value[q * npc + i] - ng) * jxw;
                          }
                      }
                  }
                cut_loads.push_back(std::move(L));
              }
          cut_loads_built = true;
        }
      if (f)
        {
          std::vector<Pt> full;
          const double    l0[3] = {0, 0, 0}, h1[3] = {1, 1, 1};
          tensor_gauss(l0, h1, dim, gauss, full);
          std::map<int, std::vector<double>> tables; // category -> phi_i(x_q)
          for (uint64_t cell = 0; cell < n_cells; ++cell)
            if (location[cell] == INSIDE)
              {
                cell_index(cell, idx);
                cell_dofs(idx, off, dofs);
                const int cat = category(idx);
                auto      it  = tables.find(cat);
                if (it == tables.end())
                  {
                    shape_at_points(idx, full, value, grads);
                    it = tables.emplace(cat, value).first;
                  }
                const std::vector<double> &phi = it->second;
                for (size_t q = 0; q < full.size(); ++q)
                  {
                    double x[3] = {0, 0, 0};
                    for (int e = 0; e < dim; ++e)
                      x[e] = lo[e] + (idx[e] + full[q].x[e]) * h[e];
                    const double fq = f(x, 0, f_user) * full[q].w * vol;
                    for (int i = 0; i < npc; ++i)
                      out[dofs[i]] += fq * phi[q * npc + i];
                  }
              }
        }
      for (const CutCellLoad &L : cut_loads)
        {
          cell_index(L.cell, idx);
          cell_dofs(idx, off, dofs);
          for (int pass = 0; pass < 2; ++pass)
            {
              gdm_function_fn            fn   = pass == 0 ? f : g;
              void                      *user = pass == 0 ? f_user : g_user;
              const std::vector<double> &pts = pass == 0 ? L.vpts : L.spts, &W = pass == 0 ? L.vW : L.sW;
              if (!fn)
                continue;
              for (size_t q = 0; q < pts.size() / dim; ++q)
                {
                  double x[3] = {0, 0, 0};
                  for (int e = 0; e < dim; ++e)
                    x[e] = pts[q * dim + e];
                  const double fq = fn(x, 0, user);
                  for (int i = 0; i < npc; ++i)
                    out[dofs[i]] += fq * W[q * npc + i];
                }
            }
        }
    }
  } // namespace cut
} // namespace gdm

using namespace gdm;

struct gdm_cut_s
{
  cut::Assembly a;
};

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
#define GDM_ARG_NOTHROW(x)                          \
  if ((x) == nullptr)                              \
    {                                              \
      gdm::set_last_error("null argument " #x);    \
      return GDM_ERR_INVALID;                      \
    }

extern "C" {

int gdm_cut_quadrature(int dim, const double *vertex_values, int n_gauss, uint64_t capacity, uint64_t *n_inside,
                       double *inside_points, double *inside_weights, uint64_t *n_surface, double *surface_points,
                       double *surface_weights, double *surface_normals)
{
  GDM_TRY
  GDM_ARG(vertex_values);
  GDM_ARG(n_inside);
  GDM_ARG(n_surface);
  GDM_REQUIRE(dim >= 1 && dim <= 3, GDM_ERR_INVALID, "dim must be 1, 2 or 3");
  GDM_REQUIRE(n_gauss >= 1 && n_gauss <= 32, GDM_ERR_INVALID, "n_gauss out of range");
  const cut::Gauss     g = cut::make_gauss(n_gauss);
  std::vector<cut::Pt> in, sf;
  cut::cut_quadrature(dim, vertex_values, g, in, sf);
  *n_inside  = in.size();
  *n_surface = sf.size();
  for (size_t q = 0; q < std::min<size_t>(in.size(), capacity); ++q)
    {
      if (inside_weights)
        inside_weights[q] = in[q].w;
      if (inside_points)
        for (int e = 0; e < dim; ++e)
          inside_points[q * dim + e] = in[q].x[e];
    }
  for (size_t q = 0; q < std::min<size_t>(sf.size(), capacity); ++q)
    {
      if (surface_weights)
        surface_weights[q] = sf[q].w;
      for (int e = 0; e < dim; ++e)
        {
          if (surface_points)
            surface_points[q * dim + e] = sf[q].x[e];
          if (surface_normals)
            surface_normals[q * dim + e] = sf[q].n[e];
        }
    }
  GDM_CATCH
}

int gdm_cut_level_set_points(const gdm_cut_desc *desc, uint64_t *n_points, double *points)
{
  GDM_TRY
  GDM_ARG(desc);
  GDM_ARG(n_points);
  GDM_REQUIRE(desc->dim >= 1 && desc->dim <= 3, GDM_ERR_INVALID, "dim must be 1, 2 or 3");
  const int     q = desc->level_set_degree > 1 ? desc->level_set_degree : 1;
  GDM_REQUIRE(q <= MAX_DEGREE, GDM_ERR_INVALID, "level_set_degree too large");
  cut::Assembly a;
  a.q = q;
  a.setup_q();
  uint64_t n[3] = {1, 1, 1}, total = 1;
  for (int e = 0; e < desc->dim; ++e)
    {
      GDM_REQUIRE(desc->n_subdivisions[e] >= 1, GDM_ERR_INVALID, "n_subdivisions");
      n[e] = (uint64_t)q * desc->n_subdivisions[e] + 1;
      total *= n[e];
    }
  *n_points = total;
  if (points)
    for (uint64_t i = 0; i < total; ++i)
      {
        uint64_t r = i;
        for (int e = 0; e < desc->dim; ++e)
          {
            const uint64_t k = r % n[e];
            r /= n[e];
            const double h = (desc->hi[e] - desc->lo[e]) / desc->n_subdivisions[e];
            points[i * desc->dim + e] =
              k == n[e] - 1 ? desc->hi[e] : desc->lo[e] + ((double)(k / q) + a.gll[k % q]) * h;
          }
      }
  GDM_CATCH
}

int gdm_cut_poisson_create(const gdm_cut_desc *desc, const double *level_set, gdm_cut_t *out)
{
  GDM_TRY
  GDM_ARG(desc);
  GDM_ARG(level_set);
  GDM_ARG(out);
  GDM_REQUIRE(desc->dim >= 1 && desc->dim <= 3, GDM_ERR_INVALID, "dim must be 1, 2 or 3");
  GDM_REQUIRE(desc->fe_degree >= 1 && desc->fe_degree <= MAX_DEGREE && desc->fe_degree % 2 == 1, GDM_ERR_INVALID,
              "fe_degree must be odd and <= 9");
  GDM_REQUIRE(desc->kind == 0 || desc->kind == 1, GDM_ERR_INVALID, "kind must be 0 (stiffness + Nitsche) or 1 (mass)");
  auto           h = std::make_unique<gdm_cut_s>();
  cut::Assembly &a = h->a;
  a.desc           = *desc;
  a.dim            = desc->dim;
  a.p              = desc->fe_degree;
  a.npc            = 1;
  a.n_dofs = a.n_cells = 1;
  for (int e = 0; e < 3; ++e)
    {
      a.N[e] = e < a.dim ? (int)desc->n_subdivisions[e] : 1;
      a.nn[e] = e < a.dim ? a.N[e] + 1 : 1;
      a.lo[e] = desc->lo[e];
      a.h[e]  = e < a.dim ? (desc->hi[e] - desc->lo[e]) / a.N[e] : 1.0;
      if (e < a.dim)
        {
          GDM_REQUIRE(a.N[e] >= a.p, GDM_ERR_INVALID, "n_subdivisions must be >= fe_degree");
          GDM_REQUIRE(desc->hi[e] > desc->lo[e], GDM_ERR_INVALID, "empty domain");
          a.npc *= a.p + 1;
          a.n_dofs *= (uint64_t)a.nn[e];
          a.n_cells *= (uint64_t)a.N[e];
        }
    }
  a.gauss = cut::make_gauss(a.p + 1);
  a.q     = desc->level_set_degree > 1 ? desc->level_set_degree : 1;
  uint64_t n_ls = a.n_dofs;
  if (a.q > 1)
    {
      GDM_REQUIRE(a.dim == 2, GDM_ERR_NOT_IMPLEMENTED, "level set of degree > 1: two dimensions only");
      GDM_REQUIRE(a.q <= MAX_DEGREE, GDM_ERR_INVALID, "level_set_degree too large");
      GDM_REQUIRE(!desc->domain_boundary_terms, GDM_ERR_NOT_IMPLEMENTED, "level set of degree > 1 with domain_boundary_terms");
      n_ls = ((uint64_t)a.q * a.N[0] + 1) * ((uint64_t)a.q * a.N[1] + 1);
      a.setup_q();
    }
  a.ls.assign(level_set, level_set + n_ls);
  a.build();
  *out = h.release();
  GDM_CATCH
}

int gdm_cut_destroy(gdm_cut_t c)
{
  delete c;
  return GDM_OK;
}

int gdm_cut_sizes(gdm_cut_t c, uint64_t *n_rows, uint64_t *nnz, uint64_t *n_identity_rows, uint64_t *n_cells_by_location)
{
  GDM_TRY
  GDM_ARG(c);
  if (n_rows)
    *n_rows = c->a.row_ids.size();
  if (nnz)
    *nnz = c->a.col.size();
  if (n_identity_rows)
    *n_identity_rows = c->a.n_identity_rows;
  if (n_cells_by_location)
    for (int i = 0; i < 3; ++i)
      n_cells_by_location[i] = c->a.counts[i];
  GDM_CATCH
}

int gdm_cut_rows(gdm_cut_t c, uint64_t *row_ids, uint64_t *rowptr, uint64_t *col, double *val)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(row_ids);
  GDM_ARG(rowptr);
  GDM_ARG(col);
  GDM_ARG(val);
  const cut::Assembly &a = c->a;
  std::copy(a.row_ids.begin(), a.row_ids.end(), row_ids);
  std::copy(a.rowptr.begin(), a.rowptr.end(), rowptr);
  std::copy(a.col.begin(), a.col.end(), col);
  std::copy(a.val.begin(), a.val.end(), val);
  GDM_CATCH
}

int gdm_cut_rhs(gdm_cut_t c, double *rhs)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(rhs);
  std::copy(c->a.rhs.begin(), c->a.rhs.end(), rhs);
  GDM_CATCH
}

int gdm_cut_load_vector(gdm_cut_t c, gdm_function_fn f, void *f_user, gdm_function_fn g, void *g_user, double *out)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(out);
  c->a.load_vector(f, f_user, g, g_user, out);
  GDM_CATCH
}

int gdm_cut_boundary_load_vector(gdm_cut_t c, gdm_function_fn g, void *user, double *out)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(g);
  GDM_ARG(out);
  const cut::Assembly &a = c->a;
  std::fill(out, out + a.n_dofs, 0.0);
  cut::Assembly::CutScratch sc;
  std::vector<uint64_t>     dofs;
  std::vector<double>       lrhs(a.npc);
  int                       idx[3], off[3];
  for (uint64_t cell = 0; cell < a.n_cells; ++cell)
    {
      if (a.location[cell] == cut::OUTSIDE)
        continue;
      a.cell_index(cell, idx);
      if (!a.at_box_boundary(idx))
        continue;
      std::fill(lrhs.begin(), lrhs.end(), 0.0);
      a.boundary_terms(idx, sc, nullptr, g, user, lrhs.data());
      a.cell_dofs(idx, off, dofs);
      for (int i = 0; i < a.npc; ++i)
        out[dofs[i]] += lrhs[i];
    }
  GDM_CATCH
}

int gdm_cut_coupling_rows(gdm_cut_t c, int which, uint64_t capacity_rows, uint64_t capacity_nnz, uint64_t *n_rows,
                          uint64_t *nnz, uint64_t *row_ids, uint64_t *rowptr, uint64_t *col, double *val)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(n_rows);
  GDM_ARG(nnz);
  GDM_REQUIRE(which >= 0 && which <= 2, GDM_ERR_INVALID, "which must be 0 (P), 1 (P^T) or 2 (Q)");
  std::vector<uint64_t> rows, rp, cols;
  std::vector<double>   vals;
  c->a.coupling(which, rows, rp, cols, vals);
  *n_rows = rows.size();
  *nnz    = cols.size();
  if (row_ids && rowptr && col && val && capacity_rows >= rows.size() && capacity_nnz >= cols.size())
    {
      std::copy(rows.begin(), rows.end(), row_ids);
      std::copy(rp.begin(), rp.end(), rowptr);
      std::copy(cols.begin(), cols.end(), col);
      std::copy(vals.begin(), vals.end(), val);
    }
  GDM_CATCH
}

int gdm_cut_locations(gdm_cut_t c, uint8_t *location)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(location);
  std::copy(c->a.location.begin(), c->a.location.end(), location);
  GDM_CATCH
}

int gdm_cut_error_norms_inside(gdm_cut_t c, const double *u, gdm_function_fn exact, void *user, double *norms)
{
  GDM_TRY
  GDM_ARG(c);
  GDM_ARG(u);
  GDM_ARG(exact);
  GDM_ARG(norms);
  const cut::Assembly  &a = c->a;
  std::vector<cut::Pt>  full, ipts, spts;
  const double          l0[3] = {0, 0, 0}, h1[3] = {1, 1, 1};
  cut::tensor_gauss(l0, h1, a.dim, a.gauss, full);
  std::vector<double>   value, grads[3];
  std::vector<uint64_t> dofs;
  int                   idx[3], off[3];
  const double          vol = a.cell_volume();
  double                acc = 0, l1 = 0, linf = 0;
  for (uint64_t cell = 0; cell < a.n_cells; ++cell)
    {
      if (a.location[cell] == cut::OUTSIDE)
        continue;
      a.cell_index(cell, idx);
      a.cell_dofs(idx, off, dofs);
      const std::vector<cut::Pt> *pts = &full;
      if (a.location[cell] == cut::INTERSECTED)
        {
          ipts.clear();
          spts.clear();
          a.cell_rules(idx, ipts, spts);
          pts = &ipts;
        }
      if (pts->empty())
        continue;
      a.shape_at_points(idx, *pts, value, grads);
      for (size_t q = 0; q < pts->size(); ++q)
        {
          double uh = 0, x[3] = {0, 0, 0};
          for (int i = 0; i < a.npc; ++i)
            uh += value[q * a.npc + i] * u[dofs[i]];
          for (int e = 0; e < a.dim; ++e)
            x[e] = a.lo[e] + (idx[e] + (*pts)[q].x[e]) * a.h[e];
          const double diff = uh - exact(x, 0, user);
          acc += diff * diff * (*pts)[q].w * vol;
          l1 += std::fabs(diff) * (*pts)[q].w * vol;
          linf = std::max(linf, std::fabs(diff));
        }
    }
  norms[0] = std::sqrt(acc);
  norms[1] = l1;
  norms[2] = linf;
  GDM_CATCH
}

int gdm_cut_l2_error_inside(gdm_cut_t c, const double *u, gdm_function_fn exact, void *user, double *error)
{
  GDM_ARG_NOTHROW(error);
  double    norms[3] = {0, 0, 0};
  const int rc = gdm_cut_error_norms_inside(c, u, exact, user, norms);
  *error       = norms[0];
  return rc;
}
}
