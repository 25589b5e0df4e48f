"""Cut geometry for a level set of degree q > 1 in 2D (oracle, test-only).

The 2D presets of `applications/wave/wave-app.cc` interpolate the level set with `FE_Q(fe_degree)`
(`level_set_fe_degree = fe_degree`, `:55,146,277`); `oracle/cut.py` and the product's generator handle Q1 level sets only.
This module restates the two deal.II pieces for a tensor-product polynomial level set on each cell so that the 2D goldens
can be compared: `NonMatching::MeshClassifier` (a face is inside / outside when all Bernstein coefficients of the level
set restricted to it are negative / positive, a cell is inside / outside when all its faces are, intersected otherwise)
and the height-function quadrature of `NonMatching::QuadratureGenerator` with `QGauss<1>(n)` (roots found numerically;
the base interval is split at the roots of the level set on the two faces across the height direction).  Cells whose
zero set is not a graph over one coordinate direction are not handled (none occur for the resolved circle of the presets).
"""
import numpy as np
from scipy.optimize import brentq

from .basis import gauss_legendre_01
from .cut import INSIDE, OUTSIDE, INTERSECTED


def gauss_lobatto_01(q):
    """Support points of `FE_Q(q)` on [0,1]: Gauss-Lobatto points."""
    if q == 1:
        return np.array([0.0, 1.0])
    inner = np.polynomial.legendre.Legendre.basis(q).deriv().roots()
    return 0.5 * (np.concatenate([[-1.0], np.sort(inner.real), [1.0]]) + 1.0)


class LevelSetQ:
    def __init__(self, system, fn, degree):
        assert system.dim == 2
        self.q = degree
        self.fn = fn
        self.nodes = gauss_lobatto_01(degree)
        V = np.vander(self.nodes, degree + 1, increasing=True)      # V[i, a] = node_i^a
        self.to_monomial = np.linalg.inv(V)                           # column k: monomial coefficients of l_k
        from math import comb
        B = np.array([[comb(degree, i) * t ** i * (1 - t) ** (degree - i) for i in range(degree + 1)] for t in self.nodes])
        self.to_bernstein = np.linalg.inv(B)
        self._cache = {}

    def local_values(self, system, cell):
        idx = system.cell_indices(cell)
        x = system.lo[0] + (idx[0] + self.nodes) * system.h[0]
        y = system.lo[1] + (idx[1] + self.nodes) * system.h[1]
        X, Y = np.meshgrid(x, y, indexing="ij")
        return np.asarray(self.fn(np.stack([X.ravel(), Y.ravel()], axis=1))).reshape(X.shape)  # [i_x, i_y]

    def classify(self, system):
        out = np.zeros(system.n_cells(), dtype=np.int8)
        for cell in range(system.n_cells()):
            c = self.local_values(system, cell)
            faces = []
            for edge in (c[0, :], c[-1, :], c[:, 0], c[:, -1]):
                b = self.to_bernstein @ edge
                faces.append(INSIDE if b.max() < 0 else (OUTSIDE if b.min() > 0 else INTERSECTED))
            out[cell] = INSIDE if all(f == INSIDE for f in faces) else (OUTSIDE if all(f == OUTSIDE for f in faces) else INTERSECTED)
        return out

    def rules(self, system, cell, n_gauss):
        key = (cell, n_gauss)
        if key not in self._cache:
            self._cache[key] = self._rules(system, cell, n_gauss)
        return self._cache[key]

    def _rules(self, system, cell, n_gauss):
        C = self.to_monomial @ self.local_values(system, cell) @ self.to_monomial.T   # psi = sum C[a, b] x^a y^b
        P = np.polynomial.polynomial
        psi = lambda x, y: P.polyval2d(x, y, C)
        Cx, Cy = P.polyder(C, axis=0), P.polyder(C, axis=1)
        gx, gy = (lambda x, y: P.polyval2d(x, y, Cx)), (lambda x, y: P.polyval2d(x, y, Cy))
        k = 0 if abs(gx(0.5, 0.5)) * (1 + 1e-8) >= abs(gy(0.5, 0.5)) else 1   # height direction; x on (near-)ties
        tt = np.linspace(0, 1, 9)
        S, T = np.meshgrid(tt, tt, indexing="ij")
        dk = gx(T, S) if k == 0 else gy(S, T)
        assert dk.min() > 0 or dk.max() < 0, "zero set is not a graph over a coordinate direction in this cell"
        at = (lambda s, t: psi(t, s)) if k == 0 else (lambda s, t: psi(s, t))   # s: base coordinate, t: height coordinate
        # base interval split at the roots of psi on the faces t = 0 and t = 1
        breaks = [0.0, 1.0]
        for t_face in (0.0, 1.0):
            coeff = (C @ np.array([t_face ** b for b in range(self.q + 1)])) if k == 1 else \
                (np.array([t_face ** a for a in range(self.q + 1)]) @ C)
            for r in P.polyroots(coeff):
                if abs(r.imag) < 1e-12 and 0.0 < r.real < 1.0:
                    breaks.append(float(r.real))
        breaks = sorted(breaks)
        xg, wg = gauss_legendre_01(n_gauss)
        ipts, iw, spts, sw, sn = [], [], [], [], []
        for s0, s1 in zip(breaks[:-1], breaks[1:]):
            if s1 - s0 <= 1e-14:
                continue
            for sq, ws in zip(s0 + (s1 - s0) * xg, (s1 - s0) * wg):
                a, b = at(sq, 0.0), at(sq, 1.0)
                if a * b < 0:
                    r = brentq(lambda t: at(sq, t), 0.0, 1.0, xtol=1e-15, rtol=1e-15)
                    t0, t1 = (0.0, r) if a < 0 else (r, 1.0)
                    pt = (r, sq) if k == 0 else (sq, r)
                    g = np.array([gx(*pt), gy(*pt)])
                    gn = np.linalg.norm(g)
                    spts.append(pt)
                    sw.append(ws * gn / abs(g[k]))
                    sn.append(g / gn)
                elif a < 0:
                    t0, t1 = 0.0, 1.0
                else:
                    continue
                for tq, wt in zip(t0 + (t1 - t0) * xg, (t1 - t0) * wg):
                    ipts.append((tq, sq) if k == 0 else (sq, tq))
                    iw.append(ws * wt)
        arr = lambda a, m: np.array(a, dtype=float).reshape(-1, m)
        return (arr(ipts, 2), np.array(iw)), (arr(spts, 2), np.array(sw), arr(sn, 2))
