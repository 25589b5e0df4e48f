"""Matrix-free oracle apply at any grid size -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`kron_operator` (assemble.py) builds  C^T A C + D  as an explicit scipy matrix; at the BASELINE size
(257^3 DoFs, 343 nnz/row) that matrix needs ~70 GB.  This module applies the *same* operator without
forming it:

    A   = sum_t alpha_t  A_t,z (x) A_t,y (x) A_t,x          (`kron_terms`, the 1D matrices of
                                                              `matrices_1d`: cell loop == Kronecker sum is
                                                              proven in tests/test_oracle_golden.py)
    C   = C_z (x) C_y (x) C_x                                (Cartesian constraints: zero-Dirichlet faces
                                                              `system.h:466-508` and periodic folds
                                                              `system.h:427-463` are per direction; C_d is
                                                              read off the closed constraint lines)
    C^T A C = sum_t alpha_t  (C_z^T A_t,z C_z) (x) (C_y^T A_t,y C_y) (x) (C_x^T A_t,x C_x)
    D   = diag(|A_ii|) on constrained rows                   (deal.II `distribute_local_to_global`)

Each factor is applied along its axis with a scipy CSR product (O(n p d) work).  It is checked against
`kron_operator` at small sizes in tests/test_oracle_golden.py::test_kron_apply_matches_matrix.
Reference semantics: `SparseMatrix::vmult` of the matrix assembled in tests/poisson_02_gdm.cc:160-216.
"""
import numpy as np
import scipy.sparse as sp

from .assemble import kron_terms
from .system import indices_to_index


def constraint_matrices_1d(system, constraints):
    """[C_d] (scipy CSR, n_d x n_d) per direction, read off the closed constraint lines.

    Probes the DoFs (i along d, an interior index in every other direction, component 0): a line without
    entries is a zero-Dirichlet node, a line (j, w) a periodic identification with node j.
    """
    dim, nn, nc = system.dim, system.n_nodes, system.n_components
    mid = [min(n // 2, n - 2) if n > 2 else 0 for n in nn]
    out = []
    for d in range(dim):
        rows, cols, vals = [], [], []
        for i in range(nn[d]):
            idx = list(mid)
            idx[d] = i
            dof = indices_to_index(idx, nn) * nc
            if dof in constraints.lines:
                entries, inh = constraints.lines[dof]
                assert inh == 0.0, "kron_apply: homogeneous constraints only"
                for (j, w) in entries:
                    jd = []
                    s = j // nc
                    for e in range(dim):
                        jd.append(s % nn[e])
                        s //= nn[e]
                    assert all(jd[e] == idx[e] for e in range(dim) if e != d), "constraint is not per direction"
                    rows.append(i); cols.append(jd[d]); vals.append(w)
            else:
                rows.append(i); cols.append(i); vals.append(1.0)
        out.append(sp.csr_matrix((vals, (rows, cols)), shape=(nn[d], nn[d])))
    return out


def _apply_axis(M, X, axis):
    """Y = M applied along `axis` of the array X (any other axes untouched)."""
    Xm = np.moveaxis(X, axis, 0)
    shp = Xm.shape
    Y = M @ Xm.reshape(shp[0], -1)
    return np.moveaxis(np.asarray(Y).reshape(shp), 0, axis)


class KronApply:
    """y = (C^T A C + D) x  without forming the matrix; `diagonal()` for Jacobi."""

    def __init__(self, system, constraints, kind, b=None, scale=1.0, constrained_diagonal="assembled"):
        self.system = system
        dim, nn = system.dim, system.n_nodes
        self.shape_grid = tuple(nn[::-1]) + ((system.n_components,) if system.n_components > 1 else ())
        Cs = constraint_matrices_1d(system, constraints)
        self.terms = []      # (alpha, [C_d^T A_d C_d])
        diag = 0.0           # unconstrained diagonal, grid shaped
        cdiag = 0.0          # diagonal of C^T A C
        for alpha, mats in kron_terms(system, kind, b):
            fold = [sp.csr_matrix(Cs[d].T @ sp.csr_matrix(mats[d]) @ Cs[d]) for d in range(dim)]
            self.terms.append((scale * alpha, fold))
            t, tc = scale * alpha, scale * alpha
            for d in range(dim):  # outer products, x fastest: axis dim-1-d
                sh = [1] * dim
                sh[dim - 1 - d] = nn[d]
                t = t * np.diag(mats[d]).reshape(sh)
                tc = tc * fold[d].diagonal().reshape(sh)
            diag, cdiag = diag + t, cdiag + tc
        free = 1.0
        for d in range(dim):
            sh = [1] * dim
            sh[dim - 1 - d] = nn[d]
            fd = np.array([1.0 if (Cs[d][i, i] == 1.0 and Cs[d][i].nnz == 1) else 0.0 for i in range(nn[d])])
            free = free * fd.reshape(sh)
        self.constrained = (free == 0.0)
        self.D = np.where(self.constrained, np.abs(diag), 0.0) if constrained_diagonal == "assembled" else np.zeros_like(diag)
        self._diag = cdiag + self.D

    def _grid(self, x):
        return np.asarray(x, dtype=float).reshape(self.shape_grid)

    def matvec(self, x):
        X = self._grid(x)
        dim = self.system.dim
        Y = np.zeros_like(X)
        for alpha, fold in self.terms:
            T = X
            for d in range(dim):
                T = _apply_axis(fold[d], T, dim - 1 - d)
            Y += alpha * T
        if self.system.n_components > 1:
            Y += self.D[..., None] * X
        else:
            Y += self.D * X
        return Y.reshape(-1)

    __matmul__ = lambda self, x: self.matvec(x)

    def diagonal(self):
        d = self._diag
        if self.system.n_components > 1:
            d = np.repeat(d[..., None], self.system.n_components, axis=-1)
        return d.reshape(-1).copy()

    def constrained_mask(self):
        m = self.constrained
        if self.system.n_components > 1:
            m = np.repeat(m[..., None], self.system.n_components, axis=-1)
        return m.reshape(-1).copy()


def kron_apply(system, constraints, kind, x, b=None, scale=1.0):
    return KronApply(system, constraints, kind, b, scale).matvec(x)
