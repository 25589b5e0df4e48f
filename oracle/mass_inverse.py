"""Oracle for the Kronecker-direct mass inverse (test infrastructure only).

The reference solves M x = b with preconditioned CG in every Runge-Kutta stage
(`applications/advection/include/gdm/advection/problem.h:236-267`, `prototypes/advection_01_gdm.cc:208-216`).
On a Cartesian grid M = (x)_d A_d on the free nodes plus deal.II's diagonal on the constrained rows, so the solve
factorises into 1D solves.  `kron_mass_solve` restates that with dense 1D matrices (numpy), `bordered_solve_1d` the
band + border elimination the CUDA path uses for a periodic (ring-banded) direction; both are checked against a sparse
direct solve of the oracle's assembled mass matrix in tests/test_oracle_golden.py."""
import numpy as np

from .assemble import matrices_1d


def free_matrix_1d(p, N, h, periodic, dirichlet_lo, dirichlet_hi):
    """(A, f0): the 1D mass matrix on the free nodes [f0, f0 + len(A)) of a direction."""
    M = np.array(matrices_1d(p, N, h)[0], dtype=float)
    M = M.toarray() if hasattr(M, "toarray") else M
    f0, f1 = 0, N + 1
    if periodic:  # C^T M C: node N folded into node 0 (`system.h:427-463`)
        M = M.copy()
        M[0, :] += M[N, :]
        M[:, 0] += M[:, N]
        f1 = N
    if dirichlet_lo:
        f0 = 1
    if dirichlet_hi:
        f1 = min(f1, N)
    return M[f0:f1, f0:f1], f0


def bordered_solve_1d(A, b, nb):
    """Solve A x = b by Cholesky of the leading block B and a Schur complement on the last nb unknowns."""
    m = A.shape[0]
    m1 = m - nb
    B, C, D = A[:m1, :m1], A[:m1, m1:], A[m1:, m1:]
    L = np.linalg.cholesky(B)
    solveB = lambda r: np.linalg.solve(L.T, np.linalg.solve(L, r))
    y = solveB(b[:m1])
    if nb == 0:
        return y
    W = solveB(C)
    S = D - C.T @ W
    z = np.linalg.solve(S, b[m1:] - C.T @ y)
    return np.concatenate([y - W @ z, z])


def kron_mass_solve(system, dirichlet, periodic, diag, b, scale=1.0):
    """x = M^-1 b.  dirichlet[d] = (lo, hi) flags, periodic[d] flag, diag = diagonal of the assembled operator
    (its entries on the constrained rows are what deal.II puts there)."""
    dim, p, nc = system.dim, system.fe_degree, system.n_components
    nn = [n + 1 for n in system.n_subdivisions]
    shape = nn[::-1] + [nc]
    X = np.array(b, dtype=float).reshape(shape)
    free = np.ones(shape, dtype=bool)
    mats = []
    for d in range(dim):
        h = (system.hi[d] - system.lo[d]) / system.n_subdivisions[d]
        A, f0 = free_matrix_1d(p, system.n_subdivisions[d], h, periodic[d], dirichlet[d][0], dirichlet[d][1])
        mats.append((A, f0))
        idx = np.arange(nn[d])
        ok = (idx >= f0) & (idx < f0 + A.shape[0])
        sh = [1] * (dim + 1)
        sh[dim - 1 - d] = nn[d]
        free &= ok.reshape(sh)
    out = np.where(free, X / scale, X / np.array(diag, dtype=float).reshape(shape))
    for d in range(dim):
        A, f0 = mats[d]
        ax = dim - 1 - d
        nb = p if periodic[d] else 0
        sl = [slice(None)] * (dim + 1)
        sl[ax] = slice(f0, f0 + A.shape[0])
        blk = np.moveaxis(out[tuple(sl)], ax, 0)
        flat = blk.reshape(A.shape[0], -1)
        sol = np.stack([bordered_solve_1d(A, flat[:, j], nb) for j in range(flat.shape[1])], axis=1)
        # lines through constrained nodes of the other directions are not part of the product: restore them below
        new = np.moveaxis(sol.reshape(blk.shape), 0, ax)
        keep = out[tuple(sl)]
        fr = free[tuple(sl)]
        out[tuple(sl)] = np.where(fr, new, keep)
    return out.reshape(-1)
