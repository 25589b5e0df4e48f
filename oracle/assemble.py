"""Operator assembly restatement (oracle, test-only).

Two independent routes to the same matrices:

1. `assemble_cell_loop` -- the reference's own loop: per cell, tensor-product
   shape values/gradients of the cell's category at `QGauss(p+1)^dim`, dense
   cell matrix, `constraints.distribute_local_to_global` into a global sparse
   matrix.  Follows `include/gdm/matrix_creator.h:21-61` (mass),
   `tests/poisson_02_gdm.cc:160-206` (stiffness + load vector),
   `tests/mass_02_gdm.cc:100-125` (component-diagonal blocks) and
   `prototypes/advection_01_gdm.cc:164-206` (advection residual).
2. `kron_operator` -- Kronecker sums of the 1D band matrices
   (`matrices_1d`), with the constraints applied algebraically.  This is the
   structure the CUDA kernels exploit; tests prove 1 == 2.
"""
import numpy as np
import scipy.sparse as sp

from .basis import basis_values, gauss_legendre_01
from .system import indices_to_index


# --------------------------------------------------------------------------- 1D
def shape_tables_1d(p):
    """[variant] -> (values[q,k], derivs[q,k]) at QGauss(p+1) on [0,1], plus weights."""
    xq, wq = gauss_legendre_01(p + 1)
    tabs = []
    for v in range(p):
        bv = basis_values(p, v, xq, n_der=1)
        tabs.append((bv[0].T.copy(), bv[1].T.copy()))
    return tabs, xq, wq


def matrices_1d(p, N, h=1.0):
    """Dense 1D matrices on N cells / N+1 nodes in PHYSICAL scaling.

    M = h * sum_c scatter(int phi_a phi_b),  K = 1/h * sum_c scatter(int phi_a' phi_b'),
    C = sum_c scatter(int phi_a phi_b')  (row = test function), f = h * int phi_a.
    """
    assert N >= p
    tabs, xq, wq = shape_tables_1d(p)
    M = np.zeros((N + 1, N + 1))
    K = np.zeros((N + 1, N + 1))
    C = np.zeros((N + 1, N + 1))
    f = np.zeros(N + 1)
    for c in range(N):
        off = 0 if c < p // 2 else min(N, c + p // 2 + 1) - p
        v = c - off
        val, der = tabs[v]
        sl = slice(off, off + p + 1)
        M[sl, sl] += h * np.einsum("q,qa,qb->ab", wq, val, val)
        K[sl, sl] += (1.0 / h) * np.einsum("q,qa,qb->ab", wq, der, der)
        C[sl, sl] += np.einsum("q,qa,qb->ab", wq, val, der)
        f[sl] += h * np.einsum("q,qa->a", wq, val)
    return M, K, C, f


# ------------------------------------------------------------------- cell loop
def _cell_tables(system):
    """Per category: tensor-product values [q, i] and physical gradients [d][q, i]."""
    p, dim = system.fe_degree, system.dim
    tabs, xq, wq = shape_tables_1d(p)
    h = system.h
    cache = {}

    def get(variants):
        key = tuple(variants)
        if key in cache:
            return cache[key]
        vals = [tabs[v][0] for v in variants]
        ders = [tabs[v][1] for v in variants]

        def tensor(fs):  # x fastest in both q and i
            out = fs[0]
            for d in range(1, dim):
                out = np.kron(fs[d], out)
            return out

        value = tensor(vals)
        grads = []
        for d in range(dim):
            fs = [ders[e] / h[e] if e == d else vals[e] for e in range(dim)]
            grads.append(tensor(fs))
        cache[key] = (value, grads)
        return cache[key]

    w = wq
    for d in range(1, dim):
        w = np.kron(wq, w)
    jxw = w * float(np.prod(h))
    xq_ref = xq
    return get, jxw, xq_ref


def _quadrature_points(system, cell, xq):
    idx = system.cell_indices(cell)
    axes = [system.lo[d] + (idx[d] + xq) * system.h[d] for d in range(system.dim)]
    grids = np.meshgrid(*axes[::-1], indexing="ij")
    return np.stack([g.ravel() for g in grids[::-1]], axis=1)  # [q, dim], x fastest


def cell_matrix_scalar(system, cell, kind, get=None, jxw=None, b=None):
    """Dense scalar cell matrix of `kind` in {'mass','stiffness','advection'}."""
    if get is None:
        get, jxw, _ = _cell_tables(system)
    idx = system.cell_indices(cell)
    variants = [system.variant(idx[d], d) for d in range(system.dim)]
    value, grads = get(variants)
    if kind == "mass":
        return np.einsum("q,qi,qj->ij", jxw, value, value)
    if kind == "stiffness":
        return sum(np.einsum("q,qi,qj->ij", jxw, g, g) for g in grads)
    if kind == "advection":  # (phi_i, b . grad phi_j)
        bg = sum(b[d] * grads[d] for d in range(system.dim))
        return np.einsum("q,qi,qj->ij", jxw, value, bg)
    raise ValueError(kind)


def assemble_cell_loop(system, constraints, kind, b=None, literal=False):
    """Global sparse matrix through the reference's cell loop.

    n_components > 1 gives component-diagonal blocks (`tests/mass_02_gdm.cc:104-110`).
    `literal=True` uses the entry-by-entry `Constraints.distribute_local_to_global`
    (slow, the faithful restatement); otherwise a vectorised equivalent that is
    valid for the constraint kinds the reference generates (homogeneous, at most
    one master with weight 1).
    """
    n = system.n_dofs()
    nc = system.n_components
    get, jxw, _ = _cell_tables(system)
    npc = (system.fe_degree + 1) ** system.dim
    mats = {}
    if literal:
        trip = {}
    else:
        target = np.arange(n, dtype=np.int64)
        for i, (entries, inh) in constraints.lines.items():
            assert inh == 0.0 and len(entries) <= 1 and all(w == 1.0 for _, w in entries)
            target[i] = entries[0][0] if entries else -1
        is_con = constraints.constrained_mask(n)
        rows, cols, vals = [], [], []
    for cell in range(system.n_cells()):
        cat = system.active_fe_index(cell)
        if cat not in mats:
            full = np.zeros((npc * nc, npc * nc))
            if kind == "elasticity":
                # 2 eps(u) : eps(v) = grad u : grad v + grad u : grad v^T  (`tests/elasticity_01_gdm.cc:143-160`):
                # block (c, d) = delta_cd sum_a (d_a phi_i, d_a phi_j) + (d_d phi_i, d_c phi_j)
                assert nc == system.dim
                idx = system.cell_indices(cell)
                _, grads = get([system.variant(idx[d], d) for d in range(system.dim)])
                lap = sum(np.einsum("q,qi,qj->ij", jxw, g, g) for g in grads)
                for c in range(nc):
                    for d in range(nc):
                        blk = np.einsum("q,qi,qj->ij", jxw, grads[d], grads[c])
                        full[c * npc:(c + 1) * npc, d * npc:(d + 1) * npc] = blk + (lap if c == d else 0.0)
            else:
                ms = cell_matrix_scalar(system, cell, kind, get, jxw, b)
                for c in range(nc):
                    full[c * npc:(c + 1) * npc, c * npc:(c + 1) * npc] = ms
            mats[cat] = full
        cm = mats[cat]
        dofs = system.get_dof_indices(cell)
        if literal:
            constraints.distribute_local_to_global(cm, None, dofs, trip, None)
        else:
            dofs = np.asarray(dofs)
            t = target[dofs]
            keep = t >= 0
            r = np.repeat(t[keep], keep.sum())
            c_ = np.tile(t[keep], keep.sum())
            rows.append(r)
            cols.append(c_)
            vals.append(cm[np.ix_(keep, keep)].ravel())
            con = is_con[dofs]
            if con.any():
                d = np.abs(np.diag(cm))[con]
                d = np.where(d == 0.0, np.mean(np.abs(np.diag(cm))), d)
                rows.append(dofs[con])
                cols.append(dofs[con])
                vals.append(d)
    if literal:
        keys = np.array(list(trip.keys()), dtype=np.int64).reshape(-1, 2)
        return sp.csr_matrix((np.array(list(trip.values())), (keys[:, 0], keys[:, 1])), shape=(n, n))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return A.tocsr()


def rhs_cell_loop(system, constraints, f):
    """b_i = sum_q f(x_q, comp) phi_i JxW, assembled with the constraints.

    `tests/poisson_02_gdm.cc:193-201`, `tests/mass_02_gdm.cc:114-125`.
    f(points[q, dim], comp) -> values[q].
    """
    n, nc = system.n_dofs(), system.n_components
    get, jxw, xq = _cell_tables(system)
    npc = (system.fe_degree + 1) ** system.dim
    rhs = np.zeros(n)
    for cell in range(system.n_cells()):
        idx = system.cell_indices(cell)
        value, _ = get([system.variant(idx[d], d) for d in range(system.dim)])
        pts = _quadrature_points(system, cell, xq)
        cv = np.zeros(npc * nc)
        for c in range(nc):
            cv[c * npc:(c + 1) * npc] = np.einsum("q,q,qi->i", jxw, np.asarray(f(pts, c), dtype=float) * np.ones(len(jxw)), value)
        constraints.distribute_local_to_global(None, cv, system.get_dof_indices(cell), None, rhs)
    return rhs


def advection_residual_cell_loop(system, constraints, u, b):
    """vec_1 of `prototypes/advection_01_gdm.cc:144-206` (scalar field).

    vec_0 = distribute(u); per cell: flux_q = b . grad u_h(x_q);
    cell_vector(i) -= flux_q phi_i JxW; distribute_local_to_global(vector).
    """
    assert system.n_components == 1
    get, jxw, _ = _cell_tables(system)
    vec0 = constraints.distribute(np.array(u, dtype=float))
    out = np.zeros_like(vec0)
    for cell in range(system.n_cells()):
        idx = system.cell_indices(cell)
        value, grads = get([system.variant(idx[d], d) for d in range(system.dim)])
        dofs = system.get_dof_indices(cell)
        ul = vec0[dofs]
        flux = sum(b[d] * (grads[d] @ ul) for d in range(system.dim))
        cv = -np.einsum("q,q,qi->i", jxw, flux, value)
        constraints.distribute_local_to_global(None, cv, dofs, None, out)
    return out


# ------------------------------------------------------------------- Kronecker
def _kron_all(mats):
    """kron with x fastest: mats = [A_x, A_y, A_z] -> A_z (x) A_y (x) A_x."""
    out = sp.csr_matrix(mats[0])
    for m in mats[1:]:
        out = sp.kron(sp.csr_matrix(m), out, format="csr")
    return out


def kron_terms(system, kind, b=None):
    """List of per-direction 1D matrix lists whose Kronecker products sum to the operator."""
    p, dim = system.fe_degree, system.dim
    one_d = [matrices_1d(p, system.n_subdivisions[d], system.h[d]) for d in range(dim)]
    M = [o[0] for o in one_d]
    K = [o[1] for o in one_d]
    C = [o[2] for o in one_d]
    if kind == "mass":
        return [(1.0, list(M))]
    if kind == "stiffness":
        return [(1.0, [K[e] if e == d else M[e] for e in range(dim)]) for d in range(dim)]
    if kind == "advection":  # (phi_i, b . grad phi_j)
        return [(float(b[d]), [C[e] if e == d else M[e] for e in range(dim)]) for d in range(dim)]
    raise ValueError(kind)


def kron_unconstrained(system, kind, b=None):
    nc = system.n_components
    A = None
    for alpha, mats in kron_terms(system, kind, b):
        T = alpha * _kron_all(mats)
        A = T if A is None else A + T
    if nc > 1:
        A = sp.kron(A, sp.identity(nc), format="csr")  # interleaved components
    return A.tocsr()


def kron_operator(system, constraints, kind, b=None, constrained_diagonal="assembled"):
    """C^T A C + D: the matrix `distribute_local_to_global` produces, built algebraically.

    D carries deal.II's positive diagonal on constrained rows: sum over cells of
    |cell_matrix(i,i)|, which for mass/stiffness equals the unconstrained
    assembled diagonal ("assembled"); "zero" leaves those rows empty (vector
    assembly semantics, e.g. the advection residual).
    """
    n = system.n_dofs()
    A = kron_unconstrained(system, kind, b)
    rows, cols, vals = [], [], []
    con = constraints.constrained_mask(n)
    for i in range(n):
        if not con[i]:
            rows.append(i); cols.append(i); vals.append(1.0)
    for i, (entries, inh) in constraints.lines.items():
        for (j, w) in entries:
            rows.append(i); cols.append(j); vals.append(w)
    Cm = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    Ac = (Cm.T @ A @ Cm).tocsr()
    if constrained_diagonal == "assembled":
        d = np.where(con, np.abs(A.diagonal()), 0.0)
        Ac = Ac + sp.diags(d)
    return Ac.tocsr()
