"""Shared set-up for the GPU parity tests: build the same problem in the product and in the oracle."""
import numpy as np

import oracle as O


def make_pair(dim, p, nc, reps, bc, hi=None, lo=None):
    import gdm_b200 as g
    lo = lo or [0.0] * dim
    hi = hi or [1.0 + 0.25 * d for d in range(dim)]
    gs = g.System(dim, p, nc)
    gs.subdivided_hyper_rectangle(reps, lo, hi)
    os_ = O.System(dim, p, nc)
    os_.subdivided_hyper_rectangle(reps, lo, hi)
    gc, oc = g.AffineConstraints(), O.Constraints()
    if bc == "dirichlet":
        gs.make_zero_boundary_constraints(gc)
        os_.make_zero_boundary_constraints(oc)
    elif bc == "periodic":
        for d in range(dim):
            gs.make_periodicity_constraints(d, gc)
            os_.make_periodicity_constraints(d, oc)
    elif bc == "mixed":  # Dirichlet on the x faces, periodic in the remaining directions
        for s in (0, 1):
            gs.make_zero_boundary_constraints(s, gc)
            os_.make_zero_boundary_constraints(oc, s)
        for d in range(1, dim):
            gs.make_periodicity_constraints(d, gc)
            os_.make_periodicity_constraints(d, oc)
    elif bc == "left":  # a single Dirichlet face
        gs.make_zero_boundary_constraints(0, gc)
        os_.make_zero_boundary_constraints(oc, 0)
    else:
        assert bc == "none"
    gc.close()
    oc.close()
    gs.categorize()
    return gs, gc, os_, oc


def make_operator(gs, gc, kind, b=None, kernel=0, scale=1.0):
    import gdm_b200 as g
    A = g.SparseMatrix()
    m, q = g.MappingQ1(), g.QGauss(gs.fe_degree + 1)
    if kind == "mass":
        g.MatrixCreator.create_mass_matrix(m, gs, q, A, gc, kernel=kernel, scale=scale)
    elif kind == "stiffness":
        g.MatrixCreator.create_laplace_matrix(m, gs, q, A, gc, kernel=kernel, scale=scale)
    elif kind == "advection":
        g.MatrixCreator.create_advection_matrix(m, gs, q, A, gc, b, kernel=kernel, scale=scale)
    elif kind == "advection_t":
        g.MatrixCreator.create_advection_matrix(m, gs, q, A, gc, b, transpose=True, kernel=kernel, scale=scale)
    return A


def oracle_operator(os_, oc, kind, b=None, scale=1.0):
    if kind == "advection":
        return scale * O.kron_operator(os_, oc, "advection", b=b, constrained_diagonal="zero")
    if kind == "advection_t":
        return scale * O.kron_operator(os_, oc, "advection", b=b, constrained_diagonal="zero").T.tocsr()
    A = O.kron_operator(os_, oc, kind, constrained_diagonal="zero")
    D = O.kron_operator(os_, oc, kind) - A  # deal.II's positive diagonal on constrained rows
    return scale * A + abs(scale) * D


def rel_err(a, ref):
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def sphere_overlay(os_, A, radius, band, gamma=0.5, centre=None):
    """Synthetic cut-cell row set in the shape BASELINE config 5 produces (prototypes/cut_poisson_01_gdm.cc:196-329,
    applications/wave/include/gdm/wave/stiffness.h:589-799): the nodes outside the sphere of `radius` (+ band) get
    identity rows (wave/mass.h:246-248), the nodes within `band` of the sphere get irregular rows
    (tensor row + a symmetric positive semi-definite penalty among the band nodes, columns to outside nodes dropped),
    everything inside keeps its tensor-product row.  Returns (rows, rowptr, col, val, A_modified) with A_modified symmetric."""
    import scipy.sparse as sp
    X = os_.node_coordinates()
    c = np.array(centre if centre is not None else [0.5 * (os_.lo[d] + os_.hi[d]) for d in range(os_.dim)])
    dist = np.linalg.norm(X - c, axis=1) - radius
    n = A.shape[0]
    outside = dist > band
    inband = np.abs(dist) <= band
    A = sp.csr_matrix(A)
    # penalty S = gamma * (band-to-band block of A): symmetric positive semi-definite, inside the pattern of A
    Pb = sp.diags(inband.astype(float))
    S = gamma * (Pb @ A @ Pb)
    keep = sp.diags((~outside).astype(float))
    Am = keep @ (A + S) @ keep + sp.diags(outside.astype(float))
    Am = sp.csr_matrix(Am)
    Am.eliminate_zeros()
    rows = np.nonzero(inband | outside)[0]
    rowptr, col, val = [0], [], []
    for r in rows:
        b, e = Am.indptr[r], Am.indptr[r + 1]
        col.extend(Am.indices[b:e].tolist())
        val.extend(Am.data[b:e].tolist())
        rowptr.append(len(col))
    return rows, rowptr, col, val, Am
