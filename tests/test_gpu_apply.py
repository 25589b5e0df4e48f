"""GPU parity: operator application through the C ABI vs the oracle (tolerance: relative 1e-12
of max|y_ref| on a single apply, as BASELINE.json's north_star states)."""
import numpy as np
import pytest

import oracle as O

from helpers import make_pair, make_operator, oracle_operator, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    # dim, p, nc, reps, bc
    (1, 1, 1, [7], "dirichlet"),
    (1, 3, 1, [10], "dirichlet"),
    (1, 5, 1, [13], "periodic"),
    (1, 9, 1, [21], "none"),
    (2, 3, 1, [20, 20], "dirichlet"),
    (2, 3, 2, [9, 7], "none"),
    (2, 5, 1, [12, 15], "periodic"),
    (2, 7, 1, [16, 15], "mixed"),
    (3, 1, 1, [5, 4, 3], "dirichlet"),
    (3, 3, 1, [17, 9, 12], "dirichlet"),
    (3, 3, 1, [9, 8, 7], "periodic"),
    (3, 3, 2, [7, 8, 9], "left"),
    (3, 5, 1, [11, 12, 13], "mixed"),
    (3, 5, 1, [12, 11, 13], "none"),
]


@pytest.mark.parametrize("dim,p,nc,reps,bc", CASES)
@pytest.mark.parametrize("kind", ["mass", "stiffness", "advection", "advection_t"])
def test_generic_apply_matches_oracle(lib, dim, p, nc, reps, bc, kind):
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(dim, p, nc, reps, bc)
    b = [1.0, 0.15, -0.05][:dim]
    scale = -1.0 if kind == "advection" else 1.0
    A = make_operator(gs, gc, kind, b=b, kernel=g.capi.KERNEL_GENERIC, scale=scale)
    Ao = oracle_operator(os_, oc, kind, b=b, scale=scale)
    rng = np.random.default_rng(0)
    xh = rng.uniform(-1, 1, gs.n_dofs())
    x, y = g.Vector(gs, xh), g.Vector(gs)
    A.vmult(y, x)
    ref = Ao @ xh
    assert rel_err(y.numpy(), ref) <= TOL
    # vmult_add and linearity
    y2 = g.Vector(gs, xh)
    A.vmult_add(y2, x)
    assert rel_err(y2.numpy(), ref + xh) <= TOL
    # host-buffer entry point
    yh = np.zeros_like(xh)
    A.vmult_host(yh, xh)
    assert rel_err(yh, ref) <= TOL


@pytest.mark.parametrize("dim,p,nc,reps,bc", [c for c in CASES if c[0] >= 2][:6])
def test_diagonal_and_jacobi(lib, dim, p, nc, reps, bc):
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(dim, p, nc, reps, bc)
    for kind in ("mass", "stiffness"):
        A = make_operator(gs, gc, kind, kernel=g.capi.KERNEL_GENERIC)
        Ao = oracle_operator(os_, oc, kind)
        d = A.diagonal().numpy()
        assert rel_err(d, Ao.diagonal()) <= TOL


def test_constraints_distribute_and_set_zero(lib):
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(3, 3, 2, [6, 7, 8], "mixed")
    rng = np.random.default_rng(1)
    xh = rng.uniform(-1, 1, gs.n_dofs())
    v = g.Vector(gs, xh)
    gc.distribute(v)
    assert np.array_equal(v.numpy(), oc.distribute(xh.copy()))
    v = g.Vector(gs, xh)
    gc.set_zero(v)
    assert np.array_equal(v.numpy(), oc.set_zero(xh.copy()))
    assert gc.n_constraints() == len(oc.lines)


def test_vector_operations(lib):
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(3, 3, 1, [9, 10, 11], "none")
    rng = np.random.default_rng(2)
    a, b = rng.uniform(-1, 1, gs.n_dofs()), rng.uniform(-1, 1, gs.n_dofs())
    va, vb = g.Vector(gs, a), g.Vector(gs, b)
    assert abs(va * vb - a @ b) <= 1e-13 * abs(a @ b) + 1e-13
    assert abs(va.l2_norm() - np.linalg.norm(a)) <= 1e-13 * np.linalg.norm(a)
    assert va.linfty_norm() == np.abs(a).max()
    va.add(0.5, vb)
    assert np.allclose(va.numpy(), a + 0.5 * b, rtol=0, atol=1e-15)
    va.sadd(2.0, -1.0, vb)
    assert np.allclose(va.numpy(), 2 * (a + 0.5 * b) - b, rtol=0, atol=1e-15)
    va.scale(3.0)
    va.scale(vb)
    assert np.allclose(va.numpy(), 3 * (2 * (a + 0.5 * b) - b) * b, rtol=0, atol=1e-14)
    vc = va.copy()
    assert np.array_equal(vc.numpy(), va.numpy())
    vc.set(1.5)
    assert np.array_equal(vc.numpy(), np.full(gs.n_dofs(), 1.5))
    assert abs(vc.l2_norm() - 1.5 * np.sqrt(gs.n_dofs())) < 1e-10  # pads stay zero


def test_csr_overlay_rows_replace_tensor_rows(lib):
    """Irregular rows (cut cells / ghost penalty) as CSR rows that replace the regular result."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(2, 3, 1, [10, 9], "dirichlet")
    A = make_operator(gs, gc, "stiffness", kernel=g.capi.KERNEL_GENERIC)
    Ao = oracle_operator(os_, oc, "stiffness").tolil()
    rng = np.random.default_rng(3)
    n = gs.n_dofs()
    rows = np.sort(rng.choice(n, 17, replace=False))
    rowptr, col, val = [0], [], []
    for r in rows:
        cols = np.sort(rng.choice(n, 23, replace=False))
        vals = rng.uniform(-1, 1, 23)
        Ao[r, :] = 0.0
        for c, v in zip(cols, vals):
            Ao[r, c] = v
        col += list(cols)
        val += list(vals)
        rowptr.append(len(col))
    A.attach_csr(rows, rowptr, col, val)
    xh = rng.uniform(-1, 1, n)
    x, y = g.Vector(gs, xh), g.Vector(gs)
    A.vmult(y, x)
    ref = Ao.tocsr() @ xh
    assert rel_err(y.numpy(), ref) <= TOL
    y2 = g.Vector(gs, xh)
    A.vmult_add(y2, x)
    assert rel_err(y2.numpy(), ref + xh) <= TOL


def test_lumped_mass(lib):
    import gdm_b200 as g
    import oracle as O
    gs, gc, os_, oc = make_pair(2, 3, 1, [8, 9], "periodic")
    inv = g.Vector(gs)
    g.MatrixCreator.create_lumped_mass_matrix(g.MappingQ1(), gs, g.QGauss(4), inv, gc)
    M = O.kron_unconstrained(os_, "mass")
    rows = np.asarray(M.sum(axis=1)).ravel()
    lumped = np.zeros_like(rows)
    for i in range(len(rows)):  # distribute_local_to_global on a vector
        if oc.is_constrained(i):
            for (j, w) in oc.lines[i][0]:
                lumped[j] += w * rows[i]
        else:
            lumped[i] += rows[i]
    ref = np.where(lumped != 0, 1.0 / np.where(lumped != 0, lumped, 1.0), 0.0)
    assert rel_err(inv.numpy(), ref) <= TOL


@pytest.mark.parametrize("dim,p,reps,bc", [(3, 3, [12, 11, 13], "dirichlet"), (2, 5, [14, 13], "periodic"), (3, 3, [35, 33, 20], "periodic")])
@pytest.mark.parametrize("kind", ["mass", "stiffness", "advection", "advection_t"])
def test_tvmult(lib, dim, p, reps, bc, kind):
    """SparseMatrix::Tvmult: the transpose, not vmult, for the (unsymmetric) advection operators."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(dim, p, 1, reps, bc)
    b = [1.0, 0.15, -0.05][:dim]
    A = make_operator(gs, gc, kind, b=b)
    Ao = oracle_operator(os_, oc, kind, b=b)
    xh = np.random.default_rng(5).uniform(-1, 1, gs.n_dofs())
    x, y = g.Vector(gs, xh), g.Vector(gs)
    A.Tvmult(y, x)
    ref = Ao.T @ xh
    assert rel_err(y.numpy(), ref) <= 1e-12
    A.vmult(y, x)
    assert rel_err(y.numpy(), Ao @ xh) <= 1e-12


@pytest.mark.parametrize("p,reps,kernel", [(3, [20, 19, 21], "fused"), (1, [24, 24, 24], "fused"), (3, [14, 13, 12], "generic")])
def test_sphere_overlay_apply_diagonal_cg(lib, p, reps, kernel):
    """BASELINE config 5 in miniature: irregular CSR rows around a sphere + identity rows outside on top of the fused
    tensor-product apply; single apply, diagonal (Jacobi with CSR rows) and CG iteration count against the oracle matrix."""
    import gdm_b200 as g
    from helpers import sphere_overlay
    gs, gc, os_, oc = make_pair(3, p, 1, reps, "none", hi=[1.0, 1.0, 1.0])
    A = make_operator(gs, gc, "stiffness", kernel=g.capi.KERNEL_FUSED if kernel == "fused" else g.capi.KERNEL_GENERIC)
    h = 1.0 / min(reps)
    rows, rowptr, col, val, Am = sphere_overlay(os_, oracle_operator(os_, oc, "stiffness"), 0.3, (p + 1) * h)
    A.attach_csr(rows, rowptr, col, val)
    n = gs.n_dofs()
    xh = np.random.default_rng(11).uniform(-1, 1, n)
    x, y = g.Vector(gs, xh), g.Vector(gs)
    A.vmult(y, x)
    assert rel_err(y.numpy(), Am @ xh) <= TOL
    d = g.Vector(gs)
    A.diagonal(d)
    assert rel_err(d.numpy(), Am.diagonal()) <= TOL
    # CG with Jacobi on the modified (symmetric positive definite) operator; rhs supported inside the sphere
    inside = np.linalg.norm(os_.node_coordinates() - 0.5, axis=1) < 0.3
    bh = np.where(inside, 1.0, 0.0) * h ** 3
    b, u = g.Vector(gs, bh), g.Vector(gs)
    ctl = g.ReductionControl(2000, 1e-12, 1e-6)
    P = g.PreconditionJacobi()
    P.initialize(A)
    g.SolverCG(ctl).solve(A, u, b, P)
    octl = O.ReductionControl(2000, 1e-12, 1e-6)
    uo = O.solver_cg(Am, np.zeros(n), bh, O.PreconditionJacobi(Am), octl)
    assert abs(ctl.last_step() - octl.last_step()) <= 1
    assert rel_err(u.numpy(), uo) <= 1e-6


@pytest.mark.parametrize("dim,p,reps,kernel", [(2, 3, [9, 8], "generic"), (3, 3, [10, 9, 12], "fused"), (3, 5, [12, 12, 12], "fused")])
def test_inhomogeneous_dirichlet(lib, dim, p, reps, kernel):
    """System::interpolate_boundary_values (system.h:511-547) + the rhs part of distribute_local_to_global: the harmonic
    polynomial g = x^2 - y^2 (+ z) lies in the GDM space (degree <= p), so -Laplace u = 0, u = g on the boundary returns
    the interpolant of g; the condensed right-hand side is checked against the oracle's matrices."""
    import gdm_b200 as g
    hi = [1.0 + 0.25 * d for d in range(dim)]
    gs = g.System(dim, p, 1)
    gs.subdivided_hyper_rectangle(reps, [0.0] * dim, hi)
    so = O.System(dim, p)
    so.subdivided_hyper_rectangle(reps, [0.0] * dim, hi)
    gfun = lambda pt, c=0: pt[0] ** 2 - pt[1] ** 2 + (0.5 * pt[2] if dim == 3 else 0.0)
    gc = g.AffineConstraints()
    gs.interpolate_boundary_values(g.MappingQ1(), 0, gfun, gc)
    gc.close()
    co = O.Constraints()
    so.make_zero_boundary_constraints(co)
    co.close()
    A = make_operator(gs, gc, "stiffness", kernel=g.capi.KERNEL_FUSED if kernel == "fused" else g.capi.KERNEL_GENERIC)
    # oracle: lifting with the unconstrained matrix, deal.II's diagonal on the constrained rows
    Au = O.kron_unconstrained(so, "stiffness")
    Ac = O.kron_operator(so, co, "stiffness")
    X = so.node_coordinates()
    gh = np.array([gfun(list(x) + [0.0] * (3 - dim)) for x in X])
    con = co.constrained_mask(so.n_dofs())
    gb = np.where(con, gh, 0.0)
    rhs_ref = np.where(con, Ac.diagonal() * gb, -(Au @ gb))
    rhs = g.Vector(gs)
    gc.condense_rhs(A, rhs)
    assert rel_err(rhs.numpy(), rhs_ref) <= TOL
    u = g.Vector(gs)
    ctl = g.ReductionControl(2000, 1e-14, 1e-12)
    g.SolverCG(ctl).solve(A, u, rhs, g.PreconditionIdentity())
    gc.distribute(u)
    assert np.abs(u.numpy() - gh).max() <= 1e-9 * np.abs(gh).max()
