"""GPU parity: CG and Runge-Kutta through the C ABI vs the reference's goldens and the oracle."""
import os

import numpy as np
import pytest

import oracle as O
from helpers import make_pair, make_operator, oracle_operator, rel_err

pytestmark = pytest.mark.gpu


def _golden(golden_dir, name):
    return open(os.path.join(golden_dir, name)).read()


@pytest.mark.parametrize("p", [1, 3, 5, 7, 9])
def test_poisson_01_gdm(lib, golden_dir, p):
    """tests/poisson_01_gdm.cc: 1D, N=10, plain CG(100,1e-10,1e-4): count, values, L2 error."""
    import gdm_b200 as g
    tok = _golden(golden_dir, "poisson_01_gdm.output").split()
    pos = 14 * [1, 3, 5, 7, 9].index(p)
    gs, gc, os_, oc = make_pair(1, p, 1, [10], "dirichlet", hi=[1.0])
    A = make_operator(gs, gc, "stiffness")
    rhs = g.Vector(gs, O.rhs_cell_loop(os_, oc, lambda pts, c: 1.0))
    sol = g.Vector(gs)
    ctl = g.ReductionControl(100, 1e-10, 1e-4)
    g.SolverCG(ctl).solve(A, sol, rhs, g.PreconditionIdentity())
    assert ctl.last_step() == int(tok[pos]) == 5
    vals = sol.numpy()
    gold = np.array([float(t) for t in tok[pos + 1:pos + 12]])
    assert np.allclose([float("%g" % v) for v in vals], gold, atol=1e-12)
    exact = lambda pt, c: 0.125 - 0.5 * (pt[0] - 0.5) ** 2
    cell = g.VectorTools.integrate_difference(g.MappingQ1(), gs, sol, exact, None, g.QGauss(p + 1), "L2_norm")
    err = g.VectorTools.compute_global_error(None, cell)
    assert "%14.8f" % err == "%14.8f" % float(tok[pos + 13])


@pytest.mark.parametrize("dim", [1, 2])
def test_poisson_02_gdm_values(lib, golden_dir, dim):
    """tests/poisson_02_gdm.cc solution values (golden printed with 6 significant digits)."""
    import gdm_b200 as g
    tok = _golden(golden_dir, "poisson_02_gdm.mpirun=1.output").split()
    gold = np.array([float(t) for t in (tok[1:22] if dim == 1 else tok[23:23 + 441])])
    gs, gc, os_, oc = make_pair(dim, 3, 1, [20] * dim, "dirichlet", hi=[1.0] * dim)
    A = make_operator(gs, gc, "stiffness")
    rhs = g.Vector(gs, O.rhs_cell_loop(os_, oc, lambda pts, c: 1.0))
    sol = g.Vector(gs)
    ctl = g.ReductionControl(1000, 1e-14, 1e-12)
    g.SolverCG(ctl).solve(A, sol, rhs, g.PreconditionIdentity())
    gc.distribute(sol)
    assert np.allclose([float("%g" % v) for v in sol.numpy()], gold, atol=1e-12)
    # plain CG with the test's own control needs 10 (1D) / 23 (2D) iterations (oracle-pinned)
    sol2 = g.Vector(gs)
    ctl = g.ReductionControl(100, 1e-10, 1e-4)
    g.SolverCG(ctl).solve(A, sol2, rhs, g.PreconditionIdentity())
    assert ctl.last_step() == (10 if dim == 1 else 23)


@pytest.mark.parametrize("nc,name", [(1, "mass_01_gdm.output"), (2, "mass_02_gdm.output")])
def test_mass_0x_gdm(lib, golden_dir, nc, name):
    """tests/mass_01_gdm.cc / mass_02_gdm.cc: Jacobi-CG L2 projection, error printed with %g."""
    import gdm_b200 as g
    gold = _golden(golden_dir, name).split()[1]
    gs, gc, os_, oc = make_pair(2, 3, nc, [40, 40], "none", hi=[1.0, 1.0])
    A = g.SparseMatrix()
    g.MatrixCreator.create_mass_matrix(g.MappingQ1(), gs, g.QGauss(4), A, gc)
    f = lambda pts, c: pts[:, 0] + c
    rhs = g.Vector(gs, O.rhs_cell_loop(os_, oc, f))
    sol = g.Vector(gs)
    pre = g.PreconditionJacobi()
    pre.initialize(A)
    ctl = g.ReductionControl(100, 1e-10, 1e-8)
    g.SolverCG(ctl).solve(A, sol, rhs, pre)
    assert ctl.last_step() == 18
    cell = g.VectorTools.integrate_difference(g.MappingQ1(), gs, sol, lambda pt, c: pt[0] + c, None, g.QGauss(4))
    assert "%g" % g.VectorTools.compute_global_error(None, cell) == gold


def test_cg_failure_is_reported(lib):
    """deal.II throws SolverControl::NoConvergence when max_steps is hit."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(2, 3, 1, [20, 20], "dirichlet")
    A = make_operator(gs, gc, "stiffness")
    rhs = g.Vector(gs, O.rhs_cell_loop(os_, oc, lambda pts, c: 1.0))
    sol = g.Vector(gs)
    ctl = g.ReductionControl(3, 1e-30, 1e-30)
    with pytest.raises(g.NoConvergence):
        g.SolverCG(ctl).solve(A, sol, rhs, g.PreconditionIdentity())
    assert ctl.last_step() == 3


@pytest.mark.parametrize("dim,p,reps,bc,pre", [(3, 3, [12, 10, 9], "dirichlet", "identity"),
                                               (3, 3, [12, 10, 9], "dirichlet", "jacobi"),
                                               (2, 5, [16, 18], "periodic", "jacobi"),
                                               (3, 5, [11, 11, 12], "left", "identity")])
def test_cg_iteration_parity_with_oracle(lib, dim, p, reps, bc, pre):
    """Same matrix, same rhs, same control => same iteration count and solution as deal.II's loop."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(dim, p, 1, reps, bc)
    kind = "mass" if bc == "periodic" else "stiffness"
    A = make_operator(gs, gc, kind)
    Ao = oracle_operator(os_, oc, kind)
    bh = O.rhs_cell_loop(os_, oc, lambda pts, c: np.sin(3 * pts[:, 0]) + 1.0)
    rhs, sol = g.Vector(gs, bh), g.Vector(gs)
    ctl = g.ReductionControl(2000, 1e-12, 1e-9)
    octl = O.ReductionControl(2000, 1e-12, 1e-9)
    if pre == "jacobi":
        P, Po = g.PreconditionJacobi(), O.PreconditionJacobi(Ao)
        P.initialize(A)
    else:
        P, Po = g.PreconditionIdentity(), O.PreconditionIdentity()
    g.SolverCG(ctl).solve(A, sol, rhs, P)
    xo = O.solver_cg(Ao, np.zeros(gs.n_dofs()), bh, Po, octl)
    # The stopping iteration may move by one when the residual crosses the threshold within rounding:
    # fused multiply-adds and a different (but fixed) summation order of the dots.  Counts pinned by the
    # reference's goldens (poisson_01: 5, mass_01/02: 18) are asserted exactly in the tests above.
    slack = max(1, octl.last_step() // 100)
    assert abs(ctl.last_step() - octl.last_step()) <= slack
    assert abs(ctl.initial_value() - octl.initial_value()) <= 1e-13 * octl.initial_value()
    # Both iterates satisfy ||r|| < 1e-9 ||r0||; when the stopping iteration moves by one (see above) they differ by
    # one CG update, i.e. by O(reduction * cond(P^-1 A)) relative -- 1e-9 only holds for identical counts.
    tol = 1e-9 if ctl.last_step() == octl.last_step() else 1e-7
    assert rel_err(sol.numpy(), xo) <= tol


def test_advection_rk4_matches_oracle(lib):
    """prototypes/advection_01_gdm.cc in small: periodic, p=5, RK4, Jacobi-CG mass inversion per stage."""
    import gdm_b200 as g
    n, p, dim = 12, 5, 2
    b = [1.0, 0.15]
    gs, gc, os_, oc = make_pair(dim, p, 1, [n, n], "periodic", hi=[1.0, 1.0])
    M = g.SparseMatrix()
    g.MatrixCreator.create_mass_matrix(g.MappingQ1(), gs, g.QGauss(p + 1), M, gc)
    R = make_operator(gs, gc, "advection", b=b, scale=-1.0)
    Mo = oracle_operator(os_, oc, "mass")
    Ro = oracle_operator(os_, oc, "advection", b=b, scale=-1.0)
    u0 = lambda pts, c: np.sin(2 * np.pi * pts[:, 0]) * np.cos(2 * np.pi * pts[:, 1])
    uh = O.interpolate(os_, u0)
    sol = g.Vector(gs)
    g.VectorTools.interpolate(g.MappingQ1(), gs, lambda pt, c: np.sin(2 * np.pi * pt[0]) * np.cos(2 * np.pi * pt[1]), sol)
    assert np.abs(sol.numpy() - uh).max() < 1e-15
    pre = g.PreconditionJacobi()
    pre.initialize(M)
    tmp0, tmp1 = g.Vector(gs), g.Vector(gs)
    iters = []

    def f(t, y, out):
        tmp0.equ(y)
        gc.distribute(tmp0)
        R.vmult(tmp1, tmp0)
        out.set(0.0)
        ctl = g.ReductionControl(100, 1e-10, 1e-8)
        g.SolverCG(ctl).solve(M, out, tmp1, pre)
        iters.append(ctl.last_step())

    oiters = []

    def fo(t, y):
        v0 = oc.distribute(y.copy())
        ctl = O.ReductionControl(100, 1e-10, 1e-8)
        out = O.solver_cg(Mo, np.zeros_like(y), Ro @ v0, O.PreconditionJacobi(Mo), ctl)
        oiters.append(ctl.last_step())
        return out

    rk, rko = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER), O.ExplicitRungeKutta4()
    time, dt = g.DiscreteTime(0.0, 0.1, 0.5 / n), 0.5 / n
    uo, t = uh.copy(), 0.0
    while not time.is_at_end():
        rk.evolve_one_time_step(f, time.get_current_time(), time.get_next_step_size(), sol)
        gc.distribute(sol)
        t, uo = rko.evolve_one_time_step(fo, t, time.get_next_step_size(), uo)
        oc.distribute(uo)
        time.advance_time()
    assert iters == oiters
    assert rel_err(sol.numpy(), uo) <= 1e-10


def _poisson3d_rhs(g, gs, gc):
    """b_i = int phi_i over the unit cube (= M 1 because the GDM basis is a partition of unity), zero on constrained rows."""
    free = g.AffineConstraints()
    free.close()
    M = g.SparseMatrix()
    g.MatrixCreator.create_mass_matrix(g.MappingQ1(), gs, g.QGauss(gs.fe_degree + 1), M, free)
    ones, b = g.Vector(gs), g.Vector(gs)
    ones.set(1.0)
    M.vmult(b, ones)
    gc.set_zero(b)
    return b


@pytest.mark.parametrize("N", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("pre", ["identity", "jacobi"])
def test_cg_poisson3d_golden_counts(lib, N, pre):
    """SURVEY 8d CG: -Laplace u = 1, ReductionControl(10000, 1e-10, 1e-8); iteration counts, residuals and the
    solution's maximum against the oracle's converged solves (tests/golden/cg_poisson3d.json, made by
    tests/golden/make_cg_golden.py; 256^3 = BASELINE config 2).  Reference call site: tests/poisson_02_gdm.cc:213-215."""
    import json
    import os
    import gdm_b200 as g
    gold = [r for r in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cg_poisson3d.json")))
            if r["N"] == N and r["p"] == 3 and r["precondition"] == pre][0]
    gs, gc, _, _ = make_pair(3, 3, 1, [N, N, N], "dirichlet", hi=[1.0, 1.0, 1.0])
    A = make_operator(gs, gc, "stiffness")
    b = _poisson3d_rhs(g, gs, gc)
    u = g.Vector(gs)
    ctl = g.ReductionControl(10000, 1e-10, 1e-8)
    if pre == "identity":
        P = g.PreconditionIdentity()
    else:
        P = g.PreconditionJacobi()
        P.initialize(A)
    g.SolverCG(ctl).solve(A, u, b, P)
    assert abs(ctl.initial_value() - gold["initial_residual"]) <= 1e-12 * gold["initial_residual"]
    # the stopping test compares a residual norm of ~1e-10 with the tolerance: rounding may move the count by one
    assert abs(ctl.last_step() - gold["iterations"]) <= 1, (ctl.last_step(), gold["iterations"])
    if ctl.last_step() == gold["iterations"]:
        assert abs(ctl.last_value() - gold["final_residual"]) <= 5e-2 * gold["final_residual"]
    assert abs(u.linfty_norm() - gold["u_max"]) <= 1e-8 * gold["u_max"]


def test_advection_rk4_3d_periodic_fused(lib):
    """BASELINE config 3 at a small size: 3D periodic p=5 advection, RK4, Jacobi-CG mass solves (app setting
    (1000, 1e-20, 1e-14) is unreachable in a few steps at this size: prototype setting (100, 1e-10, 1e-8)); fused
    kernel for both operators; per-stage CG counts identical to the oracle (prototypes/advection_01_gdm.cc:144-224,268-281)."""
    import gdm_b200 as g
    n, p = 12, 5
    b = [1.0, 0.15, -0.05]
    gs, gc, os_, oc = make_pair(3, p, 1, [n, n, n], "periodic", hi=[1.0, 1.0, 1.0])
    M = make_operator(gs, gc, "mass")
    R = make_operator(gs, gc, "advection", b=b, scale=-1.0)
    assert M.kernel_used() == g.capi.KERNEL_FUSED and R.kernel_used() == g.capi.KERNEL_FUSED
    Mo = oracle_operator(os_, oc, "mass")
    Ro = oracle_operator(os_, oc, "advection", b=b, scale=-1.0)
    u0 = lambda pts, c: np.sin(2 * np.pi * pts[:, 0]) * np.cos(2 * np.pi * pts[:, 1]) * np.cos(2 * np.pi * pts[:, 2])
    uh = O.interpolate(os_, u0)
    sol = g.Vector(gs, uh)
    pre = g.PreconditionJacobi()
    pre.initialize(M)
    tmp0, tmp1 = g.Vector(gs), g.Vector(gs)
    iters, oiters = [], []

    def f(t, y, out):
        tmp0.equ(y)
        gc.distribute(tmp0)
        R.vmult(tmp1, tmp0)
        out.set(0.0)
        ctl = g.ReductionControl(100, 1e-10, 1e-8)
        g.SolverCG(ctl).solve(M, out, tmp1, pre)
        iters.append(ctl.last_step())

    def fo(t, y):
        v0 = oc.distribute(y.copy())
        ctl = O.ReductionControl(100, 1e-10, 1e-8)
        out = O.solver_cg(Mo, np.zeros_like(y), Ro @ v0, O.PreconditionJacobi(Mo), ctl)
        oiters.append(ctl.last_step())
        return out

    rk, rko = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER), O.ExplicitRungeKutta4()
    dt = 0.5 / n
    uo, t = uh.copy(), 0.0
    for step in range(3):
        rk.evolve_one_time_step(f, t, dt, sol)
        gc.distribute(sol)
        t, uo = rko.evolve_one_time_step(fo, t, dt, uo)
        oc.distribute(uo)
    # the stopping test compares |r| with 1e-8 |r0|: rounding may move a count by one (then the stage differs by ~1e-8)
    assert len(iters) == len(oiters) and max(abs(a - b) for a, b in zip(iters, oiters)) <= 1, (iters, oiters)
    assert rel_err(sol.numpy(), uo) <= (1e-10 if iters == oiters else 2e-7)


@pytest.mark.parametrize("dim,reps", [(2, [14, 13]), (3, [12, 12, 13])])
def test_wave_rk4_block_system(lib, dim, reps):
    """BASELINE config 4 (applications/wave wave-rk, uncut): u_tt = Laplace u as the first-order block system
    [u; v]' = [v; M^-1(-K u)] (applications/wave/include/gdm/wave/problem.h:294-320), RK4 over TWO blocks
    (problem.h:330-345), Jacobi-CG mass solves; solution and per-stage CG counts against the oracle."""
    import gdm_b200 as g
    p = 3
    gs, gc, os_, oc = make_pair(dim, p, 1, reps, "dirichlet", hi=[1.0] * dim)
    M = make_operator(gs, gc, "mass")
    K = make_operator(gs, gc, "stiffness")
    Mo, Ko = oracle_operator(os_, oc, "mass"), oracle_operator(os_, oc, "stiffness")
    u0 = lambda pts, c: np.prod(np.sin(np.pi * pts), axis=1)
    uh = O.interpolate(os_, u0)
    oc.set_zero(uh)
    n = len(uh)
    u, v = g.Vector(gs, uh), g.Vector(gs)
    pre = g.PreconditionJacobi()
    pre.initialize(M)
    rhs = g.Vector(gs)
    iters, oiters = [], []

    def f(t, y, out):
        out[0].equ(y[1])                    # u' = v
        K.vmult(rhs, y[0])
        rhs.scale(-1.0)
        gc.set_zero(rhs)
        out[1].set(0.0)
        ctl = g.ReductionControl(100, 1e-10, 1e-8)
        g.SolverCG(ctl).solve(M, out[1], rhs, pre)   # v' = M^-1 (-K u)
        iters.append(ctl.last_step())

    def fo(t, y):
        uu, vv = y[:n], y[n:]
        r = -(Ko @ uu)
        oc.set_zero(r)
        ctl = O.ReductionControl(100, 1e-10, 1e-8)
        dv = O.solver_cg(Mo, np.zeros(n), r, O.PreconditionJacobi(Mo), ctl)
        oiters.append(ctl.last_step())
        return np.concatenate([vv, dv])

    rk, rko = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER), O.ExplicitRungeKutta4()
    dt = 0.3 / max(reps)
    yo, t = np.concatenate([uh, np.zeros(n)]), 0.0
    for step in range(4):
        rk.evolve_one_time_step(f, t, dt, [u, v])
        t, yo = rko.evolve_one_time_step(fo, t, dt, yo)
    assert len(iters) == len(oiters) and max(abs(a - b) for a, b in zip(iters, oiters)) <= 1, (iters, oiters)
    tol = 1e-10 if iters == oiters else 2e-7
    assert rel_err(u.numpy(), yo[:n]) <= tol
    assert np.abs(v.numpy() - yo[n:]).max() <= 10 * tol * max(np.abs(yo[n:]).max(), 1e-300)
    # energy 1/2 (v.Mv + u.Ku) is conserved by the exact flow; RK4 keeps it to O(dt^4) per step
    E0 = 0.5 * float(uh @ (Ko @ uh))
    E1 = 0.5 * float(yo[n:] @ (Mo @ yo[n:]) + yo[:n] @ (Ko @ yo[:n]))
    assert abs(E1 - E0) <= 1e-3 * E0


@pytest.mark.parametrize("dim,p,reps,bc,nc", [(3, 3, [14, 13, 15], "dirichlet", 1), (3, 5, [13, 12, 14], "periodic", 1),
                                              (3, 3, [12, 13, 11], "mixed", 1), (3, 1, [9, 8, 7], "left", 1),
                                              (2, 3, [15, 14], "none", 2), (2, 5, [17, 16], "periodic", 1), (1, 3, [20], "dirichlet", 1),
                                              (3, 3, [24, 22, 20], "dirichlet", 1)])
def test_kronecker_direct_mass_inverse(lib, dim, p, reps, bc, nc):
    """gdm_operator_mass_inverse: banded line solves per direction against a sparse direct solve of the
    oracle's assembled mass matrix; replaces the CG mass solve of the RK stages
    (applications/advection/include/gdm/advection/problem.h:236-267)."""
    import gdm_b200 as g
    import scipy.sparse.linalg as spla
    gs, gc, os_, oc = make_pair(dim, p, nc, reps, bc)
    scale = 0.75
    M = make_operator(gs, gc, "mass", scale=scale)
    Mo = oracle_operator(os_, oc, "mass", scale=scale)
    bh = np.random.default_rng(5).uniform(-1, 1, gs.n_dofs())
    ref = spla.spsolve(Mo.tocsc(), bh)
    b, x = g.Vector(gs, bh), g.Vector(gs)
    M.mass_inverse(x, b)
    assert rel_err(x.numpy(), ref) <= 1e-11
    assert np.array_equal(b.numpy(), bh)
    y = g.Vector(gs)
    M.vmult(y, x)
    assert rel_err(y.numpy(), bh) <= 1e-11
    M.mass_inverse(b, b)  # in place
    assert rel_err(b.numpy(), ref) <= 1e-11


def test_wave_rk4_with_direct_mass_inverse(lib):
    """The wave block system of test_wave_rk4_block_system with the CG mass solve replaced by the direct inverse: same
    trajectory as the oracle's RK4 with exact (sparse direct) mass solves."""
    import gdm_b200 as g
    import scipy.sparse.linalg as spla
    n, p = 12, 3
    gs, gc, os_, oc = make_pair(3, p, 1, [n, n, n], "dirichlet", hi=[1.0, 1.0, 1.0])
    M, K = make_operator(gs, gc, "mass"), make_operator(gs, gc, "stiffness")
    Mo, Ko = oracle_operator(os_, oc, "mass"), oracle_operator(os_, oc, "stiffness")
    lu = spla.splu(Mo.tocsc())
    con = oc.constrained_mask(os_.n_dofs())
    u0 = lambda pts, c: np.sin(np.pi * pts[:, 0]) * np.sin(np.pi * pts[:, 1]) * np.sin(np.pi * pts[:, 2])
    uh = O.interpolate(os_, u0)
    uh[con] = 0.0
    u, v = g.Vector(gs, uh), g.Vector(gs)
    tmp = g.Vector(gs)

    def f(t, y, out):
        out[0].equ(y[1])
        K.vmult(tmp, y[0])
        tmp.scale(-1.0)
        gc.set_zero(tmp)
        M.mass_inverse(out[1], tmp)

    def fo(t, y):
        r = -(Ko @ y[0])
        r[con] = 0.0
        return [y[1].copy(), lu.solve(r)]

    rk = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER)
    dt, t = 0.3 / n, 0.0
    yo = [uh.copy(), np.zeros_like(uh)]
    b = [1 / 6, 1 / 3, 1 / 3, 1 / 6]
    for step in range(4):
        t = rk.evolve_one_time_step(f, t, dt, [u, v])
        # classical RK4 on the oracle side
        k1 = fo(0, yo)
        k2 = fo(0, [yo[i] + 0.5 * dt * k1[i] for i in range(2)])
        k3 = fo(0, [yo[i] + 0.5 * dt * k2[i] for i in range(2)])
        k4 = fo(0, [yo[i] + dt * k3[i] for i in range(2)])
        yo = [yo[i] + dt * (b[0] * k1[i] + b[1] * k2[i] + b[2] * k3[i] + b[3] * k4[i]) for i in range(2)]
    assert rel_err(u.numpy(), yo[0]) <= 1e-10
    assert np.abs(v.numpy() - yo[1]).max() <= 1e-10 * max(np.abs(yo[1]).max(), 1.0)
