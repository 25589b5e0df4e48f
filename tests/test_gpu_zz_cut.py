"""Cut-cell problems on the GPU: cut-cell set-up on the host (`gdm_cut_*`), tensor-product apply with the cut /
ghost-penalty rows attached as CSR, CG and Runge-Kutta around it -- the reference's cut Poisson prototype (BASELINE
configuration 5 in miniature) and the explicit runs of applications/wave.

Named to run last: only `test_cut_poisson_01_gdm[False]` ran on a B200 before the round's GPU budget ended
(profiles/r2/session_ao_cut_pytest.log); the other cases were checked on the CPU piece by piece (tests/test_cut_cell.py
covers their host side against the goldens) and their logic against a numpy stand-in for the GPU classes."""
import numpy as np
import pytest

import oracle as O
from oracle import cut
from helpers import make_pair, make_operator, rel_err
from test_cut_cell import golden_errors, overlay_matrix, exact_solution

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ghost_penalty", [False, True])
def test_cut_poisson_01_gdm(lib, golden_dir, ghost_penalty):
    """prototypes/cut_poisson_01_gdm.cc (2D, 64^2 cells, p = 3, unit circle, CG(n_dofs, 1e-10, 1e-6), identity):
    the L2 error of the golden output, the oracle's CG iteration count on the same matrix."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(2, 3, 1, [64, 64], "none", lo=[-1.21] * 2, hi=[1.21] * 2)
    ls = cut.interpolate_level_set(os_, cut.sphere_level_set([0.0, 0.0], 1.0))
    c = g.CutPoisson(2, 3, [64, 64], [-1.21] * 2, [1.21] * 2, ls, ghost_penalty=ghost_penalty)
    A = make_operator(gs, gc, "stiffness")
    rows = c.rows()
    A.attach_csr(*rows)
    n = gs.n_dofs()
    b, u = g.Vector(gs, c.rhs()), g.Vector(gs)
    ctl = g.ReductionControl(n, 1e-10, 1e-6)
    g.SolverCG(ctl).solve(A, u, b, g.PreconditionIdentity())
    err = c.l2_error_inside(u.numpy(), lambda pt, comp: 1.0 - (pt[0] ** 2 + pt[1] ** 2 - 1.0))
    print(f"cut_poisson_01 gp={ghost_penalty}: L2 {err:.6e}, {ctl.last_step()} iterations")
    assert abs(err - golden_errors(golden_dir)[1 if ghost_penalty else 0]) <= 1.5e-8
    Am = overlay_matrix(os_, *rows)
    octl = O.ReductionControl(n, 1e-10, 1e-6)
    uo = O.solver_cg(Am, np.zeros(n), c.rhs(), O.PreconditionIdentity(), octl)
    print(f"  oracle CG: {octl.last_step()} iterations, solution difference {rel_err(u.numpy(), uo):.2e}")
    if ghost_penalty:  # without it the oracle's own count depends on the host's rounding (see below)
        assert abs(ctl.last_step() - octl.last_step()) <= 5, (ctl.last_step(), octl.last_step())
    # Measured on B200 (gpurun_out/ao_pytest.log): without ghost penalty 593 iterations, L2 4.230229e-04, oracle 591.
    # That variant is ill conditioned (small cut cells): perturbing the right-hand side by 1e-15 moves the oracle's own
    # count between 591 and 664 and the error between 4.2303e-04 and 4.264e-04, so its golden is pinned by rounding; the
    # GPU sums are evaluated in a fixed order, the run is reproducible.  With ghost penalty the count moves by +-2 only.
    if ghost_penalty:
        assert rel_err(u.numpy(), uo) <= 5e-3  # both stop at a residual reduction of 1e-6, at their own iteration


def test_cut_poisson_3d_fused(lib):
    """The 3D form (unit sphere in [-1.21, 1.21]^3, p = 3, ghost penalty): single apply of the fused tile kernel with
    the attached rows against the oracle matrix, Jacobi-CG against the oracle's CG, the inside L2 error."""
    import gdm_b200 as g
    n1 = 24
    gs, gc, os_, oc = make_pair(3, 3, 1, [n1] * 3, "none", lo=[-1.21] * 3, hi=[1.21] * 3)
    ls = cut.interpolate_level_set(os_, cut.sphere_level_set([0.0] * 3, 1.0))
    c = g.CutPoisson(3, 3, [n1] * 3, [-1.21] * 3, [1.21] * 3, ls, ghost_penalty=True)
    A = make_operator(gs, gc, "stiffness", kernel=g.capi.KERNEL_FUSED)
    rows = c.rows()
    A.attach_csr(*rows)
    Am = overlay_matrix(os_, *rows)
    n = gs.n_dofs()
    xh = np.random.default_rng(2).uniform(-1, 1, n)
    x, y = g.Vector(gs, xh), g.Vector(gs)
    A.vmult(y, x)
    assert rel_err(y.numpy(), Am @ xh) <= 1e-12
    b, u = g.Vector(gs, c.rhs()), g.Vector(gs)
    ctl = g.ReductionControl(n, 1e-10, 1e-8)
    P = g.PreconditionJacobi()
    P.initialize(A)
    g.SolverCG(ctl).solve(A, u, b, P)
    octl = O.ReductionControl(n, 1e-10, 1e-8)
    uo = O.solver_cg(Am, np.zeros(n), c.rhs(), O.PreconditionJacobi(Am), octl)
    print(f"cut 3D: apply {rel_err(y.numpy(), Am @ xh):.2e}, CG {ctl.last_step()} / oracle {octl.last_step()}, "
          f"solution difference {rel_err(u.numpy(), uo):.2e}")
    assert abs(ctl.last_step() - octl.last_step()) <= 4, (ctl.last_step(), octl.last_step())
    assert rel_err(u.numpy(), uo) <= 5e-3  # rounding moves the stopping iteration by +-2 and the iterate by 7e-4
    err = c.l2_error_inside(u.numpy(), lambda pt, comp: 1.0 - 2.0 / 3.0 * (pt[0] ** 2 + pt[1] ** 2 + pt[2] ** 2 - 1.0))
    erro = c.l2_error_inside(uo, lambda pt, comp: 1.0 - 2.0 / 3.0 * (pt[0] ** 2 + pt[1] ** 2 + pt[2] ** 2 - 1.0))
    assert abs(err - erro) <= 1e-4 * erro and err < 1e-2


def _gpu_wave_operators(g, params):
    """Mass and stiffness operators with the cut rows of the wave application attached (wave/mass.h:47-249,
    wave/stiffness.h:42-407), the host-side generators and the oracle system."""
    from test_cut_cell import product_wave_operators
    dim, n1 = params["dim"], params["n_subdivisions"]
    gs, gc, os_, oc = make_pair(dim, params["fe_degree"], 1, [n1] * dim, "none",
                                lo=[params["left"]] * dim, hi=[params["right"]] * dim)
    s, ls, cm, ca, Mo, Ao = product_wave_operators(params)
    M, A = make_operator(gs, gc, "mass"), make_operator(gs, gc, "stiffness")
    M.attach_csr(*cm.rows())
    A.attach_csr(*ca.rows())
    return gs, os_, cm, ca, M, A, Mo, Ao


def test_wave_app_wave_0(lib, golden_dir):
    """applications/wave, simulation "wave" in 1D (problem.h:280-345) on the GPU: RK4 over [u; v], Jacobi-CG mass
    solves with the reference's control (1000, 1e-20, 1e-14), cut rows attached to both operators; the L2 column of
    applications/wave/tests/wave_0.output for the first 25 printed steps."""
    import gdm_b200 as g
    from oracle import wave_app
    from test_cut_cell import _app_golden
    params = wave_app.wave_preset(1)
    gs, os_, cm, ca, M, A, Mo, Ao = _gpu_wave_operators(g, params)
    k = 1.5 * np.pi
    exact = lambda t: (lambda pt, c: np.cos(k * abs(pt[0])) * np.cos(k * t))
    n = gs.n_dofs()
    u, v = g.Vector(gs, O.interpolate(os_, lambda pts, c: params["exact"](pts, 0.0))), g.Vector(gs)
    pre = g.PreconditionJacobi()
    pre.initialize(M)
    rhs, load = g.Vector(gs), g.Vector(gs)
    iters = []

    def f(t, y, out):
        out[0].equ(y[1])                                  # du/dt = v
        A.vmult(rhs, y[0])
        rhs.scale(-1.0)
        load.upload(ca.load_vector(None, exact(t)))       # <gamma_D/h v - dv/dn, g(t)> (zero here: g(+-1, t) = 0)
        rhs.add(1.0, load)
        out[1].set(0.0)
        ctl = g.ReductionControl(1000, 1e-20, 1e-14)
        g.SolverCG(ctl).solve(M, out[1], rhs, pre)        # dv/dt = M^-1 rhs(u, t)
        iters.append(ctl.last_step())

    gold = _app_golden(golden_dir, "app_wave_wave_0.output")
    rk = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER)
    t, dt = 0.0, 0.3 * 2.42 / 40
    for step in range(25):
        e = cm.l2_error_inside(u.numpy(), exact(t))
        assert abs(e - gold[step][2]) <= 2e-8 * gold[step][2], (step, e, gold[step], iters[-4:])
        t = rk.evolve_one_time_step(f, t, dt, [u, v])
    print(f"wave_0 on the GPU: 25 steps, mass CG iterations {min(iters)}..{max(iters)}")


def test_cut_heat_2d_rk4(lib):
    """The 2D form of the heat-rk run (problem.h:72-127; x^9 y^8 exp(-t), wave-app.cc:95-121) with a Q1 level set: three
    RK4 steps on the GPU (Jacobi-CG mass solves to 1e-12, time-dependent volume and surface loads from gdm_cut_load_vector)
    against the oracle's run with an exact mass solve."""
    import gdm_b200 as g
    import scipy.sparse.linalg as sla
    from oracle import wave_app
    params = dict(wave_app.heat_preset(1), dim=2, n_subdivisions=24)
    gs, os_, cm, ca, M, A, Mo, Ao = _gpu_wave_operators(g, params)
    ex = lambda t: (lambda pt, c: pt[0] ** 9 * pt[1] ** 8 * np.exp(-t))
    src = lambda t: (lambda pt, c: -pt[0] ** 7 * pt[1] ** 6 * np.exp(-t) * (pt[0] ** 2 * pt[1] ** 2 + 72 * pt[1] ** 2 + 56 * pt[0] ** 2))
    n = gs.n_dofs()
    u0 = O.interpolate(os_, lambda pts, c: pts[:, 0] ** 9 * pts[:, 1] ** 8)
    u = g.Vector(gs, u0)
    pre = g.PreconditionJacobi()
    pre.initialize(M)
    rhs, load = g.Vector(gs), g.Vector(gs)

    def f(t, y, out):
        A.vmult(rhs, y)
        rhs.scale(-1.0)
        load.upload(ca.load_vector(src(t), ex(t)))
        rhs.add(1.0, load)
        out.set(0.0)
        g.SolverCG(g.ReductionControl(2000, 1e-30, 1e-12)).solve(M, out, rhs, pre)

    solve = sla.factorized(Mo.tocsc())
    fo = lambda t, y: solve(-(Ao @ y) + ca.load_vector(src(t), ex(t)))
    rk, rko = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER), O.ExplicitRungeKutta4()
    dx = 2.42 / 24
    t, dt, yo = 0.0, params["cfl"] * dx ** 2, u0.copy()
    for step in range(3):
        rk.evolve_one_time_step(f, t, dt, u)
        t, yo = rko.evolve_one_time_step(fo, t, dt, yo)
    print(f"cut heat 2D: difference to the oracle run {rel_err(u.numpy(), yo):.2e}")
    assert rel_err(u.numpy(), yo) <= 1e-9
    assert abs(cm.l2_error_inside(u.numpy(), ex(t)) - cm.l2_error_inside(yo, ex(t))) <= 1e-10


def test_cut_poisson_01_gdm_cpp_driver(lib, golden_dir):
    """examples/cut_poisson_01_gdm.cc (the reference's prototype against include/gdm): same table as
    prototypes/cut_poisson_01_gdm.output; the errors are printed with 5 digits and may differ by one unit in the last."""
    import re
    from test_gpu_examples import _run
    out = _run("cut_poisson_01_gdm")
    got = [(float(h), float(e)) for h, e in re.findall(r"^\s*([0-9.]+)\s+([0-9.e+-]+)\s*$", out, flags=re.M)]
    gold = golden_errors(golden_dir)
    assert len(got) == 2 and out.count("Mesh size  L2-Error") == 2, out
    for (h, e), ge in zip(got, gold):
        assert abs(h - 0.0378) < 1e-4 and abs(e - ge) <= 1.5e-8, (out, gold)


@pytest.mark.parametrize("dim,simulation,golden", [(1, "wave", "app_wave_wave_0.output"), (1, "heat-rk", "app_wave_heat_1.output"),
                                                   (2, "wave", "app_wave_wave_1.output")])
def test_wave_app_cpp_driver(lib, golden_dir, dim, simulation, golden):
    """examples/wave_app.cc (the explicit runs of applications/wave/wave-app.cc against include/gdm) on the GPU: every
    printed step of applications/wave/tests/{wave_0,heat_1,wave_1}.output (wave_1: 2D, level set of degree 3), all three
    error columns; the ` [L] solved in k`
    lines carry this library's Jacobi-CG counts instead of the reference's AMG / ILU counts and are not compared."""
    from test_gpu_examples import _run
    from test_cut_cell import _app_golden
    out = _run("wave_app", dim, simulation)
    rows = [l.split() for l in out.splitlines() if l.strip() and not l.startswith(" [L]")]
    gold = _app_golden(golden_dir, golden)
    assert len(rows) == len(gold), out[-2000:]
    assert out.count(" [L] solved in") == 4 * (len(gold) - 1)
    for r, g_ in zip(rows, gold):
        assert int(r[0]) == g_[0] and abs(float(r[1]) - g_[1]) <= 5.1e-6
        for i in (2, 3, 4):  # 2D: the Linf column depends on the height direction taken in the four diagonal cut cells
            assert abs(float(r[i]) - g_[i]) <= (2e-8 if dim == 1 or i < 4 else 1e-7) * g_[i], (r, g_)


def test_wave_app_wave_composite_0(lib, golden_dir):
    """applications/wave, simulation "wave-composite" in 1D (problem.h:347-437) on the GPU: two fields (inside / outside the
    unit sphere), RK4 over four blocks [u0; u1; v0; v1], each field with its own cut mass and stiffness rows (the outer
    one: negated level set + Nitsche on the box boundary), coupled on the surface through three CSR-only operators
    (scale 0 + attached rows: P, P^T, Q of wave/stiffness.h:441-574); both error tables of
    applications/wave/tests/wave_composite_0.output for the first 10 printed steps."""
    import gdm_b200 as g
    from oracle import wave_app
    from test_cut_cell import _app_golden
    prm = wave_app.wave_preset(1)
    gs, gc, os_, oc = make_pair(1, 3, 1, [40], "none", lo=[-1.21], hi=[1.21])
    ls = cut.interpolate_level_set(os_, cut.sphere_level_set([0.0], 1.0))
    n, gd, hmin = gs.n_dofs(), prm["nitsche_parameter"], 2.42 / 40
    box = ([40], [-1.21], [1.21])
    k = 1.5 * np.pi
    ex = lambda t: (lambda pt, c: np.cos(k * abs(pt[0])) * np.cos(k * t))
    fields = []
    for sign in (1.0, -1.0):
        cm = g.CutPoisson(1, 3, *box, sign * ls, ghost_parameter=prm["ghost_parameter_M"], gp_h_power=3, kind="mass", rhs_value=0.0)
        ca = g.CutPoisson(1, 3, *box, sign * ls, ghost_parameter=prm["ghost_parameter_A"], nitsche_parameter=gd, rhs_value=0.0,
                          boundary_value=0.0, outside_diagonal=0.0, surface_terms=False, domain_boundary_terms=True)
        M, A = make_operator(gs, gc, "mass"), make_operator(gs, gc, "stiffness")
        M.attach_csr(*cm.rows())
        A.attach_csr(*ca.rows())
        pre = g.PreconditionJacobi()
        pre.initialize(M)
        fields.append((cm, ca, M, A, pre))
    ci = g.CutPoisson(1, 3, *box, ls)
    coupling = {}
    for which in ("P", "PT", "Q"):
        op = make_operator(gs, gc, "stiffness", scale=0.0)  # rows outside the attached set stay zero
        op.attach_csr(*ci.coupling_rows(which))
        coupling[which] = op
    tau = 0.5 * gd / hmin
    u_init = O.interpolate(os_, lambda pts, c: prm["exact"](pts, 0.0))
    y = [g.Vector(gs, u_init), g.Vector(gs, u_init), g.Vector(gs), g.Vector(gs)]
    jump, total, c_sym, c_avg, c_pen, rhs, load = (g.Vector(gs) for _ in range(7))

    def f(t, yy, out):
        out[0].equ(yy[2])                                   # du/dt = v, both fields
        out[1].equ(yy[3])
        jump.equ(yy[0]).add(-1.0, yy[1])
        total.equ(yy[0]).add(1.0, yy[1])
        coupling["P"].vmult(c_sym, jump)                    # -1/2 P [u]
        coupling["PT"].vmult(c_avg, total)                  #  1/2 P^T (u0 + u1)
        coupling["Q"].vmult(c_pen, jump)                    #  tau / h Q [u]
        for kk, (cm, ca, M, A, pre) in enumerate(fields):
            A.vmult(rhs, yy[kk])
            rhs.scale(-1.0)
            load.upload(ca.boundary_load_vector(ex(t)))
            rhs.add(1.0, load)
            s = 1.0 if kk == 0 else -1.0                    # r0 -= c_sym - c_avg + c_pen, r1 -= c_sym + c_avg - c_pen
            rhs.add(0.5, c_sym).add(s * 0.5, c_avg).add(-s * tau, c_pen)
            out[2 + kk].set(0.0)
            g.SolverCG(g.ReductionControl(1000, 1e-20, 1e-14)).solve(M, out[2 + kk], rhs, pre)

    gold = _app_golden(golden_dir, "app_wave_wave_composite_0.output")
    rk = g.TimeStepping.ExplicitRungeKutta(g.TimeStepping.RK_CLASSIC_FOURTH_ORDER)
    t, dt = 0.0, 0.3 * hmin
    for step in range(10):
        for kk in range(2):
            e = fields[kk][0].error_norms_inside(y[kk].numpy(), ex(t))
            for i in range(3):
                assert abs(e[i] - gold[2 * step + kk][2 + i]) <= 2e-8 * gold[2 * step + kk][2 + i], (step, kk, e, gold[2 * step + kk])
        t = rk.evolve_one_time_step(f, t, dt, y)


def test_wave_app_cpp_driver_step85(lib, golden_dir):
    """`wave_app 2 step85` (wave-app.cc:13-61): 2D Poisson through the application with the assembled cut matrix (ghost
    penalty with h^3) and a level set of degree 3; the three errors of applications/wave/tests/step85_0.output are pure
    cut-cell quadrature error (the exact solution lies in the discrete space) and are reproduced to 5e-5 (host check with
    the oracle's CG: 2e-8, 3e-5, 1e-9)."""
    from test_gpu_examples import _run
    from test_cut_cell import _app_golden
    out = _run("wave_app", 2, "step85")
    rows = [l.split() for l in out.splitlines() if l.strip() and not l.startswith(" [L]")]
    g_ = _app_golden(golden_dir, "app_wave_step85_0.output")[0]
    assert len(rows) == 1 and out.count(" [L] solved in") == 1, out
    for i in (2, 3, 4):
        assert abs(float(rows[0][i]) - g_[i]) <= 5e-5 * g_[i], (rows, g_)
