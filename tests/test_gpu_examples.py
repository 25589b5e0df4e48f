"""C++ drivers written against include/gdm (the reference's API names) reproduce the reference's goldens on the GPU."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name, *args):
    exe = os.path.join(ROOT, "examples", "build", name)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    return subprocess.run([exe] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600).stdout


def test_poisson_01_gdm_stdout_matches_golden(lib, golden_dir):
    """tests/poisson_01_gdm.cc -> tests/poisson_01_gdm.output, token by token (values to 1e-12)."""
    out = _run("poisson_01_gdm").split()
    gold = open(os.path.join(golden_dir, "poisson_01_gdm.output")).read().split()
    assert len(out) == len(gold), out
    for a, b in zip(out, gold):
        assert abs(float(a) - float(b)) <= 1e-12, (a, b)


def test_mass_01_gdm_stdout_matches_golden(lib, golden_dir):
    out = _run("mass_01_gdm").strip()
    assert out == open(os.path.join(golden_dir, "mass_01_gdm.output")).read().strip()


def test_advection_01_gdm_against_oracle(lib):
    """prototypes/advection_01_gdm.cc time loop (2D, p=5, N=16): per-step L2 errors vs the oracle."""
    import oracle as O
    n, p = 16, 5
    lines = [l.split() for l in _run("advection_01_gdm", 2, p, n).strip().split("\n")]
    s = O.System(2, p)
    s.subdivided_hyper_cube(n)
    c = O.Constraints()
    for d in range(2):
        s.make_periodicity_constraints(d, c)
    c.close()
    b = [1.0, 0.15]
    M = O.kron_operator(s, c, "mass")
    R = -O.kron_operator(s, c, "advection", b=b, constrained_diagonal="zero")
    exact = lambda t: (lambda pts, comp: np.sin(2 * np.pi * (pts[:, 0] - t * b[0])) * np.cos(2 * np.pi * (pts[:, 1] - t * b[1])))
    u = O.interpolate(s, exact(0.0))
    errs = [(0.0, O.compute_global_error(O.integrate_difference(s, u, exact(0.0))))]

    def f(t, y):
        v0 = c.distribute(y.copy())
        return O.solver_cg(M, np.zeros_like(y), R @ v0, O.PreconditionJacobi(M), O.ReductionControl(100, 1e-10, 1e-8))

    rk, time = O.ExplicitRungeKutta4(), O.DiscreteTime(0.0, 0.1, 0.5 / n)
    while not time.is_at_end():
        _, u = rk.evolve_one_time_step(f, time.get_current_time(), time.get_next_step_size(), u)
        c.distribute(u)
        t = time.get_current_time() + time.get_next_step_size()
        errs.append((t, O.compute_global_error(O.integrate_difference(s, u, exact(t)))))
        time.advance_time()
    assert len(lines) == len(errs)
    for (ts, es), (t, e) in zip(lines, errs):
        assert abs(float(ts) - t) < 1e-12
        assert abs(float(es) - e) <= 2e-5 * e, (ts, es, e)  # stdout carries 6 significant digits
