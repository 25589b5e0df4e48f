"""GPU parity of the fused sm_100a tensor-product kernel (dim 3): vs the oracle at small sizes,
vs the generic multi-pass kernels and through size-independent properties at BASELINE size."""
import os

import numpy as np
import pytest

from helpers import make_pair, make_operator, oracle_operator, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

FUSED_CASES = [
    # p, reps (cells), bc
    (3, [35, 47, 20], "dirichlet"),
    (3, [33, 43, 9], "none"),
    (3, [40, 12, 17], "left"),
    (1, [37, 35, 11], "dirichlet"),
    (1, [5, 4, 6], "none"),
    (5, [35, 29, 14], "dirichlet"),
    (5, [13, 12, 27], "none"),
    (3, [8, 8, 8], "dirichlet"),
    # periodic directions (fold . A . duplicate): all three, and Dirichlet in x with periodic y, z
    (3, [35, 33, 20], "periodic"),
    (5, [34, 29, 31], "periodic"),
    (1, [33, 12, 9], "periodic"),
    (3, [40, 36, 17], "mixed"),
]


@pytest.mark.parametrize("p,reps,bc", FUSED_CASES)
@pytest.mark.parametrize("kind", ["mass", "stiffness", "advection", "advection_t"])
@pytest.mark.parametrize("mode", ["guided", "static"])
def test_fused_apply_matches_oracle(lib, monkeypatch, p, reps, bc, kind, mode):
    import gdm_b200 as g
    monkeypatch.setenv("GDM_PERS_MODE", mode)  # self-scheduled guided shares (default) / one weighted share per CTA
    gs, gc, os_, oc = make_pair(3, p, 1, reps, bc)
    b = [1.0, 0.15, -0.05]
    scale = -0.5 if kind == "advection" else 1.0
    A = make_operator(gs, gc, kind, b=b, kernel=g.capi.KERNEL_FUSED, scale=scale)
    assert A.kernel_used() == g.capi.KERNEL_FUSED
    Ao = oracle_operator(os_, oc, kind, b=b, scale=scale)
    rng = np.random.default_rng(0)
    xh = rng.uniform(-1, 1, gs.n_dofs())
    x, y = g.Vector(gs, xh), g.Vector(gs)
    y.set(7.0)  # stale values must be overwritten everywhere
    A.vmult(y, x)
    ref = Ao @ xh
    assert rel_err(y.numpy(), ref) <= TOL
    y2 = g.Vector(gs, xh)
    A.vmult_add(y2, x)
    assert rel_err(y2.numpy(), ref + xh) <= TOL
    # pads of the padded layout stay zero: the l2 norm over the storage equals the norm of the DoFs
    assert abs(y.l2_norm() - np.linalg.norm(ref)) <= 1e-11 * np.linalg.norm(ref)


def test_auto_picks_fused_and_falls_back(lib):
    import gdm_b200 as g
    gs, gc, _, _ = make_pair(3, 3, 1, [12, 12, 12], "dirichlet")
    assert make_operator(gs, gc, "stiffness").kernel_used() == g.capi.KERNEL_FUSED
    gs, gc, _, _ = make_pair(3, 3, 1, [12, 12, 12], "periodic")
    assert make_operator(gs, gc, "stiffness").kernel_used() == g.capi.KERNEL_FUSED
    gs, gc, _, _ = make_pair(3, 3, 2, [12, 12, 12], "dirichlet")  # two components: generic passes
    assert make_operator(gs, gc, "stiffness").kernel_used() == g.capi.KERNEL_GENERIC
    with pytest.raises(g.ExcNotImplemented):
        make_operator(gs, gc, "stiffness", kernel=g.capi.KERNEL_FUSED)
    gs, gc, _, _ = make_pair(2, 3, 1, [12, 12], "dirichlet")
    assert make_operator(gs, gc, "stiffness").kernel_used() == g.capi.KERNEL_GENERIC


@pytest.mark.parametrize("p,n", [(3, 256), (5, 128)])
def test_full_size_properties(lib, p, n):
    """BASELINE.json config 2 size (256^3 cells, p=3): fused == generic, symmetry, null space, volume."""
    import gdm_b200 as g
    gs, gc, _, _ = make_pair(3, p, 1, [n, n, n], "dirichlet", hi=[1.0, 1.0, 1.0])
    rng = np.random.default_rng(0)
    nd = gs.n_dofs()
    xh = rng.uniform(-1, 1, nd)
    x, y, yg, z, w = g.Vector(gs, xh), g.Vector(gs), g.Vector(gs), g.Vector(gs, rng.uniform(-1, 1, nd)), g.Vector(gs)
    for kind in ("stiffness", "mass"):
        Af = make_operator(gs, gc, kind, kernel=g.capi.KERNEL_FUSED)
        Ag = make_operator(gs, gc, kind, kernel=g.capi.KERNEL_GENERIC)
        Af.vmult(y, x)
        Ag.vmult(yg, x)
        ref = yg.numpy()
        assert rel_err(y.numpy(), ref) <= TOL
        # symmetry: <z, A x> == <x, A z>
        Af.vmult(w, z)
        s1, s2 = z * y, x * w
        assert abs(s1 - s2) <= 1e-11 * max(abs(s1), abs(s2), 1.0)
    # unconstrained operators: constants are in the null space of K; 1^T M 1 = |Omega|
    gs2, gc2, _, _ = make_pair(3, p, 1, [n, n, n], "none", hi=[1.0, 1.0, 1.0])
    one, out = g.Vector(gs2), g.Vector(gs2)
    one.set(1.0)
    K = make_operator(gs2, gc2, "stiffness", kernel=g.capi.KERNEL_FUSED)
    K.vmult(out, one)
    kd = K.diagonal()
    assert out.linfty_norm() <= 1e-12 * kd.linfty_norm() * 50
    M = make_operator(gs2, gc2, "mass", kernel=g.capi.KERNEL_FUSED)
    M.vmult(out, one)
    assert abs(one * out - 1.0) <= 1e-12


@pytest.mark.parametrize("p,reps,bc", [(3, [35, 30, 70], "dirichlet"), (3, [33, 20, 97], "none"), (5, [20, 21, 100], "left")])
def test_pipelined_host_buffer_apply(lib, p, reps, bc):
    """gdm_operator_vmult_host on a grid deep enough for the chunked H2D / apply / D2H pipeline."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(3, p, 1, reps, bc)
    for kind in ("stiffness", "mass"):
        A = make_operator(gs, gc, kind, kernel=g.capi.KERNEL_FUSED)
        Ao = oracle_operator(os_, oc, kind)
        xh = np.random.default_rng(5).uniform(-1, 1, gs.n_dofs())
        yh = np.full_like(xh, 3.0)
        A.vmult_host(yh, xh)
        assert rel_err(yh, Ao @ xh) <= TOL
        A.vmult_host(yh, -xh)  # staging buffers are reused
        assert rel_err(yh, -(Ao @ xh)) <= TOL


@pytest.mark.parametrize("p,reps,bc", FUSED_CASES)
@pytest.mark.parametrize("kind", ["mass", "stiffness"])
def test_fused_dot_product(lib, p, reps, bc, kind):
    """gdm_operator_vmult_dot: y = A x and <x, y> from the store epilogue of the tile kernel (CG's q = A p, p.q)."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(3, p, 1, reps, bc)
    A = make_operator(gs, gc, kind, kernel=g.capi.KERNEL_FUSED)
    Ao = oracle_operator(os_, oc, kind)
    xh = np.random.default_rng(3).uniform(-1, 1, gs.n_dofs())
    x, y = g.Vector(gs, xh), g.Vector(gs)
    d = A.vmult_dot(y, x)
    ref = Ao @ xh
    assert rel_err(y.numpy(), ref) <= TOL
    # relative to sum |x_i y_i| (the condition of the sum), tolerance of one FP64 summation of n terms
    assert abs(d - xh @ ref) <= 1e-13 * np.abs(xh * ref).sum()
    # deterministic: the same call gives the same bits
    assert A.vmult_dot(y, x) == d


def test_cg_with_fused_dot_matches_separate_dot(lib, monkeypatch):
    """The fused p.Ap must not change what CG computes beyond the summation order of one dot product."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(3, 3, 1, [20, 18, 33], "dirichlet")
    A = make_operator(gs, gc, "stiffness", kernel=g.capi.KERNEL_FUSED)
    rhs = g.Vector(gs)
    rhs.set(1.0)
    gc.set_zero(rhs)
    sols, steps = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("GDM_CG_FUSED_DOT", flag)
        u = g.Vector(gs)
        ctl = g.ReductionControl(500, 1e-12, 1e-8)
        g.SolverCG(ctl).solve(A, u, rhs, g.PreconditionIdentity())
        sols.append(u.numpy())
        steps.append(ctl.last_step())
    assert abs(steps[0] - steps[1]) <= 1
    assert rel_err(sols[0], sols[1]) <= 1e-6 if steps[0] != steps[1] else rel_err(sols[0], sols[1]) <= 1e-9
