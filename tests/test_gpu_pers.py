"""GPU parity of the persistent ramp-free fused kernel (kron3d_pers.cu) where its seams are exercised:
grids long in z so that tile columns are cut into several jobs (partial planes handed down through the scratch
slots), few slots so that a share runs several jobs, aligned and swept partitions; and oracle parity at the
BASELINE size (257^3 DoFs, p=3) and at 129^3 for p=5 through the matrix-free oracle apply (oracle/kron_apply.py)."""
import numpy as np
import pytest

import oracle as O
from helpers import make_pair, make_operator, oracle_operator, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

SEAM_CASES = [
    # p, reps (cells), bc, slots, aligned
    (3, [35, 33, 70], "dirichlet", None, "1"),
    (3, [35, 33, 70], "dirichlet", None, "0"),
    (3, [40, 36, 61], "none", "7", "1"),
    (3, [40, 36, 61], "left", "13", "0"),
    (1, [37, 35, 50], "dirichlet", "5", "0"),
    (1, [33, 34, 40], "none", None, "1"),
    (5, [35, 29, 90], "dirichlet", None, "1"),
    (5, [13, 40, 75], "none", "11", "0"),
    (3, [35, 33, 70], "periodic", None, "1"),
    (5, [20, 33, 64], "periodic", "9", "0"),
]


@pytest.mark.parametrize("p,reps,bc,slots,aligned", SEAM_CASES)
@pytest.mark.parametrize("kind", ["mass", "stiffness", "advection", "advection_t"])
def test_seams_match_oracle(lib, monkeypatch, p, reps, bc, slots, aligned, kind):
    import gdm_b200 as g
    if slots is not None:
        monkeypatch.setenv("GDM_PERS_SLOTS", slots)
    monkeypatch.setenv("GDM_PERS_ALIGNED", aligned)
    gs, gc, os_, oc = make_pair(3, p, 1, reps, bc)
    b = [1.0, 0.15, -0.05]
    scale = -0.5 if kind == "advection" else 1.0
    A = make_operator(gs, gc, kind, b=b, kernel=g.capi.KERNEL_FUSED, scale=scale)
    assert A.kernel_used() == g.capi.KERNEL_FUSED
    Ao = oracle_operator(os_, oc, kind, b=b, scale=scale)
    xh = np.random.default_rng(0).uniform(-1, 1, gs.n_dofs())
    x, y = g.Vector(gs, xh), g.Vector(gs)
    ref = Ao @ xh
    for rep in range(3):  # flags and tickets are reused across launches (epoch / ticket base)
        y.set(7.0)
        A.vmult(y, x)
        assert rel_err(y.numpy(), ref) <= TOL
    y2 = g.Vector(gs, xh)
    A.vmult_add(y2, x)
    assert rel_err(y2.numpy(), ref + xh) <= TOL
    assert np.array_equal(x.numpy(), xh)  # periodic directions patch the input in place and must restore it
    # fused dot product <x, A x> (CG: p . A p) and bitwise reproducibility of the apply
    y3 = g.Vector(gs)
    d = A.vmult_dot(y3, x)
    assert np.array_equal(y3.numpy(), y.numpy())
    assert abs(d - xh @ ref) <= 1e-11 * np.abs(xh * ref).sum()


@pytest.mark.parametrize("p,n,kind", [(3, 256, "stiffness"), (3, 256, "mass"), (5, 128, "stiffness"), (1, 200, "stiffness"),
                                      (3, 256, "advection")])
@pytest.mark.parametrize("kernel", ["fused", "generic"])
def test_baseline_size_matches_oracle(lib, p, n, kind, kernel):
    """BASELINE.json config 2 (256^3 cells, p=3): single apply vs the oracle, relative 1e-12 (north_star tolerance)."""
    import gdm_b200 as g
    gs, gc, os_, oc = make_pair(3, p, 1, [n, n, n], "dirichlet", hi=[1.0, 1.0, 1.0])
    b = [1.0, 0.15, -0.05]
    A = make_operator(gs, gc, kind, b=b, kernel=g.capi.KERNEL_FUSED if kernel == "fused" else g.capi.KERNEL_GENERIC)
    assert A.kernel_used() == (g.capi.KERNEL_FUSED if kernel == "fused" else g.capi.KERNEL_GENERIC)
    Ko = O.KronApply(os_, oc, kind, b=b, constrained_diagonal="zero" if kind == "advection" else "assembled")
    xh = np.random.default_rng(0).uniform(-1, 1, gs.n_dofs())
    x, y = g.Vector(gs, xh), g.Vector(gs)
    A.vmult(y, x)
    ref = Ko @ xh
    assert rel_err(y.numpy(), ref) <= TOL
