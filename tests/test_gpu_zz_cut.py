"""(Named to run last: two of the three cases could not be run on a GPU before the round's GPU budget ended.)
BASELINE configuration 5 in miniature on the GPU: the reference's cut Poisson prototype end to end -- cut-cell set-up
on the host (`gdm_cut_*`), tensor-product stiffness apply with the cut / ghost-penalty rows attached as CSR, CG."""
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
