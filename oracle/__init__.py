"""CPU oracle for the GDM hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of what the reference
(peterrum/dealii-galerkin-difference-methods on deal.II) computes on the
path `operator apply + CG + explicit RK`.  It exists to check the CUDA path.

Rules:
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` leg may import it;
  * the product (`gdm_b200`, `include/`, `csrc/`) never imports, links or
    executes anything in here and has no CPU fallback.

Pinning: the restatement is checked against the reference's committed golden
outputs (`tests/golden/*.output`, copied by `tests/golden/make_golden.py`):
poly_01, fe_02_gdm, poisson_01_gdm, poisson_02_gdm (1 and 3 ranks), mass_01_gdm,
mass_02_gdm, elasticity_01_gdm, prototypes/cut_poisson_01_gdm (cut.py), applications/wave/tests/{wave_0, heat_0, heat_1,
wave_composite_0, heat_composite_0, wave_1 (2D, cut_q.py), step85_0 (2D, 5 digits)} (wave_app.py, every printed step).  3D has no reference golden (`tests/fe_01_gdm.output` is missing
upstream): 3D parity is pinned only through the dimension-generic cell-loop
restatement validated in 1D/2D.
"""
from .basis import (lagrange_nodes, lagrange_coefficients, generate_polynomials_1D,
                    basis_values, gauss_legendre_01)
from .system import System, Constraints
from .assemble import (assemble_cell_loop, kron_operator, kron_unconstrained, matrices_1d,
                       rhs_cell_loop, advection_residual_cell_loop)
from .solvers import (ReductionControl, SolverControlNoConvergence, solver_cg,
                      PreconditionIdentity, PreconditionJacobi, DiagonalMatrix,
                      ExplicitRungeKutta4, DiscreteTime)
from .kron_apply import KronApply, kron_apply, constraint_matrices_1d
from .vector_tools import interpolate, integrate_difference, compute_global_error
from .mass_inverse import kron_mass_solve, bordered_solve_1d, free_matrix_1d
from . import cut, wave_app

__all__ = [n for n in dir() if not n.startswith("_")]
