"""deal.II solver semantics restated (oracle, test-only).

Not in the reference tree (deal.II): `SolverCG`, `ReductionControl`,
`PreconditionJacobi`, `TimeStepping::ExplicitRungeKutta(RK_CLASSIC_FOURTH_ORDER)`
and `DiscreteTime`.  Call sites in the reference: `tests/poisson_01_gdm.cc:164-170`,
`tests/mass_01_gdm.cc:124-131`, `prototypes/advection_01_gdm.cc:208-216,259-281`.
The loop structure, residual norm and iteration counting are pinned by the
goldens (poisson_01: 5 iterations; mass_01/02: the error depends on the exact
stopping iteration).
"""
import numpy as np


class SolverControlNoConvergence(RuntimeError):
    def __init__(self, last_step, last_residual):
        super().__init__(f"Iterative method reported convergence failure in step {last_step}. "
                         f"The residual in the last step was {last_residual}.")
        self.last_step, self.last_residual = last_step, last_residual


class ReductionControl:
    """`ReductionControl(max_steps, tolerance, reduce)`.

    success when  value < reduce * initial_value  (strict) or  value <= tolerance;
    failure when  step >= max_steps  (checked after the success tests) or NaN.
    """

    def __init__(self, max_steps=100, tolerance=1e-10, reduce=1e-2):
        self.max_steps, self.tol, self.reduce = max_steps, tolerance, reduce
        self._last_step, self._last_value, self._initial = 0, 0.0, 0.0

    def check(self, step, value):
        if step == 0:
            self._initial = value
            self._reduced_tol = value * self.reduce
        self._last_step, self._last_value = step, value
        if value < self._reduced_tol:
            return "success"
        if value <= self.tol:
            return "success"
        if step >= self.max_steps or np.isnan(value):
            return "failure"
        return "iterate"

    def last_step(self):
        return self._last_step

    def last_value(self):
        return self._last_value

    def initial_value(self):
        return self._initial


class PreconditionIdentity:
    def vmult(self, r):
        return r.copy()


class PreconditionJacobi:
    """`PreconditionJacobi<SparseMatrix>` with relaxation 1: z = D^-1 r."""

    def __init__(self, A):
        self.inv_diag = 1.0 / A.diagonal()

    def vmult(self, r):
        return self.inv_diag * r


class DiagonalMatrix:
    def __init__(self, diag):
        self.diag = np.asarray(diag, dtype=float)

    def vmult(self, r):
        return self.diag * r


def solver_cg(A, x, b, precond, control):
    """deal.II `SolverCG::solve(A, x, b, P)`; `A` has `.dot` or is a callable vmult.

    g = A x - b (x = 0 => g = -b);  check(0, |g|);  then
      h = P g ; d = -h (+ beta d) ; Ad = A d ; alpha = (g.h)/(d.Ad)
      x += alpha d ; g += alpha Ad ; check(it, |g|)
    Returns x; raises SolverControlNoConvergence like deal.II.
    """
    vmult = A if callable(A) else (lambda v: A @ v)
    x = np.array(x, dtype=float)
    if np.any(x != 0.0):
        g = vmult(x) - b
    else:
        g = -np.array(b, dtype=float)
    res = float(np.sqrt(g @ g))
    state = control.check(0, res)
    if state == "success":
        return x
    it, gh, d = 0, 0.0, None
    while state == "iterate":
        it += 1
        h = precond.vmult(g)
        if it > 1:
            beta = gh
            gh = float(g @ h)
            beta = gh / beta
            d = -h + beta * d
        else:
            d = -h
            gh = float(g @ h)
        Ad = vmult(d)
        alpha = float(d @ Ad)
        alpha = gh / alpha
        x += alpha * d
        g += alpha * Ad
        res = float(np.sqrt(g @ g))
        state = control.check(it, res)
    if state != "success":
        raise SolverControlNoConvergence(control.last_step(), control.last_value())
    return x


class ExplicitRungeKutta4:
    """`TimeStepping::ExplicitRungeKutta` with `RK_CLASSIC_FOURTH_ORDER`.

    c = (0, 1/2, 1/2, 1), a21 = a32 = 1/2, a43 = 1, b = (1/6, 1/3, 1/3, 1/6);
    stages k_i = f(t + c_i dt, y + dt sum_j a_ij k_j); y <- y + dt sum_i b_i k_i
    accumulated in stage order (`y.sadd(1, dt*b_i, k_i)`).
    """
    c = (0.0, 0.5, 0.5, 1.0)
    a = ((), (0.5,), (0.0, 0.5), (0.0, 0.0, 1.0))
    b = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)

    def evolve_one_time_step(self, f, t, dt, y):
        ks = []
        for i in range(4):
            Y = np.array(y, dtype=float)
            for j, aij in enumerate(self.a[i]):
                if aij != 0.0:
                    Y = Y + (dt * aij) * ks[j]
            ks.append(f(t + self.c[i] * dt, Y))
        for i in range(4):
            y = y + (dt * self.b[i]) * ks[i]
        return t + dt, y


class DiscreteTime:
    """`dealii::DiscreteTime(start, end, desired_step)`.

    next = current + step, snapped to `end` when it would pass
    `end - 0.05*step`; so the last step is shortened (or slightly stretched)
    to land exactly on `end` (evidence: applications/wave/tests/heat_0.output:11-13).
    """

    def __init__(self, start, end, step):
        self.start, self.end, self.desired = start, end, step
        self.current = start
        self.next = self._next(start)
        self.n = 0

    def _next(self, t):
        nxt = t + self.desired
        if nxt > self.end - 0.05 * self.desired:
            nxt = self.end
        return nxt

    def is_at_end(self):
        return self.current == self.end

    def get_current_time(self):
        return self.current

    def get_next_time(self):
        return self.next

    def get_next_step_size(self):
        return self.next - self.current

    def get_step_number(self):
        return self.n

    def advance_time(self):
        self.n += 1
        self.current = self.next
        self.next = self._next(self.current)
