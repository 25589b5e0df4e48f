"""Oracle-only problem set-up (no product import): shared by the CPU tests."""
import oracle as O


def make_oracle_pair(dim, p, nc, reps, bc):
    s = O.System(dim, p, nc)
    s.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0 + 0.25 * d for d in range(dim)])
    c = O.Constraints()
    if bc == "dirichlet":
        s.make_zero_boundary_constraints(c)
    elif bc == "periodic":
        for d in range(dim):
            s.make_periodicity_constraints(d, c)
    elif bc == "mixed":
        for f in (0, 1):
            s.make_zero_boundary_constraints(c, f)
        for d in range(1, dim):
            s.make_periodicity_constraints(d, c)
    elif bc == "left":
        s.make_zero_boundary_constraints(c, 0)
    else:
        assert bc == "none"
    c.close()
    return s, c
