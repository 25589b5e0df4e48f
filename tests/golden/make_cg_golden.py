"""Golden CG iteration counts of the SURVEY 8d Poisson solve, produced by the ORACLE (CPU, numpy/scipy).

    -Laplace u = 1 on [0,1]^3, u = 0 on the boundary, GDM degree p, N cells per direction,
    b_i = int phi_i (= Kronecker product of the 1D load vectors, zero on constrained rows), x0 = 0,
    deal.II SolverCG + ReductionControl(10000, 1e-10, 1e-8), PreconditionIdentity and PreconditionJacobi
    (reference call site: tests/poisson_02_gdm.cc:213-215).

The operator is applied matrix-free (oracle/kron_apply.py); at N = 256 one solve takes tens of minutes on the
CPU, which is why the counts are committed (tests/golden/cg_poisson3d.json) instead of recomputed in the tests.
usage: python tests/golden/make_cg_golden.py [N ...]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle as O  # noqa: E402
from oracle.assemble import matrices_1d  # noqa: E402

OUT = os.path.join(HERE, "cg_poisson3d.json")


def solve(N, p, precond):
    s = O.System(3, p)
    s.subdivided_hyper_cube(N)
    c = O.Constraints()
    s.make_zero_boundary_constraints(c)
    c.close()
    A = O.KronApply(s, c, "stiffness")
    f = matrices_1d(p, N, 1.0 / N)[3]
    b = np.einsum("k,j,i->kji", f, f, f).reshape(-1)
    b[A.constrained_mask()] = 0.0
    ctl = O.ReductionControl(10000, 1e-10, 1e-8)
    P = O.PreconditionIdentity() if precond == "identity" else O.PreconditionJacobi(A)
    t0 = time.time()
    u = O.solver_cg(A, np.zeros(s.n_dofs()), b, P, ctl)
    return {"N": N, "p": p, "precondition": precond, "iterations": ctl.last_step(), "initial_residual": ctl.initial_value(),
            "final_residual": ctl.last_value(), "u_max": float(np.abs(u).max()), "u_center": float(u.reshape(N + 1, N + 1, N + 1)[N // 2, N // 2, N // 2]),
            "oracle_seconds": round(time.time() - t0, 1)}


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256]
    res = json.load(open(OUT)) if os.path.exists(OUT) else []
    for N in sizes:
        for pre in ("identity", "jacobi"):
            if any(r["N"] == N and r["p"] == 3 and r["precondition"] == pre for r in res):
                continue
            r = solve(N, 3, pre)
            print(r, flush=True)
            res.append(r)
            json.dump(res, open(OUT, "w"), indent=1)
