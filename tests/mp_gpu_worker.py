"""One rank of the multi-GPU parity test (launched by torchrun, NCCL): slab-partitioned apply, dot, CG
vs the oracle on the same global inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import gdm_b200 as g
    import oracle as O
    ctx = g.init_distributed()
    rank, world = ctx.rank, ctx.n_ranks
    ok = True
    for (dim, p, reps, kernel) in [(3, 3, [12, 11, 30], g.capi.KERNEL_FUSED), (3, 3, [12, 11, 30], g.capi.KERNEL_GENERIC),
                                   (2, 3, [15, 40], g.capi.KERNEL_GENERIC), (3, 5, [13, 12, 41], g.capi.KERNEL_FUSED)]:
        gs = g.System(dim, p, 1, comm="world", context=ctx)
        gs.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
        gc = g.AffineConstraints()
        gs.make_zero_boundary_constraints(gc)
        gc.close()
        so = O.System(dim, p)
        so.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
        co = O.Constraints()
        so.make_zero_boundary_constraints(co)
        co.close()
        Ao = O.kron_operator(so, co, "stiffness")
        A = g.SparseMatrix()
        g.MatrixCreator.create_laplace_matrix(g.MappingQ1(), gs, g.QGauss(p + 1), A, gc, kernel=kernel)
        own = gs.locally_owned_dofs()
        xg = np.random.default_rng(3).uniform(-1, 1, so.n_dofs())
        x, y = g.Vector(gs, xg[own.start:own.stop]), g.Vector(gs)
        A.vmult(y, x)
        ref = Ao @ xg
        err = np.abs(y.numpy() - ref[own.start:own.stop]).max() / np.abs(ref).max()
        dot = x * y
        dref = float(xg @ ref)
        bh = O.rhs_cell_loop(so, co, lambda pts, c: 1.0)
        b, u = g.Vector(gs, bh[own.start:own.stop]), g.Vector(gs)
        ctl = g.ReductionControl(500, 1e-12, 1e-8)
        g.SolverCG(ctl).solve(A, u, b, g.PreconditionIdentity())
        octl = O.ReductionControl(500, 1e-12, 1e-8)
        uo = O.solver_cg(Ao, np.zeros(so.n_dofs()), bh, O.PreconditionIdentity(), octl)
        uerr = np.abs(u.numpy() - uo[own.start:own.stop]).max() / np.abs(uo).max()
        good = err <= 1e-12 and abs(dot - dref) <= 1e-12 * abs(dref) and abs(ctl.last_step() - octl.last_step()) <= 1 and uerr <= 1e-7
        ok &= bool(good)
        print(f"rank {rank}/{world} dim {dim} p {p} kernel {A.kernel_used()}: apply {err:.1e} dot {abs(dot - dref) / abs(dref):.1e} "
              f"cg {ctl.last_step()}/{octl.last_step()} u {uerr:.1e} {'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([0.0 if ok else 1.0], device="cuda")
    torch.distributed.all_reduce(t)
    torch.distributed.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
