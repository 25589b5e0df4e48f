"""One rank of the multi-GPU parity test (launched by torchrun, NCCL): slab-partitioned apply, fused dot, CG and the
periodic wrap of the partitioned direction vs the oracle on the same global inputs.

The z extent grows with the world size so that every rank owns more than 4p planes: the overlapped branch of the
fused apply (slab faces behind the ghost import on the communication stream, interior planes on the main stream) is
the one the SCALE benchmark runs, and it must be the one that is checked here for any number of ranks."""
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
    F, G = g.capi.KERNEL_FUSED, g.capi.KERNEL_GENERIC
    cases = [
        # dim, p, reps, kernel, bc, kind
        (3, 3, [12, 11, (4 * 3 + 6) * world], F, "dirichlet", "stiffness"),
        (3, 3, [12, 11, 30], F, "dirichlet", "stiffness"),          # thin slabs: serialised branch
        (3, 3, [12, 11, 30], G, "dirichlet", "stiffness"),
        (2, 3, [15, 40], G, "dirichlet", "stiffness"),
        (3, 5, [13, 12, (4 * 5 + 6) * world], F, "dirichlet", "stiffness"),
        (3, 1, [34, 9, 10 * world], F, "none", "mass"),
        (3, 3, [12, 11, (4 * 3 + 6) * world], F, "periodic", "stiffness"),   # wrap between the last and the first rank
        (3, 5, [13, 12, (4 * 5 + 6) * world], F, "periodic", "advection"),
        (3, 3, [35, 33, (4 * 3 + 8) * world], F, "mixed", "mass"),
        (3, 3, [12, 11, 52 * world], F, "dirichlet", "stiffness"),   # >= 16p owned planes: pipelined host-buffer apply
    ]
    for (dim, p, reps, kernel, bc, kind) in cases:
        gs = g.System(dim, p, 1, comm="world", context=ctx)
        gs.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
        so = O.System(dim, p)
        so.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
        gc, co = g.AffineConstraints(), O.Constraints()
        if bc == "dirichlet":
            gs.make_zero_boundary_constraints(gc)
            so.make_zero_boundary_constraints(co)
        elif bc == "periodic":
            for d in range(dim):
                gs.make_periodicity_constraints(d, gc)
                so.make_periodicity_constraints(d, co)
        elif bc == "mixed":  # Dirichlet on the x faces, periodic in y and z
            for s in (0, 1):
                gs.make_zero_boundary_constraints(s, gc)
                so.make_zero_boundary_constraints(co, s)
            for d in range(1, dim):
                gs.make_periodicity_constraints(d, gc)
                so.make_periodicity_constraints(d, co)
        gc.close()
        co.close()
        bvec = [1.0, 0.15, -0.05][:dim]
        A = g.SparseMatrix()
        m, q = g.MappingQ1(), g.QGauss(p + 1)
        if kind == "stiffness":
            Ao = O.kron_operator(so, co, "stiffness")
            g.MatrixCreator.create_laplace_matrix(m, gs, q, A, gc, kernel=kernel)
        elif kind == "mass":
            Ao = O.kron_operator(so, co, "mass")
            g.MatrixCreator.create_mass_matrix(m, gs, q, A, gc, kernel=kernel)
        else:
            Ao = O.kron_operator(so, co, "advection", b=bvec, constrained_diagonal="zero")
            g.MatrixCreator.create_advection_matrix(m, gs, q, A, gc, bvec, kernel=kernel)
        assert A.kernel_used() == kernel
        own = gs.locally_owned_dofs()
        xg = np.random.default_rng(3).uniform(-1, 1, so.n_dofs())
        x, y, y2 = g.Vector(gs, xg[own.start:own.stop]), g.Vector(gs), g.Vector(gs)
        A.vmult(y, x)
        ref = Ao @ xg
        err = np.abs(y.numpy() - ref[own.start:own.stop]).max() / np.abs(ref).max()
        restored = np.array_equal(x.numpy(), xg[own.start:own.stop])
        dot = x * y
        dref = float(xg @ ref)
        fdot = A.vmult_dot(y2, x)  # apply with <x, A x> in the store epilogue (CG: p . A p), reduced over the ranks
        same = np.array_equal(y2.numpy(), y.numpy())
        dscale = float(np.abs(xg * ref).sum())
        good = err <= 1e-12 and restored and same and abs(dot - dref) <= 1e-12 * dscale and abs(fdot - dref) <= 1e-12 * dscale
        line = f"apply {err:.1e} dot {abs(dot - dref) / dscale:.1e} fused-dot {abs(fdot - dref) / dscale:.1e}"
        if kind == "stiffness" and bc == "dirichlet":
            bh = O.rhs_cell_loop(so, co, lambda pts, c: 1.0)
            b, u = g.Vector(gs, bh[own.start:own.stop]), g.Vector(gs)
            ctl = g.ReductionControl(500, 1e-12, 1e-8)
            g.SolverCG(ctl).solve(A, u, b, g.PreconditionIdentity())
            octl = O.ReductionControl(500, 1e-12, 1e-8)
            uo = O.solver_cg(Ao, np.zeros(so.n_dofs()), bh, O.PreconditionIdentity(), octl)
            uerr = np.abs(u.numpy() - uo[own.start:own.stop]).max() / np.abs(uo).max()
            good = good and abs(ctl.last_step() - octl.last_step()) <= 1 and uerr <= 1e-7
            line += f" cg {ctl.last_step()}/{octl.last_step()} u {uerr:.1e}"
        if reps[-1] >= 16 * p * world:
            # host-buffer entry point (H2D, apply, D2H pipelined over z chunks; the slab faces follow the ghost exchange)
            yh = np.zeros(own.stop - own.start)
            A.vmult_host(yh, np.ascontiguousarray(xg[own.start:own.stop]))
            herr = np.abs(yh - ref[own.start:own.stop]).max() / np.abs(ref).max()
            good = good and herr <= 1e-12
            line += f" host {herr:.1e}"
        if bc in ("periodic", "mixed"):
            # AffineConstraints::distribute: the duplicate plane of the partitioned direction comes from rank 0
            v = g.Vector(gs, xg[own.start:own.stop])
            gc.distribute(v)
            vo = co.distribute(xg.copy())
            derr = np.abs(v.numpy() - vo[own.start:own.stop]).max()
            good = good and derr == 0.0
            line += f" distribute {derr:.1e}"
        ok &= bool(good)
        print(f"rank {rank}/{world} dim {dim} p {p} {bc} {kind} kernel {A.kernel_used()}: {line} {'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([0.0 if ok else 1.0], device="cuda")
    torch.distributed.all_reduce(t)
    torch.distributed.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
