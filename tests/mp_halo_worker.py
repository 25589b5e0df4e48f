"""One rank of the CPU (gloo) multi-rank test: executes the library's ghost-import plan
(gdm_system_halo_plan) on numpy slabs through torch.distributed and checks that every rank can then
apply the oracle operator to its owned rows from locally stored data only.

usage: python mp_halo_worker.py RANK WORLD PORT DIM P NX NY NZ
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port, dim, p = (int(a) for a in sys.argv[1:6])
    reps = [int(a) for a in sys.argv[6:6 + dim]]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gdm_b200
    import oracle as O
    from gdm_b200 import capi
    lib = capi.load()
    ctx = C.c_void_p()
    assert lib.gdm_context_create(-1, None, C.byref(ctx)) == 0
    d = capi.SystemDesc()
    d.dim, d.fe_degree, d.n_components = dim, p, 1
    for i in range(dim):
        d.n_subdivisions[i], d.lo[i], d.hi[i] = reps[i], 0.0, 1.0
    d.rank, d.n_ranks = rank, world
    sysh = C.c_void_p()
    assert lib.gdm_system_create(ctx, C.byref(d), C.byref(sysh)) == 0, lib.gdm_last_error()
    info = capi.LayoutInfo()
    lib.gdm_system_layout(sysh, C.byref(info))
    plan = (C.c_int32 * 10)()
    assert lib.gdm_system_halo_plan(sysh, plan) == 0, lib.gdm_last_error()
    prev, nxt, slp, slc, rlp, rlc, shp, shc, rhp, rhc = list(plan)

    # global reference
    so = O.System(dim, p)
    so.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
    co = O.Constraints()
    so.make_zero_boundary_constraints(co)
    co.close()
    A = O.kron_operator(so, co, "stiffness")
    xg = np.random.default_rng(7).uniform(-1, 1, so.n_dofs())
    yg = A @ xg

    # local slab: planes [stored_begin, stored_end) of the last direction, ghost planes poisoned
    face = int(np.prod(so.n_nodes[:-1])) if dim > 1 else 1
    sb, se, ob, oe = info.stored_begin, info.stored_end, info.owned_begin, info.owned_end
    loc = np.full((se - sb, face), np.nan)
    loc[ob - sb:oe - sb] = xg.reshape(-1, face)[ob:oe]
    t = torch.from_numpy(loc)
    reqs = []
    if prev >= 0:
        reqs.append(dist.isend(t[slp:slp + slc].clone(), prev))
        lo_buf = torch.empty(rlc, face, dtype=torch.float64)
        reqs.append(dist.irecv(lo_buf, prev))
    if nxt >= 0:
        reqs.append(dist.isend(t[shp:shp + shc].clone(), nxt))
        hi_buf = torch.empty(rhc, face, dtype=torch.float64)
        reqs.append(dist.irecv(hi_buf, nxt))
    for r in reqs:
        r.wait()
    if prev >= 0:
        t[rlp:rlp + rlc] = lo_buf
    if nxt >= 0:
        t[rhp:rhp + rhc] = hi_buf
    assert not np.isnan(loc).any(), "ghost zone not completely filled"
    # owned rows from local data only
    xl = np.zeros(so.n_dofs())
    xl.reshape(-1, face)[sb:se] = loc
    yl = (A @ xl).reshape(-1, face)[ob:oe]
    err = np.abs(yl - yg.reshape(-1, face)[ob:oe]).max() if oe > ob else 0.0
    # the slabs tile the grid
    owned = torch.tensor([float(oe - ob)])
    dist.all_reduce(owned)
    ok = err <= 1e-13 * np.abs(yg).max() and int(owned.item()) == so.n_nodes[-1]
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: owned [{ob},{oe}) stored [{sb},{se}) err {err:.2e} {'OK' if ok else 'FAIL'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
