"""One rank of the CPU (gloo) multi-rank test of the cut-cell rows: the rank generates the rows of its own slab
(gdm_cut_poisson_create with row_begin / row_end = gdm_system_locally_owned_range), imports the ghost planes of the input
with the library's plan (system created with add_ghost_layer = 1: ghost-penalty columns reach p + 1 planes) and applies
tensor-product rows + attached rows to locally stored data only; compared with the one-rank oracle matrix.

usage: python mp_cut_worker.py RANK WORLD PORT DIM P N
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port, dim, p, n1 = (int(a) for a in sys.argv[1:7])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gdm_b200 as g
    import oracle as O
    import scipy.sparse as sp
    from oracle import cut
    from gdm_b200 import capi
    lib = capi.load()
    ctx = C.c_void_p()
    assert lib.gdm_context_create(-1, None, C.byref(ctx)) == 0
    d = capi.SystemDesc()
    d.dim, d.fe_degree, d.n_components = dim, p, 1
    for i in range(dim):
        d.n_subdivisions[i], d.lo[i], d.hi[i] = n1, -1.21, 1.21
    d.rank, d.n_ranks, d.add_ghost_layer = rank, world, 1
    sysh = C.c_void_p()
    assert lib.gdm_system_create(ctx, C.byref(d), C.byref(sysh)) == 0, lib.gdm_last_error()
    info = capi.LayoutInfo()
    lib.gdm_system_layout(sysh, C.byref(info))
    plan = (C.c_int32 * 10)()
    assert lib.gdm_system_halo_plan(sysh, plan) == 0, lib.gdm_last_error()
    prev, nxt, slp, slc, rlp, rlc, shp, shc, rhp, rhc = list(plan)
    b, e = C.c_uint64(), C.c_uint64()
    assert lib.gdm_system_locally_owned_range(sysh, C.byref(b), C.byref(e)) == 0

    so = O.System(dim, p)
    so.subdivided_hyper_cube(n1, -1.21, 1.21)
    ls = cut.interpolate_level_set(so, cut.sphere_level_set([0.0] * dim, 1.0))
    A, rhs, _ = cut.assemble_cut_poisson(so, ls, ghost_penalty=True)  # the one-rank reference
    part = g.CutPoisson(dim, p, [n1] * dim, [-1.21] * dim, [1.21] * dim, ls, ghost_penalty=True, row_range=(b.value, e.value))
    rows, rowptr, col, val = part.rows()
    face = int(np.prod(so.n_nodes[:-1])) if dim > 1 else 1
    sb, se, ob, oe = info.stored_begin, info.stored_end, info.owned_begin, info.owned_end
    assert (b.value, e.value) == (ob * face, oe * face)
    ok = bool(np.all((rows >= b.value) & (rows < e.value)))
    ok &= bool(np.all((col >= sb * face) & (col < se * face)))  # what gdm_operator_attach_csr requires of the columns

    xg = np.random.default_rng(7).uniform(-1, 1, so.n_dofs())
    yg = A @ xg
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
    ok &= not np.isnan(loc).any()
    xl = np.zeros(so.n_dofs())
    xl.reshape(-1, face)[sb:se] = loc
    # tensor-product rows, replaced by the attached rows, on the owned range
    yl = O.kron_unconstrained(so, "stiffness") @ xl
    M = sp.csr_matrix((val, col.astype(np.int64), rowptr.astype(np.int64)), shape=(len(rows), so.n_dofs()))
    yl[rows.astype(np.int64)] = M @ xl
    err = np.abs(yl[b.value:e.value] - yg[b.value:e.value]).max() if e.value > b.value else 0.0
    ok &= err <= 1e-12 * np.abs(yg).max()
    ok &= bool(np.abs(part.rhs()[b.value:e.value] - rhs[b.value:e.value]).max() <= 1e-14 * np.abs(rhs).max()) if e.value > b.value else True
    counts = torch.tensor([float(len(rows)), float(len(col))])
    dist.all_reduce(counts)
    full = g.CutPoisson(dim, p, [n1] * dim, [-1.21] * dim, [1.21] * dim, ls, ghost_penalty=True).sizes()
    ok &= (int(counts[0].item()), int(counts[1].item())) == (full[0], full[1])
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: rows {len(rows)} nnz {len(col)} err {err:.2e} {'OK' if ok else 'FAIL'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
