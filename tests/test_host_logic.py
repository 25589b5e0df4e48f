"""Host-side logic of libgdm_b200 (grid, DoF windows, categories, partition, 1D band tables)
against the oracle -- runs on the CPU through a description-only context (no compute calls)."""
import ctypes as C

import numpy as np
import pytest

import oracle as O


class HostSystem:
    def __init__(self, lib, dim, p, nc, reps, lo, hi, rank=0, n_ranks=1, ghost_layer=False):
        from gdm_b200 import capi
        self.lib = lib
        self.ctx = C.c_void_p()
        assert lib.gdm_context_create(-1, None, C.byref(self.ctx)) == 0
        d = capi.SystemDesc()
        d.dim, d.fe_degree, d.n_components = dim, p, nc
        for i in range(dim):
            d.n_subdivisions[i], d.lo[i], d.hi[i] = reps[i], lo[i], hi[i]
        d.rank, d.n_ranks, d.add_ghost_layer = rank, n_ranks, int(ghost_layer)
        self.h = C.c_void_p()
        rc = lib.gdm_system_create(self.ctx, C.byref(d), C.byref(self.h))
        assert rc == 0, lib.gdm_last_error()
        self.capi = capi

    def close(self):
        self.lib.gdm_system_destroy(self.h)
        self.lib.gdm_context_destroy(self.ctx)


@pytest.mark.parametrize("p", [1, 3, 5, 7, 9])
def test_polynomials_1d(lib, p):
    import gdm_b200
    mine = gdm_b200.generate_polynomials_1D(p)
    ref = O.generate_polynomials_1D(p)
    for v in range(p):
        for k in range(p + 1):
            assert np.array_equal(mine[v, k], ref[v][k])


@pytest.mark.parametrize("p,N", [(1, 4), (3, 10), (3, 3), (5, 11), (5, 40), (7, 15), (9, 20)])
def test_band_matrices(lib, p, N):
    hs = HostSystem(lib, 1, p, 1, [N], [0.0], [2.0])
    M, K, Cm, f = O.matrices_1d(p, N, 2.0 / N)
    for kind, ref in ((0, M), (1, K), (2, Cm)):
        band = np.zeros((N + 1, 2 * p + 1))
        assert lib.gdm_system_matrix_1d(hs.h, 0, kind, band.ctypes.data_as(C.POINTER(C.c_double))) == 0
        dense = np.zeros((N + 1, N + 1))
        for i in range(N + 1):
            for t in range(2 * p + 1):
                j = i + t - p
                if 0 <= j <= N:
                    dense[i, j] = band[i, t]
                else:
                    assert band[i, t] == 0.0
        assert np.abs(dense - ref).max() <= 2e-13 * np.abs(ref).max(), (kind, np.abs(dense - ref).max())
    hs.close()


@pytest.mark.parametrize("dim,p,nc,reps", [(1, 3, 1, [10]), (2, 3, 2, [6, 9]), (3, 5, 1, [6, 7, 8]), (3, 1, 3, [3, 2, 4])])
def test_dof_indices_and_categories(lib, dim, p, nc, reps):
    hs = HostSystem(lib, dim, p, nc, reps, [0.0] * dim, [1.0] * dim)
    s = O.System(dim, p, nc)
    s.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
    assert lib.gdm_system_n_dofs(hs.h) == s.n_dofs()
    assert lib.gdm_system_n_cells(hs.h) == s.n_cells()
    npc = lib.gdm_system_dofs_per_cell(hs.h)
    assert npc == nc * (p + 1) ** dim
    out = np.zeros(npc, dtype=np.uint64)
    cat = C.c_uint32()
    for cell in range(s.n_cells()):
        assert lib.gdm_system_get_dof_indices(hs.h, cell, out.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
        assert list(out) == s.get_dof_indices(cell)
        assert lib.gdm_system_active_fe_index(hs.h, cell, C.byref(cat)) == 0
        assert cat.value == s.active_fe_index(cell)
    hs.close()


@pytest.mark.parametrize("dim,reps,n_ranks", [(1, [20], 3), (2, [20, 20], 3), (3, [8, 8, 64], 8), (2, [9, 10], 4), (3, [5, 5, 7], 2)])
def test_slab_partition(lib, dim, reps, n_ranks):
    """system.h:720-757: stride = ceil(N_last / n), rank 0 gets plane 0 too."""
    covered = 0
    for rank in range(n_ranks):
        hs = HostSystem(lib, dim, 3, 1, reps, [0.0] * dim, [1.0] * dim, rank, n_ranks)
        s = O.System(dim, 3, 1, rank, n_ranks)
        s.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
        b, e = C.c_uint64(), C.c_uint64()
        assert lib.gdm_system_locally_owned_range(hs.h, C.byref(b), C.byref(e)) == 0
        assert (b.value, e.value) == s.locally_owned_range()
        assert b.value == covered
        covered = e.value
        info = hs.capi.LayoutInfo()
        assert lib.gdm_system_layout(hs.h, C.byref(info)) == 0
        a0, a1 = s.owned_plane_range()
        assert (info.owned_begin, info.owned_end) == (a0, a1)
        if a1 > a0:
            assert info.stored_begin == max(0, a0 - 3) and info.stored_end == min(reps[-1] + 1, a1 + 3)
        assert info.pitch % 4 == 0 and info.size >= info.owned_offset + info.owned_size
        hs.close()
    assert covered == int(np.prod([r + 1 for r in reps]))


def test_constraints_bookkeeping(lib):
    hs = HostSystem(lib, 2, 3, 2, [6, 7], [0.0, 0.0], [1.0, 1.0])
    s = O.System(2, 3, 2)
    s.subdivided_hyper_rectangle([6, 7], [0.0, 0.0], [1.0, 1.0])
    c = C.c_void_p()
    assert lib.gdm_constraints_create(hs.h, C.byref(c)) == 0
    oc = O.Constraints()
    assert lib.gdm_constraints_make_zero_boundary(c, 0) == 0
    s.make_zero_boundary_constraints(oc, 0)
    assert lib.gdm_constraints_make_zero_boundary(c, 1) == 0
    s.make_zero_boundary_constraints(oc, 1)
    assert lib.gdm_constraints_make_periodicity(c, 1) == 0
    s.make_periodicity_constraints(1, oc)
    assert lib.gdm_constraints_close(c) == 0
    oc.close()
    assert lib.gdm_constraints_n_constraints(c) == len(oc.lines)
    for i in range(s.n_dofs()):
        assert bool(lib.gdm_constraints_is_constrained(c, i)) == oc.is_constrained(i)
    lib.gdm_constraints_destroy(c)
    hs.close()


def test_error_reporting(lib):
    from gdm_b200 import capi
    ctx = C.c_void_p()
    assert lib.gdm_context_create(-1, None, C.byref(ctx)) == 0
    d = capi.SystemDesc()
    d.dim, d.fe_degree, d.n_components = 2, 4, 1  # even degree: fe.h:322 AssertIndexRange
    d.n_subdivisions[0] = d.n_subdivisions[1] = 8
    d.hi[0] = d.hi[1] = 1.0
    d.n_ranks = 1
    h = C.c_void_p()
    assert lib.gdm_system_create(ctx, C.byref(d), C.byref(h)) == capi.ERR_NOT_IMPLEMENTED
    d.fe_degree = 3
    d.n_subdivisions[1] = 2  # fewer cells than the degree
    assert lib.gdm_system_create(ctx, C.byref(d), C.byref(h)) == capi.ERR_INVALID
    assert b"n_subdivisions" in lib.gdm_last_error()
    lib.gdm_context_destroy(ctx)


@pytest.mark.parametrize("dim,p,reps,nc", [(1, 3, [9], 1), (1, 5, [11], 1), (2, 3, [7, 6], 1), (2, 1, [4, 5], 2), (2, 5, [11, 12], 1),
                                           (3, 3, [7, 6, 8], 1), (3, 1, [3, 4, 3], 1)])
@pytest.mark.parametrize("flux", [False, True])
def test_sparsity_pattern_rows(lib, dim, p, reps, nc, flux):
    """System::create_sparsity_pattern / create_flux_sparsity_pattern (system.h:586-630): the rows generated on demand
    by the C ABI (boxes from the per-direction window rules) equal the oracle's literal cell / face loops."""
    import gdm_b200 as g
    import oracle as O
    desc_ctx = g.Context(device=-1)
    gs = g.System(dim, p, nc, add_ghost_layer=flux, context=desc_ctx)
    gs.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
    so = O.System(dim, p, nc)
    so.subdivided_hyper_rectangle(reps, [0.0] * dim, [1.0] * dim)
    ref = so.create_sparsity_pattern(flux=flux)
    for row in range(so.n_dofs()):
        assert gs.sparsity_row(row, flux) == ref[row], row
    if dim == 3 and p == 3 and not flux:  # interior rows couple (2p+1)^3 = 343 nodes
        assert max(len(r) for r in ref) == 343


@pytest.mark.parametrize("dim,p,reps,nc,kind,bc", [(1, 3, [9], 1, "mass", "dirichlet"), (2, 3, [7, 6], 1, "stiffness", "dirichlet"),
                                                   (2, 1, [4, 5], 2, "mass", "none"), (3, 3, [7, 6, 8], 1, "stiffness", "dirichlet"),
                                                   (2, 5, [11, 12], 1, "advection", "left")])
@pytest.mark.parametrize("binary", [False, True])
def test_write_matrix_to_file(lib, tmp_path, dim, p, reps, nc, kind, bc, binary):
    """The triplet dump of the reference's eigenvalue tool (applications/wave/wave-ev.cc:93-127) produced on the host from
    the 1D tables: pattern = create_sparsity_pattern, per row the diagonal first and then ascending columns (iteration
    order of a deal.II SparseMatrix), values = the oracle's assembled operator."""
    import gdm_b200 as g
    import oracle as O
    from gdm_b200 import capi
    from helpers_cpu import make_oracle_pair
    ctx = g.Context(device=-1)
    gs = g.System(dim, p, nc, context=ctx)
    hi = [1.0 + 0.25 * d for d in range(dim)]
    gs.subdivided_hyper_rectangle(reps, [0.0] * dim, hi)
    gc = g.AffineConstraints()
    if bc == "dirichlet":
        gs.make_zero_boundary_constraints(gc)
    elif bc == "left":
        gs.make_zero_boundary_constraints(0, gc)
    gc.close()
    so, co = make_oracle_pair(dim, p, nc, reps, bc)
    bvec = [1.0, 0.15, -0.05][:dim]
    if kind == "advection":
        Ao = O.kron_operator(so, co, "advection", b=bvec, constrained_diagonal="zero").tocsr()
        code, diag = capi.OP_ADVECTION, capi.DIAG_ZERO
    else:
        Ao = O.kron_operator(so, co, kind).tocsr()
        code, diag = (capi.OP_MASS if kind == "mass" else capi.OP_STIFFNESS), capi.DIAG_ASSEMBLED
    path = tmp_path / "matrix.txt"
    n = gs.write_matrix_to_file(gc, code, path, binary, b=bvec, constrained_diagonal=diag)
    if binary:
        rec = np.fromfile(path, dtype=np.dtype([("r", "<u4"), ("c", "<u4"), ("v", "<f8")]))
        rows, cols, vals = rec["r"].astype(np.int64), rec["c"].astype(np.int64), rec["v"]
    else:
        t = np.loadtxt(path, ndmin=2)
        rows, cols, vals = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64), t[:, 2]
    assert len(rows) == n
    pattern = so.create_sparsity_pattern()
    k = 0
    for r in range(so.n_dofs()):
        expect = [r] + [c for c in pattern[r] if c != r]
        assert list(cols[k:k + len(expect)]) == expect and np.all(rows[k:k + len(expect)] == r)
        k += len(expect)
    assert k == n
    ref = np.asarray(Ao[rows, cols]).reshape(-1)
    tol = 1e-14 if binary else 6e-6  # text: operator<< of a double prints 6 significant digits
    assert np.abs(vals - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("dim,reps,nc", [(1, [7], 1), (2, [5, 4], 2), (3, [3, 4, 2], 1)])
def test_write_vtu(lib, tmp_path, dim, reps, nc):
    """Stand-in for GDM::DataOut (include/gdm/data_out.h): points = nodes, cells = grid cells, DoFs as point data."""
    import xml.etree.ElementTree as ET
    import gdm_b200 as g
    ctx = g.Context(device=-1)
    gs = g.System(dim, 3, nc, context=ctx)
    hi = [1.0 + 0.5 * d for d in range(dim)]
    gs.subdivided_hyper_rectangle([max(r, 3) for r in reps], [0.0] * dim, hi)
    reps = [max(r, 3) for r in reps]
    vals = np.arange(gs.n_dofs(), dtype=float) * 0.5
    path = tmp_path / "field.vtu"
    gs.write_vtu(vals, "solution", path)
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    n_nodes, n_cells = int(np.prod([r + 1 for r in reps])), int(np.prod(reps))
    assert int(piece.get("NumberOfPoints")) == n_nodes and int(piece.get("NumberOfCells")) == n_cells
    pts = np.array(piece.find("Points/DataArray").text.split(), dtype=float).reshape(-1, 3)
    assert pts.shape[0] == n_nodes and abs(pts[:, 0].max() - hi[0]) < 1e-14 and (dim < 2 or abs(pts[:, 1].max() - hi[1]) < 1e-14)
    arrays = {a.get("Name"): a for a in piece.find("Cells").findall("DataArray")}
    conn = np.array(arrays["connectivity"].text.split(), dtype=int).reshape(n_cells, 2 ** dim)
    assert conn.min() == 0 and conn.max() == n_nodes - 1 and all(len(set(c)) == 2 ** dim for c in conn)
    assert set(arrays["types"].text.split()) == {str({1: 3, 2: 9, 3: 12}[dim])}
    pd = piece.find("PointData").findall("DataArray")
    assert len(pd) == nc
    for c, a in enumerate(pd):
        assert a.get("Name") == ("solution" if nc == 1 else f"solution_{c}")
        assert np.array_equal(np.array(a.text.split(), dtype=float), vals[c::nc])
