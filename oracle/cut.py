"""Cut-cell set-up restated (oracle, test-only): level-set classification, cut quadrature, CutFEM Poisson assembly.

What the reference does with deal.II's `NonMatching` classes before the hot path starts:

* `classify`            `NonMatching::MeshClassifier::reclassify` for a Q1 level set: a cell is inside when every
                        vertex value is negative, outside when every one is positive, intersected otherwise (the
                        Bernstein coefficients of a Q1 function are its vertex values).  Call site
                        `prototypes/cut_poisson_01_gdm.cc:105-121`, `applications/wave/include/gdm/wave/discretization.h:79-97`.
* `cut_quadrature`      `NonMatching::FEValues` / `QuadratureGenerator` (Saye's height-function recursion) on the unit
                        cell for the cell's multilinear level set, with `QGauss<1>(p+1)` in every direction
                        (`prototypes/cut_poisson_01_gdm.cc:176-190`).  deal.II itself is not in the reference tree; this
                        is a restatement of the published algorithm specialised to multilinear functions, where every
                        bound deal.II estimates is exact (a multilinear function and its partial derivatives take their
                        extrema at the vertices, and a line in a coordinate direction crosses the zero set at most once).
                        The point sets are not pinned by any golden, only integrals are (5 digits of the L2 error of
                        `prototypes/cut_poisson_01_gdm.output`, 9 digits of `applications/wave/tests/wave_0.output` in 1D).
* `assemble_cut_poisson` the cell loop of `prototypes/cut_poisson_01_gdm.cc:196-329`: inside volume term, symmetric Nitsche
                        terms on the surface, ghost penalty on the faces `face_has_ghost_penalty` flags (`:123-146`),
                        zero diagonal -> 1 (`:324-329`).  `gp_h_power=3` gives the scaling of the wave application's
                        matrix (`applications/wave/include/gdm/wave/stiffness.h:760-765`).
* `l2_error_inside`     `prototypes/cut_poisson_01_gdm.cc:349-398`; `error_norms_inside`: the L2 / L1 / Linf columns of
                        `applications/wave/include/gdm/wave/problem.h:531-615`.
* `load_functionals`    the data-dependent part of the residual `wave/stiffness.h:186-260` (volume source, Nitsche data).
* `domain_boundary_terms`, `coupling_matrices`  the two-domain runs: Nitsche on the box boundary (`wave/stiffness.h:262-340`)
                        and the coupling on the cut surface (`:441-574`).
`oracle/wave_app.py` drives these through the wave application's runs; pinned by `tests/test_cut_cell.py`.
"""
import numpy as np
import scipy.sparse as sp

from .assemble import _cell_tables, cell_matrix_scalar
from .basis import basis_values, gauss_legendre_01
from .system import indices_to_index

INSIDE, OUTSIDE, INTERSECTED = 0, 1, 2


# ----------------------------------------------------------------------------------------------- level set, classes
def interpolate_level_set(system, fn):
    """Nodal values of the level-set function (`VectorTools::interpolate` into FE_Q(1)), DoF order of the grid."""
    return np.asarray(fn(system.node_coordinates()), dtype=float)


def sphere_level_set(center, radius):
    """`Functions::SignedDistance::Sphere` (default: unit sphere at the origin)."""
    c = np.asarray(center, dtype=float)
    return lambda pts: np.linalg.norm(pts - c, axis=1) - radius


def cell_vertex_values(system, ls, cell):
    """Level-set values at the 2^dim vertices of a cell, array of shape (2,)*dim indexed [x][y][z]."""
    idx = system.cell_indices(cell)
    nn = system.n_nodes
    out = np.zeros((2,) * system.dim)
    for corner in np.ndindex(*out.shape):
        out[corner] = ls[indices_to_index([idx[d] + corner[d] for d in range(system.dim)], nn)]
    return out


def classify(system, ls):
    """Per cell INSIDE / OUTSIDE / INTERSECTED (`location_to_level_set`).  `ls`: nodal values of a Q1 level set, or a
    geometry object with `classify(system)` and `rules(system, cell, n_gauss)` (`oracle/cut_q.py`: level set of degree p)."""
    if hasattr(ls, "classify"):
        return ls.classify(system)
    out = np.zeros(system.n_cells(), dtype=np.int8)
    for cell in range(system.n_cells()):
        v = cell_vertex_values(system, ls, cell)
        out[cell] = INSIDE if v.max() < 0 else (OUTSIDE if v.min() > 0 else INTERSECTED)
    return out


# ------------------------------------------------------------------------------------------------- cut quadrature
# A multilinear function on a box is carried as its 2^d corner values (axis e = direction e).
def _face(c, k, side):
    return np.take(c, side, axis=k)


def _eval(c, t):
    """Multilinear interpolation, t = local coordinates in [0,1]^d."""
    for x in t:  # contract the first axis each time
        c = c[0] * (1.0 - x) + c[1] * x
    return float(c)


def _split(c, k):
    mid = 0.5 * (_face(c, k, 0) + _face(c, k, 1))
    lo = np.stack([_face(c, k, 0), mid], axis=k)
    hi = np.stack([mid, _face(c, k, 1)], axis=k)
    return lo, hi


def _tensor_gauss(lo, hi, xg, wg):
    d = len(lo)
    if d == 0:
        return np.zeros((1, 0)), np.ones(1)
    axes = [lo[e] + (hi[e] - lo[e]) * xg for e in range(d)]
    ws = [(hi[e] - lo[e]) * wg for e in range(d)]
    grids = np.meshgrid(*axes, indexing="ij")
    wgrid = np.meshgrid(*ws, indexing="ij")
    w = np.ones_like(wgrid[0])
    for a in wgrid:
        w = w * a
    return np.stack([g.ravel() for g in grids], axis=1), w.ravel()


def _height_direction(funcs):
    """Direction in which every function is strictly monotone over the box (exact for multilinear functions); the one
    with the largest worst-case |d_k psi| relative to the gradient, None if there is none."""
    d = funcs[0].ndim
    best, best_score = None, 0.0
    for k in range(d):
        score = np.inf
        for c in funcs:
            dk = _face(c, k, 1) - _face(c, k, 0)
            if not (dk.min() > 0 or dk.max() < 0):
                score = 0.0
                break
            tot = sum(np.abs(_face(c, e, 1) - _face(c, e, 0)).max() for e in range(d))
            score = min(score, np.abs(dk).min() / tot)
        if score > best_score:
            best, best_score = k, score
    return best


def _volume(funcs, signs, lo, hi, xg, wg, depth=0):
    """Quadrature of {x in box : s_i psi_i(x) > 0 for all i with s_i != 0}, partitioned along the zero sets of all
    psi_i.  Returns points [n, d] (coordinates of the enclosing unit cell) and weights."""
    d = len(lo)
    keep_f, keep_s = [], []
    for c, s in zip(funcs, signs):
        if c.min() > 0:
            if s < 0:
                return np.zeros((0, d)), np.zeros(0)
        elif c.max() < 0:
            if s > 0:
                return np.zeros((0, d)), np.zeros(0)
        else:
            keep_f.append(c)
            keep_s.append(s)
    funcs, signs = keep_f, keep_s
    if not funcs:
        return _tensor_gauss(lo, hi, xg, wg)
    if d == 1:
        return _line(funcs, signs, np.zeros((1, 0)), np.ones(1), 0, lo, hi, xg, wg)
    k = _height_direction(funcs)
    if k is None:
        if depth >= 16:  # give up: low-order, sign test per point
            pts, w = _tensor_gauss(lo, hi, xg, wg)
            ok = np.ones(len(w), dtype=bool)
            for c, s in zip(funcs, signs):
                if s != 0:
                    vals = np.array([_eval(c, (p - lo) / (hi - lo)) for p in pts])
                    ok &= s * vals > 0
            return pts[ok], w[ok]
        e = int(np.argmax(hi - lo))
        mid = 0.5 * (lo[e] + hi[e])
        halves = [_split(c, e) for c in funcs]
        hi0, lo1 = hi.copy(), lo.copy()
        hi0[e], lo1[e] = mid, mid
        p0, w0 = _volume([h[0] for h in halves], signs, lo, hi0, xg, wg, depth + 1)
        p1, w1 = _volume([h[1] for h in halves], signs, lo1, hi, xg, wg, depth + 1)
        return np.concatenate([p0, p1]), np.concatenate([w0, w1])
    base_f, base_s = [], []
    for c, s in zip(funcs, signs):
        g = 1 if (_face(c, k, 1) - _face(c, k, 0)).min() > 0 else -1
        # the column over a base point meets {s psi > 0} iff s psi > 0 on the face where s psi is largest
        base_f += [_face(c, k, 0), _face(c, k, 1)]
        base_s += [s if s * g < 0 else 0, s if s * g > 0 else 0]
    rest = [e for e in range(d) if e != k]
    bp, bw = _volume(base_f, base_s, lo[rest], hi[rest], xg, wg, depth)
    return _line(funcs, signs, bp, bw, k, lo, hi, xg, wg)


def _line(funcs, signs, bp, bw, k, lo, hi, xg, wg):
    """1D Gauss rules in direction k over every base point, on the sub-intervals between the roots where all sign
    conditions hold."""
    d = len(lo)
    rest = [e for e in range(d) if e != k]
    pts, wts = [], []
    L = hi[k] - lo[k]
    for b, w in zip(bp, bw):
        tl = (b - lo[rest]) / (hi[rest] - lo[rest]) if d > 1 else np.zeros(0)
        ends = []
        for c in funcs:
            cc = np.moveaxis(c, k, -1)
            ends.append((_eval(cc[..., 0], tl), _eval(cc[..., 1], tl)))
        roots = [0.0, 1.0]
        for a, bb in ends:
            if a * bb < 0:
                roots.append(a / (a - bb))
        roots = sorted(roots)
        for r0, r1 in zip(roots[:-1], roots[1:]):
            if r1 - r0 <= 1e-14:
                continue
            m = 0.5 * (r0 + r1)
            if all(s == 0 or s * (a + (bb - a) * m) > 0 for (a, bb), s in zip(ends, signs)):
                t = lo[k] + L * (r0 + (r1 - r0) * xg)
                p = np.zeros((len(xg), d))
                p[:, rest] = b
                p[:, k] = t
                pts.append(p)
                wts.append(w * L * (r1 - r0) * wg)
    if not pts:
        return np.zeros((0, d)), np.zeros(0)
    return np.concatenate(pts), np.concatenate(wts)


def _surface(c, lo, hi, xg, wg, depth=0):
    """Quadrature of {psi = 0} inside the box: points, weights, unit normals (grad psi / |grad psi|, unit-cell coordinates)."""
    d = len(lo)
    if c.min() > 0 or c.max() < 0:
        return np.zeros((0, d)), np.zeros(0), np.zeros((0, d))
    if d == 1:
        a, b = float(c[0]), float(c[1])
        if a * b >= 0:
            return np.zeros((0, 1)), np.zeros(0), np.zeros((0, 1))
        t = lo[0] + (hi[0] - lo[0]) * a / (a - b)
        return np.array([[t]]), np.ones(1), np.array([[np.sign(b - a)]])
    k = _height_direction([c])
    if k is None:
        if depth >= 16:
            return np.zeros((0, d)), np.zeros(0), np.zeros((0, d))
        e = int(np.argmax(hi - lo))
        mid = 0.5 * (lo[e] + hi[e])
        c0, c1 = _split(c, e)
        hi0, lo1 = hi.copy(), lo.copy()
        hi0[e], lo1[e] = mid, mid
        a = _surface(c0, lo, hi0, xg, wg, depth + 1)
        b = _surface(c1, lo1, hi, xg, wg, depth + 1)
        return tuple(np.concatenate([x, y]) for x, y in zip(a, b))
    g = 1 if (_face(c, k, 1) - _face(c, k, 0)).min() > 0 else -1
    rest = [e for e in range(d) if e != k]
    bp, bw = _volume([_face(c, k, 0), _face(c, k, 1)], [-g, g], lo[rest], hi[rest], xg, wg)
    pts, wts, nrm = [], [], []
    size = hi - lo
    for b, w in zip(bp, bw):
        tl = (b - lo[rest]) / (hi[rest] - lo[rest])
        cc = np.moveaxis(c, k, -1)
        a, bb = _eval(cc[..., 0], tl), _eval(cc[..., 1], tl)
        if a * bb >= 0:
            continue
        r = a / (a - bb)
        t = np.zeros(d)
        t[rest] = tl
        t[k] = r
        grad = np.zeros(d)
        for e in range(d):
            ce = np.moveaxis(c, e, 0)
            te = np.delete(t, e)
            grad[e] = (_eval(ce[1], te) - _eval(ce[0], te)) / size[e]
        gn = np.linalg.norm(grad)
        p = lo + size * t
        pts.append(p)
        wts.append(w * gn / abs(grad[k]))
        nrm.append(grad / gn)
    if not pts:
        return np.zeros((0, d)), np.zeros(0), np.zeros((0, d))
    return np.array(pts), np.array(wts), np.array(nrm)


def cell_rules(system, ls, cell, n_gauss):
    """Inside and surface rules of a cell for either kind of level set (see `classify`)."""
    if hasattr(ls, "rules"):
        return ls.rules(system, cell, n_gauss)
    return cut_quadrature(cell_vertex_values(system, ls, cell), n_gauss)


def cut_quadrature(vertex_values, n_gauss):
    """Inside ({psi < 0}) and surface ({psi = 0}) quadratures on the unit cell for the multilinear function with the
    given vertex values (shape (2,)*dim).  Returns (points, weights), (points, weights, normals)."""
    c = np.asarray(vertex_values, dtype=float)
    d = c.ndim
    xg, wg = gauss_legendre_01(n_gauss)
    lo, hi = np.zeros(d), np.ones(d)
    return _volume([c], [-1], lo, hi, xg, wg), _surface(c, lo, hi, xg, wg)


# --------------------------------------------------------------------------------------------------- shape values
def shape_at_points(system, cell, ref_pts):
    """values[q, i], physical gradients[d][q, i] of the cell's GDM basis at unit-cell points (i lexicographic, x fastest)."""
    p, dim = system.fe_degree, system.dim
    idx = system.cell_indices(cell)
    h = system.h
    one_d = [basis_values(p, system.variant(idx[e], e), ref_pts[:, e], n_der=1) for e in range(dim)]  # [der][k][q]

    def tensor(fs):  # fs[e] = [k][q] -> [q, i] with x fastest in i
        out = fs[0].T
        for e in range(1, dim):
            out = np.einsum("qb,qa->qba", fs[e].T, out).reshape(len(ref_pts), -1)
        return out

    value = tensor([o[0] for o in one_d])
    grads = [tensor([one_d[e][1] / h[e] if e == dd else one_d[e][0] for e in range(dim)]) for dd in range(dim)]
    return value, grads


def physical_points(system, cell, ref_pts):
    idx = system.cell_indices(cell)
    return np.array(system.lo) + (np.array(idx) + ref_pts) * np.array(system.h)


def face_has_ghost_penalty(system, location, cell, d, side):
    """`prototypes/cut_poisson_01_gdm.cc:123-146` (= `wave/mass.h:86-106`): interior face between an intersected cell
    and a neighbour that is not outside."""
    idx = system.cell_indices(cell)
    nb = list(idx)
    nb[d] += 1 if side else -1
    if nb[d] < 0 or nb[d] >= system.n_subdivisions[d]:
        return None
    ncell = indices_to_index(nb, system.n_subdivisions)
    a, b = location[cell], location[ncell]
    if (a == INTERSECTED and b != OUTSIDE) or (b == INTERSECTED and a != OUTSIDE):
        return ncell
    return None


# -------------------------------------------------------------------------------------------------- the assembly
def assemble_cut_poisson(system, ls, ghost_penalty=True, ghost_parameter=0.5, nitsche_parameter=None,
                         rhs_value=4.0, boundary_value=1.0, gp_h_power=1, kind="stiffness", outside_diagonal=1.0,
                         surface_terms=True):
    """Global matrix (CSR), right-hand side and cell locations of the CutFEM Poisson problem.

    Follows `prototypes/cut_poisson_01_gdm.cc:196-329` term by term (scalar field, no constraints).
    `rhs_value` / `boundary_value` may be callables of the physical points [q, dim].
    `kind="mass"`: the cut mass matrix of `applications/wave/include/gdm/wave/mass.h:47-249` instead (inside mass, no
    surface terms, ghost penalty with `ghost_parameter` = gamma_M and `gp_h_power=3`); the right-hand side is then
    (v, rhs_value) on the inside part.  `outside_diagonal=0` leaves the rows no active cell touches empty, which is
    what the matrix-free residual `wave/stiffness.h:42-407` amounts to.  `surface_terms=False` drops the Nitsche terms
    on the cut surface (`function_interface_dbc` unset: the two-domain runs couple there instead).
    """
    assert system.n_components == 1 and kind in ("stiffness", "mass")
    fval = rhs_value if callable(rhs_value) else (lambda pts: np.full(len(pts), float(rhs_value)))
    gval = boundary_value if callable(boundary_value) else (lambda pts: np.full(len(pts), float(boundary_value)))
    p, dim = system.fe_degree, system.dim
    n = system.n_dofs()
    if nitsche_parameter is None:
        nitsche_parameter = 5.0 * (p + 1) * p
    location = classify(system, ls)
    get, jxw_full, xq = _cell_tables(system)
    hmin = min(system.h)
    h = np.array(system.h)
    rows, cols, vals = [], [], []
    rhs = np.zeros(n)
    inside_cache = {}
    xg, wg = gauss_legendre_01(p + 1)
    fpts, fw = _tensor_gauss(np.zeros(dim - 1), np.ones(dim - 1), xg, wg)

    def add(dofs_r, mat):
        dofs_r = np.asarray(dofs_r)
        rows.append(np.repeat(dofs_r, len(dofs_r)))
        cols.append(np.tile(dofs_r, len(dofs_r)))
        vals.append(mat.ravel())

    for cell in range(system.n_cells()):
        if location[cell] == OUTSIDE:
            continue
        dofs = system.get_dof_indices(cell)
        if location[cell] == INSIDE:
            cat = system.active_fe_index(cell)
            if cat not in inside_cache:
                idx = system.cell_indices(cell)
                value, _ = get([system.variant(idx[e], e) for e in range(dim)])
                inside_cache[cat] = (cell_matrix_scalar(system, cell, kind, get, jxw_full), value)
            km, value = inside_cache[cat]
            add(dofs, km)
            grids = np.meshgrid(*([xq] * dim)[::-1], indexing="ij")
            ref = np.stack([g.ravel() for g in grids[::-1]], axis=1)  # the point order of _cell_tables (x fastest)
            rhs[dofs] += np.einsum("q,q,qi->i", jxw_full, fval(physical_points(system, cell, ref)), value)
        else:
            (ip, iw), (sp_, sw, sn) = cell_rules(system, ls, cell, p + 1)
            local = np.zeros((len(dofs), len(dofs)))
            lrhs = np.zeros(len(dofs))
            if len(iw):
                value, grads = shape_at_points(system, cell, ip)
                jxw = iw * float(np.prod(h))
                if kind == "mass":
                    local += np.einsum("q,qi,qj->ij", jxw, value, value)
                else:
                    for g in grads:
                        local += np.einsum("q,qi,qj->ij", jxw, g, g)
                lrhs += np.einsum("q,q,qi->i", jxw, fval(physical_points(system, cell, ip)), value)
            if len(sw) and kind == "stiffness" and surface_terms:
                value, grads = shape_at_points(system, cell, sp_)
                nphys = sn / h
                scale = np.linalg.norm(nphys, axis=1)
                nphys = nphys / scale[:, None]
                jxw = sw * float(np.prod(h)) * scale
                ng = sum(nphys[:, e][:, None] * grads[e] for e in range(dim))  # normal . grad phi_i
                local += -np.einsum("q,qi,qj->ij", jxw, ng, value) - np.einsum("q,qi,qj->ij", jxw, value, ng) \
                    + nitsche_parameter / hmin * np.einsum("q,qi,qj->ij", jxw, value, value)
                gq = gval(physical_points(system, cell, sp_))
                lrhs += nitsche_parameter / hmin * np.einsum("q,q,qi->i", jxw, gq, value) \
                    - np.einsum("q,q,qi->i", jxw, gq, ng)
            add(dofs, local)
            rhs[dofs] += lrhs
        if ghost_penalty:
            for d in range(dim):
                for side in (0, 1):
                    ncell = face_has_ghost_penalty(system, location, cell, d, side)
                    if ncell is None:
                        continue
                    here = np.insert(fpts, d, float(side), axis=1)
                    there = np.insert(fpts, d, float(1 - side), axis=1)
                    _, g_here = shape_at_points(system, cell, here)
                    _, g_there = shape_at_points(system, ncell, there)
                    jump = np.concatenate([g_here[d], -g_there[d]], axis=1)  # (n . n) = 1 on a Cartesian face
                    jxw = fw * float(np.prod(np.delete(h, d)))
                    stab = 0.5 * ghost_parameter * hmin ** gp_h_power * np.einsum("q,qi,qj->ij", jxw, jump, jump)
                    add(list(dofs) + list(system.get_dof_indices(ncell)), stab)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    A.sum_duplicates()
    diag = A.diagonal()
    fix = np.where(diag == 0.0)[0]  # `:324-329`: rows no active cell touches become identity rows
    A = (A + sp.coo_matrix((np.full(len(fix), float(outside_diagonal)), (fix, fix)), shape=(n, n))).tocsr()
    return A, rhs, location


def error_norms_inside(system, ls, u, exact, location=None):
    """(L2, L1, Linf) of u_h - u over the inside part, Linf as the maximum over the quadrature points: the three columns
    `applications/wave/include/gdm/wave/problem.h:609-615` prints (`postprocess`, `:531-607`)."""
    return _errors_inside(system, ls, u, exact, location)


def l2_error_inside(system, ls, u, exact, location=None):
    return _errors_inside(system, ls, u, exact, location)[0]


def _errors_inside(system, ls, u, exact, location=None):
    """sqrt(sum over non-outside cells of int_{inside part} (u_h - u)^2) (`prototypes/cut_poisson_01_gdm.cc:349-398`)."""
    p, dim = system.fe_degree, system.dim
    if location is None:
        location = classify(system, ls)
    get, jxw_full, xq = _cell_tables(system)
    full_ref, _ = _tensor_gauss(np.zeros(dim), np.ones(dim), xq, xq)
    vol = float(np.prod(system.h))
    acc = l1 = linf = 0.0
    u = np.asarray(u, dtype=float)
    for cell in range(system.n_cells()):
        if location[cell] == OUTSIDE:
            continue
        dofs = system.get_dof_indices(cell)
        if location[cell] == INSIDE:
            idx = system.cell_indices(cell)
            value, _ = get([system.variant(idx[e], e) for e in range(dim)])
            # _cell_tables orders points x fastest; build matching reference points
            grids = np.meshgrid(*([xq] * dim)[::-1], indexing="ij")
            ref = np.stack([g.ravel() for g in grids[::-1]], axis=1)
            jxw = jxw_full
        else:
            (ref, w), _ = cell_rules(system, ls, cell, p + 1)
            if not len(w):
                continue
            value, _ = shape_at_points(system, cell, ref)
            jxw = w * vol
        pts = physical_points(system, cell, ref)
        diff = value @ u[dofs] - exact(pts)
        acc += float(np.sum(diff ** 2 * jxw))
        l1 += float(np.sum(np.abs(diff) * jxw))
        linf = max(linf, float(np.abs(diff).max()))
    return float(np.sqrt(acc)), l1, linf


def load_functionals(system, ls, nitsche_parameter=None, location=None):
    """The data-dependent part of the residual `wave/stiffness.h:186-260` as two lists of (dofs, points, W[q, i]):
    volume  b_i += sum_q f(x_q) phi_i JxW  over the inside part, and
    surface b_i += sum_q g(x_q) (gamma_D / h phi_i - d_n phi_i) JxW  on the cut surface,
    so that a time-dependent f or g costs one evaluation per quadrature point and stage (`apply_load`)."""
    p, dim = system.fe_degree, system.dim
    if nitsche_parameter is None:
        nitsche_parameter = 5.0 * (p + 1) * p
    if location is None:
        location = classify(system, ls)
    get, jxw_full, xq = _cell_tables(system)
    h = np.array(system.h)
    hmin, vol = float(h.min()), float(np.prod(h))
    grids = np.meshgrid(*([xq] * dim)[::-1], indexing="ij")
    full_ref = np.stack([g.ravel() for g in grids[::-1]], axis=1)
    volume, surface = [], []
    for cell in range(system.n_cells()):
        if location[cell] == OUTSIDE:
            continue
        dofs = np.asarray(system.get_dof_indices(cell))
        if location[cell] == INSIDE:
            idx = system.cell_indices(cell)
            value, _ = get([system.variant(idx[e], e) for e in range(dim)])
            volume.append((dofs, physical_points(system, cell, full_ref), value * jxw_full[:, None]))
            continue
        (ip, iw), (sp_, sw, sn) = cell_rules(system, ls, cell, p + 1)
        if len(iw):
            value, _ = shape_at_points(system, cell, ip)
            volume.append((dofs, physical_points(system, cell, ip), value * (iw * vol)[:, None]))
        if len(sw):
            value, grads = shape_at_points(system, cell, sp_)
            nphys = sn / h
            scale = np.linalg.norm(nphys, axis=1)
            nphys = nphys / scale[:, None]
            jxw = sw * vol * scale
            ng = sum(nphys[:, e][:, None] * grads[e] for e in range(dim))
            surface.append((dofs, physical_points(system, cell, sp_), (nitsche_parameter / hmin * value - ng) * jxw[:, None]))
    return volume, surface


def apply_load(n, terms, fn):
    b = np.zeros(n)
    for dofs, pts, W in terms:
        np.add.at(b, dofs, np.asarray(fn(pts), dtype=float) @ W)
    return b


# ---------------------------------------------------------------- two-domain (composite) runs of applications/wave
def _boundary_face_rules(system, ls, cell, n_gauss):
    """[(d, side, unit-cell points, weights)] for the faces of `cell` on the box boundary: the part of the face where
    the level set is negative (`NonMatching::FEInterfaceValues::reinit(cell, f)`, `wave/stiffness.h:268-283`)."""
    dim = system.dim
    idx = system.cell_indices(cell)
    xg, wg = gauss_legendre_01(n_gauss)
    c = cell_vertex_values(system, ls, cell)
    out = []
    for d in range(dim):
        for side in (0, 1):
            if idx[d] != (0 if side == 0 else system.n_subdivisions[d] - 1):
                continue
            fc = _face(c, d, side)
            if dim == 1:
                pts, w = (np.zeros((1, 0)), np.ones(1)) if float(fc) < 0 else (np.zeros((0, 0)), np.zeros(0))
            else:
                pts, w = _volume([fc], [-1], np.zeros(dim - 1), np.ones(dim - 1), xg, wg)
            if len(w):
                out.append((d, side, np.insert(pts, d, float(side), axis=1), w))
    return out


def domain_boundary_terms(system, ls, nitsche_parameter):
    """Nitsche terms on the box boundary for the domain {level set < 0} (`wave/stiffness.h:262-340`, term IV): the
    matrix  -<d_n v, u> - <v, d_n u> + gamma_D / h <v, u>  and the load functional  <gamma_D / h v - d_n v, g>."""
    dim, n = system.dim, system.n_dofs()
    h = np.array(system.h)
    hmin = float(h.min())
    location = classify(system, ls)
    rows, cols, vals, load = [], [], [], []
    for cell in range(system.n_cells()):
        if location[cell] == OUTSIDE:
            continue
        for d, side, pts, w in _boundary_face_rules(system, ls, cell, system.fe_degree + 1):
            dofs = np.asarray(system.get_dof_indices(cell))
            value, grads = shape_at_points(system, cell, pts)
            jxw = w * float(np.prod(np.delete(h, d)))
            ng = (1.0 if side else -1.0) * grads[d]
            local = -np.einsum("q,qi,qj->ij", jxw, ng, value) - np.einsum("q,qi,qj->ij", jxw, value, ng) \
                + nitsche_parameter / hmin * np.einsum("q,qi,qj->ij", jxw, value, value)
            rows.append(np.repeat(dofs, len(dofs)))
            cols.append(np.tile(dofs, len(dofs)))
            vals.append(local.ravel())
            load.append((dofs, physical_points(system, cell, pts), (nitsche_parameter / hmin * value - ng) * jxw[:, None]))
    if not vals:
        return sp.csr_matrix((n, n)), load
    B = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    return B, load


def coupling_matrices(system, ls):
    """P = sum_q (n . grad phi_i) phi_j JxW and Q = sum_q phi_i phi_j JxW over the cut surface (n = normal of the level
    set): the interface coupling of the two-domain residual (`wave/stiffness.h:441-574`) is
        r0 -= -1/2 P [u] - 1/2 P^T (u0 + u1) + tau / h Q [u],   r1 -= -1/2 P [u] + 1/2 P^T (u0 + u1) - tau / h Q [u]
    with [u] = u0 - u1 and tau = gamma_D / 2."""
    dim, n, p = system.dim, system.n_dofs(), system.fe_degree
    h = np.array(system.h)
    vol = float(np.prod(h))
    location = classify(system, ls)
    rows, cols, pv, qv = [], [], [], []
    for cell in range(system.n_cells()):
        if location[cell] != INTERSECTED:
            continue
        _, (sp_, sw, sn) = cell_rules(system, ls, cell, p + 1)
        if not len(sw):
            continue
        dofs = np.asarray(system.get_dof_indices(cell))
        value, grads = shape_at_points(system, cell, sp_)
        nphys = sn / h
        scale = np.linalg.norm(nphys, axis=1)
        nphys = nphys / scale[:, None]
        jxw = sw * vol * scale
        ng = sum(nphys[:, e][:, None] * grads[e] for e in range(dim))
        rows.append(np.repeat(dofs, len(dofs)))
        cols.append(np.tile(dofs, len(dofs)))
        pv.append(np.einsum("q,qi,qj->ij", jxw, ng, value).ravel())
        qv.append(np.einsum("q,qi,qj->ij", jxw, value, value).ravel())
    r, c = np.concatenate(rows), np.concatenate(cols)
    return (sp.coo_matrix((np.concatenate(pv), (r, c)), shape=(n, n)).tocsr(),
            sp.coo_matrix((np.concatenate(qv), (r, c)), shape=(n, n)).tocsr())
