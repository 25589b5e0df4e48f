"""GDM::VectorTools restatement (oracle, test-only).

`interpolate` (`include/gdm/vector_tools.h:11-23`): nodal values through the
lexicographically renumbered Q1 DoFHandler (`system.h:255-335,795-797`), i.e.
u[i*nc + c] = f(x_i, c).  `integrate_difference` (`vector_tools.h:25-86`):
per cell sqrt(sum_q sum_c (u_h - u)^2 JxW) with the GDM basis of the cell's
category; `compute_global_error` (deal.II) = sqrt(sum cell^2) for L2.
"""
import numpy as np

from .assemble import _cell_tables, _quadrature_points


def interpolate(system, f):
    """f(points[n, dim], comp) -> values[n]."""
    pts = system.node_coordinates()
    nc = system.n_components
    out = np.zeros(system.n_dofs())
    for c in range(nc):
        out[c::nc] = np.asarray(f(pts, c), dtype=float) * np.ones(len(pts))
    return out


def integrate_difference(system, u, exact):
    """Cell-wise L2 error vector (length n_cells)."""
    nc = system.n_components
    get, jxw, xq = _cell_tables(system)
    npc = (system.fe_degree + 1) ** system.dim
    diff = np.zeros(system.n_cells())
    u = np.asarray(u, dtype=float)
    for cell in range(system.n_cells()):
        idx = system.cell_indices(cell)
        value, _ = get([system.variant(idx[d], d) for d in range(system.dim)])
        dofs = np.asarray(system.get_dof_indices(cell))
        pts = _quadrature_points(system, cell, xq)
        acc = 0.0
        for c in range(nc):
            uh = value @ u[dofs[c * npc:(c + 1) * npc]]
            ue = np.asarray(exact(pts, c), dtype=float) * np.ones(len(jxw))
            acc += float(np.sum((uh - ue) ** 2 * jxw))
        diff[cell] = np.sqrt(acc)
    return diff


def compute_global_error(cellwise):
    return float(np.sqrt(np.sum(np.asarray(cellwise) ** 2)))
