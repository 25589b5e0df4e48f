"""GDM::System restatement (oracle, test-only).

Follows `include/gdm/system.h`:
  * cell -> DoF window         `CellAccessor::get_dof_indices`  system.h:195-246
  * cell categories            `System::categorize`             system.h:404-424
  * periodic constraints       `make_periodicity_constraints`   system.h:427-463
  * zero Dirichlet constraints `make_zero_boundary_constraints` system.h:466-508
  * slab partition             `create_triangulation_pre`       system.h:720-757
plus the deal.II `AffineConstraints` behaviour the reference relies on
(`distribute_local_to_global`, `distribute`, `set_zero`), restated from the
deal.II documentation/implementation (not in the reference tree).
"""
import numpy as np


def index_to_indices(index, Ns):
    """Lexicographic -> multi index, x fastest (`fe.h:339-356`)."""
    dim = len(Ns)
    out = [0] * dim
    out[0] = index % Ns[0]
    if dim >= 2:
        out[1] = (index // Ns[0]) % Ns[1]
    if dim >= 3:
        out[2] = index // (Ns[0] * Ns[1])
    return out


def indices_to_index(indices, Ns):
    """Multi index -> lexicographic (`fe.h:371-386`)."""
    idx, stride = 0, 1
    for d, i in enumerate(indices):
        idx += i * stride
        stride *= Ns[d]
    return idx


class Constraints:
    """Minimal `dealii::AffineConstraints<double>`: lines dof -> ([(master, w)], inhomogeneity)."""

    def __init__(self):
        self.lines = {}

    def is_constrained(self, i):
        return i in self.lines

    def constrain_dof_to_zero(self, i):
        self.lines[i] = ([], 0.0)

    def add_line(self, i):
        self.lines.setdefault(i, ([], 0.0))

    def add_entry(self, i, j, w):
        self.lines[i][0].append((j, w))

    def add_constraint(self, i, entries, inhomogeneity):
        self.lines[i] = (list(entries), float(inhomogeneity))

    def close(self):
        # resolve chains (slave -> slave); only weight-1 periodic chains occur here
        changed = True
        while changed:
            changed = False
            for i, (entries, inh) in list(self.lines.items()):
                new, new_inh = [], inh
                for (j, w) in entries:
                    if j in self.lines:
                        ej, ij = self.lines[j]
                        new.extend((k, w * wk) for (k, wk) in ej)
                        new_inh += w * ij
                        changed = True
                    else:
                        new.append((j, w))
                self.lines[i] = (new, new_inh)

    def distribute(self, vec):
        """`AffineConstraints::distribute`: slave = sum w*master + inhomogeneity."""
        for i, (entries, inh) in self.lines.items():
            vec[i] = inh + sum(w * vec[j] for (j, w) in entries)
        return vec

    def set_zero(self, vec):
        for i in self.lines:
            vec[i] = 0.0
        return vec

    def constrained_mask(self, n):
        m = np.zeros(n, dtype=bool)
        for i in self.lines:
            m[i] = True
        return m

    def distribute_local_to_global(self, cell_matrix, cell_vector, dof_indices, trip, rhs):
        """deal.II `distribute_local_to_global(cell_matrix, cell_vector, dofs, A, b)`.

        `trip` is a dict {(row, col): value} accumulating the global matrix.
        Constrained rows/cols are redirected to their masters (weights); a
        constrained DoF keeps a positive diagonal: |local(i,i)| (or the mean
        |diagonal| of the cell matrix when that is zero) -- this is deal.II's
        convention, no reference golden depends on its value.
        """
        n = len(dof_indices)
        targets = []
        for g in dof_indices:
            if g in self.lines:
                targets.append(self.lines[g][0])
            else:
                targets.append([(g, 1.0)])
        if cell_matrix is not None:
            avg_diag = float(np.mean(np.abs(np.diag(cell_matrix))))
            for i in range(n):
                gi = dof_indices[i]
                if gi in self.lines:
                    d = abs(cell_matrix[i, i])
                    if d == 0.0:
                        d = avg_diag
                    trip[(gi, gi)] = trip.get((gi, gi), 0.0) + d
                    if rhs is not None and self.lines[gi][1] != 0.0:
                        rhs[gi] += d * self.lines[gi][1]
                for (ti, wi) in targets[i]:
                    for j in range(n):
                        v = cell_matrix[i, j]
                        if v == 0.0:
                            continue
                        gj = dof_indices[j]
                        for (tj, wj) in targets[j]:
                            key = (ti, tj)
                            trip[key] = trip.get(key, 0.0) + wi * wj * v
                        if gj in self.lines and rhs is not None and self.lines[gj][1] != 0.0:
                            rhs[ti] -= wi * v * self.lines[gj][1]
        if cell_vector is not None and rhs is not None:
            for i in range(n):
                for (ti, wi) in targets[i]:
                    rhs[ti] += wi * cell_vector[i]


class System:
    """`GDM::System<dim>` on a Cartesian grid (`system.h:339-827`)."""

    def __init__(self, dim, fe_degree, n_components=1, rank=0, n_ranks=1, add_ghost_layer=False):
        assert fe_degree % 2 == 1
        self.dim = dim
        self.fe_degree = fe_degree
        self.n_components = n_components
        self.rank, self.n_ranks = rank, n_ranks
        self.add_ghost_layer = add_ghost_layer
        self.n_subdivisions = None
        self.lo = self.hi = None

    # ---- geometry (`system.h:367-401`)
    def subdivided_hyper_cube(self, n, left=0.0, right=1.0):
        self.subdivided_hyper_rectangle([n] * self.dim, [left] * self.dim, [right] * self.dim)

    def subdivided_hyper_rectangle(self, repetitions, p1, p2):
        assert len(repetitions) == self.dim
        self.n_subdivisions = [int(r) for r in repetitions]
        assert all(n >= self.fe_degree for n in self.n_subdivisions)
        self.lo = [float(x) for x in p1]
        self.hi = [float(x) for x in p2]

    @property
    def h(self):
        return [(self.hi[d] - self.lo[d]) / self.n_subdivisions[d] for d in range(self.dim)]

    @property
    def n_nodes(self):
        return [n + 1 for n in self.n_subdivisions]

    def n_dofs(self):
        return self.n_components * int(np.prod(self.n_nodes))

    def n_cells(self):
        return int(np.prod(self.n_subdivisions))

    # ---- per-direction window/variant (`system.h:209-216`, `415-420`)
    def window_offset(self, c, d):
        p, N = self.fe_degree, self.n_subdivisions[d]
        return 0 if c < p // 2 else min(N, c + p // 2 + 1) - p

    def variant(self, c, d):
        p, N = self.fe_degree, self.n_subdivisions[d]
        if c < p // 2:
            return c
        if c < N - p // 2:
            return p // 2
        return p + c - N

    def cell_indices(self, cell):
        return index_to_indices(cell, self.n_subdivisions)

    def active_fe_index(self, cell):
        """`categorize` (`system.h:404-424`): sum variant_d * p^d."""
        idx = self.cell_indices(cell)
        v = [self.variant(idx[d], d) for d in range(self.dim)]
        return indices_to_index(v, [self.fe_degree] * self.dim)

    def get_dof_indices(self, cell):
        """Global DoFs of a cell, cell-local order comp-major then lexicographic (x fastest).

        `system.h:195-246`; the FESystem numbers interior DoFs component by
        component (`component_to_system_index(comp, c) = comp*(p+1)^dim + c`).
        """
        p, nc = self.fe_degree, self.n_components
        idx = self.cell_indices(cell)
        off = [self.window_offset(idx[d], d) for d in range(self.dim)]
        nn = self.n_nodes
        scalar = []
        rng = [range(p + 1) if d < self.dim else range(1) for d in range(3)]
        for k in rng[2]:
            for j in rng[1]:
                for i in rng[0]:
                    o = [off[0] + i] + ([off[1] + j] if self.dim >= 2 else []) + \
                        ([off[2] + k] if self.dim >= 3 else [])
                    scalar.append(indices_to_index(o, nn))
        out = []
        for comp in range(nc):
            out.extend(s * nc + comp for s in scalar)
        return out

    # ---- constraints
    def make_zero_boundary_constraints(self, constraints, surface=None):
        """`system.h:466-508`; surface = 2*d + side, or all faces."""
        surfaces = range(2 * self.dim) if surface is None else [surface]
        nn, nc = self.n_nodes, self.n_components
        for s in surfaces:
            d, side = s // 2, s % 2
            n0 = int(np.prod(nn[d + 1:])) if d + 1 < self.dim else 1
            n1 = int(np.prod(nn[:d])) if d > 0 else 1
            n2 = n1 * nn[d]
            for i in range(n0):
                for j in range(n1):
                    i0 = i * n2 + (0 if side == 0 else self.n_subdivisions[d]) * n1 + j
                    for c in range(nc):
                        constraints.constrain_dof_to_zero(i0 * nc + c)

    def make_periodicity_constraints(self, d, constraints):
        """`system.h:427-463`: node N_d == node 0 in direction d, weight 1."""
        nn, nc = self.n_nodes, self.n_components
        n0 = int(np.prod(nn[d + 1:])) if d + 1 < self.dim else 1
        n1 = int(np.prod(nn[:d])) if d > 0 else 1
        n2 = n1 * nn[d]
        for i in range(n0):
            for j in range(n1):
                i0 = i * n2 + j
                i1 = i0 + self.n_subdivisions[d] * n1
                for c in range(nc):
                    if constraints.is_constrained(i1 * nc + c):
                        continue
                    constraints.add_line(i1 * nc + c)
                    constraints.add_entry(i1 * nc + c, i0 * nc + c, 1.0)

    # ---- sparsity patterns (`system.h:586-630`), literal cell loops
    def create_sparsity_pattern(self, constraints=None, flux=False):
        """[row] -> sorted list of columns.  `create_sparsity_pattern` couples the DoFs of every cell with each other;
        `create_flux_sparsity_pattern` (flux=True) adds, for every interior face, the DoFs of the cell (rows) x the
        DoFs of the face neighbour (columns).  Constrained rows/columns are kept (deal.II's
        `add_entries_local_to_global` default `keep_constrained_entries = true`); constraints with entries (periodic)
        would add the redirected pairs as well -- not modelled (no caller on the path uses them with a flux pattern)."""
        rows = [set() for _ in range(self.n_dofs())]
        ns = self.n_subdivisions
        for cell in range(self.n_cells()):
            dofs = self.get_dof_indices(cell)
            for i in dofs:
                rows[i].update(dofs)
            if flux:
                idx = self.cell_indices(cell)
                for d in range(self.dim):
                    for s in (-1, 1):
                        nb = list(idx)
                        nb[d] += s
                        if nb[d] < 0 or nb[d] >= ns[d]:
                            continue  # at_boundary
                        nd = self.get_dof_indices(indices_to_index(nb, ns))
                        for i in dofs:
                            rows[i].update(nd)
        return [sorted(r) for r in rows]

    # ---- slab partition (`system.h:720-757`)
    def partition_stride(self):
        return (self.n_subdivisions[-1] + self.n_ranks - 1) // self.n_ranks

    def owned_plane_range(self, rank=None):
        """Node planes (last coordinate) owned by `rank`: [start, end)."""
        rank = self.rank if rank is None else rank
        stride = self.partition_stride()
        n_last = self.n_subdivisions[-1] + 1
        start = 0 if rank == 0 else stride * rank + 1
        end = stride * (rank + 1) + 1
        return min(start, n_last), min(end, n_last)

    def locally_owned_range(self, rank=None):
        face = self.n_components * int(np.prod(self.n_nodes[:-1])) if self.dim > 1 else self.n_components
        a, b = self.owned_plane_range(rank)
        return face * a, face * b

    def cell_owner(self, cell):
        return self.cell_indices(cell)[-1] // self.partition_stride()

    def node_coordinates(self):
        """[n_scalar_dofs, dim] coordinates of the grid nodes in DoF order."""
        nn = self.n_nodes
        axes = [self.lo[d] + np.arange(nn[d]) * self.h[d] for d in range(self.dim)]
        grids = np.meshgrid(*axes[::-1], indexing="ij")  # slowest first
        return np.stack([g.ravel() for g in grids[::-1]], axis=1)
