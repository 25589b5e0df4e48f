"""Host-side mirror of the reference's `include/gdm` API over the C ABI.

Names, argument order and error behaviour follow the reference so that parity
tests read like its own tests (`tests/poisson_01_gdm.cc`, `tests/mass_01_gdm.cc`,
`prototypes/advection_01_gdm.cc`):

    system = System(dim, fe_degree, n_components)            # GDM::System<dim>  system.h:343-364
    system.subdivided_hyper_cube(n)                           # system.h:367-382
    constraints = AffineConstraints()
    system.make_zero_boundary_constraints(constraints)        # system.h:502-508
    constraints.close(); system.categorize()
    A = SparseMatrix(); MatrixCreator.create_laplace_matrix(mapping, system, quadrature, A, constraints)
    SolverCG(ReductionControl(100, 1e-10, 1e-4)).solve(A, x, b, PreconditionIdentity())

deal.II types that only carry configuration here (`MappingQ1`, `QGauss`) are
checked and otherwise ignored: the kernels implement exactly MappingQ1 on a
Cartesian grid with QGauss(fe_degree+1), which is what every call site uses.
All compute happens in libgdm_b200.so on the GPU; nothing here falls back to
the CPU.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import (ExcNotImplemented, GdmError, NoConvergence)  # noqa: F401

__all__ = [
    "Context", "default_context", "System", "AffineConstraints", "Vector", "SparseMatrix", "MatrixCreator",
    "VectorTools", "SolverCG", "ReductionControl", "PreconditionIdentity", "PreconditionJacobi",
    "DiagonalMatrix", "TimeStepping", "DiscreteTime", "MappingQ1", "QGauss", "generate_polynomials_1D",
    "ExcNotImplemented", "NoConvergence", "GdmError", "init_distributed", "capi", "CutPoisson",
]


# ------------------------------------------------------------------ context
class Context:
    """One per GPU (cudaStream_t optional)."""

    def __init__(self, device=0, stream=None):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(self.lib.gdm_context_create(int(device), C.c_void_p(stream or 0), C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.n_ranks = 0, 1

    def set_stream(self, stream):
        capi.check(self.lib.gdm_context_set_stream(self.h, C.c_void_p(stream or 0)))

    def synchronize(self):
        capi.check(self.lib.gdm_context_synchronize(self.h))

    def launch_count(self):
        n = C.c_uint64()
        capi.check(self.lib.gdm_context_launch_count(self.h, C.byref(n)))
        return n.value

    def comm_init(self, unique_id: bytes, rank, n_ranks):
        buf = C.create_string_buffer(unique_id, 128)
        capi.check(self.lib.gdm_context_comm_init(self.h, buf, rank, n_ranks))
        self.rank, self.n_ranks = rank, n_ranks

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.gdm_context_destroy(self.h)
                self.h = None
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = _default_ctx.get("current", 0)
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    _default_ctx["current"] = device
    return _default_ctx[device]


def init_distributed(device=None):
    """One process per GPU: bootstrap the NCCL communicator of the context through torch.distributed.

    Reads RANK / WORLD_SIZE / LOCAL_RANK (torchrun).  torch.distributed is only the
    rendezvous (broadcast of the 128-byte ncclUniqueId); the data path is the library's own.
    """
    import os
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    device = local if device is None else device
    torch.cuda.set_device(device)
    ctx = default_context(device)
    if world > 1 and ctx.n_ranks != world:
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        lib = capi.load()
        buf = C.create_string_buffer(128)
        if rank == 0:
            capi.check(lib.gdm_comm_unique_id(buf))
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        ctx.comm_init(obj[0], rank, world)
    return ctx


# ----------------------------------------------------- configuration carriers
class MappingQ1:
    """dealii::MappingQ1 (the only mapping the GPU path implements)."""


class QGauss:
    """dealii::QGauss<dim>(n); the kernels integrate with n = fe_degree + 1 (exact on Cartesian cells)."""

    def __init__(self, n_points):
        self.n_points = int(n_points)


def _check_mapping_quadrature(system, mapping, quadrature):
    for m in (mapping if isinstance(mapping, (list, tuple)) else [mapping]):
        if m is not None and not isinstance(m, MappingQ1):
            raise ExcNotImplemented(capi.ERR_NOT_IMPLEMENTED, "only MappingQ1 is implemented")
    for q in (quadrature if isinstance(quadrature, (list, tuple)) else [quadrature]):
        if q is not None and q.n_points != system.fe_degree + 1:
            raise ExcNotImplemented(capi.ERR_NOT_IMPLEMENTED, "only QGauss(fe_degree + 1) is implemented")


def generate_polynomials_1D(fe_degree):
    """GDM::generate_polynomials_1D (fe.h:55-336): [variant][basis] -> coefficients, lowest power first."""
    lib = capi.load()
    p = int(fe_degree)
    if p < 1 or p > 9 or p % 2 == 0:
        raise ExcNotImplemented(capi.ERR_NOT_IMPLEMENTED, "fe_degree must be odd and <= 9")
    out = np.zeros((p, p + 1, p + 1))
    capi.check(lib.gdm_polynomials_1d(p, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


# ------------------------------------------------------------------- system
class System:
    """GDM::System<dim> (include/gdm/system.h:339-827)."""

    def __init__(self, dim, fe_degree, n_components=1, add_ghost_layer=False, comm=None, context=None):
        self.dim, self.fe_degree, self.n_components = int(dim), int(fe_degree), int(n_components)
        self.add_ghost_layer = bool(add_ghost_layer)
        self.ctx = context or default_context()
        self.lib = self.ctx.lib
        # comm: None = serial System(fe_degree, n_components); "world" = System(MPI_COMM_WORLD, ...)
        self.rank, self.n_ranks = (self.ctx.rank, self.ctx.n_ranks) if comm is not None else (0, 1)
        self.h = None
        self.n_subdivisions = None

    def subdivided_hyper_cube(self, n_subdivisions_1D, left=0.0, right=1.0):
        self.subdivided_hyper_rectangle([n_subdivisions_1D] * self.dim, [left] * self.dim, [right] * self.dim)

    def subdivided_hyper_rectangle(self, repetitions, p1, p2):
        assert len(repetitions) == self.dim
        d = capi.SystemDesc()
        d.dim, d.fe_degree, d.n_components = self.dim, self.fe_degree, self.n_components
        for i in range(3):
            d.n_subdivisions[i] = int(repetitions[i]) if i < self.dim else 0
            d.lo[i] = float(p1[i]) if i < self.dim else 0.0
            d.hi[i] = float(p2[i]) if i < self.dim else 1.0
        d.rank, d.n_ranks, d.add_ghost_layer = self.rank, self.n_ranks, int(self.add_ghost_layer)
        h = C.c_void_p()
        capi.check(self.lib.gdm_system_create(self.ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self.n_subdivisions = [int(r) for r in repetitions]
        self.lo, self.hi = [float(x) for x in p1], [float(x) for x in p2]

    def categorize(self):
        """system.h:404-424 -- categories are implicit in the kernels' band tables; kept for API parity."""

    def n_dofs(self):
        return int(self.lib.gdm_system_n_dofs(self.h))

    def n_cells(self):
        return int(self.lib.gdm_system_n_cells(self.h))

    def get_fe_degree(self):
        return self.fe_degree

    def locally_owned_dofs(self):
        b, e = C.c_uint64(), C.c_uint64()
        capi.check(self.lib.gdm_system_locally_owned_range(self.h, C.byref(b), C.byref(e)))
        return range(b.value, e.value)

    def n_locally_owned_dofs(self):
        r = self.locally_owned_dofs()
        return r.stop - r.start

    def get_dof_indices(self, cell):
        n = self.lib.gdm_system_dofs_per_cell(self.h)
        out = np.zeros(n, dtype=np.uint64)
        capi.check(self.lib.gdm_system_get_dof_indices(self.h, int(cell), out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def active_fe_index(self, cell):
        v = C.c_uint32()
        capi.check(self.lib.gdm_system_active_fe_index(self.h, int(cell), C.byref(v)))
        return v.value

    def sparsity_row(self, row, flux=False):
        """Columns of one row of `create_sparsity_pattern` / `create_flux_sparsity_pattern` (system.h:586-630), generated on
        demand (the pattern is never stored)."""
        n = C.c_uint64()
        capi.check(self.lib.gdm_system_sparsity_row(self.h, int(flux), int(row), None, 0, C.byref(n)))
        cols = (C.c_uint64 * max(n.value, 1))()
        capi.check(self.lib.gdm_system_sparsity_row(self.h, int(flux), int(row), cols, n.value, C.byref(n)))
        return list(cols[: n.value])

    def write_vtu(self, values, label, file_name):
        """Nodal field -> ASCII VTU (stand-in for `GDM::DataOut`, include/gdm/data_out.h): `values` = all DoFs (numpy)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.size == self.n_dofs()
        capi.check(self.lib.gdm_system_write_vtu(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), str(label).encode(),
                                                 str(file_name).encode()))

    def write_matrix_to_file(self, constraints, kind, file_name, write_binary_file=False, scale=1.0, b=(0.0, 0.0, 0.0),
                             constrained_diagonal=capi.DIAG_ASSEMBLED):
        """`write_matrix_to_file` of the reference's eigenvalue tool (applications/wave/wave-ev.cc:93-127): the assembled
        operator of `kind` (capi.OP_MASS, OP_STIFFNESS, ...) as "row column value" triplets in the iteration order of a
        deal.II SparseMatrix (text, or binary uint32/uint32/double records).  Host only.  Returns the number of entries."""
        d = capi.OperatorDesc()
        d.kind, d.scale, d.constrained_diagonal, d.kernel = kind, float(scale), constrained_diagonal, capi.KERNEL_GENERIC
        for i in range(3):
            d.b[i] = float(b[i]) if i < len(b) else 0.0
        ch = constraints._handle_for(self) if constraints is not None else None
        n = C.c_uint64()
        capi.check(self.lib.gdm_system_write_matrix(self.h, ch, C.byref(d), str(file_name).encode(), int(write_binary_file),
                                                    C.byref(n)))
        return n.value

    def matrix_1d(self, d, kind):
        p, n = self.fe_degree, self.n_subdivisions[d]
        band = np.zeros((n + 1, 2 * p + 1))
        capi.check(self.lib.gdm_system_matrix_1d(self.h, d, kind, band.ctypes.data_as(C.POINTER(C.c_double))))
        return band

    def layout(self):
        info = capi.LayoutInfo()
        capi.check(self.lib.gdm_system_layout(self.h, C.byref(info)))
        return info

    # constraints (system.h:427-508); argument order as in the reference
    def make_zero_boundary_constraints(self, *args):
        if len(args) == 1:
            surface, constraints = -1, args[0]
        else:
            surface, constraints = args
        constraints._bind(self)
        capi.check(self.lib.gdm_constraints_make_zero_boundary(constraints.h, int(surface)))

    def make_periodicity_constraints(self, d, constraints):
        constraints._bind(self)
        capi.check(self.lib.gdm_constraints_make_periodicity(constraints.h, int(d)))

    def interpolate_boundary_values(self, mapping, boundary_id, function, constraints):
        """system.h:511-547: constrain the boundary nodes (boundary id 0 = every face) to function(point, component)."""
        constraints._bind(self)

        def cb(pt, comp, _user):
            return float(function([pt[0], pt[1], pt[2]], comp))

        fn = capi.FUNCTION_FN(cb)
        capi.check(self.lib.gdm_constraints_interpolate_boundary_values(constraints.h, int(boundary_id), fn, None))

    def __del__(self):
        try:
            if self.h:
                self.lib.gdm_system_destroy(self.h)
                self.h = None
        except Exception:
            pass


class AffineConstraints:
    """dealii::AffineConstraints<double> restricted to what GDM::System generates."""

    def __init__(self, locally_active=None):
        self.h, self.system, self._closed = None, None, False

    def _bind(self, system):
        if self.h is None:
            h = C.c_void_p()
            capi.check(system.lib.gdm_constraints_create(system.h, C.byref(h)))
            self.h, self.system, self.lib = h, system, system.lib
        elif self.system is not system:
            raise GdmError(capi.ERR_INVALID, "constraints already bound to another system")

    def close(self):
        self._closed = True
        if self.h is not None:
            capi.check(self.lib.gdm_constraints_close(self.h))

    def n_constraints(self):
        return 0 if self.h is None else int(self.lib.gdm_constraints_n_constraints(self.h))

    def is_constrained(self, i):
        return False if self.h is None else bool(self.lib.gdm_constraints_is_constrained(self.h, int(i)))

    def distribute(self, vec):
        if self.h is not None:
            capi.check(self.lib.gdm_constraints_distribute(self.h, vec.h))

    def set_zero(self, vec):
        if self.h is not None:
            capi.check(self.lib.gdm_constraints_set_zero(self.h, vec.h))

    def condense_rhs(self, matrix, rhs):
        """The right-hand side part of `distribute_local_to_global` with inhomogeneous constraints
        (tests/poisson_02_gdm.cc:201): free rows b_i -= sum_j A_ij g_j, constrained rows b_j = diag_j g_j."""
        if self.h is not None:
            capi.check(self.lib.gdm_constraints_condense_rhs(self.h, matrix.h, rhs.h))

    def _handle_for(self, system):
        if self.h is None:
            return None
        if self.system is not system:
            raise GdmError(capi.ERR_INVALID, "constraints bound to another system")
        if not self._closed:
            raise GdmError(capi.ERR_INVALID, "constraints must be closed")
        return self.h

    def __del__(self):
        try:
            if self.h:
                self.lib.gdm_constraints_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ------------------------------------------------------------------ vectors
class Vector:
    """Device vector in the library's padded layout (LinearAlgebra::distributed::Vector<double>)."""

    def __init__(self, system, values=None):
        self.system, self.lib = system, system.lib
        h = C.c_void_p()
        capi.check(self.lib.gdm_vector_create(system.h, C.byref(h)))
        self.h = h
        if values is not None:
            self.upload(values)

    def reinit(self, other):
        return Vector(other.system)

    def size(self):
        return self.system.n_dofs()

    def upload(self, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        assert a.size == self.system.n_locally_owned_dofs(), (a.size, self.system.n_locally_owned_dofs())
        capi.check(self.lib.gdm_vector_upload(self.h, a.ctypes.data_as(C.c_void_p)))
        return self

    def numpy(self):
        out = np.empty(self.system.n_locally_owned_dofs())
        capi.check(self.lib.gdm_vector_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def device_ptr(self):
        return self.lib.gdm_vector_device_ptr(self.h)

    def set(self, value):
        capi.check(self.lib.gdm_vector_set(self.h, float(value)))
        return self

    def equ(self, other):  # this = other
        capi.check(self.lib.gdm_vector_copy(self.h, other.h))
        return self

    def copy(self):
        return Vector(self.system).equ(self)

    def scale(self, a):
        if isinstance(a, Vector):
            capi.check(self.lib.gdm_vector_scale_by(self.h, a.h))
        else:
            capi.check(self.lib.gdm_vector_scale(self.h, float(a)))
        return self

    def add(self, a, x):
        capi.check(self.lib.gdm_vector_add(self.h, float(a), x.h))
        return self

    def sadd(self, s, a, x):
        capi.check(self.lib.gdm_vector_sadd(self.h, float(s), float(a), x.h))
        return self

    def dot(self, other):
        r = C.c_double()
        capi.check(self.lib.gdm_vector_dot(self.h, other.h, C.byref(r)))
        return r.value

    __mul__ = dot

    def l2_norm(self):
        r = C.c_double()
        capi.check(self.lib.gdm_vector_l2_norm(self.h, C.byref(r)))
        return r.value

    def linfty_norm(self):
        r = C.c_double()
        capi.check(self.lib.gdm_vector_linfty_norm(self.h, C.byref(r)))
        return r.value

    def update_ghost_values(self):
        capi.check(self.lib.gdm_vector_update_ghost_values(self.h))

    def __del__(self):
        try:
            if self.h:
                self.lib.gdm_vector_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ---------------------------------------------------------------- operators
class SparseMatrix:
    """Stands where the reference has a SparseMatrix<double> / TrilinosWrappers::SparseMatrix.

    It is filled by `MatrixCreator.create_*` and then offers the deal.II operator
    concept (`vmult`, `vmult_add`, `m`, `n`); the "matrix" is never assembled: the GPU
    applies the Kronecker band structure directly.
    """

    def __init__(self):
        self.h, self.system = None, None

    def _create(self, system, constraints, kind, scale=1.0, b=(0.0, 0.0, 0.0), constrained_diagonal=capi.DIAG_ASSEMBLED,
                kernel=capi.KERNEL_AUTO):
        self._free()
        d = capi.OperatorDesc()
        d.kind, d.scale, d.constrained_diagonal, d.kernel = kind, float(scale), constrained_diagonal, kernel
        for i in range(3):
            d.b[i] = float(b[i]) if i < len(b) else 0.0
        ch = constraints._handle_for(system) if constraints is not None else None
        h = C.c_void_p()
        capi.check(system.lib.gdm_operator_create(system.h, ch, C.byref(d), C.byref(h)))
        self.h, self.system, self.lib = h, system, system.lib
        self._keep = constraints
        return self

    def m(self):
        return int(self.lib.gdm_operator_m(self.h))

    n = m

    def kernel_used(self):
        return int(self.lib.gdm_operator_kernel_used(self.h))

    def vmult(self, dst, src):
        capi.check(self.lib.gdm_operator_vmult(self.h, dst.h, src.h))

    def vmult_add(self, dst, src):
        capi.check(self.lib.gdm_operator_vmult_add(self.h, dst.h, src.h))

    def vmult_dot(self, dst, src):
        """dst = A src; returns <src, dst> (all ranks).  Fused path: the dot product rides in the store epilogue."""
        import ctypes as C
        out = C.c_double()
        capi.check(self.lib.gdm_operator_vmult_dot(self.h, dst.h, src.h, C.byref(out)))
        return out.value

    def Tvmult(self, dst, src):
        """dst = A^T src (`SparseMatrix::Tvmult`): vmult for mass/stiffness, the transposed operator for advection."""
        capi.check(self.lib.gdm_operator_tvmult(self.h, dst.h, src.h))

    def mass_inverse(self, dst, src):
        """dst = M^-1 src for a mass operator on a Cartesian grid: Kronecker-direct banded line solves
        instead of the CG / ILU / AMG mass solve of the reference's Runge-Kutta stages.  One rank."""
        capi.check(self.lib.gdm_operator_mass_inverse(self.h, dst.h, src.h))

    def vmult_host(self, dst, src):
        """vmult on HOST numpy buffers (H2D, apply, D2H)."""
        src = np.ascontiguousarray(src, dtype=np.float64)
        assert dst.flags["C_CONTIGUOUS"] and dst.dtype == np.float64
        capi.check(self.lib.gdm_operator_vmult_host(self.h, dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p)))

    def diagonal(self, vec=None):
        vec = vec or Vector(self.system)
        capi.check(self.lib.gdm_operator_diagonal(self.h, vec.h))
        return vec

    def attach_csr(self, row_ids, rowptr, col, val):
        row_ids = np.ascontiguousarray(row_ids, dtype=np.uint64)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        col = np.ascontiguousarray(col, dtype=np.uint64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        capi.check(self.lib.gdm_operator_attach_csr(self.h, len(row_ids), row_ids.ctypes.data_as(C.c_void_p),
                                                    rowptr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p),
                                                    val.ctypes.data_as(C.c_void_p)))

    def _free(self):
        if getattr(self, "h", None):
            self.lib.gdm_operator_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


class MatrixCreator:
    """GDM::MatrixCreator (include/gdm/matrix_creator.h) + the inline assembly loops of the tests."""

    @staticmethod
    def create_mass_matrix(mapping, system, quadrature, sparse_matrix, constraints, **kw):
        """matrix_creator.h:9-62"""
        _check_mapping_quadrature(system, mapping, quadrature)
        return sparse_matrix._create(system, constraints, capi.OP_MASS, **kw)

    @staticmethod
    def create_lumped_mass_matrix(mapping, system, quadrature, vector, constraints):
        """matrix_creator.h:64-117: vector <- 1 / (lumped mass)."""
        _check_mapping_quadrature(system, mapping, quadrature)
        op = SparseMatrix()._create(system, constraints, capi.OP_MASS, kernel=capi.KERNEL_GENERIC)
        capi.check(system.lib.gdm_operator_lumped_mass_inverse(op.h, vector.h))
        return vector

    @staticmethod
    def create_laplace_matrix(mapping, system, quadrature, sparse_matrix, constraints, **kw):
        """(grad phi_i, grad phi_j): the loop of tests/poisson_02_gdm.cc:160-206."""
        _check_mapping_quadrature(system, mapping, quadrature)
        return sparse_matrix._create(system, constraints, capi.OP_STIFFNESS, **kw)

    @staticmethod
    def create_advection_matrix(mapping, system, quadrature, sparse_matrix, constraints, velocity, scale=1.0,
                                transpose=False, **kw):
        """(phi_i, b.grad phi_j) [prototypes/advection_01_gdm.cc:164-206] or its transpose
        (b.grad phi_i, phi_j) [advection/stiffness.h:373-418]; residual semantics (constrained rows -> 0)."""
        _check_mapping_quadrature(system, mapping, quadrature)
        kind = capi.OP_ADVECTION_T if transpose else capi.OP_ADVECTION
        return sparse_matrix._create(system, constraints, kind, scale=scale, b=tuple(velocity),
                                     constrained_diagonal=capi.DIAG_ZERO, **kw)


# ------------------------------------------------------------------ solvers
class ReductionControl:
    """dealii::ReductionControl(max_steps, tolerance, reduce)."""

    def __init__(self, max_steps=100, tolerance=1e-10, reduce=1e-2):
        self.c = capi.ReductionControlC(int(max_steps), float(tolerance), float(reduce), 0, 0.0, 0.0)

    def last_step(self):
        return int(self.c.last_step)

    def last_value(self):
        return float(self.c.last_value)

    def initial_value(self):
        return float(self.c.initial_value)


class PreconditionIdentity:
    kind, vec = capi.PRECONDITION_IDENTITY, None


class PreconditionJacobi:
    """PreconditionJacobi<SparseMatrix>::initialize(A) with relaxation 1."""
    kind, vec = capi.PRECONDITION_JACOBI, None

    def initialize(self, matrix):
        self.matrix = matrix


class DiagonalMatrix:
    """dealii::DiagonalMatrix<Vector>: vmult(dst, src) = diag .* src."""
    kind = capi.PRECONDITION_DIAGONAL

    def __init__(self, vec=None):
        self.vec = vec

    def get_vector(self):
        return self.vec

    def vmult(self, dst, src):
        dst.equ(src).scale(self.vec)


class SolverCG:
    """dealii::SolverCG<VectorType>."""

    def __init__(self, control):
        self.control = control

    def solve(self, A, x, b, preconditioner):
        lib = A.lib
        pv = preconditioner.vec.h if getattr(preconditioner, "vec", None) is not None else None
        rc = lib.gdm_solver_cg(A.h, x.h, b.h, preconditioner.kind, pv, C.byref(self.control.c))
        capi.check(rc, self.control.c)


class DiscreteTime:
    """dealii::DiscreteTime(start, end, desired_step_size)."""

    def __init__(self, start, end, step):
        self.start, self.end, self.desired = float(start), float(end), float(step)
        self.current = self.start
        self.next = self._next(self.current)
        self.n = 0

    def _next(self, t):
        nxt = t + self.desired
        if nxt > self.end - 0.05 * self.desired:
            nxt = self.end
        return nxt

    def is_at_end(self):
        return self.current == self.end

    def get_current_time(self):
        return self.current

    def get_next_time(self):
        return self.next

    def get_next_step_size(self):
        return self.next - self.current

    def get_step_number(self):
        return self.n

    def advance_time(self):
        self.n += 1
        self.current = self.next
        self.next = self._next(self.current)


class TimeStepping:
    FORWARD_EULER = capi.RK_FORWARD_EULER
    RK_THIRD_ORDER = capi.RK_THIRD_ORDER
    RK_CLASSIC_FOURTH_ORDER = capi.RK_CLASSIC_FOURTH_ORDER

    class ExplicitRungeKutta:
        """dealii::TimeStepping::ExplicitRungeKutta<VectorType>.

        `evolve_one_time_step(f, t, dt, y)`: y is a Vector or a list of Vectors (block vector);
        f(t, y, out) writes f(t, y) into `out` (same shape as y).  The reference's functor returns
        the vector by value (`problem.h:302-320`); a 2-argument f(t, y) -> Vector is accepted too.
        """

        def __init__(self, method=None):
            self.method = TimeStepping.RK_CLASSIC_FOURTH_ORDER if method is None else method
            self.h = None

        def initialize(self, method):
            self.method = method

        def _ensure(self, system, n_blocks):
            if self.h is None or self._key != (id(system), n_blocks):
                self._free()
                h = C.c_void_p()
                capi.check(system.lib.gdm_rk_create(system.h, self.method, n_blocks, C.byref(h)))
                self.h, self.lib, self._key = h, system.lib, (id(system), n_blocks)

        def evolve_one_time_step(self, f, t, dt, y):
            import inspect
            blocks = list(y) if isinstance(y, (list, tuple)) else [y]
            system = blocks[0].system
            self._ensure(system, len(blocks))
            nb = len(blocks)
            n_args = len(inspect.signature(f).parameters)
            err = []

            class _View(Vector):  # non-owning wrapper around a library-owned stage vector
                def __init__(self, system, h):
                    self.system, self.lib, self.h = system, system.lib, C.c_void_p(h)

                def __del__(self):
                    pass

            def trampoline(tt, yp, outp, _user):
                try:
                    ys = [_View(system, yp[i]) for i in range(nb)]
                    outs = [_View(system, outp[i]) for i in range(nb)]
                    single = not isinstance(y, (list, tuple))
                    if n_args >= 3:
                        f(tt, ys[0] if single else ys, outs[0] if single else outs)
                    else:
                        res = f(tt, ys[0] if single else ys)
                        res = [res] if single else list(res)
                        for o, r in zip(outs, res):
                            o.equ(r)
                    return 0
                except Exception as e:  # never let an exception cross the C boundary
                    err.append(e)
                    return capi.ERR_INTERNAL

            cb = capi.RK_RHS_FN(trampoline)
            arr = (C.c_void_p * nb)(*[b.h for b in blocks])
            t_new = C.c_double()
            rc = self.lib.gdm_rk_evolve_one_time_step(self.h, cb, None, float(t), float(dt), arr, C.byref(t_new))
            if err:
                raise err[0]
            capi.check(rc)
            return t_new.value

        def _free(self):
            if getattr(self, "h", None):
                self.lib.gdm_rk_destroy(self.h)
                self.h = None

        def __del__(self):
            try:
                self._free()
            except Exception:
                pass


# ------------------------------------------------------------- vector tools
class VectorTools:
    """GDM::VectorTools (include/gdm/vector_tools.h)."""
    L2_norm = "L2_norm"

    @staticmethod
    def _wrap(function):
        def cb(pt, comp, _user):
            return float(function([pt[0], pt[1], pt[2]], comp))
        return capi.FUNCTION_FN(cb)

    @staticmethod
    def interpolate(mapping, system, function, vec):
        """vector_tools.h:11-23; function(point[3], component) -> float."""
        cb = VectorTools._wrap(function)
        capi.check(system.lib.gdm_interpolate(system.h, cb, None, vec.h))

    @staticmethod
    def integrate_difference(mapping, system, fe_function, exact_solution, difference=None, quadrature=None,
                             norm="L2_norm"):
        """vector_tools.h:25-86 (L2 only, vector_tools.h:35).  Returns the cell-wise error array."""
        if norm != VectorTools.L2_norm:
            raise ExcNotImplemented(capi.ERR_NOT_IMPLEMENTED, "only L2_norm (vector_tools.h:35)")
        _check_mapping_quadrature(system, mapping, quadrature)
        cb = VectorTools._wrap(exact_solution)
        cell = np.zeros(system.n_cells())
        g = C.c_double()
        capi.check(system.lib.gdm_integrate_difference(system.h, fe_function.h, cb, None,
                                                       cell.ctypes.data_as(C.POINTER(C.c_double)), C.byref(g)))
        if difference is not None:
            difference[:] = cell
        VectorTools._last_global = g.value
        return cell

    @staticmethod
    def compute_global_error(triangulation, cellwise_error, norm="L2_norm"):
        """dealii::VectorTools::compute_global_error for the L2 norm (serial form)."""
        return float(np.sqrt(np.sum(np.asarray(cellwise_error) ** 2)))


# ------------------------------------------------------------ cut-cell set-up
class CutPoisson:
    """Cut-cell set-up of the CutFEM Poisson problem of `prototypes/cut_poisson_01_gdm.cc` (host side).

    Classifies the cells of a Cartesian grid against a Q1 level set (`:105-121`), generates the cut quadratures
    (`:176-190`) and assembles the rows that differ from the plain stiffness operator (`:196-329`): `rows()` are the
    arguments of `SparseMatrix.attach_csr`, `rhs()` the load vector, `l2_error_inside` the error of `:349-398`.

        cut = CutPoisson(2, 3, [64, 64], [-1.21] * 2, [1.21] * 2, level_set_nodal, ghost_penalty=True)
        A = SparseMatrix(); MatrixCreator.create_laplace_matrix(mapping, system, quadrature, A, AffineConstraints())
        A.attach_csr(*cut.rows())
        SolverCG(ReductionControl(n, 1e-10, 1e-6)).solve(A, u, Vector(system, cut.rhs()), PreconditionIdentity())
    """
    INSIDE, OUTSIDE, INTERSECTED = 0, 1, 2

    def __init__(self, dim, fe_degree, n_subdivisions, lo, hi, level_set, ghost_penalty=True, ghost_parameter=0.5,
                 nitsche_parameter=None, rhs_value=4.0, boundary_value=1.0, gp_h_power=1, kind="stiffness",
                 outside_diagonal=1.0, row_range=None, surface_terms=True, domain_boundary_terms=False,
                 level_set_degree=1):
        self.lib = capi.load()
        d = capi.CutDesc()
        d.dim, d.fe_degree = dim, fe_degree
        for e in range(dim):
            d.n_subdivisions[e] = int(n_subdivisions[e])
            d.lo[e], d.hi[e] = float(lo[e]), float(hi[e])
        d.ghost_penalty, d.gp_h_power = int(bool(ghost_penalty)), int(gp_h_power)
        d.ghost_parameter = float(ghost_parameter)
        d.nitsche_parameter = float(5.0 * (fe_degree + 1) * fe_degree if nitsche_parameter is None else nitsche_parameter)
        d.rhs_value, d.boundary_value = float(rhs_value), float(boundary_value)
        d.kind, d.outside_diagonal = {"stiffness": 0, "mass": 1}[kind], float(outside_diagonal)
        d.no_surface_terms, d.domain_boundary_terms = int(not surface_terms), int(bool(domain_boundary_terms))
        if row_range is not None:  # the locally owned DoF range of a rank (System.locally_owned_range())
            d.row_begin, d.row_end = int(row_range[0]), int(row_range[1])
        self.n_dofs = int(np.prod([int(n) + 1 for n in n_subdivisions[:dim]]))
        self.n_cells = int(np.prod([int(n) for n in n_subdivisions[:dim]]))
        level_set = np.ascontiguousarray(level_set, dtype=np.float64)
        d.level_set_degree = int(level_set_degree)
        n_ls = self.n_dofs if level_set_degree <= 1 else int(np.prod([level_set_degree * int(n) + 1 for n in n_subdivisions[:dim]]))
        if level_set.size != n_ls:
            raise GdmError(capi.ERR_INVALID, f"level set has {level_set.size} values, expected {n_ls}")
        self.h = C.c_void_p()
        capi.check(self.lib.gdm_cut_poisson_create(C.byref(d), level_set.ctypes.data_as(C.c_void_p), C.byref(self.h)))

    @staticmethod
    def level_set_points(n_subdivisions, lo, hi, degree):
        """Coordinates [n, dim] (x fastest) at which a level set of degree `degree` is sampled: the support points of FE_Q(degree)
        of every cell (Gauss-Lobatto points), (degree N_e + 1) per direction."""
        inner = np.polynomial.legendre.Legendre.basis(degree).deriv().roots() if degree > 1 else np.zeros(0)
        gll = 0.5 * (np.concatenate([[-1.0], np.sort(inner.real), [1.0]]) + 1.0)
        axes = []
        for n, a, b in zip(n_subdivisions, lo, hi):
            h = (b - a) / n
            ax = np.concatenate([a + (c + gll[:-1]) * h for c in range(n)] + [[b]])
            axes.append(ax)
        grids = np.meshgrid(*axes[::-1], indexing="ij")
        return np.stack([g_.ravel() for g_ in grids[::-1]], axis=1)

    def sizes(self):
        """n_rows, nnz, n_identity_rows, (inside, outside, intersected) cell counts."""
        n_rows, nnz, n_id = C.c_uint64(), C.c_uint64(), C.c_uint64()
        cells = (C.c_uint64 * 3)()
        capi.check(self.lib.gdm_cut_sizes(self.h, C.byref(n_rows), C.byref(nnz), C.byref(n_id), cells))
        return n_rows.value, nnz.value, n_id.value, tuple(int(c) for c in cells)

    def rows(self):
        n_rows, nnz, _, _ = self.sizes()
        row_ids, rowptr = np.zeros(n_rows, dtype=np.uint64), np.zeros(n_rows + 1, dtype=np.uint64)
        col, val = np.zeros(nnz, dtype=np.uint64), np.zeros(nnz)
        capi.check(self.lib.gdm_cut_rows(self.h, row_ids.ctypes.data_as(C.c_void_p), rowptr.ctypes.data_as(C.c_void_p),
                                         col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p)))
        return row_ids, rowptr, col, val

    def rhs(self):
        out = np.zeros(self.n_dofs)
        capi.check(self.lib.gdm_cut_rhs(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def load_vector(self, f=None, g=None):
        """(v, f) over the inside part + <gamma_D / h v - dv/dn, g> on the surface (`wave/stiffness.h:186-260`);
        f, g: functions (point[3], component) -> float or None."""
        out = np.zeros(self.n_dofs)
        fcb = VectorTools._wrap(f) if f is not None else None
        gcb = VectorTools._wrap(g) if g is not None else None
        cast = lambda cb: C.cast(cb, C.c_void_p) if cb is not None else None
        capi.check(self.lib.gdm_cut_load_vector(self.h, cast(fcb), None, cast(gcb), None, out.ctypes.data_as(C.c_void_p)))
        return out

    def boundary_load_vector(self, g):
        """<gamma_D / h v - dv/dn, g> on the box boundary (`wave/stiffness.h:262-340`)."""
        out = np.zeros(self.n_dofs)
        cb = VectorTools._wrap(g)
        capi.check(self.lib.gdm_cut_boundary_load_vector(self.h, cb, None, out.ctypes.data_as(C.c_void_p)))
        return out

    def coupling_rows(self, which):
        """CSR rows (row_ids, rowptr, col, val) of the interface coupling matrices of the two-domain runs
        (`wave/stiffness.h:441-574`): which = "P", "PT" or "Q"."""
        w = {"P": 0, "PT": 1, "Q": 2}[which]
        n_rows, nnz = C.c_uint64(), C.c_uint64()
        capi.check(self.lib.gdm_cut_coupling_rows(self.h, w, 0, 0, C.byref(n_rows), C.byref(nnz), None, None, None, None))
        row_ids, rowptr = np.zeros(n_rows.value, dtype=np.uint64), np.zeros(n_rows.value + 1, dtype=np.uint64)
        col, val = np.zeros(nnz.value, dtype=np.uint64), np.zeros(nnz.value)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        capi.check(self.lib.gdm_cut_coupling_rows(self.h, w, n_rows.value, nnz.value, C.byref(n_rows), C.byref(nnz),
                                                  p(row_ids), p(rowptr), p(col), p(val)))
        return row_ids, rowptr, col, val

    def locations(self):
        out = np.zeros(self.n_cells, dtype=np.uint8)
        capi.check(self.lib.gdm_cut_locations(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def l2_error_inside(self, u, exact):
        """exact(point[3], component) -> float."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.n_dofs
        cb = VectorTools._wrap(exact)
        err = C.c_double()
        capi.check(self.lib.gdm_cut_l2_error_inside(self.h, u.ctypes.data_as(C.c_void_p), cb, None, C.byref(err)))
        return err.value

    def error_norms_inside(self, u, exact):
        """(L2, L1, Linf) over the inside part: the columns of `applications/wave` (`wave/problem.h:531-615`)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.n_dofs
        cb = VectorTools._wrap(exact)
        out = (C.c_double * 3)()
        capi.check(self.lib.gdm_cut_error_norms_inside(self.h, u.ctypes.data_as(C.c_void_p), cb, None, out))
        return tuple(out)

    @staticmethod
    def quadrature(vertex_values, n_gauss):
        """The generator on one unit cell: vertex_values of shape (2,)*dim indexed [x][y][z].
        Returns (points, weights), (points, weights, normals)."""
        lib = capi.load()
        v = np.asarray(vertex_values, dtype=np.float64)
        dim = v.ndim
        flat = np.ascontiguousarray(v.transpose(*range(dim - 1, -1, -1)).ravel())  # x fastest: bit e = direction e
        n_in, n_sf = C.c_uint64(), C.c_uint64()
        capi.check(lib.gdm_cut_quadrature(dim, flat.ctypes.data_as(C.c_void_p), n_gauss, 0, C.byref(n_in), None, None,
                                          C.byref(n_sf), None, None, None))
        cap = max(n_in.value, n_sf.value, 1)
        ip, iw = np.zeros((cap, dim)), np.zeros(cap)
        sp, sw, sn = np.zeros((cap, dim)), np.zeros(cap), np.zeros((cap, dim))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        capi.check(lib.gdm_cut_quadrature(dim, p(flat), n_gauss, cap, C.byref(n_in), p(ip), p(iw), C.byref(n_sf),
                                          p(sp), p(sw), p(sn)))
        a, b = n_in.value, n_sf.value
        return (ip[:a], iw[:a]), (sp[:b], sw[:b], sn[:b])

    def _free(self):
        if getattr(self, "h", None):
            self.lib.gdm_cut_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass
