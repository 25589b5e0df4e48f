"""ctypes binding of include/gdm/cuda/gdm_c_api.h (the C ABI of libgdm_b200.so).

This is the stub a maintainer of a Python front end would write; the C++ front
end is include/gdm/*.h.  There is no fallback: if the shared library is missing
the import fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GDM_B200_LIB: load a diagnostic build (libgdm_b200_wd.so, build.py) instead of the product library
LIB_PATH = os.environ.get("GDM_B200_LIB") or os.path.join(HERE, "libgdm_b200.so")


class GdmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[gdm status {code}] {msg}")
        self.code = code


class ExcNotImplemented(GdmError):
    """deal.II ExcNotImplemented (reference: AssertThrow(..., ExcNotImplemented()))."""


class NoConvergence(GdmError):
    """deal.II SolverControl::NoConvergence."""

    def __init__(self, code, msg, last_step=None, last_residual=None):
        super().__init__(code, msg)
        self.last_step, self.last_residual = last_step, last_residual


OK, ERR_INVALID, ERR_NOT_IMPLEMENTED, ERR_CUDA, ERR_NO_CONVERGENCE, ERR_COMM, ERR_INTERNAL = range(7)
OP_MASS, OP_STIFFNESS, OP_ADVECTION, OP_ADVECTION_T = range(4)
DIAG_ZERO, DIAG_ASSEMBLED = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED = 0, 1, 2
PRECONDITION_IDENTITY, PRECONDITION_JACOBI, PRECONDITION_DIAGONAL = 0, 1, 2
RK_FORWARD_EULER, RK_THIRD_ORDER, RK_CLASSIC_FOURTH_ORDER = 0, 1, 2


class SystemDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("fe_degree", C.c_int), ("n_components", C.c_int),
                ("n_subdivisions", C.c_uint32 * 3), ("lo", C.c_double * 3), ("hi", C.c_double * 3),
                ("rank", C.c_int), ("n_ranks", C.c_int), ("add_ghost_layer", C.c_int)]


class LayoutInfo(C.Structure):
    _fields_ = [("pitch", C.c_uint64), ("plane", C.c_uint64), ("size", C.c_uint64),
                ("owned_offset", C.c_uint64), ("owned_size", C.c_uint64),
                ("local_nodes", C.c_uint32 * 3), ("owned_begin", C.c_uint32), ("owned_end", C.c_uint32),
                ("stored_begin", C.c_uint32), ("stored_end", C.c_uint32)]


class OperatorDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("scale", C.c_double), ("b", C.c_double * 3),
                ("constrained_diagonal", C.c_int), ("kernel", C.c_int)]


class ReductionControlC(C.Structure):
    _fields_ = [("max_steps", C.c_uint32), ("tolerance", C.c_double), ("reduce", C.c_double),
                ("last_step", C.c_uint32), ("last_value", C.c_double), ("initial_value", C.c_double)]


class CutDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("fe_degree", C.c_int), ("n_subdivisions", C.c_uint32 * 3),
                ("lo", C.c_double * 3), ("hi", C.c_double * 3), ("ghost_penalty", C.c_int), ("gp_h_power", C.c_int),
                ("ghost_parameter", C.c_double), ("nitsche_parameter", C.c_double), ("rhs_value", C.c_double),
                ("boundary_value", C.c_double), ("kind", C.c_int), ("outside_diagonal", C.c_double),
                ("no_surface_terms", C.c_int), ("domain_boundary_terms", C.c_int), ("level_set_degree", C.c_int),
                ("row_begin", C.c_uint64), ("row_end", C.c_uint64)]


FUNCTION_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.c_void_p)
RK_RHS_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p)

_H = C.c_void_p
_PH = C.POINTER(C.c_void_p)
_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)

# name -> (restype, argtypes).  Every symbol declared in gdm_c_api.h appears here;
# tests/test_capi_symbols.py checks the header and this table against the library.
SIGNATURES = {
    "gdm_last_error": (C.c_char_p, []),
    "gdm_api_version": (C.c_int, []),
    "gdm_context_create": (C.c_int, [C.c_int, C.c_void_p, _PH]),
    "gdm_context_destroy": (C.c_int, [_H]),
    "gdm_context_set_stream": (C.c_int, [_H, C.c_void_p]),
    "gdm_context_synchronize": (C.c_int, [_H]),
    "gdm_context_launch_count": (C.c_int, [_H, _PU64]),
    "gdm_comm_unique_id": (C.c_int, [C.c_void_p]),
    "gdm_context_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "gdm_polynomials_1d": (C.c_int, [C.c_int, _PD]),
    "gdm_system_create": (C.c_int, [_H, C.POINTER(SystemDesc), _PH]),
    "gdm_system_destroy": (C.c_int, [_H]),
    "gdm_system_n_dofs": (C.c_uint64, [_H]),
    "gdm_system_n_cells": (C.c_uint64, [_H]),
    "gdm_system_locally_owned_range": (C.c_int, [_H, _PU64, _PU64]),
    "gdm_system_dofs_per_cell": (C.c_int, [_H]),
    "gdm_system_get_dof_indices": (C.c_int, [_H, C.c_uint64, _PU64]),
    "gdm_system_active_fe_index": (C.c_int, [_H, C.c_uint64, C.POINTER(C.c_uint32)]),
    "gdm_system_matrix_1d": (C.c_int, [_H, C.c_int, C.c_int, _PD]),
    "gdm_system_layout": (C.c_int, [_H, C.POINTER(LayoutInfo)]),
    "gdm_system_halo_plan": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "gdm_system_sparsity_row": (C.c_int, [_H, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64)]),
    "gdm_pers_partition": (C.c_int, [C.c_int] * 7 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gdm_constraints_create": (C.c_int, [_H, _PH]),
    "gdm_constraints_destroy": (C.c_int, [_H]),
    "gdm_constraints_make_zero_boundary": (C.c_int, [_H, C.c_int]),
    "gdm_constraints_make_periodicity": (C.c_int, [_H, C.c_int]),
    "gdm_constraints_interpolate_boundary_values": (C.c_int, [_H, C.c_int, FUNCTION_FN, C.c_void_p]),
    "gdm_constraints_condense_rhs": (C.c_int, [_H, _H, _H]),
    "gdm_constraints_close": (C.c_int, [_H]),
    "gdm_constraints_n_constraints": (C.c_uint64, [_H]),
    "gdm_constraints_is_constrained": (C.c_int, [_H, C.c_uint64]),
    "gdm_constraints_distribute": (C.c_int, [_H, _H]),
    "gdm_constraints_set_zero": (C.c_int, [_H, _H]),
    "gdm_vector_create": (C.c_int, [_H, _PH]),
    "gdm_vector_destroy": (C.c_int, [_H]),
    "gdm_vector_upload": (C.c_int, [_H, C.c_void_p]),
    "gdm_vector_download": (C.c_int, [_H, C.c_void_p]),
    "gdm_vector_device_ptr": (C.c_void_p, [_H]),
    "gdm_vector_set": (C.c_int, [_H, C.c_double]),
    "gdm_vector_copy": (C.c_int, [_H, _H]),
    "gdm_vector_scale": (C.c_int, [_H, C.c_double]),
    "gdm_vector_add": (C.c_int, [_H, C.c_double, _H]),
    "gdm_vector_sadd": (C.c_int, [_H, C.c_double, C.c_double, _H]),
    "gdm_vector_scale_by": (C.c_int, [_H, _H]),
    "gdm_vector_dot": (C.c_int, [_H, _H, _PD]),
    "gdm_vector_l2_norm": (C.c_int, [_H, _PD]),
    "gdm_vector_linfty_norm": (C.c_int, [_H, _PD]),
    "gdm_vector_update_ghost_values": (C.c_int, [_H]),
    "gdm_operator_create": (C.c_int, [_H, _H, C.POINTER(OperatorDesc), _PH]),
    "gdm_system_write_vtu": (C.c_int, [_H, C.POINTER(C.c_double), C.c_char_p, C.c_char_p]),
    "gdm_cut_poisson_create": (C.c_int, [C.POINTER(CutDesc), C.c_void_p, _PH]),
    "gdm_cut_level_set_points": (C.c_int, [C.POINTER(CutDesc), _PU64, C.c_void_p]),
    "gdm_cut_destroy": (C.c_int, [_H]),
    "gdm_cut_sizes": (C.c_int, [_H, _PU64, _PU64, _PU64, _PU64]),
    "gdm_cut_rows": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gdm_cut_rhs": (C.c_int, [_H, C.c_void_p]),
    "gdm_cut_locations": (C.c_int, [_H, C.c_void_p]),
    "gdm_cut_boundary_load_vector": (C.c_int, [_H, FUNCTION_FN, C.c_void_p, C.c_void_p]),
    "gdm_cut_coupling_rows": (C.c_int, [_H, C.c_int, C.c_uint64, C.c_uint64, _PU64, _PU64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "gdm_cut_load_vector": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gdm_cut_l2_error_inside": (C.c_int, [_H, C.c_void_p, FUNCTION_FN, C.c_void_p, _PD]),
    "gdm_cut_error_norms_inside": (C.c_int, [_H, C.c_void_p, FUNCTION_FN, C.c_void_p, _PD]),
    "gdm_cut_quadrature": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_uint64, _PU64, C.c_void_p, C.c_void_p, _PU64,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "gdm_system_write_matrix": (C.c_int, [_H, _H, C.POINTER(OperatorDesc), C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]),
    "gdm_operator_destroy": (C.c_int, [_H]),
    "gdm_operator_attach_csr": (C.c_int, [_H, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gdm_operator_vmult": (C.c_int, [_H, _H, _H]),
    "gdm_operator_vmult_add": (C.c_int, [_H, _H, _H]),
    "gdm_operator_tvmult": (C.c_int, [_H, _H, _H]),
    "gdm_operator_mass_inverse": (C.c_int, [_H, _H, _H]),
    "gdm_operator_vmult_dot": (C.c_int, [_H, _H, _H, C.POINTER(C.c_double)]),
    "gdm_operator_vmult_host": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "gdm_operator_diagonal": (C.c_int, [_H, _H]),
    "gdm_operator_lumped_mass_inverse": (C.c_int, [_H, _H]),
    "gdm_operator_kernel_used": (C.c_int, [_H]),
    "gdm_operator_m": (C.c_uint64, [_H]),
    "gdm_solver_cg": (C.c_int, [_H, _H, _H, C.c_int, _H, C.POINTER(ReductionControlC)]),
    "gdm_rk_create": (C.c_int, [_H, C.c_int, C.c_int, _PH]),
    "gdm_rk_destroy": (C.c_int, [_H]),
    "gdm_rk_evolve_one_time_step": (C.c_int, [_H, RK_RHS_FN, C.c_void_p, C.c_double, C.c_double, _PH, _PD]),
    "gdm_interpolate": (C.c_int, [_H, FUNCTION_FN, C.c_void_p, _H]),
    "gdm_integrate_difference": (C.c_int, [_H, _H, FUNCTION_FN, C.c_void_p, _PD, _PD]),
}

_lib = None


def load():
    """Load libgdm_b200.so and bind every ABI symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                          " -- gdm_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, control=None):
    if rc == OK:
        return
    msg = load().gdm_last_error().decode("utf-8", "replace")
    if rc == ERR_NOT_IMPLEMENTED:
        raise ExcNotImplemented(rc, msg)
    if rc == ERR_NO_CONVERGENCE:
        raise NoConvergence(rc, msg, getattr(control, "last_step", None), getattr(control, "last_value", None))
    raise GdmError(rc, msg)
