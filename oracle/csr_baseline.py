"""ctypes front end of oracle/csr_baseline.c (CPU baseline; test/bench infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "build", "libgdm_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", HERE])
        lib = C.CDLL(path)
        lib.gdm_oracle_kron_csr.restype = C.c_int64
        lib.gdm_oracle_cg.restype = C.c_int
        lib.gdm_oracle_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def _bands(system, constraints, kind, b=None):
    """Per-direction constrained band tables in the product's convention, built from the oracle's
    dense 1D matrices (Dirichlet rows/columns removed; no periodic support here)."""
    from .assemble import matrices_1d
    p, dim = system.fe_degree, system.dim
    W = 2 * p + 1
    A, B, pat = [], [], []
    for d in range(3):
        if d >= dim:
            A.append(np.zeros((1, W))); B.append(np.zeros((1, W))); pat.append(np.zeros((1, W)))
            A[-1][0, p] = 1.0; pat[-1][0, p] = 1.0
            continue
        N = system.n_subdivisions[d]
        M, K, Cm, _ = matrices_1d(p, N, system.h[d])
        Bm = {"mass": None, "stiffness": K, "advection": None if b is None else b[d] * Cm}[kind]
        mask = np.ones(N + 1)
        for s in (0, 1):
            node = 0 if s == 0 else N
            stride = int(np.prod(system.n_nodes[:d])) * system.n_components
            if constraints.is_constrained(node * stride):
                mask[node] = 0.0

        def band(mat, masked=True):
            t = np.zeros((N + 1, W))
            for i in range(N + 1):
                for k in range(W):
                    j = i + k - p
                    if 0 <= j <= N:
                        t[i, k] = mat[i, j] * (mask[i] * mask[j] if masked else 1.0)
            return t
        A.append(band(M)); pat.append(band(M, False))
        B.append(band(Bm) if Bm is not None else np.zeros((N + 1, W)))
    return A, B, pat


class CsrOperator:
    """Assembled CSR matrix of a 3D (or lower) GDM operator; constrained rows are left empty
    (homogeneous constraints with x0 = 0 never touch them in CG)."""

    def __init__(self, system, constraints, kind, b=None):
        lib = load()
        A, B, pat = _bands(system, constraints, kind, b)
        self._keep = (A, B, pat)
        n = (C.c_int * 3)(*[a.shape[0] for a in A])
        PP = C.POINTER(C.c_double)
        arr = lambda lst: (PP * 3)(*[np.ascontiguousarray(a).ctypes.data_as(PP) for a in lst])
        self.n_rows = int(np.prod([a.shape[0] for a in A]))
        self.rowptr = np.zeros(self.n_rows + 1, dtype=np.int64)
        has_b = int(kind != "mass")
        nnz = lib.gdm_oracle_kron_csr(n, system.fe_degree, has_b, arr(A), arr(B), arr(pat),
                                      self.rowptr.ctypes.data_as(C.c_void_p), None, None)
        self.col = np.zeros(nnz, dtype=np.int32)
        self.val = np.zeros(nnz)
        lib.gdm_oracle_kron_csr(n, system.fe_degree, has_b, arr(A), arr(B), arr(pat),
                                self.rowptr.ctypes.data_as(C.c_void_p), self.col.ctypes.data_as(C.c_void_p),
                                self.val.ctypes.data_as(C.c_void_p))
        self.nnz = int(nnz)

    def vmult(self, y, x):
        load().gdm_oracle_spmv(C.c_int64(self.n_rows), self.rowptr.ctypes.data_as(C.c_void_p),
                               self.col.ctypes.data_as(C.c_void_p), self.val.ctypes.data_as(C.c_void_p),
                               x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))

    def cg(self, x, b, max_steps, tol, reduce, dinv=None):
        work = np.zeros(4 * self.n_rows)
        last = C.c_double()
        it = load().gdm_oracle_cg(C.c_int64(self.n_rows), self.rowptr.ctypes.data_as(C.c_void_p),
                                  self.col.ctypes.data_as(C.c_void_p), self.val.ctypes.data_as(C.c_void_p),
                                  x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                  None if dinv is None else dinv.ctypes.data_as(C.c_void_p), C.c_int(max_steps),
                                  C.c_double(tol), C.c_double(reduce), C.byref(last), work.ctypes.data_as(C.c_void_p))
        return it, last.value

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val, self.col, self.rowptr), shape=(self.n_rows, self.n_rows))


def num_threads():
    return load().gdm_oracle_num_threads()


def set_num_threads(n):
    """Use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib = load()
    lib.gdm_oracle_set_num_threads.argtypes = [C.c_int]
    lib.gdm_oracle_set_num_threads.restype = None
    lib.gdm_oracle_set_num_threads(int(n))
