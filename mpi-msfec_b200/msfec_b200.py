"""ctypes binding of libmsfec_b200.so (C ABI: include/msfec.h).

Harness-side convenience only (tests, bench.py, smoke): the product is the shared
library and the C++ host classes in host/.  Mirrors the reference's per-pairing
interface names where it makes the parity tests read like the reference's usage:
BasisBuilder.run() ~ `for basis in cell_basis_map: basis.run()`
(reference source/Ned_RT/ned_rt_global.cc:91-98), get_global_element_matrix(),
get_global_element_rhs(), set_global_weights() (include/Ned_RT/ned_rt_basis.h:119-152).

There is no CPU fallback: if the library is missing or no B200 is present the calls
raise MsfecError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PAIRING = {"Q": 0, "Q_NED": 1, "NED_RT": 2, "RT_DQ": 3}
SOLVER = {"auto": 0, "minres": 1, "band": 2, "mf": 3}          # enum msfec_solver
SOLVER_STAT = {"minres": 0, "band": 1, "mf": 2}               # msfec_stats.solver
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsfec_b200.so")


class MsfecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"msfec error {code}: {msg}")
        self.code = code


class Problem(C.Structure):
    _fields_ = [
        ("pairing", C.c_int32), ("n_refine_local", C.c_int32), ("n_refine_global", C.c_int32),
        ("use_direct_solver_basis", C.c_int32), ("verbose_basis", C.c_int32), ("a_rotate", C.c_int32),
        ("a_freq", C.c_int32 * 3), ("b_freq", C.c_int32),
        ("a_scale", C.c_double * 3), ("a_alpha", C.c_double * 3),
        ("b_scale", C.c_double), ("b_alpha", C.c_double),
        ("b_expression", C.c_char_p), ("rhs_expression", C.c_char_p), ("rhs_constants", C.c_char_p),
        ("random_field_seed", C.c_uint64), ("random_field_sigma", C.c_double),
        ("krylov_rtol", C.c_double), ("krylov_max_iter", C.c_int32), ("cells_per_batch", C.c_int32),
        ("solver", C.c_int32), ("reserved0", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("k", C.c_int32), ("n_fine_dofs", C.c_int32), ("n_fine_dofs_interior", C.c_int32),
        ("iterations_max", C.c_int32), ("not_converged", C.c_int32), ("kernel_launches", C.c_int32),
        ("reserved", C.c_int32), ("iterations_mean", C.c_double), ("residual_max", C.c_double),
        ("ms_assemble", C.c_double), ("ms_lift", C.c_double), ("ms_solve", C.c_double), ("ms_gram", C.c_double),
        ("ms_total", C.c_double), ("krylov_matrix_bytes", C.c_double), ("krylov_ms_spmm", C.c_double),
        ("krylov_spmm_launches", C.c_int64),
        ("direct_update_launches", C.c_int64), ("direct_flops", C.c_double), ("direct_flops_timed", C.c_double),
        ("direct_ms_update", C.c_double), ("solver", C.c_int32), ("direct_timed_launches", C.c_int32),
        ("mf_flops", C.c_double), ("mf_bytes_fwd", C.c_double), ("mf_bytes_bwd", C.c_double), ("mf_ms_fwd", C.c_double),
        ("mf_ms_bwd", C.c_double), ("mf_launches", C.c_int64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EXPORTS = [
    "msfec_abi_version", "msfec_problem_defaults", "msfec_problem_from_prm", "msfec_problem_free", "msfec_k",
    "msfec_n_fine_dofs", "msfec_create", "msfec_destroy", "msfec_last_error", "msfec_build_basis",
    "msfec_build_basis_device", "msfec_set_weights", "msfec_get_fine_solution", "msfec_get_basis",
    "msfec_solution_norms", "msfec_fine_dof_layout", "msfec_debug_table", "msfec_debug_cell_values",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MsfecError(-1, f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        L = C.CDLL(LIB_PATH)
        L.msfec_last_error.restype = C.c_char_p
        L.msfec_last_error.argtypes = [C.c_void_p]
        L.msfec_create.argtypes = [C.c_int, C.POINTER(Problem), C.POINTER(C.c_void_p)]
        L.msfec_destroy.argtypes = [C.c_void_p]
        L.msfec_problem_defaults.argtypes = [C.POINTER(Problem), C.c_int]
        L.msfec_problem_from_prm.argtypes = [C.c_char_p, C.c_int, C.POINTER(Problem)]
        L.msfec_problem_free.argtypes = [C.POINTER(Problem)]
        vp = C.c_void_p
        L.msfec_build_basis.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.POINTER(Stats)]
        L.msfec_build_basis_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.POINTER(Stats)]
        L.msfec_set_weights.argtypes = [vp, C.c_int, vp]
        L.msfec_get_fine_solution.argtypes = [vp, C.c_int, vp, vp]
        L.msfec_get_basis.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.msfec_solution_norms.argtypes = [vp, C.c_int, vp]
        L.msfec_fine_dof_layout.argtypes = [vp, C.c_int, vp, vp, vp]
        L.msfec_debug_table.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.msfec_debug_cell_values.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_size_t)]
        L.msfec_n_fine_dofs.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_problem(pairing: str, **kw) -> Problem:
    """Problem with the reference's declared defaults, overridden by keywords."""
    p = Problem()
    lib().msfec_problem_defaults(C.byref(p), PAIRING[pairing])
    for k, v in kw.items():
        if k in ("a_freq", "a_scale", "a_alpha"):
            for d in range(3):
                getattr(p, k)[d] = v[d]
        elif k in ("b_expression", "rhs_expression", "rhs_constants"):
            setattr(p, k, v.encode() if isinstance(v, str) else v)
        else:
            setattr(p, k, v)
    return p


def problem_from_prm(path: str, pairing: str) -> Problem:
    p = Problem()
    rc = lib().msfec_problem_from_prm(path.encode(), PAIRING[pairing], C.byref(p))
    if rc:
        raise MsfecError(rc, lib().msfec_last_error(None).decode())
    return p


class BasisBuilder:
    """All locally owned coarse cells of one rank/GPU (one msfec_ctx)."""

    def __init__(self, problem: Problem, device: int = 0):
        self.problem = problem
        self._ctx = C.c_void_p()
        rc = lib().msfec_create(device, C.byref(problem), C.byref(self._ctx))
        if rc:
            raise MsfecError(rc, lib().msfec_last_error(None).decode())
        self.k = lib().msfec_k(problem.pairing)
        n0, n1 = C.c_int(), C.c_int()
        lib().msfec_n_fine_dofs(problem.pairing, problem.n_refine_local, C.byref(n0), C.byref(n1))
        self.n_block = (n0.value, n1.value)
        self.stats = None
        self._M = self._r = None

    def close(self):
        if self._ctx:
            lib().msfec_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc and rc not in allow:
            raise MsfecError(rc, lib().msfec_last_error(self._ctx).decode())
        return rc

    # -- reference-shaped interface ---------------------------------------------------
    def run(self, corners: np.ndarray, cell_ids: np.ndarray | None = None, allow_not_converged=False):
        corners = np.ascontiguousarray(corners, dtype=np.float64)
        n = corners.shape[0]
        assert corners.shape == (n, 8, 3)
        ids = None if cell_ids is None else np.ascontiguousarray(cell_ids, dtype=np.int64)
        self._M = np.empty((n, self.k, self.k)); self._r = np.empty((n, self.k))
        st = Stats()
        self._check(lib().msfec_build_basis(self._ctx, n, _ptr(corners), _ptr(ids), _ptr(self._M), _ptr(self._r),
                                            C.byref(st)), allow=(5,) if allow_not_converged else ())
        self.stats = st.as_dict()
        self.n_cells = n
        return self

    def run_device(self, n, d_corners: int, d_ids: int, d_M: int, d_r: int):
        st = Stats()
        self._check(lib().msfec_build_basis_device(self._ctx, n, d_corners, d_ids or None, d_M, d_r, C.byref(st)))
        self.stats = st.as_dict()
        self.n_cells = n
        return self

    def get_global_element_matrix(self):
        return self._M

    def get_global_element_rhs(self):
        return self._r

    def set_global_weights(self, w: np.ndarray):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self._check(lib().msfec_set_weights(self._ctx, w.shape[0], _ptr(w)))

    def get_fine_solution(self, cell: int):
        b0 = np.empty(self.n_block[0]); b1 = np.empty(self.n_block[1]) if self.n_block[1] else None
        self._check(lib().msfec_get_fine_solution(self._ctx, cell, _ptr(b0), _ptr(b1)))
        return b0, b1

    def solution_norms(self, n_cells: int) -> np.ndarray:
        """[n_cells, 4] squared norms (L2 / semi-norm of block 0, L2 / semi-norm of block 1) of the fine solution
        reconstructed by set_global_weights, summed on the device per cell (msfec_solution_norms)."""
        out = np.empty((n_cells, 4))
        self._check(lib().msfec_solution_norms(self._ctx, n_cells, _ptr(out)))
        return out

    def get_basis(self, cell: int, basis: int):
        b0 = np.empty(self.n_block[0]); b1 = np.empty(self.n_block[1]) if self.n_block[1] else None
        self._check(lib().msfec_get_basis(self._ctx, cell, basis, _ptr(b0), _ptr(b1)))
        return b0, b1

    # -- introspection ------------------------------------------------------------------
    def layout(self, block: int):
        n = self.n_block[block]
        pos = np.empty((n, 3)); axis = np.empty(n, np.int32); bnd = np.empty(n, np.uint8)
        self._check(lib().msfec_fine_dof_layout(self._ctx, block, _ptr(pos), _ptr(axis), _ptr(bnd)))
        return pos, axis, bnd.astype(bool)

    def table(self, name: str) -> np.ndarray:
        cnt = C.c_size_t(0); dt = C.c_int(0)
        self._check(lib().msfec_debug_table(self._ctx, name.encode(), None, C.byref(cnt), C.byref(dt)))
        out = np.empty(cnt.value, np.int32 if dt.value == 0 else np.float64)
        self._check(lib().msfec_debug_table(self._ctx, name.encode(), _ptr(out), C.byref(cnt), C.byref(dt)))
        return out

    def cell_values(self, cell: int) -> np.ndarray:
        cnt = C.c_size_t(0)
        self._check(lib().msfec_debug_cell_values(self._ctx, cell, None, C.byref(cnt)))
        out = np.empty(cnt.value)
        self._check(lib().msfec_debug_cell_values(self._ctx, cell, _ptr(out), C.byref(cnt)))
        return out


DIMS = ["pairing", "n", "nC", "k_solve", "k_gram", "k0", "two_blocks", "N0", "NI0", "N1", "NI1", "NI", "NB", "NF",
        "n_slots0", "n_slots1", "n_rhs_slots", "rhs_ncomp", "k_h_exponent", "f1_H_exponent", "asm00_h_exponent",
        "asm11_h_exponent", "asm_rhs_h_exponent", "rhs_block", "tensor_inverse", "scalar_inverse"]
