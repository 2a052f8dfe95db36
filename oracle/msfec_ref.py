"""ctypes binding of oracle/_ref/libmsfec_ref.so: the REFERENCE's own closed-form data classes (Diffusion_A,
DiffusionInverse_A, ReactionRate, BasisQ1<3>, BasisQ1Grad<3>, MyMappingQ1<3>, BasisNedelec<3>, BasisRaviartThomas<3> -- the last
two with stand-in unit-cell shape functions, i.e. what is the reference's there is the mapping and the covariant / Piola
transform), compiled unmodified from /root/reference against the
stand-in deal.II headers in oracle/ref_shim (oracle/Makefile).  TEST INFRASTRUCTURE ONLY: used by
tests/test_reference_compiled.py and tests/golden/make_reference_compiled_golden.py to pin the Python oracle."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libmsfec_ref.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def available() -> bool:
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH)
        _lib.msfec_ref_last_error.restype = ctypes.c_char_p
        _lib.msfec_ref_diffusion_a.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib.msfec_ref_reaction_rate.argtypes = [ctypes.c_int, _dp, _dp]
        _lib.msfec_ref_basis_q1.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib.msfec_ref_basis_q1_grad.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib.msfec_ref_mapping.argtypes = [_dp, ctypes.c_int, _dp, _dp, _dp]
        _lib.msfec_ref_basis_vector.argtypes = [_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _check(rc):
    if rc:
        raise RuntimeError(lib().msfec_ref_last_error().decode())


def diffusion_a(prm_path, pts, inverse=False, pointwise=False):
    """Diffusion_A / DiffusionInverse_A (eqn_coeff_A.cc) at pts[n,3] -> [n,3,3]; pointwise: value() instead of
    value_list()."""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    out = np.empty((len(pts), 3, 3))
    _check(lib().msfec_ref_diffusion_a(os.fsencode(prm_path), int(inverse) | (2 if pointwise else 0), len(pts), _p(pts),
                                       _p(out)))
    return out


def reaction_rate(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    out = np.empty(len(pts))
    _check(lib().msfec_ref_reaction_rate(len(pts), _p(pts), _p(out)))
    return out


def basis_q1(vertices, pts):
    """BasisQ1<3> on the cell with vertices[8,3] (deal.II vertex order) at pts[n,3] -> values[n,8], gradients[n,8,3]
    (BasisQ1Grad<3>)."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(8, 3)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    val = np.empty((len(pts), 8)); grad = np.empty((len(pts), 8, 3))
    for i in range(8):
        v = np.empty(len(pts)); g = np.empty((len(pts), 3))
        _check(lib().msfec_ref_basis_q1(_p(vertices), i, len(pts), _p(pts), _p(v)))
        _check(lib().msfec_ref_basis_q1_grad(_p(vertices), i, len(pts), _p(pts), _p(g)))
        val[:, i] = v; grad[:, i, :] = g
    return val, grad


def mapping(vertices, pts):
    """MyMappingQ1<3> (my_mapping_q1.tpp) of the cell: unit-cell coordinates [n,3] and the Jacobian of real -> unit [n,3,3]."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(8, 3)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    ref = np.empty((len(pts), 3)); jac = np.empty((len(pts), 3, 3))
    _check(lib().msfec_ref_mapping(_p(vertices), len(pts), _p(pts), _p(ref), _p(jac)))
    return ref, jac


def basis_vector(vertices, kind, pts, pointwise=False):
    """BasisNedelec<3> (kind 'ned', 12 functions) / BasisRaviartThomas<3> (kind 'rt', 6) on the cell at pts[n,3] -> [n, 12|6, 3]."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(8, 3)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    nb = 12 if kind == "ned" else 6
    out = np.empty((len(pts), nb, 3))
    for i in range(nb):
        v = np.empty((len(pts), 3))
        _check(lib().msfec_ref_basis_vector(_p(vertices), 0 if kind == "ned" else 1, i, 0 if pointwise else 1, len(pts), _p(pts), _p(v)))
        out[:, i, :] = v
    return out
