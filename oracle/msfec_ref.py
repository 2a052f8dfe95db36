"""ctypes binding of oracle/_ref/libmsfec_ref.so: the REFERENCE's own closed-form data classes (Diffusion_A,
DiffusionInverse_A, ReactionRate, BasisQ1<3>, BasisQ1Grad<3>), compiled unmodified from /root/reference against the
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
