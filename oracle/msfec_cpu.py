"""ctypes binding of oracle/libmsfec_cpu.so (oracle/msfec_cpu.cpp): the compiled CPU baseline of the Ned_RT basis
build -- 'exact' (banded LDL^T) and 'reference-shaped' (per right-hand side Schur-complement CG + GMRES(ILU(0)) at
1e-6, the algorithm of reference source/Ned_RT/ned_rt_basis.cc:637-847).  TEST / BENCH INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmsfec_cpu.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(path)
        L.msfec_cpu_ned_rt.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def build_basis_ned_rt(prob, corners, ids, mode="exact", n_threads=1):
    """prob: oracle.msfec_oracle.Problem (pairing NED_RT; canonical B expression; polynomial rhs family with constant
    `scale`).  Returns (M[n,18,18], r[n,18], its[n,2])."""
    assert prob.pairing == "NED_RT"
    corners = np.ascontiguousarray(corners, np.float64)
    n = corners.shape[0]
    ids = np.ascontiguousarray(ids, np.int64)
    prm = np.array(list(prob.a_scale) + list(prob.a_alpha) + [float(f) for f in prob.a_freq] + [1.0 if prob.a_rotate else 0.0,
                   prob.b_scale, prob.b_alpha, float(prob.b_freq), float(prob.rhs_constants.get("scale", 1.0)), 0.0])
    M = np.empty((n, 18, 18)); r = np.empty((n, 18)); its = np.zeros((n, 2))
    rc = lib().msfec_cpu_ned_rt(prob.n_refine_local, n, corners.ctypes.data, ids.ctypes.data, int(prob.random_field_seed),
                                float(prob.random_field_sigma), prm.ctypes.data, 0 if mode == "exact" else 1, int(n_threads),
                                M.ctypes.data, r.ctypes.data, its.ctypes.data)
    if rc:
        raise RuntimeError(f"msfec_cpu_ned_rt failed ({rc})")
    return M, r, its
