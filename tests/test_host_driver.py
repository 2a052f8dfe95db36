"""CPU tests of the host driver pieces that need no GPU: the C++ coarse assembly + solve (mpi-msfec_b200/host/coarse.cpp,
the stand-in for the reference's Trilinos solve, source/Ned_RT/ned_rt_global.cc:192-461) against the harness' scipy solve
on oracle element matrices, the deal.II CellId / file naming helpers, and that libmsfec_comm.so exports its ABI."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import coarse_solve as cs
from common import ROOT, oracle_problem
from oracle import msfec_oracle as mo

HOST = os.path.join(ROOT, "mpi-msfec_b200", "host")
PAIRING = {"Q": 0, "Q_NED": 1, "NED_RT": 2, "RT_DQ": 3}


@pytest.fixture(scope="module")
def host_built(msfec):
    subprocess.run(["make", "-C", os.path.join(ROOT, "mpi-msfec_b200")], check=True, capture_output=True)
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return HOST


def _oracle_elements(pairing, g, L):
    cells = mo.morton_cells(g)
    prob = oracle_problem(pairing, L)
    M, r = [], []
    for c in range(len(cells)):
        Mc, rc, *_ = mo.build_basis(prob, cells[c], c)
        M.append(Mc); r.append(rc)
    return cells, np.array(M), np.array(r)


@pytest.mark.parametrize("dense_limit", [6000, 0], ids=["dense-lu", "schur-cg"])
@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_cpp_coarse_solve_matches_harness(host_built, tmp_path, pairing, dense_limit):
    """Element matrices of 64 coarse cells (2 global refinements, 1 local refinement) from the oracle -> C++ assembly and
    solve (dense LU and the reference-shaped Schur-complement CG / CG) -> the same per-cell weights as the scipy harness."""
    g = 2
    cells, M, r = _oracle_elements(pairing, g, 1)
    n, k = r.shape
    order = np.random.default_rng(3).permutation(n)           # cells arrive in any order (ranks' chunks)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([PAIRING[pairing], g, n, dense_limit], np.int64).tobytes())
        for c in order:
            f.write(np.array([c], np.int64).tobytes()); f.write(M[c].tobytes()); f.write(r[c].tobytes())
    res = subprocess.run([os.path.join(host_built, "coarse_test"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    w = np.fromfile(tmp_path / "out.bin").reshape(n, k)
    ref = cs.solve_coarse(pairing, g, cells, M, r)[order]
    tol = 1e-10 if dense_limit else 1e-7
    assert np.abs(w - ref).max() <= tol * np.abs(ref).max(), (res.stdout, np.abs(w - ref).max() / np.abs(ref).max())
    assert ("dense LU" in res.stdout) == bool(dense_limit)
    resid = float(re.search(r"relative residual ([0-9.e+-]+)", res.stdout).group(1))
    assert resid < (1e-11 if dense_limit else 1e-8)


def test_comm_library_exports(host_built):
    """libmsfec_comm.so exports every symbol include/msfec_comm.h declares; without a device, creation fails with a message.
    Runs in a child process: the library links the system NCCL, which must not share a process with torch's bundled one."""
    hdr = open(os.path.join(ROOT, "include", "msfec_comm.h")).read()
    declared = sorted(set(re.findall(r"\b(msfec_comm_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 7
    code = f"""
import ctypes as C, os
lib = C.CDLL({os.path.join(ROOT, "mpi-msfec_b200", "libmsfec_comm.so")!r})
for name in {declared!r}:
    assert hasattr(lib, name), name
lib.msfec_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
lib.msfec_comm_last_error.restype = C.c_char_p
out = C.c_void_p()
if not os.path.exists("/dev/nvidiactl"):
    assert lib.msfec_comm_create(0, 1, 0, C.byref(out)) != 0 and b"CUDA" in lib.msfec_comm_last_error()
assert lib.msfec_comm_create(3, 2, 0, C.byref(out)) != 0          # rank outside the world
print("ok")
"""
    import sys
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_driver_cli_errors(host_built, tmp_path):
    exe = os.path.join(host_built, "MsFEC_Ned_RT")
    assert subprocess.run([exe, "-x"], capture_output=True).returncode == 1
    assert subprocess.run([exe], capture_output=True).returncode == 1
    r = subprocess.run([exe, "-p", str(tmp_path / "missing.prm")], capture_output=True, text=True)
    assert r.returncode == 1 and "Exception on processing" in r.stderr


def test_coarse_device_solver_exports_and_fails_loudly_without_gpu(host_built):
    """libmsfec_b200.so exports what include/msfec_coarse.h declares; without a device the call is an error (no CPU fallback
    inside it -- the host solver of coarse.cpp is a separate, explicit choice)."""
    hdr = open(os.path.join(ROOT, "include", "msfec_coarse.h")).read()
    declared = sorted(set(re.findall(r"\b(msfec_coarse_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == ["msfec_coarse_last_error", "msfec_coarse_solve_device"]
    lib = C.CDLL(os.path.join(ROOT, "mpi-msfec_b200", "libmsfec_b200.so"))
    for name in declared:
        assert hasattr(lib, name), name

    class Csr(C.Structure):
        _fields_ = [("n_rows", C.c_int32), ("n_cols", C.c_int32), ("ptr", C.POINTER(C.c_int32)), ("col", C.POINTER(C.c_int32)),
                    ("val", C.POINTER(C.c_double))]
    ptr = (C.c_int32 * 3)(0, 1, 2); col = (C.c_int32 * 2)(0, 1); val = (C.c_double * 2)(2.0, 4.0)
    A = Csr(2, 2, ptr, col, val)
    b = (C.c_double * 2)(2.0, 4.0); x = (C.c_double * 2)()
    lib.msfec_coarse_last_error.restype = C.c_char_p
    lib.msfec_coarse_solve_device.argtypes = [C.c_int, C.POINTER(Csr), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                              C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
    if not os.path.exists("/dev/nvidiactl"):
        assert lib.msfec_coarse_solve_device(0, C.byref(A), None, None, None, b, None, 1e-13, 1e-11, x, None, None) != 0
        assert b"no CUDA device" in lib.msfec_coarse_last_error()
    assert lib.msfec_coarse_solve_device(0, None, None, None, None, b, None, 1e-13, 1e-11, x, None, None) != 0
    assert b"null argument" in lib.msfec_coarse_last_error()


def test_driver_reads_both_method_sections(host_built, tmp_path):
    """`Standard method parameters` (fine-grid comparator, run first) and `Multiscale method parameters` are both read as in the
    reference's main (main_ned_rt.cxx:66-82); with `compute solution = false` each run says so and returns
    (ned_rt_ref.cc:736-742, ned_rt_global.cc:707-713) -- no GPU needed.  A malformed comparator section is an error."""
    src = open(os.path.join(ROOT, "examples", "prm", "prm_ned_rt_test-01.prm")).read()
    assert src.count("set compute solution = true") == 2
    exe = os.path.join(host_built, "MsFEC_Ned_RT")
    off = tmp_path / "off.prm"
    off.write_text(src.replace("set compute solution = true", "set compute solution = false"))
    r = subprocess.run([exe, "-p", str(off)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    a = r.stdout.index("Run of standard problem is explicitly disabled in parameter file.")
    b = r.stdout.index("Run of multiscale problem is explicitly disabled in parameter file.")
    assert a < b
    bad = tmp_path / "bad.prm"
    bad.write_text(src.replace("set refinements = 6", "set refinements = many"))
    r = subprocess.run([exe, "-p", str(bad)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Standard method parameters/Mesh/refinements" in r.stderr
