"""GPU tests of the C++ host side above the C ABI (mpi-msfec_b200/host): the reference-shaped per-cell classes in a
std::map (source/Ned_RT/ned_rt_global.cc:61-98), and the four MsFEC_* executables end to end -- basis build, coarse
assembly + solve, weight scatter, norms, VTU / PVTU output with the reference's file names -- on one rank and, when the
box has two GPUs, on two ranks over NCCL."""
import os
import re
import subprocess
import sys
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import coarse_solve as cs
from common import ROOT, lib_problem, prm_path, rel_err
from oracle import msfec_oracle as mo

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "mpi-msfec_b200", "host")
EXE = {"Q": "MsFEC_Q", "Q_NED": "MsFEC_Q_Ned", "NED_RT": "MsFEC_Ned_RT", "RT_DQ": "MsFEC_RT_DQ"}
STEM = {"Q": "basis_q", "Q_NED": "basis_q-ned", "NED_RT": "basis_ned-rt", "RT_DQ": "basis_rt-dq"}
FAMILIES = {"Q": {"": 8}, "Q_NED": {".h1": 8, ".curl": 12}, "NED_RT": {".curl": 12, ".div": 6}, "RT_DQ": {".div": 6}}
PAIRING = {"Q": 0, "Q_NED": 1, "NED_RT": 2, "RT_DQ": 3}


@pytest.fixture(scope="module")
def host_built(msfec):
    subprocess.run(["make", "-C", os.path.join(ROOT, "mpi-msfec_b200")], check=True, capture_output=True)
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return HOST


def _small_prm(tmp_path, pairing, L=2, g=2, verbose_basis=False, std_refinements=2):
    txt = open(prm_path(pairing)).read()
    # the fine-grid comparator (`Standard method parameters`, shipped: 6 refinements = 64^3 cells) on a small mesh too
    assert "set refinements = 6" in txt and txt.count("set compute solution = true") == 2
    txt = txt.replace("set refinements = 6", f"set refinements = {std_refinements}")
    txt = re.sub(r"set local refinements = \d+", f"set local refinements = {L}", txt)
    txt = re.sub(r"set global refinements = \d+", f"set global refinements = {g}", txt)
    txt = re.sub(r"set dirname output = \S+", f"set dirname output = {tmp_path}/out", txt)
    if verbose_basis:
        txt = txt.replace("set verbose basis = false", "set verbose basis = true")
    assert "set use direct solver basis = false" in txt           # verbatim flag of the shipped files
    p = tmp_path / "t.prm"
    p.write_text(txt)
    return p, re.search(r"Multiscale method parameters.*?set filename output = (\S+)", txt, re.S).group(1)


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_reference_shaped_classes_in_a_map(host_built, tmp_path, pairing):
    """host/test_basis_map.cpp: std::map<CellId, XBasis>, copy-construction before run(), run(), getters, weights --
    bitwise equal to one batched C-ABI call."""
    prm, _ = _small_prm(tmp_path, pairing, L=2, g=1)
    r = subprocess.run([os.path.join(host_built, "test_basis_map"), str(PAIRING[pairing]), str(prm)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bitwise" in r.stdout, r.stdout + r.stderr


def _run_driver(host_built, pairing, prm, env_extra, nproc=1, port=29611):
    exe = os.path.join(host_built, EXE[pairing])
    env = dict(os.environ, **env_extra)
    if nproc == 1:
        cmd = [exe, "-p", str(prm)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), "--no-python", exe, "-p", str(prm)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def _norms(stdout):
    line = [l for l in stdout.splitlines() if "Multiscale solution norms" in l][0]
    return np.array([float(x) for x in re.findall(r"= ([0-9.e+-]+)", line)])


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_driver_end_to_end_single_rank(msfec, host_built, tmp_path, pairing):
    """MsFEC_* -p file.prm (the reference's CLI) on one rank with the NCCL communicator on: element matrices equal the ctypes
    path, the C++ coarse solve equals the harness' scipy solve, norms equal the device norms of the same weights, and the
    output files carry the reference's names (ned_rt_basis.cc:1001-1022, 1254-1259; ned_rt_global.cc:652-700)."""
    prm, fname = _small_prm(tmp_path, pairing, verbose_basis=True)
    out = _run_driver(host_built, pairing, prm, {"MSFEC_NCCL": "1"})
    assert "NCCL communicator over 1 rank(s)" in out and "Outer solver completed." in out
    # the shipped .prm files ask for the fine-grid comparator too (`Standard method parameters / compute solution = true`): it runs
    # first, as in the reference's main (test_standard_method_matches_oracle checks its result)
    assert out.index("Solving >> STANDARD << problem in 3D.") < out.index("Solving >> MULTISCALE << problem in 3D.")
    assert "Standard solution norms" in out
    assert len(re.findall(r"Solving for basis in cell   0_2:\d\d   \[machine: .* \| rank: 0\]   \.\.\.\.\.done in", out)) == 64
    d = tmp_path / "out"
    name = EXE[pairing][6:]
    raw = np.fromfile(d / f"{name}_element_matrices.rank0.bin", dtype=np.uint8)
    hdr = raw[:32].view(np.int64)
    k = int(hdr[1])
    assert hdr[0] == 64 and hdr[2] == 0
    el = raw[32:].view(np.float64).reshape(64, k * k + k)
    cells = mo.morton_cells(2)
    p = msfec.problem_from_prm(str(prm), pairing)
    bb = msfec.BasisBuilder(p, device=0).run(cells, np.arange(64))
    assert rel_err(el[:, :k * k], bb.get_global_element_matrix().reshape(64, -1)) < 1e-12
    assert rel_err(el[:, k * k:], bb.get_global_element_rhs()) < 1e-12
    # coarse solve: C++ (dense LU at this size) vs scipy
    wraw = np.fromfile(d / f"{name}_coarse_weights.bin", dtype=np.uint8)
    w = wraw[16:].view(np.float64).reshape(64, k)
    w_ref = cs.solve_coarse(pairing, 2, cells, bb.get_global_element_matrix(), bb.get_global_element_rhs())
    assert rel_err(w, w_ref) < 1e-9
    # norms printed by the driver == device norms of the same weights through the ctypes path
    bb.set_global_weights(w)
    ref = np.sqrt(bb.solution_norms(64).sum(0))
    got = _norms(out)
    ref = ref[:len(got)] if pairing != "RT_DQ" else ref[[0, 1, 2]]
    assert np.allclose(got, ref[:len(got)], rtol=1e-11, atol=0)
    # files: per-cell solutions, first cell's basis functions, coarse file + two pvtu records
    sol = sorted(d.glob(f"{fname}.00000.cell-0_2:*.vtu"))
    assert len(sol) == 64 and (d / f"{fname}.00000.cell-0_2:45.vtu").exists()          # cell 37 = octal 45
    for fam, cnt in FAMILIES[pairing].items():
        files = sorted(d.glob(f"{STEM[pairing]}{fam}.00000.cell-0_2:00.index-*.vtu"))
        assert len(files) == cnt and files[0].name.endswith(".index-00.vtu"), (fam, files)
    piece = ET.parse(sol[5]).getroot().find("UnstructuredGrid").find("Piece")
    assert piece.get("NumberOfCells") == "64" and piece.get("NumberOfPoints") == "125"
    coarse = ET.parse(d / f"{fname}_n_refine-02.0000.vtu").getroot().find("UnstructuredGrid").find("Piece")
    assert coarse.get("NumberOfCells") == "64"
    assert [a.get("Name") for a in coarse.find("CellData")][-1] == "subdomain_id"
    master = ET.parse(d / f"{fname}_n_refine-02.pvtu").getroot().find("PUnstructuredGrid")
    assert [p_.get("Source") for p_ in master.findall("Piece")] == [f"{fname}_n_refine-02.0000.vtu"]
    fine = ET.parse(d / f"{fname}_fine_refine-02-02.pvtu").getroot().find("PUnstructuredGrid")
    assert sorted(p_.get("Source") for p_ in fine.findall("Piece")) == sorted(f.name for f in sol)


def test_driver_two_ranks_over_nccl(msfec, host_built, tmp_path):
    """Two ranks (one per GPU) under torch.distributed.run: the element matrices travel through ncclAllGather, the norms
    through ncclAllReduce; the printed norms equal the single-rank run to 1e-10 and both ranks write their files."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    prm1, fname = _small_prm(tmp_path / "a", "NED_RT")
    prm2, _ = _small_prm(tmp_path / "b", "NED_RT")
    one = _run_driver(host_built, "NED_RT", prm1, {})
    two = _run_driver(host_built, "NED_RT", prm2, {"NCCL_DEBUG": "WARN"}, nproc=2)
    assert "NCCL communicator over 2 rank(s)" in two and two.count("NCCL communicator over") == 1   # one communicator for both runs
    assert np.allclose(_norms(two), _norms(one), rtol=1e-10, atol=0)
    # the fine-grid comparator ran on both ranks too (its element matrices through the same ncclAllGather)
    std = lambda out: np.array([float(x) for x in re.findall(r"= ([0-9.e+-]+)", [l for l in out.splitlines() if "Standard solution norms" in l][0])])
    assert "Solving >> STANDARD << problem in 3D." in two
    assert np.allclose(std(two), std(one), rtol=1e-10, atol=1e-12 * std(one).max())
    d = tmp_path / "b" / "out"
    assert len(list(d.glob(f"{fname}.00000.cell-*.vtu"))) == 32 and len(list(d.glob(f"{fname}.00001.cell-*.vtu"))) == 32
    master = ET.parse(d / f"{fname}_n_refine-02.pvtu").getroot().find("PUnstructuredGrid")
    assert len(master.findall("Piece")) == 2
    w1 = np.fromfile(tmp_path / "a" / "out" / "Ned_RT_coarse_weights.bin", dtype=np.uint8)[16:].view(np.float64)
    w2 = np.fromfile(d / "Ned_RT_coarse_weights.bin", dtype=np.uint8)[16:].view(np.float64)
    assert np.abs(w1 - w2).max() <= 1e-12 * np.abs(w1).max()


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_coarse_solve_on_device_matches_harness(host_built, tmp_path, pairing):
    """include/msfec_coarse.h: the coarse system's nested iteration on the GPU (Schur-complement CG + inner CG; Q: CG) gives the
    harness' scipy weights (1e-7, the bar of the host iteration in tests/test_host_driver.py) and the host iteration's own
    result to 1e-9; 512 coarse cells (3 global refinements), oracle element matrices at 1 local refinement."""
    import coarse_solve as cs
    from common import oracle_problem
    g = 3
    cells = mo.morton_cells(g)
    prob = oracle_problem(pairing, 1)
    M, r = [], []
    for c in range(len(cells)):
        Mc, rc, *_ = mo.build_basis(prob, cells[c], c)
        M.append(Mc); r.append(rc)
    M = np.array(M); r = np.array(r)
    n, k = r.shape
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([PAIRING[pairing], g, n, 0], np.int64).tobytes())          # dense limit 0: iterate
        for c in range(n):
            f.write(np.array([c], np.int64).tobytes()); f.write(M[c].tobytes()); f.write(r[c].tobytes())
    exe = os.path.join(host_built, "coarse_test")
    dev = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "dev.bin"), "0"], capture_output=True, text=True, timeout=600)
    assert dev.returncode == 0 and "on device 0" in dev.stdout, dev.stdout + dev.stderr
    host = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "host.bin")], capture_output=True, text=True, timeout=600)
    assert host.returncode == 0 and "on device" not in host.stdout, host.stdout + host.stderr
    print(dev.stdout.strip()); print(host.stdout.strip())
    w_dev = np.fromfile(tmp_path / "dev.bin").reshape(n, k)
    w_host = np.fromfile(tmp_path / "host.bin").reshape(n, k)
    ref = cs.solve_coarse(pairing, g, cells, M, r)
    assert np.abs(w_dev - ref).max() <= 1e-7 * np.abs(ref).max()
    assert np.abs(w_dev - w_host).max() <= 1e-9 * np.abs(ref).max()
    resid = float(re.search(r"relative residual ([0-9.e+-]+)", dev.stdout).group(1))
    assert resid < 1e-8


def test_driver_iterates_on_the_device_by_default(host_built, tmp_path):
    """The driver hands the coarse iteration to its GPU when the system is beyond the dense-LU size (forced here with
    MSFEC_COARSE_DENSE_LIMIT=0) and prints the same norms as the dense-LU run to 1e-8; MSFEC_COARSE_SOLVER=host keeps it on
    the host."""
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir(); (tmp_path / "c").mkdir()
    prm_a, _ = _small_prm(tmp_path / "a", "NED_RT")
    prm_b, _ = _small_prm(tmp_path / "b", "NED_RT")
    prm_c, _ = _small_prm(tmp_path / "c", "NED_RT")
    lu = _run_driver(host_built, "NED_RT", prm_a, {})
    dev = _run_driver(host_built, "NED_RT", prm_b, {"MSFEC_COARSE_DENSE_LIMIT": "0"})
    hst = _run_driver(host_built, "NED_RT", prm_c, {"MSFEC_COARSE_DENSE_LIMIT": "0", "MSFEC_COARSE_SOLVER": "host"})
    assert "dense LU" in lu and "Schur-complement CG on device 0" in dev and "Schur-complement CG, " in hst
    # test-01's right-hand side is a pure gradient: sigma vanishes identically (norms ~1e-15 .. 1e-11), so the bar is
    # relative to the largest norm
    tol = dict(rtol=1e-8, atol=1e-9 * _norms(lu).max())
    assert np.allclose(_norms(dev), _norms(lu), **tol)
    assert np.allclose(_norms(hst), _norms(lu), **tol)


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_standard_method_matches_oracle(host_built, tmp_path, pairing):
    """The fine-grid comparator XStd of the reference (source/Ned_RT/ned_rt_ref.cc and siblings; `Standard method parameters` of
    the .prm): the driver runs it as the same pipeline with 0 local refinements on the mesh refined `refinements` times -- the
    device returns the standard lowest-order element matrices, the coarse solve is the fine-grid solve.  Checked against the
    oracle's element matrices at 0 local refinements + the harness' scipy solve, on 8^3 cells (device iteration for the
    pairings beyond the dense-LU size), and the output carries the reference's names (ned_rt_ref.cc:699-730)."""
    from common import oracle_problem
    R = 3
    prm, _ = _small_prm(tmp_path, pairing, L=1, g=1, std_refinements=R)
    out = _run_driver(host_built, pairing, prm, {"MSFEC_COARSE_DENSE_LIMIT": "1000"})
    assert "Solving >> STANDARD << problem in 3D." in out and "solver none (0 local refinements)" in out
    assert ("on device 0" in out) == (pairing != "Q")               # Q: 729 unknowns, dense LU
    name = EXE[pairing][6:] + "Std"
    d = tmp_path / "out"
    raw = np.fromfile(d / f"{name}_element_matrices.rank0.bin", dtype=np.uint8)
    hdr = raw[:32].view(np.int64)
    n, k = int(hdr[0]), int(hdr[1])
    assert n == 8 ** R
    el = raw[32:].view(np.float64).reshape(n, k * k + k)
    cells = mo.morton_cells(R)
    prob = oracle_problem(pairing, 0)
    Mo = np.empty((n, k, k)); ro = np.empty((n, k))
    for c in range(n):
        Mo[c], ro[c] = mo.build_basis(prob, cells[c], c)[:2]
    assert rel_err(el[:, :k * k], Mo.reshape(n, -1)) < 1e-12
    assert np.abs(el[:, k * k:] - ro).max() <= 1e-12 * max(np.abs(ro).max(), 1.0)
    w = np.fromfile(d / f"{name}_coarse_weights.bin", dtype=np.uint8)[16:].view(np.float64).reshape(n, k)
    w_ref = cs.solve_coarse(pairing, R, cells, Mo, ro)
    assert rel_err(w, w_ref) < 1e-7
    std_name = re.search(r"Standard method parameters.*?set filename output = (\S+)", open(prm).read(), re.S).group(1)
    piece = ET.parse(d / f"{std_name}_n_refine-{R:02d}.0000.vtu").getroot().find("UnstructuredGrid").find("Piece")
    assert piece.get("NumberOfCells") == str(n)
    master = ET.parse(d / f"{std_name}_n_refine-{R:02d}.pvtu").getroot().find("PUnstructuredGrid")
    assert [p_.get("Source") for p_ in master.findall("Piece")] == [f"{std_name}_n_refine-{R:02d}.0000.vtu"]
    assert not list(d.glob(f"{std_name}.00000.cell-*.vtu"))          # no per-cell files for the comparator
