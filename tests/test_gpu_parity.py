"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.
Tolerance (north star): coarse element matrices to relative 1e-9 (max-norm scaled by max |M|).
Run with:  gpurun -- python -m pytest tests -m gpu -x -q"""
import os

import numpy as np
import pytest

from common import ROOT, lib_problem, oracle_problem, rel_err
from oracle import msfec_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _check_cells(bb, prob, cells, ids, which):
    M = bb.get_global_element_matrix(); r = bb.get_global_element_rhs()
    worst = 0.0
    for c in which:
        Mo, ro, *_ = mo.build_basis(prob, cells[c], int(ids[c]))
        worst = max(worst, rel_err(M[c], Mo), float(np.abs(r[c] - ro).max() / max(np.abs(ro).max(), 1e-300)))
    return worst


SOLVERS = ["minres", "band", "mf"]


@pytest.mark.parametrize("solver", SOLVERS)
@pytest.mark.parametrize("pairing", mo.PAIRINGS)
@pytest.mark.parametrize("L", [2, 3])
def test_prm_coefficients_match_oracle(msfec, pairing, L, solver):
    """configs[0..3] coefficients (examples/prm == reference test-01 files) on the 64-cell coarse mesh, through all
    three solvers of the local problems: batched MINRES (the reference's solve_iterative), banded block LDL^T and
    multifrontal LDL^T (the reference's solve_direct)."""
    cells = mo.morton_cells(2)
    ids = np.arange(64)
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L, solver=msfec.SOLVER[solver]), device=0).run(cells, ids)
    worst = _check_cells(bb, oracle_problem(pairing, L), cells, ids, (0, 21, 37, 63))
    print(pairing, L, "worst", worst, bb.stats)
    assert worst < TOL
    assert bb.stats["not_converged"] == 0 and bb.stats["kernel_launches"] > 0
    assert bb.stats["solver"] == msfec.SOLVER_STAT[solver]


@pytest.mark.parametrize("flag", [0, 1])
def test_auto_solver_selection(msfec, flag):
    """`use direct solver basis` true and false (what every shipped .prm sets) both get an exact factorisation: the
    multifrontal solver up to 3 local refinements; the flag alone never routes to the slow MINRES fallback."""
    cells = mo.morton_cells(2)
    bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 3, use_direct_solver_basis=flag), device=0).run(cells, np.arange(64))
    assert bb.stats["solver"] == msfec.SOLVER_STAT["mf"] and bb.stats["residual_max"] < 1e-11
    assert bb.stats["mf_launches"] > 0 and bb.stats["mf_ms_fwd"] > 0 and bb.stats["mf_ms_bwd"] > 0


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_golden_fixtures(msfec, pairing):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_elem_matrices.npz"))
    cells = mo.morton_cells(2)
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, 2), device=0).run(cells)
    for c in (5, 37):
        assert rel_err(bb.get_global_element_matrix()[c], gold[f"{pairing}_M_{c}"]) < TOL
        ro = gold[f"{pairing}_r_{c}"]
        assert np.abs(bb.get_global_element_rhs()[c] - ro).max() <= TOL * max(np.abs(ro).max(), 1e-300)


@pytest.mark.parametrize("solver", SOLVERS)
def test_random_field_ragged_batches(msfec, solver):
    """C5-style rough field; 70 cells = 2 full lane groups + a ragged one, split over 2 batches."""
    cells = mo.morton_cells(3)[100:170]
    ids = np.arange(100, 170)
    p = lib_problem(msfec, "NED_RT", 2, random_seed=20261017, cells_per_batch=64, solver=msfec.SOLVER[solver])
    bb = msfec.BasisBuilder(p, device=0).run(cells, ids)
    worst = _check_cells(bb, oracle_problem("NED_RT", 2, random_seed=20261017), cells, ids, (0, 31, 32, 63, 64, 69))
    assert worst < TOL


def test_single_cell_and_basis_vectors(msfec):
    """n_cells = 1 (31 padded lanes) and the fine-scale basis functions themselves."""
    cells = mo.morton_cells(2)[37:38]
    bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 2), device=0).run(cells, np.array([37]))
    prob = oracle_problem("NED_RT", 2)
    Mo, ro, X0, X1, cs = mo.build_basis(prob, cells[0], 37)
    assert rel_err(bb.get_global_element_matrix()[0], Mo) < TOL
    # map library numbering -> oracle numbering through entity positions
    g = mo.fine_grid(prob.n)
    for blk, (opos, Xo) in enumerate(((g.e_pos, X0), (g.f_pos, X1))):
        pos, axis, bnd = bb.layout(blk)
        key = {tuple(np.round(p_ * 2).astype(int)): i for i, p_ in enumerate(opos)}
        perm = np.array([key[tuple(np.round(p_ * 2).astype(int))] for p_ in pos])
        for j in ((0, 5, 11) if blk == 0 else (12, 17)):
            b0, b1 = bb.get_basis(0, j)
            mine = b0 if blk == 0 else b1
            assert rel_err(mine, Xo[j][perm]) < 1e-8
    # set_global_weights: u_fine = sum_i w_i b_i
    w = np.random.default_rng(1).standard_normal((1, 18))
    bb.set_global_weights(w)
    u0, u1 = bb.get_fine_solution(0)
    ref0 = sum(w[0, j] * bb.get_basis(0, j)[0] for j in range(18))
    ref1 = sum(w[0, j] * bb.get_basis(0, j)[1] for j in range(18))
    assert rel_err(u0, ref0) < 1e-13 and rel_err(u1, ref1) < 1e-13


def test_constant_coefficient_reproduction_on_gpu(msfec):
    """Invariant 3 through the CUDA path: constant diagonal A -> M equals the L=0 analytic matrix."""
    cells = mo.morton_cells(1)
    for pairing in mo.PAIRINGS:
        rhs = b"1;2;3" if pairing in ("Q_NED", "NED_RT") else b"1"
        p = msfec.make_problem(pairing, n_refine_local=3, a_rotate=0, a_scale=(2.0, 3.0, 0.5), a_freq=(0, 0, 0),
                               b_expression=b"1.7", rhs_expression=rhs, use_direct_solver_basis=1)
        bb = msfec.BasisBuilder(p, device=0).run(cells)
        prob0 = mo.Problem(pairing=pairing, n_refine_local=0, a_freq=(0, 0, 0), a_scale=(2.0, 3.0, 0.5),
                           a_rotate=False, b_expr="1.7", rhs_expr=rhs.decode())
        for c in (0, 7):
            M0, r0, *_ = mo.build_basis(prob0, cells[c], c)
            assert rel_err(bb.get_global_element_matrix()[c], M0) < TOL


@pytest.mark.parametrize("solver", SOLVERS)
def test_structure_properties_at_scale(msfec, solver):
    """Size-independent properties on 1024 random-field cells (no oracle solve needed)."""
    cells = mo.morton_cells(4)[:1024]
    p = lib_problem(msfec, "NED_RT", 3, random_seed=20261017, solver=msfec.SOLVER[solver])
    bb = msfec.BasisBuilder(p, device=0).run(cells)
    M = bb.get_global_element_matrix()
    s = np.abs(M).max(axis=(1, 2))
    assert np.isfinite(M).all()
    assert (np.abs(M[:, :12, :12] - M[:, :12, :12].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    assert (np.abs(M[:, 12:, 12:] - M[:, 12:, 12:].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    assert (np.abs(M[:, :12, 12:] + M[:, 12:, :12].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    # a few cells against the oracle
    prob = oracle_problem("NED_RT", 3, random_seed=20261017)
    worst = _check_cells(bb, prob, cells, np.arange(1024), (0, 517, 1023))
    print("C5-style worst", worst, bb.stats)
    assert worst < TOL


def test_full_size_c5_properties(msfec):
    """BASELINE.json configs[4] at full size (32 768 random-field cells x 3 local refinements, direct path):
    size-independent properties over ALL cells, batch independence (a cell solved inside the full job equals the
    same cell solved in a job of its own and in a ragged 33-cell job) and three cells against the oracle."""
    cells = mo.morton_cells(5)
    assert len(cells) == 32768
    p = lib_problem(msfec, "NED_RT", 3, random_seed=20261017)               # solver: auto -> multifrontal
    ids = np.arange(32768)
    bb = msfec.BasisBuilder(p, device=0).run(cells, cell_ids=ids)
    M = bb.get_global_element_matrix().copy(); r = bb.get_global_element_rhs().copy()
    st = dict(bb.stats)
    assert st["solver"] == msfec.SOLVER_STAT["mf"]
    assert st["not_converged"] == 0 and st["residual_max"] < 1e-11          # true residual of every cell
    s = np.abs(M).max(axis=(1, 2))
    assert np.isfinite(M).all() and np.isfinite(r).all() and (s > 0).all()
    assert (np.abs(M[:, :12, :12] - M[:, :12, :12].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    assert (np.abs(M[:, 12:, 12:] - M[:, 12:, 12:].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    assert (np.abs(M[:, :12, 12:] + M[:, 12:, :12].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    # the (0,0) block is a Gram matrix of the A^-1 inner product: positive diagonal
    assert (np.einsum("cii->ci", M[:, :12, :12]) > 0).all()
    # batch independence: ids select the random field, so sub-jobs reproduce the same cells
    for sel in (np.array([20000]), np.arange(4090, 4123)):
        bb2 = msfec.BasisBuilder(p, device=0).run(cells[sel], cell_ids=ids[sel])
        M2 = bb2.get_global_element_matrix()
        assert np.abs(M2 - M[sel]).max() <= 1e-11 * s[sel].max()
        bb2.close()
    prob = oracle_problem("NED_RT", 3, random_seed=20261017)
    # 64 cells stratified over the Morton range (one per 512-cell stratum, pseudo-random offset) against the oracle
    rng = np.random.default_rng(5)
    sample = [0, 32767] + [int(512 * i + rng.integers(512)) for i in range(1, 63)]
    for c in sample:
        Mo, ro, *_ = mo.build_basis(prob, cells[c], c)
        assert rel_err(M[c], Mo) < TOL and np.abs(r[c] - ro).max() <= TOL * max(np.abs(ro).max(), 1e-300), c
    # the banded solver on one stratum gives the same matrices (two independent factorisations)
    sel = np.arange(8192, 8192 + 256)
    bb3 = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 3, random_seed=20261017, solver=msfec.SOLVER["band"]), device=0).run(cells[sel], ids[sel])
    assert np.abs(bb3.get_global_element_matrix() - M[sel]).max() <= 1e-10 * s[sel].max()
    bb3.close()
    bb.close()


def test_error_paths(msfec):
    p = lib_problem(msfec, "Q", 2)
    bb = msfec.BasisBuilder(p, device=0)
    cells = mo.morton_cells(1).copy()
    cells[3, 5, 1] += 0.01          # not a cube any more
    with pytest.raises(msfec.MsfecError) as e:
        bb.run(cells)
    assert e.value.code == 1
    with pytest.raises(msfec.MsfecError):
        bb.set_global_weights(np.zeros((3, 8)))      # wrong cell count


def test_device_pointer_entry(msfec):
    import torch
    cells = mo.morton_cells(2)
    p = lib_problem(msfec, "Q", 3)
    bb = msfec.BasisBuilder(p, device=0)
    dc = torch.tensor(cells, dtype=torch.float64, device="cuda:0").contiguous()
    dM = torch.empty((64, 8, 8), dtype=torch.float64, device="cuda:0")
    dr = torch.empty((64, 8), dtype=torch.float64, device="cuda:0")
    bb.run_device(64, dc.data_ptr(), 0, dM.data_ptr(), dr.data_ptr())
    torch.cuda.synchronize()
    host = msfec.BasisBuilder(p, device=0).run(cells)
    assert rel_err(dM.cpu().numpy(), host.get_global_element_matrix()) < 1e-12


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_reference_prm_configs_full_size(msfec, pairing):
    """BASELINE configs[0..3] at their real size and VERBATIM: the reference's prm_*_test-01.prm (64 coarse cells,
    4 local refinements, n = 16, `use direct solver basis = false` as shipped) with no override; the automatic solver
    selection takes the multifrontal LDL^T (chains of narrow fronts at the top separators).  Two cells against the
    oracle, structure on all; the banded block LDL^T on the same cells agrees to 1e-10."""
    cells = mo.morton_cells(2)
    p = msfec.problem_from_prm(os.path.join(ROOT, "examples", "prm", {"Q": "prm_q_test-01.prm", "Q_NED": "prm_q_ned_test-01.prm",
                               "NED_RT": "prm_ned_rt_test-01.prm", "RT_DQ": "prm_rt_dq_test-01.prm"}[pairing]), pairing)
    assert p.n_refine_local == 4 and p.use_direct_solver_basis == 0 and p.solver == 0
    bb = msfec.BasisBuilder(p, device=0).run(cells, np.arange(64))
    assert bb.stats["solver"] == msfec.SOLVER_STAT["mf"] and bb.stats["not_converged"] == 0 and bb.stats["residual_max"] < 1e-10
    p.solver = msfec.SOLVER["band"]
    bb2 = msfec.BasisBuilder(p, device=0).run(cells[:32], np.arange(32))
    assert bb2.stats["solver"] == msfec.SOLVER_STAT["band"] and bb2.stats["residual_max"] < 1e-10
    assert rel_err(bb2.get_global_element_matrix(), bb.get_global_element_matrix()[:32]) < 1e-10
    bb2.close()
    M = bb.get_global_element_matrix()
    assert np.isfinite(M).all()
    k0 = {"Q": 8, "Q_NED": 8, "NED_RT": 12, "RT_DQ": 6}[pairing]
    s = np.abs(M).max(axis=(1, 2))
    assert (np.abs(M[:, :k0, :k0] - M[:, :k0, :k0].transpose(0, 2, 1)).max(axis=(1, 2)) < 1e-9 * s).all()
    worst = _check_cells(bb, oracle_problem(pairing, 4), cells, np.arange(64), (37,) if pairing == "NED_RT" else (5, 37))
    print(pairing, "L=4 worst", worst, bb.stats)
    assert worst < TOL


@pytest.mark.parametrize("solver", ["auto", "mf", "minres"])
@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_one_local_refinement(msfec, pairing, solver):
    """Smallest local problems (n = 2: one 2 x 2 x 2 box, the multifrontal tree is a single root front without reached
    unknowns), ragged batch of 40 cells, rough random field."""
    cells = mo.morton_cells(2)[:40]
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, 1, random_seed=7, solver=msfec.SOLVER[solver]), device=0).run(cells, np.arange(40))
    worst = _check_cells(bb, oracle_problem(pairing, 1, random_seed=7), cells, np.arange(40), (0, 33, 39))
    assert worst < TOL and bb.stats["residual_max"] < 1e-10
    assert bb.stats["solver"] == (0 if solver == "minres" else 2)


@pytest.mark.parametrize("pairing", ["Q", "RT_DQ", "Q_NED", "NED_RT"])
def test_minres_full_size(msfec, pairing):
    """The batched MINRES path (msfec_problem.solver = MSFEC_SOLVER_MINRES; the memory fallback of the automatic selection)
    at n = 16, i.e. BASELINE configs 1-4 at their shipped size (the mixed H1-H(curl) / H(curl)-H(div) pairings need several
    thousand iterations with the diagonal preconditioner: 32 cells keep the test short)."""
    cells = mo.morton_cells(2)[:32]
    p = lib_problem(msfec, pairing, 4, solver=msfec.SOLVER["minres"])
    bb = msfec.BasisBuilder(p, device=0).run(cells, np.arange(32))
    worst = _check_cells(bb, oracle_problem(pairing, 4), cells, np.arange(32), (5,))
    print(pairing, "L=4 minres worst", worst, bb.stats["iterations_max"])
    assert worst < TOL and bb.stats["solver"] == 0


def test_bench_line_contract(tmp_path):
    """bench.py prints exactly ONE JSON line on stdout with the keys the driver reads (small workload)."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cells", "96", "--steps", "1", "--warmup", "3",
                        "--cpu-sample", "16"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["unit"] == "coarse cells/s" and d["value"] > 0 and d["gpu_launches"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert "workload" in d["config"] and d["krylov"]["residual_max"] < 1e-10


@pytest.mark.parametrize("pairing", ["Q", "NED_RT"])
def test_five_local_refinements_direct(msfec, pairing):
    """Largest local size (n = 32: 35 937 / 206 976 fine DoFs per cell).  The layer/plane band of Ned_RT exceeds 2^31 entries
    per cell there; the nested-dissection plan fits.  Checked by size-independent properties: true residual of every cell,
    partition of unity (rows of the Q matrix sum to zero), the symmetry pattern of M, independence of the batch composition."""
    cells = mo.morton_cells(2)[:3]
    ids = np.arange(3)
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, 5, use_direct_solver_basis=1), device=0).run(cells, ids)
    st = bb.stats
    assert st["not_converged"] == 0 and st["residual_max"] < 1e-10
    M = bb.get_global_element_matrix().copy()
    k0 = 8 if pairing == "Q" else 12
    scale = np.abs(M).max()
    assert np.abs(M[:, :k0, :k0] - M[:, :k0, :k0].transpose(0, 2, 1)).max() < 1e-11 * scale
    if pairing == "Q":
        assert np.abs(M.sum(2)).max() < 1e-10 * scale
    else:
        assert np.abs(M[:, k0:, k0:] - M[:, k0:, k0:].transpose(0, 2, 1)).max() < 1e-11 * scale
        assert np.abs(M[:, :k0, k0:] + M[:, k0:, :k0].transpose(0, 2, 1)).max() < 1e-11 * scale
    bb2 = msfec.BasisBuilder(lib_problem(msfec, pairing, 5, use_direct_solver_basis=1), device=0).run(cells[2:3], ids[2:3])
    assert np.abs(bb2.get_global_element_matrix()[0] - M[2]).max() < 1e-12 * scale


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_zero_local_refinements_give_the_standard_element_matrices(msfec, pairing):
    """0 local refinements: no interior unknowns (RT_DQ: only the pinned one), no local solve -- the coarse matrices are the
    standard lowest-order element matrices of the cell (what the fine-grid comparator XStd of the host driver assembles);
    against the oracle, rough random field and sine coefficients, ragged batch."""
    cells = mo.morton_cells(2)[:45]
    ids = np.arange(45)
    for seed in (0, 20261017):
        prob = oracle_problem(pairing, 0, random_seed=seed)
        bb = msfec.BasisBuilder(lib_problem(msfec, pairing, 0, random_seed=seed), device=0).run(cells, ids)
        assert bb.stats["solver"] == 3 and bb.stats["not_converged"] == 0
        M = bb.get_global_element_matrix(); r = bb.get_global_element_rhs()
        for c in (0, 17, 44):
            Mo, ro = mo.build_basis(prob, cells[c], c)[:2]
            assert rel_err(M[c], Mo) < 1e-12
            assert np.abs(r[c] - ro).max() <= 1e-12 * max(np.abs(ro).max(), 1.0)
        bb.close()
