"""GPU tests of the size-independent invariants of SURVEY.md s.8(c) through the CUDA path (C ABI), and of context re-use."""
import numpy as np
import pytest

import complex_ops as co
from common import KINDS, lib_problem, oracle_problem, perm_to_oracle, rel_err
from oracle import msfec_oracle as mo

pytestmark = pytest.mark.gpu


def _all_bases(bb, prob, pairing, cell):
    """[k, N0], [k, N1] basis functions of one cell in ORACLE numbering"""
    k = bb.k - (1 if pairing == "RT_DQ" else 0)
    p0 = perm_to_oracle(bb, prob, 0, KINDS[pairing][0]); p1 = perm_to_oracle(bb, prob, 1, KINDS[pairing][1])
    X0 = np.zeros((k, len(p0))); X1 = np.zeros((k, len(p1)))
    for j in range(k):
        b0, b1 = bb.get_basis(cell, j)
        X0[j, p0] = b0; X1[j, p1] = b1
    return X0, X1


@pytest.mark.parametrize("solver", ["mf", "band", "minres"])
def test_subcomplex_commuting_property_on_gpu(msfec, solver):
    """Invariant 6 (reference doc/pages/mainpage.dox:94-130) on the device results, rough random field, 3 local refinements:
    curl(ms-Nedelec_m) = sum_j C[j, m] ms-RT_j (Ned_RT) and grad(ms-Q1_m) = sum_j G[j, m] ms-Ned_j (Q_Ned)."""
    cells = mo.morton_cells(2)[30:40]
    ids = np.arange(30, 40)
    for pairing, op, k0 in (("NED_RT", co.curl, 12), ("Q_NED", co.gradient, 8)):
        L = 3 if solver != "minres" else 2
        prob = oracle_problem(pairing, L, random_seed=20261017)
        bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L, random_seed=20261017, solver=msfec.SOLVER[solver]), device=0).run(cells, ids)
        g = mo.fine_grid(prob.n)
        X0, X1 = _all_bases(bb, prob, pairing, 7)
        d_sigma = (op(g) @ X0[:k0].T).T
        pred = co.coarse_incidence(op).T @ X1[k0:]
        assert np.abs(d_sigma).max() > 0
        assert np.abs(d_sigma - pred).max() <= 1e-9 * np.abs(d_sigma).max(), (pairing, solver)
        bb.close()


def test_iterative_converges_to_factorisation(msfec):
    """Invariant 5, second half: the batched MINRES result tends to the exact factorisation's as its tolerance goes to 0."""
    cells = mo.morton_cells(2)[:32]
    ids = np.arange(32)
    ref = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 2, solver=msfec.SOLVER["mf"]), device=0).run(cells, ids)
    Mref = ref.get_global_element_matrix().copy()
    errs = []
    for rtol in (1e-3, 1e-6, 1e-9, 1e-13):
        bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 2, solver=msfec.SOLVER["minres"], krylov_rtol=rtol), device=0).run(cells, ids)
        errs.append(rel_err(bb.get_global_element_matrix(), Mref))
        bb.close()
    print("MINRES vs factorisation:", errs)
    assert errs[0] > errs[1] > errs[2] > errs[3] and errs[3] < 1e-9 and errs[0] < 1e-1


def test_refinement_self_consistency_on_gpu(msfec):
    """Invariant 5, first half, through the device path: element matrices of successive local refinements converge
    (Q up to 5 local refinements, whichever factorisation the automatic selection takes)."""
    cells = mo.morton_cells(1)
    Ms = []
    for L in (1, 2, 3, 4, 5):
        p = msfec.make_problem("Q", n_refine_local=L, a_freq=(1, 1, 1), a_alpha=(0.5, 0.4, 0.3), a_rotate=1, rhs_expression=b"1")
        bb = msfec.BasisBuilder(p, device=0).run(cells[5:6], np.array([5]))
        assert bb.stats["solver"] in (1, 2) and bb.stats["not_converged"] == 0 and bb.stats["residual_max"] < 1e-10
        Ms.append(bb.get_global_element_matrix()[0].copy())
        bb.close()
    d = [np.abs(Ms[i + 1] - Ms[i]).max() for i in range(4)]
    assert all(d[i + 1] < 0.5 * d[i] for i in range(3)), d


@pytest.mark.parametrize("solver", ["mf", "band", "minres"])
def test_context_reuse_with_fewer_cells(msfec, solver):
    """A second, smaller build on the same context: set_weights / solution_norms / get_* follow the LAST build (the store
    keeps its capacity but not the old cell count)."""
    cells = mo.morton_cells(2)
    p = lib_problem(msfec, "NED_RT", 2, solver=msfec.SOLVER[solver])
    bb = msfec.BasisBuilder(p, device=0).run(cells, np.arange(64))
    M64 = bb.get_global_element_matrix().copy()
    w = np.random.default_rng(7).standard_normal((64, 18))
    bb.set_global_weights(w)
    n64 = bb.solution_norms(64)
    bb.run(cells[32:], np.arange(32, 64))
    assert rel_err(bb.get_global_element_matrix(), M64[32:]) < 1e-12
    with pytest.raises(msfec.MsfecError):
        bb.get_fine_solution(0)                       # weights of the old build are gone
    with pytest.raises(msfec.MsfecError):
        bb.set_global_weights(np.ones((64, 18)))      # cell count of the old build
    bb.set_global_weights(w[32:])
    n32 = bb.solution_norms(32)
    assert np.allclose(n32, n64[32:], rtol=1e-9, atol=0)
    with pytest.raises(msfec.MsfecError):
        bb.get_basis(40, 0)                           # beyond the current build
    with pytest.raises(msfec.MsfecError):
        bb.solution_norms(64)
    # and a larger one again
    bb.run(cells, np.arange(64))
    assert rel_err(bb.get_global_element_matrix(), M64) < 1e-12
    bb.close()


def test_streamed_children_match_gathered_children(msfec, monkeypatch):
    """k_mf_forward streams the children's contribution blocks of the level-1 fronts through shared memory with bulk copies
    (cp.async.bulk + mbarrier ring, mf.cuh STG); MSFEC_MF_STAGED=0 gathers them element by element, MSFEC_MF_STAGED=2 streams every
    level that fits.  All three sum the children in the same order, so the element matrices agree to round-off, and each
    reproduces the oracle."""
    cells = mo.morton_cells(2)[:48]
    ids = np.arange(48)
    prob = oracle_problem("NED_RT", 3, random_seed=20261017)
    Mo = mo.build_basis(prob, cells[11], 11)[0]
    out = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("MSFEC_MF_STAGED", mode)
        bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 3, random_seed=20261017, solver=msfec.SOLVER["mf"]), device=0).run(cells, ids)
        out[mode] = bb.get_global_element_matrix().copy()
        assert bb.stats["solver"] == 2 and bb.stats["residual_max"] < 1e-10
        assert rel_err(out[mode][11], Mo) < 1e-9
        bb.close()
    assert rel_err(out["1"], out["0"]) < 1e-12 and rel_err(out["2"], out["0"]) < 1e-12


def test_multifrontal_sub_batches(msfec, monkeypatch):
    """The multifrontal solver works through a resident batch in sub-batches sized by its storage budget (factor records,
    contribution arena, padded solution are re-used from one sub-batch to the next).  Forcing 64-cell sub-batches on a ragged
    200-cell build must not change a single bit of the element matrices."""
    cells = np.concatenate([mo.morton_cells(2), mo.morton_cells(2), mo.morton_cells(2), mo.morton_cells(2)[:8]])
    ids = np.arange(len(cells)) % 64
    out = []
    for sub in (None, "64"):
        if sub:
            monkeypatch.setenv("MSFEC_MF_BATCH", sub)
        bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 3, random_seed=3, solver=msfec.SOLVER["mf"]), device=0).run(cells, ids)
        assert bb.stats["residual_max"] < 1e-10 and bb.stats["not_converged"] == 0
        out.append((bb.get_global_element_matrix().copy(), bb.get_global_element_rhs().copy(), bb.stats["mf_launches"]))
        bb.close()
    assert out[1][2] == 4 * out[0][2]                      # 4 sub-batches instead of 1
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0][:64], out[0][0][64:128])   # the same cells again give the same bits
