"""The oracle against the REFERENCE's own code, run in the build container: eqn_coeff_A.cc (Diffusion_A,
DiffusionInverse_A: sine family, Euler rotation, .prm parsing through its own declare / parse calls), eqn_coeff_R.cc and
the coarse Q1 shape functions basis_q1.tpp / basis_q1_grad.tpp were compiled UNMODIFIED from /root/reference
(oracle/Makefile -> oracle/_ref/libmsfec_ref.so; only the deal.II headers are stand-ins, oracle/ref_shim) and evaluated
with the reference's shipped example_parameters/*.prm.  tests/golden/reference_compiled_eqdata.npz holds their outputs
(generator: tests/golden/make_reference_compiled_golden.py); where the library itself is present (it travels with the
snapshot) the same comparison runs live on fresh random points."""
import os

import numpy as np
import pytest

from common import ROOT, oracle_problem, prm_path
from oracle import msfec_oracle as mo, msfec_ref as mr

GOLD = os.path.join(ROOT, "tests", "golden", "reference_compiled_eqdata.npz")
TOL = 1e-14   # relative; both sides evaluate the same closed forms in FP64


def _close(a, b):
    return np.abs(a - b).max() <= TOL * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_oracle_coefficients_match_compiled_reference_golden(pairing):
    gold = np.load(GOLD)
    cells = mo.morton_cells(2)
    prob = oracle_problem(pairing, 2)
    for c in (5, 37):
        x0 = cells[c].min(0); H = float(cells[c].max(0)[0] - x0[0])
        A, Ainv, _, pts = mo.coefficient_fields(prob, x0, H)
        assert np.array_equal(pts, gold[f"{pairing}_pts_{c}"])
        assert _close(A, gold[f"{pairing}_A_{c}"])
        assert _close(Ainv, gold[f"{pairing}_Ainv_{c}"])
        # the reference's zero-order term is identically zero (eqn_coeff_R.cc): the oracle has no such term
        assert not gold[f"{pairing}_R_{c}"].any()
    # the shipped test-01 files do rotate: the golden really exercises the off-diagonal entries
    if prob.a_rotate:
        assert np.abs(gold[f"{pairing}_A_5"][..., 0, 1]).max() > 1e-3


def test_oracle_coarse_q1_matches_compiled_reference_golden():
    gold = np.load(GOLD)
    val, grad = mo.coarse_q1(gold["q1_x0"], float(gold["q1_H"]), gold["q1_pts"])
    # 1e-12: the reference gets the monomial coefficients from an 8x8 inverse of the vertex matrix
    assert np.abs(val - gold["q1_val"]).max() <= 1e-12
    assert np.abs(grad - gold["q1_grad"]).max() <= 1e-12 * np.abs(gold["q1_grad"]).max()


def test_oracle_coarse_nedelec_and_rt_scaling_matches_compiled_reference_golden():
    """MyMappingQ1 + BasisNedelec / BasisRaviartThomas of the reference (compiled; the unit-cell shape functions under them are
    stand-ins restating deal.II's documented lowest-order elements): the mapping to the unit cell, its Jacobian 1/H, and the
    transforms J^-T phi (Nedelec, 1/H) and J phi / det J (Raviart-Thomas, 1/H^2) that scale the boundary data of the local
    problems (a4) -- against the oracle's coarse_ned / coarse_rt on a non-unit cube."""
    gold = np.load(GOLD)
    x0, H, pts = gold["q1_x0"], float(gold["q1_H"]), gold["q1_pts"]
    assert np.abs((pts - x0) / H - gold["map_ref"]).max() < 1e-12
    assert np.abs(gold["map_inv_jac"] - np.eye(3) / H).max() < 1e-10 / H
    ned, _ = mo.coarse_ned(x0, H, pts)
    rt, _ = mo.coarse_rt(x0, H, pts)
    assert np.abs(ned - gold["ned_val"]).max() <= 1e-11 * np.abs(gold["ned_val"]).max()
    assert np.abs(rt - gold["rt_val"]).max() <= 1e-11 * np.abs(gold["rt_val"]).max()
    assert np.abs(gold["ned_val"]).max() > 0.9 / H and np.abs(gold["rt_val"]).max() > 0.9 / H ** 2   # the scaling is exercised


@pytest.mark.skipif(not mr.available(), reason="oracle/_ref/libmsfec_ref.so not built (needs /root/reference)")
def test_oracle_matches_compiled_reference_live():
    rng = np.random.default_rng(7)
    pts = rng.random((500, 3))
    for pairing in mo.PAIRINGS:
        prob = oracle_problem(pairing, 2)
        R = mo.rotation_matrix(prob.a_rotate)
        diag = np.stack([prob.a_scale[d] * (1.0 - prob.a_alpha[d] * np.sin(2 * np.pi * prob.a_freq[d] * pts[:, d]))
                         for d in range(3)], -1)
        for inverse in (False, True):
            mine = np.einsum("ia,pa,ja->pij", R, 1.0 / diag if inverse else diag, R)
            for pointwise in (False, True):
                assert _close(mine, mr.diffusion_a(prm_path(pairing), pts, inverse=inverse, pointwise=pointwise))
        assert not mr.reaction_rate(pts).any()
    # coarse Q1 on a random axis-aligned cube, vertices in deal.II order
    x0 = rng.random(3); H = 0.3
    vertices = x0 + H * np.array([[v & 1, (v >> 1) & 1, v >> 2] for v in range(8)], dtype=float)
    p = x0 + H * rng.random((100, 3))
    val, grad = mr.basis_q1(vertices, p)
    mv, mg = mo.coarse_q1(x0, H, p)
    assert np.abs(mv - val).max() <= 1e-12 and np.abs(mg - grad).max() <= 1e-12 * np.abs(grad).max()
    # Kronecker property at the vertices: pins the vertex order the oracle's q1_ref assumes
    assert np.abs(mr.basis_q1(vertices, vertices)[0] - np.eye(8)).max() < 1e-12
    # mapping and the covariant / Piola transforms, value() and value_list() paths
    ref, jac = mr.mapping(vertices, p)
    assert np.abs(ref - (p - x0) / H).max() < 1e-12 and np.abs(jac - np.eye(3) / H).max() < 1e-10 / H
    for kind, mine in (("ned", mo.coarse_ned(x0, H, p)[0]), ("rt", mo.coarse_rt(x0, H, p)[0])):
        for pointwise in (False, True):
            theirs = mr.basis_vector(vertices, kind, p, pointwise=pointwise)
            assert np.abs(mine - theirs).max() <= 1e-11 * np.abs(theirs).max()


def test_compiled_reference_reads_a_missing_prm_loudly():
    if not mr.available():
        pytest.skip("oracle/_ref/libmsfec_ref.so not built")
    with pytest.raises(RuntimeError):
        mr.diffusion_a("/nonexistent.prm", np.zeros((1, 3)))
