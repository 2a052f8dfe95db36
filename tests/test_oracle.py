"""CPU tests of the oracle itself: the reference's only golden data at this boundary plus the
analytic invariants of SURVEY.md section 8(c)."""
import json
import os

import numpy as np
import pytest

from common import ROOT, oracle_problem, rel_err
from oracle import msfec_oracle as mo


def test_nedelec_projection_matches_reference_golden():
    """test/test_fe_projection_nedelec_mpi.mpirun=1.output: projecting (1,2,3) on an 8^3 mesh gives
    648 x {0.125, 0.25, 0.375}: Nedelec DoF = tangential component x edge length."""
    dig = json.load(open(os.path.join(ROOT, "tests", "golden", "nedelec_projection_digest.json")))
    v = mo.project_constant_on_nedelec(8, dig["field"])
    assert v.size == dig["n_values"]
    vals, counts = np.unique(np.round(v, 12), return_counts=True)
    assert {repr(float(a)): int(c) for a, c in zip(vals, counts)} == dig["histogram"]


def test_golden_fixture_regression():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_elem_matrices.npz"))
    cells = mo.morton_cells(2)
    for p in mo.PAIRINGS:
        prob = oracle_problem(p, 2)
        M, r, *_ = mo.build_basis(prob, cells[37], 37)
        assert rel_err(M, gold[f"{p}_M_37"]) < 1e-12
        assert np.abs(r - gold[f"{p}_r_37"]).max() <= 1e-12 * max(1.0, np.abs(r).max())


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_constant_coefficient_reproduces_standard_basis(pairing):
    """Invariant 3: with a constant diagonal A and constant B the multiscale basis is the standard
    coarse basis, so M(L) equals the analytic coarse element matrix M(L=0)
    (what set_to_std checks, ned_rt_basis.cc:1312-1335)."""
    cells = mo.morton_cells(1)
    out = []
    for L in (0, 2):
        prob = mo.Problem(pairing=pairing, n_refine_local=L, a_freq=(0, 0, 0), a_scale=(2.0, 3.0, 0.5),
                          a_rotate=False, b_expr="1.7",
                          rhs_expr="1;2;3" if pairing in ("Q_NED", "NED_RT") else "1")
        M, r, *_ = mo.build_basis(prob, cells[3], 3)
        out.append((M, r))
    assert rel_err(out[1][0], out[0][0]) < 1e-12
    assert np.abs(out[1][1] - out[0][1]).max() < 1e-13


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_structure_invariants(pairing):
    """Invariant 4: symmetry of the diagonal blocks, M01 = -M10^T, Q rows sum to zero,
    RT_DQ flux column = -/+ 1 (div RT_i = -/+ 1/|K|, rt_dq_basis.cc:517-537)."""
    cells = mo.morton_cells(2)
    M, r, X0, X1, cs = mo.build_basis(oracle_problem(pairing, 2), cells[21], 21)
    k0 = {"Q": 8, "Q_NED": 8, "NED_RT": 12, "RT_DQ": 6}[pairing]
    s = np.abs(M).max()
    assert np.abs(M[:k0, :k0] - M[:k0, :k0].T).max() < 1e-12 * s
    if pairing == "Q":
        assert np.abs(M.sum(1)).max() < 1e-12 * s
        assert np.abs(X0.sum(0) - 1.0).max() < 1e-12        # partition of unity
    else:
        assert np.abs(M[k0:, k0:] - M[k0:, k0:].T).max() < 1e-11 * s
        assert np.abs(M[:k0, k0:] + M[k0:, :k0].T).max() < 1e-12 * s
    if pairing == "RT_DQ":
        assert np.allclose(M[:6, 6], [1, -1, 1, -1, 1, -1], atol=1e-12)
        assert M[6, 6] == 0.0


def test_expression_parser():
    e = mo.Expr("scale*(2*x-1)*(y^2-y)*(z^2-z)", {"scale": 100})
    pts = np.random.default_rng(0).random((5, 3))
    x, y, z = pts.T
    assert np.allclose(e(pts), 100 * (2 * x - 1) * (y ** 2 - y) * (z ** 2 - z))
    assert mo.Expr("-x^2")(pts)[0] == -(x[0] ** 2)
    assert mo.Expr("2^3^2")(pts)[0] == 512.0
    assert np.allclose(mo.Expr("1/(1.0 - 0.9 * sin(2*pi*14*x))")(pts), 1 / (1 - 0.9 * np.sin(2 * np.pi * 14 * x)))
    with pytest.raises(ValueError):
        mo.Expr("foo(x)")


def test_random_field_is_partition_independent():
    a = mo.random_field_normals(20261017, np.arange(1000, dtype=np.uint64))
    b = mo.random_field_normals(20261017, np.arange(500, 1000, dtype=np.uint64))
    assert np.array_equal(a[500:], b)
    assert abs(a.mean()) < 0.1 and abs(a.std() - 1) < 0.1


def test_morton_partition():
    cells = mo.morton_cells(2)
    assert cells.shape == (64, 8, 3)
    # first 8 cells in z-order fill the first octant
    assert cells[:8, :, :].max() <= 0.5 + 1e-15
    chunks = [mo.partition(64, r, 3) for r in range(3)]
    assert chunks[0][0] == 0 and chunks[-1][1] == 64 and all(chunks[i][1] == chunks[i + 1][0] for i in range(2))


@pytest.mark.parametrize("seed", [0, 20261017])
def test_compiled_cpu_port_matches_oracle(seed):
    """oracle/msfec_cpu.cpp (the compiled CPU baseline of bench.py, an independent restatement in another language):
    its exact variant reproduces the Python oracle's Ned_RT element matrices to round-off, its reference-shaped variant
    (per right-hand side Schur-complement CG + GMRES(ILU(0)) at the reference's 1e-6, ned_rt_basis.cc:637-847) to the
    accuracy that tolerance gives."""
    from common import oracle_problem
    from oracle import msfec_cpu as mc
    prob = oracle_problem("NED_RT", 2, random_seed=seed)
    cells = mo.morton_cells(2)
    sel = np.array([5, 37, 63])
    ref = [mo.build_basis(prob, cells[c], int(c))[:2] for c in sel]
    M, r, its = mc.build_basis_ned_rt(prob, cells[sel], sel, "exact", 2)
    for i in range(3):
        assert np.abs(M[i] - ref[i][0]).max() <= 1e-12 * np.abs(ref[i][0]).max()
        assert np.abs(r[i] - ref[i][1]).max() <= 1e-12 * np.abs(ref[i][1]).max()
    M, r, its = mc.build_basis_ned_rt(prob, cells[sel], sel, "reference", 2)
    for i in range(3):
        assert np.abs(M[i] - ref[i][0]).max() <= 1e-5 * np.abs(ref[i][0]).max()
        assert np.abs(r[i] - ref[i][1]).max() <= 1e-4 * np.abs(ref[i][1]).max()
    assert (its[:, 0] >= 18).all() and (its[:, 1] > its[:, 0]).all()      # 18 outer solves, inner GMRES inside each S vmult


def test_subcomplex_commuting_property():
    """SURVEY.md s.8(c) invariant 6 (reference doc/pages/mainpage.dox:94-130: the multiscale spaces form a sub-complex).
    With rough coefficients, to round-off:
      Ned_RT:  curl of multiscale Nedelec basis m  =  sum_j C[j, m] * multiscale Raviart-Thomas basis j
      Q_Ned:   grad of multiscale Q1 basis m       =  sum_j G[j, m] * multiscale Nedelec basis j
    with C, G the curl / gradient incidence of ONE coarse cell in deal.II local order.  This ties the Nedelec and RT
    sign / orientation conventions, the boundary lifting, the basis-specific volume right-hand sides and the saddle-point
    solves together -- none of it holds if one convention is off."""
    import complex_ops as co
    from common import oracle_problem
    cells = mo.morton_cells(2)
    for seed in (0, 20261017):
        prob = oracle_problem("NED_RT", 2, random_seed=seed)
        g = mo.fine_grid(prob.n)
        M, r, X0, X1, cs = mo.build_basis(prob, cells[37], 37)
        curl_sigma = (co.curl(g) @ X0[:12].T).T                       # [12, nF]
        pred = co.coarse_incidence(co.curl).T @ X1[12:]               # [12, nF]
        assert np.abs(curl_sigma).max() > 0
        assert np.abs(curl_sigma - pred).max() <= 1e-11 * np.abs(curl_sigma).max()
        prob = oracle_problem("Q_NED", 2, random_seed=seed)
        M, r, X0, X1, cs = mo.build_basis(prob, cells[37], 37)
        grad_sigma = (co.gradient(g) @ X0[:8].T).T                    # [8, nE]
        pred = co.coarse_incidence(co.gradient).T @ X1[8:]            # [8, nE]
        assert np.abs(grad_sigma - pred).max() <= 1e-11 * np.abs(grad_sigma).max()


def test_refinement_self_consistency():
    """SURVEY.md s.8(c) invariant 5: for a coefficient the fine grids resolve, the coarse element matrices of
    successive local refinements converge (differences shrink by > 2x per level)."""
    cells = mo.morton_cells(1)
    for pairing, rhs in (("Q", "1"), ("NED_RT", "1;2;3")):
        Ms = []
        for L in (1, 2, 3, 4 if pairing == "Q" else 3):
            prob = mo.Problem(pairing=pairing, n_refine_local=L, a_freq=(1, 1, 1), a_alpha=(0.5, 0.4, 0.3), a_rotate=True,
                              b_expr="1 + 0.5*sin(2*pi*x)", rhs_expr=rhs)
            Ms.append(mo.build_basis(prob, cells[5], 5)[0])
        d = [np.abs(Ms[i + 1] - Ms[i]).max() for i in range(2)]
        assert d[1] < 0.5 * d[0] and d[1] < 0.05 * np.abs(Ms[2]).max(), (pairing, d)
