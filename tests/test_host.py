"""CPU tests of the host side of libmsfec_b200.so: ABI surface, .prm reader, shared topology and
assembly/operator tables (validated by a numpy emulation of the device pipeline against the oracle).
No compute entry point is exercised here; those need a GPU (tests/test_gpu_parity.py)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import emulate
from common import KINDS, ROOT, lib_problem, oracle_problem, perm_to_oracle, prm_path, rel_err
from oracle import msfec_oracle as mo


def test_library_exports_every_declared_symbol(msfec):
    hdr = open(os.path.join(ROOT, "include", "msfec.h")).read()
    declared = set(re.findall(r"\b(msfec_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"msfec_ctx"}
    assert declared, "no declarations parsed"
    lib = msfec.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/msfec.h but not exported"
    assert set(msfec.EXPORTS) == declared
    assert lib.msfec_abi_version() == 2


def test_struct_layout_matches_header(msfec):
    # sizeof checks guard the ctypes mirror against drift
    assert C.sizeof(msfec.Problem) == 4 * 6 + 4 * 3 + 4 + 8 * 3 * 2 + 8 * 2 + 8 * 3 + 8 + 8 + 8 + 4 + 4 + 4 + 4
    assert C.sizeof(msfec.Stats) == 4 * 8 + 8 * 9 + 8 + 8 + 8 * 3 + 4 + 4 + 8 * 5 + 8
    # field order of the ctypes mirrors == declaration order in include/msfec.h
    hdr = open(os.path.join(ROOT, "include", "msfec.h")).read()
    for name, cls in (("msfec_problem", msfec.Problem), ("msfec_stats", msfec.Stats)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        decl = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            names = re.sub(r"^(const\s+)?\w+\s+\*?", "", stmt)
            decl += [re.sub(r"\[.*?\]|\*|\s", "", x) for x in names.split(",")]
        assert decl == [f for f, _ in cls._fields_], (name, decl)


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_prm_reader_matches_oracle_reader(msfec, pairing):
    p = msfec.problem_from_prm(prm_path(pairing), pairing)
    o = mo.Problem.from_prm(prm_path(pairing), pairing)
    assert p.n_refine_local == o.n_refine_local == 4 and p.n_refine_global == 2
    assert tuple(p.a_freq) == o.a_freq and tuple(p.a_alpha) == o.a_alpha and bool(p.a_rotate) == o.a_rotate
    assert p.rhs_expression.decode() == o.rhs_expr
    if pairing in ("Q_NED", "NED_RT"):
        assert p.b_expression.decode() == o.b_expr and p.b_freq == 14


def test_prm_errors(msfec, tmp_path):
    bad = tmp_path / "bad.prm"
    bad.write_text("subsection A\n set x = 1\n")
    with pytest.raises(msfec.MsfecError) as e:
        msfec.problem_from_prm(str(bad), "Q")
    assert e.value.code == 7
    with pytest.raises(msfec.MsfecError):
        msfec.problem_from_prm(str(tmp_path / "missing.prm"), "Q")
    # manufactured-solution runs are not built: rejected, not silently changed
    exact = tmp_path / "exact.prm"
    txt = open(prm_path("NED_RT")).read()
    i = txt.index("subsection Multiscale method parameters")
    exact.write_text(txt[:i] + txt[i:].replace("set use exact solution = false", "set use exact solution = true", 1))
    with pytest.raises(msfec.MsfecError) as e:
        msfec.problem_from_prm(str(exact), "NED_RT")
    assert e.value.code == 1 and "exact solution" in str(e.value)


def test_shape_queries(msfec):
    lib = msfec.lib()
    assert [lib.msfec_k(i) for i in range(4)] == [8, 20, 18, 7]
    n0, n1 = C.c_int(), C.c_int()
    assert lib.msfec_n_fine_dofs(2, 3, C.byref(n0), C.byref(n1)) == 3672 and (n0.value, n1.value) == (1944, 1728)
    assert lib.msfec_n_fine_dofs(2, 4, C.byref(n0), C.byref(n1)) == 26928


def test_entity_counts_match_survey(msfec):
    """n=8: boundary E 768, F 384; nnz EE 53 400, FF 17 088, EF 31 488."""
    bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 3), device=-1)
    D = emulate.dims_of(bb)
    assert (D["N0"], D["N1"], D["N0"] - D["NI0"], D["N1"] - D["NI1"]) == (1944, 1728, 768, 384)
    assert 2 * D["n_slots0"] - D["N0"] == 53400 and 2 * D["n_slots1"] - D["N1"] == 17088
    # the coupling block keeps only numerically non-zero entries (structural count is 31 488)
    full_shared = len(bb.table("full.scol"))
    assert full_shared == 2 * 19200 and full_shared <= 2 * 31488


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
@pytest.mark.parametrize("L", [1, 2])
def test_tables_reproduce_oracle(msfec, pairing, L):
    """Slots, operators, boundary data and volume rhs built by csrc/topology.cpp give the oracle's
    coarse matrices when driven through the same arithmetic as the kernels."""
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L), device=-1)
    cells = mo.morton_cells(2)
    prob = oracle_problem(pairing, L)
    M, r, Z, _ = emulate.emulate_cell(bb, prob, cells[37], 37)
    Mo, ro, X0, X1, cs = mo.build_basis(prob, cells[37], 37)
    assert rel_err(M, Mo) < 1e-11
    assert np.abs(r - ro).max() <= 1e-12 * max(1.0, np.abs(ro).max())


def test_tables_random_field(msfec):
    bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 2, random_seed=20261017), device=-1)
    cells = mo.morton_cells(3)
    prob = oracle_problem("NED_RT", 2, random_seed=20261017)
    M, r, *_ = emulate.emulate_cell(bb, prob, cells[100], 100)
    Mo, ro, *_ = mo.build_basis(prob, cells[100], 100)
    assert rel_err(M, Mo) < 1e-11


def test_no_cpu_fallback(msfec):
    """Compute entry points must fail loudly without a device."""
    bb = msfec.BasisBuilder(lib_problem(msfec, "Q", 1), device=-1)
    with pytest.raises(msfec.MsfecError) as e:
        bb.run(mo.morton_cells(1))
    assert e.value.code == 2
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(msfec.MsfecError) as e2:
            msfec.BasisBuilder(lib_problem(msfec, "Q", 1), device=0)
        assert e2.value.code == 2


def test_invalid_arguments(msfec):
    p = lib_problem(msfec, "Q", 1)
    p.n_refine_local = -1          # 0 is valid: the standard basis (fine-grid comparator of the host driver)
    with pytest.raises(msfec.MsfecError) as e:
        msfec.BasisBuilder(p, device=-1)
    assert e.value.code == 1
    p = lib_problem(msfec, "NED_RT", 1, rhs_expression=b"1")      # needs 3 components
    with pytest.raises(msfec.MsfecError):
        msfec.BasisBuilder(p, device=-1)
    p = lib_problem(msfec, "Q", 1, rhs_expression=b"foo(x)")
    with pytest.raises(msfec.MsfecError) as e:
        msfec.BasisBuilder(p, device=-1)
    assert e.value.code == 7


@pytest.mark.parametrize("pairing,L", [("Q", 2), ("Q_NED", 2), ("NED_RT", 2), ("RT_DQ", 2), ("NED_RT", 3)])
def test_direct_plan_tables(msfec, pairing, L):
    """The direct solver's host plan (block ordering, symbolic block structure, band fill lists): a numpy LDL^T
    WITHOUT pivoting in the library's padded order reproduces the sparse-LU solution, pivots are positive on
    sigma-type and negative on u-type unknowns (the existence argument of DESIGN.md), padding pivots are 1."""
    seed = 20261017 if L == 3 else 0
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L, random_seed=seed), device=-1)
    cells = mo.morton_cells(2)
    prob = oracle_problem(pairing, L, random_seed=seed)
    M, r, Z, dbg = emulate.emulate_cell(bb, prob, cells[37], 37)
    D = emulate.dims_of(bb)
    h = (cells[37][7][0] - cells[37][0][0]) / D["n"]
    x, d, inv = emulate.emulate_direct(bb, dbg["vals"], h ** D["k_h_exponent"], dbg["b"])
    ref = dbg["x"]
    sel = np.arange(D["NI0"]) if pairing == "RT_DQ" else np.arange(D["NI"])     # RT_DQ: u is fixed up to a constant
    assert np.abs(x[sel] - ref[sel]).max() <= 1e-9 * np.abs(ref[sel]).max()
    real = inv >= 0
    is_u = np.zeros(len(inv), bool); is_u[real] = inv[real] >= D["NI0"]
    assert (d[real & ~is_u] > 0).all() and (d[real & is_u] < 0).all()
    pad = ~real
    assert np.isin(d[pad], (1.0, -1.0)).all() and (d[pad] == -1.0).sum() == (1 if pairing == "RT_DQ" else 0)
    # block structure sanity: every block column's front starts with itself; sizes are multiples of 32
    bs = bb.table("direct.bs"); ch_off = bb.table("direct.chunk_off"); ch_blk = bb.table("direct.chunk_blk")
    assert (bs % 32 == 0).all() and all(ch_blk[ch_off[s]] == s and ch_blk[ch_off[s + 1] - 1] == -1 for s in range(len(bs)))


@pytest.mark.parametrize("pairing", mo.PAIRINGS)
def test_norm_operator_tables_match_oracle_gram_matrices(msfec, pairing):
    """The Gram matrices msfec_solution_norms applies on the device (csrc/topology.cpp:build_norm_operators) equal the
    oracle's unit-coefficient mass / grad-grad / curl-curl / div-div matrices."""
    import scipy.sparse as sp
    import coarse_solve as cs
    L = 2
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L), device=-1)
    prob = oracle_problem(pairing, L)
    cells = mo.morton_cells(2)
    h = (cells[0][7][0] - cells[0][0][0]) / prob.n
    ref = cs.unit_norm_matrices(pairing, L, cells[0])
    hexp = bb.table("norm.h_exponents")
    names = {"Q": ("sigma_L2", "sigma_H1semi", None, None), "Q_NED": ("sigma_L2", "sigma_H1semi", "u_L2", "u_Hcurlsemi"),
             "NED_RT": ("sigma_L2", "sigma_Hcurlsemi", "u_L2", "u_Hdivsemi"), "RT_DQ": ("sigma_L2", "sigma_Hdivsemi", "cells", None)}[pairing]
    for i, name in enumerate(names):
        ptr = bb.table(f"norm{i}.ptr")
        if name is None:
            assert len(ptr) == 0 or ptr[-1] == 0
            continue
        n = len(ptr) - 1
        G = sp.csr_matrix((bb.table(f"norm{i}.val") * h ** float(hexp[i]), bb.table(f"norm{i}.col"), ptr), shape=(n, n))
        if name == "cells":
            assert abs(G - sp.identity(n) * h ** 3).max() < 1e-18
            continue
        p = perm_to_oracle(bb, prob, i // 2, KINDS[pairing][i // 2])
        Go = sp.csr_matrix(ref[name])[p][:, p]
        assert abs(G - Go).max() <= 1e-13 * abs(Go).max(), (pairing, name)
        assert abs(G - G.T).max() == 0.0


@pytest.mark.parametrize("pairing,n_blocks,n_padded", [("NED_RT", 20, 2816), ("RT_DQ", 20, 2176)])
def test_direct_plan_nested_dissection(msfec, monkeypatch, pairing, n_blocks, n_padded):
    """The nested-dissection block ordering (chosen automatically at 4 local refinements for Q_Ned / Ned_RT / RT_DQ, forced
    here at 3): the no-pivot LDL^T in the library's padded order reproduces the sparse-LU solution and pivots keep their
    signs (positive on sigma-type, negative on u-type unknowns).  RT_DQ: every box hands one cell DoF up to the plane
    that joins it with its sibling, otherwise the box interior is a pure-Neumann problem with a vanishing last pivot.
    Separator planes are split into quadrants (here 4 x 4 fine cells so that n = 8 exercises it; 8 x 8 in production) with
    the dividing lines attached to the earlier quadrant, in cyclic order -- the variant whose leading sets stay simply
    connected (a separate block for the lines makes Ned_RT pivots change sign)."""
    monkeypatch.setenv("MSFEC_DIRECT_ORDERING", "nd")
    monkeypatch.setenv("MSFEC_ND_SEP_PIECE", "4")
    L = 3
    seed = 20261017 if pairing == "NED_RT" else 0
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L, random_seed=seed), device=-1)
    info = bb.table("direct.info")
    assert (int(info[2]), int(info[3])) == (n_blocks, n_padded)   # 8 boxes of 4^3 fine cells + 4 + 2 x 2 + 4 plane pieces
    cells = mo.morton_cells(2)
    prob = oracle_problem(pairing, L, random_seed=seed)
    M, r, Z, dbg = emulate.emulate_cell(bb, prob, cells[37], 37)
    D = emulate.dims_of(bb)
    h = (cells[37][7][0] - cells[37][0][0]) / D["n"]
    x, d, inv = emulate.emulate_direct(bb, dbg["vals"], h ** D["k_h_exponent"], dbg["b"])
    ref = dbg["x"]
    sel = np.arange(D["NI0"]) if pairing == "RT_DQ" else np.arange(D["NI"])     # RT_DQ: u is fixed up to a constant
    assert np.abs(x[sel] - ref[sel]).max() <= 1e-9 * np.abs(ref[sel]).max()
    real = inv >= 0
    is_u = np.zeros(len(inv), bool); is_u[real] = inv[real] >= D["NI0"]
    assert (d[real & ~is_u] > 0).all() and (d[real & is_u] < 0).all()
    assert (d[~real] == -1.0).sum() == (1 if pairing == "RT_DQ" else 0)          # the pinned constant mode
    monkeypatch.delenv("MSFEC_DIRECT_ORDERING")
    monkeypatch.delenv("MSFEC_ND_SEP_PIECE")
    if pairing == "NED_RT":
        bb2 = msfec.BasisBuilder(lib_problem(msfec, pairing, L, random_seed=seed), device=-1)
        assert int(bb2.table("direct.info")[3]) == 2656          # layers/planes stay the default at n = 8
    bb4 = msfec.BasisBuilder(lib_problem(msfec, pairing, 4), device=-1)
    assert int(bb4.table("direct.info")[2]) == 132               # nested dissection chosen at n = 16 (127 + 3 + 2 pieces)


@pytest.mark.parametrize("pairing,L,seed", [("Q", 2, 0), ("Q_NED", 2, 0), ("NED_RT", 2, 0), ("RT_DQ", 2, 0), ("NED_RT", 1, 0), ("RT_DQ", 1, 0),
                                             ("Q", 3, 0), ("Q_NED", 3, 0), ("RT_DQ", 3, 0), ("NED_RT", 3, 20261017)])
def test_multifrontal_plan_tables(msfec, pairing, L, seed):
    """The multifrontal plan (csrc/mfplan.cpp: nested dissection to 2^3-cell boxes, supernodes, fronts, index maps,
    assembly lists) replayed in numpy front by front exactly as k_mf_forward / k_mf_backward walk it: the no-pivot
    LDL^T reproduces the sparse-LU solution, pivots are positive on sigma-type and negative on u-type unknowns, every
    front fits one SM's shared memory."""
    bb = msfec.BasisBuilder(lib_problem(msfec, pairing, L, random_seed=seed), device=-1)
    T = emulate.mf_tables(bb)
    assert T["info"]["feasible"]
    assert max(T["smem_fwd"]) <= 227 * 1024 and max(T["smem_bwd"]) <= 227 * 1024
    cells = mo.morton_cells(2)
    prob = oracle_problem(pairing, L, random_seed=seed)
    M, r, Z, dbg = emulate.emulate_cell(bb, prob, cells[37], 37)
    D = emulate.dims_of(bb)
    h = (cells[37][7][0] - cells[37][0][0]) / D["n"]
    x, d, inv = emulate.emulate_multifrontal(bb, dbg["vals"], h ** D["k_h_exponent"], dbg["b"])
    ref = dbg["x"]
    sel = np.arange(D["NI0"]) if pairing == "RT_DQ" else np.arange(D["NI"])     # RT_DQ: u is fixed up to a constant
    assert np.abs(x[sel] - ref[sel]).max() <= 1e-9 * np.abs(ref[sel]).max()
    real = inv >= 0
    is_u = np.zeros(len(inv), bool); is_u[real] = inv[real] >= D["NI0"]
    assert (d[real & ~is_u] > 0).all() and (d[real & is_u] < 0).all()
    assert (d[~real] == -1.0).sum() == (1 if pairing == "RT_DQ" else 0) and np.isin(d[~real], (1.0, -1.0)).all()
    # structure: children precede parents, a parent sits above its children, own columns of the children lead
    fr = T["fronts"]
    for f, F in enumerate(fr):
        assert F["s8"] % 8 == 0 and F["u8"] % 8 == 0 and F["m"] == F["s8"] + F["u8"] + T["info"]["kr"]
        assert F["ldx"] % 16 in (4, 12) and F["ldx"] >= F["s8"]
        if F["parent"] >= 0:
            assert F["parent"] > f and fr[F["parent"]]["level"] > F["level"]
    if L == 3 and pairing == "NED_RT":
        # C5: 34 MFLOP of exact elimination per cell (the 32-padded layer/plane band executes 420)
        assert T["info"]["flops"] < 80e6 and (T["info"]["l_doubles"] + T["info"]["c_doubles"]) * 8 < 10e6


def test_multifrontal_plan_limits(msfec):
    """At 4 local refinements the top separators are split into chains of narrow fronts so that every panel fits one SM's
    shared memory, and the contribution blocks share an arena about two tree levels deep; at 5 local refinements a Ned_RT
    front no longer fits at any width and the engine keeps the banded solver."""
    bb = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 4), device=-1)
    T = emulate.mf_tables(bb)
    info = T["info"]
    assert info["feasible"] and info["n_levels"] > 40
    assert max(T["smem_fwd"]) <= 227 * 1024 and max(T["smem_bwd"]) <= 227 * 1024
    total_c = sum((F["u8"] + info["kr"]) * F["u8"] for F in T["fronts"])
    assert info["c_doubles"] < total_c / 8 and (info["c_doubles"] + info["l_doubles"]) * 8 < 200e6
    # arena: no two live contribution blocks overlap (a block lives from its level to its parent's level)
    fr = T["fronts"]
    spans = [(F["c_off"], F["c_off"] + (F["u8"] + info["kr"]) * F["u8"], F["level"], fr[F["parent"]]["level"] if F["parent"] >= 0 else F["level"])
             for F in fr if F["u8"]]
    spans.sort()
    for i in range(len(spans)):
        for j in range(i + 1, len(spans)):
            if spans[j][0] >= spans[i][1]:
                break
            assert spans[j][2] > spans[i][3] or spans[i][2] > spans[j][3], (spans[i], spans[j])
    bb5 = msfec.BasisBuilder(lib_problem(msfec, "NED_RT", 5), device=-1)
    assert not emulate.mf_tables(bb5)["info"]["feasible"]
