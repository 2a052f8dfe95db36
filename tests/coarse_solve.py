"""Harness-side global coarse problem: assembles the coarse element
matrices of all coarse cells into the global system the reference's *Multiscale::assemble_system builds
(source/Ned_RT/ned_rt_global.cc:192-317; q_global.cc:127-139 zero Dirichlet data; q_ned_global.cc:175-207
zero essential data; rt_dq_global.cc:273-304 zero natural data), solves it with a sparse direct solver,
and returns the per-cell weights send_global_weights_to_cell would distribute (ned_rt_global.cc:464-488).
Also the fine-grid norms used to compare two multiscale solutions (the reference computes none).
Test infrastructure only."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import msfec_oracle as mo


def _cell_lex_index(corners, m):
    ijk = np.rint(np.asarray(corners)[:, 0, :] * m).astype(int)
    return ijk[:, 0] + m * (ijk[:, 1] + m * ijk[:, 2])


def coarse_dof_maps(pairing, g_ref, corners):
    """local -> global coarse DoF indices [n_cells, k] (sigma-type block first), sizes, boundary mask."""
    m = 1 << g_ref
    G = mo.fine_grid(m)                      # the coarse mesh has the topology of an m^3 grid
    lex = _cell_lex_index(corners, m)
    if pairing == "Q":
        return G.cV[lex], (G.nV, 0), G.v_bnd
    if pairing == "Q_NED":
        return np.concatenate([G.cV[lex], G.nV + G.cE[lex]], 1), (G.nV, G.nE), np.concatenate([G.v_bnd, G.e_bnd])
    if pairing == "NED_RT":
        return np.concatenate([G.cE[lex], G.nE + G.cF[lex]], 1), (G.nE, G.nF), np.zeros(G.nE + G.nF, bool)
    if pairing == "RT_DQ":
        return np.concatenate([G.cF[lex], G.nF + G.cC[lex]], 1), (G.nF, G.nC), np.zeros(G.nF + G.nC, bool)
    raise ValueError(pairing)


def solve_coarse(pairing, g_ref, corners, M, r):
    """Global weights per cell [n_cells, k] from coarse element matrices M[n,k,k] and rhs r[n,k]."""
    l2g, sizes, essential = coarse_dof_maps(pairing, g_ref, corners)
    n, k = l2g.shape
    N = sizes[0] + sizes[1]
    rows = np.repeat(l2g[:, :, None], k, 2).ravel(); cols = np.repeat(l2g[:, None, :], k, 1).ravel()
    A = sp.coo_matrix((M.ravel(), (rows, cols)), shape=(N, N)).tocsr()
    b = np.bincount(l2g.ravel(), r.ravel(), N)
    free = np.flatnonzero(~essential)
    x = np.zeros(N)                          # zero essential data
    x[free] = spla.splu(A[free][:, free].tocsc()).solve(b[free])
    return x[l2g]


def unit_norm_matrices(pairing, L, corners0):
    """Fine-grid Gram matrices of one coarse cell for L2 / H(grad|curl|div) (semi)norms, oracle numbering."""
    out = {}
    unit = dict(a_freq=(0, 0, 0), a_scale=(1.0, 1.0, 1.0), a_alpha=(1.0, 1.0, 1.0), a_rotate=False, b_expr="1")
    vec = "0;0;0"
    if pairing in ("Q", "Q_NED"):
        out["sigma_L2"] = mo.assemble_cell(mo.Problem(pairing="Q_NED", n_refine_local=L, rhs_expr=vec, **unit), corners0).A00
        out["sigma_H1semi"] = mo.assemble_cell(mo.Problem(pairing="Q", n_refine_local=L, rhs_expr="0", **unit), corners0).A00
    if pairing == "Q_NED":
        out["u_L2"] = mo.assemble_cell(mo.Problem(pairing="NED_RT", n_refine_local=L, rhs_expr=vec, **unit), corners0).A00
        out["u_Hcurlsemi"] = mo.assemble_cell(mo.Problem(pairing="Q_NED", n_refine_local=L, rhs_expr=vec, **unit), corners0).A11
    if pairing == "NED_RT":
        out["sigma_L2"] = mo.assemble_cell(mo.Problem(pairing="NED_RT", n_refine_local=L, rhs_expr=vec, **unit), corners0).A00
        out["sigma_Hcurlsemi"] = mo.assemble_cell(mo.Problem(pairing="Q_NED", n_refine_local=L, rhs_expr=vec, **unit), corners0).A11
        out["u_L2"] = mo.assemble_cell(mo.Problem(pairing="RT_DQ", n_refine_local=L, rhs_expr="0", **unit), corners0).A00
        out["u_Hdivsemi"] = mo.assemble_cell(mo.Problem(pairing="NED_RT", n_refine_local=L, rhs_expr=vec, **unit), corners0).A11
    if pairing == "RT_DQ":
        out["sigma_L2"] = mo.assemble_cell(mo.Problem(pairing="RT_DQ", n_refine_local=L, rhs_expr="0", **unit), corners0).A00
        out["sigma_Hdivsemi"] = mo.assemble_cell(mo.Problem(pairing="NED_RT", n_refine_local=L, rhs_expr=vec, **unit), corners0).A11
    return out
