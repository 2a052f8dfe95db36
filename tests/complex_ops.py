"""Discrete exterior derivatives on the fine grid and on one coarse cell, in the oracle's numbering (test
infrastructure): gradient (vertices -> edges), curl (edges -> faces), with the DoF conventions of SURVEY.md App. A
(edge DoF = int_e u.t, face DoF = int_F u.n, both in +coordinate direction).  Used by the sub-complex (commuting
diagram) tests: d of a multiscale basis function of one space is the incidence-weighted sum of the multiscale basis
functions of the next space (reference doc/pages/mainpage.dox:94-130)."""
import numpy as np
import scipy.sparse as sp

from oracle import msfec_oracle as mo


def _key(p):
    return tuple(np.round(np.asarray(p) * 2).astype(int))


def gradient(g):
    """[nE, nV]: (grad phi).t integrated along an edge = phi(end) - phi(start)"""
    vkey = {_key(p): i for i, p in enumerate(g.v_pos)}
    rows, cols, vals = [], [], []
    for e, (p, d) in enumerate(zip(g.e_pos, g.e_dir)):
        lo = p.copy(); hi = p.copy(); lo[d] -= 0.5; hi[d] += 0.5
        rows += [e, e]; cols += [vkey[_key(hi)], vkey[_key(lo)]]; vals += [1.0, -1.0]
    return sp.csr_matrix((vals, (rows, cols)), shape=(g.nE, g.nV))


def curl(g):
    """[nF, nE]: flux of curl u through a face = circulation of u along its boundary (Stokes)"""
    ekey = {_key(p): i for i, p in enumerate(g.e_pos)}
    rows, cols, vals = [], [], []
    for f, (p, d) in enumerate(zip(g.f_pos, g.f_dir)):
        a, b = (d + 1) % 3, (d + 2) % 3               # (curl u)_d = d_a u_b - d_b u_a
        for sgn, edge_axis, shift_axis, shift in ((1, b, a, 0.5), (-1, b, a, -0.5), (-1, a, b, 0.5), (1, a, b, -0.5)):
            q = p.copy(); q[shift_axis] += shift
            e = ekey[_key(q)]
            assert g.e_dir[e] == edge_axis
            rows.append(f); cols.append(e); vals.append(float(sgn))
    return sp.csr_matrix((vals, (rows, cols)), shape=(g.nF, g.nE))


def coarse_incidence(op):
    """The same operator on a single cell, rows / columns in deal.II LOCAL order (the order of the coarse basis functions)."""
    g1 = mo.fine_grid(1)
    full = op(g1).toarray()
    if op is gradient:
        return full[np.ix_(g1.cE[0], g1.cV[0])]        # [12 lines, 8 vertices]
    return full[np.ix_(g1.cF[0], g1.cE[0])]            # [6 faces, 12 lines]
