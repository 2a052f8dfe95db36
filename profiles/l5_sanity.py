"""Sanity run at 5 local refinements (n = 32) on a GPU box: python profiles/l5_sanity.py"""
import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from conftest import msfec_mod
from common import lib_problem
from oracle import msfec_oracle as mo
m = msfec_mod()
cells = mo.morton_cells(2)[:3]
bb = m.BasisBuilder(lib_problem(m, "NED_RT", 5, use_direct_solver_basis=1), device=0).run(cells, np.arange(3))
st = bb.stats; M = bb.get_global_element_matrix().copy(); k0 = 12; scale = np.abs(M).max()
print("residual", st["residual_max"], "nc", st["not_converged"], "scale", scale)
print("sym00", np.abs(M[:, :k0, :k0] - M[:, :k0, :k0].transpose(0, 2, 1)).max() / scale)
print("sym11", np.abs(M[:, k0:, k0:] - M[:, k0:, k0:].transpose(0, 2, 1)).max() / scale)
print("anti01", np.abs(M[:, :k0, k0:] + M[:, k0:, :k0].transpose(0, 2, 1)).max() / scale)
bb2 = m.BasisBuilder(lib_problem(m, "NED_RT", 5, use_direct_solver_basis=1), device=0).run(cells[2:3], np.arange(2, 3))
print("batch", np.abs(bb2.get_global_element_matrix()[0] - M[2]).max() / scale)
