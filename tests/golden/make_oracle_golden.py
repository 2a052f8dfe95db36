"""Writes tests/golden/oracle_elem_matrices.npz: coarse element matrices / rhs produced by
oracle/msfec_oracle.py (exact direct solve) for the four pairings with the coefficients of the
reference's prm_*_test-01.prm files (examples/prm/), local refinements 2, coarse cells 5 and 37 of
the 2x-refined unit cube.  Regression anchor for both the oracle and the CUDA path.

    python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import msfec_oracle as mo  # noqa: E402

FILES = {"Q": "prm_q_test-01.prm", "Q_NED": "prm_q_ned_test-01.prm", "NED_RT": "prm_ned_rt_test-01.prm",
         "RT_DQ": "prm_rt_dq_test-01.prm"}
out = {}
cells = mo.morton_cells(2)
for p, f in FILES.items():
    prob = mo.Problem.from_prm(os.path.join(ROOT, "examples", "prm", f), p)
    prob.n_refine_local = 2
    for c in (5, 37):
        M, r, *_ = mo.build_basis(prob, cells[c], c)
        out[f"{p}_M_{c}"] = M
        out[f"{p}_r_{c}"] = r
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_elem_matrices.npz"), **out)
print({k: v.shape for k, v in out.items()})
