"""Generates tests/golden/reference_compiled_eqdata.npz from the REFERENCE's own code run here:
oracle/_ref/libmsfec_ref.so is the reference's eqn_coeff_A.cc / eqn_coeff_R.cc / basis_q1(.grad).tpp compiled
unmodified (oracle/Makefile), read with the reference's own example_parameters/*.prm files.  The points are the
oracle's quadrature points of two coarse cells of the 4^3 coarse grid at 2 local refinements; what is stored are the
reference's values there.  Run from the repo root in the build container: python tests/golden/make_reference_compiled_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import msfec_oracle as mo, msfec_ref as mr  # noqa: E402

REF_PRM = "/root/reference/example_parameters"
PRM = {"Q": "prm_q_test-01.prm", "Q_NED": "prm_q_ned_test-01.prm", "NED_RT": "prm_ned_rt_test-01.prm",
       "RT_DQ": "prm_rt_dq_test-01.prm"}
CELLS = (5, 37)

out = {}
cells = mo.morton_cells(2)
for pairing, name in PRM.items():
    path = os.path.join(REF_PRM, name)
    prob = mo.Problem.from_prm(path, pairing)
    prob.n_refine_local = 2
    for c in CELLS:
        x0 = cells[c].min(0); H = float(cells[c].max(0)[0] - x0[0])
        pts = mo.coefficient_fields(prob, x0, H)[3]
        out[f"{pairing}_pts_{c}"] = pts
        out[f"{pairing}_A_{c}"] = mr.diffusion_a(path, pts).reshape(pts.shape[:-1] + (3, 3))
        out[f"{pairing}_Ainv_{c}"] = mr.diffusion_a(path, pts, inverse=True).reshape(pts.shape[:-1] + (3, 3))
        out[f"{pairing}_R_{c}"] = mr.reaction_rate(pts).reshape(pts.shape[:-1])
# coarse Q1 shape functions of a non-unit, non-origin coarse cube (deal.II vertex order: x fastest)
x0 = np.array([0.25, 0.5, 0.75]); H = 0.25
vertices = x0 + H * np.array([[v & 1, (v >> 1) & 1, v >> 2] for v in range(8)], dtype=float)
rng = np.random.default_rng(20261018)
pts = x0 + H * rng.random((64, 3))
val, grad = mr.basis_q1(vertices, pts)
out.update(q1_x0=x0, q1_H=np.array(H), q1_pts=pts, q1_val=val, q1_grad=grad)
# the same cube: the reference's Q1 mapping and its covariant (Nedelec) / Piola (Raviart-Thomas) transforms; the unit-cell shape
# functions under them are stand-ins (oracle/ref_shim/deal.II/fe), so what these arrays pin is the H-scaling and the mapping
ref, jac = mr.mapping(vertices, pts)
out.update(map_ref=ref, map_inv_jac=jac, ned_val=mr.basis_vector(vertices, "ned", pts), rt_val=mr.basis_vector(vertices, "rt", pts))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_compiled_eqdata.npz"), **out)
print("wrote", len(out), "arrays")
