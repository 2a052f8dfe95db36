"""Shared problem definitions for the tests (coefficients of the reference's test-01 .prm files)."""
import os

import numpy as np

from oracle import msfec_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRM = {"Q": "prm_q_test-01.prm", "Q_NED": "prm_q_ned_test-01.prm", "NED_RT": "prm_ned_rt_test-01.prm",
       "RT_DQ": "prm_rt_dq_test-01.prm"}
VEC_RHS = "scale*(2*x-1)*(y^2-y)*(z^2-z); scale*(2*y-1)*(x^2-x)*(z^2-z); scale*(2*z-1)*(x^2-x)*(y^2-y)"
SC_RHS = "sin(2*pi*x) * sin(2*pi*y) * sin(2*pi*z)"
B_EXPR = "scale * (1.0 - alpha * sin(2*pi*frequency*x))"


def prm_path(pairing):
    return os.path.join(ROOT, "examples", "prm", PRM[pairing])


def oracle_problem(pairing, L, random_seed=0, **kw):
    prob = mo.Problem.from_prm(prm_path(pairing), pairing)
    prob.n_refine_local = L
    prob.random_field_seed = random_seed
    for k, v in kw.items():
        setattr(prob, k, v)
    return prob


def lib_problem(msfec, pairing, L, random_seed=0, **kw):
    p = msfec.problem_from_prm(prm_path(pairing), pairing)
    p.n_refine_local = L
    p.random_field_seed = random_seed
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


KINDS = {"Q": ("V", None), "Q_NED": ("V", "E"), "NED_RT": ("E", "F"), "RT_DQ": ("F", "C")}


def perm_to_oracle(bb, prob, block, kind):
    """library fine-DoF index -> oracle fine-DoF index of one block (matched by position)."""
    g = mo.fine_grid(prob.n)
    opos = {"V": g.v_pos, "E": g.e_pos, "F": g.f_pos, "C": None}[kind]
    pos, axis, bnd = bb.layout(block)
    if kind == "C":
        n = prob.n
        return (np.floor(pos[:, 0]) + n * (np.floor(pos[:, 1]) + n * np.floor(pos[:, 2]))).astype(int)
    key = {tuple(np.round(p * 2).astype(int)): i for i, p in enumerate(opos)}
    return np.array([key[tuple(np.round(p * 2).astype(int))] for p in pos])
