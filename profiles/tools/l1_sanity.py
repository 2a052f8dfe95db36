import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import importlib.util
spec = importlib.util.spec_from_file_location("msfec_b200", "mpi-msfec_b200/msfec_b200.py")
m = importlib.util.module_from_spec(spec); sys.modules["msfec_b200"] = m; spec.loader.exec_module(m)
from common import lib_problem, oracle_problem, rel_err
from oracle import msfec_oracle as mo
cells = mo.morton_cells(2)[:40]; ids = np.arange(40)
for pairing in mo.PAIRINGS:
    for L in (1, 2):
        for solver in ("auto", "mf", "minres"):
            try:
                bb = m.BasisBuilder(lib_problem(m, pairing, L, random_seed=7, solver=m.SOLVER[solver]), device=0).run(cells, ids)
            except Exception as e:
                print(pairing, L, solver, "ERR", str(e)[:100]); continue
            Mo, ro = mo.build_basis(oracle_problem(pairing, L, random_seed=7), cells[33], 33)[:2]
            print(pairing, L, solver, "solver used", bb.stats["solver"], "err", rel_err(bb.get_global_element_matrix()[33], Mo), rel_err(bb.get_global_element_rhs()[33], ro), "res", bb.stats["residual_max"])
            bb.close()
