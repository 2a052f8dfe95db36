import os, sys
import numpy as np
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location("msfec_b200", os.path.join(ROOT, "mpi-msfec_b200", "msfec_b200.py"))
m = importlib.util.module_from_spec(spec); sys.modules["msfec_b200"] = m; spec.loader.exec_module(m)
from common import lib_problem
from oracle import msfec_oracle as mo
cells = mo.morton_cells(2)[:2]
for pairing, L, solver in (("Q", 3, "mf"), ("Q_NED", 3, "mf"), ("NED_RT", 2, "band"), ("NED_RT", 2, "minres"), ("NED_RT", 4, "mf")):
    bb = m.BasisBuilder(lib_problem(m, pairing, L, random_seed=5, solver=m.SOLVER[solver]), device=0).run(cells, np.arange(len(cells)))
    w = np.ones((len(cells), bb.k)); bb.set_global_weights(w) if hasattr(bb, "set_global_weights") else None
    print(pairing, L, solver, "residual", bb.stats["residual_max"], flush=True)
    bb.close()
