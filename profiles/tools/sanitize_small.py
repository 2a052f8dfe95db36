"""Small multifrontal builds for compute-sanitizer (memcheck / racecheck): Ned_RT and RT_DQ, 2 and 3 local refinements."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location("msfec_b200", os.path.join(ROOT, "mpi-msfec_b200", "msfec_b200.py"))
m = importlib.util.module_from_spec(spec); sys.modules["msfec_b200"] = m; spec.loader.exec_module(m)
from common import lib_problem
from oracle import msfec_oracle as mo
cells = mo.morton_cells(2)[:int(sys.argv[1]) if len(sys.argv) > 1 else 8]
for pairing, L in (("NED_RT", 2), ("NED_RT", 3), ("RT_DQ", 3)):
    bb = m.BasisBuilder(lib_problem(m, pairing, L, random_seed=5, solver=m.SOLVER["mf"]), device=0).run(cells, np.arange(len(cells)))
    print(pairing, L, "residual", bb.stats["residual_max"], flush=True)
    bb.close()
