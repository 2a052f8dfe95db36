"""Timing of BASELINE.json configs[0..3] (the reference's prm_*_test-01.prm: 64 coarse cells, 4 local
refinements) through the solvers of the local problems (auto = what the verbatim .prm gets), next to the CPU oracle.
Run on a GPU box (SOLVERS=auto,band,mf,minres selects; minres on C2 / C3 takes minutes):
    python profiles/config_table.py > gpurun_out/config_table.md
Parity of these configurations is asserted in tests/test_gpu_parity.py::test_reference_prm_configs_full_size."""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import PRM, lib_problem, oracle_problem  # noqa: E402
from oracle import msfec_oracle as mo  # noqa: E402

spec = importlib.util.spec_from_file_location("msfec_b200", os.path.join(ROOT, "mpi-msfec_b200", "msfec_b200.py"))
m = importlib.util.module_from_spec(spec); sys.modules["msfec_b200"] = m; spec.loader.exec_module(m)

cells = mo.morton_cells(2)
print("| config | pairing | N fine DoFs | k | solver | ms / 64 cells | cells/s | fine DoF-solves/s | Krylov its (max) | residual | CPU oracle s/cell (1 core) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for i, p in enumerate(("Q", "Q_NED", "NED_RT", "RT_DQ")):
    t0 = time.perf_counter(); mo.build_basis(oracle_problem(p, 4), cells[5], 5); cpu = time.perf_counter() - t0
    for solver in os.environ.get("SOLVERS", "auto,band,mf").split(","):
        try:
            bb = m.BasisBuilder(lib_problem(m, p, 4, solver=m.SOLVER[solver]), device=0)
            bb.run(cells)                                  # warm-up (allocations)
            bb.run(cells)
        except m.MsfecError as e:
            print(f"| C{i + 1} {PRM[p]} | {p} | | | {solver} | unavailable: {str(e)[:60]} | | | | | |")
            continue
        st = bb.stats
        cps = 64 / (st["ms_total"] * 1e-3)
        used = {0: "MINRES", 1: "banded LDL^T", 2: "multifrontal LDL^T"}[st["solver"]]
        print(f"| C{i + 1} {PRM[p]} | {p} | {st['n_fine_dofs']} | {st['k']} | {solver} -> {used} | {st['ms_total']:.1f} | "
              f"{cps:.1f} | {cps * st['k'] * st['n_fine_dofs']:.3g} | {st['iterations_max']} | {st['residual_max']:.1e} | {cpu:.2f} |")
        bb.close()
