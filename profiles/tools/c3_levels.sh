#!/bin/bash
# per-launch durations of the front kernels for C3 (prm_ned_rt_test-01.prm: 64 cells, 4 local refinements)
mkdir -p gpurun_out/ab
cat > /tmp/c3run.py <<'PY'
import sys, os, numpy as np
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location("msfec_b200", os.path.join(ROOT, "mpi-msfec_b200", "msfec_b200.py"))
m = importlib.util.module_from_spec(spec); sys.modules["msfec_b200"] = m; spec.loader.exec_module(m)
from common import lib_problem
from oracle import msfec_oracle as mo
pairing = sys.argv[1] if len(sys.argv) > 1 else "NED_RT"
cells = mo.morton_cells(2)
bb = m.BasisBuilder(lib_problem(m, pairing, 4), device=0)
bb.run(cells); bb.run(cells)
print(bb.stats["ms_total"], bb.stats["solver"])
PY
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mf_ --csv --log-file gpurun_out/ab/c3_levels.csv python /tmp/c3run.py NED_RT > gpurun_out/ab/c3_levels.out 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/ab/c3_levels.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); ig = hdr.index("Grid Size"); ib = hdr.index("Block Size")
half = (len(rows) - 1) // 2
fw = bw = 0.0
for i, r in enumerate(rows[1 + half:]):
    name = r[ik].split("(")[0].split("::")[-1]
    t = float(r[iv].replace(",", "")) / 1e6
    if "forward" in name: fw += t
    if "backward" in name: bw += t
    print(f"{i:3d} {name:28s} grid {r[ig]:18s} {t:.3f} ms")
print("forward", fw, "backward", bw)
PY
