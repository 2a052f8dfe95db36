#!/bin/bash
# Forward kernel with one phase switched off / redirected at a time (library built with EXTRA=-DMSFEC_MF_PHASE_SWITCHES;
# MSFEC_MF_DBG bits, mf.cuh): per-level durations of the first build.   profiles/tools/phase_cost.sh [bits ...]
mkdir -p gpurun_out/ab
bits=${@:-0 1 2 4 8 16 32 64 128 256}
for d in $bits; do
  MSFEC_MF_DBG=$d MSFEC_MF_STAGED=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mf_forward -c 9 --csv --log-file gpurun_out/ab/dbg$d.csv python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
done
python - $bits <<'PY'
import csv, sys
names = {0: "baseline", 1: "no C gather", 2: "no C store", 4: "no record store", 8: "no panel children", 16: "no pe/rhs loads", 32: "no contribution phase",
         64: "no factorisation", 3: "no C gather+store", 128: "C gather from L2 (cell 0)", 256: "C gather from L1 (2 KB)"}
for d in map(int, sys.argv[1:]):
    rows = [r for r in csv.reader(open(f"gpurun_out/ab/dbg{d}.csv")) if len(r) > 10]
    hdr = rows[0]; iv = hdr.index("Metric Value")
    t = [float(r[iv].replace(",", "")) / 1e6 for r in rows[1:]]
    print(f"{names.get(d, d):26s} " + " ".join(f"{x:6.3f}" for x in t) + f"  total {sum(t):.2f}")
PY
