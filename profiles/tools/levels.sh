#!/bin/bash
# per-launch durations of the front kernels: profiles/tools/levels.sh <tag> [ENV=VAL ...]  -> gpurun_out/ab/<tag>.levels.csv
tag=$1; shift
mkdir -p gpurun_out/ab
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mf_ --csv --log-file gpurun_out/ab/$tag.levels.csv python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ab/$tag.levels.err
python - "$tag" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/ab/{tag}.levels.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); ig = hdr.index("Grid Size")
out = []
for r in rows[1:]:
    name = r[ik].split("(")[0].split("::")[-1]
    out.append(f"{name} grid {r[ig]} {float(r[iv].replace(',', ''))/1e6:.3f} ms")
print(tag); print("\n".join(out))
PY
