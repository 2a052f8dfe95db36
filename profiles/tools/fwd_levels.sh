#!/bin/bash
# per-level durations of k_mf_forward (first build only) for a list of environment settings:
#   profiles/tools/fwd_levels.sh "tag1:ENV=VAL ENV=VAL" "tag2:..." ...
mkdir -p gpurun_out/ab
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  env $envs ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mf_forward -c 9 --csv --log-file gpurun_out/ab/lv_$tag.csv python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ab/lv_$tag.err
  python - "$tag" <<'PY'
import csv, sys
tag = sys.argv[1]
try:
    rows = [r for r in csv.reader(open(f"gpurun_out/ab/lv_{tag}.csv")) if len(r) > 10]
    hdr = rows[0]; iv = hdr.index("Metric Value")
    t = [float(r[iv].replace(",", "")) / 1e6 for r in rows[1:]]
    print(f"{tag:16s} " + " ".join(f"{x:6.3f}" for x in t) + f"  total {sum(t):.2f}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/ab/lv_{tag}.err").read()[-400:])
PY
done
