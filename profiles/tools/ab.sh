#!/bin/bash
# A/B runs of bench.py on a 4096-cell C5 slice: profiles/tools/ab.sh <tag> [ENV=VAL ...]   (results: gpurun_out/ab/<tag>.json)
tag=$1; shift
mkdir -p gpurun_out/ab
env "$@" python bench.py --cells 4096 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab/$tag.json 2> gpurun_out/ab/$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab/{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{tag}: {d['value']:.0f} cells/s, step {d['ms_per_step']:.2f} ms, fwd {r.get('ms_per_step', 0):.2f} bwd {r.get('backward', {}).get('ms_per_step', 0):.2f} "
          f"phases {d.get('phases_ms')} res {d['krylov']['residual_max']:.1e}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/ab/{tag}.err").read()[-600:])
PY
