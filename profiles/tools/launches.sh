#!/bin/bash
# launch list (durations) of one 4096-cell build: profiles/tools/launches.sh <tag> [ENV=VAL ...]
tag=$1; shift
mkdir -p gpurun_out/ab
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ab/launches_$tag.csv python bench.py --cells 4096 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ab/launches_$tag.err
python profiles/summarize_launches.py gpurun_out/ab/launches_$tag.csv 2>/dev/null | head -40
