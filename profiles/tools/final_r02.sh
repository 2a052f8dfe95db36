#!/bin/bash
# Final measurements of the round (one gpurun call): bench lines, launch list, per-level ncu tables, C1-C4 table.
set -x
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/r02_bench_c5_mf.json 2> $O/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_mf.csv python bench.py --cells 4096 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize_launches.py $O/r02_launches_mf.csv > $O/r02_launches_mf.txt
ncu --set full --clock-control none --import-source on -k regex:k_mf_forward -c 9 -o $O/fwd python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python profiles/ncu_levels.py $O/fwd.ncu-rep 2048 $O/r02_mf_forward_ncu.json > $O/r02_fwd_levels.md
ncu --set full --clock-control none --import-source on -k regex:k_mf_backward -c 9 -o $O/bwd python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python profiles/ncu_levels.py $O/bwd.ncu-rep 2048 $O/r02_mf_backward_ncu.json > $O/r02_bwd_levels.md
for i in 0 1 3; do python profiles/tools/ncu_lines.py $O/fwd.ncu-rep $i 14 > $O/r02_fwd_lines_l$i.md; done
python profiles/tools/ncu_lines.py $O/bwd.ncu-rep 8 14 > $O/r02_bwd_lines_leaf.md
rm -f $O/fwd.ncu-rep $O/bwd.ncu-rep
SOLVERS=auto,band,mf,minres python profiles/config_table.py > $O/r02_config_table_c1_c4.md 2> $O/config.err
ls -la $O
