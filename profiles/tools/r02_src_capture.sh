mkdir -p gpurun_out/r2b
ncu --set full --clock-control none --import-source on -k regex:k_mf_forward -c 4 -o gpurun_out/r2b/fwd03 python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2b/fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mf_backward -s 8 -c 1 -o gpurun_out/r2b/bwd8 python bench.py --cells 2048 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2b/bwd.log 2>&1
for i in 0 1 2 3; do ncu -i gpurun_out/r2b/fwd03.ncu-rep --launch-skip $i --launch-count 1 --page source --csv --print-source cuda > gpurun_out/r2b/fwd_l$i.src.csv 2>/dev/null; done
ncu -i gpurun_out/r2b/bwd8.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2b/bwd8.src.csv
ls -la gpurun_out/r2b
