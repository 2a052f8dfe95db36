#!/bin/bash
# compute-sanitizer (memcheck, racecheck) on the device coarse solve: Ned_RT (Schur-complement CG + inner CG) and Q (CG), 64 coarse cells
mkdir -p gpurun_out/san
for P in NED_RT Q; do
  python profiles/tools/coarse_timing.py $P 2 device 0 > gpurun_out/san/coarse_$P.log 2>&1
  for T in memcheck racecheck; do
    compute-sanitizer --tool $T mpi-msfec_b200/host/coarse_test .scratch/coarse_in.bin .scratch/san_out.bin 0 > gpurun_out/san/coarse_${P}_$T.log 2>&1
    echo "$P $T: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san/coarse_${P}_$T.log)"
  done
done
