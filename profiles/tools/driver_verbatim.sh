#!/bin/bash
# The four MsFEC_* drivers on the reference's shipped .prm files VERBATIM (only the output directory is redirected): fine-grid
# comparator XStd on 64^3 cells (6 refinements) first, then the multiscale method (4^3 coarse cells x 4 local refinements).
# Logs: gpurun_out/verbatim/<pairing>.log
mkdir -p gpurun_out/verbatim
for P in q:MsFEC_Q q_ned:MsFEC_Q_Ned ned_rt:MsFEC_Ned_RT rt_dq:MsFEC_RT_DQ; do
  f=${P%%:*}; exe=${P##*:}
  sed -e "s#set dirname output = .*#set dirname output = gpurun_out/verbatim/out_$f#" examples/prm/prm_${f}_test-01.prm > gpurun_out/verbatim/$f.prm
  ( time MSFEC_MAX_OUTPUT_CELLS=2 timeout 300 mpi-msfec_b200/host/$exe -p gpurun_out/verbatim/$f.prm ) > gpurun_out/verbatim/$f.log 2>&1
  echo "== $f: exit $?"; grep -E "Coarse solver|solution norms|real" gpurun_out/verbatim/$f.log | cut -c1-260
  rm -rf gpurun_out/verbatim/out_$f
done
