#!/bin/bash
# The MsFEC_Ned_RT host driver end to end at C5's shape (32^3 coarse cells x 3 local refinements; coefficients of the reference's
# prm_ned_rt_test-01.prm): basis build on the GPU, ncclAllGather / coarse assembly, coarse solve on the GPU, weight scatter,
# norms, coarse-level output (per-cell .vtu files limited to 2).  Log: gpurun_out/driver_c5.log
mkdir -p gpurun_out/driver_c5
sed -e 's/set global refinements = 2/set global refinements = 5/' -e 's/set local refinements = 4/set local refinements = 3/' \
    -e 's#set dirname output = .*#set dirname output = gpurun_out/driver_c5/out#' examples/prm/prm_ned_rt_test-01.prm > gpurun_out/driver_c5/c5.prm
( time MSFEC_MAX_OUTPUT_CELLS=2 MSFEC_NCCL=1 mpi-msfec_b200/host/MsFEC_Ned_RT -p gpurun_out/driver_c5/c5.prm ) > gpurun_out/driver_c5.log 2>&1
rm -rf gpurun_out/driver_c5/out/*.bin
