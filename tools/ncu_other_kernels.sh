#!/bin/bash
# ncu --set full of the sweep kernels other than the metric kernel (one launch each), for profiles/.
# usage (on the GPU box): bash tools/ncu_other_kernels.sh
set -u
mkdir -p gpurun_out
for spec in "acc 131072" "pot 131072" "kick 131072" "kickt 131072" "tangent 32768"; do
  set -- $spec
  op=$1; n=$2
  python tools/run_sweep.py $op $n 2 > gpurun_out/plain_$op.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 \
      -f -o gpurun_out/prof_${op}_r1 python tools/run_sweep.py $op $n 2 > gpurun_out/ncu_$op.log 2>&1
  tail -2 gpurun_out/plain_$op.log
done
