#!/bin/bash
# Round-2 profiles (on the GPU box, one gpurun call): launch list of the bench command, then
# ncu --set full of the metric kernel, the kick with the bit-exact time-scale, and the tangent sweep.
# Each ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --parity 0"
$B > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
for spec in "accjerk 131072 k2" "kickt 131072 kickt" "tangent 32768 tangent"; do
  set -- $spec
  op=$1; n=$2; name=$3
  python tools/run_sweep.py $op $n 2 > gpurun_out/r2_plain_$op.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 \
      -f -o gpurun_out/r2_prof_$name python tools/run_sweep.py $op $n 2 > gpurun_out/r2_ncu_$op.log 2>&1
  tail -1 gpurun_out/r2_plain_$op.log
done
ls -la gpurun_out/*.ncu-rep
