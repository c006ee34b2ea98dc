#!/bin/bash
# Profiles of the pair-once sweep (on the GPU box, one gpurun call): launch list of the bench
# command, then ncu --set full of sym_sweep_kernel at the metric size (source page included).
# Each ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --parity 0"
$B > gpurun_out/r2_sym_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_sym_launches_bench.csv $B > gpurun_out/r2_sym_ncu_launches.log 2>&1
N=${1:-131072}
python tools/time_sym.py $N > gpurun_out/r2_sym_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sym_sweep_kernel -s 2 -c 1 -f \
    -o gpurun_out/r2_prof_sym python tools/time_sym.py $N > gpurun_out/r2_sym_ncu.log 2>&1
ncu --set full --clock-control none -k regex:sym_reduce_kernel -s 2 -c 1 -f \
    -o gpurun_out/r2_prof_sym_reduce python tools/time_sym.py $N > gpurun_out/r2_sym_reduce_ncu.log 2>&1
cat gpurun_out/r2_sym_plain.log
ls -la gpurun_out/*.ncu-rep
