#!/bin/bash
# ncu --set full capture of the pair-once sweep kernel at N = 32768 (one launch), source page included.
set -e
N=${1:-32768}
python tools/time_sym.py $N > gpurun_out/ncu_sym_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_sweep_kernel -c 1 -f \
    -o gpurun_out/ncu_sym python tools/time_sym.py $N > gpurun_out/ncu_sym.log 2>&1
