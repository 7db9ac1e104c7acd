#!/bin/bash
# usage: profile_sweep.sh [size]   (ncu full capture of the two sweep kernels on the C3 recipe at that size)
export PROBE_SIZE=${1:-4000}
python tools/probe_sweep.py > gpurun_out/plain_sweep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_merge_seed|k_inflate' -s 12 -c 2 -f -o gpurun_out/prof_sweep_${PROBE_SIZE}_r1 python tools/probe_sweep.py > gpurun_out/ncu_sweep.log 2>&1
cat gpurun_out/plain_sweep.log
