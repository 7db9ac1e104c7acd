#!/bin/bash
python tools/probe_sweep.py > gpurun_out/plain_sweep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_merge_seed|k_inflate' -s 12 -c 2 -f -o gpurun_out/prof_sweep2_r1 python tools/probe_sweep.py > gpurun_out/ncu_sweep.log 2>&1
cat gpurun_out/plain_sweep.log
