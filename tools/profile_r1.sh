#!/bin/bash
# ncu captures for profiles/ (run under gpurun, one GPU). Every ncu command is preceded by the same command run plain.
set -x
export PROBE_REPS=3
python tools/probe_sweep.py > gpurun_out/plain_sweep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_update_costs_fast -s 6 -c 2 -f -o gpurun_out/prof_sweep_r1 python tools/probe_sweep.py > gpurun_out/ncu_sweep.log 2>&1
python tools/probe_dwa.py > gpurun_out/plain_dwa.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_dwa_score|k_mapgrid_prepare' -s 8 -c 4 -f -o gpurun_out/prof_dwa_r1 python tools/probe_dwa.py > gpurun_out/ncu_dwa.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/plain_*.log
