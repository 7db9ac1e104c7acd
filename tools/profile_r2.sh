#!/bin/bash
# All ncu captures behind profiles/r2_* (run under gpurun on one GPU).  Each ncu command follows a plain run of the same
# command line that exited 0 (tools/profile_kernel.sh / tools/profile_launches.sh do that).
export PROBE_REPS=3
bash tools/profile_kernel.sh 'k_merge_seed|k_inflate' prof_sweep_r2 tools/probe_sweep.py 12 2
bash tools/profile_kernel.sh k_obstacle_update prof_obstacle_r2 tools/probe_cycle.py 6 1
bash tools/profile_kernel.sh k_dwa_score prof_c4_r2 tools/probe_c4.py 1 1
bash tools/profile_kernel.sh k_mirror_diff prof_mirror_r2 tools/probe_mirror.py 3 1
PROBE_FLEET=0 PROBE_REPS=1 bash tools/profile_kernel.sh 'k_mapgrid_prepare' prof_mapgrid_r2 tools/probe_mapgrid.py 3 6
bash tools/profile_launches.sh tools/probe_fleet.py fleet
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 600 gpurun_out/plain_bench.log
