#!/bin/bash
# Turns the round-2 .ncu-rep / csv files in gpurun_out/ (scratch) into the tracked text summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
R=r2
python tools/ncu_summary.py gpurun_out/prof_sweep_$R.ncu-rep > profiles/${R}_sweep_ncu_summary.txt
python tools/ncu_lines.py gpurun_out/prof_sweep_$R.ncu-rep k_merge_seed > profiles/${R}_k_merge_seed_lines.txt
python tools/ncu_lines.py gpurun_out/prof_sweep_$R.ncu-rep k_inflate 0 k_inflateILi20 > profiles/${R}_k_inflate_lines.txt
python tools/sass_opcodes.py gpurun_out/prof_sweep_$R.ncu-rep k_inflate > profiles/${R}_k_inflate_sass.txt
python tools/ncu_summary.py gpurun_out/prof_obstacle_$R.ncu-rep > profiles/${R}_k_obstacle_update_ncu_summary.txt
python tools/ncu_summary.py gpurun_out/prof_c4_$R.ncu-rep > profiles/${R}_k_dwa_score_c4_ncu_summary.txt
python tools/ncu_lines.py gpurun_out/prof_c4_$R.ncu-rep k_dwa_score 0 k_dwa_scoreE > profiles/${R}_k_dwa_score_c4_lines.txt
python tools/ncu_summary.py gpurun_out/prof_mirror_$R.ncu-rep > profiles/${R}_k_mirror_diff_ncu_summary.txt
python tools/ncu_summary.py gpurun_out/prof_mapgrid_$R.ncu-rep > profiles/${R}_k_mapgrid_prepare_ncu_summary.txt
python tools/launch_table.py gpurun_out/launches_fleet.csv > profiles/${R}_launches_fleet.txt
python tools/launch_table.py gpurun_out/launches_bench.csv > profiles/${R}_launches_bench.txt
python tools/roofline_inputs.py > /dev/null
ls -la profiles
