python -m pytest tests -m gpu -q 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["config"]["hot_l2_ms_per_step"], d["cpu_baseline"]["value"])
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d["dwa"].items() if k in ("c2_findBestPath_us","c4_sweep_ms","c4_traj_per_s","c5_cycle_ms","c5_e2e_cycle_ms","c5_robot_cycles_per_s")})
print(d["voxel_layer"]["c3_with_voxel_layer_update_map_ms"], d["trajectory_planner"]["findBestPath_us"], d["dwa"]["cpu_baseline"]["c2_findBestPath_ms"], d["dwa"]["cpu_baseline"]["c5_all_cores"]["c5_robot_cycles_per_s"])
PY
