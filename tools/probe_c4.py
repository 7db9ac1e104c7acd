"""Measurement helper: a few C4 sweeps (844 200 samples) for ncu captures of k_dwa_score."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, navigation_b200
api = navigation_b200.load()
grid = bench.inflate_local(api, bench.local_map_c2())
d4, pose, vel = bench.dwa_setup(api, grid, bench.C4)
for _ in range(int(os.environ.get("PROBE_REPS", 3))):
    r = d4.find_best_path(pose, vel, bench.PENTAGON, want_costs=False)
print("c4 best", r["best_index"], r["cost"], r["n_samples"])
