"""Measurement helper: C5 fleet cycle (4096 robots) wall time through the C ABI; run under ncu for the kernel split."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, navigation_b200
api = navigation_b200.load()
n = int(os.environ.get("PROBE_ROBOTS", 4096))
raw, origins, poses, vels, plans = bench.fleet_inputs(range(n))
fleet = api.fleet(n, 120, 120, 0.05, bench.PENTAGON, 0.55, 10.0, vx_samples=20, vy_samples=1, vth_samples=20, max_vel_y=0.0, min_vel_y=0.0)
fleet.set_maps(raw, origins)
fleet.set_plans(poses, plans)
poses = np.ascontiguousarray(poses); vels = np.ascontiguousarray(vels)
reps = int(os.environ.get("PROBE_REPS", 5))
for _ in range(2):
    fleet.step_raw(poses, vels)
t0 = time.perf_counter()
for _ in range(reps):
    fleet.step_raw(poses, vels)
step = (time.perf_counter() - t0) / reps
t0 = time.perf_counter()
for _ in range(reps):
    fleet.set_maps(raw, origins)
maps = (time.perf_counter() - t0) / reps
print(f"robots={n} step_ms={1e3 * step:.3f} set_maps_ms={1e3 * maps:.3f}")
