"""Measurement helper (not part of the product): C2 findBestPath latency and C4 sweep time with CUDA events."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import navigation_b200  # noqa: E402

api = navigation_b200.load()
grid = bench.inflate_local(api, bench.local_map_c2())
reps = int(os.environ.get("PROBE_REPS", 10))
d2, pose, vel = bench.dwa_setup(api, grid, bench.C2)
for _ in range(3):
    d2.find_best_path(pose, vel, bench.PENTAGON, want_costs=False)
t0 = time.perf_counter()
for _ in range(reps):
    r = d2.find_best_path(pose, vel, bench.PENTAGON, want_costs=False)
c2_us = 1e6 * (time.perf_counter() - t0) / reps
d4, pose, vel = bench.dwa_setup(api, grid, bench.C4)
stream = torch.cuda.ExternalStream(d4.stream())
for _ in range(2):
    d4.find_best_path_async(pose, vel, bench.PENTAGON)
d4.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    d4.find_best_path_async(pose, vel, bench.PENTAGON)
e1.record(stream)
d4.synchronize()
torch.cuda.synchronize()
print(f"c2_findBestPath_us={c2_us:.1f} best={r['best_index']} c4_cycle_ms={e0.elapsed_time(e1) / reps:.3f}")
# C2 again without the Python result marshalling: async enqueue + synchronize, and the device span by CUDA events
s2 = torch.cuda.ExternalStream(d2.stream())
for _ in range(3):
    d2.find_best_path_async(pose, vel, bench.PENTAGON); d2.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    d2.find_best_path_async(pose, vel, bench.PENTAGON); d2.synchronize()
wall = 1e6 * (time.perf_counter() - t0) / reps
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in ev:
    a.record(s2); d2.find_best_path_async(pose, vel, bench.PENTAGON); b.record(s2); d2.synchronize()
dev = 1e3 * np.mean([a.elapsed_time(b) for a, b in ev])
t0 = time.perf_counter()
for _ in range(reps):
    d2.set_plan(pose, np.stack([np.arange(1.0, 7.0, 0.05), np.full(120, 3.0)], 1))
plan_us = 1e6 * (time.perf_counter() - t0) / reps
print(f"c2_async+sync_wall_us={wall:.1f} c2_device_span_us={dev:.1f} set_plan_us={plan_us:.1f}")
