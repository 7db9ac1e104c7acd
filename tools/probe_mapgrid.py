"""Measurement helper (not part of the product): C2 findBestPath device span and the fleet's cycle with each MapGrid
kernel variant (NAVGPU_MAPGRID = sliced | rows, NAVGPU_MAPGRID_HELPERS = 0 | 1, read by libnavgpu at every launch)."""
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
reps = int(os.environ.get("PROBE_REPS", 20))
d2, pose, vel = bench.dwa_setup(api, grid, bench.C2)
s2 = torch.cuda.ExternalStream(d2.stream())
for variant in ("sliced", "rows", "rows+h") * 2:
    os.environ["NAVGPU_MAPGRID"] = variant.split("+")[0]
    os.environ["NAVGPU_MAPGRID_HELPERS"] = str(int(variant.endswith("+h")))
    for _ in range(3):
        d2.find_best_path_async(pose, vel, bench.PENTAGON); d2.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(s2); d2.find_best_path_async(pose, vel, bench.PENTAGON); b.record(s2); d2.synchronize()
    dev = 1e3 * np.array([a.elapsed_time(b) for a, b in ev])
    print(f"{variant}: c2 device span us mean {dev.mean():.1f} min {dev.min():.1f}", flush=True)

if os.environ.get("PROBE_FLEET", "1") != "0":
    n = int(os.environ.get("PROBE_FLEET_ROBOTS", 4096))
    raw, origins, poses, vels, plans = bench.fleet_inputs(range(n))
    fleet = api.fleet(n, 120, 120, 0.05, bench.PENTAGON, 0.55, 10.0, vx_samples=20, vy_samples=1, vth_samples=20,
                      max_vel_y=0.0, min_vel_y=0.0)
    fleet.set_maps(raw, origins)
    fleet.set_plans(poses, plans)
    poses = np.ascontiguousarray(poses)
    vels = np.ascontiguousarray(vels)
    for variant in ("sliced", "rows", "rows+h") * 2:
        os.environ["NAVGPU_MAPGRID"] = variant.split("+")[0]
        os.environ["NAVGPU_MAPGRID_HELPERS"] = str(int(variant.endswith("+h")))
        fleet.step_raw(poses, vels)
        t = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fleet.step_raw(poses, vels)
            torch.cuda.synchronize()
            t.append(1e3 * (time.perf_counter() - t0))
        print(f"{variant}: fleet step wall ms min {min(t):.3f} mean {np.mean(t):.3f}", flush=True)
