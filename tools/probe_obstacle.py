"""Measurement helper: times k_obstacle_update alone (CUDA events around back-to-back cycles with the sweep disabled
is not possible, so this times whole cycles with NAVGPU_DEBUG_SKIP variants; compare differences)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, navigation_b200
api = navigation_b200.load()
size = 4000
cm, (s, o, il), sets = bench.build_c3(lambda *a: api.costmap(*a), size=size)
obs, robot = sets[0]
cm.set_observations(o, obs)
stream = torch.cuda.ExternalStream(cm.stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    cm.touch_grid_layer(s, 0, 0, size, size); cm.update_map(*robot)
n = 30
e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
for k in range(n):
    with torch.cuda.stream(stream):
        flush.zero_(); e0[k].record(stream)
    cm.touch_grid_layer(s, 0, 0, size, size); cm.update_map_async(*robot); e1[k].record(stream)
torch.cuda.synchronize()
print(f"skip={os.environ.get('NAVGPU_DEBUG_SKIP','0')} cycle_flushed_ms={np.mean([a.elapsed_time(b) for a, b in zip(e0, e1)]):.4f}")
