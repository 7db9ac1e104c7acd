"""Measurement helper: timeline of one C3 cycle's kernels on the device's global timer (needs NAVGPU_TRACE=1)."""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("NAVGPU_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import navigation_b200  # noqa: E402

api = navigation_b200.load()
size = int(os.environ.get("PROBE_SIZE", 4000))
cm, (s, o, il), sets = bench.build_c3(lambda *a: api.costmap(*a), size=size)
obs, robot = sets[0]
cm.set_observations(o, obs if os.environ.get("PROBE_NO_OBS") is None else [])
stream = torch.cuda.ExternalStream(cm.stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for k in range(12):
    if os.environ.get("PROBE_FLUSH", "1") == "1":
        with torch.cuda.stream(stream):
            flush.zero_()
    cm.touch_grid_layer(s, 0, 0, size, size)
    cm.update_map_async(*robot)
    t = cm.last_trace().astype(np.int64)
    t0 = min(t[0], t[2])
    rows.append([(int(v) - int(t0)) / 1e3 for v in t[:9]])
r = np.median(np.array(rows[3:]), axis=0)
print("us from the first kernel start: obstacle %.1f..%.1f  merge %.1f..%.1f  inflate %.1f..%.1f  "
      "last box merge tile %.1f  first/last inflate tile released %.1f / %.1f" % tuple(r))
