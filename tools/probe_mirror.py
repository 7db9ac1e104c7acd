"""Measurement helper: the host-mirror download of the C3 cycle (navgpu_costmap_get_changed) -- wall time of the call,
changed tiles and bytes per cycle with the observation set changing every cycle."""
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
size = int(os.environ.get("PROBE_SIZE", 4000))
cm, (s, o, il), sets = bench.build_c3(lambda *a: api.costmap(*a), size=size)
pinned = torch.empty((size, size), dtype=torch.uint8, pin_memory=True)
mirror = pinned.numpy()
packed = [cm.pack_observations(ob) for ob, _ in sets]
ts, tiles, nbytes, ts0 = [], [], [], []
for k in range(12):
    ob, rb = sets[k % len(sets)]
    cm.set_packed_observations(o, packed[k % len(sets)])
    cm.touch_grid_layer(s, 0, 0, size, size)
    cm.update_map(*rb)
    t0 = time.perf_counter()
    n, b, _ = cm.get_changed(mirror)
    ts.append(time.perf_counter() - t0)
    tiles.append(n)
    nbytes.append(b)
    t0 = time.perf_counter()
    cm.get_changed(mirror)   # nothing happened since: launch + wait + marshalling only
    ts0.append(time.perf_counter() - t0)
assert np.array_equal(mirror, cm.get())
print(f"size={size} get_changed_us={1e6 * np.median(ts[2:]):.1f} (empty call {1e6 * np.median(ts0[2:]):.1f}) tiles={tiles[2:]} bytes={nbytes[2:]}")


def tmed(fn, n=200):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))


print(f"binding overheads: last_window (no CUDA call) {tmed(cm.last_window):.1f} us, synchronize on an idle stream "
      f"{tmed(cm.synchronize):.1f} us, get_changed with nothing to do {tmed(lambda: cm.get_changed(mirror)):.1f} us")
