"""Measurement helper: D2H of the 4000^2 window, cudaMemcpy2DAsync (pitched -> packed) vs pack kernel + linear copy."""
import ctypes as C
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
size = 4000
cm, (s, o, il), sets = bench.build_c3(lambda *a: api.costmap(*a), size=size)
obs, robot = sets[0]
cm.set_observations(o, obs)
cm.update_map(*robot)
pin = torch.empty(size * size, dtype=torch.uint8).pin_memory()
buf = pin.numpy().reshape(size, size)
u8p, i8p = C.POINTER(C.c_uint8), C.POINTER(C.c_int8)
for name, fn, ptr in (("get_window (2D copy)", api.lib.navgpu_costmap_get_window, buf.ctypes.data_as(u8p)),
                      ("get_window_occupancy (pack kernel + linear copy)", api.lib.navgpu_costmap_get_window_occupancy,
                       buf.ctypes.data_as(i8p))):
    for _ in range(3):
        fn(cm.h, 0, 0, size, size, ptr)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        fn(cm.h, 0, 0, size, size, ptr)
    dt = (time.perf_counter() - t0) / n
    print(f"{name}: {dt * 1e3:.3f} ms = {size * size / dt / 1e9:.1f} GB/s")
