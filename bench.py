#!/usr/bin/env python
"""bench.py -- measures the two hot paths on B200 and prints ONE JSON line (contract in the task statement).

Headline metric (BASELINE.json): ms per updateMap+inflation @4k^2 grid (config C3: 4000x4000 @0.05 m warehouse map,
static + obstacle layer fed by 8 observations x 360 ray-cast beams, inflation 1.0 m => R = 20 cells, full window);
the DWA half of the metric (trajectories scored / s, configs C2 and C4) is reported in the same line under "dwa".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

* value   : device time per update cycle with every input already resident in HBM (CUDA events on the costmap's own
            stream, L2 flushed between cycles), max over ranks.  N > 1 runs N independent replicas (a single costmap
            does not shard); the DWA sweep is sharded by sample range with an all-gather.
* e2e     : the same cycle through the C ABI with HOST buffers: scan upload + update + read-back of the updated
            window into pinned memory inside the timed region.
* roofline: the fused reset+merge+inflation sweep kernel against the measured HBM copy bandwidth.
* cpu_baseline / --impl reference: the reference's own CPU code (oracle/_ref, else our port) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(size=4000, resolution=0.05, n_obs=8, n_beams=360, scan_range=10.0, inflation_radius=1.0, scaling=10.0)
ALGO_BYTES_PER_CELL = 3  # read static + read obstacle + write master
# What the roofline objects quote from ncu (DRAM traffic, executed warp instructions, issue-slot utilisation of the
# dominant kernels) is not typed here: tools/roofline_inputs.py extracts it from the .ncu-rep captures of the same
# workloads (tools/profile_r2.sh) into profiles/r2_roofline_inputs.json, which this file reads.
METRIC = "ms per updateMap+inflation @4k^2 grid; DWA trajectories scored/sec"
WORKLOAD = ("C3 full-window updateMap 4000x4000 @0.05 m: static + obstacle (8 obs x 360 beams, 10 m raytrace+mark) + "
            "inflation 1.0 m (R=20); DWA half: C2 findBestPath 20x1x20, C4 sweep 200x20x200 on a 120x120 local map, "
            "C5 fleet of 4096 robots (local-map inflation + 20x1x20 scoring each)")


def env_int(name, default):
    return int(os.environ.get(name, default))


def roofline_inputs():
    p = os.path.join(ROOT, "profiles", "r2_roofline_inputs.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for k, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[k].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def build_c3(api_costmap_factory, size=None, n_obs=None):
    """Builds the C3 layer stack on `api` and returns (costmap, ids, observation sets, robot pose)."""
    from navigation_b200 import synth
    size = size or C3["size"]
    sets = []
    static = None
    for cyc in range(4):
        static, obs, robot, fp = synth.warehouse_c3(size=size, resolution=C3["resolution"], n_obs=n_obs or C3["n_obs"],
                                                    n_beams=C3["n_beams"], scan_range=C3["scan_range"], cycle=cyc)
        sets.append((obs, robot))
    cm = api_costmap_factory(size, size, C3["resolution"])
    s = cm.add_grid_layer(0)  # StaticLayer, TrueOverwrite
    o = cm.add_obstacle_layer(1, True, 2.0)
    il = cm.add_inflation_layer(C3["inflation_radius"], C3["scaling"])
    cm.set_footprint(fp)
    cm.set_grid_layer(s, static)
    return cm, (s, o, il), sets


def voxel_numbers(api_factory, reps, sync=None, size=None):
    """The C3 recipe with a VoxelLayer (cfg/VoxelPlugin.cfg defaults: 10 voxels of 0.2 m) in place of the obstacle
    layer: wall time of one full-window update_map through the binding (device work + launch overhead)."""
    from navigation_b200 import synth
    size = size or C3["size"]
    static, obs, robot, fp = synth.warehouse_c3(size=size, resolution=C3["resolution"], n_obs=C3["n_obs"],
                                                n_beams=C3["n_beams"], scan_range=C3["scan_range"])
    for o in obs:  # give the ray-cast end points heights inside the 2 m column
        o["origin"] = (o["origin"][0], o["origin"][1], 0.5)
        o["points"][:, 2] = (0.1 + 1.7 * (np.arange(len(o["points"])) % 7) / 7.0).astype(np.float32)
    cm = api_factory(size, size, C3["resolution"])
    s = cm.add_grid_layer(0)
    v = cm.add_voxel_layer(1, True, 2.0, 0.0, 0.2, 10, 15, 0)
    cm.add_inflation_layer(C3["inflation_radius"], C3["scaling"])
    cm.set_footprint(fp)
    cm.set_grid_layer(s, static)
    cm.set_observations(v, obs)
    cm.update_map(*robot)
    t0 = time.perf_counter()
    for _ in range(reps):
        cm.touch_grid_layer(s, 0, 0, size, size)
        cm.update_map(*robot)
    if sync:
        sync()
    return 1e3 * (time.perf_counter() - t0) / reps


def obs_bytes(obs):
    return sum(o["points"].nbytes + 64 for o in obs)



C2 = dict(n=120, resolution=0.05, vx_samples=20, vy_samples=1, vth_samples=20)
C4 = dict(n=120, resolution=0.05, vx_samples=200, vy_samples=20, vth_samples=200)
PENTAGON = [(-0.325, -0.325), (-0.325, 0.325), (0.325, 0.325), (0.46, 0.0), (0.325, -0.325)]


C1 = dict(size=400, resolution=0.05, inscribed=0.325, inflation_radius=0.55, scaling=10.0)
SQUARE = [(0.325, 0.325), (0.325, -0.325), (-0.325, -0.325), (-0.325, 0.325)]
DWA_BYTES_PER_TRAJECTORY = 20  # 12 B sample in + 8 B cost out


def c1_stack(api, **kw):
    """BASELINE.json configs[0]: InflationLayer on a synthetic 400x400 @0.05 static map, inscribed 0.325 / inflation
    0.55 (R = 11), as a layered costmap (static layer + inflation)."""
    from navigation_b200 import synth
    g = synth.blocks_c1(C1["size"])
    cm = api.costmap(C1["size"], C1["size"], C1["resolution"], **kw)
    s = cm.add_grid_layer(0)
    cm.add_inflation_layer(C1["inflation_radius"], C1["scaling"])
    cm.set_footprint(SQUARE)
    cm.set_grid_layer(s, g)
    return cm, s, g


def seam_numbers(api, master0, radius_m, device, reps):
    """The costmap_2d::Layer plugin seam: navgpu_inflate_host = the body of GpuInflationLayer::updateCosts on the HOST
    master grid LayeredCostmap owns (upload of the window +- 2R rows, sweep, download), whole-map window; wall time per
    call with the master in ordinary pageable memory (what Costmap2D allocates) and page-locked (navgpu_host_register)."""
    n = master0.shape[0]
    R, costs, _ = api.build_cost_table(C1["resolution"], C1["inscribed"], radius_m, C1["scaling"])
    out = {}
    for name in ("pageable", "registered"):
        master = master0.copy()
        if name == "registered":
            api.host_register(master)
        for _ in range(2):
            master[...] = master0
            api.inflate_host(master, 0, 0, n, n, costs, R, device=device)
        ts = []
        for _ in range(reps):
            master[...] = master0
            t0 = time.perf_counter()
            api.inflate_host(master, 0, 0, n, n, costs, R, device=device)
            ts.append(time.perf_counter() - t0)
        if name == "registered":
            api.host_unregister(master)
        out[f"inflate_host_{name}_ms"] = 1e3 * float(np.median(ts))
    out["bytes_each_way"] = int(master0.size)
    out["cell_inflation_radius"] = int(R)
    return out, master


def c1_numbers(api, torch, local_rank, reps):
    """Config C1 on the device-resident path (static layer + inflation, full window, CUDA events, L2 flushed / hot) and
    through the plugin seam."""
    cm, s, g = c1_stack(api, device=local_rank)
    n = C1["size"]
    stream = torch.cuda.ExternalStream(cm.stream(), device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    for _ in range(3):
        cm.touch_grid_layer(s, 0, 0, n, n)
        cm.update_map()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        with torch.cuda.stream(stream):
            flush.zero_()
            a.record(stream)
        cm.touch_grid_layer(s, 0, 0, n, n)
        cm.update_map_async()
        b.record(stream)
    torch.cuda.synchronize()
    cold = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    a, b = ev[0]
    a.record(stream)
    for _ in range(reps):
        cm.touch_grid_layer(s, 0, 0, n, n)
        cm.update_map_async()
    b.record(stream)
    torch.cuda.synchronize()
    hot = a.elapsed_time(b) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        cm.touch_grid_layer(s, 0, 0, n, n)
        cm.update_map()
    sync_ms = 1e3 * (time.perf_counter() - t0) / reps
    seam, seam_out = seam_numbers(api, g, C1["inflation_radius"], local_rank, reps)
    same = bool(np.array_equal(seam_out, cm.get()))
    out = {"what": "BASELINE.json configs[0]: 400x400 @0.05 static map, inscribed 0.325 / inflation 0.55 (R = 11), "
                   "full-window updateMap (static layer + inflation)",
           "device_ms_l2_flushed": cold, "device_ms_hot": hot, "update_map_sync_call_ms": sync_ms,
           "algorithmic_bytes": 2 * n * n, "hbm_frac_hot": 2 * n * n / (hot * 1e-3) / 1e9 / measured_peak_gbs()[0],
           "plugin_seam": seam, "plugin_seam_equals_device_path": same}
    return out


def c1_cpu_numbers():
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    cm, s, g = c1_stack(pyoracle.load(kind))
    n = C1["size"]
    cm.update_map()
    ts = []
    for _ in range(5):
        cm.touch_grid_layer(s, 0, 0, n, n)
        t0 = time.perf_counter()
        cm.update_map()
        ts.append(time.perf_counter() - t0)
    return {"kind": kind, "cores": 1, "update_map_ms": 1e3 * float(np.median(ts)),
            "sample": "median of 5 full-window updates of the C1 stack, single thread"}


def seam_c3_numbers(api, local_rank, reps=5):
    """The same seam at C3 size: a 4000x4000 host master holding the static warehouse, R = 20."""
    from navigation_b200 import synth
    static, _, _, _ = synth.warehouse_c3(size=C3["size"], n_obs=1)
    seam, _ = seam_numbers(api, static, C3["inflation_radius"], local_rank, reps)
    return seam


def dwa_prepare_cpu_numbers():
    """MapGridCostFunction::prepare x 4 of a C2 findBestPath on one host core: as the reference is, and with the one
    line of map_grid.cpp:106 that copy-constructs the whole Costmap2D per BFS cell replaced by a const_cast (the
    `reference_hoisted` checker, oracle/Makefile) -- so that the GPU is not compared against that copy only."""
    from oracle import pyoracle
    out = {}
    for kind, key in (("reference", ""), ("reference_hoisted", "_hoisted")):
        if not pyoracle.available(kind):
            continue
        api = pyoracle.load(kind)
        grid = inflate_local(api, local_map_c2())
        d, pose, vel = dwa_setup(api, grid, C2)
        d.find_best_path(pose, vel, PENTAGON)
        t0 = time.perf_counter()
        for _ in range(10):
            d.prepare_only()
        out[f"prepare{key}_ms"] = 1e3 * (time.perf_counter() - t0) / 10
        t0 = time.perf_counter()
        for _ in range(10):
            d.find_best_path(pose, vel, PENTAGON)
        out[f"c2_findBestPath{key}_ms"] = 1e3 * (time.perf_counter() - t0) / 10
    return out


def _c4_worker(_):
    from oracle import pyoracle
    kind = "reference_hoisted" if pyoracle.available("reference_hoisted") else (
        "reference" if pyoracle.available("reference") else "port")
    api = pyoracle.load(kind)
    grid = inflate_local(api, local_map_c2())
    d4, pose, vel = dwa_setup(api, grid, dict(C4, vx_samples=C4["vx_samples"] // 4))
    t0 = time.perf_counter()
    r = d4.find_best_path(pose, vel, PENTAGON)
    return time.perf_counter() - t0, int(r["n_samples"]), kind


def c4_all_cores_numbers():
    """C4 on every host core at once: one process per core, each scoring a quarter-size sweep (50 x 20 x 200 samples of
    the same window) with the reference's code (the hoisted variant: prepare() is negligible either way at this
    size); aggregate trajectories / s over the slowest process's time."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_c4_worker, range(cores), chunksize=1)
    dt = max(r[0] for r in res)
    n = sum(r[1] for r in res)
    return {"kind": res[0][2], "cores": cores, "c4_traj_per_s": n / dt, "samples": n,
            "sample": "one 50x20x200-sample sweep per process, one process per host core concurrently; slowest "
                      "process's time"}


def local_map_c2():
    """6 m x 6 m local costmap @0.05: corridor walls + one box, inflated (radius 0.55, scaling 10) by the product's own
    costmap path when a GPU is present, else by the checker (reference arm)."""
    g = np.zeros((120, 120), np.uint8)
    g[26:29, :] = 254
    g[91:94, :] = 254
    g[52:58, 84:90] = 254
    return g


def inflate_local(api, g, **kw):
    cm = api.costmap(120, 120, 0.05, **kw)
    s = cm.add_grid_layer(0)
    cm.add_inflation_layer(0.55, 10.0)
    cm.set_footprint(PENTAGON)
    cm.set_grid_layer(s, g)
    cm.update_map(0, 0, 0)
    return cm.get()


def dwa_setup(api, grid, cfg, **kw):
    over = dict(vx_samples=cfg["vx_samples"], vy_samples=cfg["vy_samples"], vth_samples=cfg["vth_samples"])
    if cfg is C2:
        over.update(max_vel_y=0.0, min_vel_y=0.0)
    else:  # C4: wide dynamic window
        over.update(acc_lim_x=20.0, acc_lim_y=20.0, acc_lim_theta=20.0, max_vel_y=0.1, min_vel_y=-0.1)
    d = api.dwa(120, 120, 0.05, **over, **kw)
    d.set_costmap(grid, 0.0, 0.0)
    pose, vel = (1.5, 3.0, 0.0), (0.3, 0.0, 0.0)
    plan = np.stack([np.arange(1.0, 7.0, 0.05), np.full(120, 3.0)], 1)
    d.set_plan(pose, plan)
    return d, pose, vel


def tp_numbers(api, grid, reps=20, **kw):
    """Legacy base_local_planner::TrajectoryPlanner on the C2 local map: 20 x 20 forward samples + 2
    holonomic + 20 in-place + 4 strafing + 1 back-up = 427 rollouts per findBestPath, through the synchronous call."""
    tp = api.trajectory_planner(120, 120, 0.05, PENTAGON, vx_samples=20, vtheta_samples=20, sim_time=1.7, **kw)
    tp.set_costmap(grid, 0.0, 0.0)
    tp.update_plan(np.stack([np.arange(1.0, 7.0, 0.05), np.full(120, 3.0)], 1))
    pose, vel = (1.5, 3.0, 0.0), (0.3, 0.0, 0.0)
    for _ in range(3):
        r = tp.find_best_path(pose, vel)
    t0 = time.perf_counter()
    for _ in range(reps):
        r = tp.find_best_path(pose, vel)
    us = 1e6 * (time.perf_counter() - t0) / reps
    return {"findBestPath_us": us, "samples": 427, "traj_per_s": 427 / (us * 1e-6), "best": [r["xv"], r["yv"], r["thetav"]],
            "best_cost": r["cost"]}


def dwa_cpu_numbers(budget_s=25.0):
    """Reference CPU code on C2 (latency) and on C4 (throughput, full sweep once: about 20 s on one core)."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    api = pyoracle.load(kind)
    grid = inflate_local(api, local_map_c2())
    d, pose, vel = dwa_setup(api, grid, C2)
    d.find_best_path(pose, vel, PENTAGON)
    t0 = time.perf_counter()
    reps = 0
    while reps < 20 and time.perf_counter() - t0 < 3.0:
        r = d.find_best_path(pose, vel, PENTAGON)
        reps += 1
    c2_ms = 1e3 * (time.perf_counter() - t0) / reps
    d4, pose, vel = dwa_setup(api, grid, C4)
    t0 = time.perf_counter()
    r4 = d4.find_best_path(pose, vel, PENTAGON)
    c4_s = time.perf_counter() - t0
    return {"kind": kind, "cores": 1, "c2_findBestPath_ms": c2_ms, "c2_samples": int(r["n_samples"]),
            "c4_traj_per_s": r4["n_samples"] / c4_s, "c4_samples": int(r4["n_samples"]), "c4_s": c4_s,
            "c4_best_index": int(r4["best_index"]), "c2_best_index": int(r["best_index"]),
            "sample": "C2: mean of up to 20 findBestPath calls; C4: one full 200x20x200 sweep, single thread"}

def run_reference(args, rank):
    """The reference's CPU implementation of the path on the host cores (single-threaded per costmap, as the reference
    is).  Each step is one C3 update at full size when K+W of them fit in about four minutes, else the C3 recipe on a
    1000x1000 crop scaled by the cell ratio."""
    if rank != 0:
        return
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    api = pyoracle.load(kind)
    full = (args.steps + args.warmup) * 4.0 < 240.0
    sample = C3["size"] if full else 1000
    cm, (s, o, il), sets = build_c3(api.costmap, size=sample)
    scale = (C3["size"] / sample) ** 2
    times = []
    for step in range(args.warmup + args.steps):
        obs, robot = sets[step % len(sets)]
        t0 = time.perf_counter()
        cm.set_observations(o, obs)
        cm.touch_grid_layer(s, 0, 0, sample, sample)
        cm.update_map(*robot)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times)) * scale
    what = ("one full C3 update per step" if full else
            f"C3 recipe on a {sample}x{sample} crop per step, time scaled by the cell ratio x{scale:g}")
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": kind, "sample": what},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_dwa:
        line["dwa"] = dwa_cpu_numbers()
        line["dwa"]["c5"] = fleet_cpu_numbers()
        line["dwa"]["c5_all_cores"] = fleet_cpu_parallel_numbers()
    print(json.dumps(line), flush=True)


def cpu_baseline_c3():
    """One full 4000x4000 C3 update on one host core with the reference's own code (about 10-20 s)."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    api = pyoracle.load(kind)
    cm, (s, o, il), sets = build_c3(api.costmap)
    obs, robot = sets[0]
    cm.set_observations(o, obs)
    t0 = time.perf_counter()
    cm.update_map(*robot)
    ms = 1e3 * (time.perf_counter() - t0)
    return {"value": ms, "unit": "ms", "cores": 1, "kind": kind,
            "sample": "one full C3 update (4000x4000, R=20, 8x360 beams), single thread as the reference runs it"}



def run_native_dwa(api, torch, dist, rank, world, local_rank, steps):
    """C2 latency (synchronous findBestPath through the C ABI, host pose in / result out) and C4 throughput.
    With world > 1 the C4 sample range is sharded over the ranks: each scores its slice, the per-rank (cost, index)
    minima are all-gathered over NCCL and every rank finishes with the same winner."""
    dev = f"cuda:{local_rank}"
    grid = inflate_local(api, local_map_c2(), device=local_rank)
    out = {}
    d2, pose, vel = dwa_setup(api, grid, C2, device=local_rank)
    call2 = d2.prepared(pose, vel, PENTAGON)
    for _ in range(5):
        r2 = call2.find_best_path()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(20, steps)
    for _ in range(reps):
        r2 = call2.find_best_path()
    out["c2_findBestPath_us"] = 1e6 * (time.perf_counter() - t0) / reps
    out["c2_samples"] = int(r2["n_samples"])
    out["c2_best_index"] = int(r2["best_index"])
    out["c2_traj_per_s"] = r2["n_samples"] / (out["c2_findBestPath_us"] * 1e-6)
    # MapGridCostFunction::prepare x 4 alone (navgpu_dwa_prepare: the MapGrid wavefront kernel), CUDA events on the
    # planner's stream -- the device-side counterpart of cpu_baseline.prepare_ms
    s2 = torch.cuda.ExternalStream(d2.stream(), device=local_rank)
    pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in pe:
        a.record(s2)
        d2.prepare()
        b.record(s2)
        d2.synchronize()
    out["c2_prepare_device_us"] = 1e3 * float(np.mean([a.elapsed_time(b) for a, b in pe]))

    d4, pose, vel = dwa_setup(api, grid, C4, device=local_rank)
    stream = torch.cuda.ExternalStream(d4.stream(), device=local_rank)
    from navigation_b200 import sharding
    call4 = d4.prepared(pose, vel, PENTAGON)  # arguments marshalled once, as a C++ caller holds them
    if dist is None:
        one_sweep = call4.find_best_path
    else:
        # one-off: the ranks exchange the CUDA IPC handles of their 1 KB exchange buffers; from then on a sweep is one
        # call per rank and the scoring kernels trade their (cost, index) minima over NVLink themselves
        sharding.connect_shards(dist, d4)

        one_sweep = call4.find_best_path_sharded

    for _ in range(3):
        r4 = one_sweep()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_sweeps = max(5, steps)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(n_sweeps):
        r4 = one_sweep()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n_sweeps
    sweep_s = max(wall, 1e-3 * e0.elapsed_time(e1) / n_sweeps)
    if dist is not None:
        t = torch.tensor([sweep_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sweep_s = float(t.item())
    total = int(r4["n_samples"])
    # roofline of the DWA half: k_dwa_score is bound by instruction issue, not by HBM (its inputs -- a 14.4 KB costmap
    # and four 57.6 KB distance grids -- live in L1 / L2); both fractions are given
    peak_gbs, _ = measured_peak_gbs()
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_hz = 1e6 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"]) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1.965e9
    issue_peak = sms * 4 * sm_hz * world  # one warp instruction per scheduler and cycle, 4 schedulers per SM
    ncu = (roofline_inputs().get("dwa_c4") or [{}])[0]
    winst = ncu.get("warp_instructions")
    out["roofline"] = {
        "kernel": "k_dwa_score (C4 sweep)", "bound": "issue",
        "achieved": winst / sweep_s / 1e9 if winst else None, "peak": issue_peak / 1e9, "unit": "G warp-instr/s",
        "frac": winst / sweep_s / issue_peak if winst else None,
        "warp_instructions_per_sweep": winst, "ncu_issue_active_pct": ncu.get("issue_active_pct"),
        "warp_instructions_source": "smsp__inst_executed.sum of the same sweep under ncu "
                                    "(profiles/r2_roofline_inputs.json), divided by this run's sweep time",
        "hbm": {"algorithmic_bytes": DWA_BYTES_PER_TRAJECTORY * total, "traffic": ncu.get("dram_bytes"),
                "achieved_gbs": DWA_BYTES_PER_TRAJECTORY * total / sweep_s / 1e9, "peak_gbs": peak_gbs * world,
                "frac": DWA_BYTES_PER_TRAJECTORY * total / sweep_s / 1e9 / (peak_gbs * world)}}
    out.update({"c4_samples": int(total), "c4_sweep_ms": 1e3 * sweep_s, "c4_traj_per_s": total / sweep_s,
                "c4_best_index": int(r4["best_index"]), "c4_best_cost": float(r4["cost"]),
                "c4_sharding": (f"8-sample blocks dealt round-robin over {world} ranks; (cost, index) minima exchanged "
                                "by the scoring kernels' last CTAs through peer-mapped buffers (no host step, no "
                                "collective call)") if world > 1 else "single GPU"})
    return out


C5 = dict(n_robots=4096, n=120, resolution=0.05, vx_samples=20, vy_samples=1, vth_samples=20)


def fleet_inputs(ids):
    from navigation_b200 import synth
    robots = [synth.fleet_robot(i) for i in ids]
    return (np.stack([r["raw"] for r in robots]), np.array([r["origin"] for r in robots]),
            np.array([r["pose"] for r in robots]), np.array([r["vel"] for r in robots]), [r["plan"] for r in robots])


def run_native_fleet(api, torch, dist, rank, world, local_rank, steps):
    """C5: 4096 independent robots per control cycle, partitioned over the ranks (no collective): every robot's raw
    local map is inflated and its 20 x 1 x 20 velocity samples are scored.  Timed through the C ABI with host buffers:
    `step` = poses/velocities in, results out (maps already resident); `e2e` additionally uploads all raw maps."""
    from navigation_b200 import sharding
    n_total = C5["n_robots"]
    lo, hi = sharding.split_range(n_total, rank, world)
    raw, origins, poses, vels, plans = fleet_inputs(range(lo, hi))
    fleet = api.fleet(hi - lo, 120, 120, 0.05, PENTAGON, 0.55, 10.0, device=local_rank, vx_samples=C5["vx_samples"],
                      vy_samples=C5["vy_samples"], vth_samples=C5["vth_samples"], max_vel_y=0.0, min_vel_y=0.0)
    fleet.set_maps(raw, origins)
    fleet.set_plans(poses, plans)
    poses = np.ascontiguousarray(poses)
    vels = np.ascontiguousarray(vels)
    for _ in range(3):
        fleet.step_raw(poses, vels)
    reps = max(5, min(steps, 10))

    def timed(fn):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt], device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    step_s = timed(lambda: fleet.step_raw(poses, vels))

    # the raw maps of the e2e loop come from pinned host memory (as the contract asks): same bytes, page-locked
    raw_pinned_t = torch.empty(raw.shape, dtype=torch.uint8, pin_memory=True)
    raw_pinned = raw_pinned_t.numpy()
    raw_pinned[...] = raw

    def full():
        fleet.set_maps(raw_pinned, origins)
        fleet.step_raw(poses, vels)
    e2e_s = timed(full)
    res = fleet.step(poses, vels)
    n_traj = sum(r["n_samples"] for r in res)
    n_valid = sum(r["cost"] >= 0 for r in res)
    if dist is not None:
        t = torch.tensor([float(n_traj), float(n_valid)], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t)
        n_traj, n_valid = int(t[0].item()), int(t[1].item())
    return {"c5_robots": n_total, "c5_cycle_ms": 1e3 * step_s, "c5_robot_cycles_per_s": n_total / step_s,
            "c5_traj_per_s": n_traj / step_s, "c5_e2e_cycle_ms": 1e3 * e2e_s,
            "c5_e2e_robot_cycles_per_s": n_total / e2e_s, "c5_h2d_bytes_per_e2e_cycle": int(raw.nbytes) * world,
            "c5_partitioning": f"robots split over {world} rank(s), no collective" if world > 1 else "single GPU",
            "c5_valid_robots": int(n_valid)}


def fleet_cpu_numbers(n_robots=8):
    """The reference's CPU code on a few fleet robots: LayeredCostmap inflation + findBestPath per robot, one core."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    api = pyoracle.load(kind)
    raw, origins, poses, vels, plans = fleet_inputs(range(n_robots))
    t0 = time.perf_counter()
    for i in range(n_robots):
        cm = api.costmap(120, 120, 0.05, *origins[i])
        s = cm.add_grid_layer(0)
        cm.add_inflation_layer(0.55, 10.0)
        cm.set_footprint(PENTAGON)
        cm.set_grid_layer(s, raw[i])
        cm.update_map(0, 0, 0)
        d = api.dwa(120, 120, 0.05, vx_samples=C5["vx_samples"], vy_samples=C5["vy_samples"],
                    vth_samples=C5["vth_samples"], max_vel_y=0.0, min_vel_y=0.0)
        d.set_costmap(cm.get(), *origins[i])
        d.set_plan(poses[i], plans[i])
        d.find_best_path(poses[i], vels[i], PENTAGON)
    dt = time.perf_counter() - t0
    return {"kind": kind, "cores": 1, "c5_robot_cycles_per_s": n_robots / dt,
            "sample": f"{n_robots} fleet robots (inflation + findBestPath each), single thread"}


def _fleet_worker(ids):
    return fleet_cpu_numbers_ids(ids)


def fleet_cpu_numbers_ids(ids):
    """Seconds the reference's CPU code needs for the fleet robots `ids` (one process, one core)."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    api = pyoracle.load(kind)
    raw, origins, poses, vels, plans = fleet_inputs(ids)
    t0 = time.perf_counter()
    for i in range(len(ids)):
        cm = api.costmap(120, 120, 0.05, *origins[i])
        s = cm.add_grid_layer(0)
        cm.add_inflation_layer(0.55, 10.0)
        cm.set_footprint(PENTAGON)
        cm.set_grid_layer(s, raw[i])
        cm.update_map(0, 0, 0)
        d = api.dwa(120, 120, 0.05, vx_samples=C5["vx_samples"], vy_samples=C5["vy_samples"],
                    vth_samples=C5["vth_samples"], max_vel_y=0.0, min_vel_y=0.0)
        d.set_costmap(cm.get(), *origins[i])
        d.set_plan(poses[i], plans[i])
        d.find_best_path(poses[i], vels[i], PENTAGON)
    return time.perf_counter() - t0


def fleet_cpu_parallel_numbers(per_core=6):
    """The same on every host core at once (robots are independent, so the fleet is embarrassingly parallel on a CPU
    too): one process per core, `per_core` robots each; aggregate robot-cycles/s."""
    import multiprocessing as mp
    from oracle import pyoracle
    kind = "reference" if pyoracle.available("reference") else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    chunks = [list(range(c * per_core, (c + 1) * per_core)) for c in range(cores)]
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context and torch threads: no fork
    with ctx.Pool(cores) as pool:
        pool.map(_fleet_worker, [c[:1] for c in chunks])  # processes up, libraries loaded
        dt = max(pool.map(_fleet_worker, chunks, chunksize=1))
    return {"kind": kind, "cores": cores, "c5_robot_cycles_per_s": cores * per_core / dt,
            "c5_traj_per_s": cores * per_core * 420 / dt,
            "sample": f"{per_core} fleet robots per process, one process per host core running concurrently; "
                      "slowest process's compute time"}


def run_native(args, rank, world, local_rank):
    import torch
    import navigation_b200
    from navigation_b200 import build
    build.build()
    api = navigation_b200.load()
    if api.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; libnavgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    size = C3["size"]
    cm, (s, o, il), sets = build_c3(lambda *a: api.costmap(*a, device=local_rank))
    stream = torch.cuda.ExternalStream(cm.stream(), device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2
    n_cells = size * size

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input loop (value): scans already on the device, L2 flushed before every cycle, the handle's own
    # profiling events OFF (they cost a few microseconds per cycle)
    obs, robot = sets[0]
    cm.set_observations(o, obs)
    for _ in range(max(3, args.warmup)):
        cm.touch_grid_layer(s, 0, 0, size, size)
        cm.update_map(*robot)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = api.launch_count()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
            starts[k].record(stream)
        cm.touch_grid_layer(s, 0, 0, size, size)  # has_updated_data_: forces the full window
        cm.update_map_async(*robot)
        stops[k].record(stream)
    barrier()
    launches = api.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    ms_per_step = float(np.mean(step_ms))

    # ---- the same cycles again with the handle's events ON: durations of the sweep kernels for the roofline
    cm.set_profiling(True)
    sweep_ms, merge_ms, inflate_ms = [], [], []
    for k in range(max(5, args.steps)):
        with torch.cuda.stream(stream):
            flush.zero_()
        cm.touch_grid_layer(s, 0, 0, size, size)
        cm.update_map_async(*robot)
        sweep_ms.append(cm.last_timing()[1])
        m, i = cm.last_timing_split()
        merge_ms.append(m)
        inflate_ms.append(i)
    cm.set_profiling(False)

    # ---- hot-L2 loop (informational): back-to-back cycles without the flush
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        cm.touch_grid_layer(s, 0, 0, size, size)
        cm.update_map_async(*robot)
    e1.record(stream)
    barrier()
    hot_ms = e0.elapsed_time(e1) / args.steps

    # ---- end-to-end loop: host scans in, host master grid out, through the C ABI.  The host keeps a page-locked mirror
    # of the master grid (the Costmap2D the rest of the stack reads); per cycle: navgpu_obstacle_set_observations
    # (H2D), navgpu_costmap_update_map_async, navgpu_costmap_get_changed (the tiles that differ from the mirror come
    # back through mapped pinned memory and are scattered into it; one host wait per cycle).  The observation set
    # changes every cycle (4 sensor positions in rotation), so the changed tiles are real.
    pinned = torch.empty((size, size), dtype=torch.uint8, pin_memory=True)
    out_np = pinned.numpy()
    # the navgpu_observation structs a C++ caller would hand over (Python marshalling is not part of the path)
    packed = [cm.pack_observations(ob) for ob, _ in sets]

    cycle = cm.prepared_cycle(o, s, out_np)

    def e2e_cycle(k, full_window):
        ob, rb = sets[k % len(sets)]
        if full_window:
            cm.set_packed_observations(o, packed[k % len(sets)])
            cm.touch_grid_layer(s, 0, 0, size, size)  # round 1's path: synchronous update, then the whole window over PCIe
            w = cm.update_map(*rb)
            cm.get_window_into(w[0], w[2], w[1], w[3], out_np)
            return obs_bytes(ob), (w[1] - w[0]) * (w[3] - w[2]) + 32
        return obs_bytes(ob), cycle(packed[k % len(sets)], rb)  # the four C-ABI calls, arguments marshalled once

    e2e = {}
    for name, full_window in (("full_window", True), ("changed_tiles", False)):
        for k in range(4):
            e2e_cycle(k, full_window)
        barrier()
        h2d = d2h = 0
        e0.record(stream)
        t0 = time.perf_counter()
        for k in range(args.steps):
            a, b = e2e_cycle(k, full_window)
            h2d += a
            d2h += b
        e1.record(stream)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        e2e[name] = (max(e0.elapsed_time(e1) / args.steps, wall_ms), h2d // args.steps, d2h // args.steps)
    # the mirror must be the device grid, byte for byte
    assert np.array_equal(out_np, cm.get()), "host mirror differs from the device master grid"
    e2e_ms, h2d, d2h = e2e["changed_tiles"]
    e2e_full_ms = e2e["full_window"][0]
    dwa = None if args.no_dwa else run_native_dwa(api, torch, dist, rank, world, local_rank, args.steps)
    if dwa is not None:
        dwa.update(run_native_fleet(api, torch, dist, rank, world, local_rank, args.steps))
    launches = api.launch_count() - launches0 if not args.no_dwa else launches
    clocks = sampler.summary()

    if dist is not None:
        t = torch.tensor([ms_per_step, e2e_ms, hot_ms, float(np.mean(sweep_ms)), float(np.mean(merge_ms)),
                          float(np.mean(inflate_ms)), e2e_full_ms], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms, hot_ms, sweep, merge, inflate, e2e_full_ms = [float(v) for v in t.tolist()]
        lt = torch.tensor([launches], device=f"cuda:{local_rank}")
        dist.all_reduce(lt)
        launches = int(lt.item())
    else:
        sweep, merge, inflate = float(np.mean(sweep_ms)), float(np.mean(merge_ms)), float(np.mean(inflate_ms))

    c1 = c1_numbers(api, torch, local_rank, max(10, args.steps)) if rank == 0 else None
    seam_c3 = seam_c3_numbers(api, local_rank) if rank == 0 else None
    voxel_ms = voxel_numbers(lambda *a: api.costmap(*a, device=local_rank), max(5, args.steps), torch.cuda.synchronize)
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = ALGO_BYTES_PER_CELL * n_cells / (sweep * 1e-3) / 1e9
        ncu_sweep = roofline_inputs().get("sweep") or []
        sweep_traffic = sum(k["dram_bytes"] for k in ncu_sweep) or None
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        sweep_kernels = []
        for k, ms in zip(ncu_sweep, (merge, inflate)):
            issue_peak = sms * 4 * 1e6 * (clocks.get("sm_mhz") or 1965.0)
            sweep_kernels.append({"kernel": k["kernel"], "ms": ms, "traffic": k["dram_bytes"],
                                  "hbm_frac_on_traffic": k["dram_bytes"] / (ms * 1e-3) / 1e9 / peak,
                                  "warp_instructions": k["warp_instructions"],
                                  "issue_frac": k["warp_instructions"] / (ms * 1e-3) / issue_peak,
                                  "ncu_issue_active_pct": k["issue_active_pct"]})
        line = {
            "metric": METRIC, "value": ms_per_step, "unit": "ms",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "l2": "flushed (256 MiB memset) before every timed cycle",
                       "value_excludes": "the 35 KB upload of the cycle's observations (resident when the timed region "
                                         "starts); e2e includes it",
                       "multi_gpu": "replicas only for the costmap path" if world > 1 else "single GPU",
                       "hot_l2_ms_per_step": hot_ms},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "navgpu_obstacle_set_observations + navgpu_costmap_update_map_async + "
                            "navgpu_costmap_get_changed into a page-locked host mirror of the master grid (checked "
                            "byte-identical to the device grid after the loop); the observation set changes every cycle",
                    "full_window_download_ms": e2e_full_ms},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": sweep_traffic, "kernel": "k_merge_seed + k_inflate (the reset+merge+inflation sweep)",
                         "kernel_ms": sweep, "k_merge_seed_ms": merge, "k_inflate_ms": inflate,
                         "algorithmic_bytes": ALGO_BYTES_PER_CELL * n_cells, "peak_source": peak_src,
                         "frac_of_whole_cycle": ALGO_BYTES_PER_CELL * n_cells / (ms_per_step * 1e-3) / 1e9 / peak,
                         "per_kernel": sweep_kernels,
                         "note": "kernel_ms comes from a separate loop with CUDA events between the two kernels (which "
                                 "serialises them); in the timed cycles k_merge_seed overlaps k_obstacle_update and "
                                 "hands its tiles to k_inflate one by one.  traffic / issue figures: ncu capture of "
                                 "the same cycle (profiles/r2_roofline_inputs.json); the sweep is bound by instruction "
                                 "issue (k_inflate), not by HBM"},
        }
        line["voxel_layer"] = {"c3_with_voxel_layer_update_map_ms": voxel_ms,
                               "what": "C3 with costmap_2d::VoxelLayer (10 x 0.2 m voxels) instead of ObstacleLayer, "
                                       "synchronous navgpu_costmap_update_map, hot L2"}
        line["c1"] = c1
        line["plugin_seam_c3"] = seam_c3
        if dwa is not None:
            line["dwa"] = dwa
            line["trajectory_planner"] = tp_numbers(api, inflate_local(api, local_map_c2(), device=local_rank),
                                                    device=local_rank)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_c3()
            from oracle import pyoracle
            kind = "reference" if pyoracle.available("reference") else "port"
            line["voxel_layer"]["cpu_ms"] = voxel_numbers(pyoracle.load(kind).costmap, 1)
            line["voxel_layer"]["cpu_kind"] = kind
            line["c1"]["cpu_baseline"] = c1_cpu_numbers()
            if dwa is not None:
                line["dwa"]["cpu_baseline"] = dwa_cpu_numbers()
                line["dwa"]["cpu_baseline"].update(dwa_prepare_cpu_numbers())
                line["dwa"]["cpu_baseline"].update({"c5": fleet_cpu_numbers(), "c5_all_cores": fleet_cpu_parallel_numbers(),
                                                    "c4_all_cores": c4_all_cores_numbers()})
                ref_api = pyoracle.load(kind)
                line["trajectory_planner"]["cpu_baseline"] = dict(
                    tp_numbers(ref_api, inflate_local(ref_api, local_map_c2()), reps=5), kind=kind, cores=1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dwa", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
