"""Small end-to-end run of both paths, sized for a run under compute-sanitizer where that tool is available (it is closed
on the gpurun pool; there the script serves as a quick whole-surface exercise of the C ABI): a 400 x 300 layered costmap
with obstacle layer over several cycles incl. the early (whole-map) mode and the host mirror, the plugin seam, a DWA
search, the batched TrajectoryCostFunction, a 2-handle sharded sweep, a small fleet, the legacy planner and the plan
preprocessing.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import navigation_b200  # noqa: E402
from navigation_b200 import synth  # noqa: E402

api = navigation_b200.load()
size = 400
static, obs, robot, fp = synth.warehouse_c3(size=size, n_obs=2)
cm = api.costmap(size, 300, 0.05)
s = cm.add_grid_layer(0)
o = cm.add_obstacle_layer(1, True, 2.0)
il = cm.add_inflation_layer(0.55, 10.0)
cm.set_footprint(fp)
cm.set_grid_layer(s, static[:300])
mirror = np.zeros((300, size), np.uint8)
for cyc in range(4):
    cm.set_observations(o, obs if cyc % 2 == 0 else obs[:1])
    if cyc != 2:
        cm.touch_grid_layer(s, 0, 0, size, 300)   # whole-map cycles: early mode + tile hand-over
    cm.update_map_async(*robot)
    cm.get_changed(mirror)
    assert np.array_equal(mirror, cm.get())
cm.set_inflation_mode(il, 1)
cm.touch_grid_layer(s, 0, 0, size, 300)
cm.update_map(*robot)
R, costs, _ = api.build_cost_table(0.05, 0.325, 0.55, 10.0)
api.inflate_host(static[:300].copy(), 0, 0, size, 300, costs, R)

import scenarios as sc  # noqa: E402
from oracle import pyoracle  # noqa: E402  (only used to inflate the local map of the DWA scenario)
port = pyoracle.load("port")
rng = np.random.default_rng(42)
grid = sc.local_costmap(port, rng, style="corridor")
plan = np.stack([np.arange(1.0, 7.0, 0.05), np.full(120, 3.0)], 1)
pose, vel = (1.5, 3.0, 0.0), (0.3, 0.0, 0.0)
d = api.dwa(120, 120, 0.05, vx_samples=20, vy_samples=3, vth_samples=20)
d.set_costmap(grid, 0.0, 0.0)
d.set_plan(pose, plan)
r = d.find_best_path(pose, vel, sc.PENTAGON)
d.prepare()
d.score_trajectories([r["points"], r["points"][:3]], [(r["xv"], r["yv"], r["thetav"])] * 2, sc.PENTAGON)
ranks = []
for _ in range(2):
    dd = api.dwa(120, 120, 0.05, vx_samples=40, vy_samples=3, vth_samples=40, acc_lim_x=20.0, acc_lim_theta=20.0)
    dd.set_costmap(grid, 0.0, 0.0)
    dd.set_plan(pose, plan)
    ranks.append(dd)
api.shard_connect_local(ranks)
for dd in ranks:
    dd.find_best_path_sharded_async(pose, vel, sc.PENTAGON)
outs = [dd.sharded_collect(pose) for dd in ranks]
assert outs[0]["best_index"] == outs[1]["best_index"]
robots = [synth.fleet_robot(i) for i in range(6)]
fleet = api.fleet(6, 120, 120, 0.05, sc.PENTAGON, 0.55, 10.0, vx_samples=6, vy_samples=1, vth_samples=11, max_vel_y=0.0,
                  min_vel_y=0.0)
fleet.set_maps(np.stack([q["raw"] for q in robots]), np.array([q["origin"] for q in robots]))
fleet.set_plans(np.array([q["pose"] for q in robots]), [q["plan"] for q in robots])
fleet.step(np.array([q["pose"] for q in robots]), np.array([q["vel"] for q in robots]))
tp = api.trajectory_planner(120, 120, 0.05, sc.PENTAGON)
tp.set_costmap(grid, 0.0, 0.0)
tp.update_plan(plan)
tp.find_best_path(pose, vel)
api.plans_transform([np.c_[plan, np.zeros(len(plan))]], [(1.5, 3.0)], [[1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]], [2.0])
print("sanitize_small ok")
