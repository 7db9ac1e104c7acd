"""MapGrid::computeTargetDistance (base_local_planner/src/map_grid.cpp:174-310) on the GPU: every kernel variant --
the sliced shared-memory wavefront and the row kernel, with and without the helper warps a planner's launch adds --
must give the checker's four grids cell for cell, on odd map sizes, with unknown cells, with a
plan that leaves the map, and on a maze whose levels outgrow the eight bit planes held in registers.
"""
import os

import numpy as np
import pytest

import scenarios as sc

pytestmark = pytest.mark.gpu

VARIANTS = ["sliced", "rows", "rows+helpers"]
KEYS = ("NAVGPU_MAPGRID", "NAVGPU_MAPGRID_HELPERS")


@pytest.fixture(params=VARIANTS)
def variant(request):
    old = {k: os.environ.get(k) for k in KEYS}
    os.environ["NAVGPU_MAPGRID"] = request.param.split("+")[0]  # read by libnavgpu at every launch
    os.environ["NAVGPU_MAPGRID_HELPERS"] = str(int(request.param.endswith("helpers")))
    yield request.param
    for k, v in old.items():
        if v is None:
            del os.environ[k]
        else:
            os.environ[k] = v


def grids_of(api, grid, plan, pose, res=0.05, **over):
    sy, sx = grid.shape
    cfg = dict(vx_samples=3, vy_samples=1, vth_samples=5, max_vel_y=0.0, min_vel_y=0.0)
    cfg.update(over)
    d = api.dwa(sx, sy, res, **cfg)
    d.set_costmap(grid, 0.0, 0.0)
    d.set_plan(pose, plan)
    d.find_best_path(pose, (0.2, 0.0, 0.0), sc.PENTAGON)
    return [d.grid(k) for k in range(4)]


def random_map(rng, sx, sy, density):
    g = np.zeros((sy, sx), np.uint8)
    r = rng.random((sy, sx))
    g[r < density] = 254
    g[(r >= density) & (r < density * 1.3)] = 253
    g[(r >= density * 1.3) & (r < density * 1.6)] = 255
    g[(r >= density * 1.6) & (r < density * 2.5)] = rng.integers(1, 253)
    return g


@pytest.mark.parametrize("sx,sy", [(120, 120), (128, 128), (127, 113), (97, 64), (33, 128), (64, 31), (20, 7), (5, 5)])
@pytest.mark.parametrize("density", [0.0, 0.08, 0.3])
def test_grids_match_checker_on_random_maps(cuda, port, variant, sx, sy, density):
    rng = np.random.default_rng(sx * 1000 + sy + int(density * 100))
    g = random_map(rng, sx, sy, density)
    res = 0.05
    # a plan that starts off the map, wanders across it and may leave it again
    n = int(rng.integers(2, 60))
    t = np.linspace(0, 1, n)
    x0, y0 = rng.uniform(-0.2, sx * res * 0.3), rng.uniform(0, sy * res)
    x1, y1 = rng.uniform(sx * res * 0.5, sx * res * 1.2), rng.uniform(0, sy * res)
    plan = np.stack([x0 + (x1 - x0) * t, y0 + (y1 - y0) * t + 0.1 * np.sin(7 * t)], 1)
    pose = (float(np.clip(x0, 0.1, sx * res - 0.1)), float(np.clip(y0, 0.1, sy * res - 0.1)), 0.0)
    g[int(pose[1] / res), int(pose[0] / res)] = 0
    over = dict(allow_unknown=int(density > 0.2))
    a = grids_of(cuda, g, plan, pose, **over)
    b = grids_of(port, g, plan, pose, **over)
    for k in range(4):
        assert np.array_equal(a[k], b[k]), f"{variant}: grid {k} differs on {sx}x{sy}, density {density}"


def serpentine(sx, sy, gap=2):
    """Walls every `gap + 1` rows with alternating openings: the wavefront has to walk every corridor."""
    g = np.zeros((sy, sx), np.uint8)
    for i, r in enumerate(range(gap, sy, gap + 1)):
        g[r, :] = 254
        if i % 2:
            g[r, :2] = 0
        else:
            g[r, -2:] = 0
    return g


@pytest.mark.parametrize("sx,sy", [(120, 120), (128, 128), (126, 90)])
def test_maze_levels_beyond_the_register_planes(cuda, port, variant, sx, sy):
    g = serpentine(sx, sy)
    res = 0.05
    plan = np.array([[0.05, 0.03], [0.12, 0.03]])  # seeds in the first corridor only
    pose = (0.1, 0.03, 0.0)
    a = grids_of(cuda, g, plan, pose)
    b = grids_of(port, g, plan, pose)
    deepest = max(int(x[x < sx * sy].max(initial=0)) for x in b)
    assert deepest > 1000  # thousands of levels: far beyond 8 planes
    for k in range(4):
        assert np.array_equal(a[k], b[k]), f"{variant}: grid {k} differs in the maze"


def test_unknown_allowed_and_no_seed(cuda, port, variant):
    rng = np.random.default_rng(5)
    g = random_map(rng, 120, 120, 0.15)
    plan_in = np.stack([np.linspace(0.5, 5.5, 40), np.full(40, 3.0)], 1)
    plan_out = plan_in + 100.0  # nowhere near the map: every cell stays unreachable
    for plan in (plan_in, plan_out):
        a = grids_of(cuda, g, plan, (0.5, 3.0, 0.0))
        b = grids_of(port, g, plan, (0.5, 3.0, 0.0))
        for k in range(4):
            assert np.array_equal(a[k], b[k])


def test_dwa_scenarios_with_every_variant(cuda, port, variant):
    for seed in (3, 11, 27):
        x = sc.run_dwa_scenario(cuda, port, seed, cycles=2)
        y = sc.run_dwa_scenario(port, port, seed, cycles=2)
        for p, q in zip(x, y):
            assert sc.dwa_results_equal(p, q, rtol=1e-5)


def test_fleet_with_every_variant(cuda, port, variant):
    import test_gpu_fleet as tf
    tf.test_fleet_matches_checker_robot_by_robot(cuda, port, 5, tf.CFG)
