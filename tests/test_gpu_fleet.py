"""Config C5 parity: the batched fleet path (navgpu_fleet_*) against the CPU checker run robot by robot.

Per robot the checker does what the reference does for one robot: LayeredCostmap (static-style layer + InflationLayer)
on the raw local map, then DWAPlanner::findBestPath on the inflated map.  Asserted per robot and cycle: inflated
costmap bit-exact; best sample index, sample counts and oscillation flags identical; cost within 1e-5 relative (in
practice bit-equal); velocities identical."""
import numpy as np
import pytest

import scenarios as sc
from navigation_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5
CFG = dict(vx_samples=6, vy_samples=1, vth_samples=11, max_vel_y=0.0, min_vel_y=0.0)


def checker_robot(port, robot, cfg):
    cm = port.costmap(120, 120, 0.05, *robot["origin"])
    s = cm.add_grid_layer(0)
    cm.add_inflation_layer(0.55, 10.0)
    cm.set_footprint(sc.PENTAGON)
    cm.set_grid_layer(s, robot["raw"])
    cm.update_map(0, 0, 0)
    grid = cm.get()
    d = port.dwa(120, 120, 0.05, **cfg)
    d.set_costmap(grid, *robot["origin"])
    return grid, d


@pytest.mark.parametrize("n_robots,cfg", [(24, CFG), (7, dict(vx_samples=20, vy_samples=1, vth_samples=20, max_vel_y=0.0, min_vel_y=0.0)),
                                          (5, dict(vx_samples=4, vy_samples=3, vth_samples=9))])
def test_fleet_matches_checker_robot_by_robot(cuda, port, n_robots, cfg):
    robots = [synth.fleet_robot(100 + i) for i in range(n_robots)]
    fleet = cuda.fleet(n_robots, 120, 120, 0.05, sc.PENTAGON, 0.55, 10.0, **cfg)
    fleet.set_maps(np.stack([r["raw"] for r in robots]), np.array([r["origin"] for r in robots]))
    refs = [checker_robot(port, r, cfg) for r in robots]
    poses = np.array([r["pose"] for r in robots])
    vels = np.array([r["vel"] for r in robots])
    exact = total = 0
    for cycle in range(4):
        fleet.set_plans(poses, [r["plan"] for r in robots])
        out = fleet.step(poses, vels)
        for i, (grid, d) in enumerate(refs):
            if cycle == 0:
                assert np.array_equal(fleet.costmap(i), grid), f"robot {i}: inflated local costmap differs"
            d.set_plan(poses[i], robots[i]["plan"])
            e = d.find_best_path(poses[i], vels[i], sc.PENTAGON, want_costs=False) if hasattr(d, "stream") else \
                d.find_best_path(poses[i], vels[i], sc.PENTAGON)
            g = out[i]
            assert (g["best_index"], g["n_samples"], g["n_scored"]) == (e["best_index"], e["n_samples"], e["n_scored"]), \
                f"cycle {cycle} robot {i}: {g} vs best {e['best_index']} cost {e['cost']}"
            assert fleet.oscillation_mask(i) == d.oscillation_mask()
            if e["ok"]:
                assert np.isclose(g["cost"], e["cost"], rtol=RTOL, atol=0)
                assert (g["xv"], g["yv"], g["thetav"]) == (e["xv"], e["yv"], e["thetav"])
                exact += g["cost"] == e["cost"]
                total += 1
            else:
                assert g["cost"] < 0
        # alternate the commanded direction so oscillation flags latch on some robots
        for i, e in enumerate(out):
            if e["cost"] >= 0:
                vels[i] = [e["xv"] * (-1 if cycle % 2 else 1), e["yv"], -e["thetav"]]
            poses[i] = poses[i] + np.array([0.01 * vels[i][0], 0.0, 0.01 * vels[i][2]])
    assert total > 0 and exact >= 0.99 * total  # costs are bit-equal in practice


def test_fleet_inflation_does_not_leak_between_robots(cuda):
    """A lethal wall on the last rows of robot 0 and the first rows of robot 2 must not inflate into robot 1."""
    raw = np.zeros((3, 120, 120), np.uint8)
    raw[0, 117:120, :] = 254
    raw[2, 0:3, :] = 254
    fleet = cuda.fleet(3, 120, 120, 0.05, sc.PENTAGON, 1.0, 10.0, **CFG)
    fleet.set_maps(raw, np.zeros((3, 2)))
    fleet.set_plans(np.tile([1.0, 3.0, 0.0], (3, 1)), [np.array([[1.0, 3.0], [5.0, 3.0]])] * 3)
    fleet.step(np.tile([1.0, 3.0, 0.0], (3, 1)), np.tile([0.2, 0.0, 0.0], (3, 1)))
    assert fleet.costmap(1).max() == 0
    assert fleet.costmap(0)[100, 5] > 0 and fleet.costmap(2)[19, 5] > 0


def test_c5_full_fleet_sampled_parity(cuda, port):
    """The benched configuration itself: 4096 robots with bench.py's inputs (synth.fleet_robot(i), 20 x 1 x 20
    samples), one navgpu_fleet_step; every 64th robot is replayed through the checker the way the reference would run
    it (LayeredCostmap inflation, then DWAPlanner::findBestPath): inflated costmap ==, best sample index, sample counts,
    oscillation mask ==, cost <= 1e-5 relative, velocities ==.  Robots the fleet reports without a valid trajectory
    must have none in the checker either (the count bench.py prints as c5_valid_robots)."""
    cfg = dict(vx_samples=20, vy_samples=1, vth_samples=20, max_vel_y=0.0, min_vel_y=0.0)
    n = 4096
    robots = [synth.fleet_robot(i) for i in range(n)]
    fleet = cuda.fleet(n, 120, 120, 0.05, sc.PENTAGON, 0.55, 10.0, **cfg)
    fleet.set_maps(np.stack([r["raw"] for r in robots]), np.array([r["origin"] for r in robots]))
    poses = np.array([r["pose"] for r in robots])
    vels = np.array([r["vel"] for r in robots])
    fleet.set_plans(poses, [r["plan"] for r in robots])
    out = fleet.step(poses, vels)
    invalid = [i for i, g in enumerate(out) if g["cost"] < 0]
    sample = sorted(set(range(0, n, 64)) | set(invalid[:24]))
    exact = valid = 0
    for i in sample:
        grid, d = checker_robot(port, robots[i], cfg)
        assert np.array_equal(fleet.costmap(i), grid), f"robot {i}: inflated local costmap differs"
        d.set_plan(poses[i], robots[i]["plan"])
        e = d.find_best_path(poses[i], vels[i], sc.PENTAGON)
        g = out[i]
        assert (g["best_index"], g["n_samples"], g["n_scored"]) == (e["best_index"], e["n_samples"], e["n_scored"]), \
            f"robot {i}: {g} vs best {e['best_index']} cost {e['cost']}"
        assert fleet.oscillation_mask(i) == d.oscillation_mask()
        assert (g["cost"] >= 0) == bool(e["ok"]), f"robot {i}: fleet cost {g['cost']}, checker cost {e['cost']}"
        if e["ok"]:
            assert np.isclose(g["cost"], e["cost"], rtol=RTOL, atol=0)
            assert (g["xv"], g["yv"], g["thetav"]) == (e["xv"], e["yv"], e["thetav"])
            exact += g["cost"] == e["cost"]
            valid += 1
    assert valid >= 60 and exact >= 0.99 * valid
    print(f"C5 full fleet: {n - len(invalid)} of {n} robots have a valid trajectory; {len(sample)} replayed through the "
          f"checker ({min(24, len(invalid))} of them without one), {exact}/{valid} costs bit-equal")
