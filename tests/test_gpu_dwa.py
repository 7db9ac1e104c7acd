"""Path B parity: the CUDA DWA scorer (through the C ABI) against the CPU checker and the reference's goldens.

Tolerance (BASELINE.json north_star): trajectory costs within 1e-5 relative, identical best trajectory.  In practice
everything below is asserted bit-exact first (fp64 sums without FMA contraction, integer grids) and only falls back
to the 1e-5 comparison where CUDA's cos/sin (<= 2 ulp) could differ from glibc's in the last bit.
"""
import numpy as np
import pytest

import golden_util as gu
import scenarios as sc

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def assert_cycles_match(a, b):
    exact = 0
    for c, (x, y) in enumerate(zip(a, b)):
        assert sc.dwa_results_equal(x, y, rtol=RTOL), (
            f"cycle {c}: best {x['best_index']} vs {y['best_index']}, cost {x['cost']} vs {y['cost']}, "
            f"n {x['n_samples']}/{x['n_scored']} vs {y['n_samples']}/{y['n_scored']}, mask {x['mask']} vs {y['mask']}")
        exact += sc.dwa_results_equal(x, y, rtol=0.0)
    return exact


@pytest.mark.parametrize("seed", range(40))
def test_dwa_scenarios_match_checker(cuda, port, seed):
    """Multi-cycle findBestPath: 4 MapGrid wavefronts, sample enumeration, rollout, six critics, argmin, oscillation
    flags; random sample counts, holonomic / non-holonomic, use_dwa on/off, sum_scores on/off."""
    assert_cycles_match(sc.run_dwa_scenario(cuda, port, seed), sc.run_dwa_scenario(port, port, seed))


@pytest.mark.parametrize("seed,path", gu.dwa_cases())
def test_dwa_golden_fixtures(cuda, port, seed, path):
    gu.check_dwa_case(cuda, port, seed, path, rtol=RTOL)


def c2_setup(api, grid_api, **over):
    """Config C2: 6 m x 6 m rolling local costmap @0.05 (120x120), 20 x 1 x 20 samples, sim_time 1.7 s, pentagon
    footprint, corridor world inflated by the checker."""
    rng = np.random.default_rng(42)
    grid = sc.local_costmap(grid_api, rng, style="corridor")
    cfg = dict(vx_samples=20, vy_samples=1, vth_samples=20, max_vel_y=0.0, min_vel_y=0.0)
    cfg.update(over)
    d = api.dwa(120, 120, 0.05, **cfg)
    d.set_costmap(grid, 0.0, 0.0)
    pose, vel = (1.5, 3.0, 0.0), (0.3, 0.0, 0.0)
    plan = np.stack([np.arange(1.0, 7.0, 0.05), np.full(120, 3.0)], 1)
    d.set_plan(pose, plan)
    return d, pose, vel


def test_c2_bit_exact_selection_and_costs(cuda, port):
    a, pose, vel = c2_setup(cuda, port)
    b, _, _ = c2_setup(port, port)
    ra = a.find_best_path(pose, vel, sc.PENTAGON)
    rb = b.find_best_path(pose, vel, sc.PENTAGON)
    assert ra["n_samples"] == rb["n_samples"] and ra["n_samples"] in (400, 420)
    assert ra["best_index"] == rb["best_index"] and ra["ok"]
    assert np.allclose(ra["costs"], rb["costs"], rtol=RTOL, atol=0, equal_nan=True)
    assert ra["cost"] == pytest.approx(rb["cost"], rel=RTOL)
    assert (ra["xv"], ra["yv"], ra["thetav"]) == (rb["xv"], rb["yv"], rb["thetav"])
    assert np.allclose(ra["points"], rb["points"], rtol=RTOL, atol=1e-12)
    for k in range(4):
        assert np.array_equal(a.grid(k), b.grid(k))


def test_sharded_sweep_equals_single_search(cuda, port):
    """Sample-range sharding (config C4's multi-GPU scheme) on one GPU: per-shard (cost, index) minima combined with
    the lowest-index tie rule reproduce the unsharded search, whatever the shard boundaries."""
    over = dict(vx_samples=30, vy_samples=3, vth_samples=25, acc_lim_x=20.0, acc_lim_theta=20.0)
    a, pose, vel = c2_setup(cuda, port, **over)
    single = a.find_best_path(pose, vel, sc.PENTAGON, want_costs=False)
    b, _, _ = c2_setup(cuda, port, **over)
    n = single["n_samples"]
    for shards in (2, 3, 8):
        b.reset_oscillation()
        cuts = np.linspace(0, n, shards + 1).astype(np.int64)
        costs, idx = [], []
        for r in range(shards):
            c, i, total = b.score_range(pose, vel, sc.PENTAGON, int(cuts[r]), int(cuts[r + 1]))
            assert total == n
            costs.append(c)
            idx.append(i)
        res = b.finish_sharded(pose, costs, idx)
        assert res["best_index"] == single["best_index"] and res["cost"] == single["cost"]
        assert np.array_equal(res["points"], single["points"])
    ref, _, _ = c2_setup(port, port, **over)
    rr = ref.find_best_path(pose, vel, sc.PENTAGON)
    assert rr["best_index"] == single["best_index"] and rr["cost"] == pytest.approx(single["cost"], rel=RTOL)


def test_dense_sweep_reduced_c4(cuda, port):
    """Config C4 at reduced density (60 x 12 x 60 ~ 45 k samples incl. inserted zeros) against the checker."""
    over = dict(vx_samples=60, vy_samples=12, vth_samples=60, acc_lim_x=20.0, acc_lim_y=20.0, acc_lim_theta=20.0,
                max_vel_y=0.1, min_vel_y=-0.1)
    a, pose, vel = c2_setup(cuda, port, **over)
    b, _, _ = c2_setup(port, port, **over)
    ra = a.find_best_path(pose, vel, sc.PENTAGON)
    rb = b.find_best_path(pose, vel, sc.PENTAGON)
    assert ra["n_samples"] == rb["n_samples"] > 40000 and ra["n_scored"] == rb["n_scored"]
    assert ra["best_index"] == rb["best_index"]
    assert np.allclose(ra["costs"], rb["costs"], rtol=RTOL, atol=0, equal_nan=True)
    # how many per-sample costs are bit-identical (informational; cos/sin last-bit differences are the only source)
    same = np.sum((ra["costs"] == rb["costs"]) | (np.isnan(ra["costs"]) & np.isnan(rb["costs"])))
    assert same >= 0.999 * ra["n_samples"]


def test_no_valid_trajectory_keeps_stale_result(cuda, port):
    """Boxed in by lethal cells: cost -7, velocities/points keep the previous cycle's values (result_traj_ semantics)."""
    grid = np.zeros((120, 120), np.uint8)
    outs = []
    for api in (cuda, port):
        d = api.dwa(120, 120, 0.05, vx_samples=5, vy_samples=1, vth_samples=5, max_vel_y=0.0, min_vel_y=0.0)
        d.set_costmap(grid, 0.0, 0.0)
        pose, vel = (3.0, 3.0, 0.0), (0.2, 0.0, 0.0)
        plan = np.stack([np.arange(3.0, 5.5, 0.05), np.full(50, 3.0)], 1)
        d.set_plan(pose, plan)
        r1 = d.find_best_path(pose, vel, sc.PENTAGON)
        g2 = grid.copy()
        g2[50:70, 50:70] = 254
        d.set_costmap(g2, 0.0, 0.0)
        r2 = d.find_best_path(pose, vel, sc.PENTAGON)
        outs.append((r1, r2))
    (a1, a2), (b1, b2) = outs
    assert a1["ok"] and b1["ok"] and not a2["ok"] and not b2["ok"]
    assert a2["cost"] == b2["cost"] == -7.0
    assert (a2["xv"], a2["thetav"]) == (b2["xv"], b2["thetav"]) == (a1["xv"], a1["thetav"])
    assert np.allclose(a2["points"], b2["points"], rtol=RTOL, atol=1e-12) and len(a2["points"]) == len(a1["points"])
    assert np.array_equal(a2["costs"], b2["costs"], equal_nan=True)


@pytest.mark.parametrize("seed", range(8))
def test_check_trajectory_matches_checker(cuda, port, seed):
    """DWAPlanner::checkTrajectory (dwa_planner.cpp:213-237) through navgpu_dwa_check_trajectory: the cost of one
    given velocity sample against the critics' state of the last findBestPath, oscillation flags reset."""
    outs = []
    for api in (cuda, port):
        rng = np.random.default_rng(seed)
        s = sc.dwa_scenario(rng)
        grid = sc.local_costmap(port, np.random.default_rng(seed + 1000), ox=s["origin"][0], oy=s["origin"][1], style=s["style"])
        d = api.dwa(120, 120, 0.05, vx_samples=5, vy_samples=1, vth_samples=7)
        d.set_costmap(grid, *s["origin"])
        d.set_plan(s["pose"], s["plan"])
        d.find_best_path(s["pose"], s["vel"], sc.PENTAGON)
        costs = [d.check_trajectory(s["pose"], s["vel"], v, sc.PENTAGON)
                 for v in [(0.3, 0.0, 0.2), (0.55, 0.0, -1.0), (0.0, 0.0, 0.0), (0.05, 0.0, 0.1), (0.5, 0.1, 0.9), (2.0, 0.0, 0.0)]]
        outs.append((costs, d.oscillation_mask()))
    assert outs[0][1] == outs[1][1] == 0
    assert np.allclose(outs[0][0], outs[1][0], rtol=RTOL, atol=0), f"{outs}"


def test_c4_full_sweep_vs_checker(cuda, port):
    """Config C4 at BASELINE.json's full size: 200 x 20 x 200 samples (844 200 with the inserted zeros) scored against
    one local costmap; the checker needs ~10 s on one core.  Same winner, same sample counts, every reported cost
    within 1e-5 relative (and bit-identical for all but a handful)."""
    over = dict(vx_samples=200, vy_samples=20, vth_samples=200, acc_lim_x=20.0, acc_lim_y=20.0, acc_lim_theta=20.0,
                max_vel_y=0.1, min_vel_y=-0.1)
    a, pose, vel = c2_setup(cuda, port, **over)
    b, _, _ = c2_setup(port, port, **over)
    ra = a.find_best_path(pose, vel, sc.PENTAGON)
    rb = b.find_best_path(pose, vel, sc.PENTAGON)
    assert ra["n_samples"] == rb["n_samples"] == 844200 and ra["n_scored"] == rb["n_scored"]
    assert ra["best_index"] == rb["best_index"] and ra["cost"] == rb["cost"]
    assert (ra["xv"], ra["yv"], ra["thetav"]) == (rb["xv"], rb["yv"], rb["thetav"])
    assert np.allclose(ra["costs"], rb["costs"], rtol=RTOL, atol=0, equal_nan=True)
    same = np.sum((ra["costs"] == rb["costs"]) | (np.isnan(ra["costs"]) & np.isnan(rb["costs"])))
    assert same >= 0.9999 * ra["n_samples"], f"{ra['n_samples'] - same} costs differ in the last bits"
    assert np.array_equal(ra["points"], rb["points"])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_block_cyclic_shards_give_the_sequential_winner(cuda, port, world):
    """navgpu_dwa_score_strided: rank r scores the 8-sample blocks r, r + world, ...; the per-rank (cost, index) minima
    fed to navgpu_dwa_finish_sharded select the very sample (and cost, points, oscillation mask) the unsharded search
    selects -- on 40 401 samples, a count that is not a multiple of 8 x world."""
    over = dict(vx_samples=200, vy_samples=1, vth_samples=200, acc_lim_x=20.0, acc_lim_theta=20.0)
    whole, pose, vel = c2_setup(cuda, port, **over)
    rw = whole.find_best_path(pose, vel, sc.PENTAGON)
    d, _, _ = c2_setup(cuda, port, **over)
    minima = [d.score_strided(pose, vel, sc.PENTAGON, r, world) for r in range(world)]
    assert all(m[2] == rw["n_samples"] for m in minima)
    # every rank's minimum comes from one of its own blocks, and exactly one rank holds the overall winner
    assert all(i < 0 or (i // 8) % world == r for r, (c, i, _) in enumerate(minima))
    assert sum(i == rw["best_index"] for _, i, _ in minima) == 1
    rs = d.finish_sharded(pose, [m[0] for m in minima], [m[1] for m in minima])
    assert rs["best_index"] == rw["best_index"] and rs["cost"] == rw["cost"]
    assert (rs["xv"], rs["yv"], rs["thetav"]) == (rw["xv"], rw["yv"], rw["thetav"])
    assert np.array_equal(rs["points"], rw["points"])
    assert d.oscillation_mask() == whole.oscillation_mask()


@pytest.mark.parametrize("world", [2, 4])
def test_device_side_exchange_on_one_gpu(cuda, port, world):
    """navgpu_dwa_find_best_path_sharded with `world` planner handles on ONE device (navgpu_dwa_shard_connect_local):
    the last CTA of every rank's scoring kernel stores its record into every rank's exchange buffer and waits for the
    others; all ranks return the unsharded search's winner, n_scored and trajectory, over three cycles (oscillation
    state carried), without any host step between scoring and result."""
    over = dict(vx_samples=60, vy_samples=3, vth_samples=60, acc_lim_x=20.0, acc_lim_y=20.0, acc_lim_theta=20.0)
    whole, pose, vel = c2_setup(cuda, port, **over)
    ranks = [c2_setup(cuda, port, **over)[0] for _ in range(world)]
    cuda.shard_connect_local(ranks)
    pose = np.array(pose)
    vel = np.array(vel)
    for cyc in range(3):
        rw = whole.find_best_path(pose, vel, sc.PENTAGON, want_costs=False)
        for d in ranks:
            d.find_best_path_sharded_async(pose, vel, sc.PENTAGON)
        outs = [d.sharded_collect(pose) for d in ranks]
        for r, o in enumerate(outs):
            assert o["best_index"] == rw["best_index"] and o["cost"] == rw["cost"], f"cycle {cyc} rank {r}"
            assert o["n_samples"] == rw["n_samples"] and o["n_scored"] == rw["n_scored"]
            assert np.array_equal(o["points"], rw["points"])
        assert all(d.oscillation_mask() == whole.oscillation_mask() for d in ranks)
        vel = np.array([rw["xv"] * (-1 if cyc % 2 else 1), rw["yv"], -rw["thetav"]])
        pose = pose + np.array([0.01 * vel[0], 0, 0.01 * vel[2]])


def test_trajectory_cost_function_backend(cuda, port):
    """navgpu_dwa_prepare + navgpu_dwa_score_trajectories (the batched base_local_planner::TrajectoryCostFunction): the
    winner of a search, handed back as an explicit trajectory, scores exactly its search cost; the six terms add up in
    DWAPlanner's critic order; the rejecting codes of the reference's critics come out for off-map / lethal / empty
    inputs (the C++ adapter is checked against the reference's own critics in tests/cpp/dropin_harness.cpp section E)."""
    d, pose, vel = c2_setup(cuda, port)
    r = d.find_best_path(pose, vel, sc.PENTAGON)
    assert r["ok"]
    d2, _, _ = c2_setup(cuda, port)
    d2.prepare()
    win_v = (r["xv"], r["yv"], r["thetav"])
    off_map = np.array([[1.5, 3.0, 0.0], [6.5, 3.0, 0.0]])         # second point outside the 6 m map
    in_wall = np.array([[1.5, 3.0, 0.0], [1.5, 1.35, 0.0]])         # centre on the corridor wall (rows 26..28)
    empty = np.zeros((0, 3))
    costs, terms = d2.score_trajectories([r["points"], off_map, in_wall, empty, r["points"][:1]],
                                         [win_v, win_v, win_v, win_v, win_v], sc.PENTAGON, want_terms=True)
    assert costs[0] == r["cost"]
    assert terms[0, 0] == 0 and (terms[0] >= 0).all() and costs[0] == np.add.reduce(terms[0])  # left-to-right sum
    assert costs[1] == -6.0 and costs[2] == -6.0
    assert costs[3] == 0.0
    assert costs[4] >= 0
    # an empty footprint answers -9 (obstacle_cost_function.cpp:78-82); a point robot (< 3 vertices) looks at the centre
    assert d2.score_trajectories([r["points"]], [win_v], np.zeros((0, 2)))[0] == -9.0
    assert d2.score_trajectories([in_wall], [win_v], [(0.0, 0.0)])[0] == -6.0
    # latched oscillation flags reject by the sign of the trajectory's velocity (oscillation_cost_function.cpp:166-176)
    d2.update_oscillation(pose, 1.0, 0.3, 0.0, 0.0)    # forward, then backward: forward_neg_only latches (:101-164)
    d2.update_oscillation(pose, 1.0, -0.2, 0.0, 0.0)
    assert d2.oscillation_mask() == 2
    c = d2.score_trajectories([r["points"], r["points"]], [(0.3, 0.0, 0.0), (-0.3, 0.0, 0.0)], sc.PENTAGON)
    assert c[0] == -5.0 and c[1] == r["cost"]
    # 1000 trajectories in one launch: every sample's winner-style rescoring is independent of its batch position
    many = d2.score_trajectories([r["points"]] * 1000, [(-0.3, 0.0, 0.0)] * 1000, sc.PENTAGON)
    assert (many == r["cost"]).all()


def test_all_explored_velocities_available(cuda, port):
    """navgpu_dwa_get_samples: the per-axis samples behind all_explored's xv_/yv_/thetav_; the winner's velocities are
    the samples its index names."""
    d, pose, vel = c2_setup(cuda, port, vy_samples=3, max_vel_y=0.1, min_vel_y=-0.1)
    r = d.find_best_path(pose, vel, sc.PENTAGON)
    xs, ys, ths = d.samples()
    assert len(xs) * len(ys) * len(ths) == r["n_samples"]
    i = r["best_index"]
    ith, iy, ix = i % len(ths), (i // len(ths)) % len(ys), i // (len(ths) * len(ys))
    assert (float(xs[ix]), float(ys[iy]), float(ths[ith])) == (r["xv"], r["yv"], r["thetav"])
