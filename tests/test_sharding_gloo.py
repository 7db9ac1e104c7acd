"""The N > 1 host logic on CPU: world_size 2 over gloo (127.0.0.1).

Sample-range sharding of a DWA sweep: each rank takes its contiguous slice of the checker's per-sample costs, reduces
it to one (cost, global index) minimum, the minima are all-gathered and every rank picks the winner with the
reference's first-strictly-smaller rule -- which must be the checker's own best sample, including on exact ties.
Fleet partitioning: the ranks' robot ranges tile [0, n) without overlap."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenarios as sc
from navigation_b200 import sharding


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, costs, expect, out, strided):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if strided:
            c, i = sharding.local_minimum_strided(costs, rank, world)
        else:
            lo, hi = sharding.split_range(len(costs), rank, world)
            c, i = sharding.local_minimum(costs[lo:hi], lo)
        cs, idx = sharding.allgather_minima(dist, torch, c, i, "cpu")
        out[rank] = sharding.pick_winner(cs, idx)
        assert out[rank][1] == expect
    finally:
        dist.destroy_process_group()


def run_world(costs, expect, world=2, strided=False):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, free_port(), costs, expect, out, strided), nprocs=world, join=True)
    assert len(out) == world and len({v for v in out.values()}) == 1  # every rank agrees
    return out[0]


def test_sharded_argmin_equals_sequential_search(port):
    """Costs from the checker's findBestPath on a C2-style scenario (400-420 samples)."""
    rng = np.random.default_rng(5)
    grid = sc.local_costmap(port, rng, style="corridor")
    d = port.dwa(120, 120, 0.05, vx_samples=20, vy_samples=1, vth_samples=20, max_vel_y=0.0, min_vel_y=0.0)
    d.set_costmap(grid, 0.0, 0.0)
    pose, vel = (1.5, 3.0, 0.0), (0.3, 0.0, 0.0)
    d.set_plan(pose, np.stack([np.arange(1.0, 7.0, 0.05), np.full(120, 3.0)], 1))
    r = d.find_best_path(pose, vel, sc.PENTAGON)
    # the reported costs of losers may be early-exit partial sums (> best), never below the winner's
    cost, index = run_world(np.asarray(r["costs"]), r["best_index"])
    assert index == r["best_index"] and cost == r["cost"]
    # the block-cyclic partition (blocks of 8 samples dealt round-robin) selects the same sample
    cost, index = run_world(np.asarray(r["costs"]), r["best_index"], strided=True)
    assert index == r["best_index"] and cost == r["cost"]
    cost, index = run_world(np.asarray(r["costs"]), r["best_index"], world=3, strided=True)
    assert index == r["best_index"] and cost == r["cost"]


def test_sharded_argmin_ties_resolve_to_lowest_index():
    costs = np.array([5.0, np.nan, 3.0, -6.0, 3.0, 7.0, 3.0, np.nan])  # the tie spans both ranks' slices
    assert run_world(costs, 2)[1] == 2
    costs = np.array([np.nan, -6.0, -2.0, np.nan, 9.0, 4.0, 4.0, -5.0])  # rank 0 has nothing valid
    assert run_world(costs, 5) == (4.0, 5)
    costs = np.array([np.nan, -6.0, -2.0, np.nan])  # nobody has anything valid
    assert run_world(costs, -1)[1] == -1


def test_strided_ties_resolve_to_lowest_index():
    costs = np.full(40, 9.0)
    costs[[3, 11, 19, 35]] = 2.0  # equal minima in blocks 0, 1, 2, 4: ranks 0, 1, 0, 0 of a world of 2
    assert run_world(costs, 3, strided=True) == (2.0, 3)
    costs[3] = np.nan
    assert run_world(costs, 11, strided=True) == (2.0, 11)


@pytest.mark.parametrize("n,world", [(844200, 8), (40401, 3), (420, 2), (5, 4)])
def test_strided_shares_tile_without_overlap(n, world):
    shares = [sharding.strided_indices(n, r, world) for r in range(world)]
    assert sum(len(s) for s in shares) == n and len(np.unique(np.concatenate(shares))) == n
    blocks = -(-n // 8)
    for r, s in enumerate(shares):  # what launch_score computes: max(1, ceil((blocks - r) / world)) CTAs of 8 samples
        assert -(-len(s) // 8) == max(0, -(-(blocks - r) // world))


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 2), (4096, 8), (10, 4), (3, 8)])
def test_ranges_tile_without_overlap(n, world):
    r = [sharding.split_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
