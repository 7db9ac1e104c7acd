"""Host-side partitioning of the two workloads that shard; torch.distributed is plumbing only.

* Dense rollout sweep (config C4): the enumerated velocity samples (x outer, y, theta inner -- the reference's order and
  therefore its tie-break order, simple_trajectory_generator.cpp:121-133) are split into contiguous index ranges, one
  per rank; every rank scores its range on its GPU and contributes one 16-byte (cost, global index) minimum; after one
  all-gather every rank applies the same rule the sequential search applies (first strictly smaller cost ==
  smallest cost, lowest index on ties, simple_scored_sampling_planner.cpp:111-116).
* Fleet (config C5): robots are independent; rank r owns robots [n*r/G, n*(r+1)/G).  No collective on the data path.
"""
import numpy as np

SAMPLE_BLOCK = 8  # samples per CTA of k_dwa_score: the unit the block-cyclic partition deals out


def strided_indices(total, rank, world, block=SAMPLE_BLOCK):
    """Global sample indices of `rank` under the block-cyclic partition of navgpu_dwa_score_strided /
    navgpu_dwa_find_best_path_sharded: blocks rank, rank + world, rank + 2 world, ... of `block` samples each.
    Trajectory length grows with the outer (vx) sample index, so contiguous ranges are unbalanced (the last of 8
    ranks gets 1.6x the mean number of trajectory points); dealing blocks round-robin evens that out, and the
    winner rule -- smallest cost, lowest index on ties -- does not depend on the partition."""
    idx = np.arange(total, dtype=np.int64)
    return idx[(idx // block) % world == rank]


def local_minimum_strided(costs, rank, world, block=SAMPLE_BLOCK):
    """(cost, global index) of the first strictly smallest valid cost among this rank's block-cyclic share."""
    c = np.asarray(costs, dtype=np.float64)
    mine = strided_indices(len(c), rank, world, block)
    if mine.size == 0:
        return float("inf"), -1
    cost, pos = local_minimum(c[mine])
    return cost, (int(mine[pos]) if pos >= 0 else -1)


def connect_shards(dist, dwa):
    """Device-side exchange setup for one planner handle per rank (one process per GPU): every rank exports its
    exchange buffer as a CUDA IPC handle, the 64-byte handles travel once over torch.distributed, every rank maps its
    peers' buffers.  After this the per-cycle data path has no collective call: the scoring kernels exchange their
    16-byte minima themselves (navgpu_dwa_find_best_path_sharded)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    handles = [None] * world
    dist.all_gather_object(handles, dwa.shard_export())
    dwa.shard_connect(rank, world, handles)
    dist.barrier()


def split_range(total, rank, world):
    """Contiguous share [lo, hi) of `total` items for `rank` of `world`."""
    return (total * rank) // world, (total * (rank + 1)) // world


def local_minimum(costs, begin=0):
    """(cost, global index) of the first strictly smallest valid cost of a slice; (inf, -1) when none is valid.
    `costs` follows the C ABI's all_costs convention: NaN = not generated, negative = rejected."""
    c = np.asarray(costs, dtype=np.float64)
    valid = ~np.isnan(c) & (c >= 0)
    if not valid.any():
        return float("inf"), -1
    masked = np.where(valid, c, np.inf)
    i = int(np.argmin(masked))  # argmin returns the first of equal minima
    return float(masked[i]), begin + i


def pick_winner(costs, indices):
    """Winner among per-rank minima: smallest cost, lowest sample index on ties; (inf, -1) when every rank has none.
    Identical to navgpu_dwa_finish_sharded's host rule."""
    best_c, best_i = float("inf"), -1
    for c, i in zip(costs, indices):
        i = int(i)
        if i >= 0 and (best_i < 0 or c < best_c or (c == best_c and i < best_i)):
            best_c, best_i = float(c), i
    return best_c, best_i


_BUFFERS = {}


def allgather_minima(dist, torch, cost, index, device):
    """All-gather of one (cost, index) pair per rank.  The index travels as fp64 (exact below 2^53).  The send / receive
    tensors are kept across calls (a sweep is a few hundred microseconds: allocations would show)."""
    world = dist.get_world_size()
    key = (str(device), world)
    if key not in _BUFFERS:
        _BUFFERS[key] = (torch.zeros(2, dtype=torch.float64).pin_memory() if str(device) != "cpu"
                         else torch.zeros(2, dtype=torch.float64),
                         torch.zeros(2, dtype=torch.float64, device=device),
                         torch.zeros(2 * world, dtype=torch.float64, device=device))
    host, buf, out = _BUFFERS[key]
    host[0] = cost
    host[1] = float(index)
    buf.copy_(host, non_blocking=True)
    dist.all_gather_into_tensor(out, buf)
    g = out.cpu().numpy().reshape(world, 2)
    return g[:, 0].copy(), g[:, 1].astype(np.int64)
