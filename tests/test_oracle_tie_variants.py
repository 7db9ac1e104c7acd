"""What the reference's InflationLayer::updateCosts (inflation_layer.cpp:172-293) leaves to libstdc++'s heap, pinned
on the CPU: the checker's tie-policy variants (oracle_api.h navo_inflation_set_variant), the tie-variant mask they
define, and the two specifications libnavgpu's inflation modes are tested against bit for bit on the GPU.

  * variant 0 (std::priority_queue, same push order) IS the compiled reference -- bit for bit on every world here;
  * variants 1-3 (FIFO / LIFO / seeded random order among equal distances) are the same loop under other legal heaps;
    all of them coincide with the reference on axis-aligned thick worlds (the tie-free class, every gate config);
  * variant 5 (level-synchronous; = libnavgpu mode 1) is certified as a legal execution by variant 6, the sequential
    priority-queue loop with ties resolved towards variant 5's sources, and equals the reference outside the mask;
  * variant 4 (exact nearest-seed; = libnavgpu mode 0) is never lower than the reference, and differs from it also
    OUTSIDE the mask where a nearer source is blocked (the pinned three-seed case) -- a tie-independent deviation.
"""
import numpy as np
import pytest

import scenarios as sc
from test_gpu_costmap import never_lower, world_from_kind


def inflate(api, g, radius, which, seed=0, scaling=10.0, half=0.325):
    n = g.shape[0]
    cm = api.costmap(g.shape[1], n, 0.05)
    s = cm.add_grid_layer(0)
    il = cm.add_inflation_layer(radius, scaling)
    cm.set_footprint(sc.square_footprint(half))
    cm.set_grid_layer(s, g)
    sc.select_inflation(cm, il, which, seed)
    cm.update_map()
    return cm.get(), (cm.inflation_last_rounds(il) if which == "propagate" else 0)


@pytest.mark.parametrize("radius", [0.55, 1.0])
@pytest.mark.parametrize("kind", ["axis1", "ring", "salt", "diag3", "diag1"])
def test_tie_worlds(port, ref, kind, radius):
    g = world_from_kind(kind, 400, 1)
    r, _ = inflate(ref, g, radius, None)
    p, _ = inflate(port, g, radius, "reference")
    assert np.array_equal(r, p), "the restatement's heap order differs from the compiled reference"
    mask = np.zeros_like(g, bool)
    for which, seed in sc.TIE_POLICIES:
        mask |= inflate(port, g, radius, which, seed)[0] != r
    prop, rounds = inflate(port, g, radius, "propagate")
    cert, _ = inflate(port, g, radius, "certificate")
    exact, _ = inflate(port, g, radius, "exact")
    assert np.array_equal(prop, cert), "level-synchronous propagation is not reproduced by the sequential loop"
    assert not ((prop != r) & ~mask).any(), "propagation differs from the reference outside the tie-variant mask"
    assert never_lower(exact, r)
    assert ((prop == exact) | (prop < exact) | (exact == 255)).all()
    print(f"{kind} R={radius}: mask {int(mask.sum())} cells ({mask.mean():.1e}); propagate != reference on "
          f"{int((prop != r).sum())}; exact != reference on {int((exact != r).sum())}, "
          f"{int(((exact != r) & ~mask).sum())} of them outside the mask; {rounds} rounds")


@pytest.mark.parametrize("seed", range(0, 30))
def test_adversarial_scenarios(port, seed):
    """The multi-cycle scenarios of the GPU parity tests (rolling windows, merges, stateful windowed inflation)."""
    base, masks = sc.tie_mask_trace(port, seed)
    prop = sc.run_costmap_scenario(port, seed, inflation="propagate")
    cert = sc.run_costmap_scenario(port, seed, inflation="certificate")
    exact = sc.run_costmap_scenario(port, seed, inflation="exact")
    for c, (b, m, p, q, e) in enumerate(zip(base, masks, prop, cert, exact)):
        assert np.array_equal(p[1], q[1]), f"cycle {c}: certificate differs"
        assert not ((p[1] != b[1]) & ~m).any(), f"cycle {c}: propagation differs outside the mask"
        assert never_lower(e[1], b[1]), f"cycle {c}: exact inflation is lower than the reference"
        assert p[0] == b[0] == e[0] and p[3] == b[3] == e[3]


@pytest.mark.parametrize("seed", range(100, 112))
def test_tie_free_scenarios_have_an_empty_mask(port, seed):
    base, masks = sc.tie_mask_trace(port, seed, tie_free=True)
    assert not any(m.any() for m in masks)
    for which in ("exact", "propagate"):
        tr = sc.run_costmap_scenario(port, seed, tie_free=True, inflation=which)
        assert all(np.array_equal(a[1], b[1]) for a, b in zip(tr, base)), which


def test_blocked_propagation_three_seeds(port, ref):
    """seeds at (3,3), (4,1), (0,5), R = 10, scaling 3: the reference writes 158 at (-1,0), the exact
    nearest-seed value is 160, and no tie policy changes that."""
    g = np.zeros((60, 60), np.uint8)
    ox, oy = 21, 20
    for dx, dy in ((3, 3), (4, 1), (0, 5)):
        g[oy + dy, ox + dx] = 254
    r, _ = inflate(ref, g, 0.5, None, scaling=3.0, half=0.1)
    assert r[oy, ox - 1] == 158
    for which, seed in [("reference", 0), ("propagate", 0), ("certificate", 0)] + sc.TIE_POLICIES:
        assert np.array_equal(inflate(port, g, 0.5, which, seed, scaling=3.0, half=0.1)[0], r), which
    e, _ = inflate(port, g, 0.5, "exact", scaling=3.0, half=0.1)
    assert e[oy, ox - 1] == 160 and int((e != r).sum()) == 1


def test_reference_library_has_only_its_own_order(ref):
    cm = ref.costmap(10, 10, 1.0)
    il = cm.add_inflation_layer(1.0, 1.0)
    cm.set_inflation_variant(il, 0)
    with pytest.raises(ValueError):
        cm.set_inflation_variant(il, 1)
