"""Loading / comparing the golden fixtures produced by tests/golden/make_golden.py (reference outputs)."""
import glob
import os
import re

import numpy as np

import scenarios as sc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def costmap_cases():
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "costmap_*.npz"))):
        m = re.match(r"costmap_(tf|any)_(\d+)\.npz", os.path.basename(p))
        out.append((m.group(1) == "tf", int(m.group(2)), p))
    return out


def dwa_cases():
    return [(int(re.match(r"dwa_(\d+)\.npz", os.path.basename(p)).group(1)), p)
            for p in sorted(glob.glob(os.path.join(GOLDEN, "dwa_*.npz")))]


def tp_cases():
    return [(int(re.match(r"tp_(\d+)\.npz", os.path.basename(p)).group(1)), p)
            for p in sorted(glob.glob(os.path.join(GOLDEN, "tp_*.npz")))]


TP_BOXED = 9000  # the "seed" of sc.run_tp_boxed_scenario among the tp_*.npz fixtures


def run_tp_case(api, grid_api, seed):
    return sc.run_tp_boxed_scenario(api) if seed == TP_BOXED else sc.run_tp_scenario(api, grid_api, seed)


def load_tp_case(path):
    """The reference TrajectoryPlanner's outputs for sc.run_tp_scenario(seed): a list of per-cycle result dicts."""
    g = np.load(path)
    out = []
    c = 0
    while f"scalars{c}" in g:
        s = g[f"scalars{c}"]
        out.append(dict(cost=s[0], xv=s[1], yv=s[2], thetav=s[3], flags=int(s[4]), points=g[f"points{c}"],
                        scores=g[f"scores{c}"], grids=[g[f"grid{c}_{k}"].astype(np.float64) for k in range(2)]))
        c += 1
    return out


def check_costmap_case(api, tie_free, seed, path, exact=None, inflation=None, masks=None, one_sided=None):
    """Runs the scenario of a golden fixture and compares it with the compiled reference's recorded output: windows,
    origins and obstacle layer grids always ==; the master grid == everywhere (default), == outside the per-cycle
    `masks` (a scenario's tie-variant mask), or cell by cell through `one_sided(got, reference)`.  Returns the number
    of differing master cells."""
    g = np.load(path)
    tr = sc.run_costmap_scenario(api, seed, tie_free=tie_free, inflation=inflation)
    bad = 0
    for c, (w, m, o, org) in enumerate(tr):
        assert tuple(g[f"w{c}"]) == tuple(w), f"window differs in cycle {c}"
        assert tuple(g[f"org{c}"]) == tuple(org), f"origin differs in cycle {c}"
        assert np.array_equal(g[f"o{c}"], o), f"obstacle layer grid differs in cycle {c}"
        diff = g[f"m{c}"] != m
        bad += int(diff.sum())
        if masks is not None:
            assert not (diff & ~masks[c]).any(), f"cycle {c}: {int((diff & ~masks[c]).sum())} cells differ outside the mask"
        elif one_sided is not None:
            assert one_sided(m, g[f"m{c}"]), f"cycle {c}: a cell is lower than the reference"
        else:
            assert not diff.any(), f"cycle {c}: {int(diff.sum())} master cells differ from the reference"
    return bad


def check_dwa_case(api, grid_api, seed, path, rtol=0.0):
    g = np.load(path)
    out = sc.run_dwa_scenario(api, grid_api, seed)
    for c, r in enumerate(out):
        s = g[f"scalars{c}"]
        ref = dict(ok=bool(s[0]), cost=s[1], xv=s[2], yv=s[3], thetav=s[4], best_index=int(s[5]), n_samples=int(s[6]),
                   n_scored=int(s[7]), mask=int(s[8]), costs=g[f"costs{c}"], points=g[f"points{c}"],
                   grids=[g[f"grid{c}_{k}"].astype(np.float64) for k in range(4)])
        assert sc.dwa_results_equal(r, ref, rtol=rtol), f"seed {seed} cycle {c} differs from the reference"
