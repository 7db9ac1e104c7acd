"""Measurement helper: per-CTA timeline of k_merge_seed / k_inflate in one C3 cycle (needs NAVGPU_TRACE=1).
Prints how many CTAs are resident over time, the CTA duration distribution and per-SM idle time; the raw records go
to gpurun_out/cta_trace.npz."""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("NAVGPU_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import navigation_b200  # noqa: E402

api = navigation_b200.load()
size = int(os.environ.get("PROBE_SIZE", 4000))
cm, (s, o, il), sets = bench.build_c3(lambda *a: api.costmap(*a), size=size)
obs, robot = sets[0]
cm.set_observations(o, obs if os.environ.get("PROBE_NO_OBS") is None else [])
stream = torch.cuda.ExternalStream(cm.stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush64 = flush.view(torch.int64)
n_merge = ((size + 255) // 256) * ((size + 63) // 64)
n_infl = ((size + 63) // 64) * ((size + 127) // 128)
n_merge, n_infl = min(n_merge, 4096), min(n_infl, 4096)
for k in range(8):
    if os.environ.get("PROBE_FLUSH", "1") in ("1", "2"):
        with torch.cuda.stream(stream):
            flush.zero_()
            if os.environ.get("PROBE_FLUSH") == "2":  # a read pass behind the write: the L2 ends up full of CLEAN lines
                flush64.sum()
    cm.touch_grid_layer(s, 0, 0, size, size)
    cm.update_map_async(*robot)
    t = cm.last_trace().astype(np.int64)
m = cm.last_cta_trace(0, n_merge).astype(np.int64)
f = cm.last_cta_trace(1, n_infl).astype(np.int64)
t0 = min(t[0], t[2])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cta_trace.npz"), merge=m, inflate=f, summary=t)
print("summary (us): obstacle %.1f..%.1f merge %.1f..%.1f inflate %.1f..%.1f" %
      tuple((int(v) - t0) / 1e3 for v in t[:6]))
print("obstacle kernel (us): first CTA done with its rays %.1f, last %.1f, marks stored %.1f, polygon cleared %.1f, end %.1f"
      % tuple((int(v) - t0) / 1e3 for v in (t[14], t[9], t[10], t[11], t[1])))


def report(name, r):
    r = r[r[:, 0] > 0]
    st, wk, en, sm = (r[:, 0] - t0) / 1e3, (r[:, 1] - t0) / 1e3, (r[:, 2] - t0) / 1e3, r[:, 3]
    ok = en > 0
    print("%s: %d CTAs, %d SMs; start %.1f..%.1f us, end %.1f..%.1f us" %
          (name, len(r), len(np.unique(sm)), st.min(), st.max(), en[ok].min(), en[ok].max()))
    d = (en - wk)[ok]
    print("  work duration us: min %.2f p10 %.2f median %.2f p90 %.2f max %.2f mean %.2f;  wait (start->work) mean %.2f max %.2f"
          % (d.min(), np.percentile(d, 10), np.median(d), np.percentile(d, 90), d.max(), d.mean(),
             (wk - st).mean(), (wk - st).max()))
    lo, hi = st.min(), en[ok].max()
    edges = np.arange(np.floor(lo), np.ceil(hi) + 1, 2.0)
    line_res, line_work = [], []
    for a in edges[:-1]:
        mid = a + 1.0
        line_res.append(int(((st <= mid) & (en > mid)).sum()))
        line_work.append(int(((wk <= mid) & (en > mid)).sum()))
    print("  t(us)    : " + " ".join("%5d" % a for a in edges[:-1]))
    print("  resident : " + " ".join("%5d" % v for v in line_res))
    print("  working  : " + " ".join("%5d" % v for v in line_work))
    # per SM: first start, last end, number of CTAs
    firsts = np.array([st[sm == q].min() for q in np.unique(sm)])
    lasts = np.array([en[(sm == q) & ok].max() for q in np.unique(sm)])
    cnt = np.array([(sm == q).sum() for q in np.unique(sm)])
    print("  per SM: first start min/median/max %.1f %.1f %.1f; last end min/median/max %.1f %.1f %.1f; CTAs min/median/max %d %d %d"
          % (firsts.min(), np.median(firsts), firsts.max(), lasts.min(), np.median(lasts), lasts.max(), cnt.min(),
             np.median(cnt), cnt.max()))
    # the slowest CTAs
    order = np.argsort(-(en - wk))[:8]
    print("  slowest: " + ", ".join("cta %d %.1f us (start %.1f)" % (np.flatnonzero(r[:, 0] > 0)[i] if False else i, (en - wk)[i], st[i]) for i in order))


ob = cm.last_cta_trace(2, 360).astype(np.float64)
ob = ob[ob[:, 0] > 0]
if len(ob):
    rel = lambda c: (ob[:, c] - t0) / 1e3
    print("k_obstacle_update ray CTAs (us, mean / max): start %.1f / %.1f, rays traced %.1f / %.1f, marks tested %.1f / %.1f, "
          "box flushed %.1f / %.1f, ticket taken %.1f / %.1f" %
          (rel(0).mean(), rel(0).max(), rel(4).mean(), rel(4).max(), rel(5).mean(), rel(5).max(), rel(6).mean(),
           rel(6).max(), rel(2).mean(), rel(2).max()))
report("k_merge_seed", m)
report("k_inflate", f)
# k_inflate's stages per tile column class (us): seed words, pruning, phase 2, phase 3 + epilogue
nx = (size + 63) // 64
g = f[: (len(f) // nx) * nx].reshape(-1, nx, 8).astype(np.float64)
okk = g[:, :, 4] > 0
for name, sel in (("first column", g[:, 0]), ("interior columns", g[:, 1:-1].reshape(-1, 8)), ("last column", g[:, -1])):
    sel = sel[sel[:, 4] > 0]
    if len(sel):
        print("  %-17s n=%4d  seeds %.2f  prune %.2f  phase2 %.2f  phase3+epilogue %.2f" %
              (name, len(sel), ((sel[:, 4] - sel[:, 1]) / 1e3).mean(), ((sel[:, 5] - sel[:, 4]) / 1e3).mean(),
               ((sel[:, 6] - sel[:, 5]) / 1e3).mean(), ((sel[:, 2] - sel[:, 6]) / 1e3).mean()))
