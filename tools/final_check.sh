#!/bin/bash
# What the round ends with (run under gpurun on one GPU): the GPU test-suite, smoke(), the default bench line.
python -m pytest tests -m gpu -q 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d["dwa"].items() if k in ("c2_findBestPath_us","c4_sweep_ms","c4_traj_per_s","c5_cycle_ms","c5_e2e_cycle_ms")})
PY
