"""Reads an .ncu-rep (source page, SASS) and attributes executed instructions and stall samples to CUDA source lines
using nvdisasm's line info for the same kernel in libnavgpu.so (rebuild nothing between the capture and this).
Usage: ncu_lines.py REP KERNEL_SUBSTR [launch_idx] [MANGLED_SUBSTR]"""
import csv
import io
import re
import subprocess
import sys
import collections

rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mangled = sys.argv[4] if len(sys.argv) > 4 else kern
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split per kernel block
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blocks = [b for b in blocks if kern in b["name"]]
b = blocks[which]
hdr = b["rows"][0]
ia, isrc, iinst, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sass = [(r[isrc].strip(), int(r[iinst]), int(r[ismp]), [int(r[i]) for i in stall_cols]) for r in b["rows"][1:] if len(r) > ismp]

# line info from the cubin
lib = "navigation_b200/libnavgpu.so"
subprocess.run(f"cuobjdump -xelf all /root/repo/{lib} >/dev/null", shell=True, cwd="/tmp")
import glob
lines = None
for cub in glob.glob("/tmp/*.cubin"):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    # find the function section
    m = re.search(r"\.text\.[^\n]*" + re.escape(mangled) + r"[^\n]*\n", dis)
    if not m:
        continue
    body = dis[m.end():]
    nxt = re.search(r"\n//-+ \.text\.|\n\t\.section", body)
    if nxt:
        body = body[:nxt.start()]
    lines = []
    curline = "?"
    for ln in body.split("\n"):
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            curline = f"{mm.group(1).split('/')[-1]}:{mm.group(2)}"
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(curline)
    break
assert lines is not None, "kernel not found in cubin"
print(f"# {b['name']}: {len(sass)} SASS instr in report, {len(lines)} in cubin")
n = min(len(sass), len(lines))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
names = [hdr[i] for i in stall_cols]
tot_i = sum(s[1] for s in sass)
tot_s = sum(s[2] for s in sass)
for k in range(n):
    a = agg[lines[k]]
    a[0] += sass[k][1]
    a[1] += sass[k][2]
    for nm, v in zip(names, sass[k][3]):
        a[2][nm] += v
print(f"# total warp-instructions {tot_i}, samples {tot_s}")
def key(l):
    f, ln = l.rsplit(":", 1) if ":" in l else (l, "0")
    return (f, int(ln) if ln.isdigit() else 0)
for l in sorted(agg, key=key):
    i, s, st = agg[l]
    if i < 0.003 * tot_i and s < 0.003 * tot_s:
        continue
    top = ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(3) if v)
    print(f"{l:28s} inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%   {top}")
