"""Opcode histogram and the hot instruction runs of one kernel from an .ncu-rep (SASS page: executed instructions per
SASS instruction), plus the VIADDMNMX.U16x2 lines of k_inflate's phase-3 walk as evidence of the packed min-plus.
Usage: sass_opcodes.py REP KERNEL_SUBSTR"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
b = [b for b in blocks if kern in b["name"]][0]
hdr = b["rows"][0]
isrc, iinst = hdr.index("Source"), hdr.index("Instructions Executed")
rows = [(int(r[iinst]), r[isrc].strip()) for r in b["rows"][1:] if len(r) > iinst]
tot = sum(n for n, _ in rows)
ops = collections.Counter()
for n, s in rows:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    ops[(m.group(2) if m else s[:10]).split(".")[0]] += n
print(f"# {b['name']}: {len(rows)} SASS instructions, {tot} executed warp instructions")
print("## executed warp instructions by opcode")
for k, v in ops.most_common(20):
    print(f"  {k:12s} {v:12d} {100 * v / tot:5.1f}%")
v = [(n, s) for n, s in rows if "VIADDMNMX" in s]
print(f"## VIADDMNMX: {len(v)} static, {sum(n for n, _ in v)} executed ({100 * sum(n for n, _ in v) / tot:.1f}% of the kernel)")
print("## the sparse row walk of phase 3 (one VIADD + one VIADDMNMX.U16x2 per output row and cell pair):")
best = max(range(len(rows)), key=lambda i: rows[i][0] if "VIADDMNMX.U16x2" in rows[i][1] and "0x" not in rows[i][1].split(",")[2] else -1)
lo = best
while lo > 0 and best - lo < 30 and "BREV" not in rows[lo][1]:
    lo -= 1
for n, s in rows[lo:best + 3]:
    print(f"  {n:8d}  {s}")
