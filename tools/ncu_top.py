"""Summarise an `ncu --page source --csv` dump: top stalled SASS instructions and executed-instruction mix.
usage: ncu -i X.ncu-rep --page source --csv > x.csv; python tools/ncu_top.py x.csv [N]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
body = [r for r in rows[2:] if len(r) > 6]
ci, cs, ce = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(float(r[cs] or 0) for r in body)
texec = sum(float(r[ce] or 0) for r in body)
print(f"{len(body)} SASS instructions, {tot:.0f} stall samples, {texec:.3e} warp-instructions executed")
mix = Counter()
for r in body:
    op = r[ci].strip().split()[0] if r[ci].strip() else "?"
    if op.startswith("@"):
        op = r[ci].strip().split()[1]
    mix[op.split(".")[0]] += float(r[ce] or 0)
print("instruction mix:", ", ".join(f"{k} {100 * v / texec:.1f}%" for k, v in mix.most_common(18)))
print("top stalls:")
for idx, r in sorted(enumerate(body), key=lambda t: -float(t[1][cs] or 0))[:n]:
    print(f"  #{idx:5d} {float(r[cs]):7.0f} ({100 * float(r[cs]) / tot:4.1f}%) exec {float(r[ce]):10.0f}  {r[ci].strip()[:100]}")
