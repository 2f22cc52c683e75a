#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv`: instruction mix per unit of work and the hottest SASS lines.
usage: ncu -i prof.ncu-rep --page source --csv | python profiles/ncu_src_summary.py <units_of_work> [top_n]"""
import collections
import csv
import sys

units = float(sys.argv[1])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c_src, c_ex, c_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, tot, samp_tot = collections.Counter(), 0, 0
lines = []
for r in rows[hi + 1:]:
    if len(r) <= c_ex:
        continue
    try:
        n = int(r[c_ex]); s = int(r[c_samp])
    except ValueError:
        continue
    toks = r[c_src].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
    agg[op.split(".")[0]] += n
    tot += n
    samp_tot += s
    lines.append((s, n, r[c_src].strip()))
print(f"warp-instructions executed: {tot}  per unit of work (x32 lanes): {tot * 32 / units:.2f} thread-instr, {tot / (units / 32):.2f} warp-instr per 32 units")
print("opcode mix (warp-instr per 32 units of work):")
for k, v in agg.most_common(top):
    print(f"  {k:10s} {v / (units / 32):7.2f}")
print("hottest lines by stall samples:")
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(f"  {100.0 * s / max(samp_tot, 1):5.1f}%  {n:>12d}  {src}")
