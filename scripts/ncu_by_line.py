#!/usr/bin/env python3
"""Executed warp-instructions per CUDA source line: joins `ncu --page source --csv` (SASS, in address order) with
`nvdisasm -g -c` of the same cubin (line markers).  usage: ncu_by_line.py <sass.csv> <disasm> <function substring> [units] [top]"""
import collections, csv, re, sys
sass_csv, dis, fn = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c_ex = hdr.index("Instructions Executed")
execd = []
for r in rows[hi + 1:]:
    try:
        execd.append(int(r[c_ex]))
    except (ValueError, IndexError):
        pass
lines, cur, infn = [], None, False
for ln in open(dis, errors="replace"):
    if ln.startswith("//---") and ".text." in ln:
        infn = fn in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
print(f"sass rows {len(execd)}, disasm instructions {len(lines)}")
n = min(len(execd), len(lines))
agg = collections.Counter()
for k in range(n):
    agg[lines[k]] += execd[k]
tot = sum(agg.values())
print(f"total warp-instr {tot}; per 32 units {tot / (units / 32):.1f}")
for (f, l), v in agg.most_common(top):
    print(f"{v / (units / 32):9.1f}  {100.0 * v / tot:5.1f}%  {f}:{l}")
