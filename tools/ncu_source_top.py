#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
tot = sum(int(r[idx["# Samples"]]) for r in data)
ins = sum(int(r[idx["Instructions Executed"]]) for r in data)
print(f"samples {tot}, SASS lines {len(data)}, warp instructions executed {ins}")
reasons = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tots = {k: sum(int(r[idx[k]]) for r in data) for k in reasons}
for k, v in sorted(tots.items(), key=lambda x: -x[1])[:8]:
    print(f"  {k}: {100 * v / max(tot, 1):.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    rs = {k: int(r[idx[k]]) for k in reasons}
    main = max(rs, key=rs.get)
    print(f"{int(r[idx['# Samples']]):6d} {r[idx['Source']].strip()[:72]:72s} exec={r[idx['Instructions Executed']]:>9s} {main}")
