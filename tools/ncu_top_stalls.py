"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` (first profiled launch)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
first = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[ia] == "Address":
        if first:
            break
        continue
    first.append(r)
tot = sum(int(r[isamp]) for r in first)
print("instructions", len(first), "samples", tot)
for r in sorted(first, key=lambda r: -int(r[isamp]))[:n]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i] not in ("", "0")), reverse=True)[:2]
    print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  {r[isrc].strip()[:80]:80s} {st}")
