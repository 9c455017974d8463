"""Top stalled SASS instructions from `ncu --page source --csv` output (per kernel section)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) > 5:
        cur["data"].append(r)
for sec in sections[:1] if len(sys.argv) <= 3 else sections:
    hdr = sec["hdr"]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = [(int(r[i_s] or 0), r[i_src].strip(), int(r[i_ex] or 0), k) for k, r in enumerate(sec["data"])]
    tot = sum(d[0] for d in data) or 1
    print(sec["name"], "total samples", tot, "instructions", len(data))
    for s, src, ex, k in sorted(data, key=lambda d: -d[0])[:n]:
        print(f"{s:7d} {100 * s / tot:5.1f}%  ex={ex:9d} #{k:5d} {src[:100]}")
