"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # launches to skip (warm-up)
take = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
rows = rows[skip:skip + take]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in rows:
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = row["Metric Unit"]
    ns = v * 1e3 if unit in ("usecond", "us") else (v * 1e6 if unit.startswith("ms") else (v * 1e9 if unit in ("second", "s") else v))
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:80]
    agg[name][0] += 1
    agg[name][1] += ns
    tot += ns
print(f"launches {sum(a[0] for a in agg.values())}  total {tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}%  n={n:5d}  avg={t / n / 1e3:9.1f} us  {k}")
