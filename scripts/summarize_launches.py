"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  x{cnt[k]:3d}  {k}")
