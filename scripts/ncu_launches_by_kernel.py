"""Per-kernel table from an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, ms, share."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
by = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dp::", "").replace("(int)", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    by[name][0] += 1
    by[name][1] += ms
tot = sum(v[1] for v in by.values())
print(f"Sum: {tot:.2f} ms in {sum(v[0] for v in by.values())} launches\n")
print("| kernel | launches | ms (ncu) | share |\n|---|---|---|---|")
for k, (n, ms) in sorted(by.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % |")
fam = collections.defaultdict(float)
for k, (n, ms) in by.items():
    fam[re.sub(r"<.*", "", k)] += ms
print("\n| family | share under ncu |\n|---|---|")
for k, ms in sorted(fam.items(), key=lambda x: -x[1]):
    print(f"| `{k}` | {100 * ms / tot:.1f} % |")
