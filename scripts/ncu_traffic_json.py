"""profiles/rN_traffic.json from the ncu DRAM-traffic launch list of the stacked-conv family
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv3d_stack ... scripts/gpu_layers.py`)
and the per-launch labels of the same schedule (gpurun_out/layers_<tag>.csv, written by that gpu_layers.py run):
{"dp_conv3d_stack <label>": bytes per launch (mean over the launches of that shape), "family dp_conv3d_stack batch B size S": bytes per step}.
usage: python scripts/ncu_traffic_json.py <traffic.csv> <layers.csv> <batch> <size> > profiles/rN_traffic.json"""
import collections
import csv
import json
import sys

tcsv, lcsv, B, S = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.DictReader([l for l in open(tcsv) if l.startswith('"')]))
per = collections.OrderedDict()
for r in rows:
    if not r["Metric Name"].startswith("dram__bytes"):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    per[int(r["ID"])] = per.get(int(r["ID"]), 0.0) + v
labels = [r["label"] for r in csv.DictReader(open(lcsv)) if r["family"] == "dp_conv3d_stack"]
vals = list(per.values())
assert len(vals) >= len(labels), (len(vals), len(labels))
by = collections.defaultdict(list)
for lab, v in zip(labels, vals):
    by[lab].append(v)
out = {f"dp_conv3d_stack {k}": int(sum(v) / len(v)) for k, v in by.items()}
out[f"family dp_conv3d_stack batch {B} size {S}"] = int(sum(vals[:len(labels)]))
print(json.dumps(out, indent=1))
