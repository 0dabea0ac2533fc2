#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/one_step.py 4096 tf32 3 > gpurun_out/one_step.log 2>&1 && cat gpurun_out/one_step.log && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/one_step.py 4096 tf32 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu exit $?"
python scripts/parse_ncu_list.py gpurun_out/launches.csv gpurun_out/launch_summary.txt "ncu launch list"
exit 0
python - <<'PY'
import csv, collections, re
rows = []
with open("gpurun_out/launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
n = 0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"<.*", "", row["Kernel Name"]).split("(")[0]
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; n += 1
tot = sum(a[1] for a in agg.values())
print(f"launches {n} total {tot/1000:.2f} ms (both steps: warm-up + timed)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[:60]:60s} n={a[0]:5d} {a[1]/1000:9.3f} ms {100*a[1]/tot:5.1f}%")
PY
