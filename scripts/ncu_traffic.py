"""ncu CSV (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum; one training step captured
with `--profile-from-start off` around scripts/one_step.py) -> profiles/r02_ncu_traffic_<kernel>.json: DRAM traffic per
launch of ONE kernel class over every launch of the step - the same launch mix bench.py's roofline averages its
algorithmic bytes over.   usage: ncu_traffic.py <csv> <kernel substring> <batch> <math> <out.json>"""
import csv, json, re, sys
path, kern, batch, math, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
lines = [l for l in open(path) if not l.startswith("==")]
per = {}
for row in csv.DictReader(lines):
    if kern not in row["Kernel Name"]:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3,
             "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
    d = per.setdefault(int(row["ID"]), {"name": re.sub(r"\(.*", "", row["Kernel Name"])[:80]})
    d[row["Metric Name"]] = v * scale
launches = [per[k] for k in sorted(per)]
tot = [l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in launches]
res = {"kernel": kern, "batch_per_gpu": batch, "math": math, "launches": len(launches),
       "dram_bytes_per_launch_mean": sum(tot) / max(len(tot), 1), "dram_bytes_per_step": sum(tot),
       "per_launch": [{"name": l["name"], "dram_read": l.get("dram__bytes_read.sum"), "dram_write": l.get("dram__bytes_write.sum"),
                       "ns": l.get("gpu__time_duration.sum")} for l in launches],
       "how": "ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
              "--clock-control none python scripts/one_step.py <batch> <math> 1"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}))
