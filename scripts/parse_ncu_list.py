import csv, re, sys
path = sys.argv[1]; out_path = sys.argv[2] if len(sys.argv) > 2 else None; title = sys.argv[3] if len(sys.argv) > 3 else ""; ns = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
lines = [l for l in open(path) if not l.startswith("==")]
agg = {}; n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(anonymous namespace\)::", "", row["Kernel Name"]); name = re.sub(r"^void ", "", name); name = re.sub(r"\(.*", "", name)
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    us = v / 1000.0 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; n += 1
tot = sum(a[1] for a in agg.values())
out = [title, f"launches {n} total {tot/1000:.2f} ms ({ns:g} steps in the capture) -> per step {n/ns:.0f} launches, {tot/1000/ns:.2f} ms kernel time (ncu: cold-cache, serialised)"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:48]:
    out.append(f"{k[:78]:78s} n={a[0]/ns:6.1f}/step {a[1]/1000/ns:9.3f} ms/step {100*a[1]/tot:5.1f}%")
print("\n".join(out))
if out_path: open(out_path, "w").write("\n".join(out) + "\n")
