"""Prints `kernel grid duration_ns...` from an ncu --csv log (gpu__time_duration.sum rows)."""
import csv, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
out = {}
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        key = (r["Kernel Name"].split("(")[0], r["Grid Size"])
        out.setdefault(key, []).append(float(r["Metric Value"].replace(",", "")))
for k, v in out.items():
    print(k[0], k[1], " ".join(f"{x:.0f}" for x in v))
