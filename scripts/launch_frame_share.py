"""Per-frame kernel shares from an ncu launch list of bench.py: the value loop's timed steps are the segments between
consecutive L2-flush kernels (at::FillFunctor); prints the average over those frames."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
full = [r["Kernel Name"] for r in rows]
dur = [float(r["Metric Value"].replace(",", "")) / 1000.0 for r in rows]


def short(n):
    m = re.search(r"(k_[a-z0-9_]+)", n)
    return m.group(1) if m else n.split("(")[0][:40]


idx = [i for i, n in enumerate(full) if "FillFunctor" in n]
segs = [(a, b) for a, b in zip(idx[:-1], idx[1:]) if b - a < 200]
tot = collections.defaultdict(float)
cnt = collections.defaultdict(int)
for a, b in segs:
    for i in range(a + 1, b):
        tot[short(full[i])] += dur[i]
        cnt[short(full[i])] += 1
n = len(segs)
frame = sum(tot.values()) / n
print(f"{n} timed frames of the value loop, {sum(cnt.values()) / n:.0f} launches and {frame:.1f} us of kernel time per frame (serialised under ncu)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:28s} n/frame={cnt[k] / n:5.1f}  us/frame={v / n:7.1f}  avg={v / cnt[k]:6.2f} us  share={100 * v / n / frame:5.1f} %")
