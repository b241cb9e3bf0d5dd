"""Per-kernel share of a step from an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list:
    python tools/launch_list_summary.py gpurun_out/launches.csv > profiles/ncu_rNN_launch_list_summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot, n = collections.OrderedDict(), 0.0, 0
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("v2s::<unnamed>::", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
print(f"# {n} launches, {tot / 1000:.3f} ms total (cold-cache, serialised launches: compare shares, not absolutes)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={c:4d} {v:10.1f} us {100 * v / tot:5.1f} %")
