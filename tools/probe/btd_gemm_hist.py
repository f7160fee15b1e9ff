"""Per-launch GEMM times of the serial block-tridiagonal factor, grouped by (grid, flops) class."""
import collections
import csv
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("seq,") or True))
rows = [r for r in rows if r["seq"] != "seq"]
agg = collections.OrderedDict()
for r in rows:
    if "gemm" not in r["name"]:
        continue
    key = (r["name"], int(r["grid"]), float(r["flops"]))
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r["ms"])
tot = sum(a[1] for a in agg.values())
print("GEMM total ms", tot)
for (name, grid, fl), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-22s grid=%6d GF=%9.3f n=%4d ms=%8.3f share=%5.1f%% avg_us=%8.1f TF=%6.2f" % (
        name, grid, fl * 1e-9, a[0], a[1], 100 * a[1] / tot, 1e3 * a[1] / a[0], fl * a[0] / a[1] * 1e-9))
