"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py profiles/r1_launches_bench_cfg2.csv > profiles/r1_launches_bench_cfg2.txt"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r[iu], 1)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
ours = lambda k: "sigops::k_" in k and "peak" not in k   # noqa: E731
lib = sum(v[1] for k, v in agg.items() if ours(k)) or 1.0
print("# per-launch times are cold-cache and serialised: compare shares, not absolutes\n")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    note = f"{100 * v[1] / lib:5.1f}% of the library's kernels" if ours(k) else "(bench harness: torch RNG / peak micro-benchmarks)"
    print(f"{k[:92]:92s} launches={v[0]:4d} total={v[1]:10.1f} us  {note}")
