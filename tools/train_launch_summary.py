"""Aggregate ONE eager train step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of
`GCBF_TRAIN_GRAPH=0 python bench.py --train-only --T 8`: the launches from the last step's first kernel on."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = rows[1:]
names = [r[ki] for r in data]
idx = [i for i, n in enumerate(names) if "small_jobs_kernel" in n]
start = idx[-4] if len(idx) >= 4 else 0        # per step: 2 fold waves + 2 un-fold waves (both networks share them)
pre = ("gather", "graph_build", "reduce_kernel", "elementwise", "mask_count", "fill")
while start > 0 and any(t in names[start - 1] for t in pre):
    start -= 1
agg = collections.OrderedDict()
for r in data[start:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% |")
print(f"\nTotal {tot:.0f} us over {sum(a[0] for a in agg.values())} launches")
