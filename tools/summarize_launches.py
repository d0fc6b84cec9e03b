"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one forward step (between two
consecutive coord_max kernels), per-kernel totals and shares.  Usage: summarize_launches.py in.csv out.md"""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for row in r:
    v = float(row[vi].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row[ui], 1.0)
    seq.append((row[ki], v))
idx = [i for i, (k, _) in enumerate(seq) if "coord_max" in k]
a, b = idx[0], (idx[1] if len(idx) > 1 else len(seq))
step = seq[a:b]
tot = sum(v for _, v in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    k = k.split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
own = sum(v for k, (c, v) in agg.items() if "ss::" in k)
with open(dst, "w") as f:
    f.write(f"# ncu launch list, one PTv3 forward step ({len(step)} launches)\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 800 --csv python bench.py "
            "--steps 1 --warmup 3 --no-cpu-baseline` (B200, cold-cache serialised launches: compare SHARES).\n\n")
    f.write(f"Sum of kernel durations in the step: {tot / 1e6:.2f} ms; own kernels (`ss::`) {own / 1e6:.2f} ms "
            f"({100 * own / tot:.1f} %), library kernels {100 - 100 * own / tot:.1f} %.\n\n")
    f.write("| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write(f"| {v / 1e6:.3f} | {100 * v / tot:.1f} % | {c} | `{k[:110]}` |\n")
print(open(dst).read()[:1500])
