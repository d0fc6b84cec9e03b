"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list: one forward step (between two consecutive coord_max kernels), per-kernel totals, shares and DRAM bytes.
Usage: summarize_launches.py in.csv out.md ["command line shown in the header"]"""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                                             "--clock-control none -s 1200 -c 700 --csv python bench.py --steps 1 --warmup 3 "
                                             "--no-cpu-baseline")
lines = [l for l in open(src) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ii, ki, mi, vi, ui = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
launch = collections.OrderedDict()  # id -> [kernel, ns, dram_read, dram_write]
for row in r:
    rec = launch.setdefault(row[ii], [row[ki], 0.0, 0.0, 0.0])
    v = float(row[vi].replace(",", ""))
    if row[mi] == "gpu__time_duration.sum":
        rec[1] = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row[ui], 1.0)
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row[ui], 1.0)
        rec[2 if "read" in row[mi] else 3] = v
seq = list(launch.values())
idx = [i for i, rec in enumerate(seq) if "coord_max" in rec[0]]
a, b = idx[0], (idx[1] if len(idx) > 1 else len(seq))
step = seq[a:b]
tot = sum(rec[1] for rec in step)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k, ns, rd, wr in step:
    k = k.split("(")[0]
    e = agg[k]
    e[0] += 1
    e[1] += ns
    e[2] += rd
    e[3] += wr
own = sum(e[1] for k, e in agg.items() if "ss::" in k)
dram = sum(e[2] + e[3] for e in agg.values())
with open(dst, "w") as f:
    f.write(f"# ncu launch list, one PTv3 forward step ({len(step)} launches)\n\n")
    f.write(f"Command: `{cmd}` (B200, cold-cache serialised launches: compare SHARES, not absolutes).\n\n")
    f.write(f"Sum of kernel durations in the step: {tot / 1e6:.2f} ms; own kernels (`ss::`) {own / 1e6:.2f} ms "
            f"({100 * own / tot:.1f} %), library kernels {100 - 100 * own / tot:.1f} %.  DRAM traffic of the step: "
            f"{dram / 1e9:.1f} GB.\n\n")
    f.write("| ms | share | launches | DRAM read GB | DRAM write GB | kernel |\n|---:|---:|---:|---:|---:|---|\n")
    for k, (c, v, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        f.write(f"| {v / 1e6:.3f} | {100 * v / tot:.1f} % | {c} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | `{k[:110]}` |\n")
    att = [(k, e) for k, e in agg.items() if "patch_attention_tc" in k]
    if att:
        c = sum(e[0] for _, e in att)
        by = sum(e[2] + e[3] for _, e in att)
        f.write(f"\nPatch attention: {c} launches, {by / 1e9:.3f} GB DRAM in total = {by / c / 1e6:.1f} MB per launch "
                f"(algorithmic N x 4C x 2 B per launch, summed over the step: see DESIGN.md).\n")
print(open(dst).read()[:6000])
