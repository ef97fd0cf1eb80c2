#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --steps 2 ...` -> kernel-time shares of ONE step.
Usage: python tools/launch_shares.py profiles/r02_launches_b32.csv profiles/r02_launch_shares.txt"""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
names, dur = [r[4] for r in rows], [float(r[-1]) for r in rows]
is_first = lambda n: "resample_h_periodic" in n or "resample_h_rows" in n          # first launch of a step (pre.h)
is_last = lambda n: "post_horizontal" in n                                          # last launch of a step (post.h)
starts = [i for i, n in enumerate(names) if is_first(n)]
s = starts[-1]
e = [i for i, n in enumerate(names) if is_last(n) and i > s]
if not e:
    s = starts[-2]
    e = [i for i, n in enumerate(names) if is_last(n) and i > s]
e = e[0]
tot = sum(dur[s:e + 1])
agg, cnt = collections.defaultdict(float), collections.Counter()
for n, d in zip(names[s:e + 1], dur[s:e + 1]):
    k = n.split("(")[0].replace("void ", "").replace("havc::", "")[:64]
    agg[k] += d
    cnt[k] += 1
with open(dst, "w") as f:
    f.write(f"# one step (B = 32 frames, cfg2, fp16, precision auto) of the ncu launch list {src}\n"
            "# (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: compare SHARES)\n"
            f"# launches in the step: {e - s + 1}; summed kernel time {tot / 1e6:.3f} ms\n")
    conv = sum(v for k, v in agg.items() if "conv_gemm" in k)
    f.write(f"# conv_gemm_kernel (all variants): {100 * conv / tot:.1f} % of the step's kernel time\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        f.write(f"{k:66s} {cnt[k]:3d} launches {v / 1e6:9.3f} ms {100 * v / tot:5.1f} %\n")
print(open(dst).read()[:1500])
