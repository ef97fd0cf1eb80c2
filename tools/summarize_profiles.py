#!/usr/bin/env python
"""Turns the artefacts of tools/gpu_measure.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__cluster_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg']

def val(hdr, units, r, name):
    i = hdr.index(name)
    return float(r[i]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "Ghz": 1e9}.get(units[i], 1)

hdr, units, rows = raw_rows(os.path.join(G, "final_shuf8_resconv0.ncu-rep"))
shuf, res = rows[0], rows[1]
rd, wr = val(hdr, units, res, "dram__bytes_read.sum"), val(hdr, units, res, "dram__bytes_write.sum")
json.dump({"kernel": "conv_gemm_kernel<f16, generic, pair> launch res.conv0 (dominant launch of the step)", "batch": 32,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 5511315456,
           "source": "profiles/r01_ncu_final_shuf8_resconv0.txt (ncu --set full, gpurun_out/final_shuf8_resconv0.ncu-rep, tools/gpu_measure.sh)"},
          open(os.path.join(P, "r01_roofline_traffic.json"), "w"), indent=1)
def line(r, name): return val(hdr, units, r, name)
with open(os.path.join(P, "r01_ncu_final_shuf8_resconv0.txt"), "w") as f:
    f.write("# ncu --set full --import-source on --clock-control none, launches 120 and 121 of conv_gemm_kernel in\n"
            "# 'python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph' (B = 32, fp16), final code of round 1 (tools/gpu_measure.sh).\n\n")
    t, ghz = line(shuf, "gpu__time_duration.sum"), line(shuf, "sm__cycles_elapsed.avg.per_second") / 1e9
    tb = (line(shuf, "dram__bytes_read.sum") + line(shuf, "dram__bytes_write.sum")) / t / 1e12
    f.write("## 1. shuf8.conv: PixelShuffle 1x1 conv 256 -> 4x256 @192x192 (K = 256: 4 K steps), fast 16-warp epilogue, CTA pairs.\n"
            f"# Epilogue / store bound: {tb:.2f} TB/s of HBM traffic (algorithmic 2.42 GB out + 0.60 GB in) at {ghz:.2f} GHz.\n")
    for h, u, v in zip(hdr, units, shuf):
        if h in WANT or h == "Kernel Name": f.write(f"{h} [ {u} ] = {v}\n")
    ghz = line(res, "sm__cycles_elapsed.avg.per_second") / 1e9
    tp = line(res, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    ex = tp / 100 * 8192 * 148 * ghz / 1e3
    f.write("\n## 2. res.conv0: res_block 3x3 conv, K = 9 x 256 + one im2col chunk (37 K steps), N = 272 = 144 + 128, generic epilogue, CTA pairs.\n"
            f"# Tensor pipe active {tp:.1f} % of the cycles at the power-capped {ghz:.2f} GHz: {tp/100:.3f} x 8192 FLOP/clk x 148 SMs x {ghz:.3f} GHz = {ex:.0f} TFLOP/s executed\n"
            f"# (cuBLAS bf16 sustains 1373 TFLOP/s at 1.31 GHz on this pool) = {ex/1.066:.0f} algorithmic x 1.066 (N 259->272, K 2331->2368 padding).\n"
            f"# DRAM {(rd+wr)/1e9:.2f} GB vs 5.51 GB algorithmic (u 2.42 + x-col 0.60 in, r1 2.42 + r1x 0.08 out).\n")
    for h, u, v in zip(hdr, units, res):
        if h in WANT or h == "Kernel Name": f.write(f"{h} [ {u} ] = {v}\n")
for a, b in (("bench_default.json", "r01_bench_default.json"), ("bench_reference.json", "r01_bench_reference.json"),
             ("launches_b32.csv", "r01_launches_b32.csv"), ("pytest_gpu.txt", "r01_pytest_gpu.txt")):
    if os.path.exists(os.path.join(G, a)): shutil.copy(os.path.join(G, a), os.path.join(P, b))
# kernel shares of one step from the launch list
rows = [r for r in csv.reader(open(os.path.join(P, "r01_launches_b32.csv"))) if len(r) > 10 and r[0].isdigit()]
names, dur = [r[4] for r in rows], [float(r[-1]) for r in rows]
starts = [i for i, n in enumerate(names) if "resample_h_rows" in n]; ends = [i for i, n in enumerate(names) if "post_horizontal" in n]
s = starts[-1]; e = [x for x in ends if x > s]
if not e: s = starts[-2]; e = [x for x in ends if x > s]
e = e[0]; tot = sum(dur[s:e + 1]); agg = collections.defaultdict(float); cnt = collections.Counter()
for n, d in zip(names[s:e + 1], dur[s:e + 1]):
    k = n.split("(")[0].replace("void ", "").replace("havc::", "")[:60]; agg[k] += d; cnt[k] += 1
with open(os.path.join(P, "r01_launch_shares.txt"), "w") as f:
    f.write(f"# one step (B = 32) of profiles/r01_launches_b32.csv (ncu gpu__time_duration, serialised, cold caches): {e-s+1} launches, {tot/1e6:.3f} ms\n")
    conv = sum(v for k, v in agg.items() if "conv_gemm" in k)
    f.write(f"# conv_gemm_kernel (all variants): {100*conv/tot:.1f} % of the step's kernel time\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): f.write(f"{k:62s} x{cnt[k]:3d} {v/1e6:9.3f} ms {100*v/tot:5.1f} %\n")
d = json.load(open(os.path.join(P, "r01_bench_default.json")))
print("value", d["value"], "e2e", d["e2e"]["value"], "plugin", d.get("plugin_surface"), "frac", d["roofline"]["frac"], "whole", d["tensor_frac_whole_step"], d["clocks"])
print(open(os.path.join(P, "r01_launch_shares.txt")).read()[:900])
