#!/usr/bin/env python
"""Per-launch CUDA-event timing of one batch of the DeOldify engine (not a bench value: serialised launches)."""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth_weights
from vsdeoldify_b200.engine import DeoldifyEngine

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="wide"); ap.add_argument("--rf", type=int, default=24)
ap.add_argument("--batch", type=int, default=8); ap.add_argument("--w", type=int, default=1920); ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--dtype", default="fp16"); ap.add_argument("--out", default="gpurun_out/ops.json")
a = ap.parse_args()
dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
sd = synth_weights.make_unet_state_dict(a.arch, 1234)
eng = DeoldifyEngine(sd, a.w, a.h, render_factor=a.rf, batch=a.batch, dtype=dt, use_graph=False)
rows = []
with torch.cuda.stream(eng.compute):
    for rep in range(3):
        evs = []
        for op in eng.prog.ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(eng.compute); op.fn(eng.compute.cuda_stream); e1.record(eng.compute)
            evs.append((op, e0, e1))
eng.compute.synchronize()
tot = 0.0
for op, e0, e1 in evs:
    ms = e0.elapsed_time(e1); tot += ms
    d = op.fn.__self__.desc if hasattr(op.fn, "__self__") and hasattr(op.fn.__self__, "desc") else None
    rows.append(dict(name=op.name, kind=op.kind, ms=ms, gflop=op.flops / 1e9, tflops=(op.flops / ms / 1e9) if ms > 0 and op.flops else 0,
                     bn=d.BN if d else None, n=d.N_total if d else None, box=(d.box_w, d.box_h, d.box_b) if d else None,
                     taps=d.n_taps if d else None, k=(d.src0.C + (d.src1.C if d.src1.ptr else 0)) if d else None,
                     gbs=(op.bytes / ms / 1e6) if ms > 0 and op.bytes else 0))
rows_sorted = sorted(rows, key=lambda r: -r["ms"])
print(f"total serialised ms/batch {tot:.3f}  ({a.batch} frames)  gemm {sum(r['ms'] for r in rows if r['kind']=='gemm'):.3f}  aux {sum(r['ms'] for r in rows if r['kind']!='gemm'):.3f}")
for r in rows_sorted[:45]:
    print(f"{r['name']:34s} {r['ms']:8.4f} ms  {r['tflops']:7.1f} TF  {r['gbs']:7.0f} GB/s  BN={r['bn']} N={r['n']} K={r['k']} taps={r['taps']} box={r['box']}")
json.dump(rows, open(a.out, "w"))
