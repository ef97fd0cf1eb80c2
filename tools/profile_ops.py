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
# whole-batch graph time and the share outside the network launch list (pre/post/head pixel passes)
eng2 = eng
import time
g_times = []
with torch.cuda.stream(eng.compute):
    for i in range(3):
        eng._launch(0, eng.compute.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.compute)
    for i in range(5):
        eng._launch(0, eng.compute.cuda_stream)
    e1.record(eng.compute)
eng.compute.synchronize()
full = e0.elapsed_time(e1) / 5
print(f"full batch (pre + net + head + post), back-to-back launches: {full:.3f} ms -> {a.batch / full * 1000:.1f} frames/s; outside net list: {full - tot:.3f} ms")
lib = eng.lib
import ctypes
def timed(label, fn):
    with torch.cuda.stream(eng.compute):
        fn(); fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.compute)
        for _ in range(5): fn()
        e1.record(eng.compute)
    eng.compute.synchronize()
    print(f"  {label:10s} {e0.elapsed_time(e1) / 5:8.4f} ms")
st = eng.compute.cuda_stream
B, S, W, H = eng.B, eng.S, eng.W, eng.H
td, tv, uh, uv = eng.t_down_h, eng.t_down_v, eng.t_up_h, eng.t_up_v
timed("pre.h", lambda: td.resample_h(lib, eng.d_in[0].data_ptr(), eng.tmp_down.data_ptr(), B * 3 * H, st))
timed("pre.h (table kernel)", lambda: lib.havc_resample_h(eng.d_in[0].data_ptr(), eng.tmp_down.data_ptr(), B * 3 * H, W, S, td.start.data_ptr(), td.wt.data_ptr(), td.taps, st))
timed("pre.v", lambda: lib.havc_pre_vertical(eng.tmp_down.data_ptr(), eng.rgb_small.data_ptr(), eng.prog.x.data_ptr(), B, H, S, tv.start.data_ptr(), tv.w.data_ptr(), tv.taps, eng.hd, st))
timed("head", lambda: lib.havc_head(eng.prog.logits.data_ptr(), 0, None, eng.prog.b11.data_ptr(), eng.rgb_small.data_ptr(), eng.colored.data_ptr(), None, eng.skip_slots[0].data_ptr(), B, S, eng.hd, 1, st))
timed("post.v", lambda: lib.havc_resample_v(eng.colored.data_ptr(), eng.tmp_up.data_ptr(), B * 3, S, H, S, uv.start.data_ptr(), uv.w.data_ptr(), uv.taps, st))
timed("post.h", lambda: uh.post_horizontal(lib, eng.tmp_up.data_ptr(), eng.d_in[0].data_ptr(), eng.d_out[0].data_ptr(), B, H, 1, st))
timed("post.h (table kernel)", lambda: lib.havc_post_horizontal(eng.tmp_up.data_ptr(), eng.d_in[0].data_ptr(), eng.d_out[0].data_ptr(), B, S, H, W, uh.start.data_ptr(), uh.wt.data_ptr(), uh.taps, 1, st))
