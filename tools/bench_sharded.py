#!/usr/bin/env python
"""The multi-GPU PRODUCT path in one process: HAVC_colorizer(device_index=[0..N-1]) on ONE clip (vsdeoldify_b200/sharded.py: one
engine + worker thread per GPU, frames partitioned over the GPUs, frames delivered in order), end to end from host frames to
host frames through get_frame().  Also checks that the frames equal what a single GPU renders and that props / order survive.

Usage (on a multi-GPU box): python tools/bench_sharded.py --gpus 2 [--frames 1024] [--partition interleaved|block]
Not the bench contract (bench.py under torchrun is); its JSON goes to profiles/ as evidence for the product path.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import synth_weights  # noqa: E402  (weight generator only)
from vsdeoldify_b200 import havc, vs_shim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--partition", default="interleaved")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sd = synth_weights.make_unet_state_dict("wide", 1234)
    havc.register_state_dict("ColorizeVideo_gen", sd)
    havc._BATCH = a.batch
    os.environ["HAVC_B200_PARTITION"] = a.partition
    base = bench.synth_clip(4 * a.batch, bench.H1080, bench.W1080, seed=100)
    n = a.frames
    idx = np.arange(n) % base.shape[0]

    def src_clip():
        fmt = vs_shim.RGB24
        def fn(i):
            return vs_shim.VideoFrame([base[idx[i], p] for p in range(3)], fmt, {"_SceneChangePrev": int(i == 0), "frame_no": i})
        return vs_shim.VideoNode(n, bench.W1080, bench.H1080, fmt, fn)
    res = {"frames": n, "batch": a.batch, "partition": a.partition}
    ref_frames = {}
    for ndev in sorted({1, a.gpus}):
        out = havc.HAVC_colorizer(src_clip(), method=0, deoldify_p=[0, bench.RF, 1.0, 0.0], ddcolor_p=[1, bench.RF, 1.0, 0.0, True],
                                  device_index=list(range(ndev)) if ndev > 1 else 0)
        for i in range(min(n, 8 * a.batch * ndev)):        # warm: every engine has rendered and pinned its result-buffer pool
            out.get_frame(i)
        t0 = time.perf_counter()
        acc, ok_order = 0, True
        for i in range(n):
            f = out.get_frame(i)
            ok_order = ok_order and f.props["frame_no"] == i
            acc += int(np.asarray(f[0])[0, 0])
            if i in (0, 37, a.batch, n // 2 + 5, n - 1):
                planes = np.stack([np.asarray(f[p]) for p in range(3)])
                if ndev == 1:
                    ref_frames[i] = planes.copy()
                else:
                    res.setdefault("bytes_equal_single_gpu", True)
                    res["bytes_equal_single_gpu"] = bool(res["bytes_equal_single_gpu"] and np.array_equal(planes, ref_frames[i]))
        dt = time.perf_counter() - t0
        res[f"fps_{ndev}gpu"] = n / dt
        fn = getattr(out, "_frame_fn", None)
        rend = getattr(fn, "__defaults__", [None])[0] if fn is not None and getattr(fn, "__defaults__", None) else None
        if rend is not None and hasattr(rend, "stats"):
            res[f"worker_seconds_{ndev}gpu"] = [{k: round(v, 3) for k, v in st.items()} for st in rend.stats]
            res[f"out_pool_{ndev}gpu"] = [len(getattr(e, "_out_pool", [])) for e in rend.engines]
        res[f"order_ok_{ndev}gpu"] = bool(ok_order)
        del out
        torch.cuda.empty_cache()
    if a.gpus > 1:
        res["speedup"] = res[f"fps_{a.gpus}gpu"] / res["fps_1gpu"]
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
