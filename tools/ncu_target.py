#!/usr/bin/env python
"""Runs selected launches of an engine inside an NVTX range so that ncu can capture exactly them:

  ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "target/" -o gpurun_out/x \
      python tools/ncu_target.py --config cfg2 --ops res.conv0,shuf8.conv+blur,pixel

The whole launch list runs once first (buffers hold real data), then every requested op runs once inside the range, in the
order given.  `pixel` = the five pixel-pass launches (pre.h, pre.v, head, post.v, post.h).  Not a bench."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default=None)
    ap.add_argument("--ops", default="res.conv0")
    ap.add_argument("--prog", default="prog", help="prog | prog2 | zhang")
    a = ap.parse_args()
    eng, w, h, desc = bench.build_config_engine(a.config, "cuda:0", torch.float16, a.batch, precision=a.precision)
    clip = bench.synth_clip(a.batch, h, w, seed=100)
    eng.colorize_batch(clip)                                   # everything has run once on real data
    prog = {"prog": eng.prog, "prog2": eng.prog2, "zhang": eng.zhang.prog if eng.zhang is not None else None}[a.prog]
    by_name = {op.name: op for op in prog.ops}
    st = eng.compute.cuda_stream
    print("ops of", a.prog, ":", len(prog.ops))
    with torch.cuda.stream(eng.compute):
        for name in a.ops.split(","):
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push("target")
            if name == "pixel":
                eng._launch_pre(0, st)
                eng._launch_post(0, st)
            elif name in by_name:
                by_name[name].fn(st)
            else:
                print("unknown op", name, "- known:", ", ".join(list(by_name)[:200]))
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
    print("done:", desc)


if __name__ == "__main__":
    main()
