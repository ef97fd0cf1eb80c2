#!/usr/bin/env python
"""bench.py — 1080p colourised frames/s of the HAVC per-frame hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): DeOldify 'video' generator (ResNet-101 DynamicUnetWide), render_factor 24,
HAVC_colorizer(method=0) semantics on a synthetic 1080p 8-bit grayscale clip; synthetic weights (seed 1234,
reference state-dict schema).  A "step" = one batch of B frames through
   Spline64 squeeze -> network -> S x S luma transplant -> Spline64 back -> full-res luma transplant.

  value   : frames/s with the input batches already resident in HBM (CUDA-graph replays, CUDA events)
  e2e     : frames/s through the engine's public host API: inputs in pinned HOST memory, H2D of every input frame and D2H
            of every output frame into pinned host memory inside the timed region (copies overlap compute on separate
            streams), the host touches every result batch
  roofline: the tcgen05 implicit-GEMM kernel (the dominant kernel): algorithmic conv/attention FLOPs of its
            launches / the summed CUDA-event durations of those launches, vs the measured sustained bf16 peak
  cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/, torch fp32 on all host
            cores) on a bounded sample of the same workload.

N > 1 (torchrun): frames are block-partitioned over ranks, no collective on the data path; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W1080, H1080, RF = 1920, 1080, 24
GFLOP_PER_FRAME_SURVEY = 639.0
WORKLOAD = "DeOldify video rf=24 (ResNet-101 DynamicUnetWide @384x384), HAVC_colorizer(method=0), synthetic 1080p grayscale clip"


def synth_clip(n: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """Seeded synthetic grayscale clip [n,3,h,w] (R=G=B): low-frequency cosines + noise + moving rectangles,
    mean luma swept over the clip."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h, dtype=np.float32), np.linspace(0, 1, w, dtype=np.float32), indexing="ij")
    out = np.empty((n, 3, h, w), np.uint8)
    for i in range(n):
        img = np.zeros((h, w), np.float32)
        for _ in range(6):
            f = rng.uniform(1, 32)
            th = rng.uniform(0, np.pi)
            ph = rng.uniform(0, 2 * np.pi)
            img += np.cos(2 * np.pi * f * (xx * np.cos(th) + yy * np.sin(th)) + ph) * rng.uniform(0.05, 0.2)
        mean = 0.05 + 0.9 * (i + 0.5) / n
        img = img + mean + rng.normal(0, 6 / 255.0, (h, w)).astype(np.float32)
        for _ in range(3):
            x0, y0 = rng.integers(0, w - 64), rng.integers(0, h - 64)
            ww, hh = rng.integers(32, w // 4), rng.integers(32, h // 4)
            img[y0:y0 + hh, x0:x0 + ww] += rng.uniform(-0.3, 0.3)
        g = (np.clip(img, 0, 1) * 255).astype(np.uint8)
        out[i] = g[None]
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in self.rows if len(r) > 3 + j)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf_sustained=d.get("bf16_tflops_sustained", 1373.2), tf_burst=d.get("bf16_tflops", 1604.5),
                    hbm=d.get("hbm_gbs", 6535.7), src="measured")
    return dict(tf_sustained=1400.0, tf_burst=1590.0, hbm=6650.0, src="fallback")


def load_traffic(batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant conv_gemm launch (res.conv0) from the
    committed `ncu --set full` capture (profiles/r01_roofline_traffic.json), scaled linearly to this batch."""
    p = os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["dram_bytes_per_launch"] * batch / d["batch"]


def cpu_reference_fps(n_frames: int, threads: int, h: int = H1080, w: int = W1080, rf: int = RF, seed: int = 1234):
    """The CPU restatement of the reference path (oracle/pipeline_oracle.py), end to end per frame."""
    from oracle import pipeline_oracle, synth_weights
    torch.set_num_threads(threads)
    sd = synth_weights.make_unet_state_dict("wide", seed)
    clip = synth_clip(n_frames + 1, h, w, seed=0)
    pipeline_oracle.havc_colorizer_frame(sd, np.ascontiguousarray(np.transpose(clip[0], (1, 2, 0))), rf)  # warm-up
    times = []
    for i in range(1, n_frames + 1):
        t0 = time.perf_counter()
        pipeline_oracle.havc_colorizer_frame(sd, np.ascontiguousarray(np.transpose(clip[i], (1, 2, 0))), rf)
        times.append(time.perf_counter() - t0)
    return 1.0 / float(np.median(times)), times


def plugin_surface_fps(sd, clip: np.ndarray, n_frames: int, batch: int):
    """frames/s of HAVC_colorizer(method=0) itself: a clip of the in-repo VapourSynth stand-in in, frames out through
    get_frame() in order from ONE host thread (plane stacking, f.copy(), per-plane copies and the read-ahead pipeline
    included) - what a script switching from the reference calls."""
    from vsdeoldify_b200 import havc, vs_shim
    havc.register_state_dict("ColorizeVideo_gen", sd)
    havc._BATCH = batch
    n = min(n_frames, clip.shape[0])
    src = vs_shim.array_clip(clip[:n], props=[{"_SceneChangePrev": int(i == 0)} for i in range(n)])
    out = havc.HAVC_colorizer(src, method=0, deoldify_p=[0, RF, 1.0, 0.0], ddcolor_p=[1, RF, 1.0, 0.0, True])
    out.get_frame(0)                                    # first batch: engine warm, pipeline primed
    t0 = time.perf_counter()
    acc = 0
    for i in range(batch, n):
        acc += int(np.asarray(out.get_frame(i)[0])[0, 0])
    dt = time.perf_counter() - t0
    return {"value": (n - batch) / dt, "unit": "frames/s", "frames": n - batch,
            "how": "HAVC_colorizer(method=0) on a VapourSynth-stand-in clip, sequential get_frame() from one host thread"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import pipeline_oracle, synth_weights
    torch.set_num_threads(threads)
    sd = synth_weights.make_unet_state_dict("wide", 1234)
    clip = synth_clip(args.steps + args.warmup, H1080, W1080, seed=0)
    fr = lambda i: np.ascontiguousarray(np.transpose(clip[i], (1, 2, 0)))
    for i in range(args.warmup):
        pipeline_oracle.havc_colorizer_frame(sd, fr(i), RF)
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipeline_oracle.havc_colorizer_frame(sd, fr(args.warmup + i), RF)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": "1080p colorized frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1, "weights": "synthetic seed 1234 (reference state-dict schema)",
                   "sample": "each step = 1 frame of the clip on the host cores (bounded sample of the same workload)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} frames of the 1080p clip, 1 frame per step, torch CPU fp32 oracle port of the reference path"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything any library prints to fd 1 while the bench runs (NCCL's version banner,
    warnings) is sent to stderr instead, and emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("HAVC_BENCH_BATCH", "32")),
                    help="frames per step and per CUDA-graph launch (32 frames = 16 GB of activations of the 180 GB)")
    ap.add_argument("--dtype", default=os.environ.get("HAVC_BENCH_DTYPE", "fp16"), choices=["fp16", "bf16"])
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the bounded CPU-baseline sample (0 = skip)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--plugin-frames", type=int, default=128,
                    help="frames rendered through HAVC_colorizer on a VapourSynth-stand-in clip for the plugin_surface figure (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "HAVC_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["HAVC_NCCL_DEBUG"]
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = f"cuda:{local}"
    torch.cuda.set_device(dev)

    from oracle import synth_weights   # weight generator only (test infrastructure; not on the timed path)
    from vsdeoldify_b200 import _lib
    from vsdeoldify_b200.engine import DeoldifyEngine

    dtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    B, K, Wm = args.batch, args.steps, args.warmup
    sd = synth_weights.make_unet_state_dict("wide", 1234)
    eng = DeoldifyEngine(sd, W1080, H1080, render_factor=RF, batch=B, dtype=dtype, device=dev, use_graph=not args.no_graph)
    lib = _lib.lib()

    # each rank owns a contiguous block of the clip (block partition, no collective)
    n_host_batches = 4
    clip = synth_clip(n_host_batches * B, H1080, W1080, seed=100 + rank)
    host_batches = [np.ascontiguousarray(clip[i * B:(i + 1) * B]) for i in range(n_host_batches)]
    dev_batches = [torch.from_numpy(b).to(dev) for b in host_batches]
    pinned_batches = [torch.from_numpy(b).pin_memory() for b in host_batches]   # e2e inputs live in pinned host memory

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    def device_step(i):
        s = i % eng.n_slots
        with torch.cuda.stream(eng.compute):
            eng.d_in[s].copy_(dev_batches[i % n_host_batches], non_blocking=True)   # D2D: stage the next resident batch
        eng.run_slot(s)

    for i in range(Wm):
        device_step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.compute)
    for i in range(K):
        device_step(Wm + i)
    e1.record(eng.compute)
    barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_dev = float(tmax.item())

    # ---------------- end to end through the host API (e2e) ----------------
    sink = {"n": 0, "sum": 0}

    def on_result(i, out):
        sink["n"] += out.shape[0]
        sink["sum"] += int(out[0, 0, 0, 0])     # touch the result on the host

    eng.colorize_stream((pinned_batches[i % n_host_batches] for i in range(Wm)), on_result)
    barrier()
    t0 = time.perf_counter()
    eng.colorize_stream((pinned_batches[i % n_host_batches] for i in range(K)), on_result)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    sampler.stop_flag.set()
    sampler.join(timeout=3)

    # ---------------- per-kernel timing of one step (roofline of the dominant kernel) ----------------
    roof, breakdown = None, None
    if rank == 0:
        evs = []
        with torch.cuda.stream(eng.compute):
            eng.d_in[0].copy_(dev_batches[0])
            for rep in range(2):               # second pass is the measured one
                evs = []
                for op in eng.prog.ops:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(eng.compute)
                    op.fn(eng.compute.cuda_stream)
                    b.record(eng.compute)
                    evs.append((op, a, b))
        eng.compute.synchronize()
        gemm_ms = sum(a.elapsed_time(b) for op, a, b in evs if op.kind == "gemm")
        aux_ms = sum(a.elapsed_time(b) for op, a, b in evs if op.kind != "gemm")
        gemm_flops = sum(op.flops for op, a, b in evs if op.kind == "gemm")
        n_gemm = sum(1 for op, a, b in evs if op.kind == "gemm")
        peaks = load_peaks()
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM)", "achieved": ach,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"], "traffic": load_traffic(B),
                "peak_source": f"{peaks['src']} sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": n_gemm, "avg_launch_ms": gemm_ms / max(n_gemm, 1),
                "algorithmic_gflop_per_frame": gemm_flops / B / 1e9}
        top = sorted(((a.elapsed_time(b), op.name, op.flops) for op, a, b in evs), reverse=True)[:8]
        breakdown = {"gemm_ms_per_step": gemm_ms, "aux_ms_per_step": aux_ms,
                     "top": [{"op": n, "ms": round(t, 4), "tflops": round(f / (t * 1e-3) / 1e12, 1) if f else None} for t, n, f in top]}

    # ---------------- CPU baseline (rank 0, N = 1 only, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and args.cpu_frames > 0:
        threads = os.cpu_count() or 1
        fps_cpu, times = cpu_reference_fps(args.cpu_frames, threads)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_frames} frames of the same 1080p workload (after 1 warm-up frame), median per-frame wall clock, "
                         "torch CPU fp32 oracle port of the reference path"}

    # ---------------- the same path through the plugin surface (rank 0, N = 1 only; informational) ----------------
    plugin = None
    if rank == 0 and world == 1 and args.plugin_frames > 0:
        try:
            plugin = plugin_surface_fps(sd, clip, args.plugin_frames, B)
        except Exception as e:      # never lose the bench line over the informational figure
            plugin = {"error": str(e)[:200]}

    if rank == 0:
        frames = world * B * K
        value = frames / (ms_dev * 1e-3)
        e2e = frames / t_e2e
        line = {
            "metric": "1080p colorized frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if dtype == torch.float16 else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_step": B, "weights": "synthetic seed 1234 (reference state-dict schema)",
                       "cache": "inputs+activations per step (>1 GB) exceed the 126 MB L2; 4 distinct input batches rotated",
                       "partition": "contiguous frame blocks per rank, no collective"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * 3 * H1080 * W1080, "d2h_bytes_per_step": B * 3 * H1080 * W1080},
            "gpu_launches": eng.launches_per_batch * K,
            "clocks": sampler.summary(),
            "roofline": roof, "breakdown": breakdown, "cpu_baseline": cpu, "plugin_surface": plugin,
            "tensor_frac_whole_step": (GFLOP_PER_FRAME_SURVEY * 1e9 * value / world) / (load_peaks()["tf_sustained"] * 1e12),
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
