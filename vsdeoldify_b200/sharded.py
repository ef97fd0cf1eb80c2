"""Multi-GPU rendering of ONE clip inside one process: one engine + one worker thread per GPU, the frame range partitioned
over the GPUs, frames handed back strictly by frame number (BASELINE north star: "frames are block-partitioned
across the 8 GPUs of one box with no collective, and only the output frames return to the host in order").

The reference has no multi-GPU path (vsdeoldify/__init__.py:27 pins GPU 0; device_index selects ONE device,
deoldify/device_id.py:3-12).  Every frame of the per-frame path is independent, so no data-path collective exists: a job is
a batch of B consecutive frames, a partition maps jobs to GPUs, and the collector is the only place that orders anything.

  partition = "block"        rank r owns the contiguous block partition.block_range(n_frames, r, world) - the north-star
                             layout.  An in-order consumer only gets parallelism if the results of the later blocks are
                             buffered until it reaches them, i.e. with `window=None` (render a whole clip into host memory).
  partition = "interleaved"  job k goes to GPU k % world: the streaming layout for VapourSynth's frame-by-frame consumers;
                             `window` jobs are in flight ahead of the frame last asked for (bounded host memory).

Engines are anything with `next_input() / submit(None, skip=, n=) / collect(ticket, out=, pool=)` (engine.DeoldifyEngine);
tests drive the scheduler with CPU stand-ins.
"""
from __future__ import annotations

import bisect
import sys
import threading
import time
from collections import OrderedDict, deque
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

import os

from . import partition as part

ZERO_COPY = os.environ.get("HAVC_B200_ZERO_COPY", "1") != "0"          # A/B switch


def plan_jobs(n_frames: int, world: int, batch: int, mode: str = "block") -> List[Tuple[int, int, int]]:
    """[(first frame, end frame, gpu)] covering [0, n_frames) in frame order."""
    if mode not in ("block", "interleaved"):
        raise ValueError(f"unknown partition {mode!r}")
    jobs: List[Tuple[int, int, int]] = []
    if mode == "block":
        for r in range(world):
            s, e = part.block_range(n_frames, r, world)
            jobs += [(i0, i1, r) for i0, i1 in part.batches(s, e, batch)]
    else:
        jobs = [(i0, i1, k % world) for k, (i0, i1) in enumerate(part.batches(0, n_frames, batch))]
    return jobs


class ShardedRenderer:
    """frame_fn of a clip rendered by several engines.  Thread-safe; any access pattern yields the same frames."""

    NEW, RUNNING, DONE = 0, 1, 2

    def __init__(self, clip, engines: Sequence, batch: int, scenechange: bool = False, partition: str = "interleaved",
                 window: Optional[int] = None, copy_pool=None, make_frame: Optional[Callable] = None,
                 require_scene_props: bool = True):
        self.clip, self.engines, self.B, self.scenechange = clip, list(engines), batch, scenechange
        self.world = len(self.engines)
        self.jobs = plan_jobs(clip.num_frames, self.world, batch, partition)
        self.starts = [j[0] for j in self.jobs]
        self.state = [self.NEW] * len(self.jobs)
        self.results: "OrderedDict[int, list]" = OrderedDict()
        self.error: Optional[BaseException] = None
        # jobs that may be rendered ahead of the most recently requested one; None = unbounded (whole-clip rendering: every
        # result stays in host memory until close()).  Streaming default: four jobs per GPU - with two, a worker that has just
        # submitted a batch finds no second job inside the window, blocks in collect() for the whole GPU time of the first
        # and never overlaps the next upload with it (measured: 35-45 ms per 23 ms batch).
        if window is None and partition == "interleaved":
            window = 4 * self.world
        self.window = window
        self.keep_behind = self.world                                              # finished jobs kept behind the cursor
        self.cursor = 0                                                            # job of the most recent request
        self.cv = threading.Condition()
        self.stop = False
        self.copy_pool = copy_pool
        self.make_frame = make_frame or _adopt_planes
        self.require_scene_props = require_scene_props
        self._bufs: List[List[np.ndarray]] = [[] for _ in range(self.world)]
        # per-worker wall-clock accounting (seconds): where a GPU's host thread spends its time
        self.stats = [dict(fetch=0.0, copy_in=0.0, submit=0.0, collect=0.0, frames=0.0, wait=0.0, jobs=0) for _ in range(self.world)]
        self.threads = [threading.Thread(target=self._worker, args=(r,), daemon=True, name=f"havc-gpu{r}") for r in range(self.world)]
        for t in self.threads:
            t.start()

    # ---- consumer side ------------------------------------------------------------------------------------
    def job_of(self, n: int) -> int:
        return bisect.bisect_right(self.starts, n) - 1

    def __call__(self, n: int):
        j = self.job_of(n)
        # fast path, no lock: the job of the previous request, already rendered (dict / list reads are atomic under the GIL).  With
        # eight workers contending for the condition variable a per-frame lock round trip limited ONE consumer thread to ~2 k frames/s.
        if j == self.cursor:
            res = self.results.get(j)
            if res is not None:
                return res[n - self.jobs[j][0]]
        with self.cv:
            moved = j != self.cursor
            self.cursor = j
            if self.state[j] == self.DONE and j not in self.results:      # evicted: render it again
                self.state[j] = self.NEW
                moved = True
            if moved:                             # once per job, not per frame: every wake-up costs each worker a scan of its window
                self.cv.notify_all()
            while self.state[j] != self.DONE and self.error is None:
                self.cv.wait(timeout=1.0)
            if self.error is not None:
                raise self.error
            return self.results[j][n - self.jobs[j][0]]

    def frames(self):
        """All frames in order (the in-order collector)."""
        for n in range(self.clip.num_frames):
            yield self(n)

    def close(self):
        with self.cv:
            self.stop = True
            self.cv.notify_all()
        for t in self.threads:
            t.join(timeout=10)

    # ---- worker side --------------------------------------------------------------------------------------
    def _pick(self, r: int) -> Optional[int]:
        """Next job of GPU r: the first NEW one at or after the cursor, inside the window."""
        hi = len(self.jobs) if self.window is None else min(len(self.jobs), self.cursor + self.window)
        for j in range(self.cursor, hi):
            if self.jobs[j][2] == r and self.state[j] == self.NEW:
                return j
        return None

    def _skip_flags(self, i0: int, srcs) -> Optional[np.ndarray]:
        if not self.scenechange:
            return None
        return scene_skip_flags(i0, srcs, self.require_scene_props)

    def _result_buf(self, r: int) -> np.ndarray:
        if self.window is None:                              # whole-clip rendering: every result stays alive, nothing to recycle
            return np.empty((self.B, 3, self.clip.height, self.clip.width), np.uint8)
        for b in self._bufs[r]:
            if sys.getrefcount(b) <= 3:
                return b
        b = np.empty((self.B, 3, self.clip.height, self.clip.width), np.uint8)
        if len(self._bufs[r]) < 8:
            self._bufs[r].append(b)
        return b

    def _worker(self, r: int):
        eng = self.engines[r]
        try:
            dev = getattr(eng, "dev", None)
            if dev is not None:
                import torch
                torch.cuda.set_device(dev)
            inflight: deque = deque()
            depth = getattr(eng, "n_slots", 1)
            while True:
                tw = time.perf_counter()
                with self.cv:
                    j = None
                    while not self.stop:
                        j = self._pick(r) if len(inflight) < depth else None
                        if j is not None or inflight:
                            break
                        self.cv.wait(timeout=1.0)
                    if self.stop:
                        return
                    if j is not None:
                        self.state[j] = self.RUNNING
                self.stats[r]["wait"] += time.perf_counter() - tw
                if j is not None:
                    i0, i1, _ = self.jobs[j]
                    t0 = time.perf_counter()
                    srcs = [self.clip.get_frame(i) for i in range(i0, i1)]
                    t1 = time.perf_counter()
                    buf = eng.next_input()
                    planes_in = getattr(eng, "in_planes", None)

                    def put(k, buf=buf, srcs=srcs, planes_in=planes_in):
                        dst = planes_in(buf, k) if planes_in is not None else [buf[k, p] for p in range(3)]
                        for p, d in enumerate(dst):
                            np.copyto(d, np.asarray(srcs[k][p]))
                    if self.copy_pool is not None:
                        list(self.copy_pool.map(put, range(len(srcs))))
                    else:
                        for k in range(len(srcs)):
                            put(k)
                    t2 = time.perf_counter()
                    ticket = eng.submit(None, skip=self._skip_flags(i0, srcs), n=len(srcs))
                    t3 = time.perf_counter()
                    st = self.stats[r]
                    st["fetch"] += t1 - t0; st["copy_in"] += t2 - t1; st["submit"] += t3 - t2; st["jobs"] += 1
                    inflight.append((j, srcs, ticket))
                    if len(inflight) < depth:
                        continue                      # try to queue a second batch behind it before waiting
                jd, srcs, ticket = inflight.popleft()
                t4 = time.perf_counter()
                # bounded window: frames adopt views of the pinned download buffer (no host copy of the result); whole-clip rendering
                # (window None) would pin the entire clip, so it copies into pageable arrays instead
                if ZERO_COPY and self.window is not None and hasattr(eng, "collect_view"):
                    out = eng.collect_view(ticket)
                else:
                    out = eng.collect(ticket, out=self._result_buf(r), pool=self.copy_pool)
                t5 = time.perf_counter()
                planes_out = getattr(eng, "out_planes", None)
                frames = [self.make_frame(f, planes_out(out, k) if planes_out is not None else [out[k, p] for p in range(3)])
                          for k, f in enumerate(srcs)]
                self.stats[r]["collect"] += t5 - t4
                self.stats[r]["frames"] += time.perf_counter() - t5
                with self.cv:
                    self.results[jd] = frames
                    self.state[jd] = self.DONE
                    if self.window is not None:
                        for old in [k for k in self.results if k < self.cursor - self.keep_behind]:
                            del self.results[old]
                    self.cv.notify_all()
        except BaseException as e:   # surface worker failures to the consumer instead of hanging it
            with self.cv:
                self.error = e
                self.cv.notify_all()


def scene_skip_flags(i0: int, srcs, require_props: bool = True) -> np.ndarray:
    """vsslib/vsmodels.py:221-224: with scene-change gating only frames whose `_SceneChangePrev` prop is 1 (and frame 0) are
    colourised.  The reference runs SceneDetect itself (vsdeoldify/__init__.py:2496-2499) so the prop always exists; scene
    DETECTION is outside the B200 build, so a frame without the prop is an error, not an uncoloured frame."""
    flags = np.zeros(len(srcs), bool)
    for k, f in enumerate(srcs):
        n = i0 + k
        if "_SceneChangePrev" not in f.props:
            if require_props:
                raise KeyError(f"frame {n} has no '_SceneChangePrev' property: scene-change gating (sc_threshold / sc_min_freq > 0) "
                               "needs a clip that went through scene detection (the B200 build does not run SceneDetect itself)")
            flags[k] = n != 0
        else:
            flags[k] = not (n == 0 or f.props["_SceneChangePrev"] == 1)
    return flags


def _adopt_planes(src_frame, planes):
    """Output frame = copy of the source frame (all props survive, vsslib/vsutils.py:92-95) with the result planes."""
    if hasattr(src_frame, "with_planes"):
        return src_frame.with_planes(list(planes))
    g = src_frame.copy()
    for p, pl in enumerate(planes):
        np.copyto(np.asarray(g[p]), pl)
    return g
