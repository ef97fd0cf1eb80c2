"""The reference-facing plugin surface: HAVC_main / HAVC_colorizer / HAVC_deoldify / HAVC_ddeoldify.

Same names, argument order, defaults and error behaviour as the reference for the per-frame colorization
path (vsdeoldify/__init__.py:101-109 HAVC_main, :2290-2523 HAVC_colorizer, :3612-3628 HAVC_ddeoldify); clips in,
clips out, frame properties passed through bit-exactly (every output frame is `f.copy()` of the input frame with
its planes overwritten, like vsslib/vsutils.py:92-95), scene-change gating as vsslib/vsmodels.py:221-224.

What differs is underneath: frames are not colourised one by one in a `std.ModifyFrame` callback that calls
torch; the returned clip pulls B consecutive source frames at a time and hands them to the per-GPU
`DeoldifyEngine` (one CUDA graph of libhavc_b200 kernels per batch).  Out-of-order frame requests are served
from a small batch cache, so any access pattern yields the same frames.

Not built (raise `vs.Error` instead of silently doing something else): DDColor (external `vsddcolor`, method 1 /
ddcolor models 0-1), the exemplar models, scene detection itself (`sc_threshold`/`sc_min_freq` > 0 require
the `_SceneChangePrev` props to be present on the input clip already), CPU mode (device_index=99).
"""
from __future__ import annotations

import math
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import vs_shim
from .constants import (DEF_ALM_p, DEF_ARTISTIC_WEIGHT, DEF_CMC_p, DEF_CRT_p, DEF_LMM_p, DEF_STABLE_WEIGHT, DEF_THT_BLACK,
                        DEF_THT_WHITE, DEF_TWEAK_p)

vs = vs_shim.get_vs()

package_dir = os.path.dirname(os.path.realpath(__file__))
model_dir = os.path.join(package_dir, "models")

_WEIGHT_FILES = {0: "ColorizeVideo_gen", 1: "ColorizeStable_gen", 2: "ColorizeArtistic_gen"}
# torch.hub file names of the Zhang checkpoints (eccv16.py:105-107, siggraph17.py:168-170), looked up under
# <torch_dir>/checkpoints like model_zoo.load_url does after torch.hub.set_dir(torch_dir) (vsdeoldify/__init__.py:2489-2490)
_ZHANG_FILES = {"siggraph17": "siggraph17-df00044c", "eccv16": "colorization_release_v2-9b330a0b"}
_REGISTERED: Dict[str, Dict[str, torch.Tensor]] = {}
# frames per engine step (one CUDA-graph launch): the benched operating point (bench.py: B = 32 fills the 148 SMs on the 12 x 12 /
# 24 x 24 layers; 32 frames = 16 GB of activations of the 180 GB).  HAVC_B200_BATCH overrides it.
_BATCH = int(os.environ.get("HAVC_B200_BATCH", "32"))
_DTYPE = {"fp16": torch.float16, "bf16": torch.bfloat16}[os.environ.get("HAVC_B200_DTYPE", "fp16")]


# host threads for the big plane copies of the clip adapters (numpy releases the GIL while it copies)
# (one process per GPU: share the host's cores between the local ranks)
_LOCAL_RANKS = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1))
_ZERO_COPY = os.environ.get("HAVC_B200_ZERO_COPY", "1") != "0"          # A/B switch
_COPY_POOL = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) // (2 * _LOCAL_RANKS))))


def HAVC_LogMessage(level: int, *args):
    """vsslib/vsutils.py:42-47: level EXCEPTION raises vs.Error, anything else goes to the VapourSynth log."""
    text = " ".join(map(str, args))
    if level == "EXCEPTION":
        raise vs.Error(text)
    vs.core.log_message(int(level), text)


def _raise(text: str):
    raise vs.Error(text)


def register_state_dict(weights_name: str, sd: Dict[str, torch.Tensor]):
    """Provide weights in memory (tests / synthetic weights) instead of `<models>/<weights_name>.pth`."""
    _REGISTERED[weights_name] = sd


def load_state_dict(weights_name: str, models_dir: str = model_dir) -> Dict[str, torch.Tensor]:
    """Learner.load semantics (vsdeoldify/fastai/basic_train.py:264-286): `<dir>/<name>.pth`, either
    {'model': sd, 'opt': ...} or a bare state-dict."""
    if weights_name in _REGISTERED:
        return _REGISTERED[weights_name]
    path = os.path.join(models_dir, weights_name + ".pth")
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        _raise("HAVC_colorizer: model files have not been downloaded.")      # vsdeoldify/__init__.py:2477
    state = torch.load(path, map_location="cpu", weights_only=True)
    return state["model"] if isinstance(state, dict) and "model" in state else state


def _output_frame(src_frame, planes, out_format=None):
    """Output frame = copy of the source frame (every prop survives, vsslib/vsutils.py:92-95) carrying the result planes; a GRAY
    source comes back as a YUV420P8 frame with the source's props (restore_format, havc_utils.py:208-222)."""
    if out_format is not None and getattr(src_frame.format, "id", None) != getattr(out_format, "id", out_format):
        return vs_shim.VideoFrame([np.ascontiguousarray(p) for p in planes], out_format, dict(src_frame.props))
    if hasattr(src_frame, "with_planes"):             # the stand-in adopts the result planes (views of a buffer that is ours)
        return src_frame.with_planes(list(planes))
    g = src_frame.copy()
    for p, pl in enumerate(planes):
        np.copyto(np.asarray(g[p]), pl)
    return g


def _native_format(clip) -> bool:
    """RGB24, YUV420P8 (BT.709 / BT.601) or GRAY8: the formats the engines convert on the device."""
    fid = getattr(clip.format, "id", clip.format)
    ids = [getattr(getattr(vs, n, None), "id", getattr(vs, n, None)) for n in ("RGB24", "YUV420P8", "GRAY8") if hasattr(vs, n)]
    if fid not in ids:
        return False
    if fid == ids[0]:
        return True
    return int(clip.get_frame(0).props.get("_Matrix", 1)) in (1, 2, 5, 6)


def _clip_format(clip, who: str):
    """(engine fmt, matrix, out_limited, output format) of a clip, the way convert_format_RGB24 reads them (havc_utils.py:57-164):
    RGB24 as is; 8-bit YUV 4:2:0 and GRAY are converted on the device (matrix from frame 0's _Matrix, default BT.709; input
    treated as limited range; the output range is the clip's _ColorRange, default limited); anything else raises."""
    fid = getattr(clip.format, "id", clip.format)
    ids = {getattr(getattr(vs, n, None), "id", getattr(vs, n, None)): n for n in ("RGB24", "YUV420P8", "GRAY8") if hasattr(vs, n)}
    name = ids.get(fid)
    if name == "RGB24":
        return "rgb24", "709", True, None
    if name not in ("YUV420P8", "GRAY8"):
        _raise(f"{who}: only RGB24, YUV420P8 and GRAY8 clips are handled by the B200 build (convert_format_RGB24 for other formats "
               "needs zimg's generic depth / subsampling conversions)")
    props = clip.get_frame(0).props
    m = int(props.get("_Matrix", 1))
    if m in (1, 2):                                   # BT.709 / unspecified -> BT.709 (_matrixIsInvalid, havc_utils.py:76-77)
        matrix = "709"
    elif m in (5, 6):                                 # 470bg / 170m
        matrix = "601"
    else:
        _raise(f"{who}: _Matrix = {m} is not handled by the B200 build (BT.709, BT.601)")
    out_limited = int(props.get("_ColorRange", 1)) != 0                                   # 0 = full, 1 = limited (default)
    return ("yuv420p8" if name == "YUV420P8" else "gray8"), matrix, out_limited, vs.YUV420P8


def _engine_or_smaller_batch(cls, *args, batch: int, **kw):
    """Build an engine; when the device runs out of memory (activations scale with batch x render_factor^2) halve the batch
    down to 1, then give up (None): the caller applies the reference's out-of-memory convention."""
    b = batch
    while True:
        try:
            return cls(*args, batch=b, **kw)
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            if b == 1:
                return None
            b = max(1, b // 2)


class _ColorizedClip:
    """frame_fn of the output clip: batches source frames through the engine, caches results by frame number.

    With an engine that has submit() / collect() the clip reads ahead: while the batch that holds frame n is on the GPU the
    source frames of the NEXT batch are already fetched, uploaded and queued, so a sequential reader (the normal VapourSynth
    access pattern) keeps the device busy; any other access pattern still gets the same frames, batch by batch."""

    def __init__(self, clip, engine, scenechange: bool, batch: int, run=None, out_format=None):
        self.clip, self.engine = clip, engine
        self.out_format = out_format                      # None: the source clip's format (frames are copies of the source frames)
        self.run = run if run is not None else engine.colorize_batch
        self.async_ok = run is None and hasattr(engine, "submit")
        self.scenechange, self.B = scenechange, batch
        self.cache: "OrderedDict[int, object]" = OrderedDict()
        self.pending = None                               # read-ahead job: (first frame, source frames, ticket)
        self.lock = threading.Lock()
        self._bufs: List[np.ndarray] = []                 # recycled result arrays [B, 3, H, W] (see _result_buf)

    def _planes(self, f) -> np.ndarray:
        return np.stack([np.asarray(f[p]) for p in range(3)])

    def _skip(self, n: int, srcs):
        """vsslib/vsmodels.py:221-224: with scene-change gating only frames with _SceneChangePrev == 1 (and frame 0) are
        colourised.  The reference runs SceneDetect itself (vsdeoldify/__init__.py:2496-2499), so the prop always exists there;
        a frame without it is an error here (scene detection is outside this build), never a silently uncoloured frame."""
        if not self.scenechange:
            return None
        from .sharded import scene_skip_flags
        try:
            return scene_skip_flags(n, srcs)
        except KeyError as e:
            _raise("HAVC_colorizer: " + str(e.args[0]))

    def _fetch(self, n: int):
        n1 = min(n + self.B, self.clip.num_frames)
        srcs = [self.clip.get_frame(i) for i in range(n, n1)]
        batch = np.stack([self._planes(f) for f in srcs])
        return srcs, batch, self._skip(n, srcs)

    def _store(self, n: int, srcs, out):
        planes_of = getattr(self.engine, "out_planes", None)
        for i, f in zip(range(n, n + len(srcs)), srcs):
            planes = planes_of(out, i - n) if planes_of is not None else [out[i - n, p] for p in range(3)]
            self.cache[i] = _output_frame(f, planes, self.out_format)
        while len(self.cache) > 4 * self.B:
            self.cache.popitem(last=False)

    def _start(self, n: int):
        """Fetch the source frames of the batch that starts at n straight into the engine's pinned input buffer and queue it."""
        n1 = min(n + self.B, self.clip.num_frames)
        srcs = [self.clip.get_frame(i) for i in range(n, n1)]
        buf = self.engine.next_input()
        planes_of = getattr(self.engine, "in_planes", None)

        def put(j):
            dst = planes_of(buf, j) if planes_of is not None else [buf[j, p] for p in range(3)]
            for p, d in enumerate(dst):
                np.copyto(d, np.asarray(srcs[j][p]))
        list(_COPY_POOL.map(put, range(len(srcs))))
        return (n, srcs, self.engine.submit(None, skip=self._skip(n, srcs), n=len(srcs)))

    def _result_buf(self) -> np.ndarray:
        """A result array nobody references any more (the frames we hand out are views of these arrays and keep them alive
        through `.base`), else a new one: allocating 200 MB per batch costs far more in page faults than the copy."""
        for b in self._bufs:
            if sys.getrefcount(b) <= 3:                   # the list, the loop variable, getrefcount's argument
                return b
        b = np.empty((self.B, 3, self.clip.height, self.clip.width), np.uint8)
        if len(self._bufs) < 8:
            self._bufs.append(b)
        return b

    def _finish(self, job):
        n, srcs, ticket = job
        if _ZERO_COPY and hasattr(self.engine, "collect_view"):   # frames adopt views of the pinned download buffer: no host copy
            self._store(n, srcs, self.engine.collect_view(ticket))
        else:
            self._store(n, srcs, self.engine.collect(ticket, out=self._result_buf(), pool=_COPY_POOL))

    def __call__(self, n: int):
        try:
            return self._render(n)
        except RuntimeError as e:
            # deoldify/filters.py:55-63: an out-of-memory RuntimeError inside a frame is a warning and the frame comes back
            # uncoloured; anything else propagates
            if "memory" not in str(e).lower():
                raise
            vs.core.log_message(vs.MESSAGE_TYPE_WARNING, "Warning: render_factor was set too high, and out of memory error resulted. "
                                                         "Returning original image.")
            torch.cuda.empty_cache()
            self.pending = None
            return self.clip.get_frame(n).copy()

    def _render(self, n: int):
        with self.lock:
            if n in self.cache:
                return self.cache[n]
            if not self.async_ok:
                srcs, batch, skip = self._fetch(n)
                self._store(n, srcs, self.run(batch, skip=skip))
                return self.cache[n]
            job, self.pending = self.pending, None
            if job is not None and not (job[0] <= n < job[0] + len(job[1])):
                self._finish(job)                         # a read-ahead nobody asked for yet: keep its frames, free its slot
                job = None
            if job is None:
                job = self._start(n)
            nxt = job[0] + len(job[1])
            if nxt < self.clip.num_frames and nxt not in self.cache:
                self.pending = self._start(nxt)           # queued behind `job`; its upload overlaps job's compute
            self._finish(job)
            return self.cache[n]


def HAVC_colorizer(
        clip, method: int = 2, mweight: float = 0.4, deoldify_p: Sequence = (0, 24, 1.0, 0.0),
        ddcolor_p: Sequence = (1, 24, 1.0, 0.0, True), ddtweak: Sequence[bool] = (False, False, False),
        ddtweak_p: Sequence = (DEF_TWEAK_p, "300:360|0.8,0.1"),
        cmc_p: Sequence = DEF_CMC_p, lmm_p: Sequence = DEF_LMM_p, alm_p: Sequence = DEF_ALM_p,
        crt_p: Sequence = DEF_CRT_p, cmb_sw: bool = False, sc_threshold: float = 0.0, sc_tht_offset: int = 1,
        sc_min_freq: int = 0, sc_tht_ssim: float = 0.0, sc_normalize: bool = False, sc_min_int: int = 1,
        sc_tht_white: float = DEF_THT_WHITE, sc_tht_black: float = DEF_THT_BLACK, device_index: int = 0,
        torch_dir: str = model_dir, debug_level: int = 0):
    """Drop-in for vsdeoldify.HAVC_colorizer (vsdeoldify/__init__.py:2290-2523) on the DeOldify path."""
    if device_index == 99 or (isinstance(device_index, (list, tuple)) and 99 in device_index):
        _raise("HAVC_colorizer: CPU mode (device_index=99) is not available in the B200 build (no CPU fallback)")
    if not torch.cuda.is_available():
        _raise("HAVC_colorizer: CUDA is not available")                                   # :2441
    if vs is not vs_shim and not _native_format(clip):
        # real VapourSynth, a format the device conversion does not cover (10-bit, 4:2:2, 4:4:4, RGB48, ...): VapourSynth's own
        # convert_format_RGB24 / restore_format around the RGB24 path, exactly what the reference does (:2493, :2523)
        rgb, restore = convert_format_RGB24(clip, "HAVC_colorizer")
        return restore(HAVC_colorizer(rgb, method, mweight, deoldify_p, ddcolor_p, ddtweak, ddtweak_p, cmc_p, lmm_p, alm_p, crt_p, cmb_sw,
                                      sc_threshold, sc_tht_offset, sc_min_freq, sc_tht_ssim, sc_normalize, sc_min_int, sc_tht_white,
                                      sc_tht_black, device_index, torch_dir, debug_level))
    fmt, matrix, out_limited, out_format = _clip_format(clip, "HAVC_colorizer")
    if sc_threshold < 0:
        _raise("HAVC_colorizer: sc_threshold must be >= 0")                              # :2447
    if sc_min_freq < 0:
        _raise("HAVC_colorizer: sc_min_freq must be >= 0")                               # :2450
    merge_weight = 0.0 if method == 0 else (1.0 if method == 1 else mweight)              # :2452-2462
    if merge_weight == 0.0:
        method = 0
    elif merge_weight == 1.0:
        method = 1
    deoldify_model, deoldify_rf, deoldify_sat, deoldify_hue = deoldify_p[:4]
    ddcolor_model, ddcolor_rf = ddcolor_p[0], ddcolor_p[1]
    # device_index: the reference takes ONE device (GPU0...GPU7, deoldify/device_id.py:3-12).  Extension of the B200 build: a
    # list / tuple of device indices (or HAVC_B200_DEVICES="0,1,..." with the default device_index) renders the clip on
    # several GPUs of the box, frames partitioned over them and returned in order (vsdeoldify_b200/sharded.py).
    devices = list(device_index) if isinstance(device_index, (list, tuple)) else [device_index]
    if len(devices) == 1 and devices[0] == 0 and os.environ.get("HAVC_B200_DEVICES"):
        devices = [int(t) for t in os.environ["HAVC_B200_DEVICES"].split(",") if t.strip() != ""]
    if not devices or any((not isinstance(d, int)) or d > 7 or d < 0 for d in devices) or len(set(devices)) != len(devices):
        _raise("HAVC_colorizer: wrong device_index, choices are: GPU0...GPU7, CPU=99")    # :2480
    if max(devices) >= torch.cuda.device_count():
        _raise(f"HAVC_colorizer: device_index {max(devices)} but only {torch.cuda.device_count()} CUDA device(s) are visible")
    if ddcolor_rf != 0 and ddcolor_rf not in range(10, 65):
        _raise("HAVC_colorizer: ddcolor render_factor must be between: 10-64")            # :2483
    ddcolor_sat = ddcolor_p[2] if len(ddcolor_p) > 2 else 1.0
    ddcolor_hue = ddcolor_p[3] if len(ddcolor_p) > 3 else 0.0
    if method not in (0, 1, 2, 3, 4, 5, 6, 7):
        _raise("HAVC: only dd_method in (0,6) is supported")                              # mcomb.py:192
    if method != 0 and ddcolor_model not in (2, 3):
        _raise("HAVC_colorizer: DDColor (ddcolor_p model 0/1, the external vsddcolor package) is out of scope of the B200 "
               "build; use model 2 (Zhang siggraph17) or 3 (Zhang eccv16) as the second colour model")
    ddtweak = list(ddtweak) if isinstance(ddtweak, (list, tuple)) else [bool(ddtweak), False, False]
    ddtweak = (ddtweak + [False, False, False])[:3]
    scenechange = not (sc_threshold == 0 and sc_min_freq == 0)                            # :2494
    hue_adjust, tweak = "none", None
    if method != 0:                                                                       # vsmodels.py:296-336
        if ddtweak[1] or ddtweak[2]:
            _raise("HAVC_colorizer: the denoise / retinex pre-filters of the second model are not built")
        tw = list(ddtweak_p[0]) if len(ddtweak_p) == 2 else list(ddtweak_p[:8])
        hue_adjust = (ddtweak_p[1] if len(ddtweak_p) == 2 else (ddtweak_p[8] if len(ddtweak_p) > 8 else "none")).lower()
        if ddtweak[0]:
            bright, cont, gamma, constrained, luma_min, gamma_luma_min, gamma_alpha, gamma_min = tw[:8]
            if not constrained:
                _raise("HAVC_colorizer: ddtweak without luma_constrained_tweak applies gamma through vs_tweak / image_tweak "
                       "(std.Levels; the reference's image_tweak gamma LUT raises) - not built")
            if (bright != 0 or cont != 1) and not scenechange:
                _raise("HAVC_colorizer: ddtweak bright/cont without scene-change detection goes through vs_tweak (zimg YUV420 "
                       "round trip) - not built")
            tweak = dict(bright=bright, cont=cont, luma_min=luma_min, gamma=gamma, gamma_luma_min=gamma_luma_min,
                         gamma_alpha=gamma_alpha, gamma_min=gamma_min)
    if ddcolor_rf == 0:
        ddcolor_rf = min(max(math.trunc(0.4 * clip.width / 16), 16), 32)                  # :2492
    scenechange = not (sc_threshold == 0 and sc_min_freq == 0)                            # :2494
    # frame_size: HAVC_colorizer:2502 uses max(ddcolor_rf, deoldify_rf) even when method == 0
    frame_size = min(max(ddcolor_rf, deoldify_rf) * 16, clip.width)
    from .engine import DeoldifyEngine
    mdir = torch_dir or model_dir
    run_deoldify = method != 1                                            # vs_sc_deoldify returns None for method 1 (vsmodels.py:198)
    sd_video = load_state_dict(_WEIGHT_FILES[0], mdir) if run_deoldify else None   # the video generator always runs (visualize.py:120)
    sd_other = load_state_dict(_WEIGHT_FILES[deoldify_model], mdir) if (run_deoldify and deoldify_model in (1, 2)) else None
    weight = {1: DEF_STABLE_WEIGHT, 2: DEF_ARTISTIC_WEIGHT}.get(deoldify_model, 0.0)
    zhang = merge = None
    if method != 0:
        zname = "siggraph17" if ddcolor_model == 2 else "eccv16"                           # vsmodels.py:339-344
        zhang = (zname, load_state_dict(_ZHANG_FILES[zname], os.path.join(mdir, "checkpoints")))
        merge = dict(method=method, weight=merge_weight, cmc_p=list(cmc_p), lmm_p=list(lmm_p), alm_p=list(alm_p),
                     crt_p=list(crt_p), invert=bool(cmb_sw))
    from .filters import FilterError
    try:
        engines = [_engine_or_smaller_batch(DeoldifyEngine, sd_video, clip.width, clip.height, render_factor=deoldify_rf, frame_size=frame_size,
                                  batch=_BATCH, dtype=_DTYPE, device=f"cuda:{d}", sd_other=sd_other, video_weight=weight,
                                  zhang=zhang, merge=merge, hue_adjust=hue_adjust, run_deoldify=run_deoldify, ddtweak=tweak,
                                  sat=(deoldify_sat, ddcolor_sat), hue=(deoldify_hue, ddcolor_hue), fmt=fmt, matrix=matrix,
                                  out_limited=out_limited)
                   for d in devices]
    except (ValueError, FilterError) as e:
        _raise("HAVC_colorizer: " + str(e))
    if any(e is None for e in engines):
        # the reference's out-of-memory convention (deoldify/filters.py:55-63): warn and hand back uncoloured frames
        vs.core.log_message(vs.MESSAGE_TYPE_WARNING, "Warning: render_factor was set too high, and out of memory error resulted. "
                                                     "Returning original image.")
        return clip
    batch = min(e.B for e in engines)
    for e in engines:
        e.prepare_async()                 # pin the result-buffer pools before the first frame is requested
    if len(engines) == 1:
        fn = _ColorizedClip(clip, engines[0], scenechange, batch, out_format=out_format)
    else:
        from .sharded import ShardedRenderer
        renderer = ShardedRenderer(clip, engines, batch, scenechange, partition=os.environ.get("HAVC_B200_PARTITION", "interleaved"),
                                   copy_pool=_COPY_POOL, make_frame=lambda f, planes: _output_frame(f, planes, out_format))

        def fn(n, renderer=renderer):
            try:
                return renderer(n)
            except KeyError as e:                    # a frame without the scene-detection prop
                _raise("HAVC_colorizer: " + str(e.args[0]))
    return vs_shim.VideoNode(clip.num_frames, clip.width, clip.height, out_format or clip.format, fn, clip.fps_num, clip.fps_den) \
        if vs is vs_shim else _wrap_real_vs(clip, fn, out_format)


_SC_PROPS = ('_SceneChangePrev', '_SceneChangeNext', 'sc_threshold', 'sc_frequency', 'sc_luma', 'sc_ratio')    # CopySCDetect, vsscdect.py:118-120


class _MergedClip:
    """frame_fn of HAVC_merge's output clip: pulls B consecutive frames of the clips, merges them on the GPU.
    Without clip_luma the output frames are copies of clipa's frames (the mcomb.py selectors return f[0].copy()); with clip_luma
    they are copies of clip_luma's frames (_clip_chroma_resize -> vs_recover_clip_luma) that take the scene-detection props of
    the merged / selected clip (CopySCDetect)."""

    def __init__(self, clipa, clipb, engine, batch, method, weight, cmc_p, lmm_p, alm_p, crt_p, clip_luma=None, sc_src=None):
        self.clipa, self.clipb, self.clip_luma, self.engine, self.B = clipa, clipb, clip_luma, engine, batch
        self.sc_src = sc_src
        self.args = (method, weight, cmc_p, lmm_p, alm_p, crt_p)
        self.cache: "OrderedDict[int, object]" = OrderedDict()
        self.lock = threading.Lock()

    def __call__(self, n: int):
        with self.lock:
            if n in self.cache:
                return self.cache[n]
            base = self.clip_luma if self.clip_luma is not None else self.clipa
            n1 = min(n + self.B, base.num_frames)
            stack = lambda fs: np.stack([np.stack([np.asarray(f[p]) for p in range(3)]) for f in fs])
            get = lambda c: [c.get_frame(i) for i in range(n, n1)] if c is not None else None
            fa, fb = get(self.clipa), get(self.clipb)
            if self.clip_luma is None:
                out = self.engine.merge_batch(stack(fa), stack(fb), *self.args)
                fbase, fsc = fa, None
            else:
                fl = get(self.clip_luma)
                out = self.engine.merge_batch(stack(fa) if fa else None, stack(fb) if fb else None, stack(fl), *self.args)
                fbase, fsc = fl, get(self.sc_src)
            for k, (i, f) in enumerate(zip(range(n, n1), fbase)):
                g = f.copy()                              # every prop of the base frame survives
                for p in range(3):
                    np.copyto(np.asarray(g[p]), out[i - n, p])
                if fsc is not None:
                    for key in _SC_PROPS:
                        if key in fsc[k].props:
                            g.props[key] = fsc[k].props[key]
                self.cache[i] = g
            while len(self.cache) > 4 * self.B:
                self.cache.popitem(last=False)
            return self.cache[n]


def HAVC_merge(clipa=None, clipb=None, clip_luma=None, weight: float = 0.5, method: int = 2, cmc_p: Sequence = DEF_CMC_p,
               lmm_p: Sequence = DEF_LMM_p, alm_p: Sequence = DEF_ALM_p, crt_p: Sequence = DEF_CRT_p, device_index: int = 0):
    """Drop-in for vsdeoldify.HAVC_merge (vsdeoldify/__init__.py:2536-2675) on RGB24 clips: method 2 = std.Merge,
    methods 3-7 = the vsslib merges of vs_combine_models, `clip_luma` = the Spline64 squeeze + _clip_chroma_resize detour of
    :2633-2673.  clipa / clipb in other formats go through convert_format_RGB24 and the result is restored to clipa's format
    (:2651-2652, :2675), like the reference."""
    for name, c in (("clipa", clipa), ("clipb", clipb), ("clip_luma", clip_luma)):
        if c is not None and not hasattr(c, "get_frame"):
            _raise("HAVC_merge: this is not a clip: " + name)                            # :2624-2631
    rgb24 = getattr(vs.RGB24, "id", vs.RGB24)
    shortcut = method == 0 or weight == 0 or method == 1 or weight == 1
    # :2651-2652: the merge proper converts clipa / clipb with convert_format_RGB24 and restores clipa's format at the end; the
    # shortcuts (:2633-2649) and clip_luma (_clip_chroma_resize) use the clips as they are: RGB24
    for name, c in (("clipa", clipa), ("clipb", clipb), ("clip_luma", clip_luma)):
        if c is not None and (shortcut or name == "clip_luma") and getattr(c.format, "id", c.format) != rgb24:
            _raise("HAVC_merge: " + name + " must be RGB24 here (only the merge proper converts clipa / clipb, vsdeoldify/__init__.py:2651)")
    restore = lambda c: c
    if not shortcut:
        clipa, restore = convert_format_RGB24(clipa, "HAVC_merge", device_index)
        clipb, _ = convert_format_RGB24(clipb, "HAVC_merge", device_index)
    from .filters import FilterError, LumaMergeEngine, MergeEngine
    args = (list(cmc_p), list(lmm_p), list(alm_p), list(crt_p))
    mk = lambda fn, base: (vs_shim.VideoNode(base.num_frames, base.width, base.height, base.format, fn, base.fps_num, base.fps_den)
                           if vs is vs_shim else _wrap_real_vs(base, fn))

    def guarded(fn):
        def call(n):
            try:
                return fn(n)
            except FilterError as e:
                _raise("HAVC_merge: " + str(e))
        return call
    if method == 0 or weight == 0 or method == 1 or weight == 1:                         # :2633-2645
        sel = clipa if (method == 0 or weight == 0) else clipb
        if clip_luma is None:
            return sel
        if not torch.cuda.is_available():
            _raise("HAVC_merge: CUDA is not available")
        eng = LumaMergeEngine((sel.width, sel.height), (clip_luma.width, clip_luma.height), False, batch=_BATCH,
                              device=f"cuda:{device_index}")
        fn = _MergedClip(sel, None, eng, _BATCH, 0, 0.0, *args, clip_luma=clip_luma, sc_src=sel)
        return mk(guarded(fn), clip_luma)
    if (clipa.width, clipa.height) != (clipb.width, clipb.height):
        _raise("HAVC_merge: clipa and clipb must have the same size")
    if not torch.cuda.is_available():
        _raise("HAVC_merge: CUDA is not available")
    if method not in (2, 3, 4, 5, 6, 7):
        _raise("HAVC: only dd_method in (0,6) is supported")                              # mcomb.py:192
    if method == 2 or clip_luma is None:                                                  # :2648-2650: method 2 ignores clip_luma
        engine = MergeEngine(clipa.width, clipa.height, batch=_BATCH, device=f"cuda:{device_index}")
        fn = _MergedClip(clipa, clipb, engine, _BATCH, method, weight, *args)
        return restore(mk(guarded(fn), clipa))
    eng = LumaMergeEngine((clipa.width, clipa.height), (clip_luma.width, clip_luma.height), True, batch=_BATCH,
                          device=f"cuda:{device_index}")
    fn = _MergedClip(clipa, clipb, eng, _BATCH, method, weight, *args, clip_luma=clip_luma, sc_src=clipa)
    return restore(mk(guarded(fn), clip_luma))


_COLORMAP_NAMES = ['none', 'blue->brown', 'blue->red', 'blue->green', 'green->brown', 'green->red', 'green->blue', 'redrose->brown',
                   'redrose->blue', "red->brown", 'red->blue', 'yellow->rose']
_COLORMAP_HUE = ["none", "180:280|+140", "180:280|+100", "180:280|+220", "80:180|+260", "80:180|+220", "80:180|+140",
                 "300:360,0:20|+40", "300:360,0:20|+260", "320:360|+50", "300:360|+260", "30:90|+300"]
_COLORMAP_W = ["1.0", "0.90", "0.80", "0.75"]


def _get_colormap(ColorMap: str = "red->brown", ColorTune: str = "light") -> str:
    """havc_utils._get_colormap (havc_utils.py:552-581): a colour-map name -> its "chroma adjustment" string; anything
    else must already be a valid chroma adjustment and is returned as is."""
    if ColorTune not in _COLOR_TUNE:
        _raise("HAVC_main: ColorTune choice is invalid for '" + ColorTune + "'")
    cm = ColorMap.lower()
    if cm in _COLORMAP_NAMES:
        return _COLORMAP_HUE[_COLORMAP_NAMES.index(cm)] + "," + _COLORMAP_W[_COLOR_TUNE.index(ColorTune)]
    from .filters import parse_hue_adjust
    if parse_hue_adjust(cm) is None:
        _raise("HAVC_main: ColorMap choice is invalid for '" + cm + "'")
    return cm


class _ConvertedClip:
    """frame_fn of a format-converted clip under the VapourSynth stand-in: batches of frames through engine.FormatEngine
    (to_rgb = convert_format_RGB24, from_rgb = restore_format); props are carried over from the source frames."""

    def __init__(self, clip, engine, to_rgb: bool, props_from=None):
        self.clip, self.engine, self.to_rgb, self.B = clip, engine, to_rgb, engine.B
        self.props_from = props_from if props_from is not None else clip
        self.cache: "OrderedDict[int, object]" = OrderedDict()
        self.lock = threading.Lock()

    def __call__(self, n: int):
        with self.lock:
            if n in self.cache:
                return self.cache[n]
            n0 = (n // self.B) * self.B
            idx = range(n0, min(n0 + self.B, self.clip.num_frames))
            frames = [self.clip.get_frame(i) for i in idx]
            if self.to_rgb:
                rgb = self.engine.to_rgb(frames)
                outs = [vs_shim.VideoFrame([rgb[j, p] for p in range(3)], vs_shim.RGB24, dict(f.props)) for j, f in enumerate(frames)]
            else:
                planes = self.engine.from_rgb(np.stack([np.stack([np.asarray(f[p]) for p in range(3)]) for f in frames]))
                outs = [vs_shim.VideoFrame(planes[j], vs_shim.YUV420P8, dict(self.props_from.get_frame(i).props)) for j, i in enumerate(idx)]
            for i, f in zip(idx, outs):
                self.cache[i] = f
            while len(self.cache) > 4 * self.B:
                self.cache.popitem(last=False)
            return self.cache[n]


def convert_format_RGB24(clip, who: str = "convert_format_RGB24", device_index: int = 0):
    """Drop-in for havc_utils.convert_format_RGB24 (vsdeoldify/havc_utils.py:57-164, chroma_resize=False): (RGB24 clip, restore),
    `restore(rgb_clip)` being restore_format (:167-237) for this clip.  Real VapourSynth: VapourSynth's own resize.Bicubic with
    the reference's arguments (every format the reference handles).  Stand-in: YUV420P8 / GRAY8 on the device (zimg restated)."""
    fid = getattr(clip.format, "id", clip.format)
    if fid == getattr(vs.RGB24, "id", vs.RGB24):
        return clip, (lambda c: c)
    if vs is not vs_shim:
        fmt = clip.format
        props = clip.get_frame(0).props
        v74 = vs.core.core_version.release_major >= 74
        rkey = "_Range" if v74 else "_ColorRange"
        matrix = props.get("_Matrix", int(vs.MATRIX_BT709))
        if int(matrix) == 2:                                                               # unspecified (_matrixIsInvalid, :76-77)
            matrix = int(vs.MATRIX_BT709)
            clip = clip.std.SetFrameProps(_Matrix=matrix)
        full = int(props.get(rkey, 1)) == 0
        c = clip
        if fmt.bits_per_sample != 8:
            c = vs.core.resize.Bicubic(c, format=fmt.replace(bits_per_sample=8))           # :126-127
        if fmt.color_family == vs.YUV:
            c = vs.core.resize.Bicubic(c, format=vs.RGB24, matrix_in=matrix, range_in_s="limited", range_s="full",
                                       dither_type="error_diffusion")                      # :133-143
        elif fmt.color_family == vs.GRAY:
            c = vs.core.resize.Bicubic(c, format=vs.RGB24, range_in_s="limited", range_s="full")   # :145-151
        else:
            c = vs.core.resize.Bicubic(c, format=vs.RGB24, range_s="full")                 # :152-157
        c = c.std.SetFrameProps(**{rkey: 0})                                               # :160-163 (RANGE_FULL)

        def restore(rgb):                                                                  # restore_format, :167-237
            rs = "full" if full else "limited"
            if fmt.color_family == vs.YUV:
                return vs.core.resize.Bicubic(rgb, format=fmt.id, matrix_in=int(vs.MATRIX_BT709), matrix=matrix, range_in_s="full",
                                              range_s=rs, dither_type="error_diffusion")
            if fmt.color_family == vs.GRAY:
                return vs.core.resize.Bicubic(rgb, format=vs.YUV420P8, matrix=int(vs.MATRIX_BT709), range_in_s="full", range_s=rs,
                                              dither_type="error_diffusion")
            return vs.core.resize.Bicubic(rgb, format=fmt.id, range_in_s="full", range_s=rs)
        return c, restore
    from .engine import FormatEngine
    fmt, matrix, out_limited, out_format = _clip_format(clip, who)
    if not torch.cuda.is_available():
        _raise(f"{who}: CUDA is not available")
    try:
        eng = FormatEngine(fmt, clip.width, clip.height, batch=min(_BATCH, 8), device=f"cuda:{device_index}", matrix=matrix,
                           out_limited=out_limited)
    except ValueError as e:
        _raise(f"{who}: " + str(e))
    rgb = vs_shim.VideoNode(clip.num_frames, clip.width, clip.height, vs_shim.RGB24, _ConvertedClip(clip, eng, True), clip.fps_num, clip.fps_den)

    def restore(c):
        return vs_shim.VideoNode(c.num_frames, c.width, c.height, out_format, _ConvertedClip(c, eng, False), c.fps_num, c.fps_den)
    return rgb, restore


class _BatchedClip:
    """frame_fn that renders frames in aligned batches of B through `render(first, count) -> list of frames` and caches them."""

    def __init__(self, num_frames: int, B: int, render):
        self.num_frames, self.B, self.render = num_frames, B, render
        self.cache: "OrderedDict[int, object]" = OrderedDict()
        self.lock = threading.Lock()

    def __call__(self, n: int):
        with self.lock:
            if n not in self.cache:
                n0 = (n // self.B) * self.B
                for i, f in enumerate(self.render(n0, min(self.B, self.num_frames - n0))):
                    self.cache[n0 + i] = f
                while len(self.cache) > 4 * self.B:
                    self.cache.popitem(last=False)
            return self.cache[n]


_SC_COPY_PROPS = ['_SceneChangePrev', '_SceneChangeNext', 'sc_threshold', 'sc_frequency', 'sc_luma', 'sc_ratio']     # vsresize.py:124-125


def _min_hw_size(width: int, height: int, min_size=(512, 480)):
    """Target size of resize_min_HW (vsslib/vsresize.py:30-101), or None when the clip is not resized."""
    if height < width:
        if height <= min_size[1]:
            return None
        w = round(width * min_size[1] / height)
        return (w - 1 if w % 2 else w), min_size[1]                                          # resize_to_height, :52-75
    if width <= min_size[0]:
        return None
    h = round(height * min_size[0] / width)
    return min_size[0], (h + 1 if h % 2 else h)                                              # resize_to_width, :77-99


def _stack_planes(frames):
    return np.stack([np.stack([np.asarray(f[p]) for p in range(3)]) for f in frames])


def resize_min_HW(clip, device_index: int = 0):
    """Drop-in for vsresize.resize_min_HW (vsslib/vsresize.py:30-50) on RGB24 clips: Spline36 to height 480 (landscape) / width
    512 (portrait) with the aspect ratio kept and even sizes; smaller clips pass through."""
    size = _min_hw_size(clip.width, clip.height)
    if size is None:
        return clip
    if vs is not vs_shim:
        return clip.resize.Spline36(width=size[0], height=size[1])
    from .filters import ResizeEngine
    eng = ResizeEngine(clip.width, clip.height, size[0], size[1], batch=min(_BATCH, 8), device=f"cuda:{device_index}")

    def render(n0, cnt):
        src = [clip.get_frame(i) for i in range(n0, n0 + cnt)]
        out = eng.down(_stack_planes(src))
        return [vs_shim.VideoFrame([out[j, p] for p in range(3)], vs_shim.RGB24, dict(f.props)) for j, f in enumerate(src)]
    return vs_shim.VideoNode(clip.num_frames, size[0], size[1], clip.format, _BatchedClip(clip.num_frames, eng.B, render), clip.fps_num, clip.fps_den)


def resize_to_chroma(clip_highres, clip_lowres, device_index: int = 0):
    """Drop-in for vsresize.resize_to_chroma (vsslib/vsresize.py:101-127) on RGB24 clips: `clip_lowres` resized to the size of
    `clip_highres` with Spline36, both to YUV420P8 (BT.709, full range), the Y plane of `clip_highres` with the chroma of the
    resized clip, back to RGB24 with error-diffusion dither; the scene-change props come from `clip_lowres`."""
    if vs is not vs_shim:
        c = clip_lowres
        if (clip_highres.width, clip_highres.height) != (c.width, c.height):
            c = c.resize.Spline36(width=clip_highres.width, height=clip_highres.height)
        bw = clip_highres.resize.Bicubic(format=vs.YUV420P8, matrix_s="709", range_s="full")
        col = c.resize.Bicubic(format=vs.YUV420P8, matrix_s="709", range_s="full")
        yuv = vs.core.std.ShufflePlanes(clips=[bw, col, col], planes=[0, 1, 2], colorfamily=vs.YUV)
        yuv = yuv.std.CopyFrameProps(prop_src=col, props=_SC_COPY_PROPS)
        return yuv.resize.Bicubic(format=vs.RGB24, matrix_in_s="709", range_s="full", dither_type="error_diffusion")
    from .filters import FilterError, ResizeEngine
    try:
        eng = ResizeEngine(clip_highres.width, clip_highres.height, clip_lowres.width, clip_lowres.height, batch=min(_BATCH, 8),
                           device=f"cuda:{device_index}")
    except FilterError as e:
        _raise(str(e))

    def render(n0, cnt):
        hi = [clip_highres.get_frame(i) for i in range(n0, n0 + cnt)]
        lo = [clip_lowres.get_frame(i) for i in range(n0, n0 + cnt)]
        out = eng.chroma(_stack_planes(hi), _stack_planes(lo))
        frames = []
        for j, (fh, fl) in enumerate(zip(hi, lo)):
            props = dict(fh.props)
            props.update({k: fl.props[k] for k in _SC_COPY_PROPS if k in fl.props})
            frames.append(vs_shim.VideoFrame([out[j, p] for p in range(3)], vs_shim.RGB24, props))
        return frames
    return vs_shim.VideoNode(clip_highres.num_frames, clip_highres.width, clip_highres.height, vs_shim.RGB24,
                             _BatchedClip(clip_highres.num_frames, eng.B, render), clip_highres.fps_num, clip_highres.fps_den)


class _TemporalClip:
    """frame_fn of a temporally filtered clip (scope row N3): batches of B consecutive frames go to the engine together with
    their `nh` halo frames on either side, clamped to the clip's ends the way std.AverageFrames clamps its requests
    (min(max(n + d, 0), last)); source frames are cached so that the halo shared by neighbouring batches is fetched once."""

    def __init__(self, clip, engine, scene_weights: bool):
        self.clip, self.engine, self.B, self.nh = clip, engine, engine.out_B, engine.nh
        self.scene_weights = scene_weights
        self.src: "OrderedDict[int, object]" = OrderedDict()
        self.cache: "OrderedDict[int, object]" = OrderedDict()
        self.lock = threading.Lock()

    def _src(self, i: int):
        f = self.src.get(i)
        if f is None:
            f = self.src[i] = self.clip.get_frame(i)
            while len(self.src) > 2 * (self.B + 2 * self.nh):
                self.src.popitem(last=False)
        return f

    def __call__(self, n: int):
        with self.lock:
            if n in self.cache:
                return self.cache[n]
            last = self.clip.num_frames - 1
            n0 = (n // self.B) * self.B
            idx = [min(max(i, 0), last) for i in range(n0 - self.nh, n0 + self.B + self.nh)]
            frames = [self._src(i) for i in idx]
            seq = np.stack([np.stack([np.asarray(f[p]) for p in range(3)]) for f in frames])
            weights = None
            if self.scene_weights:                    # vs_clip_color_stabilizer: std.AverageFrames(scenechange=True), vsfilters.py:58
                from .filters import scene_folded_weights
                wl, K = self.engine.temporal.wl, 2 * self.nh + 1
                weights = np.array([scene_folded_weights(wl, [int(frames[b + k].props.get("_SceneChangePrev", 0)) for k in range(K)],
                                                         [int(frames[b + k].props.get("_SceneChangeNext", 0)) for k in range(K)])
                                    for b in range(self.B)], np.int32)
            out = self.engine.process_sequence(seq, n0, weights)
            for b in range(min(self.B, last + 1 - n0)):
                self.cache[n0 + b] = _output_frame(frames[self.nh + b], [out[b, p] for p in range(3)])
            while len(self.cache) > 4 * self.B:
                self.cache.popitem(last=False)
            return self.cache[n]


def _node_like(clip, fn):
    return vs_shim.VideoNode(clip.num_frames, clip.width, clip.height, clip.format, fn, clip.fps_num, clip.fps_den) \
        if vs is vs_shim else _wrap_real_vs(clip, fn)


def _stab_params(nframes, mode, sat, tht, weight, tht_scen, hue_adjust) -> dict:
    return dict(nframes=int(nframes), mode=mode, sat=float(sat), tht=int(tht), weight=float(weight), tht_scen=float(tht_scen),
                hue_adjust=(hue_adjust or "none").lower())


def vs_chroma_stabilizer_ex(clip, nframes: int = 5, mode: str = "A", sat: float = 1.0, tht: int = 0, weight: float = 0.5,
                            tht_scen: float = 0.8, hue_adjust: str = 'none', algo: int = 0, device_index: int = 0):
    """Drop-in for vsslib.vsfilters.vs_chroma_stabilizer_ex (vsslib/vsfilters.py:84-115) on RGB24 clips, algo = 0 (the value
    HAVC_stabilizer passes; algo = 1, the ModifyFrame variant, raises): the temporal chroma stabiliser, scope row N3."""
    if algo != 0:
        _raise("vs_chroma_stabilizer_ex: algo=1 (_average_frames_ex) is not built; HAVC_stabilizer uses algo=0")
    if not torch.cuda.is_available():
        _raise("vs_chroma_stabilizer_ex: CUDA is not available")
    clip, restore = convert_format_RGB24(clip, "vs_chroma_stabilizer_ex", device_index)       # the reference's function is RGB24 only
    from .filters import FilterError, TemporalEngine
    try:
        engine = TemporalEngine(clip.width, clip.height, batch=min(_BATCH, 8), device=f"cuda:{device_index}",
                                **_stab_params(nframes, mode, sat, tht, weight, tht_scen, hue_adjust))
    except (ValueError, FilterError) as e:
        _raise(str(e))
    return restore(_node_like(clip, _TemporalClip(clip, engine, scene_weights=int(tht) == 0)))


def _reduce_flicker(clip, strength: int = 2, aggressive: int = 0):
    """vs_reduce_flicker (vsslib/vsplugins.py:263-272): the external ReduceFlicker VapourSynth plugin (`core.rdfl`) that ends
    HAVC_stabilizer's temporal stage.  It is not part of the reference tree (a binary the reference loads from its plugin
    directory), so it is called when the host provides it and raises the reference's error otherwise."""
    rdfl = getattr(vs.core, "rdfl", None)
    try:
        if rdfl is None:
            raise RuntimeError("no plugin with the namespace 'rdfl' is loaded")
        return rdfl.ReduceFlicker(clip=clip, strength=strength, aggressive=aggressive)
    except Exception as error:
        _raise("vs_retinex: plugin 'ReduceFlicker.dll' not properly loaded/installed -> " + str(error))     # vsplugins.py:270


def HAVC_stabilizer(clip, dark: bool = False, dark_p: Sequence = (0.2, 0.8), smooth: bool = False,
                    smooth_p: Sequence = (0.3, 0.7, 0.9, 0.0, "none"), stab: bool = False,
                    stab_p: Sequence = (5, 'A', 1, 15, 0.2, 0.8), colormap: str = "none", render_factor: int = 24,
                    device_index: int = 0):
    """Drop-in for vsdeoldify.HAVC_stabilizer (vsdeoldify/__init__.py:2748-2873) for its per-frame stages: Spline64 squeeze
    to render_factor*16, vs_dark_tweak (`dark`), vs_chroma_bright_tweak (`smooth`), vs_colormap (`colormap`), then
    _clip_chroma_resize back to the clip's size with the original luma.  `stab=True` adds the temporal chroma stabiliser
    (vs_chroma_stabilizer_ex, row N3 of the scope table) on the squeezed clip, followed - as in the reference - by the external
    ReduceFlicker plugin (`core.rdfl`), which must be provided by the host (the reference's error is raised without it)."""
    if render_factor != 0 and render_factor not in range(16, 65):
        _raise("HAVC_stabilizer: render_factor must be between: 16-64")                   # :2796
    if stab and getattr(vs.core, "rdfl", None) is None:
        _reduce_flicker(clip)                         # raises at graph-build time, like the reference without the plugin (:2861)
    if not torch.cuda.is_available():
        _raise("HAVC_stabilizer: CUDA is not available")
    clip, restore = convert_format_RGB24(clip, "HAVC_stabilizer", device_index)           # :2787 (restore_format at :2871)
    if render_factor == 0:
        render_factor = min(max(math.trunc(0.4 * clip.width / 16), 16), 32)               # :2799
    frame_size = min(render_factor * 16, clip.width)                                      # :2803
    cm = colormap.lower()
    colormap_adjust = _get_colormap(cm) if cm not in ("none", "") else "none"             # :2827-2832
    stages = dict(dark=bool(dark), dark_p=list(dark_p), smooth=bool(smooth), smooth_p=list(smooth_p), colormap_adjust=colormap_adjust)
    from .filters import FilterError, StabilizerEngine
    stab_kw = None
    if stab:                                                                              # :2835-2845 (stab_algo = 0)
        stab_kw = _stab_params(stab_p[0], stab_p[1], stab_p[2], stab_p[3], stab_p[4], stab_p[5], stab_p[6] if len(stab_p) > 6 else "none")
    try:
        engine = StabilizerEngine(clip.width, clip.height, frame_size, stages, batch=min(_BATCH, 16) if stab else _BATCH,
                                  device=f"cuda:{device_index}", stab=stab_kw)
    except (ValueError, FilterError) as e:
        _raise("HAVC_stabilizer: " + str(e))
    if stab:
        return restore(_reduce_flicker(_node_like(clip, _TemporalClip(clip, engine, scene_weights=stab_kw["tht"] == 0))))
    return restore(_node_like(clip, _ColorizedClip(clip, engine, False, _BATCH, run=engine.process_batch)))


class ModelImageRender:
    """Drop-in for vsdeoldify.deoldify.visualize.ModelImageRender (deoldify/visualize.py:41-137): the reference's own
    per-image entry point (`get_transformed_image(PIL image) -> PIL image`; BASELINE cfg1 runs it on still images).
    The filter-level Pillow BILINEAR squeeze / un-squeeze and the full-resolution luma transplant run on the GPU
    (engine.ImageRenderEngine); one engine per image size is kept."""

    def __init__(self, package_dir: Optional[str] = None, modelname: str = 'video', render_factor: int = 24,
                 video_weight: float = 0, device_index: int = 0):
        self.package_dir = package_dir
        self._modelname, self._video_weight, self._render_factor = modelname, video_weight, render_factor
        mdir = os.path.join(package_dir, "models") if package_dir else model_dir             # generators.py:18-19
        self._sd_video = load_state_dict(_WEIGHT_FILES[0], mdir)
        other = {"stable": 1, "artistic": 2}.get(modelname)
        self._sd_other = load_state_dict(_WEIGHT_FILES[other], mdir) if other is not None else None
        self._device = f"cuda:{device_index}"
        self._engines: Dict[tuple, object] = {}

    def get_transformed_image(self, img_orig, post_process: bool = True):
        from PIL import Image
        from .engine import ImageRenderEngine
        if not post_process:
            _raise("ModelImageRender: post_process=False is not built")
        arr = np.asarray(img_orig.convert("RGB") if img_orig.mode != "RGB" else img_orig)
        H, W = arr.shape[:2]
        eng = self._engines.get((W, H))
        if eng is None:
            eng = self._engines[(W, H)] = ImageRenderEngine(self._sd_video, W, H, self._render_factor, batch=1, dtype=_DTYPE,
                                                            device=self._device, sd_other=self._sd_other,
                                                            video_weight=self._video_weight)
        out = eng.render_batch(np.ascontiguousarray(np.transpose(arr, (2, 0, 1)))[None])
        return Image.fromarray(np.ascontiguousarray(np.transpose(out[0], (1, 2, 0))), "RGB")


def _wrap_real_vs(clip, fn, out_format=None):
    """Real VapourSynth: serve frames through std.ModifyFrame; the selector ignores `f` and returns our frame.  When the output
    format differs from the clip's (a GRAY clip comes back as YUV420P8) the node is built on a blank clip of that format."""
    base = clip
    if out_format is not None and clip.format.id != out_format:
        base = vs.core.std.BlankClip(clip, format=out_format)
    return base.std.ModifyFrame(clips=[base], selector=lambda n, f: fn(n))


def HAVC_deoldify(clip, model: int = 0, render_factor: int = 24, sat: float = 1.0, hue: float = 0.0, **kw):
    """DeOldify only: HAVC_colorizer(method=0) (vsdeoldify/__init__.py:2452-2462).  The name is in the north star;
    the reference tree has no function of this name."""
    return HAVC_colorizer(clip, method=0, deoldify_p=[model, render_factor, sat, hue], **kw)


def HAVC_ddeoldify(
        clip, method: int = 2, mweight: float = 0.4, deoldify_p: Sequence = (0, 24, 1.0, 0.0),
        ddcolor_p: Sequence = (1, 24, 1.0, 0.0, True), ddtweak: bool = False,
        ddtweak_p: Sequence = (DEF_TWEAK_p, "300:360|0.8,0.1"),
        cmc_tresh: float = 0.2, lmm_p: Sequence = (0.2, 0.8, 1.0), alm_p: Sequence = (0.8, 1.0, 0.15), cmb_sw: bool = False,
        sc_threshold: float = 0.0, sc_tht_offset: int = 1, sc_min_freq: int = 0, sc_tht_ssim: float = 0.0,
        sc_normalize: bool = False, sc_min_int: int = 1, sc_tht_white: float = DEF_THT_WHITE,
        sc_tht_black: float = DEF_THT_BLACK, device_index: int = 0, torch_dir: str = model_dir, sc_debug: bool = False):
    """Deprecated positional alias, same mapping as vsdeoldify/__init__.py:3612-3628."""
    vs.core.log_message(vs.MESSAGE_TYPE_WARNING,
                        "Warning: HAVC_ddeoldify is deprecated and may be removed in the future, please use 'HAVC_colorizer' instead.")
    debug_level = 2 if sc_debug else 0
    return HAVC_colorizer(clip, method, mweight, deoldify_p, ddcolor_p, [ddtweak, False, False], ddtweak_p, [cmc_tresh], lmm_p,
                          alm_p, DEF_CRT_p, cmb_sw, sc_threshold, sc_tht_offset, sc_min_freq, sc_tht_ssim, sc_normalize,
                          sc_min_int, sc_tht_white, sc_tht_black, device_index, torch_dir, debug_level)


# ---- deprecated names the reference still exports (vsdeoldify/__init__.py:3631-3664): same forwarding, same warnings ----------
def _deprecated(old: str, new: str):
    vs.core.log_message(vs.MESSAGE_TYPE_WARNING,
                        f"Warning: {old} is deprecated and may be removed in the future, please use '{new}' instead.")


def ddeoldify_main(clip, Preset: str = 'Fast', VideoTune: str = 'Stable', ColorFix: str = 'Violet/Red', ColorTune: str = 'Light',
                   ColorMap: str = 'None', degrain_strength: int = 0, enable_fp16: bool = True):
    _deprecated("ddeoldify_main", "HAVC_main")
    return HAVC_main(clip=clip, Preset=Preset, VideoTune=VideoTune, ColorFix=ColorFix, ColorTune=ColorTune, ColorMap=ColorMap,
                     enable_fp16=enable_fp16)


def ddeoldify(clip, method: int = 2, mweight: float = 0.4, deoldify_p: Sequence = (0, 24, 1.0, 0.0),
              ddcolor_p: Sequence = (1, 24, 1.0, 0.0, True), dotweak: bool = False,
              dotweak_p: Sequence = (0.0, 1.0, 1.0, False, 0.2, 0.5, 1.5, 0.5), ddtweak: bool = False,
              ddtweak_p: Sequence = (DEF_TWEAK_p, "300:360|0.8,0.1"), degrain_strength: int = 0, cmc_tresh: float = 0.2,
              lmm_p: Sequence = (0.2, 0.8, 1.0), alm_p: Sequence = (0.8, 1.0, 0.15), cmb_sw: bool = False, device_index: int = 0,
              torch_dir: str = model_dir):
    _deprecated("ddeoldify", "HAVC_colorizer")
    return HAVC_colorizer(clip, method, mweight, deoldify_p, ddcolor_p, [ddtweak, False, False], ddtweak_p, [cmc_tresh], lmm_p, alm_p,
                          DEF_CRT_p, cmb_sw, sc_threshold=0, sc_min_freq=0, device_index=device_index, torch_dir=torch_dir)


def ddeoldify_stabilizer(clip, dark: bool = False, dark_p: Sequence = (0.2, 0.8), smooth: bool = False,
                         smooth_p: Sequence = (0.3, 0.7, 0.9, 0.0, "none"), stab: bool = False,
                         stab_p: Sequence = (5, 'A', 1, 15, 0.2, 0.80), colormap: str = "none", render_factor: int = 24):
    _deprecated("ddeoldify_stabilizer", "HAVC_stabilizer")
    return HAVC_stabilizer(clip, dark, dark_p, smooth, smooth_p, stab, stab_p, colormap, render_factor)


# ---- HAVC_main: preset tables of vsdeoldify/havc_utils.py:335-581 (values, not code) -----------------------------------
_PRESETS = ['placebo', 'veryslow', 'slower', 'slow', 'medium', 'fast', 'faster', 'veryfast']
_PRESET_RF = [32, 32, 32, 28, 24, 22, 20, 16]
_DEOLDIFY_MODELS = ["video", "stable", "artistic"]
_DDCOLOR_MODELS = ["modelscope", "artistic", "siggraph17", "eccv16"]
_VIDEO_TUNE = {'verystable': 0.2, 'morestable': 0.3, 'stable': 0.4, 'balanced': 0.5, 'vivid': 0.6, 'morevivid': 0.7, 'veryvivid': 0.8}
_COMB_METHOD = {'simple': 2, 'constrained-chroma': 3, 'luma-masked': 4, 'adaptive-luma': 5, 'chroma-retention': 6,
                'chromabound adaptive': 7}
_COLOR_TUNE = ['none', 'light', 'medium', 'strong']
_COLOR_FIX = ['none', 'magenta', 'magenta/violet', 'violet', 'violet/red', 'blue/magenta', 'yellow', 'yellow/orange', 'yellow/green',
              'retinex/red']
_HUE_FIX = ["none", "270:300", "250:360", "300:330", "300:360", "220:280", "60:90", "30:90", "60:120", "none"]


def _get_render_factor(Preset: str) -> int:
    try:
        return _PRESET_RF[_PRESETS.index(Preset.lower())]
    except ValueError:
        _raise("HAVC_main: Preset choice is invalid for '" + str(Preset) + "'")


def _get_color_model(ColorModel: str):
    """havc_utils._get_color_model (havc_utils.py:402-441) -> (deoldify model, second model, method)."""
    cm = ColorModel.lower()
    try:
        if '+' in cm:
            a, b = cm.split("+")
            return _DEOLDIFY_MODELS.index(a), _DDCOLOR_MODELS.index(b), 2
        if "deoldify" in cm:
            return _DEOLDIFY_MODELS.index(cm.replace("deoldify", "").replace("(", "").replace(")", "")), 0, 0
        if "ddcolor" in cm or "zhang" in cm:
            name = cm.replace("ddcolor", "").replace("zhang", "").replace("(", "").replace(")", "")
            return 0, _DDCOLOR_MODELS.index(name), 1
    except ValueError:
        pass
    _raise("HAVC_main: ColorModel choice is invalid for '" + ColorModel + "'")


def _get_color_tune(ColorTune: str, ColorFix: str, dd_model: int):
    """The ddtweak / hue-range part of havc_utils._get_color_tune (havc_utils.py:451-517)."""
    tune = (ColorTune or "none").lower()
    fix = (ColorFix or "none").lower()
    if tune not in _COLOR_TUNE:
        _raise("HAVC_main: ColorTune choice is invalid for '" + tune + "'")
    if fix not in _COLOR_FIX:
        _raise("HAVC_main: ColorFix choice is invalid for '" + fix + "'")
    tn, co = _COLOR_TUNE.index(tune), _COLOR_FIX.index(fix)
    hue_tune = {0: ["1.0,0.0", "0.7,0.1", "0.5,0.1", "0.2,0.1"], 2: ["1.0,0.0", "0.6,0.1", "0.4,0.2", "0.2,0.1"],
                3: ["1.0,0.0", "0.7,0.1", "0.6,0.1", "0.3,0.1"]}.get(dd_model, ["1.0,0.0", "0.8,0.1", "0.5,0.1", "0.2,0.1"])
    dd_tweak = [False, False, False]
    if tn == 0:
        return dd_tweak, "none"
    if co == 0:
        return [True, True, False], "none"
    if co == 9:
        return [True, False, True], _HUE_FIX[4] + "|" + hue_tune[2]
    return [True, False, False], _HUE_FIX[co] + "|" + hue_tune[tn]


def HAVC_main(clip, Preset: str = 'Medium', FrameInterp: int = 0, ColorModel: str = 'Video+Artistic', CombMethod: str = 'Simple',
              VideoTune: str = 'Stable', ColorFix: str = 'Magenta/Violet', ColorTune: str = 'Light', ColorMap: str = 'None',
              ColorTemp: str = 'None', BlackWhiteTune: str = 'None', BlackWhiteMode: int = 0, BlackWhiteBlend: bool = True,
              EnableDeepEx: bool = False, DeepExMethod: int = 0, DeepExPreset: str = 'Medium', DeepExRefMerge: int = 0,
              DeepExOnlyRefFrames: bool = False, ScFrameDir: Optional[str] = None, ScThreshold: float = 0.10, ScThtOffset: int = 1,
              ScMinFreq: int = 0, ScMinInt: int = 1, ScThtSSIM: float = 0.0, ScNormalize: bool = False, DeepExModel: int = 0,
              DeepExVivid: bool = True, DeepExEncMode: int = 0, DeepExMaxMemFrames=0, RefRange: Sequence[int] = (0, 0),
              enable_fp16: bool = True, debug_level: int = 0, device_index=0):
    """Drop-in for vsdeoldify.HAVC_main (vsdeoldify/__init__.py:101-330 -> HAVC_main_presets :469-912): the reference's full
    positional signature in the reference's order (`device_index`, keyword only in spirit, is this build's one addition at the
    end), restricted to the per-frame image-model branch.  The string presets are turned into HAVC_colorizer's numeric
    arguments exactly as the reference does (colour-model split, CombMethod, VideoTune weight, ColorTune / ColorFix ->
    ddtweak + hue range, ColorMap -> "chroma adjustment"), followed by the HAVC_stabilizer step the presets append (:896-910):
    colormap only for the fast presets, dark + smooth + colormap for slower / slow / medium.  The DeepEx* / Sc* / RefRange
    parameters only feed the exemplar branch in the reference (they are not passed to HAVC_colorizer, :849-853) and are
    accepted and ignored here exactly when EnableDeepEx is False; EnableDeepEx / FrameInterp / the tiled presets / ColorTemp /
    BlackWhiteTune / DDColor models raise vs.Error.  Where the reference would also switch on the TEMPORAL chroma stabiliser
    (two-model presets with ColorTune != 'none') it runs when the host provides the ReduceFlicker plugin its chain ends in
    (`core.rdfl`); without the plugin the per-frame stages run and a warning says the temporal one is skipped."""
    rf = _get_render_factor(Preset)
    speed_id = _PRESETS.index(Preset.lower())
    if EnableDeepEx or FrameInterp != 0:
        _raise("HAVC_main: exemplar-based models are sequential and out of scope of the B200 build")
    if speed_id in (0, 1):
        _raise("HAVC_main: the tiled 'placebo' / 'veryslow' presets (HAVC_clip_slice) are not built")
    do_model, dd_model, dd_method = _get_color_model(ColorModel)
    if dd_method != 0 and dd_model in (0, 1):
        _raise("HAVC_main: ColorModel '" + ColorModel + "' needs DDColor (external vsddcolor package), which is out of scope of the "
               "B200 build; use 'DeOldify(...)', '<Video|Stable|Artistic>+Siggraph17', '...+ECCV16' or 'Zhang(...)'")
    if VideoTune.lower() not in _VIDEO_TUNE:
        _raise("HAVC_main: VideoTune choice is invalid for '" + VideoTune.lower() + "'")
    weight = _VIDEO_TUNE[VideoTune.lower()]
    if dd_method == 2:
        if CombMethod.lower() not in _COMB_METHOD:
            _raise("HAVC_main: CombMethod choice is invalid for '" + CombMethod + "'")
        dd_method = _COMB_METHOD[CombMethod.lower()]
    dd_tweak, hue_range = _get_color_tune(ColorTune, ColorFix, dd_model)
    for label, v in (("ColorTemp", ColorTemp), ("BlackWhiteTune", BlackWhiteTune)):
        if str(v).lower() != "none":
            _raise(f"HAVC_main: {label} post filters are not built (temporal / B&W tuning filters are outside the per-frame path)")
    tune = (ColorTune or "none").lower()
    chroma_adjust = "none" if str(ColorMap).lower() in ("none", "") else _get_colormap(str(ColorMap), tune)   # havc_utils.py:519-548
    # :492-494 chroma_resize = speed_id > 1 (every preset built here): the clip is reduced with resize_min_HW (Spline36 to height
    # 480 / width 512) before the colour models run and the result goes back through resize_to_chroma (luma of the full-size clip,
    # chroma of the result) inside restore_format (havc_utils.py:183-184).  Real VapourSynth: the reference's order (the reduction
    # happens in the clip's own format); stand-in: the clip is converted to RGB24 first (the reduction is built for RGB24 only)
    dev0 = device_index[0] if isinstance(device_index, (list, tuple)) else device_index
    if vs is vs_shim:
        high, restore = convert_format_RGB24(clip, "HAVC_main", dev0)
        low = resize_min_HW(high, dev0)
    else:
        high = clip
        low, restore = convert_format_RGB24(resize_min_HW(clip), "HAVC_main")
    finish = lambda c: restore(resize_to_chroma(high, c, dev0))
    clip_colored = HAVC_colorizer(low, method=dd_method, mweight=weight, deoldify_p=[do_model, rf, 1.0, 0.0],
                                  ddcolor_p=[dd_model, rf, 1.0, 0.0, enable_fp16], ddtweak=dd_tweak, ddtweak_p=[DEF_TWEAK_p, hue_range],
                                  device_index=device_index, debug_level=debug_level)
    if speed_id > 4:                     # 'fast', 'faster', 'veryfast': only the colormap (:896-897)
        return finish(HAVC_stabilizer(clip_colored, colormap=chroma_adjust, device_index=dev0))
    stab_enabled = dd_method != 0 and tune != "none"       # :903-906 stab=stab_enabled
    if stab_enabled and getattr(vs.core, "rdfl", None) is None:
        # the preset's temporal stage ends in the external ReduceFlicker plugin (vsplugins.py:263-272); without it the reference
        # raises - here the per-frame stages still run and the log says what was left out
        vs.core.log_message(vs.MESSAGE_TYPE_WARNING, "HAVC_main (B200 build): plugin 'ReduceFlicker' (core.rdfl) is not loaded; the "
                                                     "temporal chroma stabilizer of this preset (HAVC_stabilizer stab=True) is not "
                                                     "applied, its per-frame stages are")
        stab_enabled = False
    return finish(HAVC_stabilizer(clip_colored, dark=True, dark_p=[0.2, 0.8], colormap=chroma_adjust, smooth=True,
                                  smooth_p=[0.3, 0.7, 0.9, 0.0, "none"], stab=stab_enabled, stab_p=[5, 'A', 1, 15, 0.2, 0.8],
                                  device_index=dev0))
