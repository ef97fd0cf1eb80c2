"""A minimal in-process stand-in for the part of the VapourSynth API the HAVC hot path touches
(SURVEY.md Appendix E), used when the real `vapoursynth` module cannot be imported (it is not installable
in the build/bench environment).  With real VapourSynth present, `get_vs()` returns it instead and the
plugin surface runs on real clips.

Only what the colorizer path needs is provided: VideoNode / VideoFrame with numpy planes and a props dict,
`std.ModifyFrame`, `std.SetFrameProp(s)`, `std.CopyFrameProps`, `std.BlankClip`, `vs.Error`, `core.log_message`
and a numpy-backed source clip.  Frame objects follow the reference's selector protocol
(vsslib/vsutils.py:60-110): `np.asarray(frame[plane])` is a writable view on copies, `frame.copy()` keeps props.
"""
from __future__ import annotations

import types
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

MESSAGE_TYPE_DEBUG, MESSAGE_TYPE_INFORMATION, MESSAGE_TYPE_WARNING, MESSAGE_TYPE_CRITICAL, MESSAGE_TYPE_FATAL = range(5)
RGB, GRAY, YUV = 2000000, 1000000, 3000000          # color families (ids are arbitrary but distinct)
MATRIX_BT709, RANGE_FULL, RANGE_LIMITED = 1, 0, 1


class Error(Exception):
    pass


@dataclass(frozen=True, eq=False)
class VideoFormat:
    id: int
    name: str
    color_family: int
    bits_per_sample: int
    num_planes: int
    subsampling_w: int = 0
    subsampling_h: int = 0

    # real VapourSynth: `clip.format.id == vs.RGB24` holds (the preset constant is an int enum); keep that true here
    def __eq__(self, other):
        if isinstance(other, VideoFormat):
            return self.id == other.id
        if isinstance(other, int):
            return self.id == other
        return NotImplemented

    def __hash__(self):
        return hash(self.id)

    def __int__(self):
        return self.id


RGB24 = VideoFormat(1, "RGB24", RGB, 8, 3)
GRAY8 = VideoFormat(2, "Gray8", GRAY, 8, 1)
YUV420P8 = VideoFormat(3, "YUV420P8", YUV, 8, 3, 1, 1)
YUV444P8 = VideoFormat(4, "YUV444P8", YUV, 8, 3)


class VideoFrame:
    def __init__(self, planes: Sequence[np.ndarray], fmt: VideoFormat, props: Optional[dict] = None, readonly=False):
        self._planes = [np.ascontiguousarray(p) for p in planes]
        self.format = fmt
        self.props = dict(props or {})
        self.height, self.width = self._planes[0].shape
        self.readonly = readonly
        if readonly:
            for p in self._planes:
                p.setflags(write=False)

    def __getitem__(self, plane: int) -> np.ndarray:
        return self._planes[plane]

    def copy(self) -> "VideoFrame":
        return VideoFrame([p.copy() for p in self._planes], self.format, dict(self.props))

    def get_read_array(self, plane):
        return self._planes[plane]

    def with_planes(self, planes: Sequence[np.ndarray]) -> "VideoFrame":
        """A new frame with this frame's format and (copied) props that ADOPTS the given plane arrays instead of copying and
        overwriting them: what `g = f.copy(); np.copyto(np.asarray(g[p]), plane)` yields, without the two plane copies.
        (Stand-in only: real VapourSynth frames are copy-on-write and are filled through np.copyto.)"""
        return VideoFrame(planes, self.format, dict(self.props))


class VideoNode:
    """Lazily evaluated clip: `get_frame(n)` calls the frame function (like a VapourSynth filter node)."""

    def __init__(self, num_frames: int, width: int, height: int, fmt: VideoFormat, frame_fn: Callable[[int], VideoFrame],
                 fps_num: int = 24, fps_den: int = 1):
        self.num_frames, self.width, self.height, self.format = num_frames, width, height, fmt
        self.fps_num, self.fps_den = fps_num, fps_den
        self._frame_fn = frame_fn
        self.std = _Std(self)

    def get_frame(self, n: int) -> VideoFrame:
        if not 0 <= n < self.num_frames:
            raise Error(f"frame {n} out of range [0, {self.num_frames})")
        return self._frame_fn(n)

    def frames(self):
        for n in range(self.num_frames):
            yield self.get_frame(n)

    def __len__(self):
        return self.num_frames


class _Std:
    def __init__(self, clip: Optional[VideoNode] = None):
        self._clip = clip

    def _c(self, clip):
        return clip if clip is not None else self._clip

    def ModifyFrame(self, clip=None, clips=None, selector=None):
        base = self._c(clip)
        srcs = ([clips] if isinstance(clips, VideoNode) else list(clips)) if clips is not None else [base]

        def fn(n):
            fs = [c.get_frame(n) for c in srcs]
            out = selector(n=n, f=fs[0] if len(fs) == 1 else fs)
            return out
        return VideoNode(base.num_frames, base.width, base.height, base.format, fn, base.fps_num, base.fps_den)

    def SetFrameProp(self, clip=None, prop=None, intval=None, floatval=None, data=None):
        base = self._c(clip)
        val = intval if intval is not None else (floatval if floatval is not None else data)

        def fn(n):
            f = base.get_frame(n).copy()
            f.props[prop] = val
            return f
        return VideoNode(base.num_frames, base.width, base.height, base.format, fn, base.fps_num, base.fps_den)

    def SetFrameProps(self, clip=None, **props):
        base = self._c(clip)

        def fn(n):
            f = base.get_frame(n).copy()
            f.props.update(props)
            return f
        return VideoNode(base.num_frames, base.width, base.height, base.format, fn, base.fps_num, base.fps_den)

    def CopyFrameProps(self, clip=None, prop_src=None, props=None):
        base = self._c(clip)

        def fn(n):
            f = base.get_frame(n).copy()
            src = prop_src.get_frame(n).props
            for k, v in src.items():
                if props is None or k in props:
                    f.props[k] = v
            return f
        return VideoNode(base.num_frames, base.width, base.height, base.format, fn, base.fps_num, base.fps_den)

    def BlankClip(self, clip=None, width=640, height=480, format=RGB24, length=240, color=None, fpsnum=24, fpsden=1):
        color = color or [0] * format.num_planes

        def fn(n):
            return VideoFrame([np.full((height, width), c, np.uint8) for c in color], format)
        return VideoNode(length, width, height, format, fn, fpsnum, fpsden)


class _Core:
    def __init__(self):
        self.std = _Std()
        self.messages: List[tuple] = []
        self.core_version = types.SimpleNamespace(release_major=70)

    def log_message(self, level, text):
        self.messages.append((int(level), str(text)))


core = _Core()


def array_clip(frames: np.ndarray, props: Optional[List[dict]] = None, fps_num=24, fps_den=1) -> VideoNode:
    """Source clip over a uint8 array [n, 3, H, W] (planar RGB24)."""
    n, planes, h, w = frames.shape
    fmt = RGB24 if planes == 3 else GRAY8

    def fn(i):
        return VideoFrame([frames[i, p] for p in range(planes)], fmt, props[i] if props else {"_DurationNum": fps_den, "_DurationDen": fps_num},
                          readonly=False)
    return VideoNode(n, w, h, fmt, fn, fps_num, fps_den)


def get_vs():
    """The real `vapoursynth` module if importable, else this shim."""
    try:
        import vapoursynth as vs  # type: ignore
        return vs
    except Exception:
        import sys
        return sys.modules[__name__]
