"""TEST INFRASTRUCTURE (parity oracle) — CPU restatement of the whole per-frame DeOldify path.

  havc_colorizer_frame: HAVC_colorizer(method=0) on one RGB24 frame (vsdeoldify/__init__.py:2290-2523):
      Spline64 squeeze -> ModelImageRender/ColorizerFilter.filter (deoldify/filters.py:81-110) ->
      _clip_chroma_resize (vsdeoldify/__init__.py:3545-3554).
  colorizer_filter: MasterFilter([ColorizerFilter]).filter on a PIL-sized image (cfg1 path, Pillow BILINEAR).
"""
from __future__ import annotations

import numpy as np
import torch

from . import pixel_oracle as px
from . import unet_oracle


def model_process_square(sd, rgb_sq: np.ndarray) -> np.ndarray:
    """BaseFilter._model_process (filters.py:48-68) on an already square S x S uint8 RGB image -> uint8 RGB."""
    L = px.pil_luma(rgb_sq)
    x = torch.from_numpy(px.normalize_gray(L))[None]
    y = unet_oracle.unet_forward(sd, x)[0].numpy()
    return px.denorm_quantize(y)


def net_output_square(sd, rgb_sq: np.ndarray) -> np.ndarray:
    L = px.pil_luma(rgb_sq)
    x = torch.from_numpy(px.normalize_gray(L))[None]
    return unet_oracle.unet_forward(sd, x)[0].numpy()


def havc_colorizer_frame(sd, frame: np.ndarray, render_factor: int = 24, kernel: str = "spline64",
                         return_stages: bool = False, skip: bool = False, sd_other=None, video_weight: float = 0.5,
                         zhang=None, method: int = 0, merge_weight: float = 0.4, hue_adjust: str = "none", cmc_p=None,
                         lmm_p=None, alm_p=None, crt_p=None, invert: bool = False, ddtweak=None, frame_size=None, memo=None,
                         sat=(1.0, 1.0), hue=(0.0, 0.0)):
    """frame: uint8 [H,W,3].  Returns uint8 [H,W,3] (and the intermediate stages on request).
    skip: the scene-change gate returned the squeezed frame unchanged (vsslib/vsmodels.py:221-224).
    sd_other: 'stable'/'artistic' generator blended with the video one at S x S (visualize.py:118-137).
    zhang = (name, state_dict): second colour model (vs_sc_ddcolor models 2/3, vsmodels.py:339-344) with its hue
    adjustment (vsmodels.py:361-362), merged by vs_sc_combine_models(method, merge_weight, ...) (mcomb.py:125-192);
    method 1 = second model only.
    memo: optional dict that keeps the expensive stages of THIS frame (squeeze, generator results, second model) so that
    several merge methods can be checked against one set of network evaluations."""
    from . import filters_oracle as fo
    from . import zhang_oracle
    from . import zimg_oracle as zo
    H, W = frame.shape[:2]
    # frame_size (vsdeoldify/__init__.py:2502) = min(max(ddcolor_rf, deoldify_rf)*16, W); the DeOldify filter itself renders at
    # render_factor*16 and stretches with Pillow BILINEAR when the two differ (deoldify/filters.py:37-41,70-73,82-84)
    S = min(render_factor * 16, W) if frame_size is None else frame_size
    N = render_factor * 16

    memo = memo if memo is not None else {}

    def cached(key, fn):
        if key not in memo:
            memo[key] = fn()
        return memo[key]

    def deoldify(sd_):
        if N == S:
            return cached(("deoldify", id(sd_)), lambda: px.chroma_post_process(model_process_square(sd_, small), small))
        return cached(("deoldify", id(sd_)), lambda: colorizer_filter(sd_, small, render_factor))

    small = cached("small", lambda: px.resize_plane_u8(frame, S, S, kernel))   # clip.resize.Spline64(S, S)
    if skip:        # the selectors returned the frame unchanged; the clip-level vs_tweak of vs_sc_combine_models still runs on it
        k = 1 if method == 1 else 0
        colored = zo.vs_tweak(small, hue=hue[k], sat=sat[k])
    else:
        colored = None
        if method != 1:
            colored = deoldify(sd)                                    # filter.filter(): render + _post_process at S x S
            if sd_other is not None:
                colored = px.pil_blend(deoldify(sd_other), colored, video_weight)
        if zhang is not None and method != 0:
            src_b = small
            if ddtweak is not None:            # vs_sc_tweak(bright, cont) + sc_constrained_tweak (vsmodels.py:326-332)
                if ddtweak.get("bright", 0) != 0 or ddtweak.get("cont", 1) != 1:
                    src_b = fo.image_tweak(src_b, cont=ddtweak["cont"], bright=ddtweak["bright"])
                src_b = fo.luma_adjusted_levels(src_b, ddtweak["luma_min"], ddtweak["gamma"], ddtweak["gamma_luma_min"],
                                                ddtweak["gamma_alpha"], ddtweak["gamma_min"])
            clipb = cached(("zhang", zhang[0], ddtweak is not None), lambda: zhang_oracle.colorize_frame(zhang[1], zhang[0], src_b))
            clipb = fo.adjust_hue_range(clipb, hue_adjust)
            if ddtweak is not None:            # vs_recover_clip_luma(clip, clipb_rgb) (vsmodels.py:367-368)
                clipb = px.chroma_post_process(clipb, small)
            if method == 1:
                colored = zo.vs_tweak(clipb, hue=hue[1], sat=sat[1])           # mcomb.py:166-169
            else:
                a, b = (clipb, colored) if invert else (colored, clipb)
                a, b = zo.vs_tweak(a, hue=hue[0], sat=sat[0]), zo.vs_tweak(b, hue=hue[1], sat=sat[1])   # mcomb.py:154-169
                kw = {k: v for k, v in dict(cmc_p=cmc_p, lmm_p=lmm_p, alm_p=alm_p, crt_p=crt_p).items() if v is not None}
                colored = fo.combine_models(a, b, method, merge_weight, **kw)
        if method == 0 or zhang is None:
            colored = zo.vs_tweak(colored, hue=hue[0], sat=sat[0])             # mcomb.py:161-164 (clipb is None)
    up = px.resize_plane_u8(colored, W, H, kernel)                    # clip_lowres.resize.Spline64(W, H)
    out = px.chroma_post_process(up, frame)                           # vs_recover_clip_luma
    if return_stages:
        return out, dict(small=small, colored=colored, up=up)
    return out


def havc_stabilizer_frame(frame: np.ndarray, dark=False, dark_p=(0.2, 0.8), smooth=False, smooth_p=(0.3, 0.7, 0.9, 0.0, "none"),
                          colormap_adjust: str = "none", render_factor: int = 24, kernel: str = "spline64") -> np.ndarray:
    """HAVC_stabilizer with stab=False on one frame (vsdeoldify/__init__.py:2792-2871): Spline64 squeeze to
    min(render_factor*16, W) squared, vs_dark_tweak / vs_chroma_bright_tweak / vs_colormap, _clip_chroma_resize."""
    from . import filters_oracle as fo
    H, W = frame.shape[:2]
    S = min(render_factor * 16, W)
    small = px.resize_plane_u8(frame, S, S, kernel)
    colored = fo.stabilizer_stages(small, dark, dark_p, smooth, smooth_p, colormap_adjust)
    up = px.resize_plane_u8(colored, W, H, kernel)
    return px.chroma_post_process(up, frame)


def havc_stabilizer_clip(frames: np.ndarray, only, dark=False, dark_p=(0.2, 0.8), smooth=False, smooth_p=(0.3, 0.7, 0.9, 0.0, "none"),
                         colormap_adjust: str = "none", stab_p=(5, 'A', 1, 15, 0.2, 0.8), render_factor: int = 24,
                         kernel: str = "spline64", props=None) -> dict:
    """HAVC_stabilizer with stab=True up to (not including) vs_reduce_flicker, the external ReduceFlicker plugin
    (vsdeoldify/__init__.py:2792-2871): squeeze, per-frame stages, vs_chroma_stabilizer_ex on the squeezed clip (scope row N3,
    oracle/temporal_oracle.py), _clip_chroma_resize.  frames uint8 [T,H,W,3]; returns {n: uint8 [H,W,3]} for n in `only`."""
    from . import filters_oracle as fo, temporal_oracle as to
    T, H, W = frames.shape[:3]
    S = min(render_factor * 16, W)
    small = np.stack([fo.stabilizer_stages(px.resize_plane_u8(frames[n], S, S, kernel), dark, dark_p, smooth, smooth_p, colormap_adjust)
                      for n in range(T)])
    hue = stab_p[6] if len(stab_p) > 6 else "none"
    st = to.chroma_stabilizer_ex(small, nframes=stab_p[0], mode=stab_p[1], sat=stab_p[2], tht=stab_p[3], weight=stab_p[4],
                                 tht_scen=stab_p[5], hue_adjust=hue.lower(), props=props, only=list(only))
    return {n: px.chroma_post_process(px.resize_plane_u8(st[n], W, H, kernel), frames[n]) for n in only}


def min_hw_size(width: int, height: int, min_size=(512, 480)):
    """vsslib/vsresize.py:30-99: the size resize_min_HW resizes to, or None."""
    if height < width:
        if height <= min_size[1]:
            return None
        w = round(width * min_size[1] / height)
        return (w - 1 if w % 2 else w), min_size[1]
    if width <= min_size[0]:
        return None
    h = round(height * min_size[0] / width)
    return min_size[0], (h + 1 if h % 2 else h)


def resize_min_hw(frame: np.ndarray) -> np.ndarray:
    """resize_min_HW (vsslib/vsresize.py:30-50) on a uint8 [H,W,3] frame: zimg Spline36 (restated, unpinned)."""
    size = min_hw_size(frame.shape[1], frame.shape[0])
    return frame if size is None else px.resize_plane_u8(frame, size[0], size[1], "spline36")


def resize_to_chroma(high: np.ndarray, low: np.ndarray) -> np.ndarray:
    """resize_to_chroma (vsslib/vsresize.py:101-127): Spline36 of `low` to the size of `high`, both to YUV420P8 (BT.709 full range,
    no dither), luma of `high` + chroma of `low`, RGB24 with error diffusion (zimg restated, unpinned)."""
    from . import zimg_oracle as zo
    H, W = high.shape[:2]
    if low.shape[:2] != (H, W):
        low = px.resize_plane_u8(low, W, H, "spline36")
    y = zo.rgb24_to_yuv420p8(high, "709", False, False)[0]
    _, u, v = zo.rgb24_to_yuv420p8(low, "709", False, False)
    return zo.yuv420p8_to_rgb24(y, u, v, True, "709", False)


def colorizer_filter(sd, img: np.ndarray, render_factor: int) -> np.ndarray:
    """MasterFilter([ColorizerFilter]).filter(img, img, rf) (deoldify/filters.py:81-124) on a uint8 [H,W,3] image:
    Pillow-BILINEAR squeeze to S x S, network, Pillow-BILINEAR back, luma transplant.  This is the path
    ModelImageRender.get_transformed_image takes when it is handed a non-square image (BASELINE cfg1)."""
    H, W = img.shape[:2]
    S = render_factor * 16
    sq = px.pil_resize(img, S, S, "bilinear")                          # _scale_to_square
    model_img = model_process_square(sd, sq)                           # _transform + _model_process
    raw = px.pil_resize(model_img, W, H, "bilinear")                   # _unsquare
    return px.chroma_post_process(raw, img)                            # _post_process


def model_image_render(sd_video, sd_other, img: np.ndarray, render_factor: int, video_weight: float = 0.5) -> np.ndarray:
    """ModelImageRender.get_transformed_image (deoldify/visualize.py:118-137): the video net always runs;
    'stable'/'artistic' blend a second net: Image.blend(img_other, img_video, video_weight)."""
    v = colorizer_filter(sd_video, img, render_factor)
    if sd_other is None:
        return v
    o = colorizer_filter(sd_other, img, render_factor)
    return px.pil_blend(o, v, video_weight)


def havc_colorizer_yuv_frame(sd, planes, render_factor: int = 24, matrix: str = "709", out_limited: bool = True, **kw):
    """HAVC_colorizer on one 8-bit YUV 4:2:0 frame ([Y, U, V] planes) or GRAY frame ([Y]): convert_format_RGB24 (havc_utils.py:57-164:
    limited-range input, the clip's matrix, error-diffusion dither for YUV; plain range expansion for GRAY), the RGB24 path above,
    restore_format (:167-237: back to YUV420P8 with the clip's matrix / range - BT.709 for a GRAY source - and error-diffusion
    dither).  Returns [Y, U, V].  The zimg steps are the restatement of oracle/zimg_oracle.py (parity unpinned)."""
    from . import zimg_oracle as zo
    if len(planes) == 1:
        rgb = zo.gray8_to_rgb24(planes[0], limited=True)
        out_matrix = "709"
    else:
        rgb = zo.yuv420p8_to_rgb24(planes[0], planes[1], planes[2], dither=True, matrix=matrix, limited=True)
        out_matrix = matrix
    res = havc_colorizer_frame(sd, rgb, render_factor, **kw)
    return list(zo.rgb24_to_yuv420p8(res, matrix=out_matrix, limited=out_limited, dither=True))
