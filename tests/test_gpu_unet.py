"""GPU parity of the full DeOldify generators and of the per-frame HAVC_colorizer(method=0) pipeline against
the CPU fp32 oracle (oracle/unet_oracle.py + oracle/pipeline_oracle.py) on seeded synthetic weights.

Gates (BASELINE.json north_star, tests/parity_gate.py): mean CIEDE2000 <= 0.5 per frame - asserted unmodified for every
generator; max 8-bit channel error <= 2 - asserted unmodified in the xfail-marked *_strict_max_gate tests (the reference
violates it against itself).
"""
import numpy as np
import pytest
import torch

from parity_gate import XFAIL_REASON, assert_mean_gate, assert_outlier_guard, strict_max_gate

pytestmark = pytest.mark.gpu

_CACHE = {}


def _sd(arch, seed=1234):
    from oracle import synth_weights
    key = (arch, seed)
    if key not in _CACHE:
        _CACHE[key] = synth_weights.make_unet_state_dict(arch, seed)
    return _CACHE[key]


def _frames(n, h, w, seed=3):
    from oracle import synth_weights
    out = np.zeros((n, 3, h, w), np.uint8)
    for i in range(n):
        g = synth_weights.make_test_frame(seed + i, h, w).numpy()
        out[i] = g[None]
    return out


def _nchw(t, c):
    t = t.float() if hasattr(t, "hi") else t          # split-precision tensors (ops.Pair): hi + lo
    return t[..., :c].permute(0, 3, 1, 2).float().cpu()


@pytest.mark.parametrize("arch,dtype", [("wide", torch.float16), ("deep", torch.float16), ("wide", torch.bfloat16)])
def test_unet_layers_vs_oracle(arch, dtype):
    from oracle import pixel_oracle as px, unet_oracle
    from vsdeoldify_b200.engine import DeoldifyEngine
    sd = _sd(arch)
    S, B = 64, 2
    eng = DeoldifyEngine(sd, S, S, render_factor=S // 16, batch=B, dtype=dtype, use_graph=False, keep_taps=True,
                         debug_net_out=True)
    frames = _frames(B, S, S)
    eng.colorize_batch(frames)
    # oracle on the same (identity-resized) input
    x = torch.stack([torch.from_numpy(px.normalize_gray(px.pil_luma(np.transpose(f, (1, 2, 0))))) for f in frames])
    taps = {}
    y_ref = unet_oracle.unet_forward(sd, x, taps=taps)
    report, worst = [], 0.0
    for name, ref in taps.items():
        if name not in eng.prog.taps:
            continue
        t = eng.prog.taps[name]
        got = _nchw(t, ref.shape[1])
        if name == "logits":   # the fused head stores fp32 logits without the 1x1 conv's bias
            got = got + sd["layers.11.0.bias"].view(1, 3, 1, 1)
        rel = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
        rms = float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt().clamp_min(1e-6))
        report.append(f"{name}: max-rel {rel:.2e} rms-rel {rms:.2e}")
        worst = max(worst, rms)
    y = eng.net_out.cpu()
    err = (y - y_ref).abs()
    report.append(f"net_out: max abs {err.max():.3e} rms {float((err ** 2).mean().sqrt()):.3e}")
    print("\n".join(report))
    # 16-bit activations: every layer injects ~2^-11 (fp16) / 2^-8 (bf16) relative rounding noise which the
    # network itself amplifies ~10x on the way to the logits (measured on the CPU with emulated rounding)
    tol = 1.5e-2 if dtype == torch.float16 else 1.2e-1
    assert worst < tol, "\n".join(report)
    # the artistic generator has two 3x3 convs per U-Net block and is ~1.5x noisier than the wide one
    out_tol = (1e-2 if arch == "wide" else 1.6e-2) if dtype == torch.float16 else 8e-2
    assert float((err ** 2).mean().sqrt()) < out_tol, "\n".join(report)


def _colorizer_frames_vs_oracle(arch):
    from oracle import metrics, pipeline_oracle
    from vsdeoldify_b200.engine import DeoldifyEngine
    sd = _sd(arch)
    H, W, rf, B = 180, 320, 8, 2
    key = ("frames", arch)
    if key not in _CACHE:
        eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=B, dtype=torch.float16)
        frames = _frames(B, H, W, seed=11)
        out = eng.colorize_batch(frames)
        res = []
        for i in range(B):
            ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(frames[i], (1, 2, 0)), rf)
            got = np.transpose(out[i], (1, 2, 0))
            res.append((metrics.frame_parity(got, ref), got, np.transpose(frames[i], (1, 2, 0))))
        _CACHE[key] = res
    return _CACHE[key]


@pytest.mark.parametrize("arch", ["wide", "deep"])
def test_colorizer_frame_vs_oracle(arch):
    for i, (m, got, src) in enumerate(_colorizer_frames_vs_oracle(arch)):
        print(arch, i, m)
        assert_mean_gate(m, (arch, i))                          # north-star gate, both generators
        assert_outlier_guard(m, 1.0e-2, (arch, i))              # regression guard (measured: wide 3.5e-3 / deep 4.3e-3)
        # the colourised frame keeps the source luma: the path is not a pass-through
        assert np.abs(got.astype(int) - src.astype(int)).max() > 8


@pytest.mark.xfail(reason=XFAIL_REASON, strict=False)
@pytest.mark.parametrize("arch", ["wide", "deep"])
def test_colorizer_frame_strict_max_gate(arch):
    for i, (m, got, src) in enumerate(_colorizer_frames_vs_oracle(arch)):
        strict_max_gate(m, (arch, i))


def test_graph_replay_is_deterministic_and_ordered():
    from vsdeoldify_b200.engine import DeoldifyEngine
    sd = _sd("wide")
    H, W, rf, B = 96, 128, 4, 2
    eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=B, dtype=torch.float16)
    batches = [_frames(B, H, W, seed=20 + 2 * k) for k in range(5)]
    ref = [eng.colorize_batch(b) for b in batches]
    got = {}
    n = eng.colorize_stream(iter(batches), lambda i, o: got.__setitem__(i, o.copy()))
    assert n == 5 and sorted(got) == list(range(5))
    for i in range(5):
        assert np.array_equal(got[i], ref[i]), f"batch {i} differs between sync and pipelined paths"


@pytest.mark.parametrize("H,W,rf", [(83, 150, 8), (97, 131, 6), (64, 100, 8)])
def test_colorizer_frame_odd_sizes_scalar_pixel_kernels(H, W, rf):
    """Frame widths that are not multiples of 4 (and a frame narrower than render_factor*16: S = W) take the scalar
    fall-backs of the squeeze / un-squeeze kernels instead of the 4-pixel / multi-row variants: same parity bar."""
    from oracle import metrics, pipeline_oracle
    from vsdeoldify_b200.engine import DeoldifyEngine
    sd = _sd("wide")
    eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=2, dtype=torch.float16)
    frames = _frames(2, H, W, seed=31)
    out = eng.colorize_batch(frames)
    for i in range(2):
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(frames[i], (1, 2, 0)), rf)
        m = metrics.frame_parity(np.transpose(out[i], (1, 2, 0)), ref)
        assert_mean_gate(m, (H, W, rf, i))


@pytest.mark.parametrize("H,W,rf,frame_size", [(90, 80, 6, None), (96, 160, 4, 96), (120, 200, 6, 160)])
def test_colorizer_frame_size_differs_from_render_size(H, W, rf, frame_size):
    """frame_size != render_factor*16: a clip narrower than rf*16 (W = 80 < 96) and a second model with a bigger render factor
    (frame_size = ddcolor_rf*16 > deoldify_rf*16).  The filter's own Pillow BILINEAR stretch to rf*16 and back
    (deoldify/filters.py:37-41,70-73) runs around the generator; the oracle branch is pinned against the real
    ModelImageRender by tests/test_oracle_golden.py (model_image_render_narrow.npz)."""
    from oracle import metrics, pipeline_oracle
    from vsdeoldify_b200.engine import DeoldifyEngine
    sd = _sd("wide")
    eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=2, dtype=torch.float16, frame_size=frame_size)
    assert eng.N == rf * 16 and eng.S == (frame_size or min(rf * 16, W)) and eng.N != eng.S
    frames = _frames(2, H, W, seed=51)
    out = eng.colorize_batch(frames, skip=np.array([False, True]))
    ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(frames[0], (1, 2, 0)), rf, frame_size=frame_size)
    m = metrics.frame_parity(np.transpose(out[0], (1, 2, 0)), ref)
    assert_mean_gate(m)
    ref_skip = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(frames[1], (1, 2, 0)), rf, frame_size=frame_size, skip=True)
    assert np.array_equal(np.transpose(out[1], (1, 2, 0)), ref_skip), "a scene-change-skipped frame is the uncoloured squeeze path"


@pytest.mark.parametrize("arch,rf", [("wide", 5), ("deep", 5), ("wide", 7)])
def test_odd_render_factor(arch, rf):
    """Odd render factors (legal in the reference: 10..64, every preset value of HAVC_main is even but ddcolor_rf = 0 / custom
    values are not): S = 16 * odd, the encoder's last stage is ceil(odd / 2) wide, the first U-Net block's up path comes out one
    pixel larger than its skip and the reference resizes it with F.interpolate(mode='nearest') (unet.py:201-203) - a crop for a
    one-pixel mismatch.  The attention runs on (S / 8)^2 = 4 * odd^2 tokens (not a multiple of 8)."""
    from oracle import metrics, pipeline_oracle
    from vsdeoldify_b200.engine import DeoldifyEngine
    sd = _sd(arch)
    H, W, B = 130, 176, 2
    eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=B, dtype=torch.float16)
    assert eng.N == rf * 16 and eng.N % 32 == 16
    frames = _frames(B, H, W, seed=71)
    out = eng.colorize_batch(frames)
    for i in range(B):
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(frames[i], (1, 2, 0)), rf)
        m = metrics.frame_parity(np.transpose(out[i], (1, 2, 0)), ref)
        print(arch, rf, i, m)
        assert_mean_gate(m, (arch, rf, i))
