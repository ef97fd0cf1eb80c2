"""CPU tests: the oracle's integer pixel formulas against the libraries the reference calls (OpenCV, Pillow)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
Image = pytest.importorskip("PIL.Image")


def _rand(shape, seed):
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def test_cv_yuv_roundtrip_formulas_exact():
    from oracle import pixel_oracle as px
    a = _rand((211, 307, 3), 0)
    assert np.array_equal(px.cv_rgb2yuv(a), cv2.cvtColor(a, cv2.COLOR_RGB2YUV))
    assert np.array_equal(px.cv_yuv2rgb(a), cv2.cvtColor(a, cv2.COLOR_YUV2RGB))
    # saturated corners
    edge = np.array([[[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0]]], np.uint8)
    assert np.array_equal(px.cv_rgb2yuv(edge), cv2.cvtColor(edge, cv2.COLOR_RGB2YUV))
    assert np.array_equal(px.cv_yuv2rgb(edge), cv2.cvtColor(edge, cv2.COLOR_YUV2RGB))


def test_pil_luma_exact():
    from oracle import pixel_oracle as px
    a = _rand((97, 131, 3), 1)
    ref = np.asarray(Image.fromarray(a).convert("LA").convert("RGB"))
    assert np.array_equal(px.pil_luma(a), ref[..., 0]) and np.array_equal(ref[..., 0], ref[..., 2])


@pytest.mark.parametrize("shape,out", [((767, 1090), (384, 384)), ((384, 384), (767, 1090)), ((33, 47), (90, 20)),
                                       ((64, 64), (64, 64))])
@pytest.mark.parametrize("filt", ["bilinear", "bicubic"])
def test_pil_resize_exact(shape, out, filt):
    from oracle import pixel_oracle as px
    a = _rand(shape + (3,), 2)
    pf = {"bilinear": Image.BILINEAR, "bicubic": Image.BICUBIC}[filt]
    ref = np.asarray(Image.fromarray(a).resize((out[1], out[0]), resample=pf))
    assert np.array_equal(px.pil_resize(a, out[1], out[0], filt), ref)


@pytest.mark.parametrize("alpha", [0.0, 0.15, 0.3, 0.5, 0.6, 1.0, 1.4, -0.2])
def test_pil_blend_exact(alpha):
    from oracle import pixel_oracle as px
    a, b = _rand((40, 50, 3), 3), _rand((40, 50, 3), 4)
    ref = np.asarray(Image.blend(Image.fromarray(a), Image.fromarray(b), alpha))
    assert np.array_equal(px.pil_blend(a, b, alpha), ref)


def test_spline_resize_properties():
    """zimg is absent (parity unpinned): check the properties any correct Spline64 resampler has."""
    from oracle import pixel_oracle as px
    m = px.resize_matrix(1920, 384)
    assert np.allclose(m.sum(1), 1.0)
    flat = np.full((108, 192), 77, np.uint8)
    assert np.array_equal(px.resize_plane_u8(flat, 64, 64), np.full((64, 64), 77, np.uint8))
    same = _rand((48, 48), 5)
    assert np.array_equal(px.resize_plane_u8(same, 48, 48), same)            # identity at equal size
    ramp = np.tile(np.linspace(0, 255, 96).astype(np.uint8), (8, 1))
    up = px.resize_plane_u8(ramp, 384, 8).astype(int)
    assert (np.diff(up[0, 8:-8]) >= -1).all()                                # monotone ramp stays monotone (+-1 rounding)


def test_ciede2000_reference_pairs():
    from oracle import metrics
    # Sharma, Wu, Dalal (2005) test data
    assert abs(metrics.ciede2000(np.array([50, 2.6772, -79.7751]), np.array([50, 0, -82.7485])) - 2.0425) < 1e-4
    assert abs(metrics.ciede2000(np.array([50, 2.5, 0]), np.array([73, 25, -18])) - 27.1492) < 1e-4
    assert abs(metrics.ciede2000(np.array([2.0776, 0.0795, -1.1350]), np.array([0.9033, -0.0636, -0.5514])) - 0.9082) < 1e-4
