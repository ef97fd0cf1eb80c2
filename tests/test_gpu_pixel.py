"""GPU tests of the phase-periodic horizontal resampling kernels (csrc/pixel.cu) against the table-driven kernels they replace
for integer ratios: every output bit equal (the periodic kernels run the same fmaf chain per output, with the weights in the
parameter bank and the inputs in registers)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tables(src, dst):
    from vsdeoldify_b200.engine import _Tables
    return _Tables(src, dst, "spline64", torch.device("cuda:0"))


@pytest.mark.parametrize("src,dst,rows", [(1920, 384, 2 * 3 * 13), (1920, 480, 3 * 9), (3840, 640, 3 * 7), (1920, 320, 8), (1920, 384, 8 * 148 * 2 + 3)])
def test_periodic_squeeze_equals_table_kernel(src, dst, rows):
    from vsdeoldify_b200 import _lib
    lib = _lib.lib()
    t = _tables(src, dst)
    assert t.plan is not None, "integer ratio: a periodic plan is expected"
    g = torch.Generator().manual_seed(src + dst + rows)
    x = torch.randint(0, 256, (rows, src), dtype=torch.uint8, generator=g).cuda()
    a = torch.full((rows, dst), -1.0, dtype=torch.float32, device="cuda")
    b = torch.full((rows, dst), -2.0, dtype=torch.float32, device="cuda")
    _lib.check(lib.havc_resample_h(x.data_ptr(), a.data_ptr(), rows, src, dst, t.start.data_ptr(), t.wt.data_ptr(), t.taps, 0))
    _lib.check(lib.havc_resample_h_periodic(x.data_ptr(), b.data_ptr(), rows, src, dst, t.start.data_ptr(), t.wt.data_ptr(), t.taps,
                                            C.byref(t.plan), 0))
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)), int((a.view(torch.int32) != b.view(torch.int32)).sum())


@pytest.mark.parametrize("S,W,B,H,transplant", [(384, 1920, 2, 13, 1), (384, 1920, 1, 5, 0), (480, 1920, 2, 7, 1), (640, 3840, 1, 6, 1),
                                                (320, 1920, 3, 4, 1), (384, 1920, 4, 301, 1)])
def test_periodic_way_back_equals_table_kernel(S, W, B, H, transplant):
    from vsdeoldify_b200 import _lib
    lib = _lib.lib()
    t = _tables(S, W)
    assert t.plan is not None
    g = torch.Generator().manual_seed(S + W + H)
    x = (torch.rand(B, 3, H, S, generator=g) * 270.0 - 8.0).cuda()           # the vertical pass leaves values slightly outside [0, 255]
    orig = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).cuda()
    a = torch.full((B, 3, H, W), 7, dtype=torch.uint8, device="cuda")
    b = torch.full((B, 3, H, W), 9, dtype=torch.uint8, device="cuda")
    op = orig.data_ptr() if transplant else None
    _lib.check(lib.havc_post_horizontal(x.data_ptr(), op, a.data_ptr(), B, S, H, W, t.start.data_ptr(), t.wt.data_ptr(), t.taps, transplant, 0))
    _lib.check(lib.havc_post_horizontal_periodic(x.data_ptr(), op, b.data_ptr(), B, S, H, W, t.start.data_ptr(), t.wt.data_ptr(), t.taps,
                                                 transplant, C.byref(t.plan), 0))
    torch.cuda.synchronize()
    assert torch.equal(a, b), (int((a != b).sum()), (a != b).nonzero()[:5].tolist())


def test_non_integer_ratio_has_no_plan_and_uses_the_table_kernels():
    assert _tables(1280, 384).plan is None and _tables(384, 1280).plan is None and _tables(1920, 256).plan is None


@pytest.mark.parametrize("B,S,c0", [(2, 384, 4), (1, 96, 4), (3, 64, 0), (1, 130, 4)])
def test_stem_im2col_rows_are_the_gathered_input_words(B, S, c0):
    """havc_im2col_small for the stem (7 x 7, stride 2, pad 3, two channels, Kp = 128 -> im2col_stem_kernel): pure data movement,
    every byte compared with a torch gather (zero padding outside the image, zero tail of every K row)."""
    from vsdeoldify_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(S + c0)
    x = torch.randn(B, S, S, 8, generator=g).half().cuda()
    OH = (S + 6 - 7) // 2 + 1
    out = torch.full((B, OH, OH, 128), 7.0, dtype=torch.float16, device="cuda")
    _lib.check(lib.havc_im2col_small(x.data_ptr(), out.data_ptr(), B, S, S, 8, c0, 2, 7, 2, 3, 128, 0, 0))
    torch.cuda.synchronize()
    xp = torch.nn.functional.pad(x[..., c0:c0 + 2].permute(0, 3, 1, 2), (3, 3, 3, 3))          # [B, 2, S + 6, S + 6]
    want = torch.zeros(B, OH, OH, 128, dtype=torch.float16, device="cuda")
    for kh in range(7):
        for kw in range(7):
            patch = xp[:, :, kh:kh + 2 * OH:2, kw:kw + 2 * OH:2]                                # [B, 2, OH, OH]
            want[..., kh * 16 + kw * 2:kh * 16 + kw * 2 + 2] = patch.permute(0, 2, 3, 1)
    assert torch.equal(out.view(torch.int16), want.view(torch.int16)), int((out.view(torch.int16) != want.view(torch.int16)).sum())
