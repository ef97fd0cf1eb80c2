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
