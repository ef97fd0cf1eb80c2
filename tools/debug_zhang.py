"""Stage-by-stage comparison of ZhangColorizer against the oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import zhang_oracle as z, pixel_oracle as px, synth_weights, metrics
from vsdeoldify_b200.zhang import ZhangColorizer
for name, S in (("siggraph17", 96), ("eccv16", 256)):
    B = 2
    sd = z.make_zhang_state_dict(name, 1234)
    frames = [np.stack([synth_weights.make_test_frame(300 + 7 * i + c, S, S).numpy() for c in range(3)], -1) for i in range(B)]
    col = ZhangColorizer(sd, name, B, S, torch.float16, device="cuda:0", keep_taps=True)
    rgb = torch.from_numpy(np.ascontiguousarray(np.stack([np.transpose(f, (2, 0, 1)) for f in frames]))).cuda()
    out = torch.empty_like(rgb)
    col.run(rgb, out, 0)
    torch.cuda.synchronize()
    f0 = frames[0]
    rs = px.pil_resize(f0, 256, 256, "bicubic")
    if col.resize:
        got_rs = np.transpose(col.rs[0].cpu().numpy(), (1, 2, 0))
        print(name, "resize mismatches", int((got_rs != rs).sum()))
    L_o = z.rgb2lab(f0)[..., 0]
    print(name, "L_orig max|d|", float(np.abs(col.L[0].cpu().numpy() - L_o.astype(np.float32)).max()))
    L_rs = torch.from_numpy(z.rgb2lab(rs)[..., 0]).float()[None, None]
    xn = ((L_rs - 50) / 100)[0, 0]
    print(name, "x max|d|", float((col.prog.x[0, :, :, 0].float().cpu() - xn).abs().max()), "other ch", float(col.prog.x[0, :, :, 1:].abs().max()))
    taps = {}
    with torch.no_grad():
        ab = (z.eccv16_forward if name == "eccv16" else z.siggraph17_forward)(sd, L_rs, taps=taps)
    for k, t in col.prog.taps.items():
        if k in taps and taps[k].dim() == 4:
            got = t[0:1].float().cpu()[..., :taps[k].shape[1]].permute(0, 3, 1, 2)
            w = taps[k]
            if got.shape != w.shape:
                print(name, k, "shape", got.shape, w.shape); continue
            print(f"{name} tap {k}: rms rel {float((got-w).pow(2).mean().sqrt()/w.pow(2).mean().sqrt()):.5f} max {float((got-w).abs().max()):.4f} ref max {float(w.abs().max()):.3f}")
    gab = col.prog.ab[0].cpu().permute(2, 0, 1)
    print(name, "ab max|d|", float((gab - ab[0]).abs().max()), "mean|d|", float((gab - ab[0]).abs().mean()), "ab std", float(ab.std()))
    want = z.colorize_frame(sd, name, f0)
    got = np.transpose(out[0].cpu().numpy(), (1, 2, 0))
    print(name, metrics.frame_parity(got, want))
    # post only: feed the oracle ab into the post kernel
    from vsdeoldify_b200 import _lib
    col.prog.ab[0].copy_(ab[0].permute(1, 2, 0).cuda())
    _lib.check(col.lib.havc_zhang_post(col.prog.ab.data_ptr(), 256, 256, col.L.data_ptr(), out.data_ptr(), B, S, S, 0))
    torch.cuda.synchronize()
    got = np.transpose(out[0].cpu().numpy(), (1, 2, 0))
    print(name, "post-only (oracle ab):", metrics.frame_parity(got, want))
