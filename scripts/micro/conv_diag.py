#!/usr/bin/env python
"""Operand-mapping probe of the TMA im2col convolution kernel (csrc/agx_conv_tma.cu): every activation carries its own index as its
value and the weights are one-hot rows, so y[pixel, o] names the input element the kernel actually multiplied for GEMM column
k0 + o.  Prints, per geometry, the number of wrong (pixel, k) pairs and a few decoded examples (expected vs fetched source index),
then times the CNN's conv2 / conv3 on both kernels.  python scripts/micro/conv_diag.py"""
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from airgym_b200 import _capi  # noqa: E402
from airgym_b200.lib.network import tc_encoders as T  # noqa: E402


def decode(idx, H, W, C):
    idx = int(idx)
    c = idx % C; idx //= C
    w = idx % W; idx //= W
    h = idx % H; idx //= H
    return (idx, h, w, c)


def probe(Cin, Cout, k, s, p, H, W, N):
    lib = _capi.load()
    kh, kw = (k, k) if isinstance(k, int) else k
    sy, sx = (s, s) if isinstance(s, int) else s
    py, px = (p, p) if isinstance(p, int) else p
    x = (torch.arange(N * H * W * Cin, device="cuda", dtype=torch.float32) + 1).reshape(N, H, W, Cin)  # 0 = padding
    K = kh * kw * Cin
    # expected im2col matrix [M, K] with k = (ky*kw + kx)*Cin + c
    cols = F.unfold(x.permute(0, 3, 1, 2).double(), (kh, kw), padding=(py, px), stride=(sy, sx))  # [N, Cin*kh*kw, L], (c, ky, kx) order
    L = cols.shape[2]
    cols = cols.reshape(N, Cin, kh * kw, L).permute(0, 3, 2, 1).reshape(N * L, K)
    bad_total, examples = 0, []
    for k0 in range(0, K, Cout):
        w = torch.zeros(Cout, K, device="cuda")
        for o in range(min(Cout, K - k0)):
            w[o, k0 + o] = 1.0
        hi, lo = T.split_tf32(w)
        Lw = {"hi": hi, "lo": lo, "bias": None, "Cin": Cin, "Cout": Cout, "k": (kh, kw), "s": (sy, sx), "p": (py, px)}
        y = T.conv2d_nhwc(x, Lw, _capi.ACT_NONE)
        torch.cuda.synchronize()
        got = y.reshape(-1, Cout).double().round()
        want = cols[:, k0:k0 + Cout]
        if want.shape[1] < Cout:
            want = torch.cat([want, torch.zeros(want.shape[0], Cout - want.shape[1], device="cuda", dtype=torch.float64)], 1)
        bad = (got != want).nonzero()
        bad_total += bad.shape[0]
        for b in bad[:3]:
            m, o = int(b[0]), int(b[1])
            examples.append({"pixel": m, "k": k0 + o, "tap": (k0 + o) // Cin, "c": (k0 + o) % Cin,
                             "want": decode(want[m, o] - 1, H, W, Cin) if want[m, o] > 0 else "pad",
                             "got": decode(got[m, o] - 1, H, W, Cin) if 0 < got[m, o] <= N * H * W * Cin else float(got[m, o])})
    return {"geom": [Cin, Cout, k, s, p, H, W, N], "M": N * L, "K": K, "wrong_pairs": bad_total, "examples": examples[:12]}


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


if __name__ == "__main__":
    lib = _capi.load()
    out = {"probes": []}
    for impl in (1,):
        _capi.check(lib.agx_set_option(b"conv_impl", impl), "conv_impl")
        for g in [(16, 32, 3, 2, 1, 10, 12, 3), (32, 64, 3, 2, 1, 9, 10, 5), (64, 64, 5, 2, 2, 15, 26, 2), (64, 128, 5, 4, (2, 1), 15, 26, 3),
                  (128, 128, 3, 1, 1, 4, 7, 9)]:
            try:
                out["probes"].append(probe(*g))
            except Exception as e:  # keep going: the timing below is still wanted
                out["probes"].append({"geom": list(g), "error": repr(e)})
    # timing, CNN conv2 / conv3 at 2048 images
    torch.manual_seed(0)
    res = {}
    for name, (Cin, Cout, H, W) in {"conv2": (16, 32, 106, 60), "conv3": (32, 64, 53, 30)}.items():
        conv = nn.Conv2d(Cin, Cout, 3, stride=2, padding=1).cuda()
        x = torch.randn(2048, H, W, Cin, device="cuda")
        ref = None
        for precise in (True, False):
            Lw = T._conv_weight(conv, precise)
            for impl in (0, 1):
                _capi.check(lib.agx_set_option(b"conv_impl", impl), "conv_impl")
                y = T.conv2d_nhwc(x, Lw, _capi.ACT_RELU)
                torch.cuda.synchronize()
                if ref is None:
                    ref = y
                res[f"{name}_{'3xtf32' if precise else 'tf32'}_impl{impl}"] = {
                    "ms_per_2048": timed(lambda: T.conv2d_nhwc(x, Lw, _capi.ACT_RELU)), "max_abs_diff_vs_first": float((y - ref).abs().max())}
    _capi.check(lib.agx_set_option(b"conv_impl", 1), "conv_impl")
    out["timing"] = res
    print(json.dumps(out, indent=1, default=str))
