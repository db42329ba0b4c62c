#!/usr/bin/env python
"""Depth-encoder timing probe (row f3): CNNFeatureExtractor forward on [N,1,212,120] images, eval-mode BatchNorm.
python scripts/enc_bench.py [--n 8192]  → ms per encode for the libagx kernel (with and without the fused input normalisation) and for torch/cuDNN in fp32 and TF32."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from airgym_b200.lib.network.cnn import CNNFeatureExtractor  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--skip_cudnn", action="store_true")
    ap.add_argument("--libs", default="", help="comma-separated variant builds of libagx.so (scripts/build_variant.sh) to A/B")
    a = ap.parse_args()
    torch.manual_seed(0)
    net = CNNFeatureExtractor(30).cuda().eval()
    x = torch.rand(a.n, 1, 212, 120, device="cuda")
    out = {"n": a.n}
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        from airgym_b200.lib.network.cnn import native_encode
        ref = net.forward_torch(x[:256])
        got = native_encode(net, x[:256])
        out["native_max_abs_err_vs_cudnn_fp32"] = float((got - ref).abs().max())
        out["native_ms"] = timed(lambda: native_encode(net, x))
        mean, rstd = torch.rand(212 * 120, device="cuda"), torch.rand(212 * 120, device="cuda") + 0.5
        out["native_fused_norm_ms"] = timed(lambda: native_encode(net, x, mean, rstd))
        if a.libs:
            import ctypes as C
            from airgym_b200 import _capi
            torch.backends.cudnn.allow_tf32 = False
            ref = net.forward_torch(x[:296])
            for path in a.libs.split(","):
                lib = _capi.bind(C.CDLL(os.path.abspath(path)))
                got = native_encode(net, x[:296], lib=lib)
                torch.cuda.synchronize()
                out[os.path.basename(path)] = {"max_abs_err": float((got - ref).abs().max()),
                                               "ms": timed(lambda: native_encode(net, x, lib=lib))}
        if not a.skip_cudnn:
            out["cudnn_fp32_ms"] = timed(lambda: net.forward_torch(x))
            torch.backends.cudnn.allow_tf32 = True
            out["cudnn_tf32_ms"] = timed(lambda: net.forward_torch(x))
    print(json.dumps(out))
