#!/usr/bin/env python
"""Stage-depth / slabs-per-stage sweep of the TMA convolution kernel on the CNN's conv2 / conv3 (2048 images): is the pipeline bound by
the latency of a stage round trip (time ~ 1 / stages) or by a throughput?  python scripts/micro/conv_sweep.py"""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from airgym_b200 import _capi  # noqa: E402
from airgym_b200.lib.network import tc_encoders as T  # noqa: E402


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
    return round(a.elapsed_time(b) / iters, 4)


lib = _capi.load()
torch.manual_seed(0)
out = {}
for name, (Cin, Cout, H, W) in {"conv2": (16, 32, 106, 60), "conv3": (32, 64, 53, 30)}.items():
    conv = nn.Conv2d(Cin, Cout, 3, stride=2, padding=1).cuda()
    x = torch.randn(2048, H, W, Cin, device="cuda")
    for precise in (True, False):
        Lw = T._conv_weight(conv, precise)
        for spp in (1, 2, 3, 4):
            for stages in (2, 3, 4, 6, 8):
                lib.agx_set_option(b"conv_spp", spp)
                lib.agx_set_option(b"conv_stages", stages)
                try:
                    out[f"{name}_{'split' if precise else 'tf32'}_spp{spp}_st{stages}"] = timed(lambda: T.conv2d_nhwc(x, Lw, _capi.ACT_RELU))
                except Exception as e:
                    out[f"{name}_{'split' if precise else 'tf32'}_spp{spp}_st{stages}"] = repr(e)[:60]
lib.agx_set_option(b"conv_spp", 0)
lib.agx_set_option(b"conv_stages", 0)
print(json.dumps(out, indent=0))
