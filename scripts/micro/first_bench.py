#!/usr/bin/env python
"""First-layer kernel alone at 2048 images (ncu target / timing): python scripts/micro/first_bench.py [norm 0|1] [conv_first mode]"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from airgym_b200 import _capi  # noqa: E402
from airgym_b200.lib.network import tc_encoders as T  # noqa: E402

norm = len(sys.argv) > 1 and sys.argv[1] == "1"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
_capi.load().agx_set_option(b"conv_first", mode)
torch.manual_seed(0)
conv = nn.Conv2d(1, 16, 5, stride=2, padding=2).cuda()
img = torch.rand(2048, 212, 120, device="cuda") * 10
mean, rstd = torch.rand(212 * 120, device="cuda") * 5, torch.rand(212 * 120, device="cuda") + 0.2
sc, sh = torch.rand(16, device="cuda") + 0.5, torch.randn(16, device="cuda")
for _ in range(3):
    T.conv2d_first(img, conv, _capi.ACT_RELU, mean if norm else None, rstd if norm else None, sc, sh)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    T.conv2d_first(img, conv, _capi.ACT_RELU, mean if norm else None, rstd if norm else None, sc, sh)
b.record()
torch.cuda.synchronize()
print("ms per 2048:", a.elapsed_time(b) / 10)
