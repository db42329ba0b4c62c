#!/usr/bin/env python
"""First-layer tensor-core kernel alone (debug target for compute-sanitizer): python scripts/micro/first_one.py [n] [norm]"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from airgym_b200 import _capi  # noqa: E402
from airgym_b200.lib.network import tc_encoders as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
norm = len(sys.argv) > 2 and sys.argv[2] == "1"
torch.manual_seed(0)
for Cout, H, W in ((16, 212, 120), (32, 120, 212)):
    conv = nn.Conv2d(1, Cout, 5, stride=2, padding=2).cuda()
    img = torch.rand(n, H, W, device="cuda") * 10
    mean, rstd = torch.rand(H * W, device="cuda") * 5, torch.rand(H * W, device="cuda") + 0.2
    y = T.conv2d_first(img, conv, _capi.ACT_RELU, mean if norm else None, rstd if norm else None)
    torch.cuda.synchronize()
    xn = torch.clamp((img - mean.view(H, W)) * rstd.view(H, W), -5, 5) if norm else img
    with torch.no_grad():
        ref = torch.relu(F.conv2d(xn.unsqueeze(1).double(), conv.weight.double(), conv.bias.double(), stride=2, padding=2))
    d = (y.permute(0, 3, 1, 2).double() - ref).abs()
    print(Cout, "max err", float(d.max()), "ref max", float(ref.abs().max()), "bad frac", float((d > 1e-4).double().mean()))
    if float(d.max()) > 1e-4:
        bad = (d > 1e-4).nonzero()
        print(" first bad (n,c,y,x):", bad[:6].tolist(), " bad per channel:", (d > 1e-4).sum((0, 2, 3)).tolist()[:16], " bad per x:", (d > 1e-4).sum((0, 1, 2)).tolist()[:12])
