#!/usr/bin/env python
"""The CNN's conv2 / conv3 (2048 images) on the TMA im2col kernel, a few launches each: the target of `ncu -k regex:agx_conv2d_tma`."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from airgym_b200 import _capi  # noqa: E402
from airgym_b200.lib.network import tc_encoders as T  # noqa: E402

torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for Cin, Cout, H, W in ((16, 32, 106, 60), (32, 64, 53, 30)):
    conv = nn.Conv2d(Cin, Cout, 3, stride=2, padding=1).cuda()
    x = torch.randn(n, H, W, Cin, device="cuda")
    Lw = T._conv_weight(conv, True)
    for _ in range(reps):
        T.conv2d_nhwc(x, Lw, _capi.ACT_RELU)
torch.cuda.synchronize()
print("ok")
