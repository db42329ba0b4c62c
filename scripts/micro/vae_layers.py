#!/usr/bin/env python
"""Per-layer timing of the VAE ImgEncoder's libagx layer sequence (2048 images): python scripts/micro/vae_layers.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from airgym_b200 import _capi  # noqa: E402
from airgym_b200.lib.network import tc_encoders as T  # noqa: E402
from airgym_b200.lib.network.vae_image_encoder import VAEImageEncoder  # noqa: E402


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


torch.manual_seed(0)
vae = VAEImageEncoder({"latent_dims": 64, "image_res": [120, 212], "interpolation_mode": "bilinear", "allow_random_init": True}).cuda()
enc = vae.encoder if hasattr(vae, "encoder") else vae.vae.encoder
W_ = T._prep_vae(enc, True)
lib = _capi.load()
n = 2048
x = torch.rand(n, 212, 120, device="cuda") * 9
r = torch.empty(n, 120, 212, device="cuda")
E, R = _capi.ACT_ELU, _capi.ACT_NONE
out = {}
out["resize"] = timed(lambda: lib.agx_resize_bilinear(x.data_ptr(), r.data_ptr(), n, 212, 120, 120, 212, None))
t0 = T.conv2d_first(r, enc.conv0, R); out["conv0 1->32 5x5s2"] = timed(lambda: T.conv2d_first(r, enc.conv0, R))
a = T.conv2d_nhwc(t0, W_["conv0_1"], E); out["conv0_1 32->32 3x3s2 %s" % (tuple(t0.shape[1:3]),)] = timed(lambda: T.conv2d_nhwc(t0, W_["conv0_1"], E))
t1 = T.conv2d_nhwc(a, W_["conv1_0"], R); out["conv1_0 32->32 5x5s2 %s" % (tuple(a.shape[1:3]),)] = timed(lambda: T.conv2d_nhwc(a, W_["conv1_0"], R))
j2 = T.conv2d_nhwc(a, W_["conv0_jump_2"], R); out["jump_2 32->64 4x4s2"] = timed(lambda: T.conv2d_nhwc(a, W_["conv0_jump_2"], R))
b = T.conv2d_nhwc(t1, W_["conv1_1"], E, res=j2); out["conv1_1 32->64 3x3s1 %s" % (tuple(t1.shape[1:3]),)] = timed(lambda: T.conv2d_nhwc(t1, W_["conv1_1"], E, res=j2))
t2 = T.conv2d_nhwc(b, W_["conv2_0"], R); out["conv2_0 64->64 5x5s2 %s" % (tuple(b.shape[1:3]),)] = timed(lambda: T.conv2d_nhwc(b, W_["conv2_0"], R))
j3 = T.conv2d_nhwc(b, W_["conv1_jump_3"], R); out["jump_3 64->128 5x5s4"] = timed(lambda: T.conv2d_nhwc(b, W_["conv1_jump_3"], R))
c = T.conv2d_nhwc(t2, W_["conv2_1"], E, res=j3); out["conv2_1 64->128 3x3s2 %s" % (tuple(t2.shape[1:3]),)] = timed(lambda: T.conv2d_nhwc(t2, W_["conv2_1"], E, res=j3))
t3 = T.conv2d_nhwc(c, W_["conv3_0"], R); out["conv3_0 128->128 3x3s1 %s" % (tuple(c.shape[1:3]),)] = timed(lambda: T.conv2d_nhwc(c, W_["conv3_0"], R))
f = t3.reshape(n, 1, 1, -1)
d0 = T.conv2d_nhwc(f, W_["dense0"], E); out["dense0 3584->512"] = timed(lambda: T.conv2d_nhwc(f, W_["dense0"], E))
out["dense1 512->128"] = timed(lambda: T.conv2d_nhwc(d0, W_["dense1"], R))
out["sum"] = round(sum(out.values()), 3)
out["vae_encode_2048"] = timed(lambda: T.vae_encode(enc, (120, 212), x.view(n, 1, 212, 120)), iters=5)
print(json.dumps(out, indent=1))
