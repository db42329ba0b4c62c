"""Encoder layers on the tensor cores (csrc/agx_conv.cu, lib/network/tc_encoders.py) against torch's float64 convolutions:
every layer geometry of the CNN and of the VAE ImgEncoder (strides 1/2/4, 3x3/4x4/5x5, asymmetric padding, cropped and broadcast
skip branches, the dense layers as 1x1 convolutions), the 3xTF32 split at fp32-level accuracy and the single-pass TF32 mode at
TF32-level accuracy; then the two encoders end to end against their torch mirrors and the committed golden latents."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from airgym_b200 import _capi
from airgym_b200.lib.network import tc_encoders as T

pytestmark = pytest.mark.gpu

GEOMS = [  # (Cin, Cout, k, s, p, H, W, act, with_res)
    (16, 32, 3, 2, 1, 106, 60, "relu", False),   # CNN conv2
    (32, 64, 3, 2, 1, 53, 30, "relu", False),    # CNN conv3
    (32, 32, 3, 2, 2, 60, 106, "elu", False),    # VAE conv0_1
    (32, 32, 5, 2, 1, 31, 54, "none", False),    # conv1_0
    (32, 64, 4, 2, 1, 31, 54, "none", False),    # conv0_jump_2
    (32, 64, 3, 1, 1, 15, 26, "elu", True),      # conv1_1 + skip
    (64, 64, 5, 2, 2, 15, 26, "none", False),    # conv2_0
    (64, 128, 5, 4, (2, 1), 15, 26, "none", False),  # conv1_jump_3
    (64, 128, 3, 2, 1, 8, 13, "elu", True),      # conv2_1 + broadcast skip
    (128, 128, 3, 1, 1, 4, 7, "none", False),    # conv3_0
    (3584, 512, 1, 1, 0, 1, 1, "elu", False),    # dense0
    (512, 128, 1, 1, 0, 1, 1, "none", False),    # dense1
]
ACT = {"none": (_capi.ACT_NONE, lambda t: t), "relu": (_capi.ACT_RELU, torch.relu), "elu": (_capi.ACT_ELU, F.elu)}


@pytest.mark.parametrize("g", GEOMS, ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}s{g[3]}_{g[5]}x{g[6]}")
@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("impl", [1, 0], ids=["tma", "gather"])
def test_conv_layer_vs_float64(built, g, precise, impl):
    """impl 1: the TMA im2col kernel (csrc/agx_conv_tma.cu) where the geometry allows — every spatial layer of both encoders; impl 0:
    the cp.async gather kernel (csrc/agx_conv.cu), which also serves the dense layers under impl 1."""
    Cin, Cout, k, s, p, H, W, act, with_res = g
    if impl == 0 and H == 1:
        pytest.skip("dense layers run on the gather kernel under both settings")
    _capi.check(_capi.load().agx_set_option(b"conv_impl", impl), "conv_impl")
    torch.manual_seed(Cin * 7 + Cout)
    N = 37 if H > 1 else 300  # M = N * Ho * Wo is not a multiple of 128: partial last tile
    conv = nn.Conv2d(Cin, Cout, k, stride=s, padding=p).cuda()
    x = torch.randn(N, Cin, H, W, device="cuda")
    L = T._conv_weight(conv, precise)
    y_ref = F.conv2d(x.double(), conv.weight.double(), conv.bias.double(), stride=s, padding=p)
    res = None
    if with_res:  # a skip tensor one column wider (cropped) or one column narrower (broadcast of its last column), like the VAE's
        wr = y_ref.shape[3] + (1 if Cout == 64 else -1)
        res = torch.randn(N, y_ref.shape[2], wr, Cout, device="cuda")
        r = res.permute(0, 3, 1, 2).double()
        d = (wr - y_ref.shape[3]) // 2
        y_ref = y_ref + r[:, :, :, d:d + y_ref.shape[3]]
    scale = shift = None
    y_ref = ACT[act][1](y_ref)
    if act == "relu":
        scale, shift = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
        y_ref = y_ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    try:
        y = T.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous(), L, ACT[act][0], res=res, scale=scale, shift=shift)
        torch.cuda.synchronize()
    finally:
        _capi.load().agx_set_option(b"conv_impl", 1)
    y_ref = y_ref.detach()
    err = float((y.permute(0, 3, 1, 2).double() - y_ref).abs().max())
    ref_scale = float(y_ref.abs().max())
    # 3xTF32 products are fp32-exact to ~2^-22; what remains is the tensor core's truncating accumulator (split over 2-4 chains by the kernel)
    K = Cin * (k * k if isinstance(k, int) else k[0] * k[1])
    assert err <= ((1e-5 if K <= 2048 else 3e-5) if precise else 4e-3) * max(1.0, ref_scale), (err, ref_scale, K)


@pytest.mark.parametrize("Cout,k,s,p,H,W", [(16, 5, 2, 2, 212, 120), (32, 5, 2, 2, 120, 212)])
@pytest.mark.parametrize("impl", [1, 3, 2, 0], ids=["const-bank-x2", "const-bank", "tensor-core", "generic"])
@pytest.mark.parametrize("norm", [True, False])
def test_first_layer_and_resize_vs_torch(built, Cout, k, s, p, H, W, impl, norm):
    """conv_first 1 (default): unrolled direct kernel with constant-bank weights, two pixels per thread; 3: one pixel per thread;
    2: operand rows built from a TMA-loaded strip, tcgen05 3xTF32 (csrc/agx_conv_tma.cu); 0: the generic direct fp32 kernel —
    1, 3 and 0 must agree bit for bit (same fmaf chain)."""
    torch.manual_seed(1)
    lib = _capi.load()
    conv = nn.Conv2d(1, Cout, k, stride=s, padding=p).cuda()
    img = torch.rand(9, H, W, device="cuda") * 10
    mean, rstd = torch.rand(H * W, device="cuda") * 5, torch.rand(H * W, device="cuda") + 0.2
    scale, shift = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
    _capi.check(lib.agx_set_option(b"conv_first", impl), "conv_first")
    try:
        y = T.conv2d_first(img, conv, _capi.ACT_RELU, mean if norm else None, rstd if norm else None, scale, shift)
        torch.cuda.synchronize()
        if impl in (1, 3):
            lib.agx_set_option(b"conv_first", 0)
            y0 = T.conv2d_first(img, conv, _capi.ACT_RELU, mean if norm else None, rstd if norm else None, scale, shift)
            assert torch.equal(y, y0)
    finally:
        lib.agx_set_option(b"conv_first", 1)
    xn = torch.clamp((img - mean.view(H, W)) * rstd.view(H, W), -5, 5) if norm else img
    with torch.no_grad():
        ref = torch.relu(F.conv2d(xn.unsqueeze(1).double(), conv.weight.double(), conv.bias.double(), stride=s, padding=p))
        ref = ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    assert float((y.permute(0, 3, 1, 2).double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    out = torch.empty(9, W, H, device="cuda")
    _capi.check(lib.agx_resize_bilinear(img.data_ptr(), out.data_ptr(), 9, H, W, W, H, None))
    ref = F.interpolate(img.unsqueeze(1), (W, H), mode="bilinear", align_corners=False).squeeze(1)
    assert float((out - ref).abs().max()) < 5e-5 * float(ref.abs().max())  # torch's GPU kernel interpolates with a different association


def test_cnn_encoder_tc_vs_float64_and_the_fused_kernel(built):
    from airgym_b200.lib.network.cnn import CNNFeatureExtractor, native_encode

    torch.manual_seed(2)
    net = CNNFeatureExtractor(30).cuda().eval()
    with torch.no_grad():
        for bn in (net.features[2], net.features[5], net.features[8]):
            bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    img = torch.rand(67, 1, 212, 120, device="cuda") * 9
    mean, rstd = torch.rand(212 * 120, device="cuda") * 4, torch.rand(212 * 120, device="cuda") * 0.5 + 0.2
    with torch.no_grad():
        tc = native_encode(net, img, mean, rstd, impl="tc")
        fused = native_encode(net, img, mean, rstd, impl="fused")
        xn = torch.clamp((img - mean.view(1, 1, 212, 120)) * rstd.view(1, 1, 212, 120), -5, 5)
        ref = net.double().forward_torch(xn.double())
        net.float()
    sc = max(1.0, float(ref.abs().max()))
    assert float((tc.double() - ref).abs().max()) <= 2e-5 * sc
    assert float((tc - fused).abs().max()) <= 2e-5 * sc
    wide = torch.zeros(67, 46, device="cuda")
    with torch.no_grad():
        native_encode(net, img, mean, rstd, out=wide[:, 16:], impl="tc")  # straight into a trunk-input row
    assert torch.equal(wide[:, 16:], tc) and float(wide[:, :16].abs().max()) == 0.0


def test_cnn_encoder_train_mode_batchnorm_vs_torch(built):
    """model.train() without autograd: conv -> ReLU -> BatchNorm with BATCH statistics (agx_col_sums + agx_bn_train between the layer
    kernels), running_mean / running_var / num_batches_tracked updated like torch's modules — against the torch module in float64,
    two consecutive calls (the second starts from the first one's running statistics), then eval mode on the updated statistics."""
    import copy

    from airgym_b200.lib.network.cnn import CNNFeatureExtractor

    torch.manual_seed(5)
    net = CNNFeatureExtractor(30).cuda()
    with torch.no_grad():
        for bn in (net.features[2], net.features[5], net.features[8]):
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    ref = copy.deepcopy(net).double()
    net.train(); ref.train()
    for it in range(2):
        img = torch.rand(41, 1, 212, 120, device="cuda") * (9 + it)
        with torch.no_grad():
            got = net(img)
            want = ref.forward_torch(img.double())
        sc = max(1.0, float(want.abs().max()))
        assert float((got.double() - want).abs().max()) <= 3e-5 * sc, it
        for k in (2, 5, 8):
            a, b = net.features[k], ref.features[k]
            assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == it + 1
            assert float((a.running_mean.double() - b.running_mean).abs().max()) <= 1e-5 * max(1.0, float(b.running_mean.abs().max()))
            assert float((a.running_var.double() - b.running_var).abs().max()) <= 1e-5 * max(1.0, float(b.running_var.abs().max()))
    net.eval(); ref.eval()
    with torch.no_grad():
        got, want = net(img), ref.forward_torch(img.double())
    assert float((got.double() - want).abs().max()) <= 3e-5 * max(1.0, float(want.abs().max()))


def test_cnn_encoder_replays_from_a_cuda_graph(built):
    """The layer sequence is capturable (the first layer's constant-bank image travels as a memcpy node, the tensor maps as kernel
    arguments): a replay after the weights and the images changed in place reproduces the eager result bit for bit."""
    from airgym_b200.lib.network.cnn import CNNFeatureExtractor, native_encode

    torch.manual_seed(7)
    net = CNNFeatureExtractor(30).cuda().eval()
    img = torch.rand(300, 1, 212, 120, device="cuda") * 9
    mean, rstd = torch.rand(212 * 120, device="cuda") * 4, torch.rand(212 * 120, device="cuda") * 0.5 + 0.2
    out = torch.zeros(300, 30, device="cuda")
    with torch.no_grad():
        native_encode(net, img, mean, rstd, out=out)  # warm-up: weight preparation cached, attributes set
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            native_encode(net, img, mean, rstd, out=out)
        img.mul_(0.5).add_(1.0)
        net.features[0].weight.mul_(1.5)  # first-layer weights are re-read from the parameter on every replay
        net.features[0].bias.add_(0.1)
        g.replay()
        torch.cuda.synchronize()
        got = out.clone()
        want = native_encode(net, img, mean, rstd)
    assert torch.equal(got, want)


def test_vae_encoder_tc_vs_mirror_and_golden(built):
    from airgym_b200.lib.network.vae_image_encoder import VAEImageEncoder
    from tests.util_vae import procedural_images, procedural_state

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae_encoder.npz"))
    cfg = {"latent_dims": 64, "image_res": [120, 212], "interpolation_mode": "bilinear", "return_sampled_latent": False, "allow_random_init": True}
    enc = VAEImageEncoder(cfg)
    shapes = {"encoder." + k: tuple(v.shape) for k, v in enc.encoder.state_dict().items()}
    dec = {"img_decoder.dense.weight": (512, 64), "img_decoder.dense.bias": (512,), "img_decoder.dense1.weight": (11648, 512),
           "img_decoder.dense1.bias": (11648,), "img_decoder.deconv1.weight": (128, 128, 3, 3), "img_decoder.deconv1.bias": (128,),
           "img_decoder.deconv2.weight": (128, 64, 4, 4), "img_decoder.deconv2.bias": (64,), "img_decoder.deconv3.weight": (64, 32, 4, 4),
           "img_decoder.deconv3.bias": (32,), "img_decoder.deconv4.weight": (32, 16, 4, 4), "img_decoder.deconv4.bias": (16,),
           "img_decoder.deconv5.weight": (16, 1, 4, 4), "img_decoder.deconv5.bias": (1,)}
    enc.load_weights(procedural_state({**shapes, **dec}))
    enc = enc.cuda()
    imgs = procedural_images(4).cuda()
    z = enc.encode(imgs)  # native: tc_encoders.vae_encode
    assert z.shape == (4, 64)
    assert np.abs(z.cpu().numpy() - G["procedural"]).max() <= 2e-5 * max(1.0, np.abs(G["procedural"]).max())  # the reference's own modules
    # a larger batch in the camera's own layout [N,1,212,120] (goes through the resize), vs the torch mirror in float64
    torch.manual_seed(3)
    big = torch.rand(130, 1, 212, 120, device="cuda")
    z2 = enc.encode(big)
    with torch.no_grad():
        r = F.interpolate(big.double(), (120, 212), mode="bilinear")
        ref = enc.encoder.double()(r)[:, :64]
    assert float((z2.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
