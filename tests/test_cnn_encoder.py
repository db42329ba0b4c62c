"""Depth-image encoder: the kernel's per-thread phase functions, replayed on the CPU by
tests/hostsim/hostsim_cnn.cpp, against torch's fp64 convolutions (the module layout of lib/network/cnn.py:3-33); the same
comparison for the sm_100a kernel through the C ABI is the gpu-marked test below."""
import ctypes as C

import numpy as np
import pytest
import torch

from airgym_b200 import _capi
from airgym_b200.lib.network.cnn import CNNFeatureExtractor, encoder_params, native_encode


def make_net(feature_dim, seed):
    torch.manual_seed(seed)
    net = CNNFeatureExtractor(feature_dim)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):  # a trained checkpoint's statistics are not the identity
                m.running_mean.uniform_(-0.5, 0.5)
                m.running_var.uniform_(0.3, 2.0)
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    return net.eval()


def make_images(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 1, _capi.AGX_CAM_W, _capi.AGX_CAM_H, generator=g)
    x[0, :, :, :] = 0.0          # an all-zero image: only biases and the padding path
    if n > 1:
        x[1, :, 0, :] = 7.0       # strong first / last rows and columns: the zero-padding borders of all three layers
        x[1, :, -1, :] = -7.0
        x[1, :, :, 0] = 5.0
        x[1, :, :, -1] = -5.0
    return x


def reference(net, x, mean=None, var=None, eps=1e-5):
    """fp64: RunningMeanStd.forward (running_mean_std.py:75-81) then the module as written."""
    x = x.double()
    if mean is not None:
        x = torch.clamp((x - mean.double()) / torch.sqrt(var.double() + eps), -5.0, 5.0)
    import copy
    with torch.no_grad():
        return copy.deepcopy(net).double().forward_torch(x)


@pytest.mark.parametrize("feature_dim,normalise", [(30, True), (30, False), (12, True), (64, False)])
def test_encoder_schedule_on_cpu_matches_torch(feature_dim, normalise):
    from tests.hostsim import driver
    lib = driver.build_cnn()
    net = make_net(feature_dim, seed=feature_dim)
    x = make_images(3, seed=5)
    mean = var = None
    args = [None, None]
    if normalise:
        g = torch.Generator().manual_seed(9)
        mean = torch.rand(1, _capi.AGX_CAM_W, _capi.AGX_CAM_H, generator=g)
        var = torch.rand(1, _capi.AGX_CAM_W, _capi.AGX_CAM_H, generator=g) * 0.2 + 0.01
        m32 = mean.reshape(-1).float().contiguous()
        r32 = torch.rsqrt(var.float() + 1e-5).reshape(-1).contiguous()
        args = [m32.data_ptr(), r32.data_ptr()]
    p, keep = encoder_params(net)
    ld = feature_dim + 3  # a wider row: the kernel must leave the other columns alone
    out = torch.full((x.shape[0], ld), 123.0)
    assert lib.hostsim_cnn_encode(C.byref(p), x.shape[0], x.data_ptr(), args[0], args[1], out.data_ptr(), ld) == 0
    ref = reference(net, x, mean, var)
    got = out[:, :feature_dim].double()
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max())), float((got - ref).abs().max())
    assert torch.equal(out[:, feature_dim:], torch.full((x.shape[0], 3), 123.0))


@pytest.mark.gpu
@pytest.mark.parametrize("impl", ["fused", "tc"])  # fused: agx_cnn_encode (fp32 FMA, one persistent kernel); tc: tcgen05 layers (3xTF32)
@pytest.mark.parametrize("n,feature_dim,normalise", [(3, 30, True), (149, 30, False), (300, 12, True), (1, 64, False)])
def test_encoder_kernel_matches_torch(built, n, feature_dim, normalise, impl):
    from airgym_b200.lib.network import cnn as cnn_mod
    cnn_mod.ENCODER_IMPL, saved = impl, cnn_mod.ENCODER_IMPL
    try:
        _encoder_kernel_matches_torch(n, feature_dim, normalise, impl)
    finally:
        cnn_mod.ENCODER_IMPL = saved


def _encoder_kernel_matches_torch(n, feature_dim, normalise, impl):
    net = make_net(feature_dim, seed=feature_dim).cuda()
    x = make_images(n, seed=n).cuda()
    mean = var = m32 = r32 = None
    if normalise:
        g = torch.Generator().manual_seed(9)
        mean = torch.rand(1, _capi.AGX_CAM_W, _capi.AGX_CAM_H, generator=g).cuda()
        var = (torch.rand(1, _capi.AGX_CAM_W, _capi.AGX_CAM_H, generator=g) * 0.2 + 0.01).cuda()
        m32, r32 = mean.reshape(-1), torch.rsqrt(var + 1e-5).reshape(-1)
    wide = torch.full((n, 16 + feature_dim), 123.0, device="cuda")
    with torch.no_grad():
        got = native_encode(net, x, m32, r32, out=wide[:, 16:])
    torch.cuda.synchronize()
    ref = reference(net, x, mean, var)
    err = float((got.double() - ref).abs().max())
    assert err < 2e-5 * max(1.0, float(ref.abs().max())), err
    assert torch.equal(wide[:, :16], torch.full((n, 16), 123.0, device="cuda"))
    with torch.no_grad():  # the module's own forward takes the same path in eval mode, and is deterministic
        again = net(x) if not normalise else native_encode(net, x, m32, r32)
    if not normalise:
        assert torch.equal(again, got)
    if impl != "fused":
        return
    # the CPU replay of the same phase functions agrees to rounding (fmaf on both sides)
    from tests.hostsim import driver
    lib = driver.build_cnn()
    k = min(n, 2)
    net_c = make_net(feature_dim, seed=feature_dim)
    p, keep = encoder_params(net_c)
    xc = x[:k].cpu().contiguous()
    out = torch.zeros(k, feature_dim)
    mc = m32.cpu().contiguous() if normalise else None
    rc = r32.cpu().contiguous() if normalise else None
    lib.hostsim_cnn_encode(C.byref(p), k, xc.data_ptr(), mc.data_ptr() if normalise else None, rc.data_ptr() if normalise else None,
                           out.data_ptr(), feature_dim)
    assert float((out - got[:k].cpu()).abs().max()) < 1e-6 * max(1.0, float(ref.abs().max()))
