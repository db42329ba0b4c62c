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
        mean, rstd = torch.rand(212 * 120, device="cuda"), torch.rand(212 * 120, device="cuda") + 0.5
        for impl in ("fused", "tc"):  # fused: one persistent fp32-FMA kernel; tc: layer by layer, conv2 / conv3 on tcgen05 (3xTF32)
            got = native_encode(net, x[:256], impl=impl)
            out[impl + "_max_abs_err_vs_cudnn_fp32"] = float((got - ref).abs().max())
            out[impl + "_ms"] = timed(lambda: native_encode(net, x, impl=impl))
            out[impl + "_fused_norm_ms"] = timed(lambda: native_encode(net, x, mean, rstd, impl=impl))
        from airgym_b200.lib.network import tc_encoders as T
        out["tc_single_pass_tf32_ms"] = timed(lambda: T.cnn_encode(net, x, precise=False))
        # per layer (2048-image chunk)
        xc = x[:2048].reshape(-1, 212, 120)
        from airgym_b200 import _capi as K
        W_ = T._prep_cnn(net, True)
        a1 = T.conv2d_first(xc, net.features[0], K.ACT_RELU, None, None, W_["s1"], W_["t1"])
        a2 = T.conv2d_nhwc(a1, W_["c2"], K.ACT_RELU, scale=W_["s2"], shift=W_["t2"])
        lib_ = K.load()
        lib_.agx_set_option(b"conv_impl", 0)
        out["gather_layers_ms_per_2048"] = {
            "conv1_direct": timed(lambda: T.conv2d_first(xc, net.features[0], K.ACT_RELU, None, None, W_["s1"], W_["t1"])),
            "conv2_cp_async": timed(lambda: T.conv2d_nhwc(a1, W_["c2"], K.ACT_RELU, scale=W_["s2"], shift=W_["t2"])),
            "conv3_cp_async": timed(lambda: T.conv2d_nhwc(a2, W_["c3"], K.ACT_RELU, scale=W_["s3"], shift=W_["t3"]))}
        lib_.agx_set_option(b"conv_impl", 1)
        for mode, nm in ((2, "tcgen05"), (3, "const_bank_1px"), (1, "const_bank_2px")):
            lib_.agx_set_option(b"conv_first", mode)
            out.setdefault("conv1_ms_per_2048", {})[nm] = timed(lambda: T.conv2d_first(xc, net.features[0], K.ACT_RELU, None, None, W_["s1"], W_["t1"]))
            out["conv1_ms_per_2048"][nm + "_fused_norm"] = timed(lambda: T.conv2d_first(xc, net.features[0], K.ACT_RELU, mean, rstd, W_["s1"], W_["t1"]))
        a3 = T.conv2d_nhwc(a2, W_["c3"], K.ACT_RELU, scale=W_["s3"], shift=W_["t3"])
        fo = torch.empty(a3.shape[0], 30, device="cuda")
        out["pool_fc_ms_per_2048"] = timed(lambda: lib_.agx_pool_fc(a3.data_ptr(), a3.shape[0], a3.shape[1] * a3.shape[2], a3.shape[3], W_["wfc"].data_ptr(),
                                                                     W_["bfc"].data_ptr(), 30, fo.data_ptr(), 30, None))
        out["tc_layers_ms_per_2048"] = {
            "conv1": timed(lambda: T.conv2d_first(xc, net.features[0], K.ACT_RELU, None, None, W_["s1"], W_["t1"])),
            "conv1_fused_norm": timed(lambda: T.conv2d_first(xc, net.features[0], K.ACT_RELU, mean, rstd, W_["s1"], W_["t1"])),
            "conv2_tcgen05": timed(lambda: T.conv2d_nhwc(a1, W_["c2"], K.ACT_RELU, scale=W_["s2"], shift=W_["t2"])),
            "conv3_tcgen05": timed(lambda: T.conv2d_nhwc(a2, W_["c3"], K.ACT_RELU, scale=W_["s3"], shift=W_["t3"]))}
        # the VAE ImgEncoder: libagx layers vs torch / cuDNN
        from airgym_b200.lib.network.vae_image_encoder import VAEImageEncoder
        vae = VAEImageEncoder({"latent_dims": 64, "image_res": [120, 212], "interpolation_mode": "bilinear", "allow_random_init": True}).cuda()
        nv = min(a.n, 8192)
        out["vae_n"] = nv
        zn = vae.encode(x[:64])
        vae.native = False
        zt = vae.encode(x[:64])
        out["vae_native_max_abs_err_vs_cudnn_fp32"] = float((zn - zt).abs().max())
        out["vae_cudnn_fp32_ms"] = timed(lambda: vae.encode(x[:nv]), iters=3)
        torch.backends.cudnn.allow_tf32 = True
        out["vae_cudnn_tf32_ms"] = timed(lambda: vae.encode(x[:nv]), iters=3)
        torch.backends.cudnn.allow_tf32 = False
        vae.native = True
        out["vae_native_ms"] = timed(lambda: vae.encode(x[:nv]), iters=3)
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
