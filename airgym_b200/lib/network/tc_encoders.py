"""Depth-image encoders on the tensor cores: the layer sequences of CNNFeatureExtractor (reference lib/network/cnn.py:3-33) and of the
frozen VAE ImgEncoder (lib/network/VAE.py:52-148 behind lib/network/vae_image_encoder.py:34-53) driven layer by layer through
libagx's channels-last convolution kernels (csrc/agx_conv.cu: tcgen05 implicit GEMM with a 3xTF32 operand split, a direct first
layer, bilinear resize, pool + fc).  Host side = geometry and weight preparation only; no torch / cuDNN op touches the images.

Weights are prepared once per parameter version: OIHW → O(HW)I rows (K = kh*kw*Cin, the order the kernel gathers its im2col
operand in), split into w_hi (exactly representable in tf32) + w_lo (the fp32 remainder)."""
import ctypes as C

import torch

from ... import _capi

CHUNK = 8192  # images per pass: the widest intermediate (first-layer output: 0.4 MB per image for the CNN, 0.8 MB for the VAE) stays below 7 GB


def split_tf32(w):
    """w = hi + lo with hi carrying the top 19 bits (round-to-nearest on the dropped 13), both float32."""
    bits = w.contiguous().view(torch.int32)
    hi = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi.contiguous(), (w - hi).contiguous()


class _Prepared:
    """Per-module cache of kernel-ready weights, refreshed when any parameter / buffer version changes."""

    def __init__(self):
        self.key, self.data = None, None

    def get(self, tensors, build):
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self.key:
            self.data, self.key = build(), key
        return self.data


def _conv_weight(conv, precise):
    w = conv.weight.detach().float()
    o, i, kh, kw = w.shape
    rows = w.permute(0, 2, 3, 1).reshape(o, kh * kw * i).contiguous()
    hi, lo = split_tf32(rows)
    return {"hi": hi, "lo": lo if precise else None, "bias": conv.bias.detach().float().contiguous() if conv.bias is not None else None,
            "Cin": i, "Cout": o, "k": (kh, kw), "s": tuple(conv.stride), "p": tuple(conv.padding)}


def _out_hw(h, w, k, s, p):
    return (h + 2 * p[0] - k[0]) // s[0] + 1, (w + 2 * p[1] - k[1]) // s[1] + 1


def _crop_window(t_len, ref_len):
    """Python-slice semantics of ImgEncoder._crop_like along one axis: t[d : d + ref_len] with d = (t_len - ref_len) // 2 — a negative
    d wraps like a negative slice start, and a one-element result broadcasts in the addition that follows.  → (start, step)."""
    d = (t_len - ref_len) // 2
    start = d if d >= 0 else max(t_len + d, 0)
    stop = min(d + ref_len, t_len) if d + ref_len >= 0 else max(t_len + d + ref_len, 0)
    n = max(stop - start, 0)
    if n == ref_len:
        return start, 1
    if n == 1:
        return start, 0
    raise ValueError(f"skip branch of length {t_len} cannot be cropped / broadcast to {ref_len}")


def conv2d_nhwc(x, L, act, res=None, scale=None, shift=None, stream=None):
    """x [N,H,W,Cin] → [N,Ho,Wo,Cout] through agx_conv2d_nhwc; res = NHWC skip tensor added before the activation."""
    lib = _capi.load()
    N, H, W, Cin = x.shape
    assert Cin == L["Cin"] and x.is_contiguous()
    Ho, Wo = _out_hw(H, W, L["k"], L["s"], L["p"])
    y = torch.empty(N, Ho, Wo, L["Cout"], device=x.device, dtype=torch.float32)
    P = _capi.AgxConvParams()
    P.x, P.N, P.H, P.W, P.Cin = x.data_ptr(), N, H, W, Cin
    P.w_hi, P.w_lo = L["hi"].data_ptr(), (L["lo"].data_ptr() if L["lo"] is not None else None)
    P.bias = L["bias"].data_ptr() if L["bias"] is not None else None
    if scale is not None:
        P.scale, P.shift = scale.data_ptr(), shift.data_ptr()
    if res is not None:
        assert res.is_contiguous() and res.shape[0] == N and res.shape[3] == L["Cout"]
        P.res, P.rH, P.rW = res.data_ptr(), res.shape[1], res.shape[2]
        (P.ry0, P.rsy), (P.rx0, P.rsx) = _crop_window(res.shape[1], Ho), _crop_window(res.shape[2], Wo)
    P.y, P.Ho, P.Wo, P.Cout = y.data_ptr(), Ho, Wo, L["Cout"]
    P.kh, P.kw, P.sy, P.sx, P.py, P.px, P.act = L["k"][0], L["k"][1], L["s"][0], L["s"][1], L["p"][0], L["p"][1], act
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream)
    _capi.check(lib.agx_conv2d_nhwc(C.byref(P), st), "agx_conv2d_nhwc")
    return y


def conv2d_first(img, conv, act, px_mean=None, px_rstd=None, scale=None, shift=None):
    """img [N,H,W] (one channel) → NHWC [N,Ho,Wo,Cout] through the direct first-layer kernel."""
    lib = _capi.load()
    N, H, W = img.shape
    w = conv.weight.detach().float().contiguous()
    b = conv.bias.detach().float().contiguous()
    o, _, kh, kw = w.shape
    Ho, Wo = _out_hw(H, W, (kh, kw), conv.stride, conv.padding)
    y = torch.empty(N, Ho, Wo, o, device=img.device, dtype=torch.float32)
    P = _capi.AgxConvFirstParams()
    P.x, P.N, P.H, P.W = img.data_ptr(), N, H, W
    P.w, P.bias = w.data_ptr(), b.data_ptr()
    if scale is not None:
        P.scale, P.shift = scale.data_ptr(), shift.data_ptr()
    if px_mean is not None:
        P.px_mean, P.px_rstd = px_mean.data_ptr(), px_rstd.data_ptr()
    P.y, P.Ho, P.Wo, P.Cout = y.data_ptr(), Ho, Wo, o
    P.kh, P.kw, P.sy, P.sx, P.py, P.px, P.act = kh, kw, conv.stride[0], conv.stride[1], conv.padding[0], conv.padding[1], act
    _capi.check(lib.agx_conv2d_first(C.byref(P), C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)), "agx_conv2d_first")
    return y


# ---- CNNFeatureExtractor --------------------------------------------------------------------------------------------------------
def _prep_cnn(net, precise):
    f = net.features
    out = {"c2": _conv_weight(f[3], precise), "c3": _conv_weight(f[6], precise)}
    for i, bn in ((1, f[2]), (2, f[5]), (3, f[8])):  # eval-mode BatchNorm folded to scale / shift (applied after the ReLU)
        s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
        out[f"s{i}"], out[f"t{i}"] = s, (bn.bias - bn.running_mean * s).detach().float().contiguous()
    out["wfc"], out["bfc"] = net.fc.weight.detach().float().contiguous(), net.fc.bias.detach().float().contiguous()
    return out


def cnn_encode(net, x, px_mean=None, px_rstd=None, out=None, precise=True):
    """features [N, feature_dim] of images x [N,1,212,120]: conv1 direct (+ fused input normalisation), conv2 / conv3 on tcgen05,
    pool + fc.  `out` may be a column slice of a wider row-major buffer."""
    lib = _capi.load()
    n = x.shape[0]
    H, W = x.shape[2], x.shape[3]
    if out is None:
        out = torch.empty(n, net.fc.out_features, device=x.device, dtype=torch.float32)
    if not hasattr(net, "_tc_prep"):
        net._tc_prep = {True: _Prepared(), False: _Prepared()}
    tensors = list(net.parameters()) + list(net.buffers())
    W_ = net._tc_prep[precise].get(tensors, lambda: _prep_cnn(net, precise))
    if px_mean is not None:
        px_mean, px_rstd = px_mean.float().contiguous(), px_rstd.float().contiguous()
    x = x.contiguous()
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    for i0 in range(0, n, CHUNK):
        xc = x[i0:i0 + CHUNK].reshape(-1, H, W)
        a1 = conv2d_first(xc, net.features[0], _capi.ACT_RELU, px_mean, px_rstd, W_["s1"], W_["t1"])
        a2 = conv2d_nhwc(a1, W_["c2"], _capi.ACT_RELU, scale=W_["s2"], shift=W_["t2"])
        a3 = conv2d_nhwc(a2, W_["c3"], _capi.ACT_RELU, scale=W_["s3"], shift=W_["t3"])
        o = out[i0:i0 + CHUNK]
        _capi.check(lib.agx_pool_fc(a3.data_ptr(), a3.shape[0], a3.shape[1] * a3.shape[2], a3.shape[3], W_["wfc"].data_ptr(),
                                    W_["bfc"].data_ptr(), net.fc.out_features, o.data_ptr(), o.stride(0), st), "agx_pool_fc")
    return out


def _bn_train_inplace(act, bn, ws):
    """Train-mode BatchNorm2d on a channels-last activation, in place: batch statistics over all N*H*W rows (agx_col_sums in float64,
    deterministic), normalisation + affine + running-statistics update in agx_bn_train; num_batches_tracked bumped like torch does."""
    lib = _capi.load()
    rows, Cc = act.shape[0] * act.shape[1] * act.shape[2], act.shape[3]
    st = C.c_void_p(torch.cuda.current_stream(act.device).cuda_stream)
    sums = torch.empty(2, Cc, device=act.device, dtype=torch.float64)
    _capi.check(lib.agx_col_sums(act.data_ptr(), rows, Cc, Cc, sums.data_ptr(), ws.data_ptr(), st), "agx_col_sums")
    track = bn.track_running_stats and bn.running_mean is not None
    if track:
        bn.num_batches_tracked += 1
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
    else:
        momentum = 0.0
    _capi.check(lib.agx_bn_train(act.data_ptr(), rows, Cc, sums.data_ptr(), bn.weight.data_ptr() if bn.affine else None,
                                 bn.bias.data_ptr() if bn.affine else None, float(bn.eps), float(momentum),
                                 bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None, st), "agx_bn_train")
    return act


def cnn_encode_train(net, x, px_mean=None, px_rstd=None, out=None):
    """CNNFeatureExtractor forward in TRAIN mode (BatchNorm with batch statistics over the whole call's batch, running statistics and
    num_batches_tracked updated) — what the reference's update pass runs on the minibatch images (lib/network/cnn.py:3-33 under
    model.train()).  Inference only (no autograd: the reference's encoder receives no gradient either, base_model.py:29-31).  The batch
    is processed in one piece (batch statistics do not chunk): 0.72 MB of activations per image."""
    lib = _capi.load()
    n, H, W = x.shape[0], x.shape[2], x.shape[3]
    if out is None:
        out = torch.empty(n, net.fc.out_features, device=x.device, dtype=torch.float32)
    f = net.features
    if not hasattr(net, "_tc_prep"):
        net._tc_prep = {True: _Prepared(), False: _Prepared()}
    W_ = net._tc_prep[True].get(list(net.parameters()) + list(net.buffers()), lambda: _prep_cnn(net, True))
    if px_mean is not None:
        px_mean, px_rstd = px_mean.float().contiguous(), px_rstd.float().contiguous()
    ws = torch.zeros(int(lib.agx_col_sums_workspace_doubles()), device=x.device, dtype=torch.float64)
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    a = conv2d_first(x.contiguous().reshape(-1, H, W), f[0], _capi.ACT_RELU, px_mean, px_rstd)
    a = _bn_train_inplace(a, f[2], ws)
    a = _bn_train_inplace(conv2d_nhwc(a, W_["c2"], _capi.ACT_RELU), f[5], ws)
    a = _bn_train_inplace(conv2d_nhwc(a, W_["c3"], _capi.ACT_RELU), f[8], ws)
    _capi.check(lib.agx_pool_fc(a.data_ptr(), a.shape[0], a.shape[1] * a.shape[2], a.shape[3], W_["wfc"].data_ptr(), W_["bfc"].data_ptr(),
                                net.fc.out_features, out.data_ptr(), out.stride(0), st), "agx_pool_fc")
    return out


# ---- VAE ImgEncoder ---------------------------------------------------------------------------------------------------------------
def _prep_vae(enc, precise):
    names = ("conv0_1", "conv1_0", "conv1_1", "conv2_0", "conv2_1", "conv3_0", "conv0_jump_2", "conv1_jump_3")
    out = {k: _conv_weight(getattr(enc, k), precise) for k in names}

    def dense(lin, perm=None):
        w = lin.weight.detach().float()
        if perm is not None:  # NCHW flatten order (c, y, x) of the reference → the NHWC order (y, x, c) the kernels produce
            c, hh, ww = perm
            w = w.reshape(w.shape[0], c, hh, ww).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
        hi, lo = split_tf32(w.contiguous())
        return {"hi": hi, "lo": lo if precise else None, "bias": lin.bias.detach().float().contiguous(), "Cin": w.shape[1], "Cout": w.shape[0],
                "k": (1, 1), "s": (1, 1), "p": (0, 0)}

    out["dense0"], out["dense1"] = dense(enc.dense0, (128, 4, 7)), dense(enc.dense1)
    return out


def vae_encode(enc, image_res, x, precise=True):
    """[means | log-variances] [N, 2*latent] of images x [N,1,H,W] through the ImgEncoder's layers (VAE.py:110-148), resized to
    `image_res` first like the reference's wrapper does (vae_image_encoder.py:36-38)."""
    lib = _capi.load()
    n = x.shape[0]
    if not hasattr(enc, "_tc_prep"):
        object.__setattr__(enc, "_tc_prep", {True: _Prepared(), False: _Prepared()})
    W_ = enc._tc_prep[precise].get(list(enc.parameters()), lambda: _prep_vae(enc, precise))
    out = torch.empty(n, enc.dense1.out_features, device=x.device, dtype=torch.float32)
    x = x.contiguous()
    H, W = x.shape[2], x.shape[3]
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    E, R = _capi.ACT_ELU, _capi.ACT_NONE
    for i0 in range(0, n, CHUNK):
        xc = x[i0:i0 + CHUNK].reshape(-1, H, W)
        if (H, W) != tuple(image_res):
            r = torch.empty(xc.shape[0], image_res[0], image_res[1], device=x.device, dtype=torch.float32)
            _capi.check(lib.agx_resize_bilinear(xc.data_ptr(), r.data_ptr(), xc.shape[0], H, W, image_res[0], image_res[1], st), "agx_resize_bilinear")
            xc = r
        t = conv2d_first(xc, enc.conv0, R)
        a = conv2d_nhwc(t, W_["conv0_1"], E)
        t = conv2d_nhwc(a, W_["conv1_0"], R)
        j2 = conv2d_nhwc(a, W_["conv0_jump_2"], R)
        b = conv2d_nhwc(t, W_["conv1_1"], E, res=j2)
        t = conv2d_nhwc(b, W_["conv2_0"], R)
        j3 = conv2d_nhwc(b, W_["conv1_jump_3"], R)
        c = conv2d_nhwc(t, W_["conv2_1"], E, res=j3)
        t = conv2d_nhwc(c, W_["conv3_0"], R)
        t = conv2d_nhwc(t.reshape(t.shape[0], 1, 1, -1), W_["dense0"], E)
        t = conv2d_nhwc(t, W_["dense1"], R)
        out[i0:i0 + CHUNK] = t.reshape(t.shape[0], -1)
    return out
