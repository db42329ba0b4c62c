"""CNNFeatureExtractor — the depth-image encoder of the avoid/planning policies, with the reference's module layout so that
its checkpoints load key for key (lib/network/cnn.py:3-33: `features.{0,3,6}` convolutions 1→16 (5x5, s2) →32 (3x3, s2) →64
(3x3, s2), each followed by ReLU then BatchNorm2d `features.{2,5,8}`, global average pool, `fc` 64→feature_dim).

Inference (eval-mode BatchNorm — the only mode the trainer uses: the reference's encoder never receives a gradient, DESIGN.md
§4.4) runs on the libagx kernel `agx_cnn_encode`: one persistent launch, the whole network per env in
shared memory, fp32 FMA arithmetic, optional fused per-pixel input normalisation.  In train mode without autograd (what the
reference's update pass amounts to: its encoder output is cut from the graph) the forward runs natively too, with BATCH statistics
and the running-statistics update (tc_encoders.cnn_encode_train: agx_col_sums + agx_bn_train between the convolution layers);
with autograd enabled the module falls back to torch's library convolutions."""
import ctypes as C

import torch
import torch.nn as nn

from airgym_b200 import _capi


def encoder_params(net):
    """AgxCnnParams over the module's own parameter storage (+ the tensors that must outlive the call).  The eval-mode
    BatchNorm of each block is folded to scale / shift: s = weight / sqrt(running_var + eps), t = bias - running_mean * s."""
    f, keep = net.features, []
    p = _capi.AgxCnnParams()
    for i, (conv, bn) in enumerate(((f[0], f[2]), (f[3], f[5]), (f[6], f[8])), start=1):
        with torch.no_grad():
            s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
            t = (bn.bias - bn.running_mean * s).float().contiguous()
        w, b = conv.weight.detach().contiguous(), conv.bias.detach().contiguous()
        keep += [w, b, s, t]
        for name, ten in (("w", w), ("b", b), ("s", s), ("t", t)):
            setattr(p, f"{name}{i}", ten.data_ptr())
    wfc, bfc = net.fc.weight.detach().contiguous(), net.fc.bias.detach().contiguous()
    keep += [wfc, bfc]
    p.wfc, p.bfc, p.feature_dim = wfc.data_ptr(), bfc.data_ptr(), net.fc.out_features
    return p, keep


# which libagx implementation native_encode uses: "tc" = layer by layer, conv2 / conv3 as implicit GEMMs on tcgen05 with a 3xTF32
# split (tc_encoders.cnn_encode); "fused" = the single persistent fp32-FMA kernel agx_cnn_encode (whole network per env in shared memory)
ENCODER_IMPL = "tc"
# 3xTF32 operand split (fp32-level results, the default) or single-pass TF32 — the precision torch / cuDNN convolutions run at by default
# on this GPU (torch.backends.cudnn.allow_tf32 = True), i.e. what the reference's own encoder forward computes; ~25 % faster
# (CNN 3.6 vs 4.8 ms per 8192 images).  Set through the network config key `encoder_precise: False` (model) or this module attribute.
ENCODER_PRECISE = True


def native_encode(net, x, px_mean=None, px_rstd=None, out=None, lib=None, impl=None, precise=None):
    """features [N, feature_dim] of images x [N,1,212,120] through agx_cnn_encode; px_mean / px_rstd [212*120] fuse the
    RunningMeanStd normalisation clamp((x - mean) * rstd, +-5) into the image load.  `out` may be a column slice of a wider
    row-major buffer (the trunk-input rows); `lib` substitutes another build of libagx (tuning variants, scripts/enc_bench.py)."""
    assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (1, _capi.AGX_CAM_W, _capi.AGX_CAM_H), x.shape
    x = x.contiguous()
    n = x.shape[0]
    if out is None:
        out = torch.empty(n, net.fc.out_features, device=x.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.shape == (n, net.fc.out_features) and out.stride(1) == 1
    if (impl or ENCODER_IMPL) == "tc" and lib is None and n > 0:
        from .tc_encoders import cnn_encode
        precise = getattr(net, "encoder_precise", ENCODER_PRECISE) if precise is None else precise
        return cnn_encode(net, x, px_mean, px_rstd, out, precise=bool(precise))
    p, keep = encoder_params(net)
    if px_mean is not None:
        px_mean, px_rstd = px_mean.float().contiguous(), px_rstd.float().contiguous()
        keep += [px_mean, px_rstd]
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    _capi.check((lib or _capi.load()).agx_cnn_encode(C.byref(p), n, x.data_ptr(), px_mean.data_ptr() if px_mean is not None else None,
                                            px_rstd.data_ptr() if px_rstd is not None else None, out.data_ptr(),
                                            out.stride(0) if n > 0 else net.fc.out_features, st), "agx_cnn_encode")
    return out  # temporaries were allocated on the launch stream: the caching allocator's stream order keeps them alive long enough


class CNNFeatureExtractor(nn.Module):
    def __init__(self, feature_dim=12):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(1, 16, kernel_size=5, stride=2, padding=2), nn.ReLU(), nn.BatchNorm2d(16),   # (16, 106, 60)
            nn.Conv2d(16, 32, kernel_size=3, stride=2, padding=1), nn.ReLU(), nn.BatchNorm2d(32),  # (32, 53, 30)
            nn.Conv2d(32, 64, kernel_size=3, stride=2, padding=1), nn.ReLU(), nn.BatchNorm2d(64),  # (64, 27, 15)
            nn.AdaptiveAvgPool2d((1, 1)),
        )
        self.fc = nn.Linear(64, feature_dim)

    def native_ok(self, x):
        return (x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()
                and tuple(x.shape[1:]) == (1, _capi.AGX_CAM_W, _capi.AGX_CAM_H) and self.fc.out_features <= 64)

    def forward_torch(self, x):
        x = self.features(x)
        return self.fc(x.view(x.size(0), -1))

    def forward(self, x):
        if self.native_ok(x):
            if self.training:  # BatchNorm with batch statistics + running-statistics update, no autograd
                from .tc_encoders import cnn_encode_train
                return cnn_encode_train(self, x)
            return native_encode(self, x)
        return self.forward_torch(x)
