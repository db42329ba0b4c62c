"""Host-side logic added in round 2 that needs no GPU: weight preparation and skip-branch geometry of the tensor-core encoders,
struct mirrors of the new ABI blocks, the packed results block layout, the PPO config plumbing for the multi-GPU collectives."""
import ctypes as C

import pytest
import torch

from airgym_b200 import _capi
from airgym_b200.lib.network import tc_encoders as T


def test_split_tf32_is_exact_and_hi_is_tf32():
    torch.manual_seed(0)
    w = torch.randn(4096) * torch.logspace(-6, 3, 4096)
    hi, lo = T.split_tf32(w)
    assert torch.equal(hi + lo, w)  # the remainder is exact in fp32
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0  # 13 low mantissa bits clear: what kind::tf32 reads is all there is
    assert float((lo.abs() / w.abs()).max()) <= 2.0 ** -11 + 1e-9  # round to nearest


@pytest.mark.parametrize("t_len,ref_len", [(27, 26), (6, 7), (4, 4), (15, 15), (9, 7), (5, 6)])
def test_crop_window_replicates_python_slicing_of_crop_like(t_len, ref_len):
    """ImgEncoder._crop_like: t[d : d + ref_len], d = (t_len - ref_len) // 2 — incl. the negative start that leaves ONE column, which
    then broadcasts in the addition (VAE.py:102-108 as the reference's conv1_jump_3 / conv2_1 shapes exercise it)."""
    t = torch.arange(t_len)
    d = (t_len - ref_len) // 2
    sl = t[d:d + ref_len]
    start, step = T._crop_window(t_len, ref_len)
    want = sl if sl.numel() == ref_len else sl.expand(ref_len)
    got = torch.tensor([t[start + i * step] for i in range(ref_len)])
    assert torch.equal(got, want)


def test_crop_window_rejects_shapes_that_do_not_broadcast():
    with pytest.raises(ValueError):
        T._crop_window(5, 8)  # t[-2:6] leaves 2 columns: neither a crop nor a broadcast


def test_conv_weight_rows_are_in_gather_order():
    conv = torch.nn.Conv2d(8, 32, (3, 5), stride=2, padding=(1, 2))
    L = T._conv_weight(conv, True)
    rows = (L["hi"] + L["lo"]).reshape(32, 3, 5, 8)
    assert torch.equal(rows.permute(0, 3, 1, 2), conv.weight.detach())  # K index = (ky * kw + kx) * Cin + c
    assert L["Cin"] == 8 and L["Cout"] == 32 and L["k"] == (3, 5) and L["s"] == (2, 2) and L["p"] == (1, 2)
    assert T._out_hw(53, 30, (3, 3), (2, 2), (1, 1)) == (27, 15) and T._out_hw(15, 26, (5, 5), (4, 4), (2, 1)) == (4, 6)


def test_new_struct_mirrors_match_the_library(built):
    lib = _capi.load()
    assert lib.agx_sizeof_policy_io() == C.sizeof(_capi.AgxPolicyIO)
    assert lib.agx_sizeof_post_io() == C.sizeof(_capi.AgxPostIO)
    assert lib.agx_sizeof_conv_params() == C.sizeof(_capi.AgxConvParams)
    assert lib.agx_sizeof_conv_first_params() == C.sizeof(_capi.AgxConvFirstParams)
    assert lib.agx_sizeof_step_io() == C.sizeof(_capi.AgxStepIO) and lib.agx_sizeof_render_io() == C.sizeof(_capi.AgxRenderIO)
    assert C.sizeof(_capi.AgxComm) == 16 + 8 * _capi.AGX_COMM_MAX_RANKS
    # error paths return codes, never raise, and need no device
    assert lib.agx_comm_region_bytes(0, 1024) == -1 and lib.agx_comm_region_bytes(8, 1024) == 256 + 2 * 8 * 2 * 1024
    assert lib.agx_comm_allreduce(None, None, 4, 0, None) == -1
    assert lib.agx_policy_step(None, None, 4, None, None) == -1
    assert lib.agx_rollout_post(None, 4, None) == -1
    assert lib.agx_conv2d_nhwc(None, None) == -1 and lib.agx_conv2d_first(None, None) == -1
    assert lib.agx_observe(None, 4, None, 1, None) == -1
    p = _capi.AgxMlpParams()
    assert lib.agx_mlp_train_supported(C.byref(p)) == 0
    assert lib.agx_col_sums(None, 4, 4, 4, None, None, None) == -1 and lib.agx_rms_merge(None, 4, 10.0, None, None, None, None) == -1
    # second half of round 2: loss block of the fused loss + backward, train-mode BatchNorm, the option keys of the TMA kernels
    assert lib.agx_sizeof_loss_io() == C.sizeof(_capi.AgxLossIO)
    assert lib.agx_ppo_loss_backward_train(None, None, None, None, 128, *([None] * 10)) == -1
    assert lib.agx_bn_train(None, 4, 16, None, None, None, 1e-5, 0.1, None, None, None) == -1
    for key, good, bad in ((b"conv_impl", 1, 2), (b"conv_first", 1, 4), (b"conv_spp", 0, 5), (b"conv_stages", 0, 1), (b"mlp_wgrad_tma", 1, None)):
        assert lib.agx_set_option(key, good) == 0
        if bad is not None:
            assert lib.agx_set_option(key, bad) == -1
            assert lib.agx_set_option(key, good) == 0


def test_encoder_precision_switch_reaches_the_modules():
    """network.cnn.encoder_precise: False = single-pass TF32 convolutions (cuDNN's default precision); default 3xTF32."""
    import copy

    from airgym_b200.lib.config import default_ppo_config
    from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    params = copy.deepcopy(default_ppo_config("planning")["params"])
    shape = {"image": (1, 212, 120), "observation": (16,)}
    m = ModelA2CContinuousLogStd(params, {"actions_num": 4, "input_shape": shape})
    assert m.actor_cnn.encoder_precise is True
    params["network"]["cnn"]["encoder_precise"] = False
    m = ModelA2CContinuousLogStd(params, {"actions_num": 4, "input_shape": shape})
    assert m.actor_cnn.encoder_precise is False


def test_train_padding_keeps_a_spare_input_plane():
    from airgym_b200.lib.config import default_ppo_config
    from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    for obs, pad_plain, pad_train in ((18, 32, 32), (48, 48, 64), (46, 48, 48), (80, 80, 96), (16, 16, 32)):
        m = ModelA2CContinuousLogStd(default_ppo_config("hovering")["params"], {"actions_num": 4, "input_shape": (obs,)})
        assert m.fused_params().in_pad == pad_plain and m.fused_params(train=True).in_pad == pad_train
        assert m.fused_params(train=True).in_pad > obs  # plane `in_dim` = ones: the bias-gradient column
