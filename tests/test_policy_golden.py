"""The camera-task policy network against the REFERENCE's own model: tests/golden/
policy_planning_cnn.npz holds what lib/model/a2c_continuous_logstd_model.py computes with the shipped planning_cnn_rate.pth on
procedural inputs (tests/golden/make_golden_policy.py).  CPU: the module mirror and the CPU replay of the encoder kernel's
schedule; GPU: the libagx encoder kernel with the fused image normalisation and the model's product path."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from airgym_b200 import _capi
from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd
from airgym_b200.lib.network.cnn import encoder_params
from tests.util_vae import policy_inputs as inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_model():
    network = {"name": "actor_critic", "separate": False, "space": {"continuous": {"fixed_sigma": True}},
               "mlp": {"units": [64, 128, 64], "activation": "elu"}, "cnn": {"output_dim": 30}}
    params = {"network": network, "config": {"normalize_value": True, "normalize_input": True, "value_size": 1}}
    keys = {"actions_num": 4, "input_shape": {"image": (1, 212, 120), "observation": (16,)}, "num_seqs": 1, "value_size": 1,
            "normalize_value": True, "normalize_input": True}
    model = ModelA2CContinuousLogStd(params, keys)
    ck = torch.load(os.path.join(GOLD, "planning_cnn_rate_model.pth"), map_location="cpu", weights_only=False)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(ck["model"][k])
    return model.eval()


def golden():
    return {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(GOLD, "policy_planning_cnn.npz")).items()}


def test_inputs_are_the_golden_inputs():
    g = golden()
    img, obs = inputs()
    assert torch.equal(obs, g["observation"]) and float(g["image_scale"]) == 18.0 and img.shape == (6, 1, 212, 120)


def test_module_mirror_reproduces_the_reference_model():
    g, model = golden(), load_model()
    img, obs = inputs()
    with torch.no_grad():
        feat = model.encode_image(img)
        res = model({"is_train": True, "prev_actions": torch.zeros(6, 4), "obs": {"image": img, "observation": obs}})
    assert torch.allclose(feat, g["cnn_features"], rtol=1e-4, atol=2e-5), float((feat - g["cnn_features"]).abs().max())
    assert torch.allclose(res["mus"], g["mus"], rtol=1e-4, atol=1e-4), float((res["mus"] - g["mus"]).abs().max())
    assert torch.allclose(res["values"], g["values"], rtol=1e-4, atol=1e-4)
    assert torch.allclose(res["sigmas"], g["sigmas"], rtol=1e-6, atol=1e-7)


def test_encoder_kernel_schedule_reproduces_the_reference_features():
    """The encoder kernel's phase functions (CPU replay) with the shipped weights and the fused per-pixel normalisation."""
    from tests.hostsim import driver
    lib = driver.build_cnn()
    g, model = golden(), load_model()
    img, _ = inputs()
    rms = model.running_mean_std.running_mean_std["image"]
    mean = rms.running_mean.float().reshape(-1).contiguous()
    rstd = torch.rsqrt(rms.running_var.float() + rms.epsilon).reshape(-1).contiguous()
    p, keep = encoder_params(model.actor_cnn)
    out = torch.zeros(6, 30)
    img = img.contiguous()
    assert lib.hostsim_cnn_encode(C.byref(p), 6, img.data_ptr(), mean.data_ptr(), rstd.data_ptr(), out.data_ptr(), 30) == 0
    assert torch.allclose(out, g["cnn_features"], rtol=1e-4, atol=2e-5), float((out - g["cnn_features"]).abs().max())


@pytest.mark.gpu
def test_gpu_policy_path_reproduces_the_reference_model(built):
    g, model = golden(), load_model().cuda()
    img, obs = inputs()
    img, obs = img.cuda(), obs.cuda()
    with torch.no_grad():
        assert model.actor_cnn.native_ok(img)
        feat = model.encode_image(img)  # agx_cnn_encode, normalisation fused
        res = model({"is_train": True, "prev_actions": torch.zeros(6, 4, device="cuda"), "obs": {"image": img, "observation": obs}})
    torch.cuda.synchronize()
    assert torch.allclose(feat.cpu(), g["cnn_features"], rtol=1e-4, atol=2e-5), float((feat.cpu() - g["cnn_features"]).abs().max())
    # the trunk behind the encoder is torch fp32 here (module forward); the fused TF32 trunk is compared in test_gpu_ppo.py
    assert torch.allclose(res["mus"].cpu(), g["mus"], rtol=2e-3, atol=2e-3), float((res["mus"].cpu() - g["mus"]).abs().max())
    assert torch.allclose(res["values"].cpu(), g["values"], rtol=2e-3, atol=2e-3)
