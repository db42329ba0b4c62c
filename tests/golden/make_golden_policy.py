"""Golden outputs of the REFERENCE's policy network for the camera tasks, produced by the reference's
own modules — lib/model/a2c_continuous_logstd_model.py (ModelA2CContinuousLogStd.forward :139-150,159-168), lib/network/cnn.py,
lib/network/mlp.py, lib/core/running_mean_std.py: pure torch, importable in the build container — with the reference's shipped
weights trained/planning_cnn_rate.pth, on procedural inputs the tests can rebuild anywhere (tests/util_vae.procedural_images
scaled to the depth range the env produces, and a seeded observation batch).

Stored per sample: the CNN features of the normalised image (the encoder alone), mu, the normalised value, sigma.
Writes tests/golden/policy_planning_cnn.npz.  The weights travel separately as tests/golden/planning_cnn_rate_model.pth
(make_golden_ckpt.py; its per-pixel image statistics are stored in float32, which moves these outputs by < 1e-6)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.util_vae import POLICY_IMAGE_SCALE, policy_inputs  # noqa: E402

REF = "/root/reference"
N = 6

if __name__ == "__main__":
    sys.path.insert(0, REF)
    from lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd  # the reference's own model (build container only)

    network = {"name": "actor_critic", "separate": False,
               "space": {"continuous": {"mu_activation": "None", "sigma_activation": "None", "mu_init": {"name": "default"},
                                        "sigma_init": {"name": "const_initializer", "val": 0}, "fixed_sigma": True}},
               "mlp": {"units": [64, 128, 64], "d2rl": False, "activation": "elu", "initializer": {"name": "default", "scale": 2}},
               "cnn": {"output_dim": 30}}  # scripts/config/ppo_planning.yaml:11-33
    params = {"network": network, "config": {"normalize_value": True, "normalize_input": True, "value_size": 1}}
    keys = {"actions_num": 4, "input_shape": {"image": (1, 212, 120), "observation": (16,)}, "num_seqs": 1}
    model = ModelA2CContinuousLogStd(params, keys)
    ck = torch.load(os.path.join(REF, "trained", "planning_cnn_rate.pth"), map_location="cpu", weights_only=False)
    model.load_state_dict(ck["model"])
    model.eval()
    img, obs = policy_inputs(N)
    with torch.no_grad():
        feat = model.actor_cnn(model.norm_image(img))
        res = model({"is_train": True, "prev_actions": torch.zeros(N, 4), "obs": {"image": img, "observation": obs}})
    np.savez_compressed(os.path.join(HERE, "policy_planning_cnn.npz"), cnn_features=feat.numpy(), mus=res["mus"].numpy(),
                        values=res["values"].numpy(), sigmas=res["sigmas"].numpy(), observation=obs.numpy(),
                        image_scale=np.float32(POLICY_IMAGE_SCALE))
    print("features", feat.shape, float(feat.abs().mean()), "mus", res["mus"][0].tolist(), "values", res["values"][:3, 0].tolist())
