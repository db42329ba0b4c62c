"""Strip the reference's shipped policy checkpoint (trained/planning_cnn_rate.pth: model + optimizer + bookkeeping) down to
its `model` state dict + scalars and store it as tests/golden/planning_cnn_rate_model.pth, so that the player / checkpoint
compatibility tests can run on a box without /root/reference.  Weights are data, not code."""
import os
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ck = torch.load("/root/reference/trained/planning_cnn_rate.pth", map_location="cpu", weights_only=False)
out = {"model": {k: v.clone() for k, v in ck["model"].items()}, "epoch": int(ck["epoch"]), "frame": int(ck["frame"]),
       "last_mean_rewards": float(ck["last_mean_rewards"]), "env_state": None}
# the per-pixel image statistics (2 x 25 440 float64) dominate the size; float32 keeps 7 digits of them
for k in list(out["model"]):
    if "image.running" in k:
        out["model"][k] = out["model"][k].float()
torch.save(out, os.path.join(HERE, "planning_cnn_rate_model.pth"))
print({k: tuple(v.shape) for k, v in out["model"].items() if "cnn" not in k and "mlp" not in k})
