import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from airgym_b200.lib.config import default_ppo_config
from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd
sys.path.insert(0, "/root/repo/scripts")
from mlp_bench import timeit
model = ModelA2CContinuousLogStd(default_ppo_config("hovering")["params"], {"actions_num": 4, "input_shape": (18,)}).cuda()
model.flatten_parameters(); model.eval()
out = {}
from airgym_b200 import _capi
lib=_capi.load()
for dbg in (0,1):
  lib.agx_mlp_debug(dbg)
  for B in (128, 32768):
    obs = torch.randn(B, 18, device="cuda"); mu, val = torch.zeros(B, 4, device="cuda"), torch.zeros(B, device="cuda")
    with torch.no_grad():
        out[f"dbg{dbg}_fwd_B{B}"] = round(timeit(lambda: model.fused_heads(obs, mu, val)), 2)
print(json.dumps(out))
