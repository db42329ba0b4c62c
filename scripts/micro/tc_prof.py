import sys, os, ctypes
sys.path.insert(0, '/root/repo')
import torch
from airgym_b200.lib.config import default_ppo_config
from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd
model = ModelA2CContinuousLogStd(default_ppo_config("hovering")["params"], {"actions_num": 4, "input_shape": (18,)}).cuda()
model.flatten_parameters(); model.eval()
B = 65536
obs = torch.randn(B, 18, device="cuda"); mu, val = torch.zeros(B, 4, device="cuda"), torch.zeros(B, device="cuda")
for _ in range(5):
    model.fused_heads(obs, mu, val)
torch.cuda.synchronize()
