"""Actor-critic with a state-independent log-std — same forward contract and state_dict key layout as the reference's
ModelA2CContinuousLogStd (lib/model/a2c_continuous_logstd_model.py:14-198, lib/network/mlp.py:4-39) for the non-separate
MLP case of the shipped yamls: `logstd`, `actor_mlp.layers.{i}.{weight,bias}`, `mu.*`, `value_head.*`,
`value_mean_std.*`, `running_mean_std.*`.  All trainable parameters are views into ONE flat fp32 buffer (and their grads
into one flat grad buffer), which is what the fused clip+Adam kernel and the NCCL all-reduce operate on."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..core.running_mean_std import RunningMeanStd

_ACTS = {"elu": F.elu, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "sin": torch.sin}


class MLP(nn.Module):
    def __init__(self, input_size, units, activation):
        super().__init__()
        if activation not in _ACTS:
            raise ValueError(f"Unsupported activation: {activation}")
        self.activation = _ACTS[activation]
        self.layers = nn.ModuleList()
        in_dim = int(input_size)
        for out_dim in units:
            layer = nn.Linear(in_dim, out_dim)  # default torch init for the weight; bias zeroed (mlp.py:28-35)
            nn.init.zeros_(layer.bias)
            self.layers.append(layer)
            in_dim = out_dim

    def forward(self, x):
        for layer in self.layers:
            x = self.activation(layer(x))
        return x


class ModelA2CContinuousLogStd(nn.Module):
    def __init__(self, params, keys):
        super().__init__()
        net = params["network"]
        if net.get("separate", False) or "cnn" in net or "vae" in net or "resnet" in net:
            raise NotImplementedError("only the non-separate MLP network of ppo_hovering/tracking.yaml is built in this round")
        self.actions_num = keys["actions_num"]
        input_shape = keys["input_shape"]
        self.normalize_value = params["config"].get("normalize_value", False)
        self.normalize_input = params["config"].get("normalize_input", False)
        self.value_size = params["config"].get("value_size", 1)
        units, act = net["mlp"]["units"], net["mlp"]["activation"]
        assert net["space"]["continuous"].get("fixed_sigma", True), "fixed_sigma: True is the only shipped configuration"
        self.actor_mlp = MLP(input_shape[0], units, act)
        self.mu = nn.Linear(units[-1], self.actions_num)
        self.mu.weight.data.mul_(0.1)
        self.mu.bias.data.mul_(0.0)
        self.logstd = nn.Parameter(torch.zeros(self.actions_num, dtype=torch.float32))
        self.value_head = nn.Linear(units[-1], 1)
        self.value_head.weight.data.mul_(0.1)
        self.value_head.bias.data.mul_(0.0)
        if self.normalize_value:
            self.value_mean_std = RunningMeanStd((self.value_size,))
        if self.normalize_input:
            self.running_mean_std = RunningMeanStd(tuple(input_shape))
        self.flat_params = self.flat_grads = None

    # ---- flat parameter / gradient storage ------------------------------------------------------------------------
    def flatten_parameters(self, extra_grad_slots=0):
        """Re-home every parameter (and a pre-allocated grad) as a view of one flat buffer; `extra_grad_slots` floats
        are appended to the grad buffer so scalars (the KL) can ride along in the same all-reduce."""
        ps = list(self.parameters())
        n = sum(p.numel() for p in ps)
        dev = ps[0].device
        flat = torch.empty(n, device=dev, dtype=torch.float32)
        grads = torch.zeros(n + extra_grad_slots, device=dev, dtype=torch.float32)
        off = 0
        for p in ps:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view_as(p.data)
            p.grad = grads[off:off + k].view_as(p.data)
            off += k
        self.flat_params, self.flat_grads, self.num_flat = flat, grads, n
        return flat, grads

    # ---- forward ----------------------------------------------------------------------------------------------------
    def norm_obs(self, obs):
        return self.running_mean_std(obs) if self.normalize_input else obs

    def denorm_value(self, value):
        return self.value_mean_std(value, denorm=True) if self.normalize_value else value

    def heads(self, obs):
        h = self.actor_mlp(self.norm_obs(obs))
        return self.mu(h), self.value_head(h)

    @staticmethod
    def neglogp(x, mean, std, logstd):
        return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * math.log(2.0 * math.pi) * x.size()[-1] + logstd.sum(dim=-1)

    def forward(self, input_dict):
        is_train = input_dict.get("is_train", True)
        mu, value = self.heads(input_dict["obs"])
        logstd = mu * 0.0 + self.logstd
        sigma = torch.exp(logstd)
        if is_train:
            entropy = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum(dim=-1)
            prev_neglogp = self.neglogp(input_dict["prev_actions"], mu, sigma, logstd)
            return {"prev_neglogp": torch.squeeze(prev_neglogp), "values": value, "entropy": entropy, "mus": mu, "sigmas": sigma}
        action = mu + sigma * torch.randn_like(mu)  # Normal(mu, sigma).sample()
        return {"neglogpacs": torch.squeeze(self.neglogp(action, mu, sigma, logstd)), "values": self.denorm_value(value),
                "actions": action, "mus": mu, "sigmas": sigma}
