"""RunningMeanStd with the reference's buffers and update rule (lib/core/running_mean_std.py:8-81): float64
`running_mean`, `running_var`, `count` (initial 0 / 1 / 1), parallel-variance merge with the batch mean, UNBIASED batch
variance and batch size; forward normalises with clamp(+-5) or de-normalises."""
import torch
import torch.nn as nn


class RunningMeanStd(nn.Module):
    def __init__(self, insize, epsilon=1e-05):
        super().__init__()
        self.insize, self.epsilon = insize, epsilon
        self.register_buffer("running_mean", torch.zeros(insize, dtype=torch.float64))
        self.register_buffer("running_var", torch.ones(insize, dtype=torch.float64))
        self.register_buffer("count", torch.ones((), dtype=torch.float64))

    @torch.no_grad()
    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.running_mean
        tot = self.count + batch_count
        new_mean = self.running_mean + delta * batch_count / tot
        m2 = self.running_var * self.count + batch_var * batch_count + delta**2 * self.count * batch_count / tot
        self.running_mean.copy_(new_mean)
        self.running_var.copy_(m2 / tot)
        self.count.copy_(tot)

    def forward(self, x, denorm=False):
        if self.training:
            var, mean = torch.var_mean(x, dim=0)  # unbiased
            self.update_from_moments(mean, var, x.size(0))
        mean, var = self.running_mean.float(), self.running_var.float()
        if denorm:
            return torch.sqrt(var + self.epsilon) * torch.clamp(x, min=-5.0, max=5.0) + mean
        return torch.clamp((x - mean) / torch.sqrt(var + self.epsilon), min=-5.0, max=5.0)


class RunningMeanStdObs(nn.Module):
    """Per-key RunningMeanStd for dict observations (lib/core/running_mean_std.py:84-95): state_dict keys
    `running_mean_std.<key>.{running_mean,running_var,count}`."""

    def __init__(self, insize, epsilon=1e-05):
        assert isinstance(insize, dict)
        super().__init__()
        self.running_mean_std = nn.ModuleDict({k: RunningMeanStd(tuple(v), epsilon) for k, v in insize.items()})

    def forward(self, input, denorm=False):
        return {k: self.running_mean_std[k](v, denorm) for k, v in input.items()}
