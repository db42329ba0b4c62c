"""Cross-rank merging of batch statistics for the sharded PPO update: every rank contributes the float64 sums
(Σx, Σx²) of its shard, one SUM all-reduce later every rank derives the moments of the GLOBAL batch — the quantities
RunningMeanStd.update (lib/core/running_mean_std.py:45-60: batch mean, UNBIASED batch variance, batch size) and the advantage
normalisation (lib/agent/a2c_continuous.py:160-164: mean, unbiased std) are defined on.  Device-agnostic torch code."""
import torch


def batch_sums(x):
    """x [n, k] (or [n]) → float64 [2, k]: column sums and sums of squares of this rank's shard."""
    x64 = x.double().reshape(x.shape[0], -1)
    return torch.stack((x64.sum(0), (x64 * x64).sum(0)))


def moments_from_sums(s, n):
    """all-reduced sums [2, k] over n samples in total → (mean [k], unbiased variance [k])."""
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return mean, var


def sums_from_moments(mean, var, n):
    """inverse of moments_from_sums for a shard whose (mean, unbiased var, n) are already known (cached image moments)."""
    return torch.stack((mean * n, var * (n - 1) + mean * mean * n))
