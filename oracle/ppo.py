"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the reference's PPO iteration, each function citing what it follows.
Pinned by `tests/golden/make_golden_ppo.py`, which imports the reference's own `lib/` modules (they are pure torch and DO
import here) and records seeded inputs/outputs into tests/golden/ppo_*.npz.
"""
import math

import torch
import torch.nn.functional as F


# ---- model: lib/model/a2c_continuous_logstd_model.py:159-198, lib/network/mlp.py:37-39 --------------------------------
def mlp_forward(sd, x, prefix="actor_mlp.layers."):
    i = 0
    while f"{prefix}{i}.weight" in sd:
        x = F.elu(F.linear(x, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"]))
        i += 1
    return x


def rms_normalize(x, mean, var, eps=1e-5):
    """RunningMeanStd.forward, eval branch (lib/core/running_mean_std.py:76-80)."""
    y = (x - mean.float()) / torch.sqrt(var.float() + eps)
    return torch.clamp(y, min=-5.0, max=5.0)


def rms_denorm(x, mean, var, eps=1e-5):
    """denorm branch (:72-74): clamp first, then scale."""
    y = torch.clamp(x, min=-5.0, max=5.0)
    return torch.sqrt(var.float() + eps) * y + mean.float()


def rms_update(mean, var, count, batch):
    """_update_mean_var_count_from_moments with the batch mean / UNBIASED var / batch size (:33-53); f64 buffers."""
    bm, bv, bc = batch.mean(0), batch.var(0), batch.shape[0]
    delta = bm - mean
    tot = count + bc
    new_mean = mean + delta * bc / tot
    m2 = var * count + bv * bc + delta**2 * count * bc / tot
    return new_mean, m2 / tot, tot


def neglogp(x, mean, std, logstd):
    """a2c_continuous_logstd_model.py:195-198"""
    return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * math.log(2.0 * math.pi) * x.size()[-1] + logstd.sum(dim=-1)


def model_forward(sd, obs, normalize_input=True):
    """Non-separate, no image.  Returns mu, logstd (broadcast), sigma, value (normalised head output)."""
    x = rms_normalize(obs, sd["running_mean_std.running_mean"], sd["running_mean_std.running_var"]) if normalize_input else obs
    h = mlp_forward(sd, x)
    mu = F.linear(h, sd["mu.weight"], sd["mu.bias"])
    logstd = mu * 0.0 + sd["logstd"]
    value = F.linear(h, sd["value_head.weight"], sd["value_head.bias"])
    return mu, logstd, torch.exp(logstd), value


# ---- losses: lib/core/common_losses.py:10-48, a2c_continuous.py:378-390, lib/core/torch_ext.py:27-36 ---------------------
def actor_loss(old_neglogp, new_neglogp, advantage, e_clip):
    ratio = torch.exp(old_neglogp - new_neglogp)
    surr1 = advantage * ratio
    surr2 = advantage * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)
    return torch.max(-surr1, -surr2)


def critic_loss(values, returns):
    return (returns - values) ** 2  # clip_value: False (ppo_hovering.yaml:53)


def bound_loss(mu, soft_bound=1.1):
    hi = torch.clamp_min(mu - soft_bound, 0.0) ** 2
    lo = torch.clamp_max(mu + soft_bound, 0.0) ** 2
    return (lo + hi).sum(axis=-1)


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma):
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma**2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma**2 + 1e-5))
    return (c1 + c2 - 0.5).sum(dim=-1).mean()


def total_loss(mu, logstd, sigma, value, batch, hp):
    """calc_gradients (a2c_continuous.py:299-349): returns (loss, dict of the mean terms)."""
    nlp = neglogp(batch["actions"], mu, sigma, logstd)
    a = actor_loss(batch["old_logp_actions"], nlp, batch["advantages"], hp["e_clip"]).mean()
    c = critic_loss(value, batch["returns"]).mean()
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum(dim=-1).mean()
    b = bound_loss(mu).mean()
    loss = a + 0.5 * c * hp["critic_coef"] - ent * hp["entropy_coef"] + b * hp["bounds_loss_coef"]
    kl = policy_kl(mu.detach(), sigma.detach(), batch["mu"], batch["sigma"])
    return loss, {"a_loss": a, "c_loss": c, "entropy": ent, "b_loss": b, "kl": kl}


def adaptive_lr(lr, kl, kl_threshold=0.008, min_lr=1e-6, max_lr=1e-2):
    """AdaptiveScheduler.update (lib/core/schedulers.py:26-32)"""
    new = lr
    if kl > 2.0 * kl_threshold:
        new = max(lr / 1.5, min_lr)
    if kl < 0.5 * kl_threshold:
        new = min(lr * 1.5, max_lr)
    return new


# ---- GAE: lib/agent/a2c_base.py:463-478 (time-major [H,N,1]) ------------------------------------------------------------
def discount_values(fdones, last_values, mb_fdones, mb_values, mb_rewards, gamma, tau):
    H = mb_rewards.shape[0]
    lastgaelam = 0
    mb_advs = torch.zeros_like(mb_rewards)
    for t in reversed(range(H)):
        if t == H - 1:
            nextnonterminal = 1.0 - fdones
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - mb_fdones[t + 1]
            nextvalues = mb_values[t + 1]
        nextnonterminal = nextnonterminal.unsqueeze(1)
        delta = mb_rewards[t] + gamma * nextvalues * nextnonterminal - mb_values[t]
        mb_advs[t] = lastgaelam = delta + gamma * tau * nextnonterminal * lastgaelam
    return mb_advs


def swap_and_flatten01(arr):
    """a2c_base.py:26-33: [H,N,...] → [N*H,...] env-major."""
    s = arr.size()
    return arr.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


# ---- optimizer step: a2c_base.py:293-316 (clip_grad_norm_ + torch.optim.Adam eps 1e-8) ---------------------------------
def clip_and_adam(params, grads, exp_avg, exp_avg_sq, step, lr, grad_norm=1.5, beta1=0.9, beta2=0.999, eps=1e-8):
    """Flat-tensor restatement; returns (new_params, new_m, new_v, total_norm)."""
    total_norm = torch.sqrt((grads.double() ** 2).sum()).float()
    clip = torch.clamp(grad_norm / (total_norm + 1e-6), max=1.0)
    g = grads * clip
    m = beta1 * exp_avg + (1 - beta1) * g
    v = beta2 * exp_avg_sq + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1**step, 1 - beta2**step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return params - (lr / bc1) * m / denom, m, v, total_norm
