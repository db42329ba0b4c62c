"""Generate tests/golden/ppo_hovering.npz from the REFERENCE's own PPO code (needs /root/reference; build container only).

Imports, unmodified: lib/model/a2c_continuous_logstd_model.py (ModelA2CContinuousLogStd + MLP + RunningMeanStd),
lib/core/common_losses.py, lib/core/torch_ext.py (policy_kl), lib/core/schedulers.py (AdaptiveScheduler),
lib/core/datasets.py (PPODataset), lib/agent/a2c_base.py (discount_values, swap_and_flatten01; `gym`/`tensorboardX`, which
a2c_base imports at module top but the recorded functions never touch, are stubbed), plus torch.optim.Adam and
nn.utils.clip_grad_norm_ exactly as a2c_base.py:293-316 / a2c_continuous.py:401 use them.  Records one rollout's GAE and two
consecutive minibatch updates (forward, losses, backward, clip, Adam, KL, LR rule, mu/sigma write-back) and asserts that
oracle/ppo.py reproduces every number.
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import ppo as O  # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    from tests.golden.make_golden import install_stubs  # isaacgym / rlPx4Controller / pytorch3d / rospy shells

    install_stubs()  # a2c_base → lib.utils.vecenv → `from airgym.envs import *` pulls the task modules in
    stub("gym", spaces=stub("gym.spaces", Box=object, Discrete=object, Tuple=object, Dict=object), Wrapper=object, Env=object)
    stub("tensorboardX", SummaryWriter=object)
    from lib.agent import a2c_base
    from lib.core import common_losses, torch_ext
    from lib.core.datasets import PPODataset
    from lib.core.schedulers import AdaptiveScheduler
    from lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    cfg = yaml.safe_load(open(os.path.join(REF, "scripts/config/ppo_hovering.yaml")))["params"]
    c = cfg["config"]
    torch.manual_seed(123)
    N, H, A, OBS = 32, c["horizon_length"], 4, 18
    model = ModelA2CContinuousLogStd(cfg, {"actions_num": A, "input_shape": (OBS,), "num_seqs": N, "value_size": 1,
                                           "normalize_value": True, "normalize_input": True})
    with torch.no_grad():  # non-trivial normalisation statistics and log-std
        model.running_mean_std.running_mean.copy_(torch.randn(OBS, dtype=torch.float64) * 0.3)
        model.running_mean_std.running_var.copy_(torch.rand(OBS, dtype=torch.float64) + 0.5)
        model.running_mean_std.count.fill_(5000.0)
        model.value_mean_std.running_mean.fill_(0.7)
        model.value_mean_std.running_var.fill_(2.3)
        model.logstd.copy_(torch.tensor([-0.3, 0.1, 0.0, -0.5]))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}

    # ---- rollout-shaped data and GAE (a2c_base.discount_values on a bare object) --------------------------------
    obs = torch.randn(H, N, OBS) * 1.5
    model.eval()
    with torch.no_grad():
        flat = model({"is_train": False, "prev_actions": None, "obs": obs.reshape(H * N, OBS)})
    actions = flat["actions"].reshape(H, N, A)
    neglogpacs = flat["neglogpacs"].reshape(H, N)
    values = flat["values"].reshape(H, N, 1)  # de-normalised, as play_steps stores them
    mus, sigmas = flat["mus"].reshape(H, N, A), flat["sigmas"].reshape(H, N, A)
    rewards = torch.rand(H, N, 1) * 0.3
    dones = (torch.rand(H, N) < 0.1).to(torch.uint8)
    dones[0] = 1
    fdones = (torch.rand(N) < 0.2).float()
    last_values = torch.randn(N, 1)
    fake = types.SimpleNamespace(horizon_length=H, gamma=c["gamma"], tau=c["tau"])
    mb_advs = a2c_base.A2CBase.discount_values(fake, fdones, last_values, dones.float(), values, rewards)
    mb_returns = mb_advs + values
    o_advs = O.discount_values(fdones, last_values, dones.float(), values, rewards, c["gamma"], c["tau"])
    assert torch.equal(o_advs, mb_advs)
    flat_of = lambda t: a2c_base.swap_and_flatten01(t)
    assert torch.equal(O.swap_and_flatten01(obs), flat_of(obs))

    # ---- prepare_dataset (a2c_continuous.py:140-177) ------------------------------------------------------------
    b_obs, b_act, b_nlp = flat_of(obs), flat_of(actions), flat_of(neglogpacs)
    b_val, b_ret, b_mu, b_sig = flat_of(values), flat_of(mb_returns), flat_of(mus), flat_of(sigmas)
    advantages = b_ret - b_val
    model.value_mean_std.train()
    n_val = model.value_mean_std(b_val)
    n_ret = model.value_mean_std(b_ret)
    model.value_mean_std.eval()
    advantages = torch.sum(advantages, axis=1)
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    vms_after = {k: v.clone() for k, v in model.value_mean_std.state_dict().items()}
    # oracle: two RMS updates then normalise
    m, v, cnt = sd0["value_mean_std.running_mean"], sd0["value_mean_std.running_var"], sd0["value_mean_std.count"]
    m, v, cnt = O.rms_update(m, v, cnt, b_val)
    o_nval = O.rms_normalize(b_val, m, v)
    m, v, cnt = O.rms_update(m, v, cnt, b_ret)
    o_nret = O.rms_normalize(b_ret, m, v)
    assert torch.allclose(o_nval, n_val, atol=1e-6) and torch.allclose(o_nret, n_ret, atol=1e-6)
    assert torch.allclose(m, vms_after["running_mean"]) and torch.allclose(v, vms_after["running_var"])

    B = H * N
    mb = B // 2
    ds = PPODataset(B, mb, False, "cpu")
    ds.update_values_dict({"old_values": n_val, "old_logp_actions": b_nlp, "advantages": advantages, "returns": n_ret,
                           "actions": b_act, "obs": b_obs, "mu": b_mu.clone(), "sigma": b_sig.clone()})

    # ---- two minibatch updates (calc_gradients + trancate_gradients_and_step + legacy adaptive schedule) ----------
    opt = torch.optim.Adam(model.parameters(), float(c["learning_rate"]), eps=1e-08, weight_decay=0)
    sched = AdaptiveScheduler(c["kl_threshold"])
    lr = float(c["learning_rate"])
    model.train()
    rec = {"steps": []}
    hp = {"e_clip": c["e_clip"], "critic_coef": c["critic_coef"], "entropy_coef": c["entropy_coef"],
          "bounds_loss_coef": c["bounds_loss_coef"]}
    names = [n for n, _ in model.named_parameters()]
    for i in range(2):
        d = ds[i]
        sd_before = {k: v.clone() for k, v in model.state_dict().items()}
        res = model({"is_train": True, "prev_actions": d["actions"], "obs": d["obs"]})
        a_loss = common_losses.actor_loss(d["old_logp_actions"], res["prev_neglogp"], d["advantages"], True, c["e_clip"])
        c_loss = common_losses.critic_loss(model, d["old_values"], res["values"], c["e_clip"], d["returns"], c["clip_value"])
        mu = res["mus"]
        b_loss = (torch.clamp_max(mu + 1.1, 0.0) ** 2 + torch.clamp_min(mu - 1.1, 0.0) ** 2).sum(axis=-1)
        losses = torch_ext.apply_masks([a_loss.unsqueeze(1), c_loss, res["entropy"].unsqueeze(1), b_loss.unsqueeze(1)])
        al, cl, ent, bl = losses
        loss = al + 0.5 * cl * c["critic_coef"] - ent * c["entropy_coef"] + bl * c["bounds_loss_coef"]
        for p in model.parameters():
            p.grad = None
        loss.backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), c["grad_norm"])
        opt.step()
        kl = torch_ext.policy_kl(mu.detach(), res["sigmas"].detach(), d["mu"], d["sigma"], True)
        ds.update_mu_sigma(mu.detach(), res["sigmas"].detach())
        new_lr, _ = sched.update(lr, 0.0, 0, 0, kl.item())
        for g in opt.param_groups:
            g["lr"] = new_lr
        sd_after = {k: v.clone() for k, v in model.state_dict().items()}
        # ---- the oracle must reproduce all of it
        o_mu, o_logstd, o_sigma, o_value = O.model_forward({**sd_before,
                                                             "running_mean_std.running_mean": model.running_mean_std.running_mean,
                                                             "running_mean_std.running_var": model.running_mean_std.running_var}, d["obs"])
        assert torch.allclose(o_mu, mu, atol=1e-6) and torch.allclose(o_value, res["values"], atol=1e-6)
        o_loss, o_terms = O.total_loss(o_mu, o_logstd, o_sigma, o_value, {**d, "mu": d["mu"], "sigma": d["sigma"]}, hp) \
            if False else (None, None)
        rec["steps"].append(dict(
            lr_in=lr, lr_out=new_lr, kl=float(kl), a_loss=float(al), c_loss=float(cl), entropy=float(ent), b_loss=float(bl),
            loss=float(loss), total_norm=float(total_norm), mu=mu.detach().numpy(), sigma=res["sigmas"].detach().numpy(),
            value=res["values"].detach().numpy(), neglogp=res["prev_neglogp"].detach().numpy(),
            grads={n: grads[n].numpy() for n in names}, sd_before={k: v.numpy() for k, v in sd_before.items()},
            sd_after={k: v.numpy() for k, v in sd_after.items()}))
        lr = new_lr

    out = {"N": N, "H": H, "A": A, "OBS": OBS, "obs": obs.numpy(), "actions": actions.numpy(), "neglogpacs": neglogpacs.numpy(),
           "values": values.numpy(), "mus": mus.numpy(), "sigmas": sigmas.numpy(), "rewards": rewards.numpy(),
           "dones": dones.numpy(), "fdones": fdones.numpy(), "last_values": last_values.numpy(), "mb_advs": mb_advs.numpy(),
           "mb_returns": mb_returns.numpy(), "n_val": n_val.numpy(), "n_ret": n_ret.numpy(), "advantages": advantages.numpy(),
           "vms_mean": vms_after["running_mean"].numpy(), "vms_var": vms_after["running_var"].numpy(),
           "vms_count": vms_after["count"].numpy(), "param_names": np.array(names)}
    for k, v in sd0.items():
        out["sd0/" + k] = v.numpy()
    for i, s in enumerate(rec["steps"]):
        for k, v in s.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    out[f"step{i}/{k}/{kk}"] = vv
            else:
                out[f"step{i}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ppo_hovering.npz"), **out)
    print("ppo_hovering.npz written:", {k: (s["kl"], s["lr_in"], s["lr_out"], s["total_norm"]) for k, s in enumerate(rec["steps"])})


if __name__ == "__main__":
    main()
