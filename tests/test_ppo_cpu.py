"""PPO row (a13) on CPU: oracle/ppo.py against the golden recorded from the reference's own lib/ code, and the kernel
arithmetic (agx_ppo_math.cuh compiled by g++) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from airgym_b200 import _capi
from oracle import ppo as O
from tests.util import GOLDEN_DIR, assert_close

HP = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ppo_hovering.npz"), allow_pickle=False))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_gae_and_flatten_match_reference(gold):
    g = gold
    t = lambda k: torch.from_numpy(g[k])
    advs = O.discount_values(t("fdones"), t("last_values"), t("dones").float(), t("values"), t("rewards"), 0.99, 0.95)
    assert_close(advs, g["mb_advs"], "GAE", rtol=0, atol=0)
    assert_close(advs + t("values"), g["mb_returns"], "returns", rtol=0, atol=0)


def test_value_rms_and_advantage_normalisation(gold):
    g = gold
    t = lambda k: torch.from_numpy(g[k])
    sd0 = _sd(g, "sd0/")
    b_val, b_ret = O.swap_and_flatten01(t("values")), O.swap_and_flatten01(t("mb_returns"))
    m, v, c = sd0["value_mean_std.running_mean"], sd0["value_mean_std.running_var"], sd0["value_mean_std.count"]
    m, v, c = O.rms_update(m, v, c, b_val)
    n_val = O.rms_normalize(b_val, m, v)
    m, v, c = O.rms_update(m, v, c, b_ret)
    n_ret = O.rms_normalize(b_ret, m, v)
    assert_close(n_val, g["n_val"], "normalised values", rtol=1e-6, atol=1e-6)
    assert_close(n_ret, g["n_ret"], "normalised returns", rtol=1e-6, atol=1e-6)
    assert_close(m, g["vms_mean"], "value RMS mean", rtol=1e-12, atol=0)
    assert_close(c, g["vms_count"], "value RMS count", rtol=0, atol=0)
    adv = torch.sum(b_ret - b_val, axis=1)
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    assert_close(adv, g["advantages"], "advantages", rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("i", [0, 1])
def test_minibatch_update_matches_reference(gold, i):
    g = gold
    N, H, A = int(g["N"]), int(g["H"]), int(g["A"])
    B = N * H
    mb = B // 2
    sl = slice(i * mb, (i + 1) * mb)
    t = lambda k: torch.from_numpy(g[k])
    fl = O.swap_and_flatten01
    sd = _sd(g, f"step{i}/sd_before/")
    sd_after = _sd(g, f"step{i}/sd_after/")
    obs = fl(t("obs"))[sl]
    # train-mode forward: the obs RMS is updated with the minibatch first (running_mean_std.py:45-53)
    m, v, c = O.rms_update(sd["running_mean_std.running_mean"], sd["running_mean_std.running_var"], sd["running_mean_std.count"], obs)
    assert_close(m, sd_after["running_mean_std.running_mean"], "obs RMS mean", rtol=1e-12, atol=0)
    sd_fwd = {k: (x.clone().requires_grad_(True) if x.dtype == torch.float32 else x) for k, x in sd.items()}
    sd_fwd["running_mean_std.running_mean"], sd_fwd["running_mean_std.running_var"] = m, v
    mu, logstd, sigma, value = O.model_forward(sd_fwd, obs)
    assert_close(mu, g[f"step{i}/mu"], "mu", rtol=1e-5, atol=1e-6)
    assert_close(value, g[f"step{i}/value"], "value", rtol=1e-5, atol=1e-6)
    old_mu, old_sigma = fl(t("mus"))[sl], fl(t("sigmas"))[sl]
    batch = {"actions": fl(t("actions"))[sl], "old_logp_actions": fl(t("neglogpacs"))[sl], "advantages": t("advantages")[sl],
             "returns": t("n_ret")[sl], "mu": old_mu, "sigma": old_sigma}
    loss, terms = O.total_loss(mu, logstd, sigma, value, batch, HP)
    for k in ("a_loss", "c_loss", "entropy", "b_loss", "kl"):
        assert_close(terms[k].detach(), g[f"step{i}/{k}"], k, rtol=2e-5, atol=1e-7)
    assert_close(loss.detach(), g[f"step{i}/loss"], "loss", rtol=2e-5, atol=1e-7)
    loss.backward()
    names = [str(n) for n in g["param_names"]]
    for n in names:
        assert_close(sd_fwd[n].grad, g[f"step{i}/grads/{n}"], f"grad {n}", rtol=1e-4, atol=1e-7)
    # clip + Adam on the flat vector, LR rule
    flat = lambda d: torch.cat([d[n].detach().reshape(-1) for n in names])
    grads = torch.cat([sd_fwd[n].grad.reshape(-1) for n in names])
    if i == 0:
        m0 = v0 = torch.zeros_like(grads)
        newp, _, _, norm = O.clip_and_adam(flat(sd), grads, m0, v0, 1, float(g[f"step{i}/lr_in"]))
        assert_close(norm, g[f"step{i}/total_norm"], "grad norm", rtol=1e-5, atol=0)
        assert_close(newp, flat(sd_after), "params after Adam", rtol=1e-5, atol=1e-7)
    assert O.adaptive_lr(float(g[f"step{i}/lr_in"]), float(g[f"step{i}/kl"])) == pytest.approx(float(g[f"step{i}/lr_out"]), rel=1e-12)


def _hyper():
    hp = _capi.AgxPpoHyper()
    hp.e_clip, hp.critic_coef, hp.entropy_coef, hp.bounds_loss_coef = 0.2, 2.0, 0.0, 1e-4
    hp.kl_threshold, hp.grad_norm, hp.beta1, hp.beta2, hp.eps, hp.weight_decay, hp.adaptive_lr = 0.008, 1.5, 0.9, 0.999, 1e-8, 0.0, 1
    return hp


def test_kernel_math_gae_vs_oracle(built):
    from tests.hostsim.driver import build

    lib = build()
    torch.manual_seed(0)
    N, H = 70, 24
    rewards, values = torch.rand(N, H), torch.randn(N, H)
    dones = (torch.rand(N, H) < 0.15).to(torch.uint8)
    last_v, last_d = torch.randn(N), (torch.rand(N) < 0.3).to(torch.uint8)
    adv, ret = np.zeros((N, H), np.float32), np.zeros((N, H), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    r_np, v_np, d_np, lv, ld = rewards.numpy(), values.numpy(), dones.numpy(), last_v.numpy(), last_d.numpy()
    lib.hostsim_gae.argtypes = [C.c_int64, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 7
    lib.hostsim_gae(N, H, 0.99, 0.95, p(r_np), p(v_np), p(d_np), p(lv), p(ld), p(adv), p(ret))
    tm = lambda x: x.t().unsqueeze(-1)  # [N,H] → [H,N,1]
    ref = O.discount_values(last_d.float(), last_v.unsqueeze(1), dones.t().float(), tm(values), tm(rewards), 0.99, 0.95)
    assert_close(adv, ref.squeeze(-1).t(), "gae", rtol=1e-5, atol=1e-6)
    assert_close(ret, (ref + tm(values)).squeeze(-1).t(), "returns", rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("A", [4, 5])
def test_kernel_math_loss_and_grads_vs_autograd(built, A):
    from tests.hostsim.driver import build

    lib = build()
    torch.manual_seed(1)
    B = 300
    mu = (torch.randn(B, A) * 0.8).requires_grad_(True)
    mu.data[:20] *= 2.0  # exercise the bound loss
    logstd = (torch.randn(A) * 0.3).requires_grad_(True)
    value = torch.randn(B, 1, requires_grad=True)
    actions = mu.detach() + torch.randn(B, A) * 0.7
    old_mu, old_sigma = mu.detach() + 0.05 * torch.randn(B, A), torch.exp(logstd.detach() + 0.05 * torch.randn(B, A))
    old_nlp = O.neglogp(actions, old_mu, old_sigma, torch.log(old_sigma)) + 0.3 * torch.randn(B)  # push ratios out of the clip range
    adv, ret = torch.randn(B), torch.randn(B, 1)
    hp = dict(HP, entropy_coef=0.01)
    ls_b = mu * 0.0 + logstd
    loss, terms = O.total_loss(mu, ls_b, torch.exp(ls_b), value, {"actions": actions, "old_logp_actions": old_nlp, "advantages": adv,
                                                                  "returns": ret, "mu": old_mu, "sigma": old_sigma}, hp)
    loss.backward()
    H = _hyper()
    H.entropy_coef = 0.01
    g_mu, g_val, g_ls, stats = np.zeros((B, A), np.float32), np.zeros(B, np.float32), np.zeros(A, np.float32), np.zeros(8, np.float32)
    om, os_ = old_mu.numpy().copy(), old_sigma.numpy().copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    arrs = [mu.detach().numpy(), logstd.detach().numpy(), value.detach().numpy().reshape(-1).copy(), actions.numpy(), old_nlp.numpy(),
            adv.numpy(), ret.numpy().reshape(-1).copy()]
    lib.hostsim_ppo_loss.argtypes = [C.POINTER(_capi.AgxPpoHyper), C.c_int64, C.c_int] + [C.c_void_p] * 13
    lib.hostsim_ppo_loss(C.byref(H), B, A, *[p(a) for a in arrs], p(om), p(os_), p(g_mu), p(g_val), p(g_ls), p(stats))
    for j, k in enumerate(("a_loss", "c_loss", "entropy", "b_loss", "kl")):
        assert_close(stats[j], terms[k].detach(), k, rtol=1e-5, atol=1e-7)
    assert_close(g_mu, mu.grad, "grad mu", rtol=1e-4, atol=1e-8)
    assert_close(g_val, value.grad.reshape(-1), "grad value", rtol=1e-5, atol=1e-9)
    assert_close(g_ls, logstd.grad, "grad logstd", rtol=1e-4, atol=1e-7)
    assert_close(om, mu.detach(), "mu write-back", rtol=0, atol=0)
    assert_close(os_, torch.exp(logstd.detach()).expand(B, A), "sigma write-back", rtol=1e-6, atol=0)
