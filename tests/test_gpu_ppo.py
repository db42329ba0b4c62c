"""GPU tests of the PPO row (a13): the three update kernels through the C ABI against the oracle (which is pinned to the
reference's lib/ code by tests/golden/ppo_hovering.npz), and the trainer end to end."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from airgym_b200 import _capi
from oracle import ppo as O
from tests.util import GOLDEN_DIR, assert_close

pytestmark = pytest.mark.gpu


def _hyper(entropy_coef=0.0):
    hp = _capi.AgxPpoHyper()
    hp.e_clip, hp.critic_coef, hp.entropy_coef, hp.bounds_loss_coef = 0.2, 2.0, entropy_coef, 1e-4
    hp.kl_threshold, hp.grad_norm, hp.beta1, hp.beta2, hp.eps, hp.weight_decay, hp.adaptive_lr = 0.008, 1.5, 0.9, 0.999, 1e-8, 0.0, 1
    return hp


@pytest.mark.parametrize("N,H", [(1000, 24), (65536, 24), (300, 64)])
def test_gae_kernel_vs_oracle(built, N, H):
    lib = _capi.load()
    torch.manual_seed(0)
    rewards, values = torch.rand(N, H), torch.randn(N, H)
    dones = (torch.rand(N, H) < 0.1).to(torch.uint8)
    last_v, last_d = torch.randn(N), (torch.rand(N) < 0.3).to(torch.uint8)
    d = lambda t: t.cuda()
    r, v, dn, lv, ld = d(rewards), d(values), d(dones), d(last_v), d(last_d)
    adv, ret = torch.zeros(N, H, device="cuda"), torch.zeros(N, H, device="cuda")
    _capi.check(lib.agx_gae(N, H, 0.99, 0.95, r.data_ptr(), v.data_ptr(), dn.data_ptr(), lv.data_ptr(), ld.data_ptr(),
                            adv.data_ptr(), ret.data_ptr(), None))
    tm = lambda x: x.t().unsqueeze(-1)
    ref = O.discount_values(last_d.float(), last_v.unsqueeze(1), dones.t().float(), tm(values), tm(rewards), 0.99, 0.95)
    assert_close(adv.cpu(), ref.squeeze(-1).t(), "gae", rtol=1e-5, atol=1e-6)
    assert_close(ret.cpu(), (ref + tm(values)).squeeze(-1).t(), "returns", rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("A,B", [(4, 300), (5, 4096), (4, 200000)])
def test_ppo_loss_kernel_vs_autograd(built, A, B):
    lib = _capi.load()
    torch.manual_seed(1)
    mu = (torch.randn(B, A) * 0.8).requires_grad_(True)
    mu.data[:20] *= 2.0
    logstd = (torch.randn(A) * 0.3).requires_grad_(True)
    value = torch.randn(B, 1, requires_grad=True)
    actions = mu.detach() + torch.randn(B, A) * 0.7
    old_mu, old_sigma = mu.detach() + 0.05 * torch.randn(B, A), torch.exp(logstd.detach() + 0.05 * torch.randn(B, A))
    old_nlp = O.neglogp(actions, old_mu, old_sigma, torch.log(old_sigma)) + 0.3 * torch.randn(B)
    adv, ret = torch.randn(B), torch.randn(B, 1)
    hp = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.01, bounds_loss_coef=1e-4)
    ls_b = mu * 0.0 + logstd
    loss, terms = O.total_loss(mu, ls_b, torch.exp(ls_b), value, {"actions": actions, "old_logp_actions": old_nlp, "advantages": adv,
                                                                  "returns": ret, "mu": old_mu, "sigma": old_sigma}, hp)
    loss.backward()
    d = lambda t: t.detach().cuda().contiguous()
    g_mu, g_val, g_ls = torch.zeros(B, A, device="cuda"), torch.zeros(B, device="cuda"), torch.zeros(A, device="cuda")
    stats = torch.zeros(8, device="cuda")
    ws = torch.zeros(int(lib.agx_ppo_workspace_floats()), device="cuda")
    om, os_ = d(old_mu), d(old_sigma)
    ins = [d(mu), d(logstd), d(value.reshape(-1)), d(actions), d(old_nlp), d(adv), d(ret.reshape(-1))]
    H = _hyper(0.01)
    for rep in range(2):  # second call: the ticket in the workspace must have been re-armed; write-back makes KL ~ 0
        _capi.check(lib.agx_ppo_loss(C.byref(H), B, A, *[t.data_ptr() for t in ins], om.data_ptr(), os_.data_ptr(), g_mu.data_ptr(),
                                     g_val.data_ptr(), g_ls.data_ptr(), stats.data_ptr(), ws.data_ptr(), None))
        torch.cuda.synchronize()
        if rep == 0:
            for j, k in enumerate(("a_loss", "c_loss", "entropy", "b_loss", "kl")):
                assert_close(stats[j].cpu(), terms[k].detach(), k, rtol=2e-5, atol=1e-7)
            assert_close(g_mu.cpu(), mu.grad, "grad mu", rtol=1e-4, atol=1e-9)
            assert_close(g_val.cpu(), value.grad.reshape(-1), "grad value", rtol=1e-5, atol=1e-10)
            assert_close(g_ls.cpu(), logstd.grad, "grad logstd", rtol=2e-4, atol=1e-7)
            assert_close(om.cpu(), mu.detach(), "mu write-back", rtol=0, atol=0)
        else:
            assert abs(float(stats[4])) < 1e-4


def test_adam_kernel_matches_torch_adam_and_lr_rule(built):
    lib = _capi.load()
    torch.manual_seed(2)
    n = 18121
    p0 = torch.randn(n)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], 3e-4, eps=1e-8)
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    lr_dev = torch.tensor([3e-4], device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    norm = torch.zeros(1, device="cuda")
    H = _hyper()
    lr = 3e-4
    for it, kl in enumerate((0.001, 0.03, 0.006, 0.0001)):
        g = torch.randn(n) * (3.0 if it % 2 == 0 else 0.001)  # alternate clipped / unclipped
        ref_p.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([ref_p], 1.5)
        for grp in opt.param_groups:
            grp["lr"] = lr
        opt.step()
        kl_dev = torch.tensor([kl * 2.0], device="cuda")  # grad_scale 0.5 halves both grads and KL (2-rank all-reduce SUM)
        _capi.check(lib.agx_adam_step(C.byref(H), n, p.data_ptr(), (g * 2.0).cuda().data_ptr(), m.data_ptr(), v.data_ptr(), lr_dev.data_ptr(),
                                      step.data_ptr(), kl_dev.data_ptr(), 0.5, norm.data_ptr(), None))
        torch.cuda.synchronize()
        assert_close(norm.cpu()[0], total, "grad norm", rtol=1e-5, atol=0)
        assert_close(p.cpu(), ref_p.detach(), f"params after step {it}", rtol=2e-5, atol=2e-7)
        lr = O.adaptive_lr(lr, kl)
        assert float(lr_dev) == pytest.approx(lr, rel=1e-6)
        assert int(step) == it + 1


def test_minibatch_update_matches_reference_golden(built):
    """The recorded reference minibatch (reference model + losses + clip + Adam + scheduler) through the product model
    and the kernels: parameters after the step and the next learning rate must match."""
    from airgym_b200.lib.config import default_ppo_config
    from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    g = dict(np.load(os.path.join(GOLDEN_DIR, "ppo_hovering.npz")))
    lib = _capi.load()
    N, Hn, A, OBS = int(g["N"]), int(g["H"]), int(g["A"]), int(g["OBS"])
    cfg = default_ppo_config("hovering")["params"]
    model = ModelA2CContinuousLogStd(cfg, {"actions_num": A, "input_shape": (OBS,)}).cuda()
    sd = {k[len("step0/sd_before/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("step0/sd_before/")}
    assert set(sd) == set(model.state_dict()), "state_dict key layout must equal the reference's"
    model.load_state_dict(sd)
    flat, grads = model.flatten_parameters(extra_grad_slots=8)
    names = [n for n, _ in model.named_parameters()]
    assert names == [str(x) for x in g["param_names"]]
    fl = O.swap_and_flatten01
    t = lambda k: torch.from_numpy(g[k])
    B = N * Hn
    mb = B // 2
    sl = slice(0, mb)
    obs = fl(t("obs"))[sl].cuda()
    model.running_mean_std.train()
    model.running_mean_std(obs)  # train-mode statistics update, then eval forward
    model.eval()
    mu, value = model.heads(obs)
    assert_close(mu.detach().cpu(), g["step0/mu"], "mu", rtol=1e-4, atol=1e-5)
    assert_close(value.detach().cpu(), g["step0/value"], "value", rtol=1e-4, atol=1e-5)
    d = lambda x: x.cuda().contiguous()
    om, os_ = d(fl(t("mus"))[sl]), d(fl(t("sigmas"))[sl])
    g_mu, g_val, g_ls = torch.zeros(mb, A, device="cuda"), torch.zeros(mb, device="cuda"), torch.zeros(A, device="cuda")
    stats = grads[model.num_flat:]
    ws = torch.zeros(int(lib.agx_ppo_workspace_floats()), device="cuda")
    Hh = _hyper()
    ins = [mu, model.logstd, value, d(fl(t("actions"))[sl]), d(fl(t("neglogpacs"))[sl]), d(t("advantages")[sl]), d(t("n_ret")[sl].reshape(-1))]
    _capi.check(lib.agx_ppo_loss(C.byref(Hh), mb, A, *[x.data_ptr() for x in ins], om.data_ptr(), os_.data_ptr(), g_mu.data_ptr(),
                                 g_val.data_ptr(), g_ls.data_ptr(), stats.data_ptr(), ws.data_ptr(), None))
    for j, k in enumerate(("a_loss", "c_loss", "entropy", "b_loss", "kl")):
        assert_close(stats[j].cpu(), g[f"step0/{k}"], k, rtol=1e-4, atol=1e-7)
    grads[: model.num_flat].zero_()
    torch.autograd.backward((mu, value), (g_mu, g_val.view(-1, 1)))
    model.logstd.grad += g_ls
    for n, p in model.named_parameters():
        assert_close(p.grad.cpu(), g[f"step0/grads/{n}"], f"grad {n}", rtol=1e-3, atol=1e-7)
    m, v = torch.zeros(model.num_flat, device="cuda"), torch.zeros(model.num_flat, device="cuda")
    lr_dev, step, norm = torch.tensor([float(g["step0/lr_in"])], device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, device="cuda")
    _capi.check(lib.agx_adam_step(C.byref(Hh), model.num_flat, flat.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(), lr_dev.data_ptr(),
                                  step.data_ptr(), stats[4:5].data_ptr(), 1.0, norm.data_ptr(), None))
    torch.cuda.synchronize()
    assert_close(norm.cpu()[0], g["step0/total_norm"], "grad norm", rtol=1e-4, atol=0)
    for n, p in model.named_parameters():
        assert_close(p.detach().cpu(), g[f"step0/sd_after/{n}"], f"param {n} after the step", rtol=1e-4, atol=1e-6)
    assert float(lr_dev) == pytest.approx(float(g["step0/lr_out"]), rel=1e-6)


@pytest.mark.parametrize("graph,fused", [(False, False), (True, False), (False, True), (True, True)])
def test_trainer_end_to_end_learns_and_checkpoints(built, tmp_path, graph, fused):
    """A few PPO epochs on Hovering/CTBR through Runner: losses finite, KL-driven LR moves, reward improves, checkpoint
    round-trips with the reference's key layout; CUDA-graph mode equals eager mode in distribution (same code path)."""
    from airgym_b200.lib.config import default_ppo_config, scale_minibatch
    from airgym_b200.lib.torch_runner import Runner

    cfg = scale_minibatch(default_ppo_config("hovering"), 2048)
    c = cfg["params"]["config"]
    c.update(max_epochs=12, train_dir=str(tmp_path), use_cuda_graph=graph, fused_mlp=fused, print_stats=False, save_best_after=1)
    c["env_config"].update(ctl_mode="rate", num_envs=2048, seed=3)
    cfg["params"]["seed"] = 3
    r = Runner()
    r.load(cfg)
    r.run({"train": True})
    hist = r.agent.history
    assert len(hist) == 12 and all(np.isfinite([h["a_loss"], h["c_loss"], h["kl"]]).all() for h in hist)
    assert hist[-1]["frame"] == 12 * 2048 * 24
    rewards = [h["mean_reward"] for h in hist if h["mean_reward"] is not None]
    assert len(rewards) >= 2
    assert hist[-1]["c_loss"] < hist[0]["c_loss"]  # the critic fits the normalised returns
    ck = r.agent.save(os.path.join(str(tmp_path), "ck"))
    w = torch.load(ck, map_location="cpu", weights_only=False)
    assert set(w) == {"model", "epoch", "frame", "optimizer", "last_mean_rewards", "env_state"}
    assert {"logstd", "actor_mlp.layers.0.weight", "mu.weight", "value_head.bias", "value_mean_std.running_mean",
            "running_mean_std.count"} <= set(w["model"])
    before = r.agent.flat_params.clone()
    r.agent.flat_params.add_(1.0)
    r.agent.restore(ck)
    assert torch.equal(r.agent.flat_params, before)
    # TensorBoard scalars under the reference's tags (a2c_base.py:318-336, isaacgym_utils.py:86-99)
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    r.agent.writer.flush()
    acc = EventAccumulator(r.agent.summaries_dir)
    acc.Reload()
    tags = set(acc.Tags()["scalars"])
    assert {"performance/step_inference_rl_update_fps", "losses/a_loss", "losses/c_loss", "losses/entropy", "losses/bounds_loss",
            "info/last_lr", "info/kl", "info/epochs", "rewards/step", "rewards/iter", "episode_lengths/step",
            "Episode/pos_reward", "Episode/ups_reward", "Episode/thrust_reward", "Episode/reward"} <= tags
    assert len(acc.Scalars("losses/a_loss")) == 12 and acc.Scalars("info/epochs")[-1].step == hist[-1]["frame"]
    ev = acc.Scalars("Episode/pos_reward")
    assert len(ev) == 12 and all(np.isfinite(e.value) and 0.0 < e.value <= 1.5 for e in ev)


@pytest.mark.parametrize("task,B", [("hovering", 2048), ("tracking", 1000), ("hovering", 65536), ("hovering", 8)])
def test_fused_mlp_forward_backward_vs_autograd(built, task, B):
    """agx_mlp_forward / agx_mlp_backward (TF32 tensor cores) against the fp32 torch model + autograd: outputs to ~1e-3,
    parameter gradients to ~1e-2 of their scale (TF32 has a 10-bit mantissa)."""
    from airgym_b200.lib.config import default_ppo_config
    from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    torch.manual_seed(4)
    OBS, A = (48, 4) if task == "tracking" else (18, 4)
    model = ModelA2CContinuousLogStd(default_ppo_config(task)["params"], {"actions_num": A, "input_shape": (OBS,)}).cuda()
    with torch.no_grad():
        model.running_mean_std.running_mean.copy_(torch.randn(OBS, dtype=torch.float64) * 0.3)
        model.running_mean_std.running_var.copy_(torch.rand(OBS, dtype=torch.float64) + 0.5)
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    model.flatten_parameters()
    model.eval()
    obs = torch.randn(B, OBS, device="cuda") * 2.0
    mu_ref, v_ref = model.heads(obs)
    dims = model.fused_keep_dims()
    keep = tuple(torch.zeros(B, d, device="cuda") for d in dims)
    mu, value = torch.zeros(B, A, device="cuda"), torch.zeros(B, device="cuda")
    # every forward implementation: mma.sync (0), tcgen05 for inference calls only (1), tcgen05 always (2 = default, last)
    for mode, with_keep in ((0, True), (1, False), (1, True), (2, False), (2, True)):
        _capi.check(_capi.load().agx_set_option(b"mlp_forward", mode))
        mu.zero_(); value.zero_()
        for k in keep:
            k.zero_()
        model.fused_heads(obs, mu, value, keep=keep if with_keep else None)
        assert_close(mu.cpu(), mu_ref.detach().cpu(), f"mu (mode {mode})", rtol=5e-3, atol=2e-3)
        assert_close(value.cpu(), v_ref.detach().squeeze(-1).cpu(), f"value (mode {mode})", rtol=5e-3, atol=2e-3)
        if with_keep:
            assert_close(keep[0][:, :OBS].cpu(), model.norm_obs(obs).cpu(), "normalised input", rtol=1e-6, atol=1e-6)
            with torch.no_grad():
                h1_ref = torch.nn.functional.elu(model.actor_mlp.layers[0](model.norm_obs(obs)))
            assert_close(keep[1].cpu(), h1_ref.cpu(), f"h1 (mode {mode})", rtol=5e-3, atol=3e-3)
    assert_close(mu.cpu(), mu_ref.detach().cpu(), "mu", rtol=5e-3, atol=2e-3)
    assert_close(value.cpu(), v_ref.detach().squeeze(-1).cpu(), "value", rtol=5e-3, atol=2e-3)
    assert_close(keep[0][:, :OBS].cpu(), model.norm_obs(obs).cpu(), "normalised input", rtol=1e-6, atol=1e-6)
    assert float(keep[0][:, OBS:].abs().max()) == 0.0 if dims[0] > OBS else True
    g_mu, g_v = torch.randn(B, A, device="cuda") / B, torch.randn(B, device="cuda") / B
    for p in model.parameters():
        p.grad.zero_()
    torch.autograd.backward((mu_ref, v_ref), (g_mu, g_v.view(-1, 1)))
    ref_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad.fill_(7.0)  # the kernel overwrites (does not accumulate)
    dz = tuple(torch.zeros(B, d, device="cuda") for d in dims[1:])
    dout = torch.zeros(B, 16, device="cuda")
    ws = model.fused_workspace("cuda")
    model.fused_backward(g_mu, g_v, keep, dz, dout, ws)
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        if n == "logstd":
            continue
        scale = float(ref_grads[n].abs().max()) + 1e-12
        assert_close((p.grad / scale).cpu(), (ref_grads[n] / scale).cpu(), f"grad {n}", rtol=2e-2, atol=1e-2)
    first = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.fused_backward(g_mu, g_v, keep, dz, dout, ws)  # deterministic: bitwise identical on a second run
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        assert torch.equal(p.grad, first[n]), n


def _rows(planes):
    """feature-major keep tensor, blocked by 128-row tile ([B/128][W][128], stored as a [W, B] allocation) → row-major [B, W]"""
    W, B = planes.shape
    return planes.reshape(B // 128, W, 128).permute(0, 2, 1).reshape(B, W)


@pytest.fixture(params=[1, 0], ids=["wgrad-tma", "wgrad-gather"])
def wgrad_impl(request):
    """weight-gradient operands by TMA boxes (default) or by the cp.async gather kernel (agx_set_option("mlp_wgrad_tma"))"""
    from airgym_b200 import _capi

    lib = _capi.load()
    _capi.check(lib.agx_set_option(b"mlp_wgrad_tma", request.param), "mlp_wgrad_tma")
    yield request.param
    lib.agx_set_option(b"mlp_wgrad_tma", 1)


@pytest.mark.parametrize("OBS,B", [(18, 2048), (18, 65536), (48, 1024), (46, 256), (18, 128), (80, 512)])
def test_tcgen05_train_path_vs_autograd(built, OBS, B, wgrad_impl):
    """agx_mlp_forward_train / agx_mlp_backward_train — forward, activation-gradient chain, weight AND bias gradients all on
    tcgen05.mma (feature-major intermediates) — against the fp32 torch model + autograd, at the TF32 bound of the mma.sync path:
    outputs 5e-3 rel + 2e-3 abs, parameter gradients 2e-2 of their scale; bitwise reproducible.  Both weight-gradient kernels."""
    from airgym_b200.lib.config import default_ppo_config
    from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    torch.manual_seed(5)
    A = 4
    model = ModelA2CContinuousLogStd(default_ppo_config("hovering")["params"], {"actions_num": A, "input_shape": (OBS,)}).cuda()
    with torch.no_grad():
        model.running_mean_std.running_mean.copy_(torch.randn(OBS, dtype=torch.float64) * 0.3)
        model.running_mean_std.running_var.copy_(torch.rand(OBS, dtype=torch.float64) + 0.5)
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    model.flatten_parameters()
    model.eval()
    assert model.train_supported(B) and not model.train_supported(B + 8)
    P = model.train_params()
    assert P.in_pad == {18: 32, 48: 64, 46: 48, 80: 96}[OBS]  # 80 = 16 + the VAE's 64 latents
    obs = torch.randn(B, OBS, device="cuda") * 2.0
    mu_ref, v_ref = model.heads(obs)
    keep, dz, dout = model.train_buffers(B, "cuda")
    mu, value = torch.zeros(B, A, device="cuda"), torch.zeros(B, device="cuda")
    model.fused_heads_train(obs, mu, value, keep)
    assert_close(mu.cpu(), mu_ref.detach().cpu(), "mu", rtol=5e-3, atol=2e-3)
    assert_close(value.cpu(), v_ref.detach().squeeze(-1).cpu(), "value", rtol=5e-3, atol=2e-3)
    xn = model.norm_obs(obs)
    x_rows = _rows(keep[0])
    assert_close(x_rows[:, :OBS].cpu(), xn.cpu(), "normalised input planes", rtol=1e-6, atol=1e-6)
    assert bool((x_rows[:, OBS] == 1.0).all()) and float(x_rows[:, OBS + 1:].abs().max() if P.in_pad > OBS + 1 else 0.0) == 0.0
    with torch.no_grad():
        h1_ref = torch.nn.functional.elu(model.actor_mlp.layers[0](xn))
        h2_ref = torch.nn.functional.elu(model.actor_mlp.layers[1](h1_ref))
    assert_close(_rows(keep[1]).cpu(), h1_ref.cpu(), "h1 planes", rtol=5e-3, atol=3e-3)
    assert_close(_rows(keep[2]).cpu(), h2_ref.cpu(), "h2 planes", rtol=5e-3, atol=5e-3)
    g_mu, g_v = torch.randn(B, A, device="cuda") / B, torch.randn(B, device="cuda") / B
    for p in model.parameters():
        p.grad.zero_()
    torch.autograd.backward((mu_ref, v_ref), (g_mu, g_v.view(-1, 1)))
    ref_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad.fill_(7.0)  # the kernels overwrite (do not accumulate)
    ws = model.fused_workspace("cuda")
    model.fused_backward_train(g_mu, g_v, keep, dz, dout, ws)
    torch.cuda.synchronize()
    assert_close(_rows(dout)[:, :A].cpu(), g_mu.cpu(), "dout planes (mu)", rtol=0, atol=0)
    assert_close(_rows(dout)[:, A].cpu(), g_v.cpu(), "dout plane (value)", rtol=0, atol=0)
    for n, p in model.named_parameters():
        if n == "logstd":
            continue
        scale = float(ref_grads[n].abs().max()) + 1e-12
        assert_close((p.grad / scale).cpu(), (ref_grads[n] / scale).cpu(), f"grad {n}", rtol=2e-2, atol=1e-2)
    first = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.fused_backward_train(g_mu, g_v, keep, dz, dout, ws)
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        assert torch.equal(p.grad, first[n]), n
    if OBS == 80:
        return  # the warp-level mma.sync kernels do not cover this width (shared memory): nothing to cross-check against
    # and against the mma.sync backward on the same kept activations (row-major copies): same TF32 products, different summation order
    keep_r = tuple(_rows(k).contiguous() for k in keep)
    keep_r = (torch.nn.functional.pad(xn, (0, (OBS + 15) // 16 * 16 - OBS)),) + keep_r[1:]
    dims = model.fused_keep_dims()
    dz_r, dout_r = tuple(torch.zeros(B, d, device="cuda") for d in dims[1:]), torch.zeros(B, 16, device="cuda")
    model.fused_heads(obs, mu, value, keep=tuple(torch.zeros(B, d, device="cuda") for d in dims))  # builds model._fused
    model.fused_backward(g_mu, g_v, keep_r, dz_r, dout_r, ws)
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        if n == "logstd":
            continue
        scale = float(first[n].abs().max()) + 1e-12
        assert_close((p.grad / scale).cpu(), (first[n] / scale).cpu(), f"tcgen05 vs mma.sync grad {n}", rtol=5e-3, atol=3e-3)
    assert_close(dz_r[2].cpu(), _rows(dz[2]).cpu(), "dz3", rtol=5e-3, atol=1e-7 + 5e-3 * float(dz_r[2].abs().max()))
    assert_close(dz_r[0].cpu(), _rows(dz[0]).cpu(), "dz1", rtol=5e-3, atol=1e-7 + 1e-2 * float(dz_r[0].abs().max()))


def _rollout_agent(fused_rollout, N=4096, H=8, seed=4):
    from airgym_b200.lib.agent.a2c_continuous import A2CAgent
    from airgym_b200.lib.config import default_ppo_config, scale_minibatch
    from airgym_b200.lib.utils import tr_helpers

    cfg = scale_minibatch(default_ppo_config("hovering"), N)
    c = cfg["params"]["config"]
    c.update(horizon_length=H, minibatch_size=N * H // 4, print_stats=False, write_summaries=False, train_dir="/tmp/agx_ro",
             save_frequency=0, save_best_after=10**9, fused_rollout=fused_rollout, use_cuda_graph=False)
    c["env_config"].update(ctl_mode="rate", num_envs=N, seed=seed)
    c["reward_shaper"] = tr_helpers.DefaultRewardsShaper(**c["reward_shaper"])
    torch.manual_seed(seed)
    return A2CAgent("ro", cfg["params"])


def test_fused_rollout_matches_the_torch_rollout_and_the_oracle(built):
    """agx_policy_step + agx_rollout_post (3 launches per step) against the per-op torch rollout on identical explicit noise, and
    the sampling / neglogp / value de-normalisation against oracle/ppo.py (the reference's model code) on the recorded inputs."""
    N, H, A = 4096, 8, 4
    gen = torch.Generator().manual_seed(9)
    noise = torch.randn(H, N, A, generator=gen).cuda()
    agents = [_rollout_agent(False), _rollout_agent(True)]
    assert not agents[0].fused_rollout and agents[1].fused_rollout
    agents[1].flat_params.copy_(agents[0].flat_params)
    for ag in agents:
        with torch.no_grad():  # non-trivial normalisation statistics, identical in both
            ag.model.running_mean_std.running_mean.copy_(torch.linspace(-0.2, 0.2, 18, dtype=torch.float64))
            ag.model.running_mean_std.running_var.copy_(torch.linspace(0.5, 1.5, 18, dtype=torch.float64))
            ag.value_mean_std.running_mean.fill_(0.3)
            ag.value_mean_std.running_var.fill_(2.0)
        ag.noise_table = noise
        ag.env_reset()
        ag.env.progress_buf[::7] = ag.env.max_episode_length - 4  # time-outs inside the horizon → dones, bootstrap, episode statistics
        ag.ep_stats.zero_()
        ag.play_steps()
    torch.cuda.synchronize()
    a, b = agents
    for k in ("obses", "actions", "mus", "sigmas", "neglogpacs", "values", "rewards"):
        assert_close(b.buf[k].cpu(), a.buf[k].cpu(), "rollout buffer " + k, rtol=1e-5, atol=1e-5)
    assert torch.equal(a.buf["dones"], b.buf["dones"]) and torch.equal(a.dones, b.dones)
    assert int(a.buf["dones"].sum()) > 0
    assert_close(b.advs.cpu(), a.advs.cpu(), "advantages", rtol=1e-4, atol=1e-5)
    assert_close(b.ep_stats.cpu(), a.ep_stats.cpu(), "episode statistics", rtol=1e-6, atol=1e-6)
    assert float(a.ep_stats[3]) > 0
    assert_close(b.current_lengths.cpu(), a.current_lengths.cpu(), "running episode lengths", rtol=0, atol=0)
    # oracle: the reference's sampling / neglogp / denormalisation on what the fused path recorded
    from oracle import ppo as O
    mu, sigma = b.buf["mus"].cpu(), b.buf["sigmas"].cpu()
    z = noise.permute(1, 0, 2).cpu()
    act = mu + sigma * z
    assert_close(b.buf["actions"].cpu(), act, "sampled action", rtol=1e-6, atol=1e-6)
    nlp = O.neglogp(act, mu, sigma, torch.log(sigma))
    assert_close(b.buf["neglogpacs"].cpu(), nlp, "neglogp vs oracle", rtol=1e-5, atol=1e-5)
    sd = {k: v.cpu() for k, v in b.model.state_dict().items()}
    obs_flat = b.buf["obses"].reshape(-1, 18).cpu()
    mu_o, _ls, _sg, v_o = O.model_forward(sd, obs_flat, normalize_input=True)
    assert_close(mu.reshape(-1, A), mu_o, "mu vs oracle (TF32 bound)", rtol=5e-3, atol=2e-3)
    v_den = O.rms_denorm(v_o, sd["value_mean_std.running_mean"], sd["value_mean_std.running_var"])
    assert_close(b.buf["values"].reshape(-1, 1).cpu(), v_den, "value vs oracle (TF32 bound)", rtol=5e-3, atol=5e-3)


def test_fused_rollout_philox_sampling_is_standard_normal_and_partition_invariant(built):
    ag = _rollout_agent(True, N=8192, H=8)
    ag.env_reset()
    ag.play_steps()
    torch.cuda.synchronize()
    z = ((ag.buf["actions"] - ag.buf["mus"]) / ag.buf["sigmas"]).flatten()
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01 and float(z.abs().max()) < 6.5
    assert abs(float((z ** 4).mean()) - 3.0) < 0.1  # kurtosis of a normal
    # two shards of the env axis (env_offset) draw the same numbers as one env of twice the size
    half = [_rollout_agent(True, N=4096, H=8) for _ in range(2)]
    half[1].env.set_seed(half[1].env.rng_seed, env_offset=4096)
    for h in half:
        h.flat_params.copy_(ag.flat_params)
        h.env_reset()
        h.play_steps()
    torch.cuda.synchronize()
    assert torch.equal(torch.cat((half[0].buf["actions"], half[1].buf["actions"])), ag.buf["actions"])
    assert torch.equal(torch.cat((half[0].buf["rewards"], half[1].buf["rewards"])), ag.buf["rewards"])


@pytest.mark.parametrize("n,k", [(32768, 18), (1000, 1), (4096, 48), (777, 80), (300000, 46)])
def test_fused_rms_update_matches_the_reference_update(built, n, k):
    """agx_col_sums + agx_rms_merge against oracle/ppo.py rms_update (lib/core/running_mean_std.py:45-60) on float64 inputs, twice in
    a row so the merge with non-trivial running statistics is covered; deterministic."""
    import ctypes as C
    lib = _capi.load()
    torch.manual_seed(n + k)
    ws = torch.zeros(int(lib.agx_col_sums_workspace_doubles()), device="cuda", dtype=torch.float64)
    sums = torch.zeros(2 * k, device="cuda", dtype=torch.float64)
    mean, var, count = (torch.zeros(k, device="cuda", dtype=torch.float64), torch.ones(k, device="cuda", dtype=torch.float64),
                        torch.ones((), device="cuda", dtype=torch.float64))
    m, v, c = torch.zeros(k, dtype=torch.float64), torch.ones(k, dtype=torch.float64), torch.ones((), dtype=torch.float64)
    wide = torch.randn(n, k + 5, device="cuda") * 3.0 + 1.5
    for it in range(2):
        x = (wide[:, :k] * (1.0 + it)).contiguous() if it else wide[:, :k]  # first pass: a strided view (row stride k + 5)
        _capi.check(lib.agx_col_sums(x.data_ptr(), n, k, x.stride(0), sums.data_ptr(), ws.data_ptr(), None))
        first = sums.clone()
        _capi.check(lib.agx_col_sums(x.data_ptr(), n, k, x.stride(0), sums.data_ptr(), ws.data_ptr(), None))
        assert torch.equal(first, sums)
        _capi.check(lib.agx_rms_merge(sums.data_ptr(), k, float(n), mean.data_ptr(), var.data_ptr(), count.data_ptr(), None))
        m, v, c = O.rms_update(m, v, c, x.double().cpu())
        torch.cuda.synchronize()
        assert float(count) == float(c)
        assert float((mean.cpu() - m).abs().max()) < 1e-10 and float((var.cpu() - v).abs().max()) < 1e-9


def test_reference_golden_minibatch_through_the_fused_trainer_path(built):
    """The same recorded reference minibatch through the kernels the TRAINER actually runs by default — agx_mlp_forward_train (tcgen05,
    TF32 operands), agx_ppo_loss, agx_mlp_backward_train (tcgen05), agx_adam_step — with the TF32 bound stated: network outputs
    5e-3 rel + 2e-3 abs, loss statistics 2e-2 rel, gradients 2e-2 of each tensor's scale, gradient norm 1e-2; after the Adam step
    (first step: update = lr * g / |g| per element) parameters within 1e-5 for all but near-zero-gradient entries, which may move
    by at most 2 lr."""
    from airgym_b200.lib.config import default_ppo_config
    from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

    g = dict(np.load(os.path.join(GOLDEN_DIR, "ppo_hovering.npz")))
    lib = _capi.load()
    N, Hn, A, OBS = int(g["N"]), int(g["H"]), int(g["A"]), int(g["OBS"])
    model = ModelA2CContinuousLogStd(default_ppo_config("hovering")["params"], {"actions_num": A, "input_shape": (OBS,)}).cuda()
    model.load_state_dict({k[len("step0/sd_before/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("step0/sd_before/")})
    flat, grads = model.flatten_parameters(extra_grad_slots=8)
    fl, t = O.swap_and_flatten01, (lambda k: torch.from_numpy(g[k]))
    mb = N * Hn // 2
    sl = slice(0, mb)
    assert model.train_supported(mb)
    obs = fl(t("obs"))[sl].cuda().contiguous()
    model.running_mean_std.train()
    model.running_mean_std(obs)
    model.eval()
    keep, dz, dout = model.train_buffers(mb, "cuda")
    mu, value = torch.zeros(mb, A, device="cuda"), torch.zeros(mb, device="cuda")
    model.fused_heads_train(obs, mu, value, keep)
    assert_close(mu.cpu(), g["step0/mu"], "mu (TF32)", rtol=5e-3, atol=2e-3)
    assert_close(value.cpu(), g["step0/value"].reshape(-1), "value (TF32)", rtol=5e-3, atol=2e-3)
    d = lambda x: x.cuda().contiguous()
    om, os_ = d(fl(t("mus"))[sl]), d(fl(t("sigmas"))[sl])
    g_mu, g_val, g_ls = torch.zeros(mb, A, device="cuda"), torch.zeros(mb, device="cuda"), torch.zeros(A, device="cuda")
    stats = grads[model.num_flat:]
    ws = torch.zeros(int(lib.agx_ppo_workspace_floats()), device="cuda")
    Hh = _hyper()
    ins = [mu, model.logstd, value, d(fl(t("actions"))[sl]), d(fl(t("neglogpacs"))[sl]), d(t("advantages")[sl]), d(t("n_ret")[sl].reshape(-1))]
    _capi.check(lib.agx_ppo_loss(C.byref(Hh), mb, A, *[x.data_ptr() for x in ins], om.data_ptr(), os_.data_ptr(), g_mu.data_ptr(),
                                 g_val.data_ptr(), g_ls.data_ptr(), stats.data_ptr(), ws.data_ptr(), None))
    for j, k in enumerate(("a_loss", "c_loss", "entropy", "b_loss", "kl")):
        assert_close(stats[j].cpu(), g[f"step0/{k}"], k + " (TF32)", rtol=2e-2, atol=1e-5)
    model.fused_backward_train(g_mu, g_val, keep, dz, dout, model.fused_workspace("cuda"))
    model.logstd.grad.copy_(g_ls)
    for n, p in model.named_parameters():
        ref = torch.from_numpy(g[f"step0/grads/{n}"])
        scale = float(ref.abs().max()) + 1e-12
        assert_close((p.grad.cpu() / scale), ref / scale, f"grad {n} (TF32)", rtol=2e-2, atol=1e-2)
    # the trainer's default: the loss folded into the backward's first stage (agx_ppo_loss_backward_train) — same per-row arithmetic, so
    # the parameter gradients and the old_mu / old_sigma updates are bit-identical, the statistics equal up to the summation order
    sep = {n: p.grad.clone() for n, p in model.named_parameters()}
    sep_stats, sep_gls, sep_om, sep_os = stats[:5].clone(), g_ls.clone(), om.clone(), os_.clone()
    om2, os2 = d(fl(t("mus"))[sl]), d(fl(t("sigmas"))[sl])
    g_ls2, stats2 = torch.zeros(A, device="cuda"), torch.zeros(_capi.AGX_PPO_STATS, device="cuda")
    assert lib.agx_sizeof_loss_io() == C.sizeof(_capi.AgxLossIO)
    lio = _capi.AgxLossIO()
    lio.mu, lio.logstd, lio.value, lio.actions, lio.old_neglogp, lio.adv, lio.returns = [x.data_ptr() for x in ins]
    lio.old_mu, lio.old_sigma, lio.grad_logstd, lio.stats, lio.workspace, lio.a = om2.data_ptr(), os2.data_ptr(), g_ls2.data_ptr(), stats2.data_ptr(), ws.data_ptr(), A
    for p in model.parameters():
        p.grad.fill_(3.0)
    model.fused_loss_backward_train(Hh, lio, keep, dz, dout, model.fused_workspace("cuda"))
    torch.cuda.synchronize()
    assert torch.equal(om2, sep_om) and torch.equal(os2, sep_os)
    assert_close(stats2[:5].cpu(), sep_stats.cpu(), "fused loss statistics", rtol=2e-5, atol=1e-7)
    assert_close(g_ls2.cpu(), sep_gls.cpu(), "fused grad_logstd", rtol=2e-5, atol=1e-8)
    for n, p in model.named_parameters():
        if n != "logstd":
            assert torch.equal(p.grad, sep[n]), f"fused loss + backward: grad {n}"
    model.logstd.grad.copy_(g_ls)
    m, v = torch.zeros(model.num_flat, device="cuda"), torch.zeros(model.num_flat, device="cuda")
    lr_in = float(g["step0/lr_in"])
    lr_dev, step, norm = torch.tensor([lr_in], device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, device="cuda")
    _capi.check(lib.agx_adam_step(C.byref(Hh), model.num_flat, flat.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(), lr_dev.data_ptr(),
                                  step.data_ptr(), stats[4:5].data_ptr(), 1.0, norm.data_ptr(), None))
    torch.cuda.synchronize()
    assert_close(norm.cpu()[0], g["step0/total_norm"], "grad norm (TF32)", rtol=1e-2, atol=0)
    worst, loose, total = 0.0, 0, 0
    for n, p in model.named_parameters():
        diff = (p.detach().cpu() - torch.from_numpy(g[f"step0/sd_after/{n}"])).abs()
        worst, loose, total = max(worst, float(diff.max())), loose + int((diff > 1e-5).sum()), total + diff.numel()
    assert worst <= 2.1 * lr_in and loose <= 0.01 * total, (worst, loose, total)
    assert float(lr_dev) == pytest.approx(float(g["step0/lr_out"]), rel=1e-6)
