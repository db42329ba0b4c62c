"""GPU parity tests proper: the sm_100a kernel, called through the C ABI via the host-side env mirror, against
(1) the oracle on identical seeded inputs, (2) the golden fixtures recorded from the reference's own code, and
(3) size-independent properties at BASELINE.json's full size (65 536 envs)."""
import ctypes as C

import numpy as np
import pytest
import torch

from airgym_b200 import _capi
from oracle import QuadSpec, make_oracle
from tests.util import assert_close, golden_cases, load_golden, pokes_at, task_tols

pytestmark = pytest.mark.gpu
MODES = ["pos", "vel", "atti", "rate", "prop"]


def make_env(task, mode, N, seed=0, **cfg_over):
    from airgym_b200.envs import task_registry
    from airgym_b200.utils.helpers import get_args

    env, _ = task_registry.make_env(task, get_args(["--ctl_mode", mode, "--num_envs", str(N), "--headless", "--seed", str(seed)]))
    return env


def sync_from_oracle(env, orc, K):
    env.root_states.copy_(orc.root_states)
    env.pre_actions.copy_(orc.pre_actions)
    env.progress_buf.copy_(orc.progress_buf)
    env.reset_buf.copy_(orc.reset_buf)
    if K:
        env.ctrl_state[:K].copy_(orc.controller.state.T[:K])
    if hasattr(orc, "aux_matrix"):
        env.aux.copy_(orc.aux_matrix())
    if hasattr(orc, "asset_matrix"):
        env.assets.copy_(orc.asset_matrix())


def well_conditioned(orc, pre_q, mode):
    """Modes that run the attitude loop (CTA, LV, PY).  The reduced-attitude law (oracle/px4_controller.py attitude_loop) has two ill-conditioned
    corners: the shortest rotation between current and commanded thrust axis, normalize(cross(ez,ezd), 1+ez.ezd), is
    conditioned like 1/(1+ez.ezd); and the yaw part uses asin(q_mix.z), conditioned like 1/sqrt(1-z^2) = 1/|q_mix.w|
    (yaw error near 180 deg).  Even the oracle in fp32 vs fp64 differs by >1e-4 there.  Uniform random quaternion
    actions land in those corners a few percent of the time; such envs are compared at a looser bound."""
    if mode in ("rate", "prop"):
        return torch.ones(orc.num_envs, dtype=torch.bool)
    d, mw = orc.controller.last_conditioning
    return (d > -0.75) & (mw > 0.15)


@pytest.fixture(autouse=True)
def _default_options(built):
    lib = _capi.load()
    lib.agx_set_option(b"block", 128)
    lib.agx_set_option(b"use_bulk", 1)
    yield
    lib.agx_set_option(b"block", 128)
    lib.agx_set_option(b"use_bulk", 1)


@pytest.mark.parametrize("task", ["hovering", "tracking", "balloon"])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("variant", [(128, 1), (64, 1), (128, 0)])
def test_per_step_parity_vs_oracle(task, mode, variant):
    """Identical state, action and random draws in → state/obs/reward/reset out within 1e-4 rel (N=1000 has a
    partial last tile, so both the TMA bulk path and the cooperative-copy path run)."""
    block, use_bulk = variant
    lib = _capi.load()
    lib.agx_set_option(b"block", block)
    lib.agx_set_option(b"use_bulk", use_bulk)
    torch.manual_seed(11)
    N, T = 1000, 12
    spec = QuadSpec(task=task, ctl_mode=mode)
    orc = make_oracle(spec, N, rng="torch")
    env = make_env(task, mode, N)
    K = spec.ctrl_state_dim
    for t in range(T):
        a = torch.rand(N, spec.num_actions) * 2 - 1
        if task == "balloon" and mode in ("rate", "atti"):
            a[:, -1] = a[:, -1] * 0.3 - 0.5
        if t == 5:
            orc.progress_buf[:37] = spec.max_episode_length - 2
        sync_from_oracle(env, orc, K)
        a_dev = a.cuda()
        orc.step(a)
        d = orc.last_draws
        pre_q = orc.pre_step_quat
        obs, _, rew, reset, extras = env.step(a_dev, rand_reset=d["reset"].cuda(), rand_noise=d["noise"].cuda())
        tag = f"{task}/{mode}/{variant} t={t}"
        ok = well_conditioned(orc, pre_q, mode)
        # measured on the oracle with these seeds: every env qualifies in pos / vel / rate / prop; with uniformly random quaternion
        # set-points (atti) 83-86 % do (P(ez.ezd < -0.75) alone is 12.5 %) — a regression cannot hide behind a shrinking mask
        assert ok.float().mean() >= (0.82 if mode == "atti" else 1.0), float(ok.float().mean())
        assert_close(env.root_states.cpu()[ok], orc.root_states[ok], tag + " state")
        assert_close(obs.cpu()[ok], orc.obs_buf[ok], tag + " obs")
        rr, ra = task_tols(task)
        assert_close(rew.cpu()[ok], orc.rew_buf[ok], tag + " rew", rtol=rr, atol=ra)
        assert_close(env.cmd_thrusts.cpu()[ok], orc.cmd_thrusts[ok], tag + " cmd")
        assert_close(env._reward_terms.cpu()[:len(type(orc).REWARD_KEYS), ok], orc.reward_terms_matrix()[:, ok], tag + " terms", rtol=rr, atol=ra)
        if hasattr(orc, "aux_matrix"):
            assert_close(env.aux.cpu()[ok], orc.aux_matrix()[ok], tag + " aux")
        # ill-conditioned attitude set-points (see well_conditioned) still agree, just not to 1e-4
        assert_close(env.cmd_thrusts.cpu(), orc.cmd_thrusts, tag + " cmd (all)", rtol=5e-2, atol=5e-3)
        assert_close(env.root_states.cpu(), orc.root_states, tag + " state (all)", rtol=5e-2, atol=5e-3)
        assert_close(env.actions.cpu(), orc.actions, tag + " actions", rtol=0, atol=0)
        assert_close(env.pre_actions.cpu(), orc.pre_actions, tag + " pre_actions", rtol=0, atol=0)
        assert_close(a_dev.cpu(), a, tag + " in-place remap (Q4)", rtol=0, atol=0)
        if K:
            assert_close(env.ctrl_state[:K].T.cpu()[ok], orc.controller.state[:, :K][ok], tag + " ctrl")
        assert torch.equal(reset.cpu()[ok], orc.reset_buf[ok]), tag
        assert torch.equal(env.progress_buf.cpu()[ok], orc.progress_buf[ok]), tag
        assert torch.equal(extras["time_outs"].cpu(), orc.time_out_buf), tag


@pytest.mark.parametrize("name", golden_cases())
def test_trajectory_vs_reference_golden(name):
    """Free-running trajectory from construction, fed the draws the reference consumed; compared with what the
    reference's own task code produced (tests/golden/make_golden.py)."""
    g, task, mode, N, T, A, max_len = load_golden(name)
    env = make_env(task, mode, N)
    env.params.max_episode_length = max_len
    r_idx = 0
    for t in range(T):
        a = torch.from_numpy(g["action_in"][t].copy()).cuda()
        tag = f"{name} t={t}"
        for env_i, st in pokes_at(g, t):
            env.root_states[env_i] = torch.from_numpy(st).cuda()
        if "rendered" in g:  # depth-camera tasks: dict observation; the image noise of a render step is explicit too
            img = None
            if g["rendered"][t]:
                img = {k: torch.from_numpy(g["img_" + k][r_idx]).cuda().contiguous() for k in ("add", "mul", "kern")}
            obs, _, rew, reset, extras = env.step(a, rand_reset=torch.from_numpy(g["draw_reset"][t]).cuda(), rand_image=img)
            if g["rendered"][t]:
                assert_image_close(obs["image"].cpu(), g["image"][r_idx], tag + " image")
                r_idx += 1
            obs = obs["observation"]
            if "assets" in g:
                assert_close(env.assets.cpu(), g["assets"][t], tag + " assets", rtol=1e-5, atol=2e-6)
        else:
            obs, _, rew, reset, extras = env.step(a, rand_reset=torch.from_numpy(g["draw_reset"][t]).cuda(),
                                                  rand_noise=torch.from_numpy(g["draw_noise"][t]).cuda())
        assert_close(env.root_states.cpu(), g["state"][t], tag + " state", rtol=3e-4, atol=1e-4)
        assert_close(obs.cpu(), g["obs"][t], tag + " obs", rtol=3e-4, atol=1e-4)
        ra = 5e-3 if task == "balloon" else 1e-4
        assert_close(rew.cpu(), g["rew"][t], tag + " rew", rtol=3e-4, atol=ra)
        assert_close(env._reward_terms.cpu()[:g["terms"][t].shape[0]], g["terms"][t], tag + " terms", rtol=3e-4, atol=ra)
        if "aux" in g:
            assert_close(env.aux.cpu(), g["aux"][t], tag + " aux", rtol=3e-4, atol=1e-4)
        assert np.array_equal(reset.cpu().numpy(), g["reset"][t]), tag
        assert np.array_equal(env.progress_buf.cpu().numpy(), g["progress"][t]), tag
        assert np.array_equal(extras["time_outs"].cpu().numpy(), g["timeout"][t]), tag
        assert_close(a.cpu(), g["action_in_after"][t], tag + " Q4", rtol=0, atol=0)


def assert_image_close(got, ref, what, frac=2e-3):
    """Depth images agree except at silhouette pixels (a grazing ray may hit in one fp32 evaluation order and miss in the
    other; the 5x5 blur spreads each such pixel over 25 outputs)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    bad = np.abs(got - ref) > 2e-4 + 1e-4 * np.abs(ref)
    assert bad.mean() <= frac, f"{what}: {bad.mean():.2%} of pixels differ (allowed {frac:.2%}); worst {np.abs(got - ref).max():.3e}"


@pytest.mark.parametrize("task", ["avoid", "planning"])
@pytest.mark.parametrize("mode", ["pos", "vel", "rate", "prop"])
def test_image_tasks_per_step_parity_vs_oracle(task, mode):
    """Avoid / Planning: identical state, action, reset draws and image noise in → state, obs16, reward, reset, aux, asset
    scatter and the depth image out.  Steps 4, 8, 12 render (PHYSICS half → agx_render_depth → TASK half)."""
    torch.manual_seed(17)
    N, T = 200, 13
    spec = QuadSpec(task=task, ctl_mode=mode)
    orc = make_oracle(spec, N, rng="explicit")
    env = make_env(task, mode, N)
    K = spec.ctrl_state_dim
    W, H = _capi.AGX_CAM_W, _capi.AGX_CAM_H
    n_render = 0
    for t in range(T):
        a = torch.rand(N, 4) * 2 - 1
        if mode == "rate":
            a[:, -1] = a[:, -1] * 0.1 - 0.69
        if t == 6:
            orc.progress_buf[:17] = spec.max_episode_length - 2
        sync_from_oracle(env, orc, K)
        rr = torch.rand(N, 2, spec.reset_draws)
        img = {"add": 0.1 * torch.randn(N, W, H), "mul": 0.3 * torch.randn(N, W, H) + 1.0,
               "kern": torch.randint(0, 256, (N, 25)).float() / 256.0}
        a_dev = a.cuda()
        orc.step(a, rr, torch.zeros(N, 18), img)
        obs, _, rew, reset, extras = env.step(a_dev, rand_reset=rr.cuda(), rand_image={k: v.cuda() for k, v in img.items()})
        tag = f"{task}/{mode} t={t}"
        ok = well_conditioned(orc, None, mode)
        assert_close(env.root_states.cpu()[ok], orc.root_states[ok], tag + " state")
        assert_close(obs["observation"].cpu()[ok], orc.obs_buf[ok], tag + " obs")
        assert_close(rew.cpu()[ok], orc.rew_buf[ok], tag + " rew")
        nt = len(type(orc).REWARD_KEYS)
        assert_close(env._reward_terms.cpu()[:nt, ok], orc.reward_terms_matrix()[:, ok], tag + " terms")
        assert_close(env.aux.cpu()[ok], orc.aux_matrix()[ok], tag + " aux")
        assert_close(env.actions.cpu(), orc.actions, tag + " actions", rtol=0, atol=0)
        assert_close(a_dev.cpu(), a, tag + " in-place remap (Q4)", rtol=0, atol=0)
        assert torch.equal(reset.cpu()[ok], orc.reset_buf[ok]), tag
        assert torch.equal(env.progress_buf.cpu()[ok], orc.progress_buf[ok]), tag
        if task == "planning":
            assert_close(env.assets.cpu(), orc.asset_matrix(), tag + " assets", rtol=1e-5, atol=2e-6)
        if orc.rendered:
            n_render += 1
            assert_image_close(obs["image"].cpu(), orc.full_camera_array, tag + " image")
    assert n_render == 3


@pytest.mark.parametrize("task", ["avoid", "planning"])
def test_image_tasks_philox_mode_at_scale(task):
    """Perf mode (in-kernel Philox for resets and image noise), 4096 envs, 150 steps: finite outputs, images in range, resets
    happen, the reset sampler's ranges hold, determinism across two identically seeded envs."""
    N = 4096
    envs = [make_env(task, "rate", N, seed=3) for _ in range(2)]
    torch.manual_seed(0)
    n_reset = 0
    for t in range(150):
        a = (torch.rand(N, 4) * 2 - 1).cuda()
        a[:, 3] = a[:, 3] * 0.2 - 0.7
        outs = [e.step(a.clone()) for e in envs]
        n_reset += int(outs[0][3].sum())
    o0, o1 = outs
    assert torch.equal(o0[0]["image"], o1[0]["image"]) and torch.equal(o0[0]["observation"], o1[0]["observation"])
    assert torch.equal(envs[0].root_states, envs[1].root_states) and torch.equal(o0[2], o1[2])
    img = o0[0]["image"]
    assert torch.isfinite(img).all() and float(img.min()) >= 0.0 and float(img.max()) < 25.0 and float(img.mean()) > 1.0
    assert torch.isfinite(o0[0]["observation"]).all() and torch.isfinite(o0[2]).all()
    assert n_reset > N // 8, n_reset
    e = envs[0]
    if task == "planning":
        A = _capi.AGX_NUM_ASSETS
        x, y = e.assets[:, 1:A], e.assets[:, A + 1:2 * A]
        assert float(x.abs().max()) <= 8.0 and float(y.abs().max()) <= 4.0 and float(x.std()) > 3.0
        c, s_ = e.assets[:, 2 * A:3 * A], e.assets[:, 3 * A:4 * A]
        assert float((c * c + s_ * s_ - 1).abs().max()) < 1e-5
        assert float((e.goal_positions[:, 0] - 8.5).abs().max()) == 0 and float(e.goal_positions[:, 1].abs().max()) <= 1.5
    else:
        parked = e.object_positions[:, 0] == -999.0
        assert 0.1 < float(parked.float().mean()) < 0.3  # 20 % of episodes have no cube (avoid.py:96-99)


def test_config1_hovering_64_ctbr_200_steps():
    """BASELINE config 1: Hovering, 64 envs, CTBR, 200 oracle steps; kernel free-runs next to the oracle."""
    torch.manual_seed(0)
    N = 64
    spec = QuadSpec(task="hovering", ctl_mode="rate")
    orc = make_oracle(spec, N, rng="torch")
    env = make_env("hovering", "rate", N)
    for t in range(200):
        a = torch.rand(N, 4) * 2 - 1
        a[:, 3] = a[:, 3] * 0.2 - 0.6
        sync_from_oracle(env, orc, 6)  # resync each step: chaos would otherwise amplify 1-ulp differences
        a_dev = a.cuda()
        orc.step(a)
        d = orc.last_draws
        obs, _, rew, reset, _ = env.step(a_dev, rand_reset=d["reset"].cuda(), rand_noise=d["noise"].cuda())
        assert_close(env.root_states.cpu(), orc.root_states, f"t={t} state")
        assert_close(obs.cpu(), orc.obs_buf, f"t={t} obs")
        assert_close(rew.cpu(), orc.rew_buf, f"t={t} rew")
        assert torch.equal(reset.cpu(), orc.reset_buf)


def test_philox_stream_matches_host_build():
    """The in-kernel Philox draws are the documented stream (same integers as the g++ build of the same header)."""
    from tests.hostsim.driver import build

    lib, host = _capi.load(), build()
    n, seed, step, off = 4096, 0xABCDEF0123456789, 12345, 7
    for sid, width in ((0, 12), (1, 12), (2, 18)):
        out = torch.zeros(n, width, device="cuda")
        _capi.check(lib.agx_philox_fill(out.data_ptr(), n, width, sid, seed, step, off, None))
        ref = np.zeros((n, width), np.float32)
        host.hostsim_philox_fill(ref.ctypes.data, n, width, sid, seed, step, off)
        if sid < 2:
            assert np.array_equal(out.cpu().numpy(), ref)
        else:
            assert_close(out.cpu(), ref, "normals", rtol=1e-4, atol=1e-4)  # SFU lg2/sin/cos (~2^-21 abs; r = sqrt(-2 ln u1) amplifies it at small r) vs libm


def test_philox_mode_equals_explicit_mode():
    """Perf mode (in-kernel Philox) produces exactly what explicit mode produces when fed agx_philox_fill's numbers."""
    lib = _capi.load()
    N = 3000
    envs = [make_env("hovering", "rate", N, seed=5) for _ in range(2)]
    torch.manual_seed(1)
    for t in range(6):
        a = (torch.rand(N, 4) * 2 - 1).cuda()
        if t == 3:
            for e in envs:
                e.progress_buf[:500] = e.max_episode_length - 2
        rr = torch.zeros(N, 2, 12, device="cuda")
        tmp = torch.zeros(N, 12, device="cuda")
        for which in (0, 1):
            _capi.check(lib.agx_philox_fill(tmp.data_ptr(), N, 12, which, envs[0].rng_seed, t, 0, None))
            rr[:, which] = tmp
        nz = torch.zeros(N, 18, device="cuda")
        _capi.check(lib.agx_philox_fill(nz.data_ptr(), N, 18, 2, envs[0].rng_seed, t, 0, None))
        envs[0].step(a.clone())
        envs[1].step(a.clone(), rand_reset=rr, rand_noise=nz)
        assert torch.equal(envs[0].root_states, envs[1].root_states), t
        assert torch.equal(envs[0].obs_buf, envs[1].obs_buf), t
        assert torch.equal(envs[0].rew_buf, envs[1].rew_buf), t
        assert torch.equal(envs[0].reset_buf, envs[1].reset_buf), t


def test_full_size_properties_65536():
    """BASELINE config 2 size.  Properties that need no oracle: unit quaternions, canonical sign, determinism,
    partition invariance of the Philox stream (2 shards == 1 env of 65 536), reset-distribution moments, finite obs."""
    N = 65536
    full = make_env("hovering", "rate", N, seed=9)
    halves = [make_env("hovering", "rate", N // 2, seed=9) for _ in range(2)]
    halves[1].set_seed(9, env_offset=N // 2)
    twin = make_env("hovering", "rate", N, seed=9)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(50):
        a = torch.rand(N, 4, device="cuda", generator=g) * 2 - 1
        a[:, 3] = a[:, 3] * 0.2 - 0.6
        full.step(a.clone())
        twin.step(a.clone())
        halves[0].step(a[: N // 2].clone())
        halves[1].step(a[N // 2:].clone())
    assert torch.equal(full.root_states, twin.root_states) and torch.equal(full.obs_buf, twin.obs_buf)  # deterministic
    assert torch.equal(full.root_states[: N // 2], halves[0].root_states)
    assert torch.equal(full.root_states[N // 2:], halves[1].root_states)  # partition invariant
    assert torch.equal(full.rew_buf[N // 2:], halves[1].rew_buf)
    q = full.root_quats
    assert (q.norm(dim=-1) - 1).abs().max() < 1e-5
    assert torch.isfinite(full.obs_buf).all() and torch.isfinite(full.rew_buf).all()
    assert int(full._step_dev[0]) == 50 and int(full._step_dev[1]) == 0
    # reset distribution (hovering.py:316-329): force a reset of everything and look at the freshly drawn states
    full.reset_buf[:] = 1
    full.step(torch.zeros(N, 4, device="cuda"))
    full.progress_buf[:] = full.max_episode_length - 2
    full.step(torch.zeros(N, 4, device="cuda"))  # time-out → post-step reset_idx draws (visible in root_states)
    s = full.root_states
    assert (full.reset_buf == 1).all() and (full.progress_buf == 0).all() and (full.pre_actions == 0).all()
    assert s[:, 0:3].abs().max() <= 1.0 and abs(float(s[:, 0:3].mean())) < 0.01
    assert abs(float(s[:, 0:3].std()) - (1 / 3) ** 0.5) < 0.01
    assert s[:, 7:10].abs().max() <= 0.5 and s[:, 10:13].abs().max() <= 0.2
    assert abs(float(s[:, 7:10].std()) - 0.5 / 3**0.5) < 0.005 and abs(float(s[:, 10:13].std()) - 0.2 / 3**0.5) < 0.002
    assert (s[:, 6] > 0.99).all()


@pytest.mark.parametrize("task,mode", [("hovering", "rate"), ("tracking", "vel")])
def test_oracle_parity_at_full_size_65536(task, mode):
    """BASELINE configs 2 / 3 at their stated size against the ORACLE (not just properties): 3 explicit-randomness steps over
    65 536 envs, the second with a forced time-out wave so both reset passes run; same bar as the N = 1000 test."""
    N = 65536
    torch.manual_seed(21)
    spec = QuadSpec(task=task, ctl_mode=mode)
    orc = make_oracle(spec, N, rng="torch")
    env = make_env(task, mode, N)
    K = spec.ctrl_state_dim
    resets_seen = 0
    for t in range(3):
        a = torch.rand(N, spec.num_actions) * 2 - 1
        if t == 1:
            orc.progress_buf[::17] = spec.max_episode_length - 2
        sync_from_oracle(env, orc, K)
        a_dev = a.cuda()
        orc.step(a)
        d = orc.last_draws
        obs, _, rew, reset, extras = env.step(a_dev, rand_reset=d["reset"].cuda(), rand_noise=d["noise"].cuda())
        tag = f"{task}/{mode} N=65536 t={t}"
        assert_close(env.root_states.cpu(), orc.root_states, tag + " state")
        assert_close(obs.cpu(), orc.obs_buf, tag + " obs")
        rr, ra = task_tols(task)
        assert_close(rew.cpu(), orc.rew_buf, tag + " rew", rtol=rr, atol=ra)
        assert_close(env.cmd_thrusts.cpu(), orc.cmd_thrusts, tag + " cmd")
        assert_close(env.actions.cpu(), orc.actions, tag + " actions", rtol=0, atol=0)
        assert_close(env.ctrl_state[:K].T.cpu(), orc.controller.state[:, :K], tag + " ctrl")
        assert torch.equal(reset.cpu(), orc.reset_buf) and torch.equal(env.progress_buf.cpu(), orc.progress_buf), tag
        assert torch.equal(extras["time_outs"].cpu(), orc.time_out_buf), tag
        assert torch.equal(env.reset_u8.cpu().long(), orc.reset_buf), tag  # the byte copy of the flags in the packed results block
        resets_seen += int(orc.reset_buf.sum())
    assert resets_seen >= N // 17  # the time-out wave (and its re-reset the step after, quirk Q1) went through both reset passes


@pytest.mark.parametrize("task,mode", [("hovering", "rate"), ("hovering", "pos"), ("tracking", "vel"), ("balloon", "rate")])
def test_compute_observations_and_compute_reward_as_standalone_calls(task, mode):
    """The reference's method names do their work outside step() too (hovering.py:337-459): agx_observe = the TASK phase of the step
    kernel in observe mode, against the oracle's compute_observations() / compute_reward() on identical buffers."""
    N = 1000
    torch.manual_seed(5)
    spec = QuadSpec(task=task, ctl_mode=mode)
    orc = make_oracle(spec, N, rng="torch")
    env = make_env(task, mode, N)
    K = spec.ctrl_state_dim
    for t in range(3):  # a few real steps first, so actions / pre_actions / cmd_thrusts / progress are non-trivial
        a = torch.rand(N, spec.num_actions) * 2 - 1
        sync_from_oracle(env, orc, K)
        orc.step(a.clone())
        d = orc.last_draws
        env.step(a.cuda(), rand_reset=d["reset"].cuda(), rand_noise=d["noise"].cuda())
    sync_from_oracle(env, orc, K)
    env.actions.copy_(orc.actions)
    env.cmd_thrusts.copy_(orc.cmd_thrusts)
    before = (env.root_states.clone(), env.progress_buf.clone(), env.time_out_buf.clone(), int(env._step_dev[0]))
    orc.last_draws = {"reset": torch.zeros(N, 2, orc.RESET_DRAWS), "noise": torch.zeros(N, 18)}
    orc.compute_observations()
    obs = env.compute_observations(rand_noise=orc.last_draws["noise"].cuda())
    assert_close(obs.cpu(), orc.obs_buf, f"{task}/{mode} compute_observations")
    orc.compute_reward()
    rew = env.compute_reward()
    rr, ra = task_tols(task)
    assert_close(rew.cpu(), orc.rew_buf, f"{task}/{mode} compute_reward", rtol=rr, atol=ra)
    assert torch.equal(env.reset_buf.cpu(), orc.reset_buf) and torch.equal(env.reset_u8.cpu().long(), orc.reset_buf)
    assert_close(env.pre_actions.cpu(), orc.pre_actions, "pre_actions = actions.clone()", rtol=0, atol=0)
    assert_close(env._reward_terms.cpu()[:len(type(orc).REWARD_KEYS)], orc.reward_terms_matrix(), "item_reward_info", rtol=rr, atol=ra)
    if hasattr(orc, "aux_matrix"):
        assert_close(env.aux.cpu(), orc.aux_matrix(), "task state (pre_root_positions = root_positions.clone())")
    torch.cuda.synchronize()
    assert torch.equal(env.root_states, before[0]) and torch.equal(env.progress_buf, before[1]) and torch.equal(env.time_out_buf, before[2])
    assert int(env._step_dev[0]) == before[3] and int(env._step_dev[1]) == 0  # the Philox step counter is read, not advanced
    # a second compute_reward sees pre_actions == actions (the continuity term changes), like the reference
    orc.compute_reward()
    assert_close(env.compute_reward().cpu(), orc.rew_buf, "second compute_reward", rtol=rr, atol=ra)


def test_reset_idx_standalone_and_reset_api():
    N = 512
    env = make_env("tracking", "vel", N, seed=2)
    obs, priv = env.reset()
    assert obs.shape == (N, 48) and priv is None
    assert int(env.progress_buf.max()) == 1 and int(env.reset_buf.sum()) == 0
    ids = torch.tensor([3, 77, 500], device="cuda")
    u = torch.rand(3, 12, device="cuda")
    before = env.root_states.clone()
    env.reset_idx(ids, rand=u)
    torch.cuda.synchronize()
    spec = QuadSpec(task="tracking", ctl_mode="vel")
    orc = make_oracle(spec, N, rng="explicit")
    orc.reset_idx(ids.cpu(), u=u.cpu())
    assert_close(env.root_states[ids].cpu(), orc.root_states[ids.cpu()], "reset_idx rows")
    mask = torch.ones(N, dtype=torch.bool, device="cuda")
    mask[ids] = False
    assert torch.equal(env.root_states[mask], before[mask])
    assert env.reset_buf[ids].eq(1).all() and env.progress_buf[ids].eq(0).all()


def test_cuda_graph_replay_advances_rng():
    """The step is capturable: replaying a captured graph advances the device-side Philox step counter."""
    N = 4096
    env = make_env("hovering", "rate", N, seed=4)
    ref = make_env("hovering", "rate", N, seed=4)
    for e in (env, ref):  # both envs read the same action tensor: switch off the in-place remap write-back (Q4)
        e.params.flags &= ~_capi.FLAG_MUTATE_ACTIONS
    a = torch.zeros(N, 4, device="cuda")
    a[:, 3] = -0.6
    env.step(a)  # warm-up outside capture
    ref.step(a)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        env.step(a)
    for _ in range(5):
        graph.replay()
    # capture itself does not execute: the env saw 1 eager step + 5 replays
    for _ in range(5):
        ref.step(a)
    torch.cuda.synchronize()
    assert int(env._step_dev[0]) == int(ref._step_dev[0]) == 6
    assert torch.equal(env.root_states, ref.root_states) and torch.equal(env.obs_buf, ref.obs_buf)


def test_misaligned_and_bad_shapes_fail_loudly():
    env = make_env("hovering", "rate", 64)
    with pytest.raises(ValueError):
        env.step(torch.zeros(64, 5, device="cuda"))
    base = torch.zeros(64 * 4 + 1, device="cuda")
    off = base[1:].view(64, 4)  # 4-byte offset → not 16-B aligned
    with pytest.raises(_capi.AgxError, match="aligned"):
        env.step(off)


def test_balloon_api_and_ppo_epoch():
    """Balloon (Customized family): aux-backed attributes, privileged obs, and a PPO epoch through the trainer."""
    env = make_env("balloon", "rate", 256, seed=1)
    obs, priv = env.reset()
    assert obs.shape == (256, 18) and priv.shape == (256, 1, 13)
    assert float(env.balloon_positions[:, 0].min()) >= 2.0 and float(env.balloon_positions[:, 0].max()) <= 3.0
    assert torch.equal(priv[:, 0, 0:3], env.balloon_positions)
    assert set(env.item_reward_info) == {"guidance_reward", "hit_reward", "action_smoothness_reward", "effort_reward", "ups_reward",
                                         "yaw_reward", "reward"}
    from airgym_b200.lib.config import default_ppo_config, scale_minibatch
    from airgym_b200.lib.torch_runner import Runner

    cfg = scale_minibatch(default_ppo_config("balloon"), 1024)
    c = cfg["params"]["config"]
    c.update(max_epochs=3, train_dir="/tmp/agx_runs_test", print_stats=False, save_frequency=0, save_best_after=10**9)
    c["env_config"].update(ctl_mode="rate", num_envs=1024, seed=2)
    r = Runner()
    r.load(cfg)
    r.run({"train": True})
    assert len(r.agent.history) == 3 and all(np.isfinite(h["kl"]) for h in r.agent.history)
