"""Pin the oracle: replay every golden trajectory (recorded from the reference's own task code by
tests/golden/make_golden.py) through the oracle in explicit-randomness mode and demand agreement."""
import numpy as np
import pytest
import torch

from oracle import QuadSpec, make_oracle
from tests.util import assert_close, golden_cases, load_golden, pokes_at


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_trajectory(name):
    g, task, mode, N, T, A, max_len = load_golden(name)
    spec = QuadSpec(task=task, ctl_mode=mode)
    spec.episode_length_s = max_len * spec.dt + 1e-9
    assert spec.max_episode_length == max_len
    orc = make_oracle(spec, N, rng="explicit")
    r_idx = 0
    for t in range(T):
        a = torch.from_numpy(g["action_in"][t].copy())
        for env_i, st in pokes_at(g, t):
            orc.root_states[env_i] = torch.from_numpy(st)
        if task in ("avoid", "planning"):  # image tasks: the image noise of a render step is part of the explicit draws
            img = None
            if g["rendered"][t]:
                img = {k: torch.from_numpy(g["img_" + k][r_idx]) for k in ("add", "mul", "kern")}
            orc.step(a, torch.from_numpy(g["draw_reset"][t]), torch.from_numpy(g["draw_noise"][t]), img)
            if g["rendered"][t]:
                assert_close(orc.full_camera_array, g["image"][r_idx], f"{name} t={t} image", rtol=1e-5, atol=2e-5)
                r_idx += 1
            assert_close(orc.aux_matrix(), g["aux"][t], f"{name} t={t} aux", rtol=1e-5, atol=2e-5)
        else:
            orc.step(a, torch.from_numpy(g["draw_reset"][t]), torch.from_numpy(g["draw_noise"][t]))
        assert_close(orc.root_states, g["state"][t], f"{name} t={t} state", rtol=1e-5, atol=2e-6)
        assert_close(orc.obs_buf, g["obs"][t], f"{name} t={t} obs", rtol=1e-5, atol=2e-6)
        ra = 6e-5 if task == "balloon" else 2e-6  # balloon: 30x guidance amplification (tests/util.py task_tols)
        assert_close(orc.rew_buf, g["rew"][t], f"{name} t={t} rew", rtol=1e-5, atol=ra)
        assert_close(orc.reward_terms_matrix(), g["terms"][t], f"{name} t={t} terms", rtol=1e-5, atol=ra)
        assert_close(orc.cmd_thrusts, g["cmd"][t], f"{name} t={t} cmd", rtol=1e-5, atol=2e-6)
        assert_close(orc.actions, g["actions"][t], f"{name} t={t} actions", rtol=0, atol=0)
        assert_close(a, g["action_in_after"][t], f"{name} t={t} in-place action remap (Q4)", rtol=0, atol=0)
        assert np.array_equal(orc.reset_buf.numpy(), g["reset"][t])
        assert np.array_equal(orc.progress_buf.numpy(), g["progress"][t])
        assert np.array_equal(orc.time_out_buf.numpy(), g["timeout"][t])


def test_event_fixtures_contain_the_events():
    """The reference-pinned camera-task data holds a goal-reached, a heading < 0.25 and collision terminations (planning.py:263-287,
    avoid.py:271-284), not just free flight."""
    from oracle.image_tasks import PlanningOracle

    g, *_ = load_golden("planning_rate_events")
    k = list(PlanningOracle.REWARD_KEYS)
    terms = g["terms"]  # [T, K, N]
    assert (terms[:, k.index("reach_goal_reward")] > 0).sum() >= 2
    assert (terms[:, k.index("heading_reward")] < 0.25).sum() >= 2
    assert (g["aux"][:, :, 6] > 0).sum() >= 1 and g["reset"].sum() >= 4 + 5
    g, *_ = load_golden("avoid_rate_events")
    assert (g["aux"][:, :, 6] > 0).sum() >= 3 and g["rendered"].sum() == 5


def test_golden_covers_resets_and_timeouts():
    g, *_ = load_golden("hovering_rate_short")
    assert g["reset"].sum() > 16, "short-episode fixture must contain time-out resets beyond the initial one"
    g, *_ = load_golden("hovering_atti")
    assert g["reset"].sum() > 50, "atti fixture must contain q_w<0 resets (hovering.py:442-444)"


def test_known_answers_hover_equilibrium_and_free_fall():
    """KATs derivable from the cited constants: hover command 0.601*9.81/(4*9.59) gives zero
    acceleration; zero command gives free fall z = -g t^2 / 2; pure yaw couple spins about z only."""
    from oracle.rigid_body import simulate

    spec = QuadSpec(task="hovering", ctl_mode="prop")
    s = torch.zeros(3, 13, dtype=torch.float64)
    s[:, 6] = 1
    hover = spec.mass * spec.gravity / 4
    f = torch.tensor([[hover] * 4, [0.0] * 4, [hover] * 4], dtype=torch.float64)
    tau = torch.tensor([0.0, 0.0, 0.2 * 0.1], dtype=torch.float64)
    for _ in range(100):
        simulate(spec, s, f, tau)
    assert s[0, :3].abs().max() < 1e-9 and s[0, 7:10].abs().max() < 1e-9
    assert abs(s[1, 2].item() + 0.5 * 9.81) < 1e-9 and abs(s[1, 9].item() + 9.81) < 1e-9
    assert s[2, 10:12].abs().max() < 1e-12 and abs(s[2, 12].item() - 0.02 / spec.inertia[2]) < 1e-9
    assert abs(float(s[2, 3:7].norm()) - 1) < 1e-12
