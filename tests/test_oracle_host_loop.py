"""The per-env host-loop controller of the "reference-shaped" CPU baseline (oracle/host_loop) against the
vectorised torch restatement it stands beside (oracle/px4_controller.py) — both are test / baseline infrastructure."""
import torch

from oracle import QuadSpec, make_oracle
from oracle.host_loop import HostLoopRateControl
from oracle.px4_controller import ParallelControl


def test_host_loop_rate_controller_matches_vectorised_restatement():
    spec = QuadSpec(task="hovering", ctl_mode="rate")
    n = 257
    g = torch.Generator().manual_seed(3)
    vec = ParallelControl(n, spec, dtype=torch.float64)
    loop = HostLoopRateControl(n, spec)
    for step in range(6):
        q = torch.randn(n, 4, generator=g, dtype=torch.float64)
        q = q / q.norm(dim=-1, keepdim=True) * (1.0 + 1e-3 * torch.randn(n, 1, generator=g, dtype=torch.float64))  # not exactly unit
        a = torch.rand(n, 4, generator=g, dtype=torch.float64) * 2 - 1
        a[:, :3] *= 6.0
        a[:, 3] = 0.5 + 0.5 * a[:, 3]
        w = torch.randn(n, 3, generator=g, dtype=torch.float64) * (4.0 if step < 4 else 12.0)  # large errors exercise the integrator fade
        vec.set_q_world(q)
        ref = vec.update(a, w, 0.01)
        loop.set_q_world(q)
        got = loop.update(a, w, 0.01)
        assert torch.allclose(got, ref, rtol=0, atol=1e-12), float((got - ref).abs().max())
        assert torch.allclose(torch.from_numpy(loop.state), vec.state[:, :6], rtol=0, atol=1e-12)
    loop.reset([0, 5])
    assert (loop.state[[0, 5]] == 0).all() and (loop.state[1] != 0).any()


def test_oracle_steps_with_the_host_loop_controller():
    """Swapped into the oracle env, trajectories agree with the vectorised controller to fp32 rounding."""
    spec = QuadSpec(task="hovering", ctl_mode="rate")
    n = 64
    torch.manual_seed(0)
    a_env = make_oracle(spec, n, rng="explicit")
    b_env = make_oracle(spec, n, rng="explicit")
    b_env.controller = HostLoopRateControl(n, spec)
    g = torch.Generator().manual_seed(1)
    for t in range(20):
        act = torch.rand(n, 4, generator=g) * 2 - 1
        act[:, 3] = act[:, 3] * 0.2 - 0.6
        rr = torch.rand(n, 2, a_env.RESET_DRAWS, generator=g)
        rn = torch.randn(n, 18, generator=g)
        oa = a_env.step(act.clone(), rand_reset=rr, rand_noise=rn)
        ob = b_env.step(act.clone(), rand_reset=rr, rand_noise=rn)
        assert torch.allclose(oa[0], ob[0], rtol=1e-4, atol=1e-4)
        assert torch.equal(oa[3], ob[3])
