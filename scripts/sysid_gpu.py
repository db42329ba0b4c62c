#!/usr/bin/env python
"""Controller-convention probe on the GPU env: the reference's shipped planning policy
(trained/planning_cnn_rate.pth → tests/golden/planning_cnn_rate_model.pth) was trained against PhysX + rlPx4Controller, neither of
which is in the reference tree.  It is the only external behavioural truth available, so: fly it in the B200 Planning env under every
combination of the conventions the absent controller could have used and keep whatever lets it hold altitude and advance.

Swept: sign convention of the body-rate set-point (8 combinations; (+,-,-) is "the policy speaks FRD, the sim is FLU"), the rate the
unit command stands for (1 rad/s = customized.py:109-113 as written, 3.84 rad/s = PX4's 220 deg/s normalisation, 6 rad/s = the Hovering
limit), the collective-thrust map (k_thrust x {0.8, 1, 1.25}), the rate-loop gain (x {0.5, 1, 2}) and the D term on / off.
Per trial: 512 envs, 400 steps, deterministic policy.  Prints one JSON line per trial and the ten best by episode length.

    python scripts/sysid_gpu.py [--envs 512] [--steps 400]"""
import argparse
import copy
import importlib.util
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from airgym_b200.envs import task_registry  # noqa: E402
from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd  # noqa: E402
from airgym_b200.utils.helpers import get_args  # noqa: E402

_s = importlib.util.spec_from_file_location("pc", os.path.join(ROOT, "scripts", "play_ckpt.py"))
pc = importlib.util.module_from_spec(_s)
_s.loader.exec_module(pc)


def load_policy():
    P = copy.deepcopy(pc.PARAMS)
    keys = {"actions_num": 4, "input_shape": {"image": (1, 212, 120), "observation": (16,)}, "value_size": 1}
    m = ModelA2CContinuousLogStd(P, keys)
    m.load_state_dict(torch.load(os.path.join(ROOT, "tests", "golden", "planning_cnn_rate_model.pth"), weights_only=False)["model"])
    return m.cuda().eval()


def trial(model, signs, scale, thrust, gain, dterm, N, T, random_policy=False):
    env, _ = task_registry.make_env("planning", get_args(["--ctl_mode", "rate", "--num_envs", str(N), "--headless", "--seed", "0"]))
    p = env.params
    for i in range(3):
        p.act_lo[i], p.act_hi[i] = -scale, scale
        p.rate_p[i] *= gain
        if not dterm:
            p.rate_d[i] = 0.0
    p.k_thrust *= thrust
    F = torch.tensor([signs[0] * scale, signs[1] * scale, signs[2] * scale], device="cuda")
    obs, _, rew, reset, _ = env.step(torch.zeros(N, 4, device="cuda"))
    steps = torch.zeros(N, device="cuda")
    x0 = env.root_states[:, 0].clone()
    ep_len_sum = ep_cnt = prog_sum = 0.0
    tot_rew = 0.0
    z_err = 0.0
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(T):
        with torch.no_grad():
            if random_policy:
                mu = (torch.rand(N, 4, device="cuda", generator=g) * 2 - 1)
                mu[:, 3] = mu[:, 3] * 0.2 - 0.6
            else:
                mu = model({"is_train": False, "obs": obs})["mus"].clamp(-1, 1)
        a = mu.clone()
        a[:, :3] = mu[:, :3] * F
        x_before = env.root_states[:, 0].clone()
        obs, _, rew, reset, _ = env.step(a)
        obs["observation"][:, 12:15] = mu[:, :3]  # the policy sees its own command, whatever the controller made of it
        steps += 1
        tot_rew += float(rew.mean())
        z_err += float((env.root_states[:, 2] - 1.5).abs().mean())
        done = reset > 0
        if bool(done.any()):
            ep_len_sum += float(steps[done].sum())
            ep_cnt += float(done.sum())
            prog_sum += float((x_before[done] - x0[done]).sum())
            steps[done] = 0
            x0 = torch.where(done, env.root_states[:, 0], x0)
    ep_len_sum += float(steps.sum())
    ep_cnt += N
    prog_sum += float((env.root_states[:, 0] - x0).sum())
    return {"signs": signs, "rate_per_unit": scale, "thrust_x": thrust, "gain_x": gain, "d_term": dterm, "mean_ep_len": round(ep_len_sum / ep_cnt, 1),
            "mean_x_progress_m": round(prog_sum / ep_cnt, 3), "rew_per_step": round(tot_rew / T, 3), "mean_abs_z_err": round(z_err / T, 3)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--steps", type=int, default=400)
    a = ap.parse_args()
    model = load_policy()
    out = [dict(trial(model, (1.0, 1.0, 1.0), 1.0, 1.0, 1.0, True, a.envs, a.steps, random_policy=True), policy="random (baseline)")]
    print(json.dumps(out[0]), flush=True)
    for scale, signs, thrust in itertools.product((1.0, 3.84, 6.0), itertools.product((1.0, -1.0), repeat=3), (0.8, 1.0, 1.25)):
        r = trial(model, signs, scale, thrust, 1.0, True, a.envs, a.steps)
        out.append(r)
        print(json.dumps(r), flush=True)
    best = sorted(out[1:], key=lambda r: -r["mean_ep_len"])[:3]
    for b in best:  # around the best conventions: loop gain and D term
        for gain, dterm in ((0.5, True), (2.0, True), (1.0, False)):
            r = trial(model, tuple(b["signs"]), b["rate_per_unit"], b["thrust_x"], gain, dterm, a.envs, a.steps)
            out.append(r)
            print(json.dumps(r), flush=True)
    b = sorted(out[1:], key=lambda r: -r["mean_ep_len"])[0]  # refine thrust map and loop gain around the best convention, longer horizon
    for thrust, gain in itertools.product((0.6, 0.7, 0.8, 0.9), (0.2, 0.3, 0.5, 0.7)):
        r = trial(model, tuple(b["signs"]), b["rate_per_unit"], thrust, gain, True, a.envs, 2 * a.steps)
        r["stage"] = "refine (2x steps)"
        out.append(r)
        print(json.dumps(r), flush=True)
    print("BEST", json.dumps(sorted(out[1:], key=lambda r: -r["mean_ep_len"])[:10]))
