#!/usr/bin/env python
"""System identification by proxy (build container, CPU, oracle): the reference's shipped planning policy was trained against
rlPx4Controller, whose conventions are not in the reference tree.  Search over rate-command sign conventions and scales for the
one under which that policy flies (episode length, forward progress), using the CPU oracle.  Diagnostic only."""
import copy, itertools, json, sys, os, importlib.util
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import QuadSpec, make_oracle
from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

spec_ = importlib.util.spec_from_file_location("pc", os.path.join(ROOT, "scripts", "play_ckpt.py")); pc = importlib.util.module_from_spec(spec_); spec_.loader.exec_module(pc)
P = copy.deepcopy(pc.PARAMS)
keys = {"actions_num": 4, "input_shape": {"image": (1, 212, 120), "observation": (16,)}, "value_size": 1}
m = ModelA2CContinuousLogStd(P, keys); m.eval()
m.load_state_dict(torch.load(os.path.join(ROOT, "tests", "golden", "planning_cnn_rate_model.pth"), weights_only=False)["model"])

def trial(signs, scale, gain, N=6, T=200):
    torch.manual_seed(0)
    spec = QuadSpec(task="planning", ctl_mode="rate")
    spec.rate_p = [g * gain for g in spec.rate_p]
    o = make_oracle(spec, N, rng="torch")
    o.action_lower_limits = torch.tensor([-scale, -scale, -scale, 0.0]); o.action_upper_limits = torch.tensor([scale, scale, scale, 1.0])
    F = torch.tensor(list(signs) + [1.0]) * torch.tensor([scale, scale, scale, 1.0])
    obs, _, rew, reset, ex = o.step(torch.zeros(N, 4))
    steps = torch.zeros(N); ep_len = []; xmax = -8.5; tot = 0.0
    for t in range(T):
        with torch.no_grad():
            mu = m({"is_train": False, "obs": obs})["mus"].clamp(-1, 1)
        a = mu.clone()
        a[:, :3] = mu[:, :3] * F[:3]
        obs, _, rew, reset, ex = o.step(a)
        obs["observation"][:, 12:15] = mu[:, :3]  # the policy sees its own command
        steps += 1
        xmax = max(xmax, float(o.root_states[:, 0].max()))
        tot += float(rew.mean())
        for i in reset.nonzero().flatten().tolist():
            ep_len.append(float(steps[i])); steps[i] = 0
    ep_len += steps.tolist()
    return {"signs": signs, "scale": scale, "gain": gain, "mean_ep_len": sum(ep_len) / len(ep_len), "xmax": round(xmax, 2), "rew_per_step": round(tot / T, 3)}

if __name__ == "__main__":
    out = []
    for scale in (1.0, 3.84, 6.0):
        for signs in itertools.product((1.0, -1.0), repeat=3):
            for gain in (1.0,):
                r = trial(signs, scale, gain)
                out.append(r); print(json.dumps(r), flush=True)
    best = sorted(out, key=lambda r: -r["mean_ep_len"])[:5]
    print("BEST", json.dumps(best))
