#!/usr/bin/env python
"""Kernel A/B harness for the fused step: times CUDA-graph replay of agx_step for a list of cases in one GPU call.

    python scripts/kbench.py [--n 65536] [--steps 2000] case [case ...]
    case = name:key=value,...   keys: lib=<path of an alternative libagx build>  act=bench|hover|iid  noise=0|1
                                       n=<envs>  task=hovering|tracking  mode=rate|...  opt.<knob>=<int>
Each case runs in a child process (the library is loaded once per process).  Prints one line per case.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(spec):
    import torch

    from airgym_b200 import _capi
    from airgym_b200.envs import task_registry  # noqa: F401
    from airgym_b200.envs.base.hovering import Hovering
    from airgym_b200.envs.base.hovering_config import HoveringCfg
    from airgym_b200.envs.task.tracking import Tracking
    from airgym_b200.envs.task.tracking_config import TrackingCfg
    from airgym_b200.envs.task.avoid import Avoid
    from airgym_b200.envs.task.avoid_config import AvoidCfg
    from airgym_b200.envs.task.planning import Planning
    from airgym_b200.envs.task.planning_config import PlanningCfg

    CLS = {"hovering": (Hovering, HoveringCfg), "tracking": (Tracking, TrackingCfg), "avoid": (Avoid, AvoidCfg),
           "planning": (Planning, PlanningCfg)}

    n, steps = int(spec.get("n", 65536)), int(spec.get("steps", 2000))
    for k, v in spec.items():
        if k.startswith("opt."):
            _capi.check(_capi.load().agx_set_option(k[4:].encode(), int(v)), k)
    task, mode = spec.get("task", "hovering"), spec.get("mode", "rate")
    reps = max(2, (8 << 16) // n) if task in ("hovering", "tracking") else 1  # image tasks: 101 KB of image per env
    render_only = spec.get("render_only", "0") == "1"
    envs, acts = [], []
    g = torch.Generator(device="cuda").manual_seed(5678)
    for r in range(reps):
        cfg = CLS[task][1]()
        cfg.env.num_envs, cfg.env.ctl_mode, cfg.seed = n, mode, 1234 + r
        cfg.backend.reward_terms = False
        cfg.backend.export_cmd_thrusts = False
        cfg.backend.mutate_input_actions = False
        env = CLS[task][0](cfg, None, None, "cuda:0", True)
        if spec.get("noise", "1") == "0":
            env.params.flags |= _capi.FLAG_NO_NOISE
        envs.append(env)
        A = env.num_actions
        a = torch.rand(n, A, device="cuda", generator=g) * 2 - 1
        if spec.get("act", "bench") == "hover":
            a.zero_()
            a[:, A - 1] = -0.6934 if mode in ("rate", "atti") else 0.0
            if mode == "atti":
                a[:, 0] = 1.0
            if mode == "prop":
                a[:] = 0.1537
        acts.append(a)
    if task in ("avoid", "planning") and mode == "rate":
        for a in acts:
            a[:, 3] = a[:, 3] * 0.1 - 0.69  # near hover: episodes last
    for i in range(50):
        envs[i % reps].step(acts[i % reps])
    torch.cuda.synchronize()
    chunk = torch.cuda.CUDAGraph()
    passes = int(spec.get("chunk", 1))  # passes over the replicas per captured graph
    with torch.cuda.graph(chunk):
      for _ in range(passes):
        for r in range(reps):
            if render_only:
                envs[r].render_cameras()
            elif task in ("avoid", "planning"):
                for _ in range(4):  # one camera period: 3 fused steps + 1 split step with a render
                    envs[r].step(acts[r])
            else:
                envs[r].step(acts[r])
    for _ in range(30):
        chunk.replay()
    torch.cuda.synchronize()
    best = 1e9
    rates = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps // (reps * passes)):
            chunk.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (steps // (reps * passes) * reps * passes))
        rates.append(float(envs[0].reset_buf.float().mean()))
    if task in ("avoid", "planning") and not render_only:
        best /= 4
    print(json.dumps({"us_per_step": round(best, 3), "algo_GBs": round(288 * n / best / 1e3, 1),
                      "reset_frac": round(sum(rates) / len(rates), 4)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--child", type=str, default=None)
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    if a.child:
        child(json.loads(a.child))
        return
    for c in a.cases:
        name, _, kv = c.partition(":")
        spec = dict(x.split("=", 1) for x in kv.split(",") if x)
        spec.setdefault("steps", str(a.steps))
        env = dict(os.environ)
        if "lib" in spec:
            env["AGX_LIB"] = os.path.join(ROOT, spec["lib"])
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(spec)], env=env,
                           capture_output=True, text=True, timeout=600)
        out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("ERR " + r.stderr.strip()[-300:])
        print(f"[kbench] {name:28s} {kv:60s} {out}", flush=True)


if __name__ == "__main__":
    main()
