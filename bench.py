#!/usr/bin/env python
"""bench.py — env-steps/sec of the fused Hovering/CTBR step at 65 536 envs per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one JSON line on rank 0)
    python bench.py --impl reference ...                           the reference's CPU path (oracle port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...              N ranks, envs sharded (weak scaling, no collective)

A "step" is one fused env step (agx_step) over one batch of 65 536 envs.  To keep the inputs larger than the 126 MB
L2 the bench rotates through R=8 independent env replicas (R x ~21 MB of state/obs/action buffers = 170 MB), one launch per
step, replayed from a CUDA graph; the e2e leg goes through the public env.step() with HOST action/result buffers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_ENVS = 65536
REPLICAS = 8
GRAPH_PASSES = 8  # the timed loop replays graphs of REPLICAS x GRAPH_PASSES = 64 step launches
ALGO_BYTES_PER_ENV_STEP = 288  # reads 113 B + writes 174 B (Hovering/CTBR fp32)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of agx_step_kernel<hovering, rate, 128> at 65 536 envs, from the
# `ncu --set full` capture in profiles/r1_agx_step_hovering_rate_ncu_raw.csv (2 launches: 8.18 MB read, 1.28 / 0 MB written — the
# 11.5 MB of outputs stay in the 126 MB L2 within one launch and are written back later)
NCU_DRAM_BYTES_PER_LAUNCH = 8.82e6
WORKLOAD = "Hovering, 65536 envs/GPU, CTBR (ctl_mode=rate), fused step kernel"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=NUM_ENVS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also print an N sweep (2^16..2^22) to stderr")
    ap.add_argument("--opt", action="append", default=[], help="libagx tuning knob key=value (agx_set_option)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (tuning runs only)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------
def _calibrate_threads(orc, a):
    """torch's intra-op pool over-subscribes badly on many-core hosts for these small element-wise ops: pick the
    thread count that makes one oracle step fastest (the CPU path gets "all the host threads it can use")."""
    import torch

    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best, best_t = 1, float("inf")
    cand = sorted({c for c in (1, 2, 4, 8, 16, 32, 64, 128, avail) if c <= avail})
    for c in cand:
        torch.set_num_threads(c)
        orc.step(a.clone())
        t0 = time.perf_counter()
        orc.step(a.clone())
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        if dt > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best, avail


def cpu_reference_shaped_steps_per_s(num_envs, cores, budget_s=6.0, min_steps=3):
    """SURVEY.md 8d's second CPU number: the same oracle step, but with the controller called the way the reference's host
    calls rlPx4Controller (hovering.py:246-250) — float64 numpy marshalling and a single-threaded per-env C loop
    (oracle/host_loop) — to show what that boundary costs next to the vectorised restatement."""
    import torch

    from oracle import QuadSpec, make_oracle
    from oracle.host_loop import HostLoopRateControl

    torch.manual_seed(0)
    torch.set_num_threads(cores)
    spec = QuadSpec(task="hovering", ctl_mode="rate")
    orc = make_oracle(spec, num_envs, rng="torch")
    orc.controller = HostLoopRateControl(num_envs, spec)
    a = torch.rand(num_envs, 4) * 2 - 1
    a[:, 3] = a[:, 3] * 0.2 - 0.6
    orc.step(a.clone())
    n, t0 = 0, time.perf_counter()
    while True:
        orc.step(a.clone())
        n += 1
        el = time.perf_counter() - t0
        if (el > budget_s and n >= min_steps) or n >= 200:
            break
    return num_envs * n / el, (f"{n} oracle steps of {num_envs} envs ({el:.1f} s) with the controller as a single-threaded per-env C "
                               f"loop behind float64 numpy marshalling (oracle/host_loop); the rest torch CPU fp32 on {cores} threads")


def cpu_oracle_steps_per_s(num_envs, budget_s=12.0, min_steps=3):
    """The reference's step() restated (oracle/) timed on this box's host cores: bounded sample."""
    import torch

    from oracle import QuadSpec, make_oracle

    torch.manual_seed(0)
    orc = make_oracle(QuadSpec(task="hovering", ctl_mode="rate"), num_envs, rng="torch")
    a = torch.rand(num_envs, 4) * 2 - 1
    a[:, 3] = a[:, 3] * 0.2 - 0.6
    cores, avail = _calibrate_threads(orc, a)
    n, t0 = 0, time.perf_counter()
    while True:
        orc.step(a.clone())
        n += 1
        el = time.perf_counter() - t0
        if (el > budget_s and n >= min_steps) or n >= 400:
            break
    return num_envs * n / el, cores, (f"{n} oracle steps of {num_envs} envs ({el:.1f} s, torch CPU fp32, {cores} threads "
                                      f"= fastest of the {avail} available)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" = a bounded sample: one oracle step over the same 65 536-env batch
    import torch

    from oracle import QuadSpec, make_oracle

    torch.manual_seed(0)
    N = args.num_envs
    orc = make_oracle(QuadSpec(task="hovering", ctl_mode="rate"), N, rng="torch")
    a = torch.rand(N, 4) * 2 - 1
    a[:, 3] = a[:, 3] * 0.2 - 0.6
    cores, avail = _calibrate_threads(orc, a)
    K = min(args.steps, 200)  # bounded: the CPU path needs ~50-100 ms per step
    W = min(args.warmup, 5)
    for _ in range(W):
        orc.step(a.clone())
    t0 = time.perf_counter()
    for _ in range(K):
        orc.step(a.clone())
    el = time.perf_counter() - t0
    v = N * K / el
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * el / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + " — reference step() restated on CPU (oracle/, torch fp32)", "num_envs": N,
                   "note": "IsaacGym/rlPx4Controller/pytorch3d are absent: the reference cannot run; this is the pinned CPU port"},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{K} oracle steps of {N} envs after {W} warm-up; {cores} threads = fastest of {avail} available"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def make_envs(num_envs, replicas, rank, world, device):
    import torch

    from airgym_b200.envs.base.hovering import Hovering
    from airgym_b200.envs.base.hovering_config import HoveringCfg

    envs = []
    for r in range(replicas):
        cfg = HoveringCfg()
        cfg.env.num_envs, cfg.env.ctl_mode, cfg.seed = num_envs, "rate", 1234 + r
        cfg.backend.reward_terms = False  # extras["item_reward_info"] is optional logging
        cfg.backend.export_cmd_thrusts = False  # internal attribute of the reference env, not part of step()'s return
        cfg.backend.mutate_input_actions = False  # the bench re-feeds one action tensor; Q4's write-back would drift it
        env = Hovering(cfg, None, None, device, True)
        env.set_seed(1234 + r, env_offset=rank * num_envs)
        envs.append(env)
    g = torch.Generator(device=device).manual_seed(5678 + rank)
    acts = []
    for r in range(replicas):
        a = torch.rand(num_envs, 4, device=device, generator=g) * 2 - 1  # U(-1,1)^4
        acts.append(a)
    return envs, acts


def time_kernel_loop(envs, acts, steps, warmup, dist):
    """K launches replayed from CUDA graphs (one launch per step, replica i mod R); device-timed, max over ranks."""
    import torch

    R = len(envs)
    P = GRAPH_PASSES  # launches per captured graph = R * P (graph-launch gaps amortised; every launch is still one step)
    for i in range(max(warmup, 3)):
        envs[i % R].step(acts[i % R])
    torch.cuda.synchronize()
    chunk = torch.cuda.CUDAGraph()
    with torch.cuda.graph(chunk):
        for _ in range(P):
            for r in range(R):
                envs[r].step(acts[r])
    singles = []
    for r in range(steps % (R * P)):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            envs[r % R].step(acts[r % R])
        singles.append(g)
    chunk.replay()  # graph warm-up
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps // (R * P)):
        chunk.replay()
    for g in singles:
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def time_e2e_loop(env, num_envs, steps, warmup, dist):
    """Public API with HOST buffers: per step H2D(actions) → env.step → D2H(obs, rew, reset); copies in the timed region."""
    import torch

    a_host = (torch.rand(num_envs, 4) * 2 - 1).pin_memory()
    obs_h = torch.empty(num_envs, env.num_obs).pin_memory()
    rew_h = torch.empty(num_envs).pin_memory()
    rst_h = torch.empty(num_envs, dtype=torch.long).pin_memory()
    a_dev = [torch.empty(num_envs, 4, device="cuda") for _ in range(2)]  # double-buffered: the H2D of step t+1 rides under the D2H of step t
    h2d = a_host.numel() * 4
    d2h = obs_h.numel() * 4 + rew_h.numel() * 4 + rst_h.numel() * 8
    main = torch.cuda.current_stream()
    copy_in = torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_step = [torch.cuda.Event() for _ in range(2)]

    def stage(t):  # host -> device copy of step t's actions, on the copy stream (PCIe is full duplex)
        with torch.cuda.stream(copy_in):
            copy_in.wait_event(ev_step[t % 2])  # the step that last read this buffer has run
            a_dev[t % 2].copy_(a_host, non_blocking=True)
            ev_in[t % 2].record(copy_in)

    def one(t):
        main.wait_event(ev_in[t % 2])
        obs, _, rew, rst, _ = env.step(a_dev[t % 2])
        ev_step[t % 2].record(main)
        stage(t + 1)
        obs_h.copy_(obs, non_blocking=True)
        rew_h.copy_(rew, non_blocking=True)
        rst_h.copy_(rst, non_blocking=True)

    for e in ev_step:
        e.record(main)
    stage(0)
    nw = max(3, min(warmup, 20))
    for t in range(nw):
        one(t)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(nw, nw + steps):
        one(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert float(rew_h.abs().sum()) > 0
    return ms, h2d, d2h


def run_ours(args):
    import torch

    import __graft_entry__ as graft

    from airgym_b200.dist_utils import rank_world

    rank, local, world = rank_world()
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        dist = dist_mod
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if rank == 0:
        graft.build()
    if dist is not None:
        dist.barrier()
    N, K, W = args.num_envs, args.steps, max(args.warmup, 3)

    from airgym_b200 import _capi

    opts = {}
    for kv in args.opt:
        k, v = kv.split("=")
        _capi.check(_capi.load().agx_set_option(k.encode(), int(v)), f"agx_set_option({kv})")
        opts[k] = int(v)
    envs, acts = make_envs(N, REPLICAS, rank, world, device)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = time_kernel_loop(envs, acts, K, W, dist)
    e2e_steps = max(20, min(K, 300))
    if args.no_e2e:
        ms_e2e, h2d, d2h = float("nan"), 0, 0
    else:
        ms_e2e, h2d, d2h = time_e2e_loop(envs[0], N, e2e_steps, W, dist)
    clocks = sampler.stop() if rank == 0 else None

    if args.sweep and rank == 0:
        for logn in (16, 18, 20, 22):
            n = 1 << logn
            reps = max(2, (REPLICAS << 16) // n)
            ev, ac = make_envs(n, reps, 0, 1, device)
            m = time_kernel_loop(ev, ac, 400, 50, None)
            gbs = ALGO_BYTES_PER_ENV_STEP * n * 400 / (m * 1e-3) / 1e9
            print(f"[sweep] N=2^{logn} replicas={reps} us/step={1e3 * m / 400:.2f} env-steps/s={n * 400 / (m * 1e-3):.3e} algo GB/s={gbs:.0f}",
                  file=sys.stderr, flush=True)
            del ev, ac
            torch.cuda.empty_cache()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    value = world * N * K / (ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_ENV_STEP * N / ((ms / K) * 1e-3) / 1e9
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "num_envs_per_gpu": N, "ctl_mode": "rate", "rng": "in-kernel Philox4x32-10",
                   "l2_policy": f"inputs larger than L2: {REPLICAS} independent env replicas rotated, one launch per step",
                   "launch": f"CUDA graph replay ({REPLICAS * GRAPH_PASSES} step launches per graph), programmatic dependent launch "
                             "(noise-first, agx.h 'pdl' auto)", "options": opts, "parallelism": f"env-sharded x{world}, no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if N == NUM_ENVS else None, "traffic_unit": "B/launch (ncu dram read+write; algorithmic: "
                     f"{ALGO_BYTES_PER_ENV_STEP * N} B/launch)", "peak_source": peak_src,
                     "note": f"{ALGO_BYTES_PER_ENV_STEP} algorithmic B/env-step x {N} envs / (timed region / launches), i.e. launch gaps included"},
        "e2e": {"value": world * N * e2e_steps / (ms_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "path": "env.step() via ctypes C ABI; pinned host actions in (copy stream, overlapping the previous step's read-back), "
                        "obs+reward+reset out to pinned host; every step moves all three"},
        "gpu_launches": K,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_oracle_steps_per_s(N)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
        try:  # an extra, informational number: it must never cost the bench line
            rv, rsample = cpu_reference_shaped_steps_per_s(N, cores)
            line["cpu_baseline"]["reference_shaped"] = {"value": rv, "unit": "env-steps/s", "sample": rsample}
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"]["reference_shaped"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
