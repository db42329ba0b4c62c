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
GRAPH_PASSES = 8  # the timed loop replays graphs of REPLICAS x GRAPH_PASSES = 64 step launches (+ ONE graph of steps % 64)
MIN_TIMED_MS = 50.0  # the K-step block is repeated until this much device time has been measured; the MEDIAN block is reported
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
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO samples/s sub-record")
    ap.add_argument("--no-extras", action="store_true", help="skip the full-contract and 2^22-env side measurements")
    ap.add_argument("--ppo-envs", type=int, default=NUM_ENVS, help="envs per GPU of the PPO sub-record")
    ap.add_argument("--ppo-epochs", type=int, default=6, help="timed PPO epochs (after 3 warm-up / capture epochs)")
    ap.add_argument("--no-camera", action="store_true", help="skip the Avoid / Planning (configs 4-5) side records")
    ap.add_argument("--camera-envs", type=int, default=32768, help="envs per GPU of the Avoid / Planning side records")
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


def bench_config(num_envs, world, opts=None):
    """`config` of the JSON line — the same dict for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "num_envs_per_gpu": num_envs, "ctl_mode": "rate", "rng": "in-kernel Philox4x32-10",
            "l2_policy": f"inputs larger than L2: {REPLICAS} independent env replicas rotated, one launch per step",
            "launch": f"CUDA graph replay ({REPLICAS * GRAPH_PASSES} step launches per graph + one graph of steps % {REPLICAS * GRAPH_PASSES}), "
                      "programmatic dependent launch (noise-first, agx.h 'pdl' auto)",
            "timing": f"the K-step block is repeated until >= {MIN_TIMED_MS:.0f} ms are timed; median block, max over ranks per block",
            "options": opts or {}, "parallelism": f"env-sharded x{world}, no data-path collective"}


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
        "config": bench_config(N, args.gpus),
        "note": "reference step() restated on CPU (oracle/, torch fp32): IsaacGym/rlPx4Controller/pytorch3d are absent, the reference "
                "itself cannot run; the launch / l2_policy / rng entries of `config` describe the GPU arm's workload it is timed against",
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{K} oracle steps of {N} envs after {W} warm-up; {cores} threads = fastest of {avail} available"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def make_envs(num_envs, replicas, rank, world, device, full_contract=False):
    import torch

    from airgym_b200.envs.base.hovering import Hovering
    from airgym_b200.envs.base.hovering_config import HoveringCfg

    envs = []
    for r in range(replicas):
        cfg = HoveringCfg()
        cfg.env.num_envs, cfg.env.ctl_mode, cfg.seed = num_envs, "rate", 1234 + r
        cfg.backend.reward_terms = full_contract  # extras["item_reward_info"] is optional logging
        cfg.backend.export_cmd_thrusts = full_contract  # internal attribute of the reference env, not part of step()'s return
        cfg.backend.mutate_input_actions = full_contract  # the bench re-feeds one action tensor; Q4's write-back would drift it
        env = Hovering(cfg, None, None, device, True)
        env.set_seed(1234 + r, env_offset=rank * num_envs)
        envs.append(env)
    g = torch.Generator(device=device).manual_seed(5678 + rank)
    acts = []
    for r in range(replicas):
        a = torch.rand(num_envs, 4, device=device, generator=g) * 2 - 1  # U(-1,1)^4
        acts.append(a)
    return envs, acts


def timed_blocks(block, pre_roll, dist, what):
    """[pre-roll (untimed, keeps the stream in steady state) | e0 | block | e1] repeated until >= MIN_TIMED_MS of device time
    (at least 5, at most 400 blocks; the count is agreed across ranks); each block is bracketed by barrier + synchronize, its
    time is the MAX over ranks; returns the per-block times in ms."""
    import torch

    def once():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        pre_roll()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        block()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def reduce_max(vals):
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    first = reduce_max([once()])[0]
    n = int(min(400, max(5, -(-MIN_TIMED_MS // max(first, 1e-3)))))
    times = reduce_max([once() for _ in range(n)])
    if not all(t > 0 for t in times):
        raise RuntimeError(f"bench: non-positive {what} block time")
    return times


def median(v):
    v = sorted(v)
    return v[len(v) // 2] if len(v) % 2 else 0.5 * (v[len(v) // 2 - 1] + v[len(v) // 2])


def time_kernel_loop(envs, acts, steps, warmup, dist, fresh=None):
    """K launches replayed from CUDA graphs (one launch per step, replica i mod R): steps // 64 replays of the 64-launch graph +
    ONE graph holding the steps % 64 remaining launches; device-timed per block, max over ranks, median block.
    `fresh` (full-contract variant): pristine copies of the action tensors, copied in before every step because quirk Q4
    rewrites the caller's tensor in place — as a policy writing new actions every step would."""
    import torch

    R = len(envs)
    P = GRAPH_PASSES  # launches per captured graph = R * P (graph-launch gaps amortised; every launch is still one step)

    def step(i):
        r = i % R
        if fresh is not None:
            acts[r].copy_(fresh[r])
        envs[r].step(acts[r])

    for i in range(max(warmup, 3)):
        step(i)
    torch.cuda.synchronize()

    def capture(n_launches):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n_launches):
                step(i)
        return g

    n_chunks, n_rem = steps // (R * P), steps % (R * P)
    chunk = capture(R * P) if n_chunks else None
    rem = capture(n_rem) if n_rem else None
    pre = capture(R)
    for g in (chunk, rem, pre):
        if g is not None:
            g.replay()  # graph warm-up
    torch.cuda.synchronize()

    def block():
        for _ in range(n_chunks):
            chunk.replay()
        if rem is not None:
            rem.replay()

    times = timed_blocks(block, pre.replay, dist, "kernel")
    return median(times), times


def time_e2e_loop(env, num_envs, steps, warmup, dist, device_index):
    """Public API with HOST buffers: per step H2D(actions) → env.step → ONE D2H of the packed [obs | reward | reset flags] block
    (env.results_block; reset flags travel as bytes, reset_buf stays int64 on the device); copies in the timed region.  The pinned
    host buffers are placed on the GPU's NUMA node (agx_host_alloc_pinned)."""
    import torch

    from airgym_b200 import _capi

    a_bytes = num_envs * 4 * 4
    a_pin, info_a = _capi.pinned_host_tensor(a_bytes, device_index)
    a_host = a_pin.view(torch.float32).view(num_envs, 4)
    a_host.copy_(torch.rand(num_envs, 4) * 2 - 1)
    blk_h, info_r = _capi.pinned_host_tensor(env.results_block.numel(), device_index)
    obs_h, rew_h, rst_h = env.unpack_results(blk_h)
    a_dev = [torch.empty(num_envs, 4, device="cuda") for _ in range(2)]  # double-buffered: the H2D of step t+1 rides under the D2H of step t
    h2d, d2h = a_bytes, env.results_block.numel()
    main = torch.cuda.current_stream()
    copy_in = torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_step = [torch.cuda.Event() for _ in range(2)]
    state = {"t": 0}

    def stage(t):  # host -> device copy of step t's actions, on the copy stream (PCIe is full duplex)
        with torch.cuda.stream(copy_in):
            copy_in.wait_event(ev_step[t % 2])  # the step that last read this buffer has run
            a_dev[t % 2].copy_(a_host, non_blocking=True)
            ev_in[t % 2].record(copy_in)

    def one():
        t = state["t"]
        main.wait_event(ev_in[t % 2])
        env.step(a_dev[t % 2])
        ev_step[t % 2].record(main)
        stage(t + 1)
        blk_h.copy_(env.results_block, non_blocking=True)
        state["t"] = t + 1

    for e in ev_step:
        e.record(main)
    stage(0)
    for _ in range(max(3, min(warmup, 20))):
        one()
    torch.cuda.synchronize()

    def block():
        for _ in range(steps):
            one()

    def pre_roll():
        for _ in range(3):
            one()

    times = timed_blocks(block, pre_roll, dist, "e2e")
    assert float(rew_h.abs().sum()) > 0 and torch.isfinite(obs_h).all() and int(rst_h.max()) <= 1
    assert torch.equal(rst_h.to(torch.int64), env.reset_buf.cpu()), "the byte flags of the block must equal reset_buf"
    return median(times), h2d, d2h, {"pinned_numa_node": info_r["numa_node"], "pinned_placed": info_r["placed"] and info_a["placed"]}


def ppo_run(task, num_envs, epochs, rank, local, world, dist, vae=False, warm=3):
    """`epochs` timed PPO epochs (after `warm` eager / capture epochs) of `task` at `num_envs` envs per GPU through the trainer's public
    classes; rollout and update replay from CUDA graphs; with more than one rank the gradient all-reduce is fused into the Adam kernel
    over NVLink peer memory.  Device-timed, max over ranks."""
    import torch

    from airgym_b200.lib.agent.a2c_continuous import A2CAgent
    from airgym_b200.lib.config import default_ppo_config, scale_minibatch
    from airgym_b200.lib.utils import tr_helpers

    cfg = scale_minibatch(default_ppo_config(task), num_envs)
    c = cfg["params"]["config"]
    c.update(multi_gpu=world > 1, print_stats=False, write_summaries=False, train_dir="/tmp/agx_bench_ppo", save_frequency=0,
             save_best_after=10**9, device=f"cuda:{local}")
    c["env_config"].update(ctl_mode="rate", num_envs=num_envs, seed=1)
    c["reward_shaper"] = tr_helpers.DefaultRewardsShaper(**c["reward_shaper"])
    if vae:  # ppo_planning.yaml:33-39 — frozen VAE encoder (latent 64); trained/vae_model.pth does not travel: random frozen weights
        cfg["params"]["network"].pop("cnn", None)
        cfg["params"]["network"]["vae"] = {"latent_dims": 64, "image_res": [120, 212], "interpolation_mode": "bilinear",
                                           "return_sampled_latent": False, "allow_random_init": True}
    torch.manual_seed(1)
    agent = A2CAgent("bench", cfg["params"])
    agent.env_reset()
    agent.sync_replicas()
    for _ in range(warm):  # eager warm-up, graph capture, first replay
        agent.train_epoch()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    play = upd = 0.0
    for _ in range(epochs):
        p, u, _ = agent.train_epoch()
        play, upd = play + p, upd + u
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), 1e3 * play / epochs, 1e3 * upd / epochs], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_play, ms_upd = t.tolist()
    allreduce_us = None
    if agent.comm is not None and task == "hovering":  # the stand-alone collective on a gradient-sized message, back to back (eager launches)
        buf = torch.zeros_like(agent.flat_grads)
        for _ in range(20):
            agent.comm.all_reduce(buf)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(200):
            agent.comm.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        allreduce_us = 1e3 * a0.elapsed_time(a1) / 200
        agent.comm.check()
    n_mb = agent.num_minibatches * agent.mini_epochs_num
    rec = {"samples_per_s": world * agent.batch_size * epochs / (ms * 1e-3), "env_steps_per_s_rollout": world * agent.batch_size / (ms_play * 1e-3),
           "ms_per_epoch": ms / epochs, "ms_rollout": ms_play, "ms_update": ms_upd,
           "allreduce_us": allreduce_us, "epochs_timed": epochs, "num_envs_per_gpu": num_envs, "horizon": agent.horizon_length,
           "minibatch_per_gpu": agent.minibatch_size, "minibatches_per_epoch": n_mb,
           "rollout": "fused: policy step (tcgen05 MLP + sampling + buffer writes) + env step + post kernel" if agent.fused_rollout else "per-op",
           "mlp_backward": "tcgen05" if getattr(agent, "mlp_train_tc", False) else "mma.sync",
           "collective": ("none (1 rank)" if world == 1 else
                          ("peer-memory all-reduce fused into the Adam kernel (agx_adam_step_allreduce) + agx_comm_allreduce for the "
                           "per-epoch moments, all inside the CUDA graphs" if agent.comm is not None else "NCCL all-reduce")),
           "kl_last": float(agent.epoch_loss_sums[4]) / n_mb}
    if agent.has_cnn:
        rec["encoder"] = "frozen VAE ImgEncoder (random weights), libagx tcgen05 layers" if vae else "CNNFeatureExtractor, libagx"
        rec["render_every"] = agent.env.cam_every
    if agent.comm is not None:
        agent.comm.close()
    del agent
    torch.cuda.empty_cache()
    return rec


def ppo_record(args, rank, local, world, dist):
    """PPO samples/s on the same envs (north_star: "PPO tokens/sec scaling 1->8"): Hovering/CTBR, 65 536 envs per GPU, the reference's
    ppo_hovering.yaml hyper-parameters with the minibatch scaled to keep its 48 minibatches per mini-epoch."""
    return ppo_run("hovering", args.ppo_envs, args.ppo_epochs, rank, local, world, dist)


def camera_records(args, rank, local, world, dist):
    """BASELINE configs 4 / 5 at their per-GPU share (32 768 envs per GPU; 4 GPUs = 131 072 Avoid envs, 8 GPUs = 262 144 Planning
    envs): depth camera every 4th step, CNN(30) policy (ppo_avoid.yaml / ppo_planning.yaml) and, for Planning, the VAE variant.
    Avoid runs in rate mode: the reference itself cannot run Avoid in atti (CTA) mode (avoid.py:226 writes [N,5] into obs[12:16])."""
    out = {}
    for key, task, vae, epochs in (("avoid_cnn", "avoid", False, 2), ("planning_cnn", "planning", False, 3), ("planning_vae", "planning", True, 3)):
        try:
            out[key] = ppo_run(task, args.camera_envs, epochs, rank, local, world, dist, vae=vae, warm=3)
        except Exception as exc:  # noqa: BLE001 — side records must never cost the bench line
            out[key] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    return out


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
        graft.build_product()  # the checkers (tests/hostsim, oracle/) are not built or loaded by the GPU arm
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
    ms, blocks = time_kernel_loop(envs, acts, K, W, dist)
    e2e_steps = max(20, min(K, 300))
    if args.no_e2e:
        ms_e2e, h2d, d2h, e2e_info = float("nan"), 0, 0, {}
    else:
        ms_e2e, h2d, d2h, e2e_info = time_e2e_loop(envs[0], N, e2e_steps, W, dist, local)
    clocks = sampler.stop() if rank == 0 else None

    extras = {}
    if not args.no_extras:
        # (1) the full reference contract of step(): extras["item_reward_info"] planes + cmd_thrusts + quirk Q4's in-place remap
        fenvs, facts = make_envs(N, REPLICAS, rank, world, device, full_contract=True)
        fresh = [a.clone() for a in facts]
        fms, _ = time_kernel_loop(fenvs, facts, K, W, dist, fresh=fresh)
        extras["full_contract"] = {"us_per_step": 1e3 * fms / K, "includes": "reward_terms [12,N] + cmd_thrusts [N,4] outputs and the in-place "
                                   "action remap (Q4); a fresh 1 MB action tensor is copied in before every step (a second launch inside the timed "
                                   "region), as a policy writing new actions would — Q4 would otherwise drift the re-fed tensor"}
        del fenvs, facts, fresh
        # (2) the bandwidth-bound regime of the same kernel: 2^22 envs per launch
        n_big = 1 << 22
        benvs, bacts = make_envs(n_big, 2, rank, world, device)
        bms, _ = time_kernel_loop(benvs, bacts, 128, 16, dist)
        extras["n_2pow22"] = {"us_per_step": 1e3 * bms / 128, "algo_GBps": ALGO_BYTES_PER_ENV_STEP * n_big / ((bms / 128) * 1e-3) / 1e9,
                              "env_steps_per_s_per_gpu": n_big * 128 / (bms * 1e-3)}
        del benvs, bacts
        torch.cuda.empty_cache()
    if args.sweep and rank == 0 and world == 1:
        for logn in (16, 18, 20, 22):
            n = 1 << logn
            reps = max(2, (REPLICAS << 16) // n)
            ev, ac = make_envs(n, reps, 0, 1, device)
            m, _ = time_kernel_loop(ev, ac, 384, 50, None)
            gbs = ALGO_BYTES_PER_ENV_STEP * n * 384 / (m * 1e-3) / 1e9
            print(f"[sweep] N=2^{logn} replicas={reps} us/step={1e3 * m / 384:.2f} env-steps/s={n * 384 / (m * 1e-3):.3e} algo GB/s={gbs:.0f}",
                  file=sys.stderr, flush=True)
            del ev, ac
            torch.cuda.empty_cache()
    ppo = None
    if not args.no_ppo:
        del envs, acts
        torch.cuda.empty_cache()
        ppo = ppo_record(args, rank, local, world, dist)
    camera = None
    if not args.no_camera:
        camera = camera_records(args, rank, local, world, dist)

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
        "config": bench_config(N, world, opts),
        "blocks": {"n": len(blocks), "ms_median": ms, "ms_min": min(blocks), "ms_max": max(blocks)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if N == NUM_ENVS else None, "traffic_unit": "B/launch (ncu dram read+write; algorithmic: "
                     f"{ALGO_BYTES_PER_ENV_STEP * N} B/launch)", "peak_source": peak_src,
                     "note": f"{ALGO_BYTES_PER_ENV_STEP} algorithmic B/env-step x {N} envs / (median K-step block / K), i.e. launch gaps included"},
        "e2e": {"value": world * N * e2e_steps / (ms_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "path": "env.step() via ctypes C ABI; pinned host actions in (copy stream, overlapping the previous step's read-back); "
                        "obs + reward + reset flags out to pinned host as ONE copy of the env's packed results block (reset flags as "
                        "bytes; reset_buf stays int64 on the device); host buffers placed on the GPU's NUMA node", **e2e_info},
        "gpu_launches": K,
        "clocks": clocks,
    }
    line.update(extras)
    if ppo is not None:
        line["ppo"] = ppo
    if camera is not None:
        line["camera_tasks"] = camera
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
