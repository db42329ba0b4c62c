#!/usr/bin/env python
"""PPO throughput / learning-curve probe:  python scripts/ppo_bench.py --task hovering --ctl_mode rate --num_envs 65536 --epochs 20
Prints one JSON line (rank 0) with samples/s for rollout ("step+inference"), update, and total, plus the reward curve.
Under torchrun the envs are sharded and gradients all-reduced (multi_gpu)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from airgym_b200.lib.config import default_ppo_config, scale_minibatch  # noqa: E402
from airgym_b200.lib.torch_runner import Runner  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="hovering")
    ap.add_argument("--ctl_mode", default="rate")
    ap.add_argument("--num_envs", type=int, default=65536)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--no_graph", action="store_true")
    ap.add_argument("--graph_collectives", action="store_true", help="multi-GPU: replay the update from graphs with NCCL inside")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--skip", type=int, default=3, help="epochs excluded from the timing (graph capture / warm-up)")
    ap.add_argument("--vae", action="store_true", help="planning: the frozen depth-VAE encoder (latent 64) instead of the CNN (ppo_planning.yaml:33-39); "
                    "random frozen weights — trained/vae_model.pth does not travel to the GPU box")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--encoder_tf32", action="store_true", help="camera tasks: single-pass TF32 encoder convolutions (cuDNN's default precision) instead of 3xTF32")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = scale_minibatch(default_ppo_config(a.task), a.num_envs)
    c = cfg["params"]["config"]
    c.update(max_epochs=a.epochs, use_cuda_graph=not a.no_graph, print_stats=False, save_frequency=0, save_best_after=10**9,
             train_dir="/tmp/agx_runs", multi_gpu=world > 1, write_summaries=False,
             graph_collectives=a.graph_collectives)
    c["env_config"].update(ctl_mode=a.ctl_mode, num_envs=a.num_envs, seed=a.seed)
    c["multi_gpu_comm"] = a.comm
    if a.vae:
        cfg["params"]["network"].pop("cnn", None)
        cfg["params"]["network"]["vae"] = {"latent_dims": 64, "image_res": [120, 212], "interpolation_mode": "bilinear",
                                           "return_sampled_latent": False, "allow_random_init": True}
    if a.encoder_tf32:
        for k in ("cnn", "vae"):
            if k in cfg["params"]["network"]:
                cfg["params"]["network"][k]["encoder_precise"] = False
    cfg["params"]["seed"] = a.seed
    import contextlib
    r = Runner()
    with contextlib.redirect_stdout(sys.stderr):  # the trainer's own prints (reference wording) must not mix with the JSON line
        r.load(cfg)
        r.run({"train": True})
    if int(os.environ.get("RANK", "0")) == 0:
        h = r.agent.history[a.skip:]
        frames = sum(x["frame"] - (r.agent.history[i + a.skip - 1]["frame"] if i + a.skip > 0 else 0) for i, x in enumerate(h))
        play, upd = sum(x["play_time"] for x in h), sum(x["update_time"] for x in h)
        print(json.dumps({
            "task": a.task, "ctl_mode": a.ctl_mode, "encoder": ("vae" if a.vae else ("cnn" if r.agent.has_cnn else None)), "encoder_precision": ("tf32" if a.encoder_tf32 else "3xtf32"),
            "env_steps_per_s_rollout": frames / play, "fused_rollout": r.agent.fused_rollout, "mlp_backward_tcgen05": getattr(r.agent, "mlp_train_tc", False), "num_envs_per_gpu": a.num_envs, "n_gpus": world, "epochs_timed": len(h),
            "minibatch": c["minibatch_size"], "cuda_graph": not a.no_graph,
            "samples_per_s_rollout": frames / play, "samples_per_s_update": frames / upd, "samples_per_s_total": frames / (play + upd),
            "ms_per_epoch_rollout": 1e3 * play / len(h), "ms_per_epoch_update": 1e3 * upd / len(h),
            "reward_curve": [None if x["mean_reward"] is None else round(x["mean_reward"], 2) for x in r.agent.history][:: max(1, a.epochs // 25)],
            "kl_last": h[-1]["kl"], "lr_last": h[-1]["lr"]}), flush=True)
