"""CLI flags and cfg overrides — same flag names/defaults as the reference (airgym/utils/helpers.py:64-116 and
the vendored gymutil parser airgym/utils/gym_utils/gymutil.py:298-366)."""
import argparse


def class_to_dict(obj) -> dict:
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        out[key] = [class_to_dict(v) for v in val] if isinstance(val, list) else class_to_dict(val)
    return out


def parse_device_str(device_str: str):
    s = device_str.lower()
    if s in ("cpu", "cuda"):
        return s, 0
    kind, _, idx = s.partition(":")
    assert kind == "cuda" and idx.isdigit(), f"Invalid device string {device_str!r}"
    return kind, int(idx)


def parse_sim_params(args, cfg):
    """The reference builds a gymapi.SimParams; this backend just carries the dict through."""
    sim = dict(cfg.get("sim", {}))
    sim["use_gpu_pipeline"] = getattr(args, "use_gpu_pipeline", True)
    return sim


def update_cfg_from_args(env_cfg, args):
    if env_cfg is not None:
        if getattr(args, "num_envs", None) is not None:
            env_cfg.env.num_envs = args.num_envs
        if hasattr(args, "ctl_mode"):
            env_cfg.env.ctl_mode = args.ctl_mode
        if hasattr(args, "seed"):
            env_cfg.seed = args.seed
    return env_cfg


def build_parser():
    p = argparse.ArgumentParser(description="RL Policy")
    p.add_argument("--sim_device", type=str, default="cuda:0", help="Physics Device in PyTorch-like syntax")
    p.add_argument("--pipeline", type=str, default="gpu", help="Tensor API pipeline (cpu/gpu)")
    p.add_argument("--graphics_device_id", type=int, default=0)
    g = p.add_mutually_exclusive_group()
    g.add_argument("--flex", action="store_true")
    g.add_argument("--physx", action="store_true")
    p.add_argument("--num_threads", type=int, default=0)
    p.add_argument("--subscenes", type=int, default=0)
    p.add_argument("--slices", type=int)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--tf", action="store_true")
    p.add_argument("--train", action="store_true")
    p.add_argument("--play", action="store_true")
    p.add_argument("--checkpoint", type=str)
    p.add_argument("--num_envs", type=int, default=4096)
    p.add_argument("--sigma", type=float)
    p.add_argument("--track", action="store_true")
    p.add_argument("--wandb-project-name", type=str, default="rl_games")
    p.add_argument("--wandb-entity", type=str, default=None)
    p.add_argument("--task", type=str, default=None)
    p.add_argument("--experiment_name", type=str)
    p.add_argument("--headless", action="store_true", default=False)
    p.add_argument("--horovod", action="store_true", default=False)
    p.add_argument("--rl_device", type=str, default="cuda:0")
    p.add_argument("--ctl_mode", required=True, type=str, help="pos, vel, atti, rate, prop")
    # additions of this backend (not in the reference)
    p.add_argument("--config", type=str, default=None, help="PPO yaml, e.g. the reference's scripts/config/ppo_hovering.yaml")
    p.add_argument("--max_epochs", type=int, default=None)
    return p


def get_args(argv=None):
    args = build_parser().parse_args(argv)
    args.sim_device_type, args.compute_device_id = parse_device_str(args.sim_device)
    args.use_gpu_pipeline = args.pipeline.lower() in ("gpu", "cuda")
    args.physics_engine = None
    args.use_gpu = False
    if args.slices is None:
        args.slices = args.subscenes
    args.sim_device_id = args.compute_device_id
    args.sim_device = args.sim_device_type
    if args.sim_device == "cuda":
        args.sim_device += f":{args.sim_device_id}"
    return args
