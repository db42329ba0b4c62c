#!/usr/bin/env python
"""Train entry point, same command line as the reference (scripts/runner.py:47-70):

    python scripts/runner.py --task hovering --ctl_mode rate --headless [--num_envs N] [--seed S] [--config file.yaml]

The reference's own yaml files (scripts/config/ppo_*.yaml) can be passed with --config; without it the built-in
defaults with the same values are used."""
import os
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from airgym_b200.lib.config import default_ppo_config, scale_minibatch  # noqa: E402
from airgym_b200.lib.torch_runner import Runner  # noqa: E402
from airgym_b200.utils.helpers import get_args  # noqa: E402


def update_config(config, args):
    """scripts/runner.py:19-44 of the reference"""
    c = config["params"]["config"]
    if args["task"] is not None:
        c["env_name"] = args["task"]
    if args.get("experiment_name"):
        c["name"] = args["experiment_name"]
    for k in ("physics_engine", "sim_device", "headless", "use_gpu", "subscenes", "use_gpu_pipeline", "num_threads", "ctl_mode"):
        c["env_config"][k] = args[k]
    if args["num_envs"] > 0:
        scale_minibatch(config, args["num_envs"])
        c["env_config"]["num_envs"] = args["num_envs"]
    if args["seed"] > 0:
        config["params"]["seed"] = args["seed"]
        c["env_config"]["seed"] = args["seed"]
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        c["multi_gpu"] = True
    return config


if __name__ == "__main__":
    args = vars(get_args())
    args["task"] = args["task"] or "hovering"
    if args["config"]:
        with open(args["config"]) as f:
            config = yaml.safe_load(f)
    else:
        config = default_ppo_config(args["task"])
    config = update_config(config, args)
    if args["max_epochs"] is not None:
        config["params"]["config"]["max_epochs"] = args["max_epochs"]
    runner = Runner()
    runner.load(config)
    runner.run(args)
