"""Runner — the PPO entry point with the reference's interface (lib/torch_runner.py:17-100):
`Runner(algo_observer).load(yaml_config); runner.run(args)`; seeds per rank, builds the reward shaper, runs train or play."""
import os
import random
import time
from copy import deepcopy

import numpy as np
import torch

from .agent import a2c_continuous
from .utils import tr_helpers


def _restore(agent, args):
    if args.get("checkpoint"):
        agent.restore(args["checkpoint"])


class Runner:
    def __init__(self, algo_observer=None):
        self.algo_observer = algo_observer

    def reset(self):
        pass

    def load_config(self, params):
        self.seed = params.get("seed", None)
        if self.seed is None:
            self.seed = int(time.time())
        self.local_rank = self.global_rank = 0
        self.world_size = 1
        if params["config"].get("multi_gpu", False):
            self.local_rank = int(os.getenv("LOCAL_RANK", "0"))
            self.global_rank = int(os.getenv("RANK", "0"))
            self.world_size = int(os.getenv("WORLD_SIZE", "1"))
            self.seed += self.global_rank  # torch_runner.py:43-44
        self.algo_params = params["algo"]
        self.algo_name = self.algo_params["name"]
        if self.seed:
            torch.manual_seed(self.seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed_all(self.seed)
            np.random.seed(self.seed)
            random.seed(self.seed)
            if "env_config" in params["config"] and "seed" not in params["config"]["env_config"]:
                params["config"]["env_config"]["seed"] = self.seed
        config = params["config"]
        if isinstance(config.get("reward_shaper"), dict):
            config["reward_shaper"] = tr_helpers.DefaultRewardsShaper(**config["reward_shaper"])
        config.setdefault("features", {})["observer"] = self.algo_observer
        self.params = params

    def load(self, yaml_config):
        config = deepcopy(yaml_config)
        self.default_config = deepcopy(config["params"])
        self.load_config(params=self.default_config)

    def run_train(self, args):
        print("Started to train")
        agent = a2c_continuous.A2CAgent(base_name="run", params=self.params)
        _restore(agent, args)
        self.agent = agent
        return agent.train()

    def run_play(self, args):
        print("Started to play")  # torch_runner.py:86-90
        from .agent.players import PpoPlayerContinuous

        player = PpoPlayerContinuous(self.params)
        if args.get("checkpoint"):
            player.restore(args["checkpoint"])
        self.player = player
        return player.run()

    def run(self, args):
        if args.get("play") and not args.get("train"):
            return self.run_play(args)
        return self.run_train(args)
