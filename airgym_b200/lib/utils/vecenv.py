"""rl_games-style vec-env adapter (reference lib/utils/vecenv.py:50-119, lib/utils/env_configurations.py):
`create_vec_env(name, num_actors, **env_config)` → object with step / reset / get_env_info.  `gym` is not a dependency:
spaces are a two-field Box."""
from argparse import Namespace
from dataclasses import dataclass

import numpy as np

from ...envs import task_registry  # registers the tasks on import


@dataclass
class Box:
    low: np.ndarray
    high: np.ndarray

    @property
    def shape(self):
        return self.low.shape


class AirGymRLGPUEnv:
    def __init__(self, config_name, num_actors, **kwargs):
        self.use_image = kwargs.get("use_image", False)
        kwargs.setdefault("num_envs", num_actors)
        kwargs.setdefault("headless", True)
        kwargs.setdefault("physics_engine", None)
        kwargs.setdefault("sim_device", "cuda:0")
        kwargs.setdefault("seed", 0)
        self.env, self.env_cfg = task_registry.make_env(config_name, args=Namespace(**kwargs))

    def _obs(self, obs):
        # the camera tasks always return the dict; without use_image the policy only sees the vector part
        return obs["observation"] if (isinstance(obs, dict) and not self.use_image) else obs

    def step(self, actions):  # ExtractObsWrapper.step: drop the privileged observations (vecenv.py:59-67)
        obs, _priv, rewards, dones, infos = self.env.step(actions)
        return self._obs(obs), rewards, dones, infos

    def reset(self):
        obs, _priv = self.env.reset()
        return self._obs(obs)

    def get_number_of_agents(self):
        return 1

    def get_env_info(self):
        info = {k: v for k, v in type(self.env_cfg.env).__dict__.items() if not k.startswith("__") and not callable(v)}
        n = self.env.num_actions
        info["action_space"] = Box(-np.ones(n, np.float32), np.ones(n, np.float32))
        vec = Box(np.full(self.env.num_obs, -np.inf, np.float32), np.full(self.env.num_obs, np.inf, np.float32))
        if self.use_image:  # vecenv.py:93-98: Dict{'image': Box(0,1,(C,W,H)), 'observation': Box}
            shp = (self.env.cam_channel, self.env.cam_resolution[0], self.env.cam_resolution[1])
            info["observation_space"] = {"image": Box(np.zeros(shp, np.float32), np.ones(shp, np.float32)), "observation": vec}
        else:
            info["observation_space"] = vec
        return info

    def get_env_state(self):
        return None  # the reference never checkpoints env state (lib/utils/ivecenv.py:28-35)

    def set_env_state(self, env_state):
        pass


def create_vec_env(config_name, num_actors, **kwargs):
    if config_name not in task_registry.get_registered_tasks():
        raise ValueError(f"unknown env_name {config_name!r}; registered: {task_registry.get_registered_tasks()}")
    return AirGymRLGPUEnv(config_name, num_actors, **kwargs)
