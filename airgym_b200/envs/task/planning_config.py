"""PlanningCfg — the reference's config (airgym/envs/task/planning_config.py:7-85): 16 s episodes, no reset on collision (the
reward function resets), depth camera 212x120 every 0.04 s, goal ball + 40 `thin` trees."""
import numpy as np

from ..base.base_config import BaseConfig
from ..base.hovering_config import HoveringCfg
from .avoid_config import CAMERA_ROBOT


class PlanningCfg(BaseConfig):
    seed = -1

    class env:
        target_state = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0])
        num_envs = 4
        num_observations = 16
        headless = True
        get_privileged_obs = True
        env_spacing = 14
        episode_length_s = 16
        num_control_steps_per_env_step = 1
        reset_on_collision = False
        create_ground_plane = True
        cam_dt = 0.04

    viewer = HoveringCfg.viewer

    class sim:
        dt = 0.01
        substeps = 1
        gravity = [0.0, 0.0, -9.81]
        up_axis = 1
        physx = HoveringCfg.sim.physx

    class asset_config:
        include_robot = {"X152b": CAMERA_ROBOT}
        include_single_asset = {"balls/ball": {"color": [255, 102, 102], "num_assets": 1}}
        include_group_asset = {"thin": {"num_assets": 40, "collision_mask": 1, "color": [139, 69, 0]}}
        include_boundary = {}

    backend = HoveringCfg.backend
