"""BalloonCfg — the reference's config (airgym/envs/task/balloon_config.py:7-74): 8 s episodes, reset on collision, one
`balls/ball` asset, no onboard camera."""
import numpy as np

from ..base.base_config import BaseConfig
from ..base.hovering_config import HoveringCfg


class BalloonCfg(BaseConfig):
    seed = -1

    class env:
        target_state = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0])
        num_envs = 4
        num_observations = 18
        headless = True
        get_privileged_obs = True
        env_spacing = 10
        episode_length_s = 8
        num_control_steps_per_env_step = 1
        reset_on_collision = True
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
        include_robot = {"X152b": dict(HoveringCfg.asset_config.include_robot["X152b"], enable_tensors=True)}
        include_single_asset = {"balls/ball": {"color": [255, 102, 102], "num_assets": 1}}
        include_group_asset = {}
        include_boundary = {}

    backend = HoveringCfg.backend
