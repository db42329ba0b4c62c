"""AvoidCfg — the reference's config (airgym/envs/task/avoid_config.py:7-97): 6 s episodes, reset on collision, depth camera
212x120 every 0.04 s, one thrown 1x1 cube, hover target at z = 1."""
import numpy as np

from ..base.base_config import BaseConfig
from ..base.hovering_config import HoveringCfg

CAMERA_ROBOT = dict(HoveringCfg.asset_config.include_robot["X152b"], enable_onboard_cameras=True, cam_channel=1, enable_tensors=True,
                    width=212, height=120, far_plane=5.0, horizontal_fov=87.0, use_collision_geometry=True,
                    **{"local_transform.p": (0.15, 0.00, 0.1), "local_transform.r": (0.0, 0.0, 0.0, 1.0)}, collision_mask=1)


class AvoidCfg(BaseConfig):
    seed = -1

    class env:
        target_state = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0])
        num_envs = 4
        num_observations = 16
        headless = True
        get_privileged_obs = True
        env_spacing = 4
        episode_length_s = 6
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
        include_robot = {"X152b": CAMERA_ROBOT}
        include_single_asset = {
            "cubes/1x1": {"collision_mask": 0, "num_assets": 1, "density": 0.5, "fix_base_link": False},
            "balls/ball": {"disable_gravity": False, "color": [255, 102, 102], "collision_mask": 0, "num_assets": 0, "density": 1,
                           "fix_base_link": False},
        }
        include_group_asset = {}
        include_boundary = {}

    backend = HoveringCfg.backend
