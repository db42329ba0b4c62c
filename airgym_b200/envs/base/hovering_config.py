"""HoveringCfg — attribute-for-attribute the reference's config (airgym/envs/base/hovering_config.py:8-69).
`sim.physx` and `asset_config` are kept so code reading them keeps working; this backend only consumes
env.*, sim.dt and sim.gravity (there is no PhysX: the integrator lives in the fused CUDA step)."""
import numpy as np

from .base_config import BaseConfig


class HoveringCfg(BaseConfig):
    seed = -1

    class env:
        target_state = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0])
        num_envs = 256
        num_observations = 18
        get_privileged_obs = True
        env_spacing = 1
        episode_length_s = 24
        num_control_steps_per_env_step = 1
        reset_on_collision = False
        create_ground_plane = False

    class viewer:
        ref_env = 0
        pos = [-5, -5, 4]
        lookat = [0, 0, 0]

    class sim:
        dt = 0.01
        substeps = 1
        gravity = [0.0, 0.0, -9.81]
        up_axis = 1

        class physx:
            num_threads = 10
            solver_type = 1
            num_position_iterations = 4
            num_velocity_iterations = 0
            contact_offset = 0.01
            rest_offset = 0.0
            bounce_threshold_velocity = 0.5
            max_depenetration_velocity = 1.0
            max_gpu_contact_pairs = 2**23
            default_buffer_size_multiplier = 5
            contact_collection = 0

    class asset_config:
        include_robot = {
            "X152b": {
                "num_assets": 1, "enable_onboard_cameras": False, "cam_channel": 1, "enable_tensors": False,
                "width": 212, "height": 120, "far_plane": 5.0, "horizontal_fov": 87.0,
                "use_collision_geometry": True, "local_transform.p": (0.15, 0.00, 0.1),
                "local_transform.r": (0.0, 0.0, 0.0, 1.0), "collision_mask": 1,
            }
        }
        include_single_asset = {}
        include_group_asset = {}
        include_boundary = {}

    class backend:
        """B200 backend knobs (not in the reference)."""
        integrator = "rk4"          # "rk4" | "euler"
        ctrl_reset_on_reset = False  # reference behaviour: controller integrators survive episode resets
        mutate_input_actions = True  # reference quirk Q4 (hovering.py:212-215)
        reward_terms = True          # fill extras["item_reward_info"]
        export_cmd_thrusts = True    # keep the env attribute `cmd_thrusts` up to date (hovering.py:90,238-252)
