"""Avoid — host-side mirror of airgym/envs/task/avoid.py: hover at (0,0,1) while a cube is thrown at the drone (80 % of episodes),
16-dim yaw-aligned observation + depth image.  aux columns: 0:3 cube position (`object_positions`), 3:6 cube linear velocity
(`object_linvels`), 6 `collisions`."""
import torch

from ..base.customized import Customized


class Avoid(Customized):
    TASK = "avoid"
    REWARD_KEYS = ("pose_reward", "ups_reward", "spin_reward", "effort_reward", "action_smoothness_reward", "thrust_reward",
                   "alive_reward", "_pad0", "reward")  # avoid.py:282-291

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.object_positions = self.aux[:, 0:3]
        self.object_linvels = self.aux[:, 3:6]

    @property
    def object_states(self):
        """[N,13] root state of the cube actor (avoid.py:39-44); the stand-in cube does not rotate."""
        s = torch.zeros(self.num_envs, 13, device=self._dev)
        s[:, 0:3], s[:, 6], s[:, 7:10] = self.object_positions, 1.0, self.object_linvels
        return s

    @property
    def privileged_obs_buf(self):  # customized.py:78-79: the asset root states
        return self.object_states.unsqueeze(1) if self.get_privileged_obs else None

    @privileged_obs_buf.setter
    def privileged_obs_buf(self, value):
        pass
