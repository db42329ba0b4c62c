"""Planning — host-side mirror of airgym/envs/task/planning.py: fly from x = -8.5 to a goal ball at x = +8.5 through 40 randomly
scattered thin trees, 16-dim yaw-aligned observation + depth image.  aux columns: 0:3 goal position, 3:6 `pre_root_positions`,
6 `collisions`, 7 `esdf_dist` (min over the current image, planning.py:162-163).  `assets` [N,164] holds the scatter:
x[41] | y[41] | cos yaw[41] | sin yaw[41] (asset 0 = the goal ball, 1..40 = trees)."""
import torch

from ... import _capi
from ..base.customized import Customized, load_tree_table


class Planning(Customized):
    TASK = "planning"
    REWARD_KEYS = ("continous_action_reward", "heading_reward", "speed_reward", "forward_reward", "alive_reward", "ups_reward",
                   "z_reward", "esdf_reward", "thrust_reward", "reach_goal_reward", "reward")  # planning.py:293-305

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        N, dev = self.num_envs, self._dev
        self.num_assets = _capi.AGX_NUM_ASSETS
        self.assets = torch.zeros(N, _capi.AGX_ASSET_ROW, device=dev, dtype=torch.float32)
        self.assets[:, 2 * _capi.AGX_NUM_ASSETS:3 * _capi.AGX_NUM_ASSETS] = 1.0  # identity yaw
        self.trees = load_tree_table(dev)
        self.goal_positions = self.aux[:, 0:3]
        self.pre_root_positions = self.aux[:, 3:6]
        self.esdf_dist = self.aux[:, 7]
        self._io.assets = self.assets.data_ptr()
        self._io.trees = self.trees.data_ptr()
        self._rio.assets = self.assets.data_ptr()
        self._rio.trees = self.trees.data_ptr()

    def _assets_ptr(self):
        return self.assets.data_ptr()

    @property
    def env_asset_root_states(self):
        """[N,41,13] asset root states (customized.py:75-79), rebuilt from the scatter table."""
        A = _capi.AGX_NUM_ASSETS
        s = torch.zeros(self.num_envs, A, 13, device=self._dev)
        s[:, :, 0], s[:, :, 1] = self.assets[:, 0:A], self.assets[:, A:2 * A]
        half = 0.5 * torch.atan2(self.assets[:, 3 * A:4 * A], self.assets[:, 2 * A:3 * A])
        s[:, :, 5], s[:, :, 6] = torch.sin(half), torch.cos(half)
        s[:, 0, 0:3] = self.goal_positions
        return s

    @property
    def privileged_obs_buf(self):
        return self.env_asset_root_states if self.get_privileged_obs else None

    @privileged_obs_buf.setter
    def privileged_obs_buf(self, value):
        pass
