"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the reference's Customized task family — ``airgym/envs/base/customized.py:216-477`` (pre_physics_step
:216-295, step :308-344, check_collisions :393-397) — and of Balloon (``airgym/envs/task/balloon.py:57-225``).
Pinned, like oracle/hovering.py, by running the reference's own classes with the absent dependencies stubbed
(tests/golden/make_golden.py).  Builder-defined and PARITY UNPINNED: the contact model behind `check_collisions`
(PhysX net contact force in the reference) — here the drone's r=0.2 collision sphere (robots/X152b/model.urdf:13-18)
against the ground plane; assets that share the drone's collision mask (the balloon) never collide (IsaacGym filter rule).
"""
import math

import torch
import torch.nn.functional as F

from . import rotations as R
from .hovering import HoveringOracle
from .rigid_body import simulate
from .spec import QuadSpec


class CustomizedOracle(HoveringOracle):
    AUX = 8
    reset_on_collision = False

    def __init__(self, spec: QuadSpec, num_envs: int, dtype=torch.float32, rng: str = "torch"):
        super().__init__(spec, num_envs, dtype, rng)
        self.collisions = torch.zeros(num_envs, dtype=dtype)
        self.contact_forces = torch.zeros(num_envs, 3, dtype=dtype)
        self.counter = 0

    # builder-defined stand-in for PhysX's net contact force on the drone body
    def refresh_contact_forces(self):
        hit = self.root_states[:, 2] < 0.2
        self.contact_forces.zero_()
        self.contact_forces[hit, 2] = 1.0

    def check_collisions(self):  # customized.py:393-397
        ones, zeros = torch.ones(self.num_envs, dtype=self.dtype), torch.zeros(self.num_envs, dtype=self.dtype)
        self.collisions = torch.where(torch.norm(self.contact_forces, dim=-1) > 0.1, ones, zeros)

    def pre_physics_step(self, _actions):  # customized.py:216-295
        self.counter += 1
        reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_env_ids) > 0:
            self.reset_idx(reset_env_ids, which=0)
        self.actions = _actions
        actions = self.actions
        if self.ctl_mode in ("rate", "atti"):
            actions[..., -1] = 0.5 + 0.5 * self.actions[..., -1]  # in place: self.actions sees the remap, not the clamp
        actions = R.tensor_clamp(actions, self.action_lower_limits, self.action_upper_limits)
        self.root_states[..., 3:7] = torch.where(self.root_states[..., 6:7] < 0, -self.root_states[..., 3:7],
                                                 self.root_states[..., 3:7])
        self.pre_step_quat = self.root_states[:, 3:7].clone()
        pos, quat = self.root_states[:, 0:3].clone(), self.root_states[:, 3:7].clone()
        linvel, angvel = self.root_states[:, 7:10].clone(), self.root_states[:, 10:13].clone()
        q_wxyz = quat[:, [3, 0, 1, 2]]
        if self.ctl_mode in ("pos", "vel", "atti"):
            self.controller.set_status(pos, q_wxyz, linvel, angvel, 0.01)
            self.cmd_thrusts = self.controller.update(actions)
        elif self.ctl_mode == "rate":
            self.controller.set_q_world(q_wxyz)
            self.cmd_thrusts = self.controller.update(actions, angvel, 0.01)
        else:
            self.cmd_thrusts = actions
        if self.rng == "torch":
            torch.rand(self.num_envs, 1)  # `delta = .0*torch_rand_float(...)` consumes a draw (customized.py:267)
        thrusts = (self.cmd_thrusts * self.spec.k_thrust).clone()
        thrusts[reset_env_ids] = 0
        prop_rot = self.cmd_thrusts * self.spec.k_torque
        return thrusts, -prop_rot[:, 0] - prop_rot[:, 1] + prop_rot[:, 2] + prop_rot[:, 3]

    def step(self, actions, rand_reset=None, rand_noise=None):  # customized.py:308-344 / balloon.py:95-130
        N = self.num_envs
        if self.rng == "explicit":
            self._explicit = {"reset": rand_reset.reshape(N, 2, self.RESET_DRAWS), "noise": rand_noise}
        self.last_draws = {"reset": torch.zeros(N, 2, self.RESET_DRAWS), "noise": torch.zeros(N, 18)}
        thrusts, tau_z = self.pre_physics_step(actions)
        simulate(self.spec, self.root_states, thrusts, tau_z)
        self.refresh_contact_forces()
        self.progress_buf += 1
        self.check_collisions()
        self.compute_observations()
        self.compute_reward()
        if self.reset_on_collision:
            self.reset_buf = torch.where(self.collisions > 0, torch.ones_like(self.reset_buf), self.reset_buf)
        reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_env_ids) > 0:
            self.reset_idx(reset_env_ids, which=1)
        self.time_out_buf = self.progress_buf > self.max_episode_length
        self.extras["time_outs"] = self.time_out_buf
        self.extras["item_reward_info"] = self.item_reward_info
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras


class BalloonOracle(CustomizedOracle):
    RESET_DRAWS = 15
    reset_on_collision = True  # balloon_config.py:19
    REWARD_KEYS = ("guidance_reward", "hit_reward", "action_smoothness_reward", "effort_reward", "ups_reward", "yaw_reward",
                   "_pad0", "_pad1", "reward")

    def __init__(self, spec, num_envs, dtype=torch.float32, rng="torch"):
        super().__init__(spec, num_envs, dtype, rng)
        self.balloon_states = torch.zeros(num_envs, 13, dtype=dtype)
        self.balloon_states[:, 6] = 1.0
        self.pre_root_positions = torch.zeros(num_envs, 3, dtype=dtype)

    @property
    def balloon_positions(self): return self.balloon_states[:, 0:3]
    @property
    def balloon_quats(self): return self.balloon_states[:, 3:7]

    def _reset_uniforms(self, env_ids, which):
        n = len(env_ids)
        if self.rng == "explicit":
            u = self._explicit["reset"][env_ids, which].to(self.dtype)
        else:  # call order of balloon.py:61-83
            u = torch.cat([torch.rand(n, k) for k in (1, 1, 1, 2, 1, 1, 1, 1, 3, 3)], -1).to(self.dtype)
        if which is not None and self.last_draws is not None:
            self.last_draws["reset"][env_ids, which] = u.to(torch.float32)
        return u

    def reset_idx(self, env_ids, which=None, u=None):  # balloon.py:57-93
        if u is None:
            u = self._reset_uniforms(env_ids, which)
        pi = math.pi
        self.balloon_states[env_ids, 0:1] = 0.5 * R.rand_float(-1.0, 1.0, u[:, 0:1]) + 2.5
        self.balloon_states[env_ids, 1:2] = 2.0 * R.rand_float(-1.0, 1.0, u[:, 1:2]) + 0.0
        self.balloon_states[env_ids, 2:3] = 0.3 * R.rand_float(-1.0, 1.0, u[:, 2:3]) + 1.0
        self.root_states[env_ids, 0:2] = 0.1 * R.rand_float(-1.0, 1.0, u[:, 3:5]) + 0.0
        self.root_states[env_ids, 2:3] = 0.2 * R.rand_float(-1.0, 1.0, u[:, 5:6]) + 1.0
        ang = torch.cat((0.1 * R.rand_float(-pi, pi, u[:, 6:7]), 0.1 * R.rand_float(0.0, pi, u[:, 7:8]),
                         0.2 * R.rand_float(-pi, pi, u[:, 8:9])), -1)
        quat = R.matrix_to_quaternion(R.euler_angles_to_matrix(ang, "XYZ"))
        self.root_states[env_ids, 3:7] = quat[:, [1, 2, 3, 0]]
        self.root_states[env_ids, 7:10] = 0.5 * R.rand_float(-1.0, 1.0, u[:, 9:12])
        self.root_states[env_ids, 10:13] = 0.2 * R.rand_float(-1.0, 1.0, u[:, 12:15])
        self.reset_buf[env_ids] = 1
        self.progress_buf[env_ids] = 0
        self.pre_actions[env_ids] = 0
        self.pre_root_positions[env_ids] = 0
        if self.spec.ctrl_reset and self.controller is not None:
            self.controller.reset(env_ids)

    def compute_observations(self):  # balloon.py:132-145
        self._fill_base_obs()
        self.add_noise()
        balloon_matrix = R.quaternion_to_matrix(self.balloon_quats[:, [3, 0, 1, 2]]).reshape(self.num_envs, 9)
        self.obs_buf[..., 0:9] -= balloon_matrix
        self.obs_buf[..., 9:12] -= self.balloon_positions
        return self.obs_buf

    def compute_reward(self):  # balloon.py:147-153
        reward, reset, info = self.compute_quadcopter_reward()
        self.rew_buf[:] = reward
        self.reset_buf[:] = reset
        self.item_reward_info = info
        self.pre_actions = self.actions.clone()
        self.pre_root_positions = self.root_positions.clone()

    def compute_quadcopter_reward(self):  # balloon.py:159-225
        relative_positions = self.balloon_positions - self.root_positions
        direction_vector = F.normalize(relative_positions, dim=-1)
        direction_yaw = torch.atan2(direction_vector[..., 1], direction_vector[..., 0])
        root_matrix = R.quaternion_to_matrix(self.root_quats[:, [3, 0, 1, 2]])
        root_euler = R.matrix_to_euler_xyz(root_matrix)
        relative_heading = R.compute_yaw_diff(root_euler[..., 2], direction_yaw)
        yaw_distance = torch.norm(relative_heading.unsqueeze(-1), dim=1)
        yaw_reward = 1.0 / (1.0 + torch.square(1.6 * yaw_distance))
        guidance_reward = 30 * (torch.norm(self.balloon_positions - self.pre_root_positions, dim=-1)
                                - torch.norm(self.balloon_positions - self.root_positions, dim=-1))
        ups = R.quat_axis(self.root_quats, axis=2)
        ups_reward = 0.5 * torch.pow((ups[..., 2] + 1) / 2, 2)
        check = torch.norm(self.balloon_positions - self.root_positions, dim=-1)
        hit_reward = 800 * torch.where(check < 0.1, torch.tensor(1), torch.tensor(0))
        effort_reward = 0.1 * torch.exp(-self.actions.pow(2).sum(-1))
        action_diff = torch.norm(self.actions - self.pre_actions, dim=-1)
        action_smoothness_reward = 0.1 * torch.exp(-action_diff)
        reward = guidance_reward + yaw_reward + hit_reward + action_smoothness_reward + ups_reward + effort_reward
        ones, die = torch.ones_like(self.reset_buf), torch.zeros_like(self.reset_buf)
        reset = torch.where(self.progress_buf >= self.max_episode_length - 1, ones, die)
        reset = torch.where(self.actions[..., -1] < -1, ones, reset)
        reset = torch.where(self.actions[..., -1] > 1, ones, reset)
        reset = torch.where(relative_positions[..., 0] < -0.2, ones, reset)
        reset = torch.where(self.root_linvels[..., 0] < 0, ones, reset)
        reset = torch.where(torch.norm(relative_positions, dim=1) > 4, ones, reset)
        reset = torch.where(self.root_positions[..., 2] < 0.5, ones, reset)
        reset = torch.where(self.root_positions[..., 2] > 1.5, ones, reset)
        reset = torch.where(check < 0.1, ones, reset)
        info = {"guidance_reward": guidance_reward, "hit_reward": hit_reward, "action_smoothness_reward": action_smoothness_reward,
                "effort_reward": effort_reward, "ups_reward": ups_reward, "yaw_reward": yaw_reward, "_pad0": 0, "_pad1": 0,
                "reward": reward}
        return reward, reset, info

    def aux_matrix(self):
        """[N,8] float32 in the layout of AgxStepIO.aux for the balloon task."""
        a = torch.zeros(self.num_envs, 8)
        a[:, 0:3] = self.balloon_positions
        a[:, 3:6] = self.pre_root_positions
        a[:, 6] = self.collisions
        return a
