"""ORACLE (test infrastructure, never on the product path).

CPU restatement, in plain torch, of the reference's Hovering env step — ``airgym/envs/base/hovering.py:203-459``
followed method by method (pre_physics_step :203-281, step :286-308, reset_idx :310-335, compute_observations
:337-358, add_noise :349-358, compute_reward/compute_quadcopter_reward :360-459) — and of the Tracking deltas
(``airgym/envs/task/tracking.py:159-296``) in the subclass at the bottom.

Pinning: everything that is in-tree in the reference (action shaping, reset sampler, observation packing,
reward, termination, reset bookkeeping, draw order) is checked against the reference's OWN code by
`tests/golden/make_golden.py`, which imports /root/reference's Hovering/Tracking with the three absent
dependencies stubbed (isaacgym → a fake gym whose `simulate` is oracle/rigid_body.py, rlPx4Controller →
oracle/px4_controller.py, pytorch3d → oracle/rotations.py) and records seeded trajectories into
tests/golden/*.npz.  The three stubbed dependencies themselves are PARITY UNPINNED (see their module headers).

Randomness is explicit: every U[0,1)/N(0,1) number a step consumes is either drawn from torch's global CPU
generator in the reference's call order (rng="torch": seed-for-seed identical to the reference harness) and
recorded in `last_draws`, or supplied by the caller (rng="explicit") so the CUDA kernel can consume the same
numbers.
"""
import math

import torch

from . import rotations as R
from .px4_controller import ParallelControl
from .rigid_body import simulate
from .spec import QuadSpec


class HoveringOracle:
    RESET_DRAWS = 12
    REWARD_KEYS = (
        "continous_action_reward", "effort_reward", "thrust_reward", "pos_reward", "vel_direction_reward",
        "ups_reward", "spin_reward", "yaw_reward", "reward",
    )

    def __init__(self, spec: QuadSpec, num_envs: int, dtype=torch.float32, rng: str = "torch"):
        assert rng in ("torch", "explicit")
        self.spec, self.num_envs, self.dtype, self.rng = spec, num_envs, dtype, rng
        self.ctl_mode = spec.ctl_mode
        self.num_actions, self.num_obs = spec.num_actions, spec.num_obs
        self.max_episode_length = spec.max_episode_length
        self.dt = spec.dt
        N = num_envs
        # BaseTask buffers (base_task.py:73-76)
        self.obs_buf = torch.zeros(N, self.num_obs, dtype=dtype)
        self.rew_buf = torch.zeros(N, dtype=dtype)
        self.reset_buf = torch.ones(N, dtype=torch.long)
        self.time_out_buf = torch.zeros(N, dtype=torch.bool)
        self.progress_buf = torch.zeros(N, dtype=torch.long)  # hovering.py:164-165
        self.extras = {}
        # root state tensor: zero pose, identity quaternion (what IsaacGym hands back before any reset)
        self.root_states = torch.zeros(N, 13, dtype=dtype)
        self.root_states[:, 6] = 1.0
        self.initial_root_states = self.root_states.clone()  # hovering.py:87
        self.action_lower_limits = torch.tensor(spec.act_lo, dtype=dtype)
        self.action_upper_limits = torch.tensor(spec.act_hi, dtype=dtype)
        self.controller = ParallelControl(N, spec, dtype) if self.ctl_mode != "prop" else None
        self.cmd_thrusts = torch.zeros(N, 4, dtype=dtype)
        self.target_states = torch.tensor(spec.target_state, dtype=dtype).repeat(N, 1)  # hovering.py:132
        self.actions = torch.zeros(N, self.num_actions, dtype=dtype)
        self.pre_actions = torch.zeros(N, self.num_actions, dtype=dtype)
        self.noise_sigma = spec.noise_sigma
        self.last_draws = None
        self._explicit = None

    # views like hovering.py:74-77
    @property
    def root_positions(self): return self.root_states[:, 0:3]
    @property
    def root_quats(self): return self.root_states[:, 3:7]
    @property
    def root_linvels(self): return self.root_states[:, 7:10]
    @property
    def root_angvels(self): return self.root_states[:, 10:13]

    # ---- randomness -----------------------------------------------------------------------------------------
    def _reset_uniforms(self, env_ids, which):
        n = len(env_ids)
        if self.rng == "explicit":
            u = self._explicit["reset"][env_ids, which].to(self.dtype)
        else:  # reference call order: xy, z, roll/pitch, yaw, linvel, angvel (hovering.py:316-329)
            u = torch.cat([torch.rand(n, k) for k in (2, 1, 2, 1, 3, 3)], -1).to(self.dtype)
        if which is not None and self.last_draws is not None:
            self.last_draws["reset"][env_ids, which] = u.to(torch.float32)
        return u

    # ---- reset_idx (hovering.py:310-335) ---------------------------------------------------------------------
    def _sample_pose(self, u):
        pi = math.pi
        xy = R.rand_float(-1.0, 1.0, u[:, 0:2])
        z = R.rand_float(-1.0, 1.0, u[:, 2:3])
        ang = torch.cat((0.01 * R.rand_float(-pi, pi, u[:, 3:5]), 0.05 * R.rand_float(-pi, pi, u[:, 5:6])), -1)
        return xy, z, ang

    def reset_idx(self, env_ids, which=None, u=None):
        if u is None:
            u = self._reset_uniforms(env_ids, which)
        self.root_states[env_ids] = self.initial_root_states[env_ids]
        xy, z, ang = self._sample_pose(u)
        self.root_states[env_ids, 0:2] = xy
        self.root_states[env_ids, 2:3] = z
        quat = R.matrix_to_quaternion(R.euler_angles_to_matrix(ang, "XYZ"))  # wxyz
        self.root_states[env_ids, 3:7] = quat[:, [1, 2, 3, 0]]
        self.root_states[env_ids, 7:10] = 0.5 * R.rand_float(-1.0, 1.0, u[:, 6:9])
        self.root_states[env_ids, 10:13] = 0.2 * R.rand_float(-1.0, 1.0, u[:, 9:12])
        self.reset_buf[env_ids] = 1
        self.progress_buf[env_ids] = 0
        self.pre_actions[env_ids] = 0
        if self.spec.ctrl_reset and self.controller is not None:
            self.controller.reset(env_ids)

    # ---- pre_physics_step (hovering.py:203-281) -------------------------------------------------------------
    def pre_physics_step(self, _actions):
        reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_env_ids) > 0:
            self.reset_idx(reset_env_ids, which=0)
        self.actions = _actions  # .to(device) is the identity on the same device → aliasing (quirk Q4)
        if self.ctl_mode in ("rate", "atti"):
            self.actions[..., -1] = 0.5 + 0.5 * self.actions[..., -1]
        self.actions = R.tensor_clamp(self.actions, self.action_lower_limits, self.action_upper_limits)
        # quaternion sign canonicalisation written into the state (hovering.py:224-226)
        self.root_states[..., 3:7] = torch.where(self.root_states[..., 6:7] < 0, -self.root_states[..., 3:7],
                                                 self.root_states[..., 3:7])
        self.pre_step_quat = self.root_states[:, 3:7].clone()  # (test hook: attitude seen by the controller)
        pos, quat = self.root_states[:, 0:3].clone(), self.root_states[:, 3:7].clone()
        linvel, angvel = self.root_states[:, 7:10].clone(), self.root_states[:, 10:13].clone()
        q_wxyz = quat[:, [3, 0, 1, 2]]
        if self.ctl_mode in ("pos", "vel", "atti"):
            self.controller.set_status(pos, q_wxyz, linvel, angvel, 0.01)
            self.cmd_thrusts = self.controller.update(self.actions)
        elif self.ctl_mode == "rate":
            self.controller.set_q_world(q_wxyz)
            self.cmd_thrusts = self.controller.update(self.actions, angvel, 0.01)
        else:  # prop
            self.cmd_thrusts = self.actions
        if self.rng == "torch":
            torch.rand(self.num_envs, 1)  # `delta = .0*torch_rand_float(...) + 9.59` consumes a draw (:256)
        thrusts = self.cmd_thrusts * self.spec.k_thrust
        thrusts = thrusts.clone()
        thrusts[reset_env_ids] = 0  # hovering.py:268
        prop_rot = self.cmd_thrusts * self.spec.k_torque
        tau_z = -prop_rot[:, 0] - prop_rot[:, 1] + prop_rot[:, 2] + prop_rot[:, 3]  # :272-275, LOCAL_SPACE z
        return thrusts, tau_z

    # ---- step (hovering.py:286-308) -----------------------------------------------------------------------------
    def step(self, actions, rand_reset=None, rand_noise=None):
        N = self.num_envs
        if self.rng == "explicit":
            assert rand_reset is not None and (rand_noise is not None or self.spec.no_noise)
            self._explicit = {"reset": rand_reset.reshape(N, 2, self.RESET_DRAWS), "noise": rand_noise}
        self.last_draws = {"reset": torch.zeros(N, 2, self.RESET_DRAWS), "noise": torch.zeros(N, 18)}
        thrusts, tau_z = self.pre_physics_step(actions)
        self.root_matrix3 = simulate(self.spec, self.root_states, thrusts, tau_z)  # gym.simulate + refresh
        self.progress_buf += 1
        self.compute_observations()
        self.compute_reward()
        reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_env_ids) > 0:
            self.reset_idx(reset_env_ids, which=1)
        self.time_out_buf = self.progress_buf > self.max_episode_length
        self.extras["time_outs"] = self.time_out_buf
        self.extras["item_reward_info"] = self.item_reward_info
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras

    def reset(self, rand_reset=None, rand_noise=None):  # base_task.py:107-111
        self.last_draws = None
        ids = torch.arange(self.num_envs)
        if self.rng == "torch":
            self.reset_idx(ids)  # draws are consumed, the states are re-drawn by the step below (quirk Q1/Q3)
        else:  # the first draw is invisible (overwritten by the step's own reset): only the bookkeeping remains
            self.reset_buf[:] = 1
            self.progress_buf[:] = 0
            self.pre_actions[:] = 0
        obs, priv, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, dtype=self.dtype), rand_reset, rand_noise)
        return obs, priv

    # ---- observations (hovering.py:337-358) --------------------------------------------------------------------
    def _noise(self):
        if self.spec.no_noise:
            return torch.zeros(self.num_envs, 18, dtype=self.dtype)
        if self.rng == "explicit":
            z = self._explicit["noise"].to(self.dtype)
        else:
            z = torch.cat([torch.randn(self.num_envs, k) for k in (9, 3, 3, 3)], -1).to(self.dtype)
        self.last_draws["noise"] = z.to(torch.float32)
        return z

    def _fill_base_obs(self):
        self.root_matrix = R.quaternion_to_matrix(self.root_quats[:, [3, 0, 1, 2]]).reshape(self.num_envs, 9)
        self.obs_buf[..., 0:9] = self.root_matrix
        self.obs_buf[..., 9:12] = self.root_positions
        self.obs_buf[..., 12:15] = self.root_linvels
        self.obs_buf[..., 15:18] = self.root_angvels

    def add_noise(self):
        z = self._noise()
        s = self.noise_sigma
        self.obs_buf[..., 0:9] += s[0] * z[:, 0:9]
        self.obs_buf[..., 9:12] += s[1] * z[:, 9:12]
        self.obs_buf[..., 12:15] += s[2] * z[:, 12:15]
        self.obs_buf[..., 15:18] += s[3] * z[:, 15:18]

    def compute_observations(self):
        self._fill_base_obs()
        self.add_noise()
        self.obs_buf[..., 0:18] -= self.target_states
        return self.obs_buf

    # ---- reward (hovering.py:360-459) ---------------------------------------------------------------------------
    def compute_reward(self):
        reward, reset, info = self.compute_quadcopter_reward()
        self.rew_buf[:] = reward
        self.reset_buf[:] = reset
        self.item_reward_info = info
        self.pre_actions = self.actions.clone()

    def _common_terms(self):
        thrust_cmds = torch.clamp(self.cmd_thrusts, min=0.0, max=1.0)
        effort_reward = 0.1 * (1 - thrust_cmds).sum(-1) / 4
        action_diff = self.actions - self.pre_actions
        target_matrix = self.target_states[..., 0:9].reshape(self.num_envs, 3, 3)
        target_euler = R.matrix_to_euler_xyz(target_matrix)
        root_matrix = R.quaternion_to_matrix(self.root_quats[:, [3, 0, 1, 2]])
        root_euler = R.matrix_to_euler_xyz(root_matrix)
        yaw_diff = R.compute_yaw_diff(target_euler[..., 2], root_euler[..., 2]) / torch.pi
        spinnage = torch.square(self.root_angvels[:, -1])
        ups = R.quat_axis(self.root_quats, 2)
        ups_reward = torch.square((ups[..., 2] + 1) / 2)
        return effort_reward, action_diff, yaw_diff, spinnage, ups, ups_reward

    def compute_quadcopter_reward(self):
        effort_reward, action_diff, yaw_diff, spinnage, ups, ups_reward = self._common_terms()
        thrust_mode = self.ctl_mode in ("rate", "atti")
        if not thrust_mode:
            continous_action_reward = 0.2 * torch.exp(-torch.norm(action_diff[..., :], dim=-1))
            thrust_reward = 0
        else:
            continous_action_reward = 0.2 * torch.exp(-torch.norm(action_diff[..., :-1], dim=-1)) + 0.5 / (
                1.0 + torch.square(3 * action_diff[..., -1]))
            thrust = self.actions[..., -1]
            thrust_reward = 0.1 * (1 - torch.abs(0.1533 - thrust))
        target_positions = self.target_states[..., 9:12]
        relative_positions = target_positions - self.root_positions
        pos_diff = torch.norm(relative_positions, dim=-1)
        pos_reward = 0.7 / (1.0 + torch.square(1.6 * pos_diff))
        tar_direction = relative_positions / torch.norm(relative_positions, dim=1, keepdim=True)
        vel_direction = self.root_linvels / torch.norm(self.root_linvels, dim=1, keepdim=True)
        dot_product = (tar_direction * vel_direction).sum(dim=1)
        angle_diff = torch.acos(dot_product.clamp(-1.0, 1.0)).abs()
        vel_direction_reward = 0.1 * torch.exp(-angle_diff / torch.pi)
        yaw_reward = 1.0 / (1.0 + torch.square(3 * yaw_diff))
        spin_reward = 1.0 / (1.0 + torch.square(3 * spinnage))
        if not thrust_mode:
            reward = continous_action_reward + effort_reward + pos_reward + pos_reward * (
                vel_direction_reward + ups_reward + spin_reward + yaw_reward)
        else:
            reward = continous_action_reward + effort_reward + thrust_reward + pos_reward + pos_reward * (
                vel_direction_reward + ups_reward + spin_reward + yaw_reward)
        ones = torch.ones_like(self.reset_buf)
        die = torch.zeros_like(self.reset_buf)
        reset = torch.where(self.progress_buf >= self.max_episode_length - 1, ones, die)
        reset = torch.where(torch.norm(relative_positions, dim=1) > 4, ones, reset)
        reset = torch.where(relative_positions[..., 2] < -2, ones, reset)
        reset = torch.where(relative_positions[..., 2] > 2, ones, reset)
        reset = torch.where(ups[..., 2] < 0.0, ones, reset)
        if self.ctl_mode == "atti":
            reset = torch.where(self.actions[..., 0] < 0, ones, reset)
        info = {
            "continous_action_reward": continous_action_reward, "effort_reward": effort_reward,
            "thrust_reward": thrust_reward, "pos_reward": pos_reward, "vel_direction_reward": vel_direction_reward,
            "ups_reward": ups_reward, "spin_reward": spin_reward, "yaw_reward": yaw_reward, "reward": reward,
        }
        return reward, reset, info

    def reward_terms_matrix(self):
        """[9,N] float32 in the plane order of AgxStepIO.reward_terms."""
        rows = []
        for k in self.REWARD_KEYS:
            v = self.item_reward_info[k]
            rows.append(v.to(torch.float32) if torch.is_tensor(v) else torch.full((self.num_envs,), float(v)))
        return torch.stack(rows, 0)


class TrackingOracle(HoveringOracle):
    """Tracking deltas (tracking.py): reset sampler :159-192, lemniscate :194-200, obs :202-214, reward :223-296."""

    REWARD_KEYS = (
        "dist_norm", "dist_reward", "yaw_reward", "spin_reward", "continous_action_reward", "thrust_reward",
        "effort_reward", "ups_reward", "reward",
    )

    def _sample_pose(self, u):
        pi = math.pi
        xy = 0.1 * R.rand_float(-1.0, 1.0, u[:, 0:2])
        z = 0.1 * R.rand_float(-1.0, 1.0, u[:, 2:3]) + 1.0
        ang = torch.cat((0.1 * R.rand_float(-pi, pi, u[:, 3:5]), 0.2 * R.rand_float(-pi, pi, u[:, 5:6])), -1)
        return xy, z, ang

    def compute_traj_lemniscate(self, n_steps=10, step_size=5, scale=0.25):
        step = self.progress_buf.unsqueeze(1).expand(-1, n_steps) + torch.arange(n_steps).repeat(self.num_envs, 1) * step_size
        t = step.to(self.dtype) * self.dt * scale
        ref_x = 3 * torch.sin(t) / (1 + torch.cos(t) ** 2)
        ref_y = 3 * torch.sin(t) * torch.cos(t) / (1 + torch.cos(t) ** 2)
        ref_z = torch.ones_like(ref_x)
        return torch.stack((ref_x, ref_y, ref_z), dim=-1)

    def compute_observations(self):
        self._fill_base_obs()
        self.ref_positions = self.compute_traj_lemniscate()
        self.related_future_pos = (self.ref_positions - self.root_positions.clone().unsqueeze(1)).reshape(self.num_envs, -1)
        self.obs_buf[..., 18:48] = self.related_future_pos
        self.add_noise()
        return self.obs_buf

    def compute_quadcopter_reward(self):
        effort_reward, action_diff, yaw_diff, spinnage, ups, ups_reward = self._common_terms()
        thrust_mode = self.ctl_mode in ("rate", "atti")
        if not thrust_mode:
            continous_action_reward = 0.2 * torch.exp(-torch.norm(action_diff[..., :], dim=-1))
            thrust_reward = 0
        else:
            continous_action_reward = 0.1 * torch.exp(-torch.norm(action_diff[..., :-1], dim=-1)) + 0.5 / (
                1.0 + torch.square(2 * action_diff[..., -1]))
            thrust = self.actions[..., -1]
            thrust_reward = 0.1 * (1 - torch.abs(0.1533 - thrust))
        dist_diff = self.ref_positions[:, 0] - self.root_positions
        dist_norm = torch.norm(dist_diff, dim=-1)
        dist_reward = 1.0 / (1.0 + torch.square(1.8 * dist_norm))
        yaw_reward = 1 / (1.0 + torch.square(4 * yaw_diff))
        spin_reward = 1 / (1.0 + torch.square(2 * spinnage))
        if not thrust_mode:
            reward = continous_action_reward + effort_reward + dist_reward + dist_reward * (
                spin_reward + yaw_reward + ups_reward)
        else:
            reward = continous_action_reward + effort_reward + thrust_reward + dist_reward + dist_reward * (
                spin_reward + yaw_reward + ups_reward)
        ones = torch.ones_like(self.reset_buf)
        die = torch.zeros_like(self.reset_buf)
        reset = torch.where(self.progress_buf >= self.max_episode_length - 1, ones, die)
        reset = torch.where(dist_norm > 1.0, ones, reset)
        if self.ctl_mode == "atti":
            reset = torch.where(self.actions[..., 0] < 0, ones, reset)
        info = {
            "dist_norm": dist_norm, "dist_reward": dist_reward, "yaw_reward": yaw_reward, "spin_reward": spin_reward,
            "continous_action_reward": continous_action_reward, "thrust_reward": thrust_reward,
            "effort_reward": effort_reward, "ups_reward": ups_reward, "reward": reward,
        }
        return reward, reset, info


def make_oracle(spec: QuadSpec, num_envs: int, dtype=torch.float32, rng="torch"):
    from .customized import BalloonOracle
    from .image_tasks import AvoidOracle, PlanningOracle

    return {"hovering": HoveringOracle, "tracking": TrackingOracle, "balloon": BalloonOracle, "avoid": AvoidOracle,
            "planning": PlanningOracle}[spec.task](spec, num_envs, dtype, rng)
