"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the reference's two depth-camera tasks: Avoid (``airgym/envs/task/avoid.py:58-295``) and Planning
(``airgym/envs/task/planning.py:63-307``) on top of the Customized step (``airgym/envs/base/customized.py:216-435``,
oracle/customized.py), including the depth post-processing of `dump_images` (customized.py:399-435).  In-tree reference
logic (reset samplers and their draw order, observation packing, rewards, terminations, image noise/blur) is pinned by
tests/golden/make_golden.py against the reference's own classes; the camera, the contact model and the cube's flight
(oracle/scene.py) are builder-defined stand-ins for IsaacGym/PhysX and PARITY UNPINNED.

Compact explicit-draw layouts (what the kernel consumes; zero-weighted reference draws are dropped):
  avoid    D = 11: [u_mode, theta, aim_x, aim_y, aim_z, x, y, z, roll, pitch, yaw]
  planning D = 124: [asset x (41), asset y (41), asset yaw (41), goal y]      (asset 0 = the goal ball, 1..40 = trees)
Image noise per render: add [N,212,120] (the N(0,0.1) sample), mul [N,212,120] (the N(1,0.3) sample), kern [N,25] (randint/256).
"""
import math

import torch
import torch.nn.functional as F

from . import rotations as R
from . import scene
from .customized import CustomizedOracle
from .rigid_body import simulate

LENGTH, WIDTH, FLY_HEIGHT = 8.0, 4.0, 1.5  # planning.py:10-12


def compute_yaw_diff(a, b):  # avoid.py:26-31
    return R.compute_yaw_diff(a, b)


class ImageTaskOracle(CustomizedOracle):
    """Customized.step with the camera (customized.py:308-344, 386-435)."""
    cam_every = 4  # cam_dt / dt = 0.04 / 0.01 (avoid_config.py:22, planning_config.py:22)

    def __init__(self, spec, num_envs, dtype=torch.float32, rng="torch"):
        super().__init__(spec, num_envs, dtype, rng)
        self.full_camera_array = torch.zeros(num_envs, 1, scene.CAM_W, scene.CAM_H, dtype=dtype)
        self.pre_root_positions = torch.zeros(num_envs, 3, dtype=dtype)
        self.pre_root_angvels = torch.zeros(num_envs, 3, dtype=dtype)
        self.last_image_draws = None
        self._explicit_image = None

    # ---- scene hooks (builder-defined, oracle/scene.py) -----------------------------------------------------
    def scene_kwargs(self):
        raise NotImplementedError

    def refresh_contact_forces(self):
        hit = scene.drone_contacts(self.root_positions, **self.contact_kwargs())
        self.contact_forces.zero_()
        self.contact_forces[hit, 2] = 1.0

    def raw_depth(self):
        return scene.render_depth(self.root_positions, self.root_quats, **self.scene_kwargs())  # [N,W,H], +inf = no hit

    # ---- dump_images (customized.py:399-435), per env like the reference (draw order!) ----------------------
    def dump_images(self, depth):
        N = self.num_envs
        rec = {"add": torch.zeros(N, scene.CAM_W, scene.CAM_H), "mul": torch.zeros(N, scene.CAM_W, scene.CAM_H),
               "kern": torch.zeros(N, 25)}
        for e in range(N):
            img = depth[e].unsqueeze(0)  # [1,W,H] = -camera_tensor.T
            img = torch.where(img > 4.5, torch.tensor(4.5), img)
            img = torch.clamp(img, 0, 4.5) / 4.5
            if self.rng == "explicit":
                add, mul, kern = (self._explicit_image[k][e] for k in ("add", "mul", "kern"))
                add, mul, kern = add.unsqueeze(0), mul.unsqueeze(0), kern.reshape(5, 5)
            else:
                add = torch.normal(0.0, 0.1, size=img.shape)
            img = torch.clamp(img + add, 0.0, img.max())
            if self.rng != "explicit":
                mul = torch.normal(1.0, 0.3, size=img.shape)
            img = torch.clamp(img * mul, 0.0, img.max())
            if self.rng != "explicit":
                kern = torch.randint(0, 256, (5, 5), dtype=torch.float32) / 256.0
            img = F.conv2d(img.unsqueeze(0), kern.reshape(1, 1, 5, 5), padding=2).squeeze(0)
            self.full_camera_array[e, :] = img
            rec["add"][e], rec["mul"][e], rec["kern"][e] = add[0], mul[0], kern.reshape(25)
        self.last_image_draws = rec

    def render_cameras(self):
        self.dump_images(self.raw_depth())

    # ---- step ------------------------------------------------------------------------------------------------
    def object_physics(self):
        pass

    def pre_reward_hook(self):
        pass

    def post_step_hook(self):
        pass

    def step(self, actions, rand_reset=None, rand_noise=None, rand_image=None):
        N = self.num_envs
        if self.rng == "explicit":
            self._explicit = {"reset": rand_reset.reshape(N, 2, self.RESET_DRAWS), "noise": rand_noise}
            self._explicit_image = rand_image
        self.last_draws = {"reset": torch.zeros(N, 2, self.RESET_DRAWS), "noise": torch.zeros(N, 18)}
        self.last_image_draws = None
        self.actions_local = actions  # avoid.py:162 / planning.py:143 — the very tensor pre_physics_step remaps in place
        thrusts, tau_z = self.pre_physics_step(actions)
        simulate(self.spec, self.root_states, thrusts, tau_z)
        self.object_physics()
        self.refresh_contact_forces()
        self.rendered = self.counter % self.cam_every == 0
        if self.rendered:
            self.render_cameras()
        self.progress_buf += 1
        self.check_collisions()
        self.compute_observations()
        self.pre_reward_hook()
        self.compute_reward()
        if self.reset_on_collision:
            self.reset_buf = torch.where(self.collisions > 0, torch.ones_like(self.reset_buf), self.reset_buf)
        reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_env_ids) > 0:
            self.reset_idx(reset_env_ids, which=1)
        self.time_out_buf = self.progress_buf > self.max_episode_length
        self.extras["time_outs"] = self.time_out_buf
        self.extras["item_reward_info"] = self.item_reward_info
        self.post_step_hook()
        obs = {"image": self.full_camera_array, "observation": self.obs_buf}
        return obs, None, self.rew_buf, self.reset_buf, self.extras

    def local_frame(self):
        """world_to_local = Rz(yaw)^T with yaw = atan2(R10, R00) (avoid.py:204-214: the dim=2 stack transposes the rows)."""
        rot = R.quaternion_to_matrix(self.root_quats[:, [3, 0, 1, 2]])
        yaw = torch.atan2(rot[:, 1, 0], rot[:, 0, 0])
        c, s, z, o = torch.cos(yaw), torch.sin(yaw), torch.zeros_like(yaw), torch.ones_like(yaw)
        W = torch.stack([torch.stack([c, -s, z], dim=1), torch.stack([s, c, z], dim=1), torch.stack([z, z, o], dim=1)], dim=2)
        return rot, W


class AvoidOracle(ImageTaskOracle):
    RESET_DRAWS = 11
    reset_on_collision = True  # avoid_config.py:19
    REWARD_KEYS = ("pose_reward", "ups_reward", "spin_reward", "effort_reward", "action_smoothness_reward", "thrust_reward",
                   "alive_reward", "_pad0", "reward")

    def __init__(self, spec, num_envs, dtype=torch.float32, rng="torch"):
        super().__init__(spec, num_envs, dtype, rng)
        self.object_states = torch.zeros(num_envs, 13, dtype=dtype)
        self.object_states[:, 6] = 1.0

    @property
    def object_positions(self): return self.object_states[:, 0:3]
    @property
    def object_linvels(self): return self.object_states[:, 7:10]

    def scene_kwargs(self): return {"cube": self.object_positions}
    def contact_kwargs(self): return {"cube": self.object_positions}

    def object_physics(self):
        scene.cube_step(self.object_states[:, 0:3], self.object_states[:, 7:10], self.dt, self.spec.gravity)

    def _reset_uniforms(self, env_ids, which):
        n = len(env_ids)
        if self.rng == "explicit":
            u = self._explicit["reset"][env_ids, which].to(self.dtype)
        else:  # avoid.py:95-148 call order; theta / z / aim are drawn for the thrown subset only
            u = torch.zeros(n, self.RESET_DRAWS, dtype=self.dtype)
            u[:, 0:1] = torch.rand(n, 1)
            thrown = u[:, 0] < 0.8
            m = int(thrown.sum())
            if m > 0:
                u[thrown, 1:2] = torch.rand(m, 1)
                torch.rand(m, 1)  # `0.0 * torch_rand_float(...) + 1.4` (avoid.py:112)
                u[thrown, 2:5] = torch.rand(m, 3)
            u[:, 5:7] = torch.rand(n, 2)
            u[:, 7:8] = torch.rand(n, 1)
            u[:, 8:10] = torch.rand(n, 2)
            u[:, 10:11] = torch.rand(n, 1)
            torch.rand(n, 3); torch.rand(n, 3)  # 0.*linvel, 0.*angvel (avoid.py:146-147)
        if which is not None and self.last_draws is not None:
            self.last_draws["reset"][env_ids, which] = u.to(torch.float32)
        return u

    def calculate_object_velocity(self, positions, aim_u, v_e=4.5, g=9.81):  # avoid.py:58-89
        drone_position = 0.3 * R.rand_float(-1.0, 1.0, aim_u) + torch.tensor([0.0, 0.0, 1.0], dtype=self.dtype)
        direction = drone_position - positions
        distance_xy = torch.norm(direction[:, :2], dim=1, keepdim=True)
        unit_direction_xy = direction[:, :2] / distance_xy
        ve = torch.tensor(v_e, dtype=self.dtype).expand_as(distance_xy)
        t = distance_xy / ve
        z_c, z_u = positions[:, 2].unsqueeze(1), drone_position[:, 2].unsqueeze(1)
        v_z = (z_u - z_c + 0.5 * g * t ** 2) / t
        return torch.cat([unit_direction_xy[:, 0:1] * ve, unit_direction_xy[:, 1:2] * ve, v_z], dim=1)

    def reset_idx(self, env_ids, which=None, u=None):  # avoid.py:91-158
        if u is None:
            u = self._reset_uniforms(env_ids, which)
        pi = math.pi
        thrown = u[:, 0] < 0.8
        if torch.any(thrown):
            ids = env_ids[thrown]
            theta = pi / 6 * R.rand_float(-1.0, 1.0, u[thrown, 1:2])
            self.object_states[ids, 0:1] = 4.2 * torch.cos(theta)
            self.object_states[ids, 1:2] = 4.2 * torch.sin(theta)
            self.object_states[ids, 2:3] = 1.4
            self.object_states[ids, 7:10] = self.calculate_object_velocity(self.object_states[ids, 0:3], u[thrown, 2:5])
        if torch.any(~thrown):
            ids = env_ids[~thrown]
            self.object_states[ids, 0:3] = torch.tensor([-999.0, -999.0, 0.0], dtype=self.dtype)
            self.object_states[ids, 7:10] = 0.0
        self.root_states[env_ids] = self.initial_root_states[env_ids]
        self.root_states[env_ids, 0:2] = 0.2 * R.rand_float(-1.0, 1.0, u[:, 5:7])
        self.root_states[env_ids, 2:3] = 0.2 * R.rand_float(-1.0, 1.0, u[:, 7:8]) + 1.0
        ang = torch.cat((0.01 * R.rand_float(-pi, pi, u[:, 8:10]), 0.05 * R.rand_float(-pi, pi, u[:, 10:11])), -1)
        quat = R.matrix_to_quaternion(R.euler_angles_to_matrix(ang, "XYZ"))
        self.root_states[env_ids, 3:7] = quat[:, [1, 2, 3, 0]]
        self.root_states[env_ids, 7:13] = 0.0
        self.reset_buf[env_ids] = 1
        self.progress_buf[env_ids] = 0
        self.pre_actions[env_ids] = 0
        self.pre_root_positions[env_ids] = 0
        self.pre_root_angvels[env_ids] = 0
        if self.spec.ctrl_reset and self.controller is not None:
            self.controller.reset(env_ids)

    def compute_observations(self):  # avoid.py:203-226
        rot, W = self.local_frame()
        self.world_to_local = W
        self.euler_angles_local = R.matrix_to_euler_xyz(torch.bmm(W, rot))
        self.vel_local = torch.einsum("bij,bj->bi", W, self.root_linvels)
        self.ang_vel_local = torch.einsum("bij,bj->bi", W, self.root_angvels)
        self.obs_buf[..., 0:3] = self.root_positions - self.target_states[..., 9:12]
        self.obs_buf[..., 3:6] = self.euler_angles_local
        self.obs_buf[..., 6:9] = self.vel_local
        self.obs_buf[..., 9:12] = self.ang_vel_local
        self.obs_buf[..., 12:16] = self.actions_local  # needs A = 4: the reference cannot run these tasks in atti mode

    def compute_reward(self):  # avoid.py:228-233
        reward, reset, info = self.compute_quadcopter_reward()
        self.rew_buf[:] = reward
        self.reset_buf[:] = reset
        self.item_reward_info = info
        self.pre_actions = self.actions.clone()
        self.pre_root_positions = self.root_positions.clone()
        self.pre_root_angvels = self.root_angvels.clone()

    def compute_quadcopter_reward(self):  # avoid.py:235-295
        target_positions = self.target_states[..., 9:12]
        relative_positions = target_positions - self.root_positions
        target_euler = R.matrix_to_euler_xyz(self.target_states[..., 0:9].reshape(self.num_envs, 3, 3))
        root_euler = R.matrix_to_euler_xyz(R.quaternion_to_matrix(self.root_quats[:, [3, 0, 1, 2]]))
        relative_heading = compute_yaw_diff(target_euler[..., 2], root_euler[..., 2])
        distance = torch.norm(torch.cat((relative_positions, relative_heading.unsqueeze(-1)), dim=-1), dim=1)
        pose_reward = 1.0 / (1.0 + torch.square(1.6 * distance))
        ups = R.quat_axis(self.root_quats, axis=2)
        ups_reward = torch.square((ups[..., 2] + 1) / 2)
        spinnage = torch.square(self.root_angvels[:, -1])
        spin_reward = 1.0 / (1.0 + torch.square(spinnage))
        effort_reward = 0.1 * torch.exp(-self.actions.pow(2).sum(-1))
        action_diff = torch.norm(self.actions[..., :-1] - self.pre_actions[..., :-1], dim=-1)
        thrust_reward = 0.05 * (1 - torch.abs(0.1533 - self.actions[..., -1]))
        action_smoothness_reward = 0.1 * torch.exp(-action_diff)
        alive_reward = torch.where(self.collisions > 0, -500.0, 0.5).to(self.dtype)
        reward = (pose_reward + pose_reward * (ups_reward + spin_reward) + effort_reward + action_smoothness_reward
                  + thrust_reward + alive_reward)
        ones, die = torch.ones_like(self.reset_buf), torch.zeros_like(self.reset_buf)
        reset = torch.where(self.progress_buf >= self.max_episode_length - 1, ones, die)
        reset = torch.where(self.root_positions[..., 2] < 0.3, ones, reset)
        reset = torch.where(self.root_positions[..., 2] > 1.7, ones, reset)
        reset = torch.where(relative_positions.norm(dim=-1) > 2.0, ones, reset)
        reset = torch.where(ups[..., 2] < 0.0, ones, reset)
        info = {"pose_reward": pose_reward, "ups_reward": ups_reward, "spin_reward": spin_reward, "effort_reward": effort_reward,
                "action_smoothness_reward": action_smoothness_reward, "thrust_reward": thrust_reward,
                "alive_reward": alive_reward, "_pad0": 0, "reward": reward}
        return reward, reset, info

    def aux_matrix(self):
        """[N,8] in the layout of AgxStepIO.aux for avoid: cube xyz, cube linvel xyz, collisions, pad."""
        a = torch.zeros(self.num_envs, 8)
        a[:, 0:3] = self.object_positions
        a[:, 3:6] = self.object_linvels
        a[:, 6] = self.collisions
        return a


class PlanningOracle(ImageTaskOracle):
    NUM_ASSETS = 41  # ball (goal) + 40 thin trees, asset_manager.py:79-152 order
    RESET_DRAWS = 124
    reset_on_collision = False  # planning_config.py:19
    REWARD_KEYS = ("continous_action_reward", "heading_reward", "speed_reward", "forward_reward", "alive_reward", "ups_reward",
                   "z_reward", "esdf_reward", "thrust_reward", "reach_goal_reward", "reward")

    def __init__(self, spec, num_envs, dtype=torch.float32, rng="torch"):
        super().__init__(spec, num_envs, dtype, rng)
        self.env_asset_root_states = torch.zeros(num_envs, self.NUM_ASSETS, 13, dtype=dtype)
        self.env_asset_root_states[:, :, 6] = 1.0
        self.asset_yaw = torch.zeros(num_envs, self.NUM_ASSETS, dtype=dtype)  # what the asset quaternions encode
        self.prev_related_dist = torch.zeros(num_envs, dtype=dtype)
        self.esdf_dist = torch.ones(num_envs, dtype=dtype) * 10

    @property
    def goal_states(self): return self.env_asset_root_states[:, 0, :]
    @property
    def goal_positions(self): return self.env_asset_root_states[:, 0, 0:3]

    def _trees(self): return (self.env_asset_root_states[:, 1:, 0:2], self.asset_yaw[:, 1:])
    def scene_kwargs(self): return {"trees": self._trees(), "ball": self.goal_positions}
    def contact_kwargs(self): return {"trees": self._trees()}

    def _reset_uniforms(self, env_ids, which):
        n, A = len(env_ids), self.NUM_ASSETS
        if self.rng == "explicit":
            u = self._explicit["reset"][env_ids, which].to(self.dtype)
        else:  # planning.py:66-112 call order
            u = torch.zeros(n, self.RESET_DRAWS, dtype=self.dtype)
            u[:, 0:A] = torch.rand(n, A, 1).squeeze(-1)
            u[:, A:2 * A] = torch.rand(n, A, 1).squeeze(-1)
            torch.rand(n, A, 2)  # 0 * roll/pitch of the assets
            u[:, 2 * A:3 * A] = torch.rand(n, A, 1).squeeze(-1)
            u[:, 3 * A:3 * A + 1] = torch.rand(n, 1)
            torch.rand(n, 1)  # .0 * goal z
            torch.rand(n, 1)  # .0 * root z
            torch.rand(n, 2); torch.rand(n, 1)  # 0.* root roll/pitch, yaw
            torch.rand(n, 3); torch.rand(n, 3)  # 0.* linvel, angvel
        if which is not None and self.last_draws is not None:
            self.last_draws["reset"][env_ids, which] = u.to(torch.float32)
        return u

    def reset_idx(self, env_ids, which=None, u=None):  # planning.py:63-136
        if u is None:
            u = self._reset_uniforms(env_ids, which)
        pi, A = math.pi, self.NUM_ASSETS
        self.env_asset_root_states[env_ids, :, 0] = LENGTH * R.rand_float(-1.0, 1.0, u[:, 0:A]) + 0.0
        self.env_asset_root_states[env_ids, :, 1] = WIDTH * R.rand_float(-1.0, 1.0, u[:, A:2 * A]) + 0.0
        self.env_asset_root_states[env_ids, :, 2] = 0
        yaw = R.rand_float(-pi, pi, u[:, 2 * A:3 * A])
        ang = torch.stack((torch.zeros_like(yaw), torch.zeros_like(yaw), yaw), -1)
        quat = R.matrix_to_quaternion(R.euler_angles_to_matrix(ang.reshape(-1, 3), "XYZ")).reshape(len(env_ids), A, 4)
        self.env_asset_root_states[env_ids, :, 3:7] = quat[:, :, [1, 2, 3, 0]]
        # the simulator only knows the quaternion: the yaw the scene stand-ins (camera, contacts) use is read back from it
        self.asset_yaw[env_ids] = 2.0 * torch.atan2(quat[:, :, 3], quat[:, :, 0])
        self.env_asset_root_states[env_ids, 0, 0] = LENGTH + 0.5
        self.env_asset_root_states[env_ids, 0, 1] = 1.5 * R.rand_float(-1.0, 1.0, u[:, 3 * A]) + 0.0
        self.env_asset_root_states[env_ids, 0, 2] = FLY_HEIGHT
        self.root_states[env_ids, 0:2] = torch.tensor([-LENGTH - 0.5, 0.0], dtype=self.dtype)
        self.root_states[env_ids, 2:3] = FLY_HEIGHT
        vec = self.env_asset_root_states[env_ids, 0, 0:2] - self.root_states[env_ids, 0:2]
        init_yaw = torch.atan2(vec[..., 1], vec[..., 0]).unsqueeze(-1)
        root_angle = torch.cat((torch.zeros(len(env_ids), 2, dtype=self.dtype), init_yaw), -1)
        q = R.matrix_to_quaternion(R.euler_angles_to_matrix(root_angle, "XYZ"))
        self.root_states[env_ids, 3:7] = q[:, [1, 2, 3, 0]]
        self.root_states[env_ids, 7:13] = 0.0
        self.reset_buf[env_ids] = 1
        self.progress_buf[env_ids] = 0
        self.pre_actions[env_ids] = 0
        self.prev_related_dist[env_ids] = 0
        self.pre_root_positions[env_ids] = 0
        self.pre_root_angvels[env_ids] = 0
        if self.spec.ctrl_reset and self.controller is not None:
            self.controller.reset(env_ids)

    def compute_observations(self):  # planning.py:186-214
        forward_global = self.goal_positions - self.root_positions
        rot, W = self.local_frame()
        self.world_to_local = W
        self.euler_angles_local = R.matrix_to_euler_xyz(torch.bmm(W, rot))
        self.pos_diff_local = torch.einsum("bij,bj->bi", W, forward_global)
        self.vel_local = torch.einsum("bij,bj->bi", W, self.root_linvels)
        self.ang_vel_local = torch.einsum("bij,bj->bi", W, self.root_angvels)
        self.goal_dir = self.pos_diff_local / torch.norm(self.pos_diff_local, dim=-1, keepdim=True)
        self.related_dist = torch.norm(forward_global, dim=-1)
        self.obs_buf[..., 0:3] = self.goal_dir
        self.obs_buf[..., 3:6] = self.euler_angles_local
        self.obs_buf[..., 6:9] = self.vel_local
        self.obs_buf[..., 9:12] = self.ang_vel_local
        self.obs_buf[..., 12:16] = self.actions_local

    def pre_reward_hook(self):  # planning.py:162-163
        self.esdf_dist = torch.min(self.full_camera_array.clone().view(self.num_envs, -1), dim=1).values

    def post_step_hook(self):  # planning.py:183
        self.prev_related_dist = self.related_dist

    def compute_reward(self):  # planning.py:216-221
        reward, reset, info = self.compute_quadcopter_reward()
        self.rew_buf[:] = reward
        self.reset_buf[:] = reset
        self.item_reward_info = info
        self.pre_actions = self.actions.clone()
        self.pre_root_positions = self.root_positions.clone()
        self.pre_root_angvels = self.root_angvels.clone()

    def compute_quadcopter_reward(self):  # planning.py:223-307
        action_diff = self.actions - self.pre_actions
        continous_action_reward = 0.2 * torch.norm(self.ang_vel_local, dim=-1) + 0.2 * torch.norm(action_diff, dim=-1)
        thrust_reward = 0.5 * (1 - torch.abs(0.1533 - self.actions[..., -1]))
        forward_reward = 0.1 * (torch.norm(self.goal_positions - self.pre_root_positions, dim=-1)
                                - torch.norm(self.goal_positions - self.root_positions, dim=-1))
        forward_vec = self.pos_diff_local / torch.norm(self.pos_diff_local, dim=-1, keepdim=True)
        heading_reward = forward_vec[..., 0] * 1.0 + forward_vec[..., 1] * 0.0 + forward_vec[..., 2] * 0.0
        speed_reward = -0.5 * (1 - torch.exp(-2 * torch.square(self.vel_local[..., 0] - 1.0)))
        z_reward = torch.min(torch.min(self.root_positions[..., 2] - 1.8, torch.tensor(0.0)), 1.2 - self.root_positions[..., 2])
        ups = R.quat_axis(self.root_quats, axis=2)
        ups_reward = torch.square((ups[..., 2] + 1) / 2)
        esdf_reward = 0.5 * (1 - torch.exp(-0.5 * torch.square(self.esdf_dist)))
        alive_reward = torch.where(self.esdf_dist > 0.3, torch.tensor(0.0), torch.tensor(-1.0))
        reach_goal = self.related_dist < 0.3
        reach_goal_reward = torch.where(reach_goal, torch.tensor(200.0), torch.tensor(0.0))
        reward = (continous_action_reward + forward_reward + alive_reward + esdf_reward + ups_reward + z_reward + speed_reward
                  + heading_reward + thrust_reward + reach_goal_reward)
        ones, die = torch.ones_like(self.reset_buf), torch.zeros_like(self.reset_buf)
        reset = torch.where(self.root_positions[..., 2] < FLY_HEIGHT - 0.3, ones, die)
        reset = torch.where(self.root_positions[..., 2] > FLY_HEIGHT + 0.3, ones, reset)
        reset = torch.where(self.root_positions[..., 0] < -LENGTH - 0.5, ones, reset)
        reset = torch.where(self.root_positions[..., 0] > LENGTH + 0.5, ones, reset)
        reset = torch.where(self.root_positions[..., 1] < -WIDTH, ones, reset)
        reset = torch.where(self.root_positions[..., 1] > WIDTH, ones, reset)
        reset = torch.where(self.collisions > 0, ones, reset)
        reset = torch.where(reach_goal, ones, reset)
        reset = torch.where(heading_reward < 0.25, ones, reset)
        reset = torch.where(self.progress_buf >= self.max_episode_length - 1, ones, reset)
        info = {"continous_action_reward": continous_action_reward, "heading_reward": heading_reward, "speed_reward": speed_reward,
                "forward_reward": forward_reward, "alive_reward": alive_reward, "ups_reward": ups_reward, "z_reward": z_reward,
                "esdf_reward": esdf_reward, "thrust_reward": thrust_reward, "reach_goal_reward": reach_goal_reward, "reward": reward}
        return reward, reset, info

    def aux_matrix(self):
        """[N,8] in the layout of AgxStepIO.aux for planning: goal xyz, pre_root_positions xyz, collisions, esdf_dist."""
        a = torch.zeros(self.num_envs, 8)
        a[:, 0:3] = self.goal_positions
        a[:, 3:6] = self.pre_root_positions
        a[:, 6] = self.collisions
        a[:, 7] = self.full_camera_array.reshape(self.num_envs, -1).min(dim=1).values  # = esdf_dist once a step has run
        return a

    def asset_matrix(self):
        """[N,164] in the layout of the kernel's asset row: x (41) | y (41) | cos yaw (41) | sin yaw (41) — world frame."""
        a = torch.zeros(self.num_envs, 164)
        a[:, 0:41] = self.env_asset_root_states[:, :, 0]
        a[:, 41:82] = self.env_asset_root_states[:, :, 1]
        a[:, 82:123] = torch.cos(self.asset_yaw)
        a[:, 123:164] = torch.sin(self.asset_yaw)
        return a
