"""Hovering — host-side mirror of the reference task class (airgym/envs/base/hovering.py:40-459).

Same constructor signature, attributes and method names; the body of `step` is ONE call into libagx.so
(`agx_step`, include/agx.h) which fuses pre_physics_step (:203-281), the rlPx4Controller cascade (:217-250),
gym.simulate (:290), compute_observations (:337-358), compute_reward (:360-459), both reset_idx passes
(:209-211, :300-302) and the time-out flag (:304) into a single sm_100a kernel over the num_envs axis.
"""
import ctypes as C

import torch

from ... import _capi
from .base_task import BaseTask
from .hovering_config import HoveringCfg


class Hovering(BaseTask):
    TASK = "hovering"
    REWARD_KEYS = (  # plane order of AgxStepIO.reward_terms; hovering.py:447-457
        "continous_action_reward", "effort_reward", "thrust_reward", "pos_reward", "vel_direction_reward",
        "ups_reward", "spin_reward", "yaw_reward", "reward",
    )

    def __init__(self, cfg: HoveringCfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        self.cfg = cfg
        assert cfg.env.ctl_mode is not None, "Please specify one control mode!"
        if cfg.env.ctl_mode not in _capi.CTL_IDS:
            raise ValueError(f"Mode Error! ctl_mode must be one of {sorted(_capi.CTL_IDS)}, got {cfg.env.ctl_mode!r}")
        self.ctl_mode = cfg.env.ctl_mode
        self.cfg.env.num_actions = 5 if cfg.env.ctl_mode == "atti" else 4
        self.max_episode_length = int(self.cfg.env.episode_length_s / self.cfg.sim.dt)
        self.debug_viz = False
        super().__init__(self.cfg, sim_params, physics_engine, sim_device, headless)
        dev, N, A = self._dev, self.num_envs, self.num_actions

        # ---- parameters of the fused step (cfg → AgxParams)
        P = _capi.default_params(self.TASK, self.ctl_mode)
        P.dt = float(cfg.sim.dt)
        P.gravity = float(-cfg.sim.gravity[2])
        P.max_episode_length = self.max_episode_length
        tgt = torch.tensor(cfg.env.target_state, dtype=torch.float32)
        for i in range(18):
            P.target[i] = float(tgt[i])
        P.target_yaw = float(torch.atan2(-tgt[1], tgt[0]))  # matrix_to_euler_angles(target,'XYZ')[2]
        bk = getattr(cfg, "backend", HoveringCfg.backend)
        P.integrator = _capi.INT_EULER if bk.integrator == "euler" else _capi.INT_RK4
        P.flags = (_capi.FLAG_MUTATE_ACTIONS if bk.mutate_input_actions else 0) | (
            _capi.FLAG_CTRL_RESET if bk.ctrl_reset_on_reset else 0)
        self.params = P

        # ---- state tensors (hovering.py:70-77): one actor per env → [N,13] contiguous
        self.vec_root_tensor = torch.zeros(N, 1, 13, device=dev, dtype=torch.float32)
        self.vec_root_tensor[:, 0, 6] = 1.0
        self.root_states = self.vec_root_tensor[:, 0, :]
        self.root_positions = self.root_states[..., 0:3]
        self.root_quats = self.root_states[..., 3:7]  # x,y,z,w
        self.root_linvels = self.root_states[..., 7:10]
        self.root_angvels = self.root_states[..., 10:13]
        self.privileged_obs_buf = None
        self.initial_root_states = self.root_states.clone()
        self.counter = 0
        self.progress_buf = torch.zeros(N, device=dev, dtype=torch.long)

        self.action_lower_limits = torch.tensor(list(P.act_lo)[:A], device=dev, dtype=torch.float32)
        self.action_upper_limits = torch.tensor(list(P.act_hi)[:A], device=dev, dtype=torch.float32)
        self.cmd_thrusts = torch.zeros(N, 4, device=dev, dtype=torch.float32)
        self.target_states = tgt.to(dev).repeat(N, 1)
        self.actions = torch.zeros(N, A, device=dev, dtype=torch.float32)
        self.pre_actions = torch.zeros(N, A, device=dev, dtype=torch.float32)
        self.ctrl_state = torch.zeros(max(P.ctrl_state_dim, 1), N, device=dev, dtype=torch.float32)
        self._reward_terms = torch.zeros(_capi.AGX_REWARD_TERMS, N, device=dev, dtype=torch.float32) if bk.reward_terms else None
        self._step_dev = torch.zeros(2, device=dev, dtype=torch.int64)  # {global step, CTA ticket}
        self.rng_seed = int(cfg.seed) if int(getattr(cfg, "seed", -1)) >= 0 else int(torch.initial_seed() & 0x7FFFFFFF)
        self.env_offset = 0  # global id of env 0 (multi-GPU shards set this)
        self.item_reward_info = self._make_reward_info()

        io = _capi.AgxStepIO()
        io.state = self.root_states.data_ptr()
        io.actions_out = self.actions.data_ptr()
        io.prev_action = self.pre_actions.data_ptr()
        io.ctrl_state = self.ctrl_state.data_ptr() if P.ctrl_state_dim > 0 else None
        io.progress = self.progress_buf.data_ptr()
        io.reset = self.reset_buf.data_ptr()
        io.timeout = self.time_out_buf.data_ptr()
        io.obs = self.obs_buf.data_ptr()
        io.reward = self.rew_buf.data_ptr()
        io.reset_u8 = self.reset_u8.data_ptr()
        io.cmd = self.cmd_thrusts.data_ptr() if getattr(bk, "export_cmd_thrusts", True) else None
        io.reward_terms = self._reward_terms.data_ptr() if self._reward_terms is not None else None
        io.step_dev = self._step_dev.data_ptr()
        self._io = io

    # ------------------------------------------------------------------------------------------------------
    @property
    def reward_terms_matrix(self):
        """[AGX_REWARD_TERMS, N] planes behind extras["item_reward_info"], rows in REWARD_KEYS order (None when not exported)."""
        return self._reward_terms

    def _make_reward_info(self):
        if self._reward_terms is None:
            return {}
        info = {k: self._reward_terms[i] for i, k in enumerate(self.REWARD_KEYS)}
        if self.ctl_mode not in ("rate", "atti"):
            info["thrust_reward"] = 0  # reference quirk Q7 (hovering.py:451)
        return info

    def set_seed(self, seed: int, env_offset: int = 0):
        self.rng_seed, self.env_offset = int(seed), int(env_offset)

    def step(self, actions, rand_reset=None, rand_noise=None):
        """hovering.py:286-308.  `rand_reset` [N,2,12] U[0,1) / `rand_noise` [N,18] N(0,1) make the randomness
        explicit (parity tests); by default the kernel draws from its Philox stream."""
        if self.counter % 250 == 0 and not self.headless:
            print("self.counter:", self.counter)
        self.counter += 1
        a = actions.to(self._dev)
        if a.dtype != torch.float32 or not a.is_contiguous():
            a = a.to(torch.float32).contiguous()
        if a.shape != (self.num_envs, self.num_actions):
            raise ValueError(f"actions must be [{self.num_envs},{self.num_actions}], got {tuple(a.shape)}")
        io = self._io
        io.action = a.data_ptr()
        io.rand_reset = rand_reset.data_ptr() if rand_reset is not None else None
        io.rand_noise = rand_noise.data_ptr() if rand_noise is not None else None
        io.seed = self.rng_seed
        io.env_offset = self.env_offset
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        _capi.check(self._lib.agx_step(C.byref(self.params), self.num_envs, C.byref(io), C.c_void_p(stream)), "agx_step")
        self.extras["time_outs"] = self.time_out_buf
        self.extras["item_reward_info"] = self.item_reward_info
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def reset_idx(self, env_ids, rand=None):
        """hovering.py:310-335 as a standalone call (agx_reset_idx)."""
        env_ids = env_ids.to(self._dev, torch.long).contiguous()
        m = int(env_ids.numel())
        if m == 0:
            return
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        _capi.check(self._lib.agx_reset_idx(
            C.byref(self.params), self.num_envs, m, env_ids.data_ptr(), self.root_states.data_ptr(),
            self.pre_actions.data_ptr(), self.ctrl_state.data_ptr() if self.params.ctrl_state_dim > 0 else None,
            self.progress_buf.data_ptr(), self.reset_buf.data_ptr(), None, None,
            rand.data_ptr() if rand is not None else None, self.rng_seed, self.counter, self.env_offset,
            C.c_void_p(stream)), "agx_reset_idx")

    # The reference exposes these as separate methods; here they are byproducts of the fused step.
    def pre_physics_step(self, _actions):
        raise RuntimeError("pre_physics_step is fused into Hovering.step (agx_step); call step().")

    def post_physics_step(self):
        return None

    def _observe(self, what, rand_noise=None):
        io = self._io
        if (what & 2) and not io.cmd:
            raise RuntimeError("compute_reward() needs cmd_thrusts (cfg.backend.export_cmd_thrusts = True): the effort term reads them")
        io.rand_noise = rand_noise.data_ptr() if rand_noise is not None else None
        io.seed, io.env_offset = self.rng_seed, self.env_offset
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        _capi.check(self._lib.agx_observe(C.byref(self.params), self.num_envs, C.byref(io), what, C.c_void_p(stream)), "agx_observe")

    def compute_observations(self, rand_noise=None):
        """hovering.py:337-358 as a stand-alone call: obs_buf recomputed from the current root states (+ fresh observation noise,
        or the explicit `rand_noise` [N,18]) by the TASK phase of the step kernel in observe mode (agx_observe)."""
        self._observe(1, rand_noise)
        return self.obs_buf

    def compute_reward(self):
        """hovering.py:360-459 as a stand-alone call: rew_buf, reset_buf (overwritten), item_reward_info and
        pre_actions = actions.clone() recomputed from the current root states, `self.actions` and `cmd_thrusts`."""
        self._observe(2)
        return self.rew_buf
