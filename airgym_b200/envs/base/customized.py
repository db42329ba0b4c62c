"""Customized — host-side mirror of the camera half of the reference's Customized base (airgym/envs/base/customized.py:308-435):
dict observations {'image', 'observation'}, a depth image refreshed every cam_dt/dt steps, collisions and reset_on_collision.

On a render step the fused kernel is split around the camera exactly where the reference renders (customized.py:318-325):
    agx_step(phase=PHYSICS) → agx_render_depth → agx_step(phase=TASK)
on the other 3 of 4 steps it is one fused launch.  The per-env task state lives in an [N,8] `aux` tensor (agx.h)."""
import ctypes as C
import os

import numpy as np
import torch

from ... import _capi
from .hovering import Hovering

_ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "assets")


class Customized(Hovering):
    TASK = "customized"

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        if cfg.env.ctl_mode == "atti":
            # the reference writes its [N,A] actions into obs[12:16] (avoid.py:226, planning.py:214): A = 5 cannot run there either
            raise ValueError(f"{self.TASK}: ctl_mode 'atti' is not available (observation slot 12:16 holds the 4 actions)")
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        N, dev = self.num_envs, self._dev
        self.aux = torch.zeros(N, _capi.AGX_AUX_MAX, device=dev, dtype=torch.float32)
        self.collisions = self.aux[:, 6]
        self._io.aux = self.aux.data_ptr()
        if cfg.env.reset_on_collision:
            self.params.flags |= _capi.FLAG_RESET_ON_COLLISION
        else:
            self.params.flags &= ~_capi.FLAG_RESET_ON_COLLISION
        cam = cfg.asset_config.include_robot["X152b"]
        self.enable_onboard_cameras = bool(cam.get("enable_onboard_cameras", False))
        self.cam_resolution = (int(cam["width"]), int(cam["height"]))  # customized.py:203-212: (w, h)
        self.cam_channel = int(cam.get("cam_channel", 1))
        if self.cam_resolution != (_capi.AGX_CAM_W, _capi.AGX_CAM_H):
            raise ValueError(f"the depth camera kernel is built for {_capi.AGX_CAM_W}x{_capi.AGX_CAM_H} images")
        self.full_camera_array = torch.zeros(N, self.cam_channel, *self.cam_resolution, device=dev, dtype=torch.float32)
        self.cam_every = int(round(cfg.env.cam_dt / cfg.sim.dt))
        self.counter = 0
        rio = _capi.AgxRenderIO()
        rio.state = self.root_states.data_ptr()
        rio.aux = self.aux.data_ptr()
        rio.image = self.full_camera_array.data_ptr()
        self._rio = rio

    def _make_reward_info(self):
        info = super()._make_reward_info()
        for k in [k for k in info if k.startswith("_pad")]:
            info.pop(k)
        return info

    def render_cameras(self, rand_image=None):
        """customized.py:386-391 + dump_images (:399-435), one launch.  rand_image = {'add','mul','kern'} makes the image
        noise explicit (parity tests)."""
        rio = self._rio
        rio.rand_add = rand_image["add"].data_ptr() if rand_image is not None else None
        rio.rand_mul = rand_image["mul"].data_ptr() if rand_image is not None else None
        rio.rand_kern = rand_image["kern"].data_ptr() if rand_image is not None else None
        rio.seed, rio.step, rio.env_offset = self.rng_seed, self.counter, self.env_offset
        rio.step_dev = self._step_dev.data_ptr()  # Philox counter word from the device (fresh noise on every CUDA-graph replay)
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        _capi.check(self._lib.agx_render_depth(C.byref(self.params), self.num_envs, C.byref(rio), C.c_void_p(stream)),
                    "agx_render_depth")

    def step(self, actions, rand_reset=None, rand_noise=None, rand_image=None):
        """customized.py:308-344 (avoid.py:160-201, planning.py:138-184)."""
        self.counter += 1  # customized.py:219
        a = actions.to(self._dev)
        if a.dtype != torch.float32 or not a.is_contiguous():
            a = a.to(torch.float32).contiguous()
        if a.shape != (self.num_envs, self.num_actions):
            raise ValueError(f"actions must be [{self.num_envs},{self.num_actions}], got {tuple(a.shape)}")
        io = self._io
        io.action = a.data_ptr()
        io.rand_reset = rand_reset.data_ptr() if rand_reset is not None else None
        io.rand_noise = None
        io.seed, io.env_offset = self.rng_seed, self.env_offset
        stream = C.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)
        if self.enable_onboard_cameras and self.counter % self.cam_every == 0:
            io.phase = _capi.PHASE_PHYSICS
            _capi.check(self._lib.agx_step(C.byref(self.params), self.num_envs, C.byref(io), stream), "agx_step(physics)")
            self.render_cameras(rand_image)
            io.phase = _capi.PHASE_TASK
            _capi.check(self._lib.agx_step(C.byref(self.params), self.num_envs, C.byref(io), stream), "agx_step(task)")
        else:
            io.phase = _capi.PHASE_FUSED
            _capi.check(self._lib.agx_step(C.byref(self.params), self.num_envs, C.byref(io), stream), "agx_step")
        self.extras["time_outs"] = self.time_out_buf
        self.extras["item_reward_info"] = self.item_reward_info
        obs = {"image": self.full_camera_array, "observation": self.obs_buf}
        return obs, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _assets_ptr(self):
        return None

    def reset_idx(self, env_ids, rand=None):
        env_ids = env_ids.to(self._dev, torch.long).contiguous()
        m = int(env_ids.numel())
        if m == 0:
            return
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        _capi.check(self._lib.agx_reset_idx(
            C.byref(self.params), self.num_envs, m, env_ids.data_ptr(), self.root_states.data_ptr(), self.pre_actions.data_ptr(),
            self.ctrl_state.data_ptr() if self.params.ctrl_state_dim > 0 else None, self.progress_buf.data_ptr(),
            self.reset_buf.data_ptr(), self.aux.data_ptr(), self._assets_ptr(), rand.data_ptr() if rand is not None else None,
            self.rng_seed, self.counter, self.env_offset, C.c_void_p(stream)), "agx_reset_idx")


def load_tree_table(device):
    """[40,8] cylinder table of the `thin` asset group (scripts/gen_trees.py): slot i = tree_<i>.urdf."""
    t = np.load(os.path.join(_ASSETS, "thin_trees.npy"))[:_capi.AGX_NUM_TREES]
    return torch.from_numpy(t.copy()).to(device)
