"""Balloon — host-side mirror of airgym/envs/task/balloon.py on top of the Customized step semantics
(airgym/envs/base/customized.py:216-344): target-ball reaching, 18-dim obs relative to the ball, guidance/hit reward,
reset on collision.  The per-env task state lives in one [N,8] `aux` tensor the fused kernel reads and writes:
columns 0:3 ball position (`balloon_positions`), 3:6 `pre_root_positions`, 6 `collisions`."""
import torch

from ... import _capi
from ..base.hovering import Hovering


class Balloon(Hovering):
    TASK = "balloon"
    REWARD_KEYS = ("guidance_reward", "hit_reward", "action_smoothness_reward", "effort_reward", "ups_reward", "yaw_reward",
                   "_pad0", "_pad1", "reward")  # balloon.py:217-223 (+ yaw_reward, which the reference computes but does not export)

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        N, dev = self.num_envs, self._dev
        self.aux = torch.zeros(N, _capi.AGX_AUX_MAX, device=dev, dtype=torch.float32)
        self.balloon_positions = self.aux[:, 0:3]
        self.pre_root_positions = self.aux[:, 3:6]
        self.collisions = self.aux[:, 6]
        self._io.aux = self.aux.data_ptr()
        if cfg.env.reset_on_collision:
            self.params.flags |= _capi.FLAG_RESET_ON_COLLISION
        else:
            self.params.flags &= ~_capi.FLAG_RESET_ON_COLLISION

    def _make_reward_info(self):
        info = super()._make_reward_info()
        for k in ("_pad0", "_pad1", "thrust_reward"):
            info.pop(k, None)
        return info

    @property
    def balloon_states(self):
        """[N,13] root state of the ball actor: fixed base, identity orientation (assets/__init__.py:196-224)."""
        s = torch.zeros(self.num_envs, 13, device=self._dev)
        s[:, 0:3] = self.balloon_positions
        s[:, 6] = 1.0
        return s

    @property
    def privileged_obs_buf(self):  # customized.py:78-79: the asset root states
        return self.balloon_states.unsqueeze(1) if self.get_privileged_obs else None

    @privileged_obs_buf.setter
    def privileged_obs_buf(self, value):
        pass

    def reset_idx(self, env_ids, rand=None):
        import ctypes as C

        env_ids = env_ids.to(self._dev, torch.long).contiguous()
        m = int(env_ids.numel())
        if m == 0:
            return
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        _capi.check(self._lib.agx_reset_idx(
            C.byref(self.params), self.num_envs, m, env_ids.data_ptr(), self.root_states.data_ptr(), self.pre_actions.data_ptr(),
            self.ctrl_state.data_ptr() if self.params.ctrl_state_dim > 0 else None, self.progress_buf.data_ptr(),
            self.reset_buf.data_ptr(), self.aux.data_ptr(), None, rand.data_ptr() if rand is not None else None, self.rng_seed,
            self.counter, self.env_offset, C.c_void_p(stream)), "agx_reset_idx")
