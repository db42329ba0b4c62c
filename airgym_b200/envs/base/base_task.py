"""BaseTask — the buffers and `reset()` contract of the reference's BaseTask (airgym/envs/base/base_task.py:38-111)
without IsaacGym: no gym handle, no viewer; the simulator is the fused CUDA step in libagx.so."""
import torch

from ... import _capi


class BaseTask:
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless):
        self.sim_params = sim_params
        self.dt = float(cfg.sim.dt)
        self.physics_engine = physics_engine
        self.sim_device = sim_device if isinstance(sim_device, str) else str(sim_device)
        self.headless = headless
        dev = torch.device(self.sim_device)
        if dev.type != "cuda":
            raise RuntimeError(
                f"airgym_b200 runs on a CUDA device only (got sim_device={sim_device!r}); there is no CPU pipeline."
            )
        if not torch.cuda.is_available():
            raise RuntimeError("airgym_b200: no CUDA device available; the env step has no CPU fallback.")
        self.device = self.sim_device
        self._dev = dev
        self._lib = _capi.load()  # raises if libagx.so is missing

        self.num_envs = cfg.env.num_envs
        self.num_obs = cfg.env.num_observations
        self.get_privileged_obs = cfg.env.get_privileged_obs
        self.num_actions = cfg.env.num_actions

        # base_task.py:73-76.  obs_buf and rew_buf are views into ONE block [obs | rew | reset flags as bytes] (sections padded
        # to 256 B), so a caller that ships a step's results to the host reads all three back with a single copy
        # (`results_block` / `unpack_results`); reset_buf itself stays the reference's int64 tensor.
        N, no = self.num_envs, self.num_obs
        al = lambda b: (b + 255) // 256 * 256
        self._res_off = (0, al(N * no * 4), al(N * no * 4) + al(N * 4))
        self.results_block = torch.zeros(self._res_off[2] + al(N), device=dev, dtype=torch.uint8)
        self.obs_buf = self.results_block[: N * no * 4].view(torch.float32).view(N, no)
        self.rew_buf = self.results_block[self._res_off[1]: self._res_off[1] + N * 4].view(torch.float32)
        self.reset_u8 = self.results_block[self._res_off[2]: self._res_off[2] + N]
        self.reset_buf = torch.ones(self.num_envs, device=dev, dtype=torch.long)
        self.time_out_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.bool)
        self.extras = {}
        self.viewer = None
        self.enable_viewer_sync = False

    def unpack_results(self, block):
        """Views (obs [N,num_obs] f32, rew [N] f32, reset [N] u8) of a copy of `results_block` (e.g. in pinned host memory)."""
        N, no, (o0, o1, o2) = self.num_envs, self.num_obs, self._res_off
        return (block[o0: o0 + N * no * 4].view(torch.float32).view(N, no), block[o1: o1 + N * 4].view(torch.float32),
                block[o2: o2 + N])

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        raise NotImplementedError

    def reset(self):
        """base_task.py:107-111: reset every env, then take one zero-action step."""
        self.reset_idx(torch.arange(self.num_envs, device=self._dev))
        obs, privileged_obs, _, _, _ = self.step(
            torch.zeros(self.num_envs, self.num_actions, device=self._dev, requires_grad=False))
        return obs, privileged_obs

    def step(self, actions):
        raise NotImplementedError

    def render(self, sync_frame_time=True):
        return None  # headless backend
