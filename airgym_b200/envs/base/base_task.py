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

        # base_task.py:73-76
        self.obs_buf = torch.zeros(self.num_envs, self.num_obs, device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(self.num_envs, device=dev, dtype=torch.long)
        self.time_out_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.bool)
        self.extras = {}
        self.viewer = None
        self.enable_viewer_sync = False

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
