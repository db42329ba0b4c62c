"""PpoPlayerContinuous — the reference's `--play` path (lib/agent/players.py:24-435, lib/torch_runner.py:86-90): build the
model from the yaml, restore a checkpoint (the reference's `.pth` layout: {'model': state_dict, 'epoch', 'frame', ...}),
run `games_num` episodes over the vectorised env with deterministic (mu) or sampled actions and report average reward/steps.
A checkpoint of a plain-MLP policy loads into a CNN policy the way the reference allows (players.py:380-428)."""
import torch

from ..model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd
from ..utils import vecenv


def rescale_actions(low, high, action):  # players.py:17-21
    d = (high - low) / 2.0
    m = (high + low) / 2.0
    return action * d + m


class PpoPlayerContinuous:
    def __init__(self, params):
        self.config = config = params["config"]
        self.env_name = config["env_name"]
        self.env_config = dict(config.get("env_config", {}))
        self.player_config = config.get("player", {})
        self.device = self.player_config.get("device_name", config.get("device", "cuda:0"))
        self.env_config.setdefault("sim_device", self.device)
        self.num_actors = self.player_config.get("num_actors", config["num_actors"])
        self.vec_env = vecenv.create_vec_env(self.env_name, self.num_actors, **self.env_config)
        self.env = self.vec_env
        self.env_info = self.vec_env.get_env_info()
        space = self.env_info["observation_space"]
        self.obs_shape = {k: v.shape for k, v in space.items()} if isinstance(space, dict) else space.shape
        self.actions_num = self.env_info["action_space"].shape[0]
        self.actions_low = torch.from_numpy(self.env_info["action_space"].low.copy()).float().to(self.device)
        self.actions_high = torch.from_numpy(self.env_info["action_space"].high.copy()).float().to(self.device)
        self.clip_actions = config.get("clip_actions", True)
        self.games_num = self.player_config.get("games_num", 2000)
        self.is_deterministic = self.player_config.get("deterministic", self.player_config.get("determenistic", True))
        self.print_stats = self.player_config.get("print_stats", True)
        self.max_steps = self.player_config.get("max_steps", 108000 // 4)
        self.normalize_input = config["normalize_input"]
        self.normalize_value = config.get("normalize_value", False)
        keys = {"actions_num": self.actions_num, "input_shape": self.obs_shape, "value_size": 1,
                "normalize_value": self.normalize_value, "normalize_input": self.normalize_input}
        self.model = ModelA2CContinuousLogStd(params, keys).to(self.device)
        self.model.eval()

    # ---- checkpoints ---------------------------------------------------------------------------------------------------
    def restore(self, fn):
        self.set_full_state_weights(torch.load(fn, map_location=self.device, weights_only=False))

    def set_full_state_weights(self, checkpoint):
        w = checkpoint["model"]
        try:
            self.model.load_state_dict(w)
        except RuntimeError:
            if not self.model.has_cnn or any("cnn" in k for k in w):
                raise
            # players.py:387-428: a pretrained MLP policy initialises everything but the CNN
            print("Missing CNN part. Loading Pretrained MLP Model......")
            with torch.no_grad():
                self.model.logstd.copy_(w["logstd"])
            self.model.running_mean_std.running_mean_std["observation"].load_state_dict(
                {k: w["running_mean_std." + k] for k in ("running_mean", "running_var", "count")})
            self.model.value_mean_std.load_state_dict({k: w["value_mean_std." + k] for k in ("running_mean", "running_var", "count")})
            self.model.actor_mlp.load_state_dict({k[len("actor_mlp."):]: v for k, v in w.items() if k.startswith("actor_mlp.")}, strict=False)
            self.model.mu.load_state_dict({"weight": w["mu.weight"], "bias": w["mu.bias"]})
            self.model.value_head.load_state_dict({"weight": w["value_head.weight"], "bias": w["value_head.bias"]})

    # ---- acting ----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_action(self, obs, is_deterministic=False):
        res = self.model({"is_train": False, "prev_actions": None, "obs": obs})
        a = res["mus"] if is_deterministic else res["actions"]
        if self.clip_actions:
            return rescale_actions(self.actions_low, self.actions_high, torch.clamp(a, -1.0, 1.0))
        return a

    def run(self):
        """players.py:204-290 for a vectorised env: play until `games_num` episodes have finished (or max_steps)."""
        obses = self.env.reset()
        n = self.num_actors
        cr = torch.zeros(n, device=self.device)
        steps = torch.zeros(n, device=self.device)
        sum_rewards = sum_steps = 0.0
        games_played = 0
        for _ in range(self.max_steps):
            action = self.get_action(obses, self.is_deterministic)
            obses, r, done, info = self.env.step(action)
            cr += r
            steps += 1
            d = done.bool()
            cnt = int(d.sum())
            if cnt > 0:
                sum_rewards += float(cr[d].sum())
                sum_steps += float(steps[d].sum())
                games_played += cnt
                cr = cr * (~d)
                steps = steps * (~d)
                if games_played >= self.games_num:
                    break
        games_played = max(games_played, 1)
        self.av_reward, self.av_steps, self.games_played = sum_rewards / games_played, sum_steps / games_played, games_played
        if self.print_stats:
            print("av reward:", self.av_reward, "av steps:", self.av_steps)
        return self.av_reward, self.av_steps
