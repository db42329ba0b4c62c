"""A2CAgent — PPO (continuous actions) with the reference's algorithm and entry points
(lib/agent/a2c_continuous.py:37-476 on top of lib/agent/a2c_base.py:78-711), rebuilt around device-resident state:

* rollout buffers are env-major [N, H, ...], so `swap_and_flatten01` (a2c_base.py:26-33) is a free view and minibatches are
  the same contiguous env-major slices the reference's PPODataset cuts (lib/core/datasets.py:29-46);
* GAE, the whole loss forward+backward after the network heads, and grad-scale + clip_grad_norm_ + Adam + the adaptive-KL
  learning-rate rule are single kernels in libagx.so (agx_gae / agx_ppo_loss / agx_adam_step); the learning rate and
  Adam's step counter live on the device, so there is no `.item()` between minibatches and both the rollout (H env steps +
  policy inference) and the update pass replay from CUDA graphs;
* multi-GPU: envs sharded by rank, one NCCL all-reduce per minibatch over [flat grads ‖ loss stats incl. KL], global advantage
  moments and input/value normalisation moments all-reduced so replicas stay identical.
"""
import ctypes as C
import os
import time
from datetime import datetime

import torch
import torch.distributed as dist

from ... import _capi
from ...comm import PeerComm
from ..core.moments import batch_sums, moments_from_sums, sums_from_moments
from ..model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd
from ..utils import tr_helpers, vecenv


def rescale_actions(low, high, action):
    d = (high - low) / 2.0
    m = (high + low) / 2.0
    return action * d + m


class A2CAgent:
    def __init__(self, base_name, params):
        self.name = base_name
        self.network_config = params["network"]
        self.config = config = params["config"]
        self.experiment_name = config.get("full_experiment_name") or config["name"] + datetime.now().strftime("_%d-%H-%M-%S")
        self.multi_gpu = config.get("multi_gpu", False)
        self.local_rank = self.global_rank = 0
        self.world_size = 1
        if self.multi_gpu:
            self.local_rank = int(os.getenv("LOCAL_RANK", "0"))
            self.global_rank = int(os.getenv("RANK", "0"))
            self.world_size = int(os.getenv("WORLD_SIZE", "1"))
            if not dist.is_initialized():
                dist.init_process_group("nccl", rank=self.global_rank, world_size=self.world_size,
                                        device_id=torch.device(f"cuda:{self.local_rank}"))
            config["device"] = f"cuda:{self.local_rank}"
            if self.global_rank != 0:
                config["print_stats"] = False
        self.ppo_device = config.get("device", "cuda:0")
        torch.cuda.set_device(self.ppo_device)
        self.env_config = dict(config.get("env_config", {}))
        self.env_config.setdefault("sim_device", self.ppo_device)
        self.num_actors = config["num_actors"]
        self.env_name = config["env_name"]
        self.vec_env = vecenv.create_vec_env(self.env_name, self.num_actors, **self.env_config)
        self.env = self.vec_env.env
        if self.multi_gpu:  # shard the global env axis: partition-invariant Philox streams
            self.env.set_seed(self.env.rng_seed, env_offset=self.global_rank * self.num_actors)
        self.env_info = self.vec_env.get_env_info()
        self.value_size = 1
        space = self.env_info["observation_space"]
        self.has_cnn = isinstance(space, dict)  # avoid / planning with env_config use_image: True (ppo_avoid/planning.yaml)
        self.obs_shape = {k: v.shape for k, v in space.items()} if self.has_cnn else space.shape
        self.actions_num = self.env_info["action_space"].shape[0]
        dev = self.ppo_device
        self.actions_low = torch.from_numpy(self.env_info["action_space"].low.copy()).float().to(dev)
        self.actions_high = torch.from_numpy(self.env_info["action_space"].high.copy()).float().to(dev)
        self.clip_actions = config.get("clip_actions", True)

        self.save_freq = config.get("save_frequency", 0)
        self.save_best_after = config.get("save_best_after", 100)
        self.print_stats = config.get("print_stats", True)
        self.max_epochs = config.get("max_epochs", -1)
        self.max_frames = config.get("max_frames", -1)
        self.is_adaptive_lr = config.get("lr_schedule") == "adaptive"
        self.kl_threshold = config.get("kl_threshold", 0.008)
        self.e_clip = config["e_clip"]
        if config.get("clip_value", False):
            raise NotImplementedError("clip_value: True is not used by the shipped yamls")
        self.rewards_shaper = config["reward_shaper"]
        self.horizon_length = config["horizon_length"]
        self.normalize_advantage = config["normalize_advantage"]
        self.normalize_input = config["normalize_input"]
        self.normalize_value = config.get("normalize_value", False)
        self.truncate_grads = config.get("truncate_grads", False)
        self.critic_coef = config["critic_coef"]
        self.grad_norm = config["grad_norm"]
        self.gamma, self.tau = config["gamma"], config["tau"]
        self.entropy_coef = config["entropy_coef"]
        self.bounds_loss_coef = config.get("bounds_loss_coef", None)
        self.value_bootstrap = config.get("value_bootstrap")
        self.batch_size = self.horizon_length * self.num_actors
        self.batch_size_envs = self.batch_size
        self.minibatch_size = config.get("minibatch_size", self.num_actors * config.get("minibatch_size_per_env", 0))
        assert self.minibatch_size > 0 and self.batch_size % self.minibatch_size == 0, "batch_size % minibatch_size != 0"
        self.num_minibatches = self.batch_size // self.minibatch_size
        self.mini_epochs_num = config["mini_epochs"]
        self.last_lr = float(config["learning_rate"])
        self.frame = 0
        self.epoch_num = 0
        self.curr_frames = 0
        self.mean_rewards = self.last_mean_rewards = -1000000000
        self.train_dir = config.get("train_dir", "runs")
        self.experiment_dir = os.path.join(self.train_dir, self.experiment_name)
        self.nn_dir = os.path.join(self.experiment_dir, "nn")
        self.summaries_dir = os.path.join(self.experiment_dir, "summaries")
        self.writer = None
        if self.global_rank == 0:
            os.makedirs(self.nn_dir, exist_ok=True)
            if config.get("write_summaries", True):  # rank 0 only, a2c_base.py:262-267
                from torch.utils.tensorboard import SummaryWriter
                os.makedirs(self.summaries_dir, exist_ok=True)
                self.writer = SummaryWriter(self.summaries_dir)
        # camera tasks: the render cadence (every cam_every-th step) is a host-side decision baked into the captured rollout, which is
        # valid as long as every rollout starts at the same phase of it: horizon % cam_every == 0 (24 / 64 vs 4 in the shipped yamls)
        self.use_cuda_graph = config.get("use_cuda_graph", True) and (
            not self.has_cnn or self.horizon_length % max(int(getattr(self.env, "cam_every", 1)), 1) == 0)
        self.fused_mlp = config.get("fused_mlp", True)  # tensor-core MLP kernels (TF32) instead of torch fp32 + autograd
        # Collectives of the sharded update.  "peer" (default): libagx kernels over NVLink peer memory — the
        # per-minibatch all-reduce is fused into the Adam launch (agx_adam_step_allreduce), the per-epoch moments go through
        # agx_comm_allreduce; everything is stream-ordered and captured in the update / dataset graphs.  "nccl": torch.distributed
        # all-reduces, captured in the graphs when graph_collectives is set (default) or run eagerly around them.
        self.comm_kind = config.get("multi_gpu_comm", "peer")
        if self.comm_kind not in ("peer", "nccl"):
            raise ValueError(f"multi_gpu_comm must be 'peer' or 'nccl', got {self.comm_kind!r}")
        self.graph_collectives = bool(config.get("graph_collectives", True))
        self.comm = None
        self.noise_table = None  # optional [H, N, A] N(0,1) draws replacing torch.randn in the rollout (explicit randomness, tests)
        self.algo_observer = config.get("features", {}).get("observer", None)

        keys = {"actions_num": self.actions_num, "input_shape": self.obs_shape, "num_seqs": self.num_actors,
                "value_size": 1, "normalize_value": self.normalize_value, "normalize_input": self.normalize_input}
        self.model = ModelA2CContinuousLogStd(params, keys).to(dev)
        if self.has_cnn:
            # The reference normalises the trunk input under no_grad (base_model.py:29-31), so its CNN never receives a gradient:
            # the encoder is a fixed feature extractor.  The agent therefore encodes each observation ONCE, when it arrives
            # (eval-mode BatchNorm), and everything downstream — rollout buffer, fused MLP kernels, update — works on the
            # 46-wide trunk input [obs16 | cnn(norm(image))] instead of 101 KB images per sample.  (Deviation, DESIGN.md §9: the
            # reference re-encodes the minibatch images in train mode — BatchNorm batch statistics — during the update.)
            self.image_shape = self.obs_shape["image"]
            # static buffers (a captured CUDA graph must find the last render's features / image moments at fixed addresses)
            self._feat = torch.zeros(self.num_actors, self.model.feature_dim, device=dev)
            self._img_mean = torch.zeros(self.image_shape, device=dev, dtype=torch.float64)
            self._img_var = torch.zeros(self.image_shape, device=dev, dtype=torch.float64)
            self._img_valid = False
            self.obs_shape = (self.obs_shape["observation"][0] + self.model.feature_dim,)
        self.flat_params, self.flat_grads = self.model.flatten_parameters(extra_grad_slots=_capi.AGX_PPO_STATS)
        self.n_params = self.model.num_flat
        self.stats = self.flat_grads[self.n_params:]  # loss statistics ride behind the grads through the all-reduce
        self.exp_avg = torch.zeros(self.n_params, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_params, device=dev)
        self.lr_dev = torch.full((1,), self.last_lr, device=dev, dtype=torch.float32)
        self.opt_step = torch.zeros(1, device=dev, dtype=torch.int64)
        self.grad_norm_dev = torch.zeros(1, device=dev)
        self._lib = _capi.load()
        if self.multi_gpu and self.world_size > 1 and self.comm_kind == "peer":
            biggest = self.flat_grads.numel() * 4
            if self.has_cnn:  # merged image moments: 2 x [C,W,H] float64
                biggest = max(biggest, 2 * 8 * int(torch.Size(self.image_shape).numel()))
            self.comm = PeerComm(self.global_rank, self.world_size, biggest, dev)
        self.workspace = torch.zeros(int(self._lib.agx_ppo_workspace_floats()), device=dev)
        hp = _capi.AgxPpoHyper()
        hp.e_clip, hp.critic_coef, hp.entropy_coef = self.e_clip, self.critic_coef, self.entropy_coef
        hp.bounds_loss_coef = self.bounds_loss_coef if self.bounds_loss_coef is not None else 0.0
        hp.kl_threshold = self.kl_threshold
        hp.grad_norm = self.grad_norm if self.truncate_grads else 0.0
        hp.beta1, hp.beta2, hp.eps, hp.weight_decay = 0.9, 0.999, 1e-8, config.get("weight_decay", 0.0)
        hp.adaptive_lr = 1 if self.is_adaptive_lr else 0
        self.hyper = hp
        if self.normalize_value:
            self.value_mean_std = self.model.value_mean_std
        self._graphs = {}
        self.init_tensors()
        # Fused rollout step (agx_policy_step + env.step + agx_rollout_post: 3 launches instead of ~25) where the policy-head
        # epilogue exists (tcgen05 forward of the shipped 64-128-64 network) and the shaper is the plain affine/clamp one.
        sh = self.rewards_shaper
        self.fused_rollout = bool(config.get("fused_rollout", True)) and self.fused_mlp and self.model.policy_step_supported() and (
            isinstance(sh, tr_helpers.DefaultRewardsShaper) and not sh.log_val)
        if self.fused_rollout:
            self._setup_fused_rollout()
        if self.algo_observer is not None:  # a2c_base.py:147-148,255
            self.algo_observer.before_init(base_name, config, self.experiment_name)
            self.algo_observer.after_init(self)

    # ---- buffers --------------------------------------------------------------------------------------------------------
    def init_tensors(self):
        N, H, A, dev = self.num_actors, self.horizon_length, self.actions_num, self.ppo_device
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        self.buf = {"obses": f(N, H, *self.obs_shape), "actions": f(N, H, A), "mus": f(N, H, A), "sigmas": f(N, H, A),
                    "neglogpacs": f(N, H), "values": f(N, H), "rewards": f(N, H),
                    "dones": torch.zeros(N, H, device=dev, dtype=torch.uint8)}
        self.advs, self.returns = f(N, H), f(N, H)
        self.dones = torch.ones(N, device=dev, dtype=torch.uint8)  # a2c_base.py:406
        self.last_values = f(N)
        self.obs = torch.zeros(N, *self.obs_shape, device=dev)
        self.env_actions = f(N, A)
        self.current_rewards, self.current_shaped_rewards, self.current_lengths = f(N), f(N), f(N)
        self.ep_stats = torch.zeros(4, device=dev, dtype=torch.float64)  # Σreward, Σshaped, Σlength, #episodes (per epoch)
        self.term_keys, self.term_sums = None, None  # per-epoch Σ over envs and steps of extras["item_reward_info"] (observer)
        self.grad_mu, self.grad_value, self.grad_logstd = f(self.minibatch_size, A), f(self.minibatch_size), f(A)
        self.norm_values, self.norm_returns, self.advantages = f(N * H, 1), f(N * H, 1), f(N * H)
        if self.fused_mlp:
            mb, dims = self.minibatch_size, self.model.fused_keep_dims()
            self.mlp_ws = self.model.fused_workspace(dev)
            # "tcgen05" (default where it applies: 64-128-64 network, minibatch % 128 == 0): forward, activation-gradient chain and
            # weight gradients on tcgen05.mma with feature-major intermediates; "mma_sync": the warp-level TF32 kernels
            self.mlp_train_tc = self.config.get("mlp_backward", "tcgen05") == "tcgen05" and self.model.train_supported(mb)
            # the PPO loss inside the backward kernel's first stage (agx_ppo_loss_backward_train); False: the stand-alone agx_ppo_loss launch
            self.fuse_loss = bool(self.config.get("fuse_loss", True))
            if self.mlp_train_tc:
                self.keep, self.dz, self.dout = self.model.train_buffers(mb, dev)
            else:
                self.keep = tuple(f(mb, d) for d in dims)           # normalised (padded) input + post-ELU activations
                self.dz = tuple(f(mb, d) for d in dims[1:])         # pre-activation gradients
                self.dout = f(mb, 16)                                # [dmu | dvalue | 0] padded head gradient
            self.mb_mu, self.mb_value = f(mb, A), f(mb)
            self.ro_mu, self.ro_value = f(N, A), f(N)
        self.epoch_loss_sums = torch.zeros(_capi.AGX_PPO_STATS, device=dev)

    # ---- env interaction ----------------------------------------------------------------------------------------------
    def preprocess_actions(self, actions):
        if self.clip_actions:
            return rescale_actions(self.actions_low, self.actions_high, torch.clamp(actions, -1.0, 1.0))
        return actions

    def _trunk_rms(self):
        rms = self.model.running_mean_std
        return rms.running_mean_std["observation"] if self.has_cnn else rms

    def _ingest(self, obs, update_image_rms=True):
        """What the policy trunk sees of an env observation: the vector itself, or [obs16 | cnn(norm(image))]."""
        if not self.has_cnn:
            return obs
        with torch.no_grad():
            self.model.eval()
            # the camera refreshes every cam_every-th step (customized.py:318-321): between renders the image tensor — and with a
            # frozen encoder its features — do not change, so the CNN runs once per render and the image statistics are merged
            # from the cached batch moments on the steps in between (same counts as the reference's per-sample update)
            fresh = not self._img_valid or self.env.counter % self.env.cam_every == 0
            if fresh:
                img = obs["image"]
                var, mean = torch.var_mean(img.reshape(img.shape[0], -1), dim=0)
                self._feat.copy_(self.model.encode_image(img))
                self._img_mean.copy_(mean.reshape(self.image_shape))
                self._img_var.copy_(var.reshape(self.image_shape))
                self._img_valid = True
            feat, mean, var, n = self._feat, self._img_mean, self._img_var, obs["image"].shape[0]
            if self.normalize_input and update_image_rms:
                if self.multi_gpu and self.world_size > 1:  # merge the per-rank moments through their sums
                    s = self._allreduce(sums_from_moments(mean, var, n).contiguous())
                    nt = n * self.world_size
                    gmean, gvar = moments_from_sums(s, nt)
                    self.model.running_mean_std.running_mean_std["image"].update_from_moments(gmean, gvar, nt)
                else:
                    self.model.running_mean_std.running_mean_std["image"].update_from_moments(mean, var, n)
            return torch.cat((obs["observation"], feat), dim=-1)

    def env_reset(self):
        first = self._ingest(self.vec_env.reset(), update_image_rms=False)
        if self.fused_rollout and not self.has_cnn:
            self.obs = first  # the env's own obs_buf (returned by reference, overwritten in place each step): no per-step copy
        else:
            self.obs.copy_(first)
        return self.obs

    def _setup_fused_rollout(self):
        """Argument blocks of the two rollout kernels: one AgxPolicyIO per horizon slot (the rollout-buffer slices differ), one AgxPostIO per slot."""
        b, A, H = self.buf, self.actions_num, self.horizon_length
        p = lambda t: t.data_ptr()
        sh, env = self.rewards_shaper, self.env
        self._pol, self._post = [], []
        for n in range(H):
            io = _capi.AgxPolicyIO()
            io.logstd = p(self.model.logstd)
            if self.normalize_value:
                io.value_mean, io.value_var = p(self.value_mean_std.running_mean), p(self.value_mean_std.running_var)
            io.actions, io.ld_actions = p(b["actions"][:, n]), H * A
            io.mus, io.ld_mus = p(b["mus"][:, n]), H * A
            io.sigmas, io.ld_sigmas = p(b["sigmas"][:, n]), H * A
            io.neglogp, io.ld_neglogp = p(b["neglogpacs"][:, n]), H
            io.values, io.ld_values = p(b["values"][:, n]), H
            io.obs_out, io.ld_obs = p(b["obses"][:, n]), H * self.obs_shape[0]
            io.dones_out, io.ld_dones, io.dones_in = p(b["dones"][:, n]), H, p(self.dones)
            io.env_actions = p(self.env_actions)
            if self.clip_actions:
                io.act_lo, io.act_hi = p(self.actions_low), p(self.actions_high)
            io.seed, io.step_dev, io.env_offset = env.rng_seed, p(env._step_dev), env.env_offset
            self._pol.append(io)
            po = _capi.AgxPostIO()
            po.reward, po.reset_u8, po.timeout = p(env.rew_buf), p(env.reset_u8), p(env.time_out_buf)
            po.values, po.ld_values = p(b["values"][:, n]), H
            po.rewards_out, po.ld_rewards = p(b["rewards"][:, n]), H
            po.cur_reward, po.cur_shaped, po.cur_length = p(self.current_rewards), p(self.current_shaped_rewards), p(self.current_lengths)
            po.dones_state, po.ep_stats = p(self.dones), p(self.ep_stats)
            po.scale, po.shift, po.gamma = float(sh.scale_value), float(sh.shift_value), float(self.gamma)
            po.min_val, po.max_val = max(float(sh.min_val), -3.0e38), min(float(sh.max_val), 3.0e38)
            po.bootstrap = 1 if self.value_bootstrap else 0
            self._post.append(po)

    def _rollout_step_fused(self, n):
        """One env step of play_steps (a2c_base.py:657-695) in three launches: policy (network + sampling + buffer writes), env, post."""
        io = self._pol[n]
        io.noise = self.noise_table[n].data_ptr() if self.noise_table is not None else None
        io.seed, io.env_offset = self.env.rng_seed, self.env.env_offset
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _capi.check(self._lib.agx_policy_step(C.byref(self.model.policy_params()), C.byref(io), self.num_actors, self.obs.data_ptr(), st),
                    "agx_policy_step")
        obs, _rewards, _dones, infos = self.vec_env.step(self.env_actions)
        _capi.check(self._lib.agx_rollout_post(C.byref(self._post[n]), self.num_actors, st), "agx_rollout_post")
        new_obs = self._ingest(obs)
        if new_obs.data_ptr() != self.obs.data_ptr():
            self.obs.copy_(new_obs)
        if self.writer is not None:
            self._observe_reward_terms(infos)

    def _rollout_step(self, n):
        if self.fused_rollout:
            return self._rollout_step_fused(n)
        b = self.buf
        self.model.eval()
        self._ro_step = n
        if self.fused_mlp:
            res = self._fused_policy(self.obs)
        else:
            res = self.model({"is_train": False, "prev_actions": None, "obs": self.obs})
        b["obses"][:, n] = self.obs
        b["dones"][:, n] = self.dones
        b["actions"][:, n] = res["actions"]
        b["neglogpacs"][:, n] = res["neglogpacs"]
        b["values"][:, n] = res["values"].squeeze(-1)
        b["mus"][:, n] = res["mus"]
        b["sigmas"][:, n] = res["sigmas"]
        self.env_actions.copy_(self.preprocess_actions(res["actions"]))
        obs, rewards, dones, infos = self.vec_env.step(self.env_actions)
        shaped = self.rewards_shaper(rewards)
        if self.value_bootstrap and "time_outs" in infos:
            shaped = shaped + self.gamma * res["values"].squeeze(-1) * infos["time_outs"].float()
        b["rewards"][:, n] = shaped
        self.obs.copy_(self._ingest(obs))
        self.dones.copy_(dones)
        # episode statistics, on device (the reference keeps a 100-game window on the host, a2c_base.py:680-695)
        self.current_rewards += rewards
        self.current_shaped_rewards += shaped
        self.current_lengths += 1
        d = self.dones.float()
        self.ep_stats[0] += (self.current_rewards * d).sum()
        self.ep_stats[1] += (self.current_shaped_rewards * d).sum()
        self.ep_stats[2] += (self.current_lengths * d).sum()
        self.ep_stats[3] += d.sum()
        nd = 1.0 - d
        self.current_rewards *= nd
        self.current_shaped_rewards *= nd
        self.current_lengths *= nd
        if self.writer is not None:
            self._observe_reward_terms(infos)

    def _observe_reward_terms(self, infos):
        """RLGPUAlgoObserver.process_infos (lib/utils/isaacgym_utils.py:66-71) without the per-step host list: the reward terms
        are summed into one device vector (graph-safe); after_print_stats' mean over all envs and steps is Σ / (N · steps)."""
        info = infos.get("item_reward_info") if isinstance(infos, dict) else None
        if not info:
            return
        if self.term_keys is None:
            self.term_keys = [k for k, v in info.items() if isinstance(v, torch.Tensor)]
            self.term_zero_keys = [k for k, v in info.items() if not isinstance(v, torch.Tensor)]  # scalar 0 entries (quirk Q7)
            self.term_sums = torch.zeros(len(self.term_keys) + 1, device=self.ppo_device, dtype=torch.float64)
            keys = list(getattr(self.vec_env.env, "REWARD_KEYS", ()))
            m = getattr(self.vec_env.env, "reward_terms_matrix", None)  # [K, N], rows in REWARD_KEYS order: one reduction per step
            self.term_rows = (torch.tensor([keys.index(k) for k in self.term_keys], device=self.ppo_device)
                              if m is not None and all(k in keys for k in self.term_keys) else None)
        if self.term_rows is not None:
            self.term_sums[:-1] += self.vec_env.env.reward_terms_matrix.sum(1)[self.term_rows]
        else:
            self.term_sums[:-1] += torch.stack([info[k].float().sum() for k in self.term_keys])
        self.term_sums[-1] += 1.0

    def _fused_policy(self, obs):
        """get_action_values (a2c_base.py:357-369) through the fused MLP kernel; sampling/neglogp as in the model's forward."""
        m = self.model
        m.fused_heads(obs, self.ro_mu, self.ro_value)
        mu = self.ro_mu
        logstd = mu * 0.0 + m.logstd
        sigma = torch.exp(logstd)
        noise = self.noise_table[self._ro_step] if self.noise_table is not None else torch.randn_like(mu)
        action = mu + sigma * noise
        return {"neglogpacs": m.neglogp(action, mu, sigma, logstd), "values": m.denorm_value(self.ro_value.unsqueeze(-1)),
                "actions": action, "mus": mu, "sigmas": sigma}

    def _rollout(self):
        for n in range(self.horizon_length):
            self._rollout_step(n)
        self.model.eval()
        if self.fused_mlp:
            self.model.fused_heads(self.obs, self.ro_mu, self.ro_value)
            value = self.ro_value.unsqueeze(-1)
        else:
            _, value = self.model.heads(self.obs)
        self.last_values.copy_(self.model.denorm_value(value).squeeze(-1))
        b = self.buf
        st = torch.cuda.current_stream().cuda_stream
        _capi.check(self._lib.agx_gae(self.num_actors, self.horizon_length, self.gamma, self.tau, b["rewards"].data_ptr(),
                                      b["values"].data_ptr(), b["dones"].data_ptr(), self.last_values.data_ptr(),
                                      self.dones.data_ptr(), self.advs.data_ptr(), self.returns.data_ptr(), C.c_void_p(st)), "agx_gae")

    def play_steps(self):
        """a2c_base.py:651-711: H policy+env steps, bootstrap value, GAE — one CUDA-graph replay."""
        with torch.no_grad():
            replayed = self._run_graphed("rollout", self._rollout)
        if replayed == "replayed":  # the env's host-side step counter (render cadence, logging) did not run during the replay
            self.env.counter += self.horizon_length

    # ---- dataset ----------------------------------------------------------------------------------------------------------
    def _allreduce(self, t):
        """In-place SUM over the ranks (no-op on one rank)."""
        if self.multi_gpu and self.world_size > 1:
            if self.comm is not None and t.dtype in (torch.float32, torch.float64) and t.numel() * t.element_size() <= self.comm.slot_bytes:
                self.comm.all_reduce(t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def _rms_update(self, rms, x):
        """RunningMeanStd train-mode update; with >1 rank the batch moments are merged across ranks first so that
        replicas keep identical statistics (the reference keeps per-rank statistics)."""
        n = x.shape[0]
        if (x.dim() == 2 and x.shape[1] <= 128 and x.dtype == torch.float32 and x.stride(1) == 1 and rms.running_mean.dim() == 1
                and self.config.get("fused_rms", True)):
            # two libagx launches (column sums → [all-reduce] → merge) instead of ~12 torch kernels per update
            k, st = x.shape[1], C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if not hasattr(self, "_sums_ws"):
                self._sums_ws = torch.zeros(int(self._lib.agx_col_sums_workspace_doubles()), device=x.device, dtype=torch.float64)
                self._sums = torch.zeros(2 * 128, device=x.device, dtype=torch.float64)
            sums = self._sums[: 2 * k]
            _capi.check(self._lib.agx_col_sums(x.data_ptr(), n, k, x.stride(0), sums.data_ptr(), self._sums_ws.data_ptr(), st), "agx_col_sums")
            if self.multi_gpu and self.world_size > 1:
                self._allreduce(sums)
                n = n * self.world_size
            _capi.check(self._lib.agx_rms_merge(sums.data_ptr(), k, float(n), rms.running_mean.data_ptr(), rms.running_var.data_ptr(),
                                                rms.count.data_ptr(), st), "agx_rms_merge")
            return
        if self.multi_gpu and self.world_size > 1:
            n = n * self.world_size
            mean, var = moments_from_sums(self._allreduce(batch_sums(x)), n)
            mean, var = mean.reshape(rms.running_mean.shape), var.reshape(rms.running_mean.shape)
        else:
            var, mean = torch.var_mean(x.double(), dim=0)
        rms.update_from_moments(mean, var, n)

    def _prepare_dataset(self):
        """a2c_continuous.py:140-177"""
        values, returns = self.buf["values"].view(-1, 1), self.returns.view(-1, 1)
        adv = (returns - values).sum(dim=1)
        if self.normalize_value:
            vms = self.value_mean_std
            vms.eval()
            self._rms_update(vms, values)
            self.norm_values.copy_(vms(values))
            self._rms_update(vms, returns)
            self.norm_returns.copy_(vms(returns))
        else:
            self.norm_values.copy_(values)
            self.norm_returns.copy_(returns)
        if self.normalize_advantage:
            if self.multi_gpu and self.world_size > 1:  # global moments (north_star: all-reduce of Σadv, Σadv², n)
                n = adv.numel() * self.world_size
                mean, var = moments_from_sums(self._allreduce(batch_sums(adv)), n)
                adv = (adv - mean.float()) / (torch.sqrt(var).float() + 1e-8)
            else:
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        self.advantages.copy_(adv)

    def _graph_ok_with_collectives(self):
        return self.use_cuda_graph and (self.comm is not None or self.graph_collectives or not (self.multi_gpu and self.world_size > 1))

    def prepare_dataset(self):
        with torch.no_grad():
            if self._graph_ok_with_collectives():
                self._run_graphed("dataset", self._prepare_dataset)
            else:
                self._prepare_dataset()

    # ---- update ---------------------------------------------------------------------------------------------------------
    def _minibatch(self, i, update_rms):
        """calc_gradients (a2c_continuous.py:299-369) + trancate_gradients_and_step (a2c_base.py:293-316) + LR rule."""
        mb, A = self.minibatch_size, self.actions_num
        sl = slice(i * mb, (i + 1) * mb)
        b = self.buf
        obs = b["obses"].view(-1, *self.obs_shape)[sl]
        if self.normalize_input and update_rms:
            with torch.no_grad():
                self._rms_update(self._trunk_rms(), obs)
        self.model.eval()  # statistics are updated explicitly above; forward only normalises
        if self.fused_mlp:
            if self.mlp_train_tc:
                self.model.fused_heads_train(obs, self.mb_mu, self.mb_value, self.keep)
            else:
                self.model.fused_heads(obs, self.mb_mu, self.mb_value, keep=self.keep)
            mu, value = self.mb_mu, self.mb_value
        else:
            mu, value = self.model.heads(obs)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: t.data_ptr()
        if self.fused_mlp and self.mlp_train_tc and self.fuse_loss:
            # the loss lives in the first stage of the tcgen05 backward (agx_ppo_loss_backward_train): one launch less per minibatch
            lio = _capi.AgxLossIO()
            lio.mu, lio.logstd, lio.value = p(mu), p(self.model.logstd), p(value)
            lio.actions, lio.old_neglogp = p(b["actions"].view(-1, A)[sl]), p(b["neglogpacs"].view(-1)[sl])
            lio.adv, lio.returns = p(self.advantages[sl]), p(self.norm_returns[sl])
            lio.old_mu, lio.old_sigma = p(b["mus"].view(-1, A)[sl]), p(b["sigmas"].view(-1, A)[sl])
            # grad_logstd lands straight in logstd.grad (a view of the flat gradient buffer): no copy kernel per minibatch
            lio.grad_logstd, lio.stats, lio.workspace, lio.a = p(self.model.logstd.grad), p(self.stats), p(self.workspace), A
            self.model.fused_loss_backward_train(self.hyper, lio, self.keep, self.dz, self.dout, self.mlp_ws)
        else:
            _capi.check(self._lib.agx_ppo_loss(
                C.byref(self.hyper), mb, A, p(mu), p(self.model.logstd), p(value), p(b["actions"].view(-1, A)[sl]),
                p(b["neglogpacs"].view(-1)[sl]), p(self.advantages[sl]), p(self.norm_returns[sl]), p(b["mus"].view(-1, A)[sl]),
                p(b["sigmas"].view(-1, A)[sl]), p(self.grad_mu), p(self.grad_value), p(self.grad_logstd), p(self.stats),
                p(self.workspace), st), "agx_ppo_loss")
        if self.fused_mlp and self.mlp_train_tc and self.fuse_loss:
            pass
        elif self.fused_mlp:
            self._manual_backward()
        else:
            self.flat_grads[: self.n_params].zero_()
            torch.autograd.backward((mu, value), (self.grad_mu, self.grad_value.view(-1, 1)))
            self.model.logstd.grad += self.grad_logstd
        if self.comm is not None:  # all-reduce of [grads ‖ stats (KL)] fused into the Adam launch, over NVLink peer memory
            scale = 1.0 / self.world_size
            _capi.check(self._lib.agx_adam_step_allreduce(
                C.byref(self.hyper), C.byref(self.comm.c), self.n_params, _capi.AGX_PPO_STATS, p(self.flat_params), p(self.flat_grads),
                p(self.exp_avg), p(self.exp_avg_sq), p(self.lr_dev), p(self.opt_step), p(self.stats[4:5]), scale, p(self.grad_norm_dev), st),
                "agx_adam_step_allreduce")
            self.epoch_loss_sums += self.stats * scale
            return
        scale = 1.0
        if self.multi_gpu and self.world_size > 1:
            self._allreduce(self.flat_grads)  # grads ‖ stats (KL) in one message
            scale = 1.0 / self.world_size
        self.epoch_loss_sums += self.stats * scale
        _capi.check(self._lib.agx_adam_step(
            C.byref(self.hyper), self.n_params, p(self.flat_params), p(self.flat_grads), p(self.exp_avg), p(self.exp_avg_sq),
            p(self.lr_dev), p(self.opt_step), p(self.stats[4:5]), scale, p(self.grad_norm_dev), st), "agx_adam_step")

    def _manual_backward(self):
        """Backward of the MLP without autograd: activation-gradient chain, split-K weight gradients and their deterministic
        reduction (three launches) write straight into the flat gradient buffer (every parameter's .grad is a view of it)."""
        if self.mlp_train_tc:
            self.model.fused_backward_train(self.grad_mu, self.grad_value, self.keep, self.dz, self.dout, self.mlp_ws)
        else:
            self.model.fused_backward(self.grad_mu, self.grad_value, self.keep, self.dz, self.dout, self.mlp_ws)
        self.model.logstd.grad.copy_(self.grad_logstd)

    def _update_pass(self, update_rms):
        for i in range(self.num_minibatches):
            self._minibatch(i, update_rms)

    def _run_graphed(self, key, fn, *args):
        """Replay `fn` from a CUDA graph (captured on first use after one eager warm-up run)."""
        if not self.use_cuda_graph:
            return fn(*args)
        g = self._graphs.get(key)
        if g is None:  # first call: eager on a side stream (the warm-up cuBLAS / autograd need before capture)
            self._graphs[key] = "warm"
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = fn(*args)
            torch.cuda.current_stream().wait_stream(side)
            return out
        if g == "warm":
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            # with a process group alive, the NCCL watchdog thread's event queries must not invalidate this thread's capture
            mode = "thread_local" if (self.multi_gpu and self.world_size > 1) else "global"
            with torch.cuda.graph(graph, capture_error_mode=mode):
                fn(*args)
            self._graphs[key] = graph
            graph.replay()
            return None
        g.replay()
        return "replayed"

    def train_epoch(self):
        """a2c_continuous.py:78-138"""
        t0 = time.time()
        self.ep_stats.zero_()
        if self.term_sums is not None:
            self.term_sums.zero_()
        self.play_steps()
        torch.cuda.synchronize()
        t1 = time.time()
        self.prepare_dataset()
        self.epoch_loss_sums.zero_()
        graph_update = self._graph_ok_with_collectives()
        for mini_ep in range(self.mini_epochs_num):
            first = mini_ep == 0  # input statistics are only updated during the first mini-epoch (:130-131)
            if graph_update:
                self._run_graphed("update_rms" if first else "update", self._update_pass, first)
            else:
                self._update_pass(first)
        torch.cuda.synchronize()
        t2 = time.time()
        self.last_lr = float(self.lr_dev.item())
        if self.comm is not None:
            self.comm.check()  # a collective that timed out on a missing peer raises here instead of training on garbage
        return t1 - t0, t2 - t1, t2 - t0

    # ---- driver -----------------------------------------------------------------------------------------------------------
    def sync_replicas(self):
        """Rank 0's parameters and normalisation statistics to every rank (a2c_continuous.py:188-192) — one flat-tensor
        broadcast instead of a pickled state_dict."""
        if not (self.multi_gpu and self.world_size > 1):
            return
        dist.broadcast(self.flat_params, 0)
        rmss = [getattr(self.model, "value_mean_std", None)]
        if self.normalize_input:
            rmss += list(self.model.running_mean_std.running_mean_std.values()) if self.has_cnn else [self.model.running_mean_std]
        for rms in rmss:
            if rms is not None:
                for t in (rms.running_mean, rms.running_var, rms.count):
                    dist.broadcast(t, 0)

    def train(self):
        """a2c_continuous.py:179-294"""
        self.env_reset()
        self.sync_replicas()
        total_time = 0.0
        self.history = []
        while True:
            self.epoch_num += 1
            play_time, update_time, sum_time = self.train_epoch()
            total_time += sum_time
            curr_frames = self.batch_size * self.world_size
            self.frame += curr_frames
            ep = self.ep_stats.clone()
            if self.multi_gpu and self.world_size > 1:
                self._allreduce(ep)
            ep = ep.tolist()
            n_mb = self.num_minibatches * self.mini_epochs_num
            losses = (self.epoch_loss_sums / n_mb).tolist()
            if ep[3] > 0:
                self.mean_rewards = ep[0] / ep[3]
            rec = {"epoch": self.epoch_num, "frame": self.frame, "fps_step_inference": curr_frames / play_time,
                   "fps_total": curr_frames / sum_time, "play_time": play_time, "update_time": update_time,
                   "mean_reward": self.mean_rewards if ep[3] > 0 else None, "mean_length": ep[2] / ep[3] if ep[3] > 0 else None,
                   "episodes": ep[3], "a_loss": losses[0], "c_loss": losses[1], "entropy": losses[2], "b_loss": losses[3],
                   "kl": losses[4], "lr": self.last_lr}
            self.history.append(rec)
            should_exit = False
            if self.writer is not None:
                self.write_stats(rec, total_time, ep)
                if self.algo_observer is not None:
                    self.algo_observer.after_print_stats(self.frame, self.epoch_num, total_time)
            if self.global_rank == 0:
                if self.print_stats:
                    print(f"fps step and policy inference: {rec['fps_step_inference']:.0f} fps total: {rec['fps_total']:.0f} "
                          f"epoch: {self.epoch_num}/{self.max_epochs} frames: {self.frame} reward: {rec['mean_reward']} "
                          f"kl: {rec['kl']:.5f} lr: {self.last_lr:.2e}", flush=True)
                if ep[3] > 0:
                    name = self.config["name"] + "_ep_" + str(self.epoch_num) + "_rew_" + str(self.mean_rewards)
                    if self.save_freq > 0 and self.epoch_num % self.save_freq == 0:
                        self.save(os.path.join(self.nn_dir, "last_" + name))
                    if self.mean_rewards > self.last_mean_rewards and self.epoch_num >= self.save_best_after:
                        self.last_mean_rewards = self.mean_rewards
                        self.save(os.path.join(self.nn_dir, self.config["name"]))
                        if self.last_mean_rewards > self.config.get("score_to_win", float("inf")):
                            should_exit = True
                if self.max_epochs != -1 and self.epoch_num >= self.max_epochs:
                    self.save(os.path.join(self.nn_dir, "last_" + self.config["name"] + "_ep_" + str(self.epoch_num)))
                    should_exit = True
                if self.max_frames != -1 and self.frame >= self.max_frames:
                    should_exit = True
            if self.multi_gpu and self.world_size > 1:
                flag = torch.tensor([float(should_exit)], device=self.ppo_device)
                dist.broadcast(flag, 0)
                should_exit = bool(flag.item())
            if should_exit:
                return self.last_mean_rewards, self.epoch_num

    def write_stats(self, rec, total_time, ep):
        """TensorBoard scalars under the reference's tags: a2c_base.py:318-336 (performance/losses/info), a2c_continuous.py:
        220-242 (bounds loss, rewards, episode lengths) and the observer's Episode/<reward term> means (isaacgym_utils.py:86-99).
        step_time is not separable from policy inference inside one graph replay, so performance/step_* repeat play_time."""
        w, frame, epoch = self.writer, rec["frame"], rec["epoch"]
        curr_frames = self.batch_size * self.world_size
        w.add_scalar("performance/step_inference_rl_update_fps", rec["fps_total"], frame)
        w.add_scalar("performance/step_inference_fps", rec["fps_step_inference"], frame)
        w.add_scalar("performance/step_fps", curr_frames / rec["play_time"], frame)
        w.add_scalar("performance/rl_update_time", rec["update_time"], frame)
        w.add_scalar("performance/step_inference_time", rec["play_time"], frame)
        w.add_scalar("performance/step_time", rec["play_time"], frame)
        w.add_scalar("losses/a_loss", rec["a_loss"], frame)
        w.add_scalar("losses/c_loss", rec["c_loss"], frame)
        w.add_scalar("losses/entropy", rec["entropy"], frame)
        w.add_scalar("losses/bounds_loss", rec["b_loss"], frame)
        w.add_scalar("info/last_lr", rec["lr"], frame)
        w.add_scalar("info/lr_mul", 1.0, frame)
        w.add_scalar("info/e_clip", self.hyper.e_clip, frame)
        w.add_scalar("info/kl", rec["kl"], frame)
        w.add_scalar("info/epochs", epoch, frame)
        if self.term_sums is not None:
            sums = self.term_sums.tolist()
            denom = max(sums[-1], 1.0) * self.num_actors
            for k, v in zip(self.term_keys, sums[:-1]):
                w.add_scalar("Episode/" + k, v / denom, epoch)
            for k in self.term_zero_keys:
                w.add_scalar("Episode/" + k, 0.0, epoch)
        if ep[3] > 0:
            for tag, val in (("rewards", ep[0] / ep[3]), ("shaped_rewards", ep[1] / ep[3]), ("episode_lengths", ep[2] / ep[3])):
                w.add_scalar(tag + "/step", val, frame)
                w.add_scalar(tag + "/iter", val, epoch)
                w.add_scalar(tag + "/time", val, total_time)

    # ---- checkpoints: reference key layout (a2c_base.py:528-577, torch_ext.py:74-84) ------------------------------------------
    def get_full_state_weights(self):
        names = [n for n, _ in self.model.named_parameters()]
        state, off = {}, 0
        for idx, (n, prm) in enumerate(self.model.named_parameters()):
            k = prm.numel()
            state[idx] = {"step": self.opt_step.clone().float().squeeze(), "exp_avg": self.exp_avg[off:off + k].view_as(prm).clone(),
                          "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(prm).clone()}
            off += (k + 3) // 4 * 4  # parameters sit 16-byte aligned in the flat buffers (model.flatten_parameters)
        groups = [{"lr": self.last_lr, "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": self.hyper.weight_decay,
                   "amsgrad": False, "params": list(range(len(names)))}]
        return {"model": {k: v.clone() for k, v in self.model.state_dict().items()}, "epoch": self.epoch_num, "frame": self.frame,
                "optimizer": {"state": state, "param_groups": groups}, "last_mean_rewards": self.last_mean_rewards, "env_state": None}

    def set_full_state_weights(self, weights, set_epoch=True):
        sd = weights["model"]
        with torch.no_grad():
            self.flat_params.zero_()  # alignment padding between parameters stays zero
            for k, v in self.model.state_dict().items():
                v.copy_(sd[k])  # in place: parameters stay views of the flat buffer
        if set_epoch:
            self.epoch_num, self.frame = weights["epoch"], weights["frame"]
        opt = weights.get("optimizer")
        if opt and opt.get("state"):
            off = 0
            for idx, (_, prm) in enumerate(self.model.named_parameters()):
                k = prm.numel()
                s = opt["state"].get(idx)
                if s is not None:
                    self.exp_avg[off:off + k].copy_(s["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + k].copy_(s["exp_avg_sq"].reshape(-1))
                    self.opt_step.fill_(int(s["step"]))
                off += (k + 3) // 4 * 4
            self.last_lr = float(opt["param_groups"][0]["lr"])
            self.lr_dev.fill_(self.last_lr)
        self.last_mean_rewards = weights.get("last_mean_rewards", -1000000000)

    def save(self, fn):
        torch.save(self.get_full_state_weights(), fn + ".pth")
        return fn + ".pth"

    def restore(self, fn, set_epoch=True):
        self.set_full_state_weights(torch.load(fn, map_location=self.ppo_device, weights_only=False), set_epoch)
