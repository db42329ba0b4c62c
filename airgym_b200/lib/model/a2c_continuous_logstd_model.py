"""Actor-critic with a state-independent log-std — same forward contract and state_dict key layout as the reference's
ModelA2CContinuousLogStd (lib/model/a2c_continuous_logstd_model.py:14-198, lib/network/mlp.py:4-39) for the non-separate
MLP case of the shipped yamls: `logstd`, `actor_mlp.layers.{i}.{weight,bias}`, `mu.*`, `value_head.*`,
`value_mean_std.*`, `running_mean_std.*`.  All trainable parameters are views into ONE flat fp32 buffer (and their grads
into one flat grad buffer), which is what the fused clip+Adam kernel and the NCCL all-reduce operate on."""
import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _capi
from ..core.running_mean_std import RunningMeanStd, RunningMeanStdObs
from ..network.cnn import CNNFeatureExtractor, native_encode
from ..network.vae_image_encoder import VAEImageEncoder

_ACTS = {"elu": F.elu, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "sin": torch.sin}


class MLP(nn.Module):
    def __init__(self, input_size, units, activation):
        super().__init__()
        if activation not in _ACTS:
            raise ValueError(f"Unsupported activation: {activation}")
        self.activation = _ACTS[activation]
        self.layers = nn.ModuleList()
        in_dim = int(input_size)
        for out_dim in units:
            layer = nn.Linear(in_dim, out_dim)  # default torch init for the weight; bias zeroed (mlp.py:28-35)
            nn.init.zeros_(layer.bias)
            self.layers.append(layer)
            in_dim = out_dim

    def forward(self, x):
        for layer in self.layers:
            x = self.activation(layer(x))
        return x


class ModelA2CContinuousLogStd(nn.Module):
    def __init__(self, params, keys):
        super().__init__()
        net = params["network"]
        if net.get("separate", False) or "resnet" in net:
            raise NotImplementedError("built: the non-separate MLP network (ppo_hovering/tracking/balloon.yaml) and the non-separate "
                                      "CNN / VAE networks (ppo_avoid/planning.yaml); separate critics and the resnet encoder are not")
        self.actions_num = keys["actions_num"]
        input_shape = keys["input_shape"]
        self.has_vae = "vae" in net and "cnn" not in net  # a2c_continuous_logstd_model.py:27-35: resnet > cnn > vae
        self.has_cnn = "cnn" in net or self.has_vae        # "has an image encoder in front of the trunk"
        if self.has_cnn != isinstance(input_shape, dict):
            raise ValueError("a `cnn`/`vae` network needs a dict observation space {'image','observation'} (env_config use_image: True) and vice versa")
        self.normalize_value = params["config"].get("normalize_value", False)
        self.normalize_input = params["config"].get("normalize_input", False)
        self.value_size = params["config"].get("value_size", 1)
        units, act = net["mlp"]["units"], net["mlp"]["activation"]
        assert net["space"]["continuous"].get("fixed_sigma", True), "fixed_sigma: True is the only shipped configuration"
        if self.has_vae:  # :33-35 — a frozen, pre-trained encoder: a plain attribute in the reference (its weights are not part of
            enc = VAEImageEncoder(net["vae"])   # the policy's state_dict), so it is kept out of the module registry here too
            self.feature_dim = enc.latent_dim
            enc.encoder_precise = bool(net["vae"].get("encoder_precise", True))  # False: single-pass TF32 (cuDNN's default precision)
            object.__setattr__(self, "actor_enc", enc)
            self.actor_mlp = MLP(input_shape["observation"][0] + self.feature_dim, units, act)
        elif self.has_cnn:  # :30-32
            self.feature_dim = int(net["cnn"]["output_dim"])
            self.actor_cnn = CNNFeatureExtractor(feature_dim=self.feature_dim)
            self.actor_cnn.encoder_precise = bool(net["cnn"].get("encoder_precise", True))  # False: single-pass TF32 (cuDNN's default precision)
            self.actor_mlp = MLP(input_shape["observation"][0] + self.feature_dim, units, act)
        else:
            self.actor_mlp = MLP(input_shape[0], units, act)
        self.mu = nn.Linear(units[-1], self.actions_num)
        self.mu.weight.data.mul_(0.1)
        self.mu.bias.data.mul_(0.0)
        self.logstd = nn.Parameter(torch.zeros(self.actions_num, dtype=torch.float32))
        self.value_head = nn.Linear(units[-1], 1)
        self.value_head.weight.data.mul_(0.1)
        self.value_head.bias.data.mul_(0.0)
        if self.normalize_value:
            self.value_mean_std = RunningMeanStd((self.value_size,))
        if self.normalize_input:
            if self.has_cnn:  # the vector part is normalised AFTER the CNN features are appended (:73-78)
                shapes = {"image": tuple(input_shape["image"]), "observation": (input_shape["observation"][0] + self.feature_dim,)}
                self.running_mean_std = RunningMeanStdObs(shapes)
            else:
                self.running_mean_std = RunningMeanStd(tuple(input_shape))
        self.flat_params = self.flat_grads = None

    # ---- flat parameter / gradient storage ------------------------------------------------------------------------
    def flatten_parameters(self, extra_grad_slots=0):
        """Re-home every parameter (and a pre-allocated grad) as a view of one flat buffer; `extra_grad_slots` floats
        are appended to the grad buffer so scalars (the KL) can ride along in the same all-reduce."""
        ps = list(self.parameters())
        al = lambda k: (k + 3) // 4 * 4  # every parameter starts 16-byte aligned (vector loads / bulk copies in the kernels)
        n = sum(al(p.numel()) for p in ps)
        dev = ps[0].device
        flat = torch.zeros(n, device=dev, dtype=torch.float32)
        grads = torch.zeros(n + extra_grad_slots, device=dev, dtype=torch.float32)
        off = 0
        for p in ps:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view_as(p.data)
            p.grad = grads[off:off + k].view_as(p.data)
            off += al(k)
        self.flat_params, self.flat_grads, self.num_flat = flat, grads, n
        return flat, grads

    # ---- fused tensor-core path (libagx agx_mlp_forward / agx_mlp_backward) -------------------------------------------
    def fused_params(self, train=False):
        """AgxMlpParams pointing at this module's parameter storage (valid as long as the tensors are not re-allocated).
        train: the tcgen05 training path keeps one spare input plane (index in_dim, all ones: the bias-gradient column)."""
        layers = self.actor_mlp.layers
        if len(layers) != 3 or self.actor_mlp.activation is not F.elu:
            raise NotImplementedError("fused MLP kernels cover the shipped [h1,h2,h3]+ELU network")
        P = _capi.AgxMlpParams()
        P.in_dim = layers[0].in_features
        P.in_pad = (P.in_dim + (16 if train else 15)) // 16 * 16
        P.h1, P.h2, P.h3 = layers[0].out_features, layers[1].out_features, layers[2].out_features
        P.actions_num = self.actions_num
        for i, l in enumerate(layers, 1):
            setattr(P, f"w{i}", l.weight.data_ptr())
            setattr(P, f"b{i}", l.bias.data_ptr())
        P.w_mu, P.b_mu = self.mu.weight.data_ptr(), self.mu.bias.data_ptr()
        P.w_value, P.b_value = self.value_head.weight.data_ptr(), self.value_head.bias.data_ptr()
        if self.normalize_input:
            rms = self.running_mean_std.running_mean_std["observation"] if self.has_cnn else self.running_mean_std
            P.in_mean, P.in_var = rms.running_mean.data_ptr(), rms.running_var.data_ptr()
        return P

    def fused_heads(self, obs, mu_out, value_out, keep=None):
        """mu, value (normalised head output) of `obs` in one kernel; `keep` = (xn, h1, h2, h3) buffers for the backward."""
        if getattr(self, "_fused", None) is None:
            self._fused = self.fused_params()
        p = lambda t: t.data_ptr() if t is not None else None
        k = keep if keep is not None else (None, None, None, None)
        # inference calls take the padding of the tcgen05 kernels when the network has one (e.g. the 80-wide VAE trunk input → 96)
        P = self.train_params() if (keep is None and self.policy_step_supported()) else self._fused
        _capi.check(_capi.load().agx_mlp_forward(C.byref(P), obs.shape[0], p(obs), p(mu_out), p(value_out), p(k[0]), p(k[1]),
                                                 p(k[2]), p(k[3]), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "agx_mlp_forward")

    # ---- tcgen05 training path: feature-major intermediates (agx_mlp_forward_train / agx_mlp_backward_train) ----------
    def train_params(self):
        if getattr(self, "_fused_t", None) is None:
            self._fused_t = self.fused_params(train=True)
        return self._fused_t

    def train_supported(self, batch):
        """The tensor-core backward covers the shipped 64-128-64 network with batches that are multiples of 128."""
        try:
            P = self.train_params()
        except NotImplementedError:
            return False
        return batch % 128 == 0 and bool(_capi.load().agx_mlp_train_supported(C.byref(P)))

    def train_buffers(self, batch, device):
        """(keep_t, dz_t, dout_t): [width, batch] planes — xt/h1t/h2t/h3t kept by the forward, dz1t/dz2t/dz3t + doutt scratch."""
        P = self.train_params()
        f = lambda w: torch.zeros(w, batch, device=device, dtype=torch.float32)
        return (f(P.in_pad), f(P.h1), f(P.h2), f(P.h3)), (f(P.h1), f(P.h2), f(P.h3)), f(16)

    def fused_heads_train(self, obs, mu_out, value_out, keep_t):
        p = lambda t: t.data_ptr()
        _capi.check(_capi.load().agx_mlp_forward_train(C.byref(self.train_params()), obs.shape[0], p(obs), p(mu_out), p(value_out),
                                                       p(keep_t[0]), p(keep_t[1]), p(keep_t[2]), p(keep_t[3]),
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "agx_mlp_forward_train")

    def fused_backward_train(self, grad_mu, grad_value, keep_t, dz_t, dout_t, workspace):
        if getattr(self, "_fused_g", None) is None:
            self._fused_g = self.fused_grads()
        p = lambda t: t.data_ptr()
        _capi.check(_capi.load().agx_mlp_backward_train(
            C.byref(self.train_params()), C.byref(self._fused_g), grad_mu.shape[0], p(grad_mu), p(grad_value), p(keep_t[0]), p(keep_t[1]),
            p(keep_t[2]), p(keep_t[3]), p(dz_t[0]), p(dz_t[1]), p(dz_t[2]), p(dout_t), p(workspace),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "agx_mlp_backward_train")

    def fused_loss_backward_train(self, hyper, lio, keep_t, dz_t, dout_t, workspace):
        """agx_ppo_loss_backward_train: the PPO loss of the minibatch (lio: _capi.AgxLossIO) computed inside the first stage of the tcgen05
        backward — loss statistics, grad_logstd, old_mu / old_sigma and every parameter gradient from ONE call (3 launches)."""
        if getattr(self, "_fused_g", None) is None:
            self._fused_g = self.fused_grads()
        p = lambda t: t.data_ptr()
        _capi.check(_capi.load().agx_ppo_loss_backward_train(
            C.byref(hyper), C.byref(lio), C.byref(self.train_params()), C.byref(self._fused_g), keep_t[0].shape[1], p(keep_t[0]), p(keep_t[1]),
            p(keep_t[2]), p(keep_t[3]), p(dz_t[0]), p(dz_t[1]), p(dz_t[2]), p(dout_t), p(workspace),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "agx_ppo_loss_backward_train")

    def policy_params(self):
        """Parameters of the fused rollout step: the training-path padding is always one of the widths the tcgen05 kernel is built for."""
        return self.train_params()

    def policy_step_supported(self):
        try:
            P = self.train_params()
        except NotImplementedError:
            return False
        return bool(_capi.load().agx_mlp_train_supported(C.byref(P)))

    def fused_grads(self):
        """AgxMlpGrads pointing at the parameters' .grad tensors (views of the flat gradient buffer)."""
        G = _capi.AgxMlpGrads()
        for i, l in enumerate(self.actor_mlp.layers, 1):
            setattr(G, f"gw{i}", l.weight.grad.data_ptr())
            setattr(G, f"gb{i}", l.bias.grad.data_ptr())
        G.gw_mu, G.gb_mu = self.mu.weight.grad.data_ptr(), self.mu.bias.grad.data_ptr()
        G.gw_value, G.gb_value = self.value_head.weight.grad.data_ptr(), self.value_head.bias.grad.data_ptr()
        return G

    def fused_keep_dims(self):
        """Column counts of the kept tensors: normalised input (padded), h1, h2, h3."""
        L = self.actor_mlp.layers
        return [(L[0].in_features + 15) // 16 * 16] + [l.out_features for l in L]

    def fused_backward(self, grad_mu, grad_value, keep, dz, dout, workspace):
        """Parameter gradients of the whole MLP for d(loss)/d(mu), d(loss)/d(value): written into the .grad views."""
        if getattr(self, "_fused_g", None) is None:
            self._fused_g = self.fused_grads()
        p = lambda t: t.data_ptr()
        _capi.check(_capi.load().agx_mlp_backward(
            C.byref(self._fused), C.byref(self._fused_g), grad_mu.shape[0], p(grad_mu), p(grad_value), p(keep[0]), p(keep[1]), p(keep[2]),
            p(keep[3]), p(dz[0]), p(dz[1]), p(dz[2]), p(dout), p(workspace), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "agx_mlp_backward")

    def fused_workspace(self, device):
        if getattr(self, "_fused", None) is None:
            self._fused = self.fused_params()
        n = int(_capi.load().agx_mlp_workspace_floats(C.byref(self._fused)))
        try:  # the training path pads the input wider: size for the larger of the two
            n = max(n, int(_capi.load().agx_mlp_workspace_floats(C.byref(self.train_params()))))
        except NotImplementedError:
            pass
        return torch.zeros(n, device=device)

    # ---- forward ----------------------------------------------------------------------------------------------------
    def norm_obs(self, obs):
        return self.running_mean_std(obs) if self.normalize_input else obs

    def denorm_value(self, value):
        return self.value_mean_std(value, denorm=True) if self.normalize_value else value

    def encode_image(self, img):
        """features of the (normalised) depth image: CNN (:141-142) or the frozen VAE encoder's means (:146-147).  In eval mode
        on the GPU the CNN is the libagx kernel, with the image normalisation fused into its load."""
        if (not self.has_vae and self.actor_cnn.native_ok(img)
                and not (self.normalize_input and self.running_mean_std.training)):
            if self.normalize_input:
                rms = self.running_mean_std.running_mean_std["image"]
                return native_encode(self.actor_cnn, img, rms.running_mean.float().reshape(-1),
                                     torch.rsqrt(rms.running_var.float() + rms.epsilon).reshape(-1))
            return native_encode(self.actor_cnn, img)
        if self.normalize_input:
            with torch.no_grad():
                img = self.running_mean_std.running_mean_std["image"](img)
        return self.actor_enc.encode(img) if self.has_vae else self.actor_cnn(img)

    def trunk_input(self, obs):
        """CNN network: [observation | cnn(norm(image))] (:141-145), un-normalised — what the MLP trunk's input
        normalisation (`running_mean_std.observation`) then sees."""
        return torch.cat((obs["observation"], self.encode_image(obs["image"])), dim=-1)

    def _apply(self, fn, *a, **k):  # .to(device) / .cuda(): the unregistered VAE encoder follows the module
        out = super()._apply(fn, *a, **k)
        if getattr(self, "has_vae", False):
            self.actor_enc._apply(fn)
        return out

    def heads(self, obs):
        """mu, value of an observation.  Camera networks take the reference's dict {'image','observation'} (:141-145) or the
        already-encoded trunk input [observation | features] as a tensor (what the agent caches in its rollout buffer)."""
        if self.has_cnn:
            x = self.trunk_input(obs) if isinstance(obs, dict) else obs
            if self.normalize_input:
                with torch.no_grad():
                    xn = self.running_mean_std.running_mean_std["observation"](x.detach())
                # normalisation is applied under no_grad in the reference too (base_model.py:29-31): the CNN gets no gradient
                # through it — the features reach the trunk only through this detached, normalised copy
                x = xn
            h = self.actor_mlp(x)
        else:
            h = self.actor_mlp(self.norm_obs(obs))
        return self.mu(h), self.value_head(h)

    @staticmethod
    def neglogp(x, mean, std, logstd):
        return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * math.log(2.0 * math.pi) * x.size()[-1] + logstd.sum(dim=-1)

    def forward(self, input_dict):
        is_train = input_dict.get("is_train", True)
        mu, value = self.heads(input_dict["obs"])
        logstd = mu * 0.0 + self.logstd
        sigma = torch.exp(logstd)
        if is_train:
            entropy = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum(dim=-1)
            prev_neglogp = self.neglogp(input_dict["prev_actions"], mu, sigma, logstd)
            return {"prev_neglogp": torch.squeeze(prev_neglogp), "values": value, "entropy": entropy, "mus": mu, "sigmas": sigma}
        action = mu + sigma * torch.randn_like(mu)  # Normal(mu, sigma).sample()
        return {"neglogpacs": torch.squeeze(self.neglogp(action, mu, sigma, logstd)), "values": self.denorm_value(value),
                "actions": action, "mus": mu, "sigmas": sigma}
