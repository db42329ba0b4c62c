"""Depth-image VAE encoder of the planning policy's optional `vae` network block (ppo_planning.yaml:33-39) — the encoder half
of the reference's VAE (lib/network/VAE.py:52-148: ResNet8-style, 9 convolutions with two skip convolutions, two dense layers →
[means | log-variances]) and the inference wrapper (lib/network/vae_image_encoder.py:18-60: resize to `image_res`, return the
means).  Module/key names follow the reference so `trained/vae_model.pth` (`encoder.*`; the `img_decoder.*` half is not needed
for acting and is skipped) loads unchanged.  Convolutions are cuDNN's through torch (library path)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


class ImgEncoder(nn.Module):
    def __init__(self, input_dim=1, latent_dim=64):
        super().__init__()
        self.input_dim, self.latent_dim = input_dim, latent_dim
        c = nn.Conv2d
        self.conv0 = c(input_dim, 32, kernel_size=5, stride=2, padding=2)
        self.conv0_1 = c(32, 32, kernel_size=3, stride=2, padding=2)
        self.conv1_0 = c(32, 32, kernel_size=5, stride=2, padding=1)
        self.conv1_1 = c(32, 64, kernel_size=3, stride=1, padding=1)
        self.conv2_0 = c(64, 64, kernel_size=5, stride=2, padding=2)
        self.conv2_1 = c(64, 128, kernel_size=3, stride=2, padding=1)
        self.conv3_0 = c(128, 128, kernel_size=3, stride=1, padding=1)
        self.conv0_jump_2 = c(32, 64, kernel_size=4, stride=2, padding=1)
        self.conv1_jump_3 = c(64, 128, kernel_size=5, stride=4, padding=(2, 1))
        self.dense0 = nn.Linear(4 * 7 * 128, 512)
        self.dense1 = nn.Linear(512, 2 * latent_dim)
        for m in (self.conv0_1, self.conv1_1, self.conv2_1):  # VAE.py:75-88
            nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain("linear"))
            nn.init.zeros_(m.bias)

    @staticmethod
    def _crop_like(t, ref):  # centre crop of a skip branch to the main branch's size (VAE.py:102-108)
        dh, dw = (t.shape[2] - ref.shape[2]) // 2, (t.shape[3] - ref.shape[3]) // 2
        return t[:, :, dh:dh + ref.shape[2], dw:dw + ref.shape[3]]

    def forward(self, img):
        a = F.elu(self.conv0_1(self.conv0(img)))
        b = self.conv1_1(self.conv1_0(a))
        b = F.elu(b + self._crop_like(self.conv0_jump_2(a), b))
        c = self.conv2_1(self.conv2_0(b))
        c = F.elu(c + self._crop_like(self.conv1_jump_3(b), c))
        x = self.conv3_0(c).flatten(1)
        return self.dense1(F.elu(self.dense0(x)))  # [means (latent_dim) | log-variances (latent_dim)]


class VAEImageEncoder(nn.Module):
    """`encode(images) -> means [N, latent_dims]` (vae_image_encoder.py:34-53); frozen, eval mode."""

    def __init__(self, config, device="cuda:0"):
        super().__init__()
        get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
        self.latent_dim = int(get("latent_dims", 64))
        self.image_res = tuple(get("image_res", (120, 212)))
        self.interpolation_mode = get("interpolation_mode", "bilinear")
        self.return_sampled_latent = bool(get("return_sampled_latent", False))
        self.native = bool(get("native", True))  # False: torch / cuDNN modules (the CPU mirror the goldens are checked against)
        self.encoder = ImgEncoder(1, self.latent_dim)
        folder, fn = get("model_folder"), get("model_file")
        if get("allow_random_init", False):  # explicit opt-in (tests, throughput runs without the weight file): frozen random weights
            if folder and fn and os.path.exists(os.path.join(folder, fn)):
                self.load_weights(torch.load(os.path.join(folder, fn), map_location="cpu", weights_only=False))
        else:  # the reference's torch.load raises on a missing file (vae_image_encoder.py:24-27): a mis-set path must not train on garbage latents
            if not (folder and fn):
                raise FileNotFoundError("vae network block: model_folder / model_file are not set (allow_random_init: True to run without weights)")
            self.load_weights(torch.load(os.path.join(folder, fn), map_location="cpu", weights_only=False))
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    def load_weights(self, state_dict):
        clean = {}
        for k, v in state_dict.items():  # clean_state_dict (vae_image_encoder.py:6-14)
            k = k.replace("module.", "").replace("dronet.", "encoder.")
            if k.startswith("encoder."):
                clean[k[len("encoder."):]] = v
        self.encoder.load_state_dict(clean)

    @torch.no_grad()
    def encode(self, image_tensors):
        if (image_tensors.is_cuda and image_tensors.dtype == torch.float32 and image_tensors.shape[1] == 1
                and self.interpolation_mode == "bilinear" and self.native):
            # libagx: resize + the nine convolutions + two dense layers on the tensor cores (tc_encoders.vae_encode), no cuDNN
            from .tc_encoders import vae_encode
            out = vae_encode(self.encoder, self.image_res, image_tensors, precise=getattr(self, "encoder_precise", True))
        else:
            if tuple(image_tensors.shape[-2:]) != self.image_res:
                image_tensors = F.interpolate(image_tensors, self.image_res, mode=self.interpolation_mode)
            out = self.encoder(image_tensors)
        means, logvar = out[:, :self.latent_dim], out[:, self.latent_dim:]
        if self.return_sampled_latent:
            return means + torch.exp(0.5 * logvar) * torch.randn_like(means)
        return means
