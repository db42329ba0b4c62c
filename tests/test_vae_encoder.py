"""VAE depth encoder mirror vs latents produced by the reference's own VAE modules
(tests/golden/make_golden_vae.py).  CPU: the encoder is torch modules (cuDNN on a GPU), no libagx kernel involved."""
import os

import numpy as np
import pytest
import torch

from airgym_b200.lib.network.vae_image_encoder import ImgEncoder, VAEImageEncoder
from tests.util_vae import procedural_images, procedural_state

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae_encoder.npz"))
CFG = {"latent_dims": 64, "image_res": [120, 212], "interpolation_mode": "bilinear", "return_sampled_latent": False}


def test_key_layout_matches_reference_checkpoint():
    want = {}
    for row in G["encoder_shapes"]:
        k, s = str(row).split(":")
        want[k[len("encoder."):]] = tuple(int(x) for x in s.split(","))
    have = {k: tuple(v.shape) for k, v in ImgEncoder(1, 64).state_dict().items()}
    assert have == want


def test_missing_weight_file_raises_unless_random_init_is_opted_into():
    with pytest.raises(FileNotFoundError):
        VAEImageEncoder(CFG)  # no model_folder / model_file
    with pytest.raises(FileNotFoundError):
        VAEImageEncoder(dict(CFG, model_folder="/nonexistent", model_file="vae_model.pth"))  # torch.load raises, as the reference's does
    assert VAEImageEncoder(dict(CFG, allow_random_init=True)).latent_dim == 64


def test_procedural_weights_match_reference_latents():
    enc = VAEImageEncoder(dict(CFG, allow_random_init=True))
    shapes = {"encoder." + k: tuple(v.shape) for k, v in enc.encoder.state_dict().items()}
    # the golden generator drew weights for the whole VAE (encoder + decoder), sorted by name: rebuild with the same indices
    full = {str(r).split(":")[0]: tuple(int(x) for x in str(r).split(":")[1].split(",")) for r in G["encoder_shapes"]}
    assert full == shapes
    dec = {"img_decoder.dense.weight": (512, 64), "img_decoder.dense.bias": (512,), "img_decoder.dense1.weight": (11648, 512),
           "img_decoder.dense1.bias": (11648,), "img_decoder.deconv1.weight": (128, 128, 3, 3), "img_decoder.deconv1.bias": (128,),
           "img_decoder.deconv2.weight": (128, 64, 4, 4), "img_decoder.deconv2.bias": (64,), "img_decoder.deconv3.weight": (64, 32, 4, 4),
           "img_decoder.deconv3.bias": (32,), "img_decoder.deconv4.weight": (32, 16, 4, 4), "img_decoder.deconv4.bias": (16,),
           "img_decoder.deconv5.weight": (16, 1, 4, 4), "img_decoder.deconv5.bias": (1,)}
    enc.load_weights(procedural_state({**shapes, **dec}))
    z = enc.encode(procedural_images(4)).numpy()
    assert z.shape == (4, 64)
    assert np.abs(z - G["procedural"]).max() <= 2e-6 + 1e-5 * np.abs(G["procedural"]).max()


@pytest.mark.skipif(not os.path.exists("/root/reference/trained/vae_model.pth"), reason="the 31 MB shipped weight file does not travel")
def test_shipped_weights_match_reference_latents():
    enc = VAEImageEncoder(dict(CFG, model_folder="/root/reference/trained", model_file="vae_model.pth"))
    z = enc.encode(procedural_images(4)).numpy()
    assert np.abs(z - G["shipped"]).max() <= 1e-5 * max(1.0, np.abs(G["shipped"]).max())
