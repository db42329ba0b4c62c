"""Procedural inputs shared by tests/golden/make_golden_vae.py and tests/test_vae_encoder.py."""
import math

import torch


def procedural_images(n):
    """[n,1,212,120] depth-like images in [0,1] (the env's width-major layout)."""
    u = torch.arange(212, dtype=torch.float32)[None, :, None]
    v = torch.arange(120, dtype=torch.float32)[None, None, :]
    k = torch.arange(1, n + 1, dtype=torch.float32)[:, None, None]
    img = 0.5 + 0.25 * torch.sin(0.05 * k * u + 0.3 * k) * torch.cos(0.07 * v - 0.1 * k) + 0.25 * torch.sin(0.011 * (u + v) * k)
    return img.clamp(0, 1).unsqueeze(1)


POLICY_IMAGE_SCALE = 18.0  # the env's depth images after dump_images span roughly [0, 20] (stored image mean of the checkpoint: 9.3)


def policy_inputs(n=6):
    """Inputs of tests/golden/policy_planning_cnn.npz: procedural depth images in the env's range + a seeded observation batch."""
    g = torch.Generator().manual_seed(7)
    return procedural_images(n) * POLICY_IMAGE_SCALE, torch.randn(n, 16, generator=g) * 0.5


def procedural_state(shapes):
    """Deterministic weights from the parameter name and element index: w = scale * sin(a * i + b), scale ~ 1/sqrt(fan_in)."""
    out = {}
    for j, (name, shape) in enumerate(sorted(shapes.items())):
        n = 1
        for s in shape:
            n *= s
        fan_in = max(1, n // shape[0]) if len(shape) > 1 else 1
        i = torch.arange(n, dtype=torch.float64)
        w = torch.sin(0.61803398875 * (j + 1) * i + 0.37 * (j + 1)) * (1.0 / math.sqrt(fan_in) if len(shape) > 1 else 0.05)
        out[name] = w.reshape(shape).float()
    return out
