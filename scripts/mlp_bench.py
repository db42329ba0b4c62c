#!/usr/bin/env python
"""Micro-benchmark of the fused MLP kernels against the torch fp32 path (CUDA events, graph replay)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from airgym_b200.lib.config import default_ppo_config
from airgym_b200.lib.model.a2c_continuous_logstd_model import ModelA2CContinuousLogStd

def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters // 10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters // 10 * 10)

if __name__ == "__main__":
    OBS, A = 18, 4
    model = ModelA2CContinuousLogStd(default_ppo_config("hovering")["params"], {"actions_num": A, "input_shape": (OBS,)}).cuda()
    model.flatten_parameters(); model.eval()
    out = {}
    for B in (32768, 65536):
        obs = torch.randn(B, OBS, device="cuda")
        mu, val = torch.zeros(B, A, device="cuda"), torch.zeros(B, device="cuda")
        dims = model.fused_keep_dims(); ws = model.fused_workspace("cuda")
        keep = tuple(torch.zeros(B, d, device="cuda") for d in dims)
        dz = tuple(torch.zeros(B, d, device="cuda") for d in dims[1:]); dout = torch.zeros(B, 16, device="cuda")
        gmu, gv = torch.randn(B, A, device="cuda"), torch.randn(B, device="cuda")
        from airgym_b200 import _capi
        dbg = lambda m: _capi.load().agx_set_option(b"mlp_forward", m)
        with torch.no_grad():
            dbg(0)  # legacy mma.sync forward
            out[f"mma_sync_fwd_B{B}_us"] = timeit(lambda: model.fused_heads(obs, mu, val))
            out[f"mma_sync_fwd_keep_B{B}_us"] = timeit(lambda: model.fused_heads(obs, mu, val, keep=keep))
            dbg(2)  # tcgen05 forward also when the activations are kept
            out[f"tcgen05_fwd_keep_B{B}_us"] = timeit(lambda: model.fused_heads(obs, mu, val, keep=keep))
            dbg(2)  # default: tcgen05 always
            out[f"fused_fwd_B{B}_us"] = timeit(lambda: model.fused_heads(obs, mu, val))
            out[f"fused_fwd_keep_B{B}_us"] = timeit(lambda: model.fused_heads(obs, mu, val, keep=keep))
            out[f"fused_bwd_wgrad_B{B}_us"] = timeit(lambda: model.fused_backward(gmu, gv, keep, dz, dout, ws))
            out[f"torch_fwd_B{B}_us"] = timeit(lambda: model.heads(obs))
    print(json.dumps(out))
