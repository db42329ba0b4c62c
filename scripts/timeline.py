#!/usr/bin/env python
"""Per-CTA phase timeline of the fused step under real graph-replay conditions (tuning only).
Builds build/libagx_tl.so with -DAGX_TIMELINE (host side: `python scripts/timeline.py --build`), then on the GPU:
    AGX_LIB=build/libagx_tl.so python scripts/timeline.py [--pdl 3]
Phases: 0 kernel entry, 1 before griddepcontrol.wait (pdl3) / after it, 2 after wait (pdl3) / loads issued + noise done,
3 state tile arrived, 4 compute + per-env stores issued, 5 exit.
"""
import argparse, ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "build", "libagx_tl.so")

def build():
    import __graft_entry__ as g
    csrc = g.CSRC
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    srcs = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cu"))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DAGX_TIMELINE", "-o", LIB] + srcs)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--n", type=int, default=65536)
    a = ap.parse_args()
    if a.build:
        build(); return
    os.environ["AGX_LIB"] = LIB
    import numpy as np, torch
    from airgym_b200 import _capi
    from airgym_b200.envs.base.hovering import Hovering
    from airgym_b200.envs.base.hovering_config import HoveringCfg
    lib = _capi.load()
    lib.agx_set_option(b"pdl", a.pdl)
    n, reps = a.n, 8
    envs, acts = [], []
    g = torch.Generator(device="cuda").manual_seed(5678)
    for r in range(reps):
        cfg = HoveringCfg(); cfg.env.num_envs, cfg.env.ctl_mode, cfg.seed = n, "rate", 1234 + r
        cfg.backend.reward_terms = False; cfg.backend.export_cmd_thrusts = False; cfg.backend.mutate_input_actions = False
        envs.append(Hovering(cfg, None, None, "cuda:0", True))
        acts.append(torch.rand(n, 4, device="cuda", generator=g) * 2 - 1)
    for i in range(200):
        envs[i % reps].step(acts[i % reps])
    torch.cuda.synchronize()
    chunk = torch.cuda.CUDAGraph()
    with torch.cuda.graph(chunk):
        for r in range(reps):
            envs[r].step(acts[r])
    for _ in range(20):
        chunk.replay()
    torch.cuda.synchronize()
    # the timeline buffer holds the LAST launch (replica 7); its predecessor's end is unknown, so report phases relative to
    # the earliest CTA start, plus the per-launch period from events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        chunk.replay()
    e1.record(); torch.cuda.synchronize()
    period = e0.elapsed_time(e1) * 1e3 / 800
    ncta = (n + 127) // 128
    buf = np.zeros(ncta * 16, np.uint64)
    rc = lib.agx_debug_timeline(buf.ctypes.data_as(C.c_void_p), C.c_int(ncta * 16))
    assert rc == 0, rc
    t = buf.reshape(ncta, 16)
    gt = t[:, 0:12:2].astype(np.int64); ck = t[:, 1:12:2].astype(np.int64); sm = t[:, 14].astype(np.int64)
    t0 = gt[:, 0].min()
    print(f"period {period:.2f} us/launch, pdl={a.pdl}, CTAs={ncta}, SMs used={len(set(sm.tolist()))}, max CTAs/SM={np.bincount(sm).max()}")
    names = ["entry", "pre-wait", "post-wait/noise", "tile arrived", "compute done", "exit"]
    for p in range(6):
        x = (gt[:, p] - t0) / 1e3
        print(f"  phase {p} {names[p]:16s} globaltimer us rel. first entry: min {x.min():6.2f} p50 {np.median(x):6.2f} p90 {np.percentile(x, 90):6.2f} max {x.max():6.2f}")
    for p in range(1, 6):
        d = (ck[:, p] - ck[:, p - 1])
        print(f"  clk {names[p-1]:>16s} -> {names[p]:16s} cycles: p10 {np.percentile(d,10):8.0f} p50 {np.median(d):8.0f} p90 {np.percentile(d,90):8.0f} max {d.max():8.0f}")
    late = gt[:, 0] - t0 > 2000
    print(f"  CTAs entering > 2 us after the first: {late.sum()}")

if __name__ == "__main__":
    main()
