"""ORACLE / BASELINE (test infrastructure, never on the product path).

`HostLoopRateControl` has the call-site API of the reference's `ParallelRateControl` (hovering.py:112-115,246-250) and the
reference's calling SHAPE: tensors are marshalled to float64 numpy arrays and a single-threaded C loop walks the envs one by
one (rate_ctl_loop.c, the same rate loop + mixer as oracle/px4_controller.py).  bench.py's cpu_baseline swaps it into the oracle
to report the "reference-shaped" number next to the vectorised one; tests check it against the torch restatement.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(os.path.dirname(_HERE), "_build")
_SO = os.path.join(_BUILD, "librate_ctl_loop.so")


class RateCtlGains(C.Structure):
    _fields_ = [("rate_p", C.c_double * 3), ("rate_i", C.c_double * 3), ("rate_d", C.c_double * 3),
                ("rate_i_fade", C.c_double), ("rate_int_lim", C.c_double)]


def build(force=False):
    """gcc -O2 the C loop into oracle/_build/ (git-ignored; travels to the GPU box with the snapshot)."""
    src = os.path.join(_HERE, "rate_ctl_loop.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", _SO, src])
    lib = C.CDLL(_SO)
    lib.rate_ctl_update.argtypes = [C.POINTER(RateCtlGains), C.c_int64] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p, C.c_void_p]
    lib.rate_ctl_update.restype = None
    return lib


class HostLoopRateControl:
    def __init__(self, num_envs, spec):
        self.lib = build()
        self.n = num_envs
        self.g = RateCtlGains((C.c_double * 3)(*spec.rate_p), (C.c_double * 3)(*spec.rate_i), (C.c_double * 3)(*spec.rate_d),
                              spec.rate_i_fade, spec.rate_int_lim)
        self.state = np.zeros((num_envs, 6), np.float64)
        self.q = None

    def reset(self, env_ids):
        self.state[np.asarray(env_ids)] = 0.0

    def set_q_world(self, q_wxyz):  # hovering.py:249: root_quats_cpu.numpy().astype(np.float64)
        self.q = np.ascontiguousarray(q_wxyz.cpu().numpy().astype(np.float64))

    def update(self, actions, angvel, dt):  # hovering.py:250
        a = np.ascontiguousarray(actions.cpu().numpy().astype(np.float64))
        w = np.ascontiguousarray(angvel.cpu().numpy().astype(np.float64))
        cmd = np.empty((self.n, 4), np.float64)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        self.lib.rate_ctl_update(C.byref(self.g), self.n, p(self.q), p(a), p(w), float(dt), p(self.state), p(cmd))
        return torch.tensor(cmd, dtype=actions.dtype)  # `torch.tensor(...)` as at the call site: a copy back into torch
