"""TEST INFRASTRUCTURE: run the host build of agx_math.cuh (tests/hostsim/hostsim.cpp) on numpy buffers."""
import ctypes as C
import os
import subprocess

import numpy as np

from airgym_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "libhostsim.so")


def build(force=False):
    src = os.path.join(_HERE, "hostsim.cpp")
    hdr = os.path.join(_ROOT, "airgym_b200", "csrc", "agx_math.cuh")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + os.path.join(_ROOT, "include"), "-I" + os.path.join(_ROOT, "airgym_b200", "csrc"),
                               "-o", _SO, src])
    lib = C.CDLL(_SO)
    lib.hostsim_step.argtypes = [C.POINTER(_capi.AgxParams), C.c_int64, C.POINTER(_capi.AgxStepIO)]
    lib.hostsim_render.argtypes = [C.POINTER(_capi.AgxParams), C.c_int64, C.POINTER(_capi.AgxRenderIO)]
    lib.hostsim_philox_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int64]
    return lib


_SO_CNN = os.path.join(_HERE, "libhostsim_cnn.so")


def build_cnn(force=False):
    """Host build of the depth-encoder phase functions (airgym_b200/csrc/agx_cnn.cuh) + the schedule replay."""
    src = os.path.join(_HERE, "hostsim_cnn.cpp")
    hdr = os.path.join(_ROOT, "airgym_b200", "csrc", "agx_cnn.cuh")
    if force or not os.path.exists(_SO_CNN) or os.path.getmtime(_SO_CNN) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + os.path.join(_ROOT, "include"), "-I" + os.path.join(_ROOT, "airgym_b200", "csrc"),
                               "-o", _SO_CNN, src])
    lib = C.CDLL(_SO_CNN)
    lib.hostsim_cnn_encode.argtypes = [C.POINTER(_capi.AgxCnnParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    return lib


class HostEnv:
    """Numpy-buffer twin of airgym_b200's env state; `step` runs the host build of the kernel body."""

    def __init__(self, params, n):
        self.lib = build()
        self.P, self.n = params, n
        A, K = params.num_actions, params.ctrl_state_dim
        self.state = np.zeros((n, 13), np.float32); self.state[:, 6] = 1
        self.actions_out = np.zeros((n, A), np.float32)
        self.prev_action = np.zeros((n, A), np.float32)
        self.ctrl_state = np.zeros((max(K, 1), n), np.float32)
        self.progress = np.zeros(n, np.int64)
        self.reset = np.ones(n, np.int64)
        self.timeout = np.zeros(n, np.uint8)
        self.obs = np.zeros((n, params.num_obs), np.float32)
        self.reward = np.zeros(n, np.float32)
        self.cmd = np.zeros((n, 4), np.float32)
        self.terms = np.zeros((_capi.AGX_REWARD_TERMS, n), np.float32)
        self.aux = np.zeros((n, _capi.AGX_AUX_MAX), np.float32)
        self.assets = np.zeros((n, _capi.AGX_ASSET_ROW), np.float32)
        self.trees = np.load(os.path.join(_ROOT, "airgym_b200", "assets", "thin_trees.npy"))[:_capi.AGX_NUM_TREES].copy()
        self.image = np.zeros((n, _capi.AGX_CAM_W, _capi.AGX_CAM_H), np.float32)

    def render(self, rand_add, rand_mul, rand_kern):
        """Host twin of agx_render_depth (explicit noise)."""
        io = _capi.AgxRenderIO()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        self._keep_r = (rand_add, rand_mul, rand_kern)
        io.state, io.aux, io.assets, io.trees, io.image = ptr(self.state), ptr(self.aux), ptr(self.assets), ptr(self.trees), ptr(self.image)
        io.rand_add, io.rand_mul, io.rand_kern = ptr(rand_add), ptr(rand_mul), ptr(rand_kern)
        rc = self.lib.hostsim_render(C.byref(self.P), self.n, C.byref(io))
        assert rc == 0, rc

    def step(self, action, rand_reset=None, rand_noise=None, seed=0, step=0, env_offset=0, phase=0):
        io = _capi.AgxStepIO()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        self._keep = (action, rand_reset, rand_noise)
        io.state, io.action, io.actions_out, io.prev_action = ptr(self.state), ptr(action), ptr(self.actions_out), ptr(self.prev_action)
        io.ctrl_state, io.progress, io.reset, io.timeout = ptr(self.ctrl_state), ptr(self.progress), ptr(self.reset), ptr(self.timeout)
        io.obs, io.reward, io.cmd, io.reward_terms = ptr(self.obs), ptr(self.reward), ptr(self.cmd), ptr(self.terms)
        io.rand_reset, io.rand_noise = ptr(rand_reset), ptr(rand_noise)
        io.aux = ptr(self.aux)
        io.assets, io.trees, io.phase = ptr(self.assets), ptr(self.trees), phase
        io.seed, io.step, io.env_offset = seed, step, env_offset
        rc = self.lib.hostsim_step(C.byref(self.P), self.n, C.byref(io))
        assert rc == 0, rc
