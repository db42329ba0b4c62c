"""The C-ABI library loads and exports every symbol include/agx.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from airgym_b200 import _capi
from oracle import QuadSpec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "agx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agx_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = _capi.load()
    names = _declared_symbols()
    assert set(names) == set(_capi.EXPORTS), (names, _capi.EXPORTS)
    for n in names:
        assert hasattr(lib, n), f"libagx.so does not export {n}"
    assert lib.agx_version() == 200


def test_struct_mirrors_match_library(built):
    lib = _capi.load()
    assert lib.agx_sizeof_params() == C.sizeof(_capi.AgxParams)
    assert lib.agx_sizeof_step_io() == C.sizeof(_capi.AgxStepIO)
    assert lib.agx_sizeof_render_io() == C.sizeof(_capi.AgxRenderIO)


def test_error_paths_return_codes_not_exceptions(built):
    lib = _capi.load()
    p = _capi.AgxParams()
    assert lib.agx_params_default(C.byref(p), 0, 99) == -1
    assert b"ctl_mode" in lib.agx_error_string()
    assert lib.agx_params_default(C.byref(p), 4, 2) == -4  # planning/avoid have no atti mode (the reference cannot run it either)
    assert lib.agx_params_default(C.byref(p), 5, 3) == -1  # unknown task
    assert lib.agx_params_default(C.byref(p), 0, 3) == 0
    io = _capi.AgxStepIO()
    assert lib.agx_step(C.byref(p), 8, C.byref(io), None) == -1  # null buffers
    assert lib.agx_step(C.byref(p), -1, C.byref(io), None) == -1
    rio = _capi.AgxRenderIO()
    assert lib.agx_render_depth(C.byref(p), 8, C.byref(rio), None) == -1  # null buffers
    assert lib.agx_render_depth(None, 8, C.byref(rio), None) == -1
    io.phase = 1
    assert lib.agx_step(C.byref(p), 0, C.byref(io), None) == -1  # hovering has no phases (and null buffers)
    io.phase = 0
    assert lib.agx_set_option(b"pdl", 7) == -1 and lib.agx_set_option(b"pdl", -1) == 0
    assert lib.agx_set_option(b"block", 100) == -1
    assert lib.agx_set_option(b"block", 128) == 0
    with pytest.raises(_capi.AgxError):
        _capi.check(-1, "x")


@pytest.mark.parametrize("task", ["hovering", "tracking", "balloon", "avoid", "planning"])
@pytest.mark.parametrize("mode", ["pos", "vel", "atti", "rate", "prop"])
def test_params_default_equals_oracle_spec(built, task, mode):
    """The constants are written down twice (agx_params_default in C, oracle/spec.py); they must agree."""
    if task in ("avoid", "planning") and mode == "atti":
        pytest.skip("no atti mode for the depth-camera tasks")
    P = _capi.default_params(task, mode)
    s = QuadSpec(task=task, ctl_mode=mode)
    f32 = lambda x: C.c_float(x).value
    assert (P.num_actions, P.num_obs, P.max_episode_length, P.ctrl_state_dim, P.reset_draws) == (
        s.num_actions, s.num_obs, s.max_episode_length, s.ctrl_state_dim, s.reset_draws)
    for name, ref in (("dt", s.dt), ("gravity", s.gravity), ("mass", s.mass), ("arm", s.arm), ("k_thrust", s.k_thrust),
                      ("k_torque", s.k_torque), ("max_lin_vel", s.max_lin_vel), ("max_ang_vel", s.max_ang_vel),
                      ("rate_int_lim", s.rate_int_lim), ("rate_i_fade", s.rate_i_fade), ("att_yaw_w", s.att_yaw_w),
                      ("hover_thrust", s.hover_thrust), ("tilt_max_tan", s.tilt_max_tan), ("thr_min", s.thr_min),
                      ("thr_max", s.thr_max)):
        assert getattr(P, name) == f32(ref), name
    for name, ref in (("inertia", s.inertia), ("rate_p", s.rate_p), ("rate_i", s.rate_i), ("rate_d", s.rate_d),
                      ("att_p", s.att_p), ("att_rate_lim", s.att_rate_lim), ("vel_p", s.vel_p), ("vel_i", s.vel_i),
                      ("vel_d", s.vel_d), ("vel_int_lim", s.vel_int_lim), ("pos_p", s.pos_p), ("vel_sp_lim", s.vel_sp_lim),
                      ("target", s.target_state), ("noise_sigma", s.noise_sigma)):
        assert list(getattr(P, name)) == [f32(x) for x in ref], name
    A = s.num_actions
    assert list(P.act_lo)[:A] == s.act_lo and list(P.act_hi)[:A] == s.act_hi


def test_product_has_no_cpu_path(built):
    """Without a CUDA device the env must refuse to construct (no silent fallback, no oracle import)."""
    import torch

    from airgym_b200.envs import task_registry
    from airgym_b200.utils.helpers import get_args

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU"):
        task_registry.make_env("hovering", get_args(["--ctl_mode", "rate", "--num_envs", "8", "--sim_device", "cpu"]))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        task_registry.make_env("hovering", get_args(["--ctl_mode", "rate", "--num_envs", "8"]))
    src = ""
    for d, _, fs in os.walk(os.path.join(ROOT, "airgym_b200")):
        for f in fs:
            if f.endswith(".py"):
                src += open(os.path.join(d, f)).read()
    assert "import oracle" not in src and "from oracle" not in src and "hostsim" not in src
