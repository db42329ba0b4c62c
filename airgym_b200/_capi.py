"""ctypes binding of libagx.so (include/agx.h).  This is the only place Python touches the C ABI.

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU fallback:
if the shared object is missing or a struct mirror disagrees with the library, importing fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AGX_LIB: tuning runs (scripts/kbench.py) point this at an alternative build of the same sources
LIB_PATH = os.environ.get("AGX_LIB") or os.path.join(_HERE, "libagx.so")

AGX_MAX_ACTIONS = 5
AGX_CTRL_STATE_MAX = 12
AGX_RESET_DRAWS_MAX = 16
AGX_NOISE_DRAWS = 18

TASK_IDS = {"hovering": 0, "tracking": 1, "balloon": 2, "avoid": 3, "planning": 4}
CTL_IDS = {"pos": 0, "vel": 1, "atti": 2, "rate": 3, "prop": 4}
FLAG_MUTATE_ACTIONS, FLAG_CTRL_RESET, FLAG_NO_NOISE, FLAG_RESET_ON_COLLISION = 1, 2, 4, 8
AGX_AUX_MAX = 8
AGX_REWARD_TERMS = 12
AGX_NUM_TREES, AGX_NUM_ASSETS, AGX_ASSET_ROW, AGX_PLANNING_DRAWS = 40, 41, 164, 124
AGX_CAM_W, AGX_CAM_H = 212, 120
PHASE_FUSED, PHASE_PHYSICS, PHASE_TASK = 0, 1, 2
INT_RK4, INT_EULER = 0, 1

_f3 = C.c_float * 3
_f5 = C.c_float * AGX_MAX_ACTIONS


class AgxParams(C.Structure):
    _fields_ = [
        ("task", C.c_int32), ("ctl_mode", C.c_int32), ("num_actions", C.c_int32), ("num_obs", C.c_int32),
        ("integrator", C.c_int32), ("flags", C.c_int32), ("max_episode_length", C.c_int32),
        ("ctrl_state_dim", C.c_int32), ("reset_draws", C.c_int32), ("_pad0", C.c_int32),
        ("dt", C.c_float), ("gravity", C.c_float), ("mass", C.c_float), ("inertia", _f3), ("arm", C.c_float),
        ("k_thrust", C.c_float), ("k_torque", C.c_float), ("max_lin_vel", C.c_float), ("max_ang_vel", C.c_float),
        ("act_lo", _f5), ("act_hi", _f5),
        ("rate_p", _f3), ("rate_i", _f3), ("rate_d", _f3), ("rate_int_lim", C.c_float), ("rate_i_fade", C.c_float),
        ("att_p", _f3), ("att_yaw_w", C.c_float), ("att_rate_lim", _f3),
        ("vel_p", _f3), ("vel_i", _f3), ("vel_d", _f3), ("vel_int_lim", _f3), ("pos_p", _f3), ("vel_sp_lim", _f3),
        ("hover_thrust", C.c_float), ("tilt_max_tan", C.c_float), ("thr_min", C.c_float), ("thr_max", C.c_float),
        ("target", C.c_float * 18), ("target_yaw", C.c_float), ("noise_sigma", C.c_float * 4),
        ("collision_radius", C.c_float), ("_pad1", C.c_float),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class AgxStepIO(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("action", C.c_void_p), ("actions_out", C.c_void_p), ("prev_action", C.c_void_p),
        ("ctrl_state", C.c_void_p), ("progress", C.c_void_p), ("reset", C.c_void_p), ("timeout", C.c_void_p),
        ("obs", C.c_void_p), ("reward", C.c_void_p), ("cmd", C.c_void_p), ("reward_terms", C.c_void_p),
        ("aux", C.c_void_p), ("rand_reset", C.c_void_p), ("rand_noise", C.c_void_p),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("step_dev", C.c_void_p), ("env_offset", C.c_int64),
        ("assets", C.c_void_p), ("trees", C.c_void_p), ("phase", C.c_int32), ("_pad", C.c_int32),
        ("reset_u8", C.c_void_p),
    ]


class AgxRenderIO(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("aux", C.c_void_p), ("assets", C.c_void_p), ("trees", C.c_void_p), ("image", C.c_void_p),
        ("rand_add", C.c_void_p), ("rand_mul", C.c_void_p), ("rand_kern", C.c_void_p),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("env_offset", C.c_int64), ("step_dev", C.c_void_p),
    ]


class AgxPpoHyper(C.Structure):
    _fields_ = [
        ("e_clip", C.c_float), ("critic_coef", C.c_float), ("entropy_coef", C.c_float), ("bounds_loss_coef", C.c_float),
        ("kl_threshold", C.c_float), ("grad_norm", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("weight_decay", C.c_float), ("adaptive_lr", C.c_int32), ("_pad", C.c_int32),
    ]


AGX_PPO_STATS = 8


class AgxLossIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mu", "logstd", "value", "actions", "old_neglogp", "adv", "returns", "old_mu", "old_sigma",
                                          "grad_logstd", "stats", "workspace")] + [("a", C.c_int32), ("_pad", C.c_int32)]


class AgxMlpParams(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("in_pad", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32), ("h3", C.c_int32),
                ("actions_num", C.c_int32)] + [(n, C.c_void_p) for n in (
                    "w1", "b1", "w2", "b2", "w3", "b3", "w_mu", "b_mu", "w_value", "b_value", "in_mean", "in_var")]


class AgxMlpGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("gw1", "gb1", "gw2", "gb2", "gw3", "gb3", "gw_mu", "gb_mu", "gw_value", "gb_value")]


class AgxCnnParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "s1", "t1", "w2", "b2", "s2", "t2", "w3", "b3", "s3", "t3", "wfc", "bfc")] + [
        ("feature_dim", C.c_int32), ("_pad", C.c_int32)]


AGX_IPC_HANDLE_BYTES, AGX_COMM_MAX_RANKS, AGX_F32, AGX_F64 = 64, 8, 0, 1


class AgxComm(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("slot_bytes", C.c_int64), ("region", C.c_void_p * AGX_COMM_MAX_RANKS)]


class AgxPolicyIO(C.Structure):
    _fields_ = [("logstd", C.c_void_p), ("value_mean", C.c_void_p), ("value_var", C.c_void_p),
                ("actions", C.c_void_p), ("ld_actions", C.c_int64), ("mus", C.c_void_p), ("ld_mus", C.c_int64),
                ("sigmas", C.c_void_p), ("ld_sigmas", C.c_int64), ("neglogp", C.c_void_p), ("ld_neglogp", C.c_int64),
                ("values", C.c_void_p), ("ld_values", C.c_int64), ("obs_out", C.c_void_p), ("ld_obs", C.c_int64),
                ("dones_out", C.c_void_p), ("ld_dones", C.c_int64), ("dones_in", C.c_void_p), ("env_actions", C.c_void_p),
                ("act_lo", C.c_void_p), ("act_hi", C.c_void_p), ("noise", C.c_void_p), ("seed", C.c_uint64),
                ("step_dev", C.c_void_p), ("env_offset", C.c_int64)]


class AgxPostIO(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("reset_u8", C.c_void_p), ("timeout", C.c_void_p), ("values", C.c_void_p), ("ld_values", C.c_int64),
                ("rewards_out", C.c_void_p), ("ld_rewards", C.c_int64), ("cur_reward", C.c_void_p), ("cur_shaped", C.c_void_p),
                ("cur_length", C.c_void_p), ("dones_state", C.c_void_p), ("ep_stats", C.c_void_p),
                ("scale", C.c_float), ("shift", C.c_float), ("min_val", C.c_float), ("max_val", C.c_float), ("gamma", C.c_float),
                ("bootstrap", C.c_int32)]


class AgxConvParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
                ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("bias", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("res", C.c_void_p), ("rH", C.c_int32), ("rW", C.c_int32), ("ry0", C.c_int32), ("rx0", C.c_int32), ("rsy", C.c_int32),
                ("rsx", C.c_int32), ("y", C.c_void_p), ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("sy", C.c_int32), ("sx", C.c_int32), ("py", C.c_int32), ("px", C.c_int32),
                ("act", C.c_int32), ("_pad", C.c_int32)]


class AgxConvFirstParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("_p0", C.c_int32),
                ("w", C.c_void_p), ("bias", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("px_mean", C.c_void_p),
                ("px_rstd", C.c_void_p), ("y", C.c_void_p), ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("sy", C.c_int32), ("sx", C.c_int32), ("py", C.c_int32), ("px", C.c_int32),
                ("act", C.c_int32)]


ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2


class AgxError(RuntimeError):
    pass


def bind(lib):
    """Attach prototypes of every symbol include/agx.h declares."""
    lib.agx_version.restype = C.c_int
    lib.agx_error_string.restype = C.c_char_p
    lib.agx_sizeof_params.restype = C.c_int
    lib.agx_sizeof_step_io.restype = C.c_int
    lib.agx_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.agx_params_default.argtypes = [C.POINTER(AgxParams), C.c_int, C.c_int]
    lib.agx_step.argtypes = [C.POINTER(AgxParams), C.c_int64, C.POINTER(AgxStepIO), C.c_void_p]
    lib.agx_observe.argtypes = [C.POINTER(AgxParams), C.c_int64, C.POINTER(AgxStepIO), C.c_int, C.c_void_p]
    lib.agx_reset_idx.argtypes = [
        C.POINTER(AgxParams), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p,
    ]
    lib.agx_render_depth.argtypes = [C.POINTER(AgxParams), C.c_int64, C.POINTER(AgxRenderIO), C.c_void_p]
    lib.agx_sizeof_render_io.restype = C.c_int
    lib.agx_philox_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]
    lib.agx_gae.argtypes = [C.c_int64, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 8
    lib.agx_ppo_workspace_floats.restype = C.c_int64
    lib.agx_ppo_loss.argtypes = [C.POINTER(AgxPpoHyper), C.c_int64, C.c_int] + [C.c_void_p] * 15
    lib.agx_adam_step.argtypes = [C.POINTER(AgxPpoHyper), C.c_int64] + [C.c_void_p] * 7 + [C.c_float, C.c_void_p, C.c_void_p]
    lib.agx_mlp_forward.argtypes = [C.POINTER(AgxMlpParams), C.c_int64] + [C.c_void_p] * 8
    lib.agx_mlp_workspace_floats.argtypes = [C.POINTER(AgxMlpParams)]
    lib.agx_mlp_workspace_floats.restype = C.c_int64
    lib.agx_mlp_backward.argtypes = [C.POINTER(AgxMlpParams), C.POINTER(AgxMlpGrads), C.c_int64] + [C.c_void_p] * 12
    lib.agx_mlp_train_supported.argtypes = [C.POINTER(AgxMlpParams)]
    lib.agx_mlp_forward_train.argtypes = [C.POINTER(AgxMlpParams), C.c_int64] + [C.c_void_p] * 8
    lib.agx_mlp_backward_train.argtypes = [C.POINTER(AgxMlpParams), C.POINTER(AgxMlpGrads), C.c_int64] + [C.c_void_p] * 12
    lib.agx_sizeof_loss_io.restype = C.c_int
    lib.agx_ppo_loss_backward_train.argtypes = [C.POINTER(AgxPpoHyper), C.POINTER(AgxLossIO), C.POINTER(AgxMlpParams), C.POINTER(AgxMlpGrads), C.c_int64] + [C.c_void_p] * 10
    lib.agx_col_sums_workspace_doubles.restype = C.c_int64
    lib.agx_col_sums.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.agx_rms_merge.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.agx_sizeof_policy_io.restype = C.c_int
    lib.agx_sizeof_post_io.restype = C.c_int
    lib.agx_policy_step.argtypes = [C.POINTER(AgxMlpParams), C.POINTER(AgxPolicyIO), C.c_int64, C.c_void_p, C.c_void_p]
    lib.agx_rollout_post.argtypes = [C.POINTER(AgxPostIO), C.c_int64, C.c_void_p]
    lib.agx_sizeof_conv_params.restype = C.c_int
    lib.agx_sizeof_conv_first_params.restype = C.c_int
    lib.agx_conv2d_nhwc.argtypes = [C.POINTER(AgxConvParams), C.c_void_p]
    lib.agx_conv2d_first.argtypes = [C.POINTER(AgxConvFirstParams), C.c_void_p]
    lib.agx_resize_bilinear.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.agx_bn_train.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.agx_pool_fc.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    lib.agx_sizeof_cnn_params.restype = C.c_int
    lib.agx_cnn_encode.argtypes = [C.POINTER(AgxCnnParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p]
    lib.agx_comm_region_bytes.argtypes = [C.c_int, C.c_int64]
    lib.agx_comm_region_bytes.restype = C.c_int64
    lib.agx_comm_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]
    lib.agx_comm_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.agx_comm_close.argtypes = [C.c_void_p]
    lib.agx_comm_free.argtypes = [C.c_void_p]
    lib.agx_comm_allreduce.argtypes = [C.POINTER(AgxComm), C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.agx_comm_status.argtypes = [C.POINTER(AgxComm), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]
    lib.agx_adam_step_allreduce.argtypes = [C.POINTER(AgxPpoHyper), C.POINTER(AgxComm), C.c_int64, C.c_int64] + [C.c_void_p] * 7 + [
        C.c_float, C.c_void_p, C.c_void_p]
    lib.agx_device_numa_node.argtypes = [C.c_int]
    lib.agx_host_alloc_pinned.argtypes = [C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.agx_host_free_pinned.argtypes = [C.c_void_p, C.c_int64]
    return lib


EXPORTS = (
    "agx_version", "agx_error_string", "agx_sizeof_params", "agx_sizeof_step_io", "agx_sizeof_render_io", "agx_render_depth", "agx_set_option",
    "agx_params_default", "agx_step", "agx_observe", "agx_reset_idx", "agx_philox_fill", "agx_gae", "agx_ppo_workspace_floats",
    "agx_ppo_loss", "agx_adam_step", "agx_mlp_forward", "agx_mlp_workspace_floats", "agx_mlp_backward",
    "agx_sizeof_cnn_params", "agx_cnn_encode",
    "agx_sizeof_conv_params", "agx_conv2d_nhwc", "agx_sizeof_conv_first_params", "agx_conv2d_first", "agx_resize_bilinear", "agx_pool_fc", "agx_bn_train",
    "agx_col_sums_workspace_doubles", "agx_col_sums", "agx_rms_merge",
    "agx_sizeof_policy_io", "agx_policy_step", "agx_sizeof_post_io", "agx_rollout_post", "agx_mlp_train_supported", "agx_mlp_forward_train", "agx_mlp_backward_train", "agx_sizeof_loss_io", "agx_ppo_loss_backward_train",
    "agx_comm_region_bytes", "agx_comm_alloc", "agx_comm_open", "agx_comm_close", "agx_comm_free", "agx_comm_allreduce",
    "agx_comm_status", "agx_adam_step_allreduce", "agx_device_numa_node", "agx_host_alloc_pinned", "agx_host_free_pinned",
)

_lib = None


def load():
    """Load libagx.so once; raise (never fall back) if it is absent or ABI-incompatible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). airgym_b200 has no CPU fallback."
        )
    lib = bind(C.CDLL(LIB_PATH))
    if (lib.agx_sizeof_params() != C.sizeof(AgxParams) or lib.agx_sizeof_step_io() != C.sizeof(AgxStepIO)
            or lib.agx_sizeof_render_io() != C.sizeof(AgxRenderIO) or lib.agx_sizeof_cnn_params() != C.sizeof(AgxCnnParams)
            or lib.agx_sizeof_policy_io() != C.sizeof(AgxPolicyIO) or lib.agx_sizeof_post_io() != C.sizeof(AgxPostIO)
            or lib.agx_sizeof_conv_params() != C.sizeof(AgxConvParams) or lib.agx_sizeof_conv_first_params() != C.sizeof(AgxConvFirstParams)):
        raise ImportError("libagx.so struct layout differs from airgym_b200/_capi.py (rebuild the library)")
    _lib = lib
    return lib


def check(code: int, what: str = "agx call"):
    if code != 0:
        raise AgxError(f"{what} failed with code {code}: {load().agx_error_string().decode()}")


def default_params(task: str, ctl_mode: str) -> AgxParams:
    p = AgxParams()
    check(load().agx_params_default(C.byref(p), TASK_IDS[task], CTL_IDS[ctl_mode]), "agx_params_default")
    return p


def pinned_host_tensor(nbytes: int, device_index: int):
    """uint8 CPU tensor over a page-locked buffer placed on the NUMA node of GPU `device_index` (agx_host_alloc_pinned);
    returns (tensor, info) — keep the tensor alive as long as copies are in flight; the buffer lives until process exit."""
    import torch

    lib = load()
    node = int(lib.agx_device_numa_node(device_index))
    ptr, placed = C.c_void_p(), C.c_int(0)
    check(lib.agx_host_alloc_pinned(nbytes, node, C.byref(ptr), C.byref(placed)), "agx_host_alloc_pinned")
    buf = (C.c_ubyte * nbytes).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    return t, {"numa_node": node, "placed": bool(placed.value)}
