"""ORACLE (test infrastructure, never on the product path).

Constants of the hot path, each with the reference line it comes from.  The controller gains and the
rigid-body model are builder-defined (the reference delegates them to rlPx4Controller / PhysX, both absent
from /root/reference — SURVEY.md §8c-2, §8c-3); they are written down here once for the oracle and
independently in ``agx_params_default`` (airgym_b200/csrc/agx_step.cu); `tests/test_params.py` asserts
the two agree field by field.
"""
import math
from dataclasses import dataclass, field
from typing import List

CTL_MODES = ("pos", "vel", "atti", "rate", "prop")  # helpers.py:103 ; = PY / LV / CTA / CTBR / SRT
TASKS = ("hovering", "tracking", "balloon", "avoid", "planning")  # envs/__init__.py:5-62


def _limits(task: str, ctl_mode: str):
    """Hovering.__init__ hovering.py:93-123 ; Tracking.__init__ tracking.py:95-123."""
    if ctl_mode == "pos":
        lim = 6.0 if task == "tracking" else 3.0
        return [-lim, -lim, -lim, -6.0], [lim, lim, lim, 6.0]
    if ctl_mode == "vel":
        return [-6.0] * 4, [6.0] * 4
    if ctl_mode == "atti":
        return [-1.0, -1.0, -1.0, -1.0, 0.0], [1.0] * 5
    if ctl_mode == "rate":  # Customized family: +-1 rad/s (customized.py:109-113)
        r = 1.0 if task in ("balloon", "avoid", "planning") else 6.0
        return [-r, -r, -r, 0.0], [r, r, r, 1.0]
    if ctl_mode == "prop":
        return [0.0] * 4, [1.0] * 4
    raise ValueError(f"unknown ctl_mode {ctl_mode!r}")


@dataclass
class QuadSpec:
    task: str = "hovering"
    ctl_mode: str = "rate"
    integrator: str = "rk4"            # "rk4" | "euler"
    ctrl_reset: bool = False           # reference never resets the controller objects
    no_noise: bool = False
    dt: float = 0.01                   # hovering_config.py:29
    gravity: float = 9.81              # hovering_config.py:31
    episode_length_s: float = 24.0     # hovering_config.py:17 (tracking_config.py:17 → 36)
    # X152b URDF (assets/robots/X152b/model.urdf:19-24,36-39,86-105)
    m_base: float = 0.585
    m_prop: float = 0.004
    arm: float = 0.05374
    prop_z: float = 0.024
    i_base: float = 0.04
    i_prop: float = 1e-6
    k_thrust: float = 9.59             # hovering.py:256
    k_torque: float = 0.2              # hovering.py:270
    max_lin_vel: float = 100.0         # assets/__init__.py:34-35
    max_ang_vel: float = 100.0
    # PX4 defaults (builder-defined)
    rate_p: List[float] = field(default_factory=lambda: [0.15, 0.15, 0.2])
    rate_i: List[float] = field(default_factory=lambda: [0.2, 0.2, 0.1])
    rate_d: List[float] = field(default_factory=lambda: [0.003, 0.003, 0.0])
    rate_int_lim: float = 0.3
    rate_i_fade: float = math.radians(400.0)
    att_p: List[float] = field(default_factory=lambda: [6.5, 6.5, 2.8])
    att_yaw_w: float = 0.4
    att_rate_lim: List[float] = field(default_factory=lambda: [math.radians(220.0), math.radians(220.0), math.radians(200.0)])
    vel_p: List[float] = field(default_factory=lambda: [1.8, 1.8, 4.0])
    vel_i: List[float] = field(default_factory=lambda: [0.4, 0.4, 2.0])
    vel_d: List[float] = field(default_factory=lambda: [0.2, 0.2, 0.0])
    vel_int_lim: List[float] = field(default_factory=lambda: [1.0, 1.0, 2.0])
    pos_p: List[float] = field(default_factory=lambda: [0.95, 0.95, 1.0])
    vel_sp_lim: List[float] = field(default_factory=lambda: [6.0, 6.0, 6.0])
    tilt_max_tan: float = 1.0
    thr_min: float = 0.0
    thr_max: float = 1.0
    target_state: List[float] = field(  # hovering_config.py:12
        default_factory=lambda: [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    )
    noise_sigma: List[float] = field(default_factory=lambda: [1e-3, 5e-3, 2e-2, 4e-1])  # hovering.py:350-353

    def __post_init__(self):
        if self.episode_length_s == 24.0:  # *_config.py episode_length_s
            self.episode_length_s = {"tracking": 36.0, "balloon": 8.0, "avoid": 6.0, "planning": 16.0}.get(self.task, 24.0)
        self.act_lo, self.act_hi = _limits(self.task, self.ctl_mode)
        if self.task == "avoid":  # avoid_config.py:11: hover target at z = 1
            self.target_state = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0]

    @property
    def num_actions(self) -> int:
        return 5 if self.ctl_mode == "atti" else 4  # hovering.py:46

    @property
    def num_obs(self) -> int:
        return {"hovering": 18, "tracking": 48, "balloon": 18, "avoid": 16, "planning": 16}[self.task]

    @property
    def max_episode_length(self) -> int:
        return int(self.episode_length_s / self.dt)  # hovering.py:48

    @property
    def mass(self) -> float:
        return self.m_base + 4 * self.m_prop

    @property
    def inertia(self):
        ixx = self.i_base + 4 * (self.i_prop + self.m_prop * (self.arm**2 + self.prop_z**2))
        izz = self.i_base + 4 * (self.i_prop + self.m_prop * (2 * self.arm**2))
        return [ixx, ixx, izz]

    @property
    def hover_thrust(self) -> float:
        return self.mass * 9.81 / (4 * 9.59)

    @property
    def ctrl_state_dim(self) -> int:
        return {"prop": 0, "rate": 6, "atti": 6, "vel": 12, "pos": 12}[self.ctl_mode]

    @property
    def reset_draws(self) -> int:
        return {"balloon": 15, "avoid": 11, "planning": 124}.get(self.task, 12)
