"""ORACLE (test infrastructure, never on the product path).

Stand-in for ``rlPx4Controller.pyParallelControl.Parallel{Pos,Vel,Atti,Rate}Control`` — a third-party
C++/pybind11 dependency that is NOT under /root/reference (cloned from GitHub HEAD with no pinned commit,
reference ``configuration.sh:94-111``; ``setup.py:19`` lists it unversioned).  PARITY UNPINNED: only the
call sites are in-tree (``airgym/envs/base/hovering.py:98-116,235-250``); the reference holds no test or
golden vector for the controller.  What is restated here is the published PX4 cascade the package is named
after (RateControl → AttitudeControl → PositionControl velocity/position loops → quad-X mixer), with PX4's
default gains, in the simulator's FLU/ENU frames (equivalent to PX4's FRD/NED for its diagonal gains).

The object mirrors the call-site API: ``set_status(pos, q_wxyz, linvel, angvel, dt)``,
``set_q_world(q_wxyz)``, ``update(actions[, angvel, dt]) -> cmd[N,4]`` and — like the reference's objects —
keeps per-env integrator state that is never reset on episode resets unless told to.
"""
import torch

from . import rotations as R
from .spec import QuadSpec


def _clamp(x, lo, hi):
    # comparison-based so NaN propagates, like torch.clamp
    return torch.where(x < lo, torch.as_tensor(lo, dtype=x.dtype), torch.where(x > hi, torch.as_tensor(hi, dtype=x.dtype), x))


class ParallelControl:
    """State layout (columns of ``self.state[N,12]``): [0:3] rate integrator, [3:6] previous body rate,
    [6:9] velocity integrator, [9:12] previous world velocity."""

    def __init__(self, num_envs: int, spec: QuadSpec, dtype=torch.float32):
        self.n = num_envs
        self.spec = spec
        self.dtype = dtype
        self.state = torch.zeros(num_envs, 12, dtype=dtype)
        t = lambda v: torch.tensor(v, dtype=dtype)
        self.rate_p, self.rate_i, self.rate_d = t(spec.rate_p), t(spec.rate_i), t(spec.rate_d)
        self.att_p, self.att_rate_lim = t(spec.att_p), t(spec.att_rate_lim)
        self.vel_p, self.vel_i, self.vel_d = t(spec.vel_p), t(spec.vel_i), t(spec.vel_d)
        self.vel_int_lim, self.pos_p, self.vel_sp_lim = t(spec.vel_int_lim), t(spec.pos_p), t(spec.vel_sp_lim)

    def reset(self, env_ids):
        self.state[env_ids] = 0

    # -- call-site API -----------------------------------------------------------------------------------
    def set_status(self, pos, q_wxyz, linvel, angvel, dt):
        self.pos, self.linvel, self.angvel_w, self.dt = pos, linvel, angvel, dt
        self.set_q_world(q_wxyz)

    def set_q_world(self, q_wxyz):
        self.q = q_wxyz[:, [1, 2, 3, 0]]  # xyzw internally
        self.Rm = R.quaternion_to_matrix(q_wxyz)

    def body_rates(self, angvel_w):
        return torch.einsum("nji,nj->ni", self.Rm, angvel_w)  # R^T w

    # -- loops ---------------------------------------------------------------------------------------------
    def mixer(self, thrust, tau):
        T = thrust.unsqueeze(-1)
        tx, ty, tz = tau[:, 0:1], tau[:, 1:2], tau[:, 2:3]
        cmd = torch.cat((T - tx - ty - tz, T + tx + ty - tz, T + tx - ty + tz, T - tx + ty + tz), -1)
        return _clamp(cmd, 0.0, 1.0)

    def rate_loop(self, w_sp, w_b, dt):
        s = self.spec
        integ, prev = self.state[:, 0:3], self.state[:, 3:6]
        e = w_sp - w_b
        wdot = (w_b - prev) / dt
        tau = self.rate_p * e + integ - self.rate_d * wdot
        ef = e / s.rate_i_fade
        fade = 1.0 - ef * ef
        fade = torch.where(fade < 0, torch.zeros_like(fade), fade)
        self.state[:, 0:3] = _clamp(integ + fade * self.rate_i * e * dt, -s.rate_int_lim, s.rate_int_lim)
        self.state[:, 3:6] = w_b
        return tau

    def attitude_loop(self, q, qd):
        s = self.spec
        ez, ezd = R.quat_body_z(q), R.quat_body_z(qd)
        d = (ez * ezd).sum(-1)
        c = torch.cross(ez, ezd, dim=-1)
        tilt = R.qnormalize(torch.cat((c, (d + 1.0).unsqueeze(-1)), -1))
        qd_red = torch.where((d < -1.0 + 1e-5).unsqueeze(-1), qd, R.qmul(tilt, q))
        qmix = R.qmul(R.qconj(qd_red), qd)
        qmix = torch.where(qmix[:, 3:4] < 0, -qmix, qmix)
        mw = _clamp(qmix[:, 3], -1.0, 1.0)
        mz = _clamp(qmix[:, 2], -1.0, 1.0)
        self.last_conditioning = (d.clone(), mw.clone())  # test hook: 1+d and asin(mz) near |mz|=1 amplify rounding
        zero = torch.zeros_like(mw)
        yawq = torch.stack((zero, zero, torch.sin(s.att_yaw_w * torch.asin(mz)), torch.cos(s.att_yaw_w * torch.acos(mw))), -1)
        qdd = R.qmul(qd_red, yawq)
        qe = R.qmul(R.qconj(q), qdd)
        sgn = torch.where(qe[:, 3:4] < 0, -2.0 * torch.ones_like(qe[:, 3:4]), 2.0 * torch.ones_like(qe[:, 3:4]))
        rate = sgn * qe[:, 0:3] * self.att_p
        return torch.max(torch.min(rate, self.att_rate_lim), -self.att_rate_lim)

    def velocity_loop(self, v_sp, yaw_sp, v, dt):
        s = self.spec
        integ, prev = self.state[:, 6:9], self.state[:, 9:12]
        e = v_sp - v
        vdot = (v - prev) / dt
        acc = self.vel_p * e + integ - self.vel_d * vdot
        self.state[:, 6:9] = torch.max(torch.min(integ + self.vel_i * e * dt, self.vel_int_lim), -self.vel_int_lim)
        self.state[:, 9:12] = v
        fx, fy, fz = acc[:, 0], acc[:, 1], acc[:, 2] + s.gravity
        fz_min = torch.as_tensor(0.1, dtype=self.dtype) * s.gravity
        fz = torch.where(fz < fz_min, fz_min.expand_as(fz), fz)
        h = torch.sqrt(fx * fx + fy * fy)
        hmax = fz * s.tilt_max_tan
        k = torch.where(h > hmax, hmax / h, torch.ones_like(h))
        fx, fy = torch.where(h > hmax, fx * k, fx), torch.where(h > hmax, fy * k, fy)
        fn = torch.sqrt(fx * fx + fy * fy + fz * fz)
        bz = torch.stack((fx / fn, fy / fn, fz / fn), -1)
        thrust = _clamp(s.hover_thrust * fn / s.gravity, s.thr_min, s.thr_max)
        yc = torch.stack((-torch.sin(yaw_sp), torch.cos(yaw_sp), torch.zeros_like(yaw_sp)), -1)
        bx = torch.cross(yc, bz, dim=-1)
        bx = bx / torch.sqrt((bx * bx).sum(-1, keepdim=True))
        by = torch.cross(bz, bx, dim=-1)
        m = torch.stack((bx, by, bz), -1)  # columns
        q_sp = R.matrix_to_quaternion(m)[:, [1, 2, 3, 0]]
        return q_sp, thrust

    # -- update: actions are the shaped + clamped actions of hovering.py:212-216 -------------------------
    def update(self, actions, angvel=None, dt=None):
        s = self.spec
        mode = s.ctl_mode
        a = actions.to(self.dtype)
        if mode == "rate":  # update(actions, ang_vel, dt) after set_q_world (hovering.py:249-250)
            w_b = self.body_rates(angvel.to(self.dtype))
            tau = self.rate_loop(a[:, 0:3], w_b, dt)
            return self.mixer(a[:, 3], tau)
        dt = self.dt
        w_b = self.body_rates(self.angvel_w)
        if mode == "atti":
            t = a[:, [1, 2, 3, 0]]  # action is (w,x,y,z,thrust)
            n2 = (t * t).sum(-1, keepdim=True)
            ident = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=self.dtype).expand_as(t)
            q_sp = R.qnormalize(torch.where(n2 < 1e-12, ident, t))
            thrust = a[:, 4]
        else:
            if mode == "pos":
                v_sp = torch.max(torch.min(self.pos_p * (a[:, 0:3] - self.pos), self.vel_sp_lim), -self.vel_sp_lim)
            else:
                v_sp = a[:, 0:3]
            q_sp, thrust = self.velocity_loop(v_sp, a[:, 3], self.linvel, dt)
        w_sp = self.attitude_loop(R.qnormalize(self.q), q_sp)
        tau = self.rate_loop(w_sp, w_b, dt)
        return self.mixer(thrust, tau)
