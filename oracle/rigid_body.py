"""ORACLE (test infrastructure, never on the product path).

Stand-in for ``gym.simulate`` (reference ``airgym/envs/base/hovering.py:290``) — IsaacGym Preview 4 / PhysX,
a closed binary absent from /root/reference (``configuration.sh:124-150``).  PARITY UNPINNED against PhysX:
the reference has no test or recorded trajectory.  The model is the one BASELINE.json's north_star asks
for: a single composite rigid body (X152b URDF: base 0.585 kg, I=diag(0.04); 4 rotors of
0.004 kg, I=1e-6 at (±0.05374, ±0.05374, 0.024)), body-frame rotor forces 9.59·cmd along +z at the rotor
positions (hovering.py:256-268), rotor reaction torques ∓0.2·cmd about z (hovering.py:270-275), gravity
(0,0,-9.81), no damping (assets/__init__.py:30-31), |v|,|w| ≤ 100 (:34-35); classic RK4 over dt=0.01 with the
wrench held constant in the body frame (optional semi-implicit Euler for A/B against PhysX-like stepping).
State row = [p, q_xyzw, v_world, w_world] as IsaacGym's root-state tensor (hovering.py:73-77).
"""
import torch

from . import rotations as R
from .spec import QuadSpec


def _deriv(spec, inertia, v, q, w, fz_over_m, tau):
    bz = R.quat_body_z(q)
    dv = bz * fz_over_m.unsqueeze(-1)
    dv = torch.stack((dv[:, 0], dv[:, 1], dv[:, 2] - spec.gravity), -1)
    x, y, z, qw = q.unbind(-1)
    wx, wy, wz = w.unbind(-1)
    dq = 0.5 * torch.stack(
        (qw * wx + y * wz - z * wy, qw * wy + z * wx - x * wz, qw * wz + x * wy - y * wx, -x * wx - y * wy - z * wz), -1
    )
    Iw = inertia * w
    dw = (tau - torch.cross(w, Iw, dim=-1)) / inertia
    return v, dv, dq, dw


def simulate(spec: QuadSpec, state: torch.Tensor, rotor_force: torch.Tensor, tau_z: torch.Tensor):
    """state [N,13] (modified in place), rotor_force [N,4] in newtons, tau_z [N]. Returns R(q_new) [N,3,3]."""
    dt_ = state.dtype
    inertia = torch.tensor(spec.inertia, dtype=dt_)
    p, q, v, w_world = state[:, 0:3].clone(), state[:, 3:7].clone(), state[:, 7:10].clone(), state[:, 10:13].clone()
    Rm = R.quaternion_to_matrix(q[:, [3, 0, 1, 2]])
    w = torch.einsum("nji,nj->ni", Rm, w_world)
    f = rotor_force.to(dt_)
    fz_over_m = (f[:, 0] + f[:, 1] + f[:, 2] + f[:, 3]) / spec.mass
    tau = torch.stack(
        (spec.arm * (-f[:, 0] + f[:, 1] + f[:, 2] - f[:, 3]), spec.arm * (-f[:, 0] + f[:, 1] - f[:, 2] + f[:, 3]), tau_z.to(dt_)), -1
    )
    h = spec.dt
    if spec.integrator == "euler":
        _, dv, _, dw = _deriv(spec, inertia, v, q, w, fz_over_m, tau)
        v = v + h * dv
        p = p + h * v
        w = w + h * dw
        _, _, dq, _ = _deriv(spec, inertia, v, q, w, fz_over_m, tau)
        q = q + h * dq
    else:
        k1 = _deriv(spec, inertia, v, q, w, fz_over_m, tau)
        k2 = _deriv(spec, inertia, v + 0.5 * h * k1[1], q + 0.5 * h * k1[2], w + 0.5 * h * k1[3], fz_over_m, tau)
        k3 = _deriv(spec, inertia, v + 0.5 * h * k2[1], q + 0.5 * h * k2[2], w + 0.5 * h * k2[3], fz_over_m, tau)
        k4 = _deriv(spec, inertia, v + h * k3[1], q + h * k3[2], w + h * k3[3], fz_over_m, tau)
        h6 = h / 6.0
        p = p + h6 * (k1[0] + 2.0 * k2[0] + 2.0 * k3[0] + k4[0])
        v = v + h6 * (k1[1] + 2.0 * k2[1] + 2.0 * k3[1] + k4[1])
        q = q + h6 * (k1[2] + 2.0 * k2[2] + 2.0 * k3[2] + k4[2])
        w = w + h6 * (k1[3] + 2.0 * k2[3] + 2.0 * k3[3] + k4[3])
    q = R.qnormalize(q)
    Rn = R.quaternion_to_matrix(q[:, [3, 0, 1, 2]])
    ww = torch.einsum("nij,nj->ni", Rn, w)
    vn = torch.sqrt((v * v).sum(-1, keepdim=True))
    v = torch.where(vn > spec.max_lin_vel, v * (spec.max_lin_vel / vn), v)
    wn = torch.sqrt((ww * ww).sum(-1, keepdim=True))
    ww = torch.where(wn > spec.max_ang_vel, ww * (spec.max_ang_vel / wn), ww)
    state[:, 0:3], state[:, 3:7], state[:, 7:10], state[:, 10:13] = p, q, v, ww
    return Rn
