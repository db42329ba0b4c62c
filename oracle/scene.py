"""ORACLE (test infrastructure, never on the product path).

Builder-defined stand-ins for what the reference gets from IsaacGym's cameras and PhysX contacts in the Customized task
family — PARITY UNPINNED: the reference delegates all of it to a closed binary.

* depth camera (customized.py:386-391 `render_cameras`, asset config `avoid_config.py:55-68`): pin-hole, 212x120, horizontal
  fov 87 deg, mounted at body (0.15, 0, 0.1) looking along body +x; value = planar depth (distance along the optical axis),
  "no hit" (nothing closer than the 5 m far plane) = +inf.  The reference then stores `-camera_tensor.T`, i.e. a [212,120]
  (width-major) array of positive depths (customized.py:402).
* scene primitives: ground plane z = 0 (`create_ground_plane`), capped cylinders (the `thin` trees:
  airgym_b200/assets/thin_trees.npy, slot i of the 40 uses tree_<i>.urdf — the reference picks with an unseeded
  `random.choice`, asset_manager.py:141), sphere r = 0.2 (balls/ball/model.urdf), axis-aligned box of half-extent 0.15
  (cubes/1x1: +-1 mesh scaled 0.15).
* contacts (customized.py:393-397 `check_collisions`, ||net contact force|| > 0.1): the drone's r = 0.2 collision sphere
  (robots/X152b/model.urdf:13-18) against the ground, the tree capsules and the cube; the goal ball does not collide.
* thrown cube (avoid.py:91-158 leaves its flight to PhysX): ballistic semi-implicit Euler, comes to rest on the ground.
"""
import math
import os

import numpy as np
import torch

CAM_W, CAM_H = 212, 120
CAM_HFOV_DEG = 87.0
CAM_FAR = 5.0
CAM_OFFSET = (0.15, 0.0, 0.1)
CAM_F = (CAM_W / 2) / math.tan(math.radians(CAM_HFOV_DEG) / 2)  # focal length in pixels
DRONE_RADIUS = 0.2
BALL_RADIUS = 0.2
CUBE_HALF = 0.15
NUM_TREES = 40
T_MIN = 1e-3

_TREES = None


def tree_table():
    """[40,8] float32: centre(3), unit axis(3), radius, half-length in the asset frame (slot i = tree_<i>.urdf)."""
    global _TREES
    if _TREES is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "airgym_b200", "assets", "thin_trees.npy")
        _TREES = torch.from_numpy(np.load(p)[:NUM_TREES].copy())
    return _TREES


def place_trees(xy, yaw):
    """World-frame cylinders of one env batch: xy [N,40,2], yaw [N,40] -> centre [N,40,3], axis [N,40,3], r [40], h [40]."""
    t = tree_table().to(xy.dtype)
    c, s = torch.cos(yaw), torch.sin(yaw)
    cx = c * t[:, 0] - s * t[:, 1] + xy[..., 0]
    cy = s * t[:, 0] + c * t[:, 1] + xy[..., 1]
    cz = t[:, 2].expand_as(cx)
    ax = c * t[:, 3] - s * t[:, 4]
    ay = s * t[:, 3] + c * t[:, 4]
    az = t[:, 5].expand_as(ax)
    return torch.stack((cx, cy, cz), -1), torch.stack((ax, ay, az), -1), t[:, 6], t[:, 7]


def quat_to_mat(q):  # xyzw, unit
    x, y, z, w = q.unbind(-1)
    return torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)), -1).reshape(q.shape[:-1] + (3, 3))


def camera_rays(pos, quat):
    """Origin [N,3] and world directions [N,W,H,3] (x_cam component = 1, so the ray parameter IS the planar depth)."""
    R = quat_to_mat(quat)
    off = torch.tensor(CAM_OFFSET, dtype=pos.dtype)
    o = pos + (R @ off)
    u = torch.arange(CAM_W, dtype=pos.dtype)
    v = torch.arange(CAM_H, dtype=pos.dtype)
    dy = (CAM_W / 2 - u - 0.5) / CAM_F
    dz = (CAM_H / 2 - v - 0.5) / CAM_F
    d_body = torch.stack((torch.ones(CAM_W, CAM_H, dtype=pos.dtype), dy[:, None].expand(CAM_W, CAM_H),
                          dz[None, :].expand(CAM_W, CAM_H)), -1)  # [W,H,3]
    d = torch.einsum("nij,whj->nwhi", R, d_body)
    return o, d


def _hit_ground(o, d):
    dz = d[..., 2]
    t = -o[:, None, None, 2] / dz
    return torch.where((dz < 0) & (t > T_MIN), t, torch.full_like(t, float("inf")))


def _hit_sphere(o, d, c, r):
    oc = o - c  # [N,3]
    a = (d * d).sum(-1)
    b = (d * oc[:, None, None, :]).sum(-1)
    cc = (oc * oc).sum(-1)[:, None, None] - r * r
    disc = b * b - a * cc
    t = (-b - torch.sqrt(disc.clamp_min(0))) / a
    return torch.where((disc > 0) & (t > T_MIN), t, torch.full_like(t, float("inf")))


def _hit_box(o, d, c, half):
    inv = 1.0 / d
    t1 = (c[:, None, None, :] - half - o[:, None, None, :]) * inv
    t2 = (c[:, None, None, :] + half - o[:, None, None, :]) * inv
    tn = torch.minimum(t1, t2).amax(-1)
    tf = torch.maximum(t1, t2).amin(-1)
    return torch.where((tn <= tf) & (tn > T_MIN), tn, torch.full_like(tn, float("inf")))


def _hit_cylinder(o, d, c, a, r, h):
    """Capped cylinder: centre c [N,3], unit axis a [N,3], radius r, half-length h (scalars); side wall, then end caps."""
    oc = (o - c)[:, None, None, :]
    an = a[:, None, None, :]
    card = (an * d).sum(-1)
    caoc = (an * oc).sum(-1)
    A = (d * d).sum(-1) - card * card
    B = (oc * d).sum(-1) - caoc * card
    C = (oc * oc).sum(-1) - caoc * caoc - r * r
    disc = B * B - A * C
    inf = torch.full_like(A, float("inf"))
    sq = torch.sqrt(disc.clamp_min(0))
    t = (-B - sq) / A
    y = caoc + t * card
    side = torch.where((disc > 0) & (y.abs() < h) & (t > T_MIN), t, inf)
    # caps: the one facing the ray origin
    sgn = torch.where(y < 0, -torch.ones_like(y), torch.ones_like(y))
    tc = (sgn * h - caoc) / card
    cap_ok = (disc > 0) & ((B + A * tc).abs() < sq) & (tc > T_MIN)
    cap = torch.where(cap_ok, tc, inf)
    return torch.minimum(side, cap)


def render_depth(pos, quat, *, trees=None, ball=None, cube=None):
    """Planar depth [N,W=212,H=120] (float32, +inf = no hit within the far plane) — what `-camera_tensor.T` holds.
    trees = (xy [N,40,2], yaw [N,40]); ball = centre [N,3]; cube = centre [N,3]."""
    o, d = camera_rays(pos, quat)
    t = _hit_ground(o, d)
    if trees is not None:
        c, a, r, h = place_trees(*trees)
        for j in range(c.shape[1]):
            t = torch.minimum(t, _hit_cylinder(o, d, c[:, j], a[:, j], float(r[j]), float(h[j])))
    if ball is not None:
        t = torch.minimum(t, _hit_sphere(o, d, ball, BALL_RADIUS))
    if cube is not None:
        t = torch.minimum(t, _hit_box(o, d, cube, CUBE_HALF))
    return torch.where(t > CAM_FAR, torch.full_like(t, float("inf")), t)


def drone_contacts(pos, *, trees=None, cube=None):
    """bool [N]: the drone's collision sphere touches the ground, a tree capsule or the cube."""
    hit = pos[:, 2] < DRONE_RADIUS
    if trees is not None:
        c, a, r, h = place_trees(*trees)
        rel = pos[:, None, :] - c
        s = (rel * a).sum(-1).clamp(-h, h)
        dist = torch.norm(rel - s[..., None] * a, dim=-1)
        hit = hit | (dist < (r + DRONE_RADIUS)).any(-1)
    if cube is not None:
        q = (pos - cube).abs() - CUBE_HALF
        dist = torch.norm(q.clamp_min(0), dim=-1)
        hit = hit | (dist < DRONE_RADIUS)
    return hit


def cube_step(p, v, dt=0.01, g=9.81):
    """One dt of the thrown cube (in place): semi-implicit Euler; lands and stays at z = half-extent; parked cubes
    (x = -999, avoid.py:124-129) do not move."""
    parked = p[:, 0] == -999.0
    vz = v[:, 2] - g * dt
    v_new = torch.stack((v[:, 0], v[:, 1], vz), -1)
    p_new = p + v_new * dt
    landed = p_new[:, 2] < CUBE_HALF
    p_new[:, 2] = torch.where(landed, torch.full_like(vz, CUBE_HALF), p_new[:, 2])
    v_new = torch.where(landed[:, None], torch.zeros_like(v_new), v_new)
    keep = parked[:, None]
    p.copy_(torch.where(keep, p, p_new))
    v.copy_(torch.where(keep, v, v_new))
