"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the rotation helpers the reference pulls from two places:

* ``pytorch3d.transforms`` [EXT, un-vendored, unpinned — SURVEY.md §8c-1]: ``euler_angles_to_matrix``,
  ``matrix_to_quaternion``, ``quaternion_to_matrix``, ``matrix_to_euler_angles`` as called at
  reference ``airgym/envs/base/hovering.py:323-324,338,401-403``.  Restated from the published pytorch3d
  algorithm (wxyz quaternions, 'XYZ' intrinsic convention).  PARITY UNPINNED for these four: the reference
  tree holds no test or golden vector for them; `tests/test_oracle_rotations.py` checks them against closed
  forms instead.
* ``airgym/utils/torch_utils.py`` (in-tree): ``quat_rotate`` :58-68 (re-stated in hovering.py:464-474),
  ``quat_axis`` :476-481, ``torch_rand_float`` :192-193, ``tensor_clamp`` :199-201.  These are pinned:
  `tests/golden/make_golden.py` runs the reference's own copies.
"""
import torch


def quaternion_to_matrix(q_wxyz: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(q_wxyz, -1)
    two_s = 2.0 / (q_wxyz * q_wxyz).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q_wxyz.shape[:-1] + (3, 3))


def _axis_rotation(axis: str, angle: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, c, -s, zero, s, c)
    elif axis == "Y":
        flat = (c, zero, s, zero, one, zero, -s, zero, c)
    else:
        flat = (c, -s, zero, s, c, zero, zero, zero, one)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(angles: torch.Tensor, convention: str = "XYZ") -> torch.Tensor:
    mats = [_axis_rotation(c, a) for c, a in zip(convention, torch.unbind(angles, -1))]
    return torch.matmul(torch.matmul(mats[0], mats[1]), mats[2])


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(m: torch.Tensor) -> torch.Tensor:
    """max-of-four-candidates; returns wxyz with w >= 0 (standardised, as newer pytorch3d does)."""
    batch = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(batch + (9,)), -1)
    q_abs = _sqrt_positive_part(
        torch.stack((1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22), -1)
    )
    cand = torch.stack(
        (
            torch.stack((q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01), -1),
            torch.stack((m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20), -1),
            torch.stack((m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21), -1),
            torch.stack((m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2), -1),
        ),
        -2,
    )
    floor = torch.tensor(0.1, dtype=q_abs.dtype)
    cand = cand / (2.0 * q_abs[..., None].max(floor))
    idx = q_abs.argmax(-1)
    out = torch.gather(cand, -2, idx[..., None, None].expand(batch + (1, 4))).squeeze(-2)
    return torch.where(out[..., 0:1] < 0, -out, out)


def matrix_to_euler_xyz(m: torch.Tensor) -> torch.Tensor:
    """matrix_to_euler_angles(M, 'XYZ') = (atan2(-M12, M22), asin(M02), atan2(-M01, M00))."""
    return torch.stack(
        (torch.atan2(-m[..., 1, 2], m[..., 2, 2]), torch.asin(m[..., 0, 2]), torch.atan2(-m[..., 0, 1], m[..., 0, 0])), -1
    )


# ---- in-tree helpers (airgym/utils/torch_utils.py, hovering.py:464-486) -------------------------------
def quat_rotate(q_xyzw: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    q_w = q_xyzw[:, -1]
    q_vec = q_xyzw[:, :3]
    a = v * (2.0 * q_w**2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(-1, 1, 3), v.view(-1, 3, 1)).squeeze(-1) * 2.0
    return a + b + c


def quat_axis(q_xyzw: torch.Tensor, axis: int = 0) -> torch.Tensor:
    basis = torch.zeros(q_xyzw.shape[0], 3, dtype=q_xyzw.dtype)
    basis[:, axis] = 1
    return quat_rotate(q_xyzw, basis)


def rand_float(lower, upper, u: torch.Tensor) -> torch.Tensor:
    """torch_rand_float with the U[0,1) draw `u` made explicit."""
    return (upper - lower) * u + lower


def tensor_clamp(t, lo, hi):
    return torch.max(torch.min(t, hi), lo)


def compute_yaw_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """hovering.py:33-38"""
    diff = b - a
    diff = torch.where(diff < -torch.pi, diff + 2 * torch.pi, diff)
    diff = torch.where(diff > torch.pi, diff - 2 * torch.pi, diff)
    return diff


# ---- xyzw quaternion algebra used by the builder-defined controller/integrator ------------------------
def qmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack(
        (
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ),
        -1,
    )


def qconj(a: torch.Tensor) -> torch.Tensor:
    return torch.cat((-a[..., :3], a[..., 3:]), -1)


def qnormalize(a: torch.Tensor) -> torch.Tensor:
    return a / torch.sqrt((a * a).sum(-1, keepdim=True))


def quat_body_z(q: torch.Tensor) -> torch.Tensor:
    x, y, z, w = q.unbind(-1)
    return torch.stack((2.0 * (x * z + y * w), 2.0 * (y * z - x * w), 1.0 - 2.0 * (x * x + y * y)), -1)
