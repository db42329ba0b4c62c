"""The pytorch3d-convention restatements in oracle/rotations.py against closed forms."""
import math

import torch

from oracle import rotations as R


def test_euler_xyz_is_rx_ry_rz():
    a = torch.tensor([[0.3, -0.2, 0.9], [0.0, 0.0, 0.5], [1.2, 0.7, -2.0]], dtype=torch.float64)
    m = R.euler_angles_to_matrix(a, "XYZ")
    for row, (x, y, z) in zip(m, a.tolist()):
        cx, sx, cy, sy, cz, sz = math.cos(x), math.sin(x), math.cos(y), math.sin(y), math.cos(z), math.sin(z)
        rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float64)
        ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
        rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float64)
        assert torch.allclose(row, rx @ ry @ rz, atol=1e-12)


def test_matrix_to_euler_inverts_euler_to_matrix():
    a = (torch.rand(200, 3, dtype=torch.float64) - 0.5) * torch.tensor([6.0, 3.0, 6.0])
    a[:, 1].clamp_(-1.5, 1.5)
    back = R.matrix_to_euler_xyz(R.euler_angles_to_matrix(a, "XYZ"))
    assert torch.allclose(back, a, atol=1e-9)
    # quirk Q6: the "yaw" used by the reward is the third INTRINSIC-XYZ angle atan2(-R01, R00)
    m = R.euler_angles_to_matrix(a, "XYZ")
    assert torch.allclose(back[:, 2], torch.atan2(-m[:, 0, 1], m[:, 0, 0]))


def test_quaternion_matrix_round_trip_and_conventions():
    q = torch.randn(500, 4, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    q = torch.where(q[:, :1] < 0, -q, q)  # w >= 0 (wxyz)
    m = R.quaternion_to_matrix(q)
    assert torch.allclose(m @ m.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand(500, 3, 3), atol=1e-12)
    assert torch.allclose(torch.linalg.det(m), torch.ones(500, dtype=torch.float64))
    assert torch.allclose(R.matrix_to_quaternion(m), q, atol=1e-9)
    # rotation about z by +90deg maps x to y
    qz = torch.tensor([[math.cos(math.pi / 4), 0, 0, math.sin(math.pi / 4)]], dtype=torch.float64)
    assert torch.allclose(R.quaternion_to_matrix(qz)[0] @ torch.tensor([1.0, 0, 0], dtype=torch.float64),
                          torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64), atol=1e-12)
    # quaternion_to_matrix is scale invariant (2/|q|^2)
    assert torch.allclose(R.quaternion_to_matrix(3.0 * q), m, atol=1e-12)


def test_quat_axis_matches_matrix_column():
    q = torch.randn(100, 4, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)  # xyzw
    m = R.quaternion_to_matrix(q[:, [3, 0, 1, 2]])
    for axis in range(3):
        assert torch.allclose(R.quat_axis(q, axis), m[:, :, axis], atol=1e-12)
    assert torch.allclose(R.quat_body_z(q), m[:, :, 2], atol=1e-12)


def test_yaw_diff_wraps():
    a = torch.tensor([3.0, -3.0, 0.1])
    b = torch.tensor([-3.0, 3.0, 0.3])
    d = R.compute_yaw_diff(a, b)
    assert torch.allclose(d, torch.tensor([2 * math.pi - 6.0, 6.0 - 2 * math.pi, 0.2]), atol=1e-6)
