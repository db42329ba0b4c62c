#!/usr/bin/env python
"""Extract the collision cylinders of the reference's `thin` asset group (airgym/assets/env_assets/thin/tree_<i>.urdf,
one tilted cylinder each) into airgym_b200/assets/thin_trees.npy: [100,8] float32 rows
(cx, cy, cz, ax, ay, az, radius, half_length) in the asset frame — centre, unit axis = R(rpy) e_z (URDF rpy:
R = Rz(yaw) Ry(pitch) Rx(roll)).  Run in the build container only (needs /root/reference); the table is data, not code.
"""
import math, os, re, sys
import numpy as np
REF = "/root/reference/airgym/assets/env_assets/thin"
rows = []
for i in range(100):
    s = open(os.path.join(REF, f"tree_{i}.urdf")).read()
    m = re.search(r'<collision.*?<cylinder radius="([^"]+)" length="([^"]+)".*?<origin xyz="([^"]+)" rpy="([^"]+)"', s, re.S)
    r, L = float(m.group(1)), float(m.group(2))
    c = [float(x) for x in m.group(3).split()]
    roll, pitch, yaw = [float(x) for x in m.group(4).split()]
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    # third column of Rz Ry Rx
    a = [cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr]
    rows.append(c + a + [r, 0.5 * L])
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "airgym_b200", "assets", "thin_trees.npy")
np.save(out, np.asarray(rows, np.float32))
print(out, np.asarray(rows).shape)
