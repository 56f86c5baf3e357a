"""Shared helpers for tests: reference fixtures (tests/golden/reference_fixtures.json, extracted from the reference's
own tests/tutorials by tests/golden/extract_fixtures.py) and small pose utilities (numpy only)."""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, "golden", "reference_fixtures.json")))


def quat_to_R(qr, qx, qy, qz):
    return np.array([
        [qr * qr + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - qr * qz), 2 * (qz * qx + qr * qy)],
        [2 * (qx * qy + qr * qz), qr * qr - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - qr * qx)],
        [2 * (qz * qx - qr * qy), 2 * (qy * qz + qr * qx), qr * qr - qx * qx - qy * qy + qz * qz]])


def T_from_xyzq(v):
    T = np.eye(4)
    T[:3, :3] = quat_to_R(*v[3:7])
    T[:3, 3] = v[:3]
    return T


def T_from_xyzypr(v):
    x, y, z, yaw, pitch, roll = v
    cy, sy, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    T = np.eye(4)
    T[:3, :3] = [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                 [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                 [-sp, cp * sr, cp * cr]]
    T[:3, 3] = [x, y, z]
    return T


def T_from_flat12(p):
    T = np.eye(4)
    T[:3, :3] = np.asarray(p[:9]).reshape(3, 3)
    T[:3, 3] = p[9:12]
    return T


def flat12_from_T(T):
    return np.concatenate([T[:3, :3].ravel(), T[:3, 3]])


def ypr_vec_from_T(T):
    R = T[:3, :3]
    pitch = math.atan2(-R[2, 0], math.hypot(R[0, 0], R[1, 0]))
    roll = math.atan2(R[2, 1], R[2, 2])
    yaw = math.atan2(R[1, 0], R[0, 0])
    return np.array([T[0, 3], T[1, 3], T[2, 3], yaw, pitch, roll])


def se2_noise_lambda(sxy, syaw):
    nl = np.zeros(16)
    nl[:9] = np.diag([1 / sxy ** 2, 1 / sxy ** 2, 1 / syaw ** 2]).ravel()
    return nl


def inject_state(eo, ep):
    """Copies the oracle's numeric state (edge poses, unknown landmark positions, numeric spanning tree) into the
    product engine so that the next optimize call of both starts from bit-identical inputs."""
    _, poses = eo.edges()
    for i in range(eo.num_edges):
        ep.set_edge_pose(i, poses[i])
    ids_, _, pos = eo.unknown_lms()
    for i, q in zip(ids_, pos):
        ep.set_lm_pos(int(i), q)
    for src in range(eo.num_keyframes):
        tt, pp, _ = eo.num(src)
        for dst, q in zip(tt, pp):
            ep.set_num(src, int(dst), q)
