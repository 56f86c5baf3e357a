"""Analytic-vs-finite-difference check of every dh_dAp / dh_df formula of the oracle, following the reference's own
numeric-Jacobian recipe (impl/jacobians.h:90-133 numeric_dh_dAp, :170-186 numeric_dh_df): normal and inverse edge
direction, with and without the A pose (A==NULL means d+1 == observer)."""
import math

import numpy as np
import pytest

from fixtures import T_from_xyzypr, flat12_from_T
from oracle import pyoracle as po


def rand_pose(kind, rng, scale=1.0):
    if kind == po.POSE_SE2:
        return np.array([rng.uniform(-3, 3) * scale, rng.uniform(-3, 3) * scale, rng.uniform(-math.pi, math.pi)])
    return flat12_from_T(T_from_xyzypr([rng.uniform(-2, 2) * scale, rng.uniform(-2, 2) * scale, rng.uniform(-2, 2) * scale,
                                        rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4)]))


def identity(kind):
    return np.zeros(3) if kind == po.POSE_SE2 else flat12_from_T(np.eye(4))


def base_from_obs(kind, A, D, p, normal_dir, eps):
    comp, inv, ex = po.pose_compose, po.pose_inverse, po.pose_exp
    incr = ex(kind, eps)
    if normal_dir:
        X = comp(kind, incr, D) if A is None else comp(kind, comp(kind, A, incr), D)
    else:
        Dp = comp(kind, p, D)
        pm = inv(kind, comp(kind, incr, p))
        X = comp(kind, pm, Dp) if A is None else comp(kind, comp(kind, A, pm), Dp)
    return X


CASES = [
    ("relpose2d", dict(pose_kind=po.POSE_SE2, obs_kind=po.OBS_RELPOSES2D)),
    ("cartesian3d", dict(pose_kind=po.POSE_SE3, obs_kind=po.OBS_CARTESIAN3D)),
    ("cartesian3d_sensorpose", dict(pose_kind=po.POSE_SE3, obs_kind=po.OBS_CARTESIAN3D, sensor_pose_kind=po.SENSORPOSE_SE3,
                                    sensor_pose=[0.1, 0.2, 0.3, 0.17, 0.35, 0.52])),
    ("mono", dict(pose_kind=po.POSE_SE3, obs_kind=po.OBS_MONO, cam_left=[200, 200, 400, 320], sensor_pose_kind=po.SENSORPOSE_SE3,
                  sensor_pose=[0, 0, 0, math.radians(-90), 0, math.radians(-90)])),
    ("stereo", dict(pose_kind=po.POSE_SE3, obs_kind=po.OBS_STEREO, cam_left=[200, 150, 512, 384], cam_right=[200, 150, 512, 384],
                    right_cam_pose=[0.2, 0, 0, 1, 0, 0, 0], sensor_pose_kind=po.SENSORPOSE_SE3,
                    sensor_pose=[0, 0, 0, math.radians(-90), 0, math.radians(-90)])),
]


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("normal_dir", [True, False])
@pytest.mark.parametrize("with_A", [True, False])
def test_dh_dAp_vs_finite_differences(name, kw, normal_dir, with_A):
    params = po.default_params(**kw)
    kind = params.pose_kind
    P = 3 if kind == po.POSE_SE2 else 6
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(40):
        scale = 0.2 if kind == po.POSE_SE3 else 1.0
        A = rand_pose(kind, rng, scale) if with_A else None
        D = rand_pose(kind, rng, scale)
        p = rand_pose(kind, rng, scale)
        if kw["obs_kind"] in (po.OBS_MONO, po.OBS_STEREO):
            lm = np.array([rng.uniform(4, 8), rng.uniform(-1, 1), rng.uniform(-1, 1)])  # in front of the camera (robot +X)
        else:
            lm = rng.uniform(-3, 3, 3)
        J, valid = po.eval_dh_dAp(params, A, D, p, normal_dir, lm)
        assert valid == 1
        X0 = base_from_obs(kind, A, D, p, normal_dir, np.zeros(P))
        if kw["obs_kind"] == po.OBS_RELPOSES2D:
            z = X0.copy()   # analytic J is exact only at zero residual
            h = lambda eps: -po.predict_err(params, base_from_obs(kind, A, D, p, normal_dir, eps), lm, z)
        else:
            z = np.zeros(po.OBS_DIMS[kw["obs_kind"]])
            h = lambda eps: -po.predict_err(params, base_from_obs(kind, A, D, p, normal_dir, eps), lm, z)
        Jn = np.zeros_like(J)
        step = 1e-6
        for k in range(P):
            e = np.zeros(P)
            e[k] = step
            Jn[:, k] = (h(e) - h(-e)) / (2 * step)
        err = np.abs(J - Jn).max() / max(1.0, np.abs(Jn).max())
        worst = max(worst, err)
    assert worst < 2e-6, worst


@pytest.mark.parametrize("name,kw", [c for c in CASES if c[0] != "relpose2d"])
@pytest.mark.parametrize("with_pose", [True, False])
def test_dh_df_vs_finite_differences(name, kw, with_pose):
    params = po.default_params(**kw)
    rng = np.random.default_rng(13)
    for _ in range(40):
        pose = rand_pose(po.POSE_SE3, rng, 0.2) if with_pose else None
        if kw["obs_kind"] in (po.OBS_MONO, po.OBS_STEREO):
            lm = np.array([rng.uniform(4, 8), rng.uniform(-1, 1), rng.uniform(-1, 1)])
        else:
            lm = rng.uniform(-3, 3, 3)
        J, valid = po.eval_dh_df(params, pose, lm)
        assert valid == 1
        Pm = identity(po.POSE_SE3) if pose is None else pose
        z = np.zeros(po.OBS_DIMS[kw["obs_kind"]])
        Jn = np.zeros_like(J)
        step = 1e-6
        for k in range(3):
            e = np.zeros(3)
            e[k] = step
            Jn[:, k] = (-po.predict_err(params, Pm, lm + e, z) + po.predict_err(params, Pm, lm - e, z)) / (2 * step)
        assert np.abs(J - Jn).max() / max(1.0, np.abs(Jn).max()) < 2e-6


def test_pinhole_behind_camera_is_invalid():
    # models/sensors.h:92-93: z<=0 marks the observation invalid and zero-fills the block (jacobians.h:322-327)
    params = po.default_params(pose_kind=po.POSE_SE3, obs_kind=po.OBS_MONO, cam_left=[200, 200, 400, 320])
    J, valid = po.eval_dh_dAp(params, None, identity(po.POSE_SE3), None, True, np.array([0.1, 0.2, -1.0]))
    assert valid == 0 and np.all(J == 0)
