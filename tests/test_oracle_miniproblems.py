"""End-to-end pins of the oracle against the reference's integration tests:
tests/fixed-transformations_unittest.cpp (1e-3), tests/sensor-pose_unittest.cpp (1e-2),
tests/submaps_edge_init_values.cpp (loop-closure bookkeeping, rmse<1e-6)."""
import math

import numpy as np
import pytest

from fixtures import FX, T_from_flat12, T_from_xyzq, T_from_xyzypr, se2_noise_lambda, ypr_vec_from_T
from oracle import pyoracle as po


def make_engine(**kw):
    return po.OracleEngine(po.default_params(**kw))


@pytest.mark.parametrize("inverse", [False, True])
def test_fixed_transformations(inverse):
    fx = FX["fixed_transformations"]
    pts = np.array(fx["dummy_obs"])
    for tr in fx["test_fixed_transfs"]:
        e = make_engine(pose_kind=po.POSE_SE3, obs_kind=po.OBS_CARTESIAN3D, use_robust_kernel=0, max_tree_depth=3, max_optimize_depth=3)
        ids = pts[:, 0].astype(np.uint64)
        e.define_new_keyframe(ids, pts[:, 1:4])
        T = T_from_xyzypr(tr)
        if inverse:
            obs1 = (np.linalg.inv(T) @ np.c_[pts[:, 1:4], np.ones(len(pts))].T).T[:, :3]
        else:
            obs1 = (T @ np.c_[pts[:, 1:4], np.ones(len(pts))].T).T[:, :3]
        r = e.define_new_keyframe(ids, obs1)
        est = T_from_flat12(e.get_edge(0)[2])
        if inverse:
            est = np.linalg.inv(est)
        assert np.abs(T - est).sum() < 1e-3, (tr, inverse, r["optimize_results"])


@pytest.mark.parametrize("displaced", [False, True])
def test_sensor_pose(displaced):
    fx = FX["sensor_pose_unittest"]
    d0 = np.array(fx["observations_0_displ" if displaced else "observations_0"])
    d1 = np.array(fx["observations_10_displ" if displaced else "observations_10"])
    kw = dict(pose_kind=po.POSE_SE3, obs_kind=po.OBS_CARTESIAN3D, use_robust_kernel=1, noise_std=0.1, max_tree_depth=3, max_optimize_depth=3)
    if displaced:
        kw.update(sensor_pose_kind=po.SENSORPOSE_SE3, sensor_pose=fx["sensor_pose_xyzypr"])
    e = make_engine(**kw)
    rng = np.random.default_rng(3)
    for d in (d0, d1):
        e.define_new_keyframe(d[:, 0].astype(np.uint64), d[:, 1:4] + rng.normal(0, 1e-9, (len(d), 3)))
    P = np.linalg.inv(T_from_flat12(e.get_edge(0)[2]))  # -inv_pose = pose of KF#1 wrt KF#0
    P_GT = np.linalg.inv(T_from_xyzq(fx["GT0_xyzq"])) @ T_from_xyzq(fx["GT1_xyzq"])
    assert np.abs(ypr_vec_from_T(P) - ypr_vec_from_T(P_GT)).sum() < 1e-2


def test_submaps_edge_init_values():
    ds = FX["submaps_edge_init_values"]["dataset"]
    sxy, syaw = 0.001, math.radians(0.05)
    e = make_engine(pose_kind=po.POSE_SE2, obs_kind=po.OBS_RELPOSES2D, noise_kind=po.NOISE_CONST_MATRIX, noise_lambda=se2_noise_lambda(sxy, syaw),
                    max_tree_depth=3, max_optimize_depth=3, submap_size=5, min_obs_to_loop_closure=1)
    rng = np.random.default_rng(5)
    idx, n_lc = 0, 0
    for kf in range(17):
        ids, obs, fixed = [kf], [[0, 0, 0]], [1]
        while idx < len(ds) and int(ds[idx][0]) == kf:
            d = ds[idx]
            ids.append(int(d[1]))
            obs.append([d[2] + rng.normal(0, sxy), d[3] + rng.normal(0, sxy), d[4] + rng.normal(0, syaw)])
            fixed.append(0)
            idx += 1
        r = e.define_new_keyframe(ids, obs, is_fixed=fixed)
        if len(r["created_edges"]) == 2:
            n_lc += 1
            ce = r["created_edges"]
            assert ce[0]["loopclosure_base_kf"] != po.INVALID or ce[1]["loopclosure_base_kf"] != po.INVALID
            assert ce[0]["loopclosure_observer_kf"] != po.INVALID or ce[1]["loopclosure_observer_kf"] != po.INVALID
            assert r["optimize_results"]["num_observations"] > 1
            assert r["optimize_results"]["obs_rmse"] < 1e-6
    assert n_lc == 1  # KF#11 observing KF#1 raises the 10<=>0 loop-closure edge (submaps_edge_init_values.cpp:38-48)
    ft, _ = e.edges()
    assert [0, 10] in ft.tolist()
