"""Synthetic restatements of the reference's datasets (the data files are not shipped: datasets/ only holds RWT
generator configs, /root/reference/datasets/README.txt:1-5).  Shapes follow SURVEY.md §8(d) G1..G5.  numpy only.

Every generator returns a list with one entry per keyframe: dict(feat_id, obs, is_fixed, is_unknown_with_init_val,
feat_rel_pos) in the exact form RbaEngine.define_new_keyframe() takes (apps/srba-slam/srba-run-generic-impl.h:337-465).
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------------ SE(2) helpers
def se2_compose(a, b):
    c, s = math.cos(a[2]), math.sin(a[2])
    return np.array([a[0] + b[0] * c - b[1] * s, a[1] + b[0] * s + b[1] * c, wrap(a[2] + b[2])])


def se2_inverse(a):
    c, s = math.cos(a[2]), math.sin(a[2])
    return np.array([-a[0] * c - a[1] * s, a[0] * s - a[1] * c, -a[2]])


def wrap(a):
    return (a + math.pi) % (2 * math.pi) - math.pi


def graphslam_se2_tour(n_kfs=3000, seed=1, cell=100.0, ncells=3, step=2.0, max_turn_deg=30.0, sensor_range=9.0, sigma_xy=0.001,
                       sigma_yaw_deg=0.2, add_noise=True, noise_seed=None):
    """G1: relative SE(2) graph-SLAM on a corridor tour of a `ncells` x `ncells` grid of `cell`-metre blocks
    (mirrors datasets/world-2d-30k-landmarks.wrl:8-53 and world-2d-30k-rel-graph-slam.cfg:26-27,44-45): 2 m / <=30 deg
    steps, each keyframe observes every earlier keyframe within `sensor_range` as a relative pose
    (CDatasetParser_RelGraphSLAM2D.h:30-43), Gaussian noise as `srba-slam --noise 0.001 --noise-ang 0.2 --add-noise`.
    Each keyframe's list starts with its own fixed fake landmark (srba-run-generic-impl.h:370-377)."""
    rng = np.random.default_rng(seed)
    # noise_seed: draw the observation noise from its own stream, so that several runs share the tour (hence the shape of
    # every local area) and differ only in the noise realisation; None keeps the single stream of the original generator
    nrng = rng if noise_seed is None else np.random.default_rng(noise_seed)
    nodes = [(i * cell, j * cell) for i in range(ncells + 1) for j in range(ncells + 1)]

    def neighbours(n):
        i, j = divmod(n, ncells + 1)
        out = []
        if i > 0: out.append(n - (ncells + 1))
        if i < ncells: out.append(n + (ncells + 1))
        if j > 0: out.append(n - 1)
        if j < ncells: out.append(n + 1)
        return out

    prev_node, cur_target = 0, 1
    pose = np.array([0.0, 0.0, math.atan2(nodes[1][1] - nodes[0][1], nodes[1][0] - nodes[0][0])])
    gt = [pose.copy()]
    max_turn = math.radians(max_turn_deg)
    while len(gt) < n_kfs:
        tx, ty = nodes[cur_target]
        if math.hypot(tx - pose[0], ty - pose[1]) < step:
            nb = [n for n in neighbours(cur_target) if n != prev_node] or neighbours(cur_target)
            prev_node, cur_target = cur_target, int(rng.choice(nb))
            tx, ty = nodes[cur_target]
        want = math.atan2(ty - pose[1], tx - pose[0])
        turn = max(-max_turn, min(max_turn, wrap(want - pose[2])))
        pose = np.array([pose[0], pose[1], wrap(pose[2] + turn)])
        pose = np.array([pose[0] + step * math.cos(pose[2]), pose[1] + step * math.sin(pose[2]), pose[2]])
        gt.append(pose.copy())
    gt = np.array(gt)
    sig_yaw = math.radians(sigma_yaw_deg)
    frames = []
    # spatial hash for the range queries
    cellsz = sensor_range
    grid = {}
    for k in range(n_kfs):
        ids, obs, fixed = [k], [[0.0, 0.0, 0.0]], [1]
        cx, cy = int(math.floor(gt[k, 0] / cellsz)), int(math.floor(gt[k, 1] / cellsz))
        cand = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                cand.extend(grid.get((cx + dx, cy + dy), ()))
        inv_k = se2_inverse(gt[k])
        for j in sorted(cand):
            if math.hypot(gt[j, 0] - gt[k, 0], gt[j, 1] - gt[k, 1]) < sensor_range:
                rel = se2_compose(inv_k, gt[j])
                if add_noise:
                    rel = rel + np.array([nrng.normal(0, sigma_xy), nrng.normal(0, sigma_xy), nrng.normal(0, sig_yaw)])
                ids.append(j)
                obs.append([rel[0], rel[1], rel[2]])
                fixed.append(0)
        grid.setdefault((cx, cy), []).append(k)
        frames.append(dict(feat_id=np.array(ids, np.uint64), obs=np.array(obs, np.float64), is_fixed=np.array(fixed, np.uint8),
                           is_unknown_with_init_val=np.zeros(len(ids), np.uint8), feat_rel_pos=np.zeros((len(ids), 3))))
    return dict(frames=frames, gt=gt, sigma_xy=sigma_xy, sigma_yaw=sig_yaw)


def sample_keyframes(n_kfs, areas, first=100):
    """The keyframes whose local areas make up a G4 batch: `areas` keyframes spread evenly over [first, n_kfs) -- the whole
    tour, revisits and loop closures included (with repetition when the tour has fewer keyframes than `areas`)."""
    first = min(first, max(0, n_kfs - 1))
    return np.linspace(first, n_kfs - 1, areas).round().astype(int)


def se2_noise_lambda(sigma_xy, sigma_yaw):
    """Constant information matrix diag(1/sxy^2, 1/sxy^2, 1/syaw^2) (CDatasetParser_RelGraphSLAM2D.h:45-52), row-major 3x3 in a 16-vector."""
    nl = np.zeros(16)
    nl[:9] = np.diag([1 / sigma_xy ** 2, 1 / sigma_xy ** 2, 1 / sigma_yaw ** 2]).ravel()
    return nl


# ------------------------------------------------------------------------------------------------ SE(3) landmark SLAM
def rot_zyx(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def landmark_slam_se3(n_kfs=500, n_lms=5000, seed=1, obs_kind="cartesian3d", step=0.3, max_range=5.0, pixel_sigma=0.5, xyz_sigma=1e-3,
                      cam=(200.0, 200.0, 400.0, 300.0), img=(800, 600), baseline=0.2, lm_init_sigma=None):
    """G2/G3/G5: keyframes on a closed 3-D loop (radius chosen so that consecutive keyframes are `step` apart, gentle
    vertical wave), landmarks uniform in a tube of half-width `max_range` around the path; the sensor looks along the
    robot's +X (sensor pose (0,0,0,-90deg,0,-90deg) for cameras, srba-run-generic-impl.h:79; none for Cartesian_3D).
    Observation models: models/sensors.h (Cartesian_3D :348-364, MonocularCamera :49-66, StereoCamera :174-211).
    Returns frames + ground truth (robot poses as 4x4, landmark positions)."""
    rng = np.random.default_rng(seed)
    radius = n_kfs * step / (2 * math.pi)
    T = []
    for k in range(n_kfs):
        a = 2 * math.pi * k / n_kfs
        pos = np.array([radius * math.cos(a), radius * math.sin(a), 0.3 * math.sin(3 * a)])
        yaw = a + math.pi / 2
        M = np.eye(4)
        M[:3, :3] = rot_zyx(yaw, 0.05 * math.sin(2 * a), 0.03 * math.cos(5 * a))
        M[:3, 3] = pos
        T.append(M)
    # landmarks in a tube around the circle
    ang = rng.uniform(0, 2 * math.pi, n_lms)
    rr = radius + rng.uniform(-max_range, max_range, n_lms)
    lms = np.stack([rr * np.cos(ang), rr * np.sin(ang), rng.uniform(-2.0, 2.0, n_lms)], axis=1)
    is_cam = obs_kind in ("mono", "stereo")
    S = np.eye(4)
    if is_cam:
        S[:3, :3] = rot_zyx(math.radians(-90), 0, math.radians(-90))
    fx, fy, cx, cy = cam
    frames = []
    seen = set()
    for k in range(n_kfs):
        Tc = T[k] @ S
        loc = (np.linalg.inv(Tc) @ np.c_[lms, np.ones(n_lms)].T).T[:, :3]
        dist = np.linalg.norm(loc, axis=1)
        if is_cam:
            z = loc[:, 2]
            with np.errstate(divide="ignore", invalid="ignore"):
                u = cx + fx * loc[:, 0] / z
                v = cy + fy * loc[:, 1] / z
            vis = (z > 0.3) & (dist < max_range) & (u > 0) & (u < img[0]) & (v > 0) & (v < img[1])
        else:
            vis = (loc[:, 0] > 0.2) & (dist < max_range) & (np.abs(np.arctan2(loc[:, 1], loc[:, 0])) < math.radians(60))
        idx = np.nonzero(vis)[0]
        ids, obs, init_flag, init_pos = [], [], [], []
        for i in idx:
            if obs_kind == "cartesian3d":
                o = loc[i] + rng.normal(0, xyz_sigma, 3)
            elif obs_kind == "mono":
                o = np.array([u[i], v[i]]) + rng.normal(0, pixel_sigma, 2)
            else:
                ur = cx + fx * (loc[i, 0] - baseline) / loc[i, 2]
                o = np.array([u[i], v[i], ur, v[i]]) + rng.normal(0, pixel_sigma, 4)
                if o[0] - o[2] <= 0.05:
                    continue
            ids.append(int(i))
            obs.append(o)
            first = int(i) not in seen
            if first and lm_init_sigma is not None:
                rel = (np.linalg.inv(T[k]) @ np.r_[lms[i], 1.0])[:3] + rng.normal(0, lm_init_sigma, 3)
                init_flag.append(1)
                init_pos.append(rel)
            else:
                init_flag.append(0)
                init_pos.append(np.zeros(3))
            seen.add(int(i))
        frames.append(dict(feat_id=np.array(ids, np.uint64), obs=np.array(obs, np.float64).reshape(len(ids), -1),
                           is_fixed=np.zeros(len(ids), np.uint8), is_unknown_with_init_val=np.array(init_flag, np.uint8),
                           feat_rel_pos=np.array(init_pos, np.float64).reshape(len(ids), 3)))
    return dict(frames=frames, gt=np.array(T), lms=lms, sensor_pose_xyzypr=[0, 0, 0, math.radians(-90), 0, math.radians(-90)] if is_cam else None,
                cam=cam, baseline=baseline)


def feed(engine, frame, run_local_optimization=True):
    """engine.define_new_keyframe() with one generated frame (works for the product engine and for the oracle mirror)."""
    return engine.define_new_keyframe(frame["feat_id"], frame["obs"], is_fixed=frame["is_fixed"],
                                      is_unknown_with_init_val=frame["is_unknown_with_init_val"], feat_rel_pos=frame["feat_rel_pos"],
                                      run_local_optimization=run_local_optimization)
