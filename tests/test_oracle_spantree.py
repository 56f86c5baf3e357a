"""Pins the oracle's incremental symbolic spanning tree, numeric pose-chain composition and BFS helpers against
each other and against ground truth, as the reference's tests/spantree_unittest.cpp:38-245 does (1e-6 at :208,216)."""
import numpy as np
import pytest

from oracle import pyoracle as po


def ypr_vec(p12):
    R = p12[:9].reshape(3, 3)
    pitch = np.arctan2(-R[2, 0], np.hypot(R[0, 0], R[1, 0]))
    roll = np.arctan2(R[2, 1], R[2, 2])
    yaw = np.arctan2(R[1, 0], R[0, 0])
    return np.array([p12[9], p12[10], p12[11], yaw, pitch, roll])


def pose_diff(a, b):
    """sum |x,y,z,yaw,pitch,roll| difference with angle wrap (getAsVectorVal() comparison of the reference)."""
    d = ypr_vec(a) - ypr_vec(b)
    d[3:] = (d[3:] + np.pi) % (2 * np.pi) - np.pi
    return np.abs(d).sum()


def run_topology(topo, nKFs, max_depth, seed):
    rng = np.random.default_rng(seed)
    p = po.default_params(pose_kind=po.POSE_SE3, obs_kind=po.OBS_MONO, max_tree_depth=max_depth)
    e = po.OracleEngine(p)
    GT = [po.pose_from_xyzypr(np.concatenate([rng.uniform(-10, 10, 3), [rng.uniform(-np.pi, np.pi), rng.uniform(-0.5 * np.pi, 0.5 * np.pi),
                                                                           rng.uniform(-np.pi, np.pi)]])) for _ in range(nKFs)]
    rel = lambda a, b: po.pose_compose(po.POSE_SE3, po.pose_inverse(po.POSE_SE3, b), a)  # a - b
    for kf in range(nKFs):
        new_kf = e.alloc_keyframe()
        if new_kf == 0:
            continue
        new_edges = []
        if topo in (0,):
            new_edges.append((new_kf - 1, new_kf))
        elif topo == 100:
            new_edges.append((new_kf, new_kf - 1))
        else:
            is_inv = topo > 100
            new_edges.append((new_kf - 1, new_kf))
            if new_kf > 2:
                while rng.uniform() < 0.05:
                    i = int(rng.integers(0, new_kf - 1))
                    if not e.are_keyframes_connected(i, new_kf) and (i, new_kf) not in new_edges and (new_kf, i) not in new_edges:
                        new_edges.append((new_kf, i) if is_inv else (i, new_kf))
        for (a, b) in new_edges:
            e.create_kf2kf_edge(new_kf, a, b, rel(GT[a], GT[b]))
    e.update_numeric_all()
    for kf in range(nKFs):
        ids, poses = e.complete_spanning_tree(kf, max_depth)
        t, nx, d = e.next_edge(kf)
        assert len(ids) >= 1
        assert len(t) == len(ids) - 1
        nt, nposes, upd = e.num(kf)
        num = {int(a): nposes[i] for i, a in enumerate(nt)}
        dist = {int(a): int(d[i]) for i, a in enumerate(t)}
        for i, dst in enumerate(ids):
            dst = int(dst)
            if dst == kf:
                continue
            plen = e.find_path_bfs_len(kf, dst)
            assert plen > 0
            assert dst in dist and dist[dst] == plen
            assert dst in num
            assert pose_diff(poses[i], num[dst]) < 1e-6
            assert pose_diff(rel(GT[dst], GT[kf]), num[dst]) < 1e-6


@pytest.mark.parametrize("topo", [0, 100, 1, 101])
def test_spantree_topologies(topo):
    for N in (10, 50, 300):
        for depth in range(1, 5):
            seeds = range(1, 10) if N < 300 else range(1, 3)  # the 300-KF cases are trimmed to keep the CPU suite short
            for seed in seeds:
                run_topology(topo, N, depth, seed)
