"""Helper for test_gpu_parity.py: solves a fixed set of flattened SE(2)/SE(3) problems through the C ABI and prints the
results as JSON (run as a subprocess so that SRBA_B200_SOLVER, which the library reads once, can differ between runs)."""
import json
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), ".."))
import srba_b200 as sb
from srba_b200 import datasets as D


def problems():
    ds = D.graphslam_se2_tour(70, seed=9)
    kw = dict(pose_kind=0, obs_kind=0, noise_kind=1, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]), max_tree_depth=3,
              max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1, max_error_per_obs_to_stop=1e-8)
    e = sb.RbaEngine(sb.default_engine_params(**kw))
    e.record_problems(True)
    for f in ds["frames"]:
        D.feed(e, f)
    out = [f for f in e.take_recorded() if f.p.n_unk_edges >= 3][-6:]
    ds = D.landmark_slam_se3(n_kfs=16, n_lms=120, seed=3, obs_kind="cartesian3d")
    kw = dict(pose_kind=1, obs_kind=1, noise_kind=0, noise_std=1e-3, max_tree_depth=3, max_optimize_depth=3, submap_size=5, min_obs_to_loop_closure=4)
    e = sb.RbaEngine(sb.default_engine_params(**kw))
    e.record_problems(True)
    for f in ds["frames"]:
        D.feed(e, f)
    out += [f for f in e.take_recorded() if f.p.n_unk_edges >= 2][-3:]
    return out


if __name__ == "__main__":
    res = []
    for fl in problems():
        r = sb.optimize(fl, trace_cap=4)
        res.append(dict(n_unk=int(fl.p.n_unk_edges), trials=int(r.stats["num_trials"]), err=float(r.stats["total_sqr_error_final"]),
                        first_delta=np.asarray(r.trace_delta)[0].tolist() if r.stats["num_trials"] > 0 else [], edge=np.asarray(r.edge_pose).ravel().tolist()))
    print(json.dumps(res))
