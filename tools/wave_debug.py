"""Debug aid: one stereo / cartesian3d / se2 problem through the wave pipeline (SRBA_WAVE_DEBUG=1 prints the views)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import srba_b200 as sb
from srba_b200 import _capi as K, datasets as D
name = sys.argv[1] if len(sys.argv) > 1 else "stereo"
K.set_option("wave", 1)
ds = D.landmark_slam_se3(n_kfs=16, n_lms=160, seed=9, obs_kind=name)
kind = {"cartesian3d": 1, "mono": 2, "stereo": 3}[name]
kw = dict(pose_kind=1, obs_kind=kind, noise_kind=0, noise_std=1e-3 if kind == 1 else 0.5, max_tree_depth=3, max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=4)
if kind != 1:
    kw.update(sensor_pose_kind=1, sensor_pose=ds["sensor_pose_xyzypr"], cam_left=ds["cam"], cam_right=ds["cam"],
              right_cam_pose=[ds["baseline"], 0, 0, 1, 0, 0, 0], use_robust_kernel=1)
ep = sb.RbaEngine(sb.default_engine_params(**kw))
for f in ds["frames"]:
    D.feed(ep, f, run_local_optimization=False)
flat = ep.flatten_local_area(15, 3)
print(flat.sizes(), flush=True)
rw = sb.optimize(flat, trace_cap=32)
print("wave:", rw.stats, flush=True)
K.set_option("wave", 0)
rc = sb.optimize(flat, trace_cap=32)
print("cta :", rc.stats, flush=True)
tw, dw = rw.trace_rows(); tc, dc = rc.trace_rows()
print(np.abs(np.asarray(dw[0]) - np.asarray(dc[0])).max() / np.abs(dc[0]).max())
