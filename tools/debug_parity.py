"""Debug helper: runs the oracle and the GPU engine side by side and prints the first diverging optimize call."""
import sys

import numpy as np

sys.path.insert(0, ".")
import srba_b200 as sb
from oracle import pyoracle as po
from srba_b200 import datasets as D

np.set_printoptions(linewidth=220, precision=9)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
model = sys.argv[2] if len(sys.argv) > 2 else "se2"
if model == "se2":
    ds = D.graphslam_se2_tour(n, seed=7)
    kw = dict(pose_kind=0, obs_kind=0, noise_kind=1, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]), max_tree_depth=3,
              max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1, max_error_per_obs_to_stop=1e-8)
else:
    ok = {"cartesian3d": 1, "mono": 2, "stereo": 3}[model]
    ds = D.landmark_slam_se3(n_kfs=n, n_lms=300, seed=2, obs_kind=model, lm_init_sigma=0.05 if model == "mono" else None)
    kw = dict(pose_kind=1, obs_kind=ok, noise_kind=0, noise_std=1e-3 if ok == 1 else 0.5, max_tree_depth=3, max_optimize_depth=3, submap_size=10,
              min_obs_to_loop_closure=4)
    if ok != 1:
        kw.update(sensor_pose_kind=1, sensor_pose=ds["sensor_pose_xyzypr"], cam_left=ds["cam"], cam_right=ds["cam"],
                  right_cam_pose=[ds["baseline"], 0, 0, 1, 0, 0, 0], use_robust_kernel=1)
eo, ep = po.OracleEngine(po.default_params(**kw)), sb.RbaEngine(sb.default_engine_params(**kw))
eo.trace_enable(True)
ep.trace(64)
for k, f in enumerate(ds["frames"]):
    eo.trace_clear()
    ro, rp = D.feed(eo, f), D.feed(ep, f)
    tr = eo.traces()
    a, b = ro["optimize_results"], rp["optimize_results"]
    bad = any(a[x] != b[x] for x in ("num_observations",))
    if tr and len(tr[-1]["trials"]):
        tp_, _ = ep.last_trace()
        m_ = min(len(tp_), len(tr[-1]["trials"]))
        bad = bad or np.abs(tp_[:m_, 3] - tr[-1]["trials"][:m_, 3]).max() > 1e-6 * np.abs(tr[-1]["trials"][:m_, 3]).max()
    if tr and len(tr[-1]["trials"]):
        to = tr[-1]
        tp, dp = ep.last_trace()
        m = min(len(tp), len(to["trials"]))
        scale = max(1e-30, np.abs(to["deltas"]).max())
        dd = max([np.abs(dp[j] - to["deltas"][j]).max() / max(scale, np.abs(to["deltas"][j]).max()) for j in range(m) if to["trials"][j, 1]] or [0])
    else:
        dd = 0
    if bad or dd > 1e-6 or k % 10 == 0:
      print("kf", k, "trials o/g", a["num_lm_trials"], b["num_lm_trials"], "err_init", a["total_sqr_error_init"], b["total_sqr_error_init"], "final",
          a["total_sqr_error_final"], b["total_sqr_error_final"], "max rel delta diff %.2e" % dd)
    if bad or dd > 1e-6:
        print("ORACLE trace:\n", to["trials"], "\nlambda0", to["lambda0"], "\ninit grad", to["init_grad"][:12])
        print("GPU trace:\n", tp)
        print("oracle delta0", to["deltas"][0][:12], "\ngpu delta0   ", dp[0][:12] if len(dp) else None)
        break
