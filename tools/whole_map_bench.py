"""BASELINE configs #2/#3 "single local area" in their whole-map reading: builds an SE(3) landmark map
with the product engine, then optimises EVERY keyframe-to-keyframe edge and landmark in one optimize_local_area(root, 10^6)
problem (one heavy area -> an 8-CTA cluster, block-sparse L D L^t with minimum-degree order on L2-resident tiles).
Usage: python tools/whole_map_bench.py [obs_kind] [n_kfs] [n_lms] [depth] [--cpu]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import srba_b200 as sb
from srba_b200 import _capi as K
from srba_b200 import datasets as D

name = sys.argv[1] if len(sys.argv) > 1 else "cartesian3d"
n_kfs = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n_lms = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
depth = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 3
kind = {"cartesian3d": 1, "mono": 2, "stereo": 3}[name]
ds = D.landmark_slam_se3(n_kfs=n_kfs, n_lms=n_lms, seed=1, obs_kind=name, lm_init_sigma=0.05 if name == "mono" else None)
kw = dict(pose_kind=1, obs_kind=kind, noise_kind=0, noise_std=1e-3 if kind == 1 else 0.5, max_tree_depth=depth, max_optimize_depth=depth, submap_size=10,
          min_obs_to_loop_closure=4)
if kind != 1:
    kw.update(sensor_pose_kind=1, sensor_pose=ds["sensor_pose_xyzypr"], cam_left=ds["cam"], cam_right=ds["cam"], right_cam_pose=[ds["baseline"], 0, 0, 1, 0, 0, 0],
              use_robust_kernel=1)
eng = sb.RbaEngine(sb.default_engine_params(**kw))
t0 = time.perf_counter()
for f in ds["frames"]:
    D.feed(eng, f)
print("map of %d keyframes / %d landmarks built in %.1f s" % (n_kfs, n_lms, time.perf_counter() - t0))
t0 = time.perf_counter()
flat = eng.flatten_local_area(n_kfs - 1, 10 ** 6)
p = flat.p
print("whole-map problem flattened in %.2f s: n_p=%d n_f=%d n_obs=%d blkA=%d blkF=%d n_S=%d S_terms=%d pairs=%d (dense reduced system would be %d^2)" %
      (time.perf_counter() - t0, p.n_unk_edges, p.n_unk_lms, p.n_obs, p.n_blkA, p.n_blkF, p.n_S, p.n_S_terms, p.n_pairs, 6 * p.n_unk_edges))
out = (C.c_int64 * 8)()
t0 = time.perf_counter()
K.lib().srba_b200_problem_plan(flat.problem, out)
cnt = (C.c_int32 * 6)()
K.lib().srba_b200_debug_ldl_schedule(flat.problem, cnt, None, None, None, None, None, None, 0, 0, 800)
print("plan (%.2f s): image %.1f MB, workspace %.1f MB, solver mode %d (0 block-sparse, +2 cluster), tiles %d, updates %d, phases %d" %
      (time.perf_counter() - t0, out[0] / 1e6, out[1] / 1e6, out[4], cnt[1], cnt[2], cnt[4]))
import torch
b = sb.Batch([flat])
b.upload(); b.run(); b.download()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); b.run(); ev[1].record(); torch.cuda.synchronize()
b.download()
st = b.stats(0)
ms = ev[0].elapsed_time(ev[1])
print("whole-map LM on the GPU: %.2f ms, %d trials (%d accepted) -> %.2f ms per LM iteration, %.1f LM it/s; error %.4g -> %.4g, rmse %.3g" %
      (ms, st["num_trials"], st["num_accepted"], ms / max(1, st["num_trials"]), st["num_trials"] / (ms * 1e-3), st["total_sqr_error_init"], st["total_sqr_error_final"],
       st["obs_rmse"]))
if "--cpu" in sys.argv:
    from oracle import pyoracle as po
    eo = po.OracleEngine(po.default_params(**kw))
    for f in ds["frames"]:
        D.feed(eo, f)
    s0, t0_ = eo.opt_timer()
    info = eo.optimize_local_area(n_kfs - 1, 10 ** 6)
    s1, t1_ = eo.opt_timer()
    print("oracle (1 CPU thread) whole-map optimize_local_area: %.2f s, %d trials -> %.1f ms per LM iteration; final error %.4g" %
          (s1 - s0, t1_ - t0_, 1e3 * (s1 - s0) / max(1, t1_ - t0_), info["total_sqr_error_final"]))
