"""Times a batch of SE(3) landmark-SLAM local areas (BASELINE configs #2/#3/#5 shapes) through the staged batch API.
Usage: python tools/se3_area_bench.py [obs_kind cartesian3d|mono|stereo] [n_kfs] [n_lms] [areas] [depth]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import srba_b200 as sb
from srba_b200 import datasets as D

name = sys.argv[1] if len(sys.argv) > 1 else "cartesian3d"
n_kfs = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n_lms = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
areas = int(sys.argv[4]) if len(sys.argv) > 4 else 32
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 3
kind = {"cartesian3d": 1, "mono": 2, "stereo": 3}[name]
ds = D.landmark_slam_se3(n_kfs=n_kfs, n_lms=n_lms, seed=1, obs_kind=name, lm_init_sigma=0.05 if name == "mono" else None)
kw = dict(pose_kind=1, obs_kind=kind, noise_kind=0, noise_std=1e-3 if kind == 1 else 0.5, max_tree_depth=depth, max_optimize_depth=depth, submap_size=10,
          min_obs_to_loop_closure=4)
if kind != 1:
    kw.update(sensor_pose_kind=1, sensor_pose=ds["sensor_pose_xyzypr"], cam_left=ds["cam"], cam_right=ds["cam"], right_cam_pose=[ds["baseline"], 0, 0, 1, 0, 0, 0],
              use_robust_kernel=1)
eng = sb.RbaEngine(sb.default_engine_params(**kw))
flats = []
t0 = time.perf_counter()
for k, f in enumerate(ds["frames"]):
    if k >= n_kfs - areas:
        eng.record_problems(True)
    D.feed(eng, f)
    if k >= n_kfs - areas:
        rec = eng.take_recorded()
        eng.record_problems(False)
        if rec:
            flats.append(rec[-1])
print("sequential run of %d keyframes: %.1f s; %d local areas recorded" % (n_kfs, time.perf_counter() - t0, len(flats)))
a = np.array([[f.p.n_unk_edges, f.p.n_unk_lms, f.p.n_obs, f.p.n_blkA, f.p.n_blkF, f.p.n_S, f.p.n_S_terms, f.p.n_pairs] for f in flats])
for i, nm in enumerate("n_p n_f n_obs n_blkA n_blkF n_S n_S_terms n_pairs".split()):
    print("  %-9s median %6d max %6d" % (nm, np.median(a[:, i]), a[:, i].max()))
import ctypes as C
from srba_b200 import _capi as K
plans = []
for f in flats:
    out = (C.c_int64 * 8)(); K.lib().srba_b200_problem_plan(f.problem, out); plans.append(list(out))
pl = np.array(plans)
print("  plan: smem median %d B, solver modes %s, image %.1f KB, workspace %.1f KB per area" % (np.median(pl[:, 3]), np.bincount(pl[:, 4].astype(int)).tolist(), pl[:, 0].mean() / 1e3, pl[:, 1].mean() / 1e3))
import torch
b = sb.Batch(flats)
b.upload(); b.run(); b.download()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); b.run(); ev[1].record(); torch.cuda.synchronize()
b.download()
st = [b.stats(i) for i in range(b.n)]
trials = sum(s["num_trials"] for s in st)
ms = ev[0].elapsed_time(ev[1])
print("batch of %d %s areas (depth %d): %.3f ms, %d LM trials -> %.0f LM it/s; trials/area mean %.1f max %d; final rmse median %.3g" %
      (b.n, name, depth, ms, trials, trials / (ms * 1e-3), trials / b.n, max(s["num_trials"] for s in st), np.median([s["obs_rmse"] for s in st])))
if "--cpu" in sys.argv:   # the oracle (CPU restatement of the reference) on the same keyframes, 1 thread, 'opt' span only
    from oracle import pyoracle as po
    eo = po.OracleEngine(po.default_params(**kw))
    sec, tr = 0.0, 0
    for k, f in enumerate(ds["frames"]):
        s0, t0_ = eo.opt_timer()
        D.feed(eo, f)
        s1, t1_ = eo.opt_timer()
        if k >= n_kfs - areas:
            sec += s1 - s0; tr += t1_ - t0_
    print("oracle (1 CPU thread) on the same %d keyframes: %d LM trials in %.2f s -> %.0f LM it/s" % (areas, tr, sec, tr / max(sec, 1e-9)))
if "--stages" in sys.argv:
    rows = sorted(((sum(v for k2, v in b.stage_cycles(i).items() if not k2.startswith("chol.")), i) for i in range(b.n)), reverse=True)
    for tot, i in rows[:2]:
        sc = b.stage_cycles(i); tr = max(1, st[i]["num_trials"])
        print("area n_p=%d n_f=%d trials=%d: %.0f kcycles/trial:" % (flats[i].p.n_unk_edges, flats[i].p.n_unk_lms, tr, tot / tr / 1e3), {k2: int(v / tr) for k2, v in sc.items() if v})
