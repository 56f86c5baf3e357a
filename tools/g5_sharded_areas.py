"""BASELINE config #5: SE(3) + Cartesian3D landmarks, one large map (default 10 000 keyframes / 100 000 landmarks,
depth 5) whose local areas -- optimize_local_area(kf, depth) of every `--stride`-th keyframe -- are solved as INDEPENDENT problems with
snapshot semantics (every area reads the pre-batch state) and sharded over the ranks with no data-path collective.
Single process:  python tools/g5_sharded_areas.py [--kfs N --lms M --depth D --stride S --max-areas A]
N GPUs:          python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/g5_sharded_areas.py ...
--plan-only: host side only (map, flattening, memory/solver plan), no device needed."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import srba_b200 as sb
from srba_b200 import _capi as K
from srba_b200 import datasets as D
from srba_b200 import parallel as PL

ap = argparse.ArgumentParser()
ap.add_argument("--kfs", type=int, default=10000)
ap.add_argument("--lms", type=int, default=100000)
ap.add_argument("--depth", type=int, default=5)
ap.add_argument("--stride", type=int, default=10)
ap.add_argument("--max-areas", type=int, default=0)
ap.add_argument("--wave", type=int, default=-1, help="-1 auto | 0 one CTA per area | 1 wave pipeline for every area")
ap.add_argument("--plan-only", action="store_true")
ap.add_argument("--cpu-areas", type=int, default=0, help="also time the oracle (1 thread) on the first K areas of rank 0")
a = ap.parse_args()
K.set_option("wave", a.wave)
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))

ds = D.landmark_slam_se3(n_kfs=a.kfs, n_lms=a.lms, seed=1, obs_kind="cartesian3d")
kw = dict(pose_kind=1, obs_kind=1, noise_kind=0, noise_std=1e-3, max_tree_depth=a.depth, max_optimize_depth=a.depth, submap_size=10, min_obs_to_loop_closure=4)
eng = sb.RbaEngine(sb.default_engine_params(**kw))
t0 = time.perf_counter()
for f in ds["frames"]:
    D.feed(eng, f, run_local_optimization=False)   # graph + initial values only (host); the areas below do all the optimisation
t_map = time.perf_counter() - t0
roots = list(range(a.stride - 1, a.kfs, a.stride))
if a.max_areas:
    roots = roots[:a.max_areas]
mine = [roots[i] for i in PL.shard_indices(len(roots), rank, world)]
t0 = time.perf_counter()
flats = eng.flatten_local_areas(mine, a.depth, max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", world))))   # host threads, engine read-only
flats = [f for f in flats if f is not None]
t_flat = time.perf_counter() - t0
sz = np.array([[f.p.n_unk_edges, f.p.n_unk_lms, f.p.n_obs, f.p.n_blkA, f.p.n_S, f.p.n_S_terms] for f in flats])
plans = []
for f in flats:
    out = (C.c_int64 * 8)(); K.lib().srba_b200_problem_plan(f.problem, out); plans.append(list(out))
pl = np.array(plans)
if rank == 0:
    print("map: %d keyframes / %d landmarks / %d k2k edges / %d observations built in %.1f s (host only)" % (a.kfs, a.lms, eng.num_edges, eng.num_observations, t_map))
    print("rank 0: %d of %d areas (depth %d, every %dth keyframe) flattened in %.1f s" % (len(flats), len(roots), a.depth, a.stride, t_flat))
    for i, nm in enumerate("n_p n_f n_obs n_blkA n_S n_S_terms".split()):
        print("  %-9s median %6d max %6d" % (nm, np.median(sz[:, i]), sz[:, i].max()))
    print("  plan: image %.1f MB, workspace %.1f MB in total; solver modes %s" % (pl[:, 0].sum() / 1e6, pl[:, 1].sum() / 1e6, np.bincount(pl[:, 4].astype(int)).tolist()))
if a.plan_only:
    sys.exit(0)

import torch
import torch.distributed as dist
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    # NCCL prints its version banner on STDOUT when NCCL_DEBUG (env or nccl.conf) is VERSION, and honours NCCL_DEBUG_FILE only above
    # that level: raise VERSION/unset to WARN and send everything to stderr, so that stdout stays the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
b = sb.Batch(flats)
b.upload(); b.run(); b.download()   # warm-up (same inputs: the batch keeps the packed image)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); b.run(); ev[1].record(); torch.cuda.synchronize()
b.download()
st = [b.stats(i) for i in range(b.n)]
trials = sum(s["num_trials"] for s in st)
ms, tot = PL.reduce_timing(ev[0].elapsed_time(ev[1]), trials, device="cuda")
n_areas = len(roots)
if rank == 0:
    e0 = sum(s["total_sqr_error_init"] for s in st); e1 = sum(s["total_sqr_error_final"] for s in st)
    line = dict(config="G5: SE3+Cartesian3D, %d KF / %d LM, depth %d, %d areas sharded over %d GPU(s), snapshot semantics" % (a.kfs, a.lms, a.depth, n_areas, world),
                n_gpus=world, ms=ms, lm_trials=tot, lm_iterations_per_s=tot / (ms * 1e-3), areas_per_s=n_areas / (ms * 1e-3),
                rank0=dict(areas=b.n, trials_mean=trials / b.n, trials_max=max(s["num_trials"] for s in st), err_init=e0, err_final=e1,
                           rmse_median=float(np.median([s["obs_rmse"] for s in st])), h2d_bytes=int(b.h2d_bytes), d2h_bytes=int(b.d2h_bytes)))
    import bench as B   # SURVEY 8(d) unique-touch byte model, evaluated on this rank's exact counts and trial counts
    ab = B.algorithmic_bytes(flats, st)
    peak, peak_src = B.measured_peaks()
    line["roofline_rank0"] = dict(bound="hbm", algorithmic_bytes=ab["total_bytes"], achieved_gbs=ab["total_bytes"] / (ms * 1e-3) / 1e9, peak_gbs=peak,
                                  frac=ab["total_bytes"] / (ms * 1e-3) / 1e9 / peak, peak_source=peak_src,
                                  stage_bytes={k: ab[k] for k in ("chain", "assem_res", "assem_jh", "jacmat", "schur", "solve", "update")}, dense_equiv_chol_flop=ab["chol_flop"])
    print(json.dumps(line))
    if a.cpu_areas:
        from oracle import pyoracle as po
        eo = po.OracleEngine(po.default_params(**kw))
        for f in ds["frames"]:
            D.feed(eo, f, run_local_optimization=False)
        _, poses0 = eo.edges()
        sec, tr = 0.0, 0
        for r in mine[:a.cpu_areas]:
            s0, t0_ = eo.opt_timer()
            eo.optimize_local_area(r, a.depth)
            s1, t1_ = eo.opt_timer()
            sec += s1 - s0; tr += t1_ - t0_
        print("oracle (1 CPU thread), first %d areas of rank 0 (sequential, NOT snapshot semantics): %d LM trials in %.2f s -> %.1f LM it/s" % (a.cpu_areas, tr, sec, tr / max(sec, 1e-9)))
if world > 1:
    dist.destroy_process_group()
