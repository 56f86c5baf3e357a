"""bench.py --config 2 | 3 | 5: the other BASELINE.json configs behind the same bench contract (one JSON line, rank 0).

  2  SE3 + MonocularCamera, 5 000 landmarks / 500 keyframes, ONE local area holding the whole map (SURVEY 8d, G2 whole-map reading)
  3  SE3 + StereoCamera, 20 000 landmarks / 2 000 keyframes, depth 4, ONE area holding the whole map (G3)
  5  SE3 + Cartesian3D, 100 000 landmarks / 10 000 keyframes, depth 5: the local areas of every 10th keyframe = 1 000 independent problems
     with snapshot semantics, SHARDED over the ranks -- strong scaling, no data-path collective (G5)

A step = one pass of the hot path over the config's problem set (every problem runs its whole LM loop).  N > 1: configs 2/3 do not shard
(one coupled LM recurrence) -> every rank solves a replica ("replicas only", weak); config 5 deals its areas round-robin to the ranks.
`--size-scale s` shrinks the map (keyframes and landmarks x s) for quick runs; the JSON names the actual sizes.
`value` device-timed with the packed problems resident in HBM; `e2e` through srba_b200_optimize_batch with host buffers."""
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

CONFIGS = {
    2: dict(obs="mono", kfs=500, lms=5000, depth=3, whole_map=True, text="SE3 + MonocularCamera, %d landmarks / %d keyframes, ONE local area holding the whole map (BASELINE config #2)"),
    3: dict(obs="stereo", kfs=2000, lms=20000, depth=4, whole_map=True, text="SE3 + StereoCamera, %d landmarks / %d keyframes, depth 4, ONE local area holding the whole map (BASELINE config #3)"),
    5: dict(obs="cartesian3d", kfs=10000, lms=100000, depth=5, whole_map=False,
            text="SE3 + Cartesian3D, %d landmarks / %d keyframes, depth 5: local areas of every 10th keyframe as independent problems, snapshot semantics, sharded over the ranks (BASELINE config #5)"),
}


def main(args):
    import torch
    import torch.distributed as dist
    import srba_b200 as sb
    from srba_b200 import _capi as K
    from srba_b200 import datasets as D
    from srba_b200 import parallel as PL
    import bench as B
    cfg = CONFIGS[args.config]
    scale = float(os.environ.get("SRBA_BENCH_SIZE_SCALE", "1"))
    n_kfs, n_lms = max(30, int(cfg["kfs"] * scale)), max(300, int(cfg["lms"] * scale))
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or K.lib().srba_b200_device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device -- the srba_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind = {"cartesian3d": 1, "mono": 2, "stereo": 3}[cfg["obs"]]
    ds = D.landmark_slam_se3(n_kfs=n_kfs, n_lms=n_lms, seed=1, obs_kind=cfg["obs"], lm_init_sigma=0.05 if cfg["obs"] == "mono" else None)
    kw = dict(pose_kind=1, obs_kind=kind, noise_kind=0, noise_std=1e-3 if kind == 1 else 0.5, max_tree_depth=cfg["depth"], max_optimize_depth=cfg["depth"],
              submap_size=10, min_obs_to_loop_closure=4)
    if kind != 1:
        kw.update(sensor_pose_kind=1, sensor_pose=ds["sensor_pose_xyzypr"], cam_left=ds["cam"], cam_right=ds["cam"], right_cam_pose=[ds["baseline"], 0, 0, 1, 0, 0, 0],
                  use_robust_kernel=1)
    eng = sb.RbaEngine(sb.default_engine_params(**kw))
    t0 = time.perf_counter()
    if cfg["obs"] == "mono":
        # monocular BA is scale-free and landmark_matcher<MonocularCamera> gives no edge initial value, so the map is built with
        # define_new_keyframe(..., run_local_optimization=false), landmark guesses (is_unknown_with_init_val, sigma 5 cm) and perturbed
        # ground-truth edge values (sigma 1 cm / 0.5 deg); the whole-map call below does all the optimisation
        for f in ds["frames"]:
            D.feed(eng, f, run_local_optimization=False)
        rng = np.random.default_rng(7)
        ft, _ = eng.edges()
        for i, (a_, b_) in enumerate(ft.astype(int)):
            Tab = np.linalg.inv(ds["gt"][b_]) @ ds["gt"][a_]        # inv_pose: pose of `from` as seen from `to` (srba_types.h:88)
            w = rng.normal(0, math.radians(0.5), 3)
            Wm = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            R = Tab[:3, :3] @ (np.eye(3) + Wm + 0.5 * Wm @ Wm)
            U, _, Vt = np.linalg.svd(R)
            eng.set_edge_pose(i, np.concatenate([(U @ Vt).ravel(), Tab[:3, 3] + rng.normal(0, 0.01, 3)]))
    else:
        for f in ds["frames"]:
            D.feed(eng, f, run_local_optimization=cfg["whole_map"])   # whole-map stereo: a real sequential SLAM run builds the map first
    t_map = time.perf_counter() - t0
    t0 = time.perf_counter()
    if cfg["whole_map"]:
        flats = [eng.flatten_local_area(n_kfs - 1, 10 ** 6)]
        n_units = 1
    else:
        roots = list(range(9, n_kfs, 10))
        n_units = len(roots)
        # snapshot semantics: the areas of one map state, flattened by this rank's share of the host cores (srba_engine_flatten_local_areas)
        mine = [roots[i] for i in PL.shard_indices(len(roots), rank, world)]
        flats = [f for f in eng.flatten_local_areas(mine, cfg["depth"], max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", world)))) if f is not None]
    t_flat = time.perf_counter() - t0
    batch = sb.Batch(flats)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    batch.upload(sp)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    B.time_batch(batch, stream, flush, 0, args.warmup, torch)
    if world > 1:
        dist.barrier()
    sampler = B.ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    dev_ms, _ = B.time_batch(batch, stream, flush, args.steps, 0, torch)
    if world > 1:
        dist.barrier()
    batch.download(sp)
    stats = [batch.stats(i) for i in range(batch.n)]
    trials = sum(s["num_trials"] for s in stats)
    launches = int(batch.kernel_launches)
    # e2e through the C ABI with host buffers
    n = len(flats)
    parr = (C.POINTER(K.Problem) * n)(*[f.problem for f in flats])
    bufs = [sb.ResultBuffers(f.p) for f in flats]
    rarr = (C.POINTER(K.Result) * n)(*[C.pointer(b_.r) for b_ in bufs])
    L = K.lib()
    assert L.srba_b200_optimize_batch(n, parr, rarr) == 0, L.srba_b200_last_error()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        L.srba_b200_optimize_batch(n, parr, rarr)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    e2e_trials = sum(b_.r.stats.num_trials for b_ in bufs)
    (dev_ms_max, e2e_s_max), (trials_all, e2e_trials_all) = PL.reduce_max_sum([dev_ms, e2e_s], [trials, e2e_trials], device="cuda")
    if rank == 0:
        peak, peak_src = B.measured_peaks()
        ab = B.algorithmic_bytes(flats, stats)
        launch_s = dev_ms * 1e-3 / args.steps
        sz = {k: {"median": int(np.median(v)), "max": int(np.max(v))} for k, v in
              (("n_p", [f.p.n_unk_edges for f in flats]), ("n_f", [f.p.n_unk_lms for f in flats]), ("n_obs", [f.p.n_obs for f in flats]),
               ("n_blkA", [f.p.n_blkA for f in flats]), ("n_S_terms", [f.p.n_S_terms for f in flats]))}
        plans = np.array([batch.plan(i) for i in range(batch.n)])
        line = {"metric": B.METRIC, "value": trials_all * args.steps / (dev_ms_max * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak" if cfg["whole_map"] else "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": cfg["text"] % (n_lms, n_kfs), "problems_total": n_units if not cfg["whole_map"] else world, "problems_rank0": n,
                           "multi_gpu": "replicas only (one coupled LM recurrence does not shard)" if cfg["whole_map"] else "areas dealt round-robin to the ranks, no data-path collective",
                           "lm_trials_per_step_rank0": trials, "ms_per_lm_iteration_rank0": dev_ms / args.steps / max(1, trials) if cfg["whole_map"] else None,
                           "sizes_rank0": sz, "execution_modes_rank0 (solver + 2 cluster + 4 wave)": np.bincount(plans[:, 4].astype(int)).tolist(),
                           "map_build_seconds": round(t_map, 2), "flatten_seconds": round(t_flat, 2),
                           "error_rank0": {"init": sum(s["total_sqr_error_init"] for s in stats), "final": sum(s["total_sqr_error_final"] for s in stats)},
                           "l2": "flushed between timed iterations (256 MiB fill)", "timing": "CUDA events around each step on the launching stream",
                           "kernel_launches_per_step": launches},
                "e2e": {"value": e2e_trials_all * e2e_steps / e2e_s_max, "unit": B.UNIT, "h2d_bytes_per_step": int(batch.h2d_bytes), "d2h_bytes_per_step": int(batch.d2h_bytes)},
                "gpu_launches": launches * args.steps, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "all kernels of a step", "achieved": ab["total_bytes"] / launch_s / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": ab["total_bytes"] / launch_s / 1e9 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab["total_bytes"],
                             "stage_bytes": {k: ab[k] for k in ("chain", "assem_res", "assem_jh", "jacmat", "schur", "solve", "update")},
                             "dense_equiv_chol_flop": ab["chol_flop"], "solve_flop_per_launch": ab["solve_flop"]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
