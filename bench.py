#!/usr/bin/env python
"""bench.py — LM iterations/sec of the SRBA local-area optimiser (BASELINE.json metric) on B200.

Workload (config.workload): G4 of SURVEY.md §8(d) — a batch of independent SE(2) relative graph-SLAM local-area
problems (submap 10, depth 3), each one the exact optimize_local_area() input of one keyframe of a world-2d-30k-like
tour, snapshotted from a sequential run of the product engine.  One STEP = one pass of the hot path over the whole batch
(every problem runs its full Levenberg–Marquardt loop: Jacobians + Hessian + Schur + Cholesky + update per trial).
One LM iteration = one solve(lambda) trial (impl/optimize_edges.h:471-476), accepted or rejected.

  value : whole-job LM iterations/sec with the packed inputs already resident in HBM (CUDA events around each launch)
  e2e   : the same through srba_b200_optimize_batch() with HOST buffers (pack + H2D + kernel + D2H + unpack in the timed region)
  roofline / cpu_baseline / clocks : see DESIGN.md §Measurement
Usage: python bench.py --gpus N --steps K --warmup W [--impl reference]
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LM iterations/sec (Jacobian+Hessian+Schur solve) on 30k-LM local area"
UNIT = "LM iterations/s"
TOUR_PARAMS = dict(max_tree_depth=3, max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1, max_error_per_obs_to_stop=1e-8,
                   pose_kind=0, obs_kind=0, noise_kind=1)


def workload_name(areas):
    return ("G4: %d independent SE2 relative graph-SLAM local areas (world-2d-30k-like tour, submap 10, depth 3), each the "
            "optimize_local_area() input of one keyframe" % areas)


def engine_kwargs(ds):
    from srba_b200 import datasets as D
    return dict(TOUR_PARAMS, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(flats, stats):
    """SURVEY.md §8(d) unique-touch model, per launch: every trial pays chain + residual half of assembly + Schur + Cholesky
    bytes; the Jacobian/Hessian half of the assembly is paid once at start and once per relinearisation."""
    tot = dict(chain=0.0, assem_res=0.0, assem_jh=0.0, schur=0.0, chol=0.0, chol_flop=0.0)
    for f, st in zip(flats, stats):
        p = f.p
        P = 3 if p.model.pose_kind == 0 else 6
        S = 24 if P == 3 else 96
        o = {0: 3, 1: 3, 2: 2, 3: 4}[p.model.obs_kind]
        NT, trials, relin = p.n_pairs, st["num_trials"], st["num_relinearized"]
        b_chain = NT * 4 + p.n_path * 5 + p.n_edges * S + 2 * NT * S
        b_res = p.n_obs * (o * 8 + 8) + 2 * NT * S + p.n_lms * 24 + p.n_obs * (o * 8 + 1)
        b_jh = p.n_blkA * 20 + p.n_blkF * 8 + (p.n_HAp * P * P + p.n_HApf * P * 3 + p.n_unk_lms * 9) * 8 + (p.n_unk_edges * P + p.n_unk_lms * 3) * 8
        b_schur = (p.n_HApf * P * 3 + 2 * p.n_unk_lms * 9 + p.n_HAp * P * P) * 8 + 2 * p.n_S * P * P * 8 + (p.n_unk_edges * P + p.n_unk_lms * 3) * 16
        n = P * p.n_unk_edges
        tot["chain"] += b_chain * (trials + 1)
        tot["assem_res"] += b_res * (trials + 1)
        tot["assem_jh"] += b_jh * (relin + 1)
        tot["schur"] += b_schur * trials
        tot["chol"] += n * n * 8.0 * trials
        tot["chol_flop"] += (n ** 3 / 3.0 + 2 * n * n) * trials
    tot["total_bytes"] = tot["chain"] + tot["assem_res"] + tot["assem_jh"] + tot["schur"] + tot["chol"]
    return tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        return out


def build_problems(seed, areas, n_kfs, noise_seed=None):
    """Sequential product-engine run over the tour; returns the flattened local-area problem of the last `areas` keyframes."""
    import srba_b200 as sb
    from srba_b200 import datasets as D
    ds = D.graphslam_se2_tour(n_kfs, seed=seed, noise_seed=noise_seed)
    eng = sb.RbaEngine(sb.default_engine_params(**engine_kwargs(ds)))
    flats = []
    first = n_kfs - areas
    for k, f in enumerate(ds["frames"]):
        if k >= first:
            eng.record_problems(True)
        D.feed(eng, f)
        if k >= first:
            rec = eng.take_recorded()
            eng.record_problems(False)
            assert rec, "keyframe %d produced no optimisation problem" % k
            flats.append(rec[-1])   # the last optimize_edges call of define_new_keyframe is the local area (define_new_keyframe.h:97)
    return ds, eng, flats


def cpu_pass(args):
    """One oracle pass over the tour; returns (opt seconds, LM trials) of the sampled keyframes (every `stride`-th of the
    last keyframes, starting at `offset`). Runs in a worker process."""
    seed, n_kfs, first = args[:3]
    stride, offset = (args[3], args[4]) if len(args) > 3 else (1, 0)
    from oracle import pyoracle as po
    from srba_b200 import datasets as D
    ds = D.graphslam_se2_tour(n_kfs, seed=seed)
    e = po.OracleEngine(po.default_params(**engine_kwargs(ds)))
    sec, trials = 0.0, 0
    for k, f in enumerate(ds["frames"]):
        s0, t0 = e.opt_timer()
        D.feed(e, f)
        s1, t1 = e.opt_timer()
        if k >= first and (k - first) % stride == offset:
            sec += s1 - s0
            trials += t1 - t0
    return sec, trials


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be built here
    (MRPT + Eigen absent), so this times the oracle port, one independent engine per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 16))
    n_kfs, areas = args.ref_kfs, min(args.areas, args.ref_kfs - 40)
    first = n_kfs - areas
    ctx = mp.get_context("spawn")
    per_step = []
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            # the SAME tour/areas as the GPU arm (seed 1), the areas dealt round-robin to the cores
            res = pool.map(cpu_pass, [(1, n_kfs, first, cores, w) for w in range(cores)])
            if step >= args.warmup:
                per_step.append((sum(t for _, t in res), max(s for s, _ in res)))
    trials = sum(t for t, _ in per_step)
    secs = sum(s for _, s in per_step)
    val = trials / secs
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(1, len(per_step)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(areas), "areas_per_gpu": areas, "tour_keyframes": n_kfs, "note": "oracle port of the reference CPU path (the reference needs MRPT+Eigen, absent here)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d single-threaded engines in parallel (one per core) share the same %d local areas (round-robin) of the %d-keyframe tour; 'opt' span only, value = all LM trials / slowest core" % (cores, areas, n_kfs)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="srba_b200")
    ap.add_argument("--areas", type=int, default=256, help="independent local areas per GPU (BASELINE config #4)")
    ap.add_argument("--kfs", type=int, default=700, help="keyframes of the tour the areas are snapshotted from")
    ap.add_argument("--ref-kfs", type=int, default=0, help="tour length of the CPU legs (default: same as --kfs, i.e. the very same problems)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tour-seed", type=int, default=1, help="seed of the tour geometry (1 = the workload every number in profiles/ is quoted on)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.ref_kfs <= 0:
        args.ref_kfs = args.kfs
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import srba_b200 as sb
    from srba_b200 import _capi as K
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or K.lib().srba_b200_device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device — the srba_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG (env or nccl.conf) is VERSION, and honours NCCL_DEBUG_FILE only above
        # that level: raise VERSION/unset to WARN and send everything to stderr, so that stdout stays the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # weak scaling: every rank gets the SAME tour (seed 1: same shape of every local area, i.e. the same work per GPU) with
    # its own observation-noise realisation; rank 0 is bit-identical to the single-GPU run
    ds, eng, flats = build_problems(seed=args.tour_seed, areas=args.areas, n_kfs=args.kfs, noise_seed=None if rank == 0 else 1000 + rank)
    # BASELINE config #1 reading (one sequential srba-slam run): the tour above WAS such a run through the product
    # engine -- every define_new_keyframe a flatten + one-problem srba_b200_optimize call + scatter.  Reported as context only.
    tm = eng.timers()
    seq = {"keyframes": args.kfs, "lm_trials": int(tm["total_trials"]), "gpu_call_seconds": round(tm["gpu_seconds"], 4),
           "flatten_scatter_seconds": round(tm["flatten_seconds"], 4),
           "lm_iterations_per_s": tm["total_trials"] / max(1e-9, tm["gpu_seconds"] + tm["flatten_seconds"])}
    batch = sb.Batch(flats)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    batch.upload(sp)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step():
        flush.fill_(1)                      # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.run(sp)
        e1.record(stream)
        return e0, e1

    torch.cuda.profiler.start()   # cudaProfilerStart: `ncu --profile-from-start off` then skips the set-up launches
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    evs = [step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    batch.download(sp)
    stats = [batch.stats(i) for i in range(batch.n)]
    trials_per_step = sum(s["num_trials"] for s in stats)

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (pack + H2D + kernel + D2H + unpack), every step
    n = len(flats)
    parr = (C.POINTER(K.Problem) * n)(*[f.problem for f in flats])
    bufs = [sb.ResultBuffers(f.p) for f in flats]
    rarr = (C.POINTER(K.Result) * n)(*[C.pointer(b.r) for b in bufs])
    L = K.lib()
    for _ in range(2):
        rc = L.srba_b200_optimize_batch(n, parr, rarr)
        assert rc == 0, L.srba_b200_last_error()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        rc = L.srba_b200_optimize_batch(n, parr, rarr)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    e2e_trials = sum(b.r.stats.num_trials for b in bufs)

    tvec = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    cvec = torch.tensor([float(trials_per_step), float(e2e_trials)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cvec, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max = tvec.tolist()
    trials_all, e2e_trials_all = cvec.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = trials_all * args.steps / (dev_ms_max * 1e-3)
    e2e_value = e2e_trials_all * args.steps / e2e_s_max

    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(flats, stats)
    launch_s = dev_ms * 1e-3 / args.steps
    achieved = ab["total_bytes"] / launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("srba_lm_batch_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.areas), "areas_per_gpu": args.areas, "tour_keyframes": args.kfs, "lm_trials_per_step_per_gpu": trials_per_step,
                       "ranks": "same tour (seed 1) on every rank, independent observation noise per rank",
                       "sequential_run_of_the_tour": seq,
                       "l2": "flushed between timed iterations (256 MiB fill)", "timing": "CUDA events around each launch on the launching stream"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(batch.h2d_bytes), "d2h_bytes_per_step": int(batch.d2h_bytes)},
            "gpu_launches": int(batch.kernel_launches) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "srba_lm_batch_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab["total_bytes"],
                         "stage_bytes": {k: ab[k] for k in ("chain", "assem_res", "assem_jh", "schur", "chol")}, "chol_flop_per_launch": ab["chol_flop"]}}
    if not args.no_cpu_baseline and world == 1:
        first = args.ref_kfs - min(args.areas, args.ref_kfs - 40)
        t = time.perf_counter()
        sec, tr = cpu_pass((1, args.ref_kfs, first))
        line["cpu_baseline"] = {"value": tr / sec, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "oracle (g++ -O2, 1 thread): 'opt' span of the last %d local areas of a %d-keyframe tour with the same generator/parameters (%.1f s wall)"
                                          % (args.ref_kfs - first, args.ref_kfs, time.perf_counter() - t)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
