#!/usr/bin/env python
"""bench.py — LM iterations/sec of the SRBA local-area optimiser (BASELINE.json metric) on B200.

Workload (config.workload): G4 of SURVEY.md §8(d) on the G1 dataset shape — a batch of independent SE(2) relative graph-SLAM
local-area problems (submap 10, depth 3), each one the exact optimize_local_area() input of one keyframe of a
world-2d-30k-like tour of 3 000 keyframes, snapshotted from a sequential run of the product engine.  The areas are
sampled EVENLY OVER THE WHOLE TOUR, i.e. with the revisits and loop closures a 6 km tour of a 3x3 grid of 100 m blocks
produces (n_p up to ~250 unknown edges; the distribution is printed in config.area_sizes).
One STEP = one pass of the hot path over the whole batch (every problem runs its full Levenberg–Marquardt loop:
Jacobians + Hessian + Schur + Cholesky + update per trial).  One LM iteration = one solve(lambda) trial
(impl/optimize_edges.h:471-476), accepted or rejected.

  value : whole-job LM iterations/sec with the packed inputs already resident in HBM (CUDA events around each step)
  e2e   : the same through srba_b200_optimize_batch() with HOST buffers (pack + H2D + kernel + D2H + unpack in the timed region)
  roofline / cpu_baseline / clocks : see DESIGN.md §Measurement
  config.chain_like_case   : the round-1 workload (last 256 keyframes of a 700-keyframe tour: chain-like areas, n_p ~ 27)
  config.areas_sweep       : throughput at 1 024 / 4 096 areas per GPU (N=1 only)
  config.strong_scaling_256_areas_total : BASELINE config #4 read literally (256 areas IN TOTAL over the N ranks)
Other BASELINE configs: `--config 2|3|5` (SE(3) mono / stereo whole-map single area, SE(3) Cartesian3D sharded depth-5 areas).
Usage: python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 4]
"""
import argparse
import ctypes as C
import json
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
FIXTURE = os.path.join(ROOT, "tests", "golden", "g4_tour3000_states.npz")


def workload_name(areas, kfs):
    return ("G4: %d independent SE2 relative graph-SLAM local areas sampled evenly over a %d-keyframe world-2d-30k-like tour (G1: submap 10, "
            "depth 3, revisits and loop closures included), each the optimize_local_area() input of one keyframe" % (areas, kfs))


def engine_kwargs(ds):
    from srba_b200 import datasets as D
    return dict(TOUR_PARAMS, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ldl_counts(flat):
    """(tiles, updates) of the block-sparse L D L^t schedule the packer ships with this problem (host symbolic factorisation)."""
    from srba_b200 import _capi as K
    cnt = (C.c_int64 * 4)()
    K.lib().srba_b200_problem_ldl_counts(flat.problem, cnt)
    return int(cnt[0]), int(cnt[1])


def algorithmic_bytes(flats, stats):
    """Unique-touch byte model of SURVEY.md §8(d), per launch, charging only what a trial really touches: the pose pairs
    that hold a listed pose (update_numeric(set,true), not all pairs), the tiles of the filled block-sparse pattern (not a
    dense n^2 system), no Schur stage when there are no landmark unknowns; the Jacobian/Hessian half of the assembly is paid
    once at the start and once per relinearisation, incl. the materialised Jacobians (written, re-read: B_jacmat)."""
    tot = dict(chain=0.0, assem_res=0.0, assem_jh=0.0, jacmat=0.0, schur=0.0, solve=0.0, update=0.0, dense_equiv_chol=0.0, chol_flop=0.0, solve_flop=0.0)
    for f, st in zip(flats, stats):
        p = f.p
        P = 3 if p.model.pose_kind == 0 else 6
        S = 24 if P == 3 else 96
        o = {0: 3, 1: 3, 2: 2, 3: 4}[p.model.obs_kind]
        trials, relin = st["num_trials"], st["num_relinearized"]
        sa, sd = f.array("blkA_slotA", p.n_blkA, np.int64), f.array("blkA_slotD", p.n_blkA, np.int64)
        sf = f.array("blkF_slot", p.n_blkF, np.int64) if p.n_blkF else np.zeros(0, np.int64)
        listed = np.unique(np.concatenate([sa[sa >= 0], sd, sf[sf >= 0]]))
        refresh = np.unique(listed // 2)
        ptr = f.array("pair_path_ptr", p.n_pairs + 1, np.int64)
        plen = np.diff(ptr)
        osl = f.array("obs_slot", p.n_obs, np.int64)
        n_obs_slots = len(np.unique(osl[osl >= 0]))
        tiles, updates = ldl_counts(f)
        nsc = p.n_unk_edges * P + p.n_unk_lms * 3
        b_chain_all = p.n_pairs * 4 + p.n_path * 5 + p.n_edges * S + 2 * p.n_pairs * S
        b_chain_trial = len(refresh) * 4 + int(plen[refresh].sum()) * 5 + p.n_edges * S + 2 * len(refresh) * S
        b_res = p.n_obs * (o * 8 + 8) + n_obs_slots * S + (p.n_lms * 24 if P == 6 else 0) + p.n_obs * (o * 8 + 1)
        b_jh = p.n_blkA * 20 + p.n_blkF * 8 + len(listed) * S + (p.n_HAp * P * P + p.n_HApf * P * 3 + p.n_unk_lms * 9) * 8 + nsc * 8
        b_jm = 2 * (p.n_blkA * o * P + p.n_blkF * o * 3) * 8
        b_schur = 0 if p.n_unk_lms == 0 else ((p.n_HApf * P * 3 + 2 * p.n_unk_lms * 9 + p.n_HAp * P * P) * 8 + 2 * p.n_S * P * P * 8 + nsc * 16)
        b_solve = 2 * tiles * 72 + 2 * P * p.n_unk_edges * 8
        b_upd = 3 * p.n_unk_edges * S + 2 * len(listed) * S + 3 * p.n_unk_lms * 24
        n = P * p.n_unk_edges
        tot["chain"] += b_chain_all + b_chain_trial * trials
        tot["assem_res"] += b_res * (trials + 1)
        tot["assem_jh"] += b_jh * (relin + 1)
        tot["jacmat"] += b_jm * (relin + 1)
        tot["schur"] += b_schur * trials
        tot["solve"] += b_solve * trials
        tot["update"] += b_upd * trials
        tot["dense_equiv_chol"] += n * n * 8.0 * trials
        tot["chol_flop"] += (n ** 3 / 3.0 + 2 * n * n) * trials
        tot["solve_flop"] += (updates * 81.0 + tiles * 27.0) * trials   # 27 FMA per tile update + 27 for W on the fly, 27 per W tile
    tot["total_bytes"] = sum(tot[k] for k in ("chain", "assem_res", "assem_jh", "jacmat", "schur", "solve", "update"))
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


def build_problems(seed, areas, n_kfs, noise_seed=None, last=False, extra_samples=()):
    """Sequential product-engine run over the tour (every define_new_keyframe a flatten + one-problem GPU call + scatter);
    returns the flattened local-area problems of `areas` keyframes sampled evenly over the whole tour (last=True: the last
    `areas` keyframes, the chain-like round-1 workload).  extra_samples: further batch sizes to sample (areas sweep)."""
    import srba_b200 as sb
    from srba_b200 import datasets as D
    ds = D.graphslam_se2_tour(n_kfs, seed=seed, noise_seed=noise_seed)
    eng = sb.RbaEngine(sb.default_engine_params(**engine_kwargs(ds)))
    picks = {areas: (np.arange(n_kfs - areas, n_kfs) if last else D.sample_keyframes(n_kfs, areas))}
    for a in extra_samples:
        picks[a] = D.sample_keyframes(n_kfs, a)
    wanted = set(int(k) for v in picks.values() for k in v)
    by_kf = {}
    for k, f in enumerate(ds["frames"]):
        if k in wanted:
            eng.record_problems(True)
        D.feed(eng, f)
        if k in wanted:
            rec = eng.take_recorded()
            eng.record_problems(False)
            assert rec, "keyframe %d produced no optimisation problem" % k
            by_kf[k] = rec[-1]   # the last optimize_edges call of define_new_keyframe is the local area (define_new_keyframe.h:97)
    sets = {a: [by_kf[int(k)] for k in v] for a, v in picks.items()}
    return ds, eng, sets


# ---------------------------------------------------------------------------------------------------------- CPU legs (oracle)
def cpu_fixture_worker(args):
    """Oracle on a share of the fixture's areas, ONE pass: the graph is rebuilt from the tour's frames (no optimisation), at each of
    this worker's keyframes the touched edges get the stored pre-optimisation poses and optimize_local_area(k, 3) -- the problem of
    the batch -- is timed ('opt' span, impl/optimize_edges.h:56,786).  Returns per area, in keyframe order: (seconds, LM trials, n_p)."""
    mine = sorted(set(args))
    from oracle import pyoracle as po
    from srba_b200 import datasets as D
    fx = np.load(FIXTURE)
    ds = D.graphslam_se2_tour(int(fx["kfs"]), seed=int(fx["seed"]))
    e = po.OracleEngine(po.default_params(**engine_kwargs(ds)))
    at = {int(fx["kf"][i]): i for i in mine}
    last = max(at) if at else -1
    out = []
    for k, f in enumerate(ds["frames"]):
        if k > last:
            break
        D.feed(e, f, run_local_optimization=False)
        if k not in at:
            continue
        i = at[k]
        ids = fx["edge_id"][fx["edge_ptr"][i]:fx["edge_ptr"][i + 1]]
        poses = fx["edge_in"][fx["edge_ptr"][i]:fx["edge_ptr"][i + 1]]
        for eid, q in zip(ids, poses):
            e.set_edge_pose(int(eid), q)
        s0, t0 = e.opt_timer()
        e.optimize_local_area(k, 3)
        s1, t1 = e.opt_timer()
        out.append((s1 - s0, t1 - t0, int(fx["n_p"][i]) if "n_p" in fx else -1))
    return out


def cpu_sample(n_areas, workers, per_worker):
    """Bounded sample of the batch for the CPU legs: workers x per_worker areas evenly spaced over the batch (with wrap-around when
    that exceeds the batch), dealt round-robin; every worker solves its areas once, in keyframe order, concurrently with the others.
    Returns a list (per worker) of per-area (seconds, trials, n_p)."""
    import multiprocessing as mp
    m = workers * per_worker
    idx = [int(j * n_areas / m) % n_areas for j in range(m)]
    shares = [sorted(set(idx[w::workers])) for w in range(workers)]
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(cpu_fixture_worker, shares)


def cpu_pass(args):
    """Fallback when the workload is not the fixture's (other --kfs / --areas / --tour-seed): one oracle pass over the whole
    tour WITH optimisation; returns (opt seconds, LM trials) of this worker's sampled keyframes."""
    seed, n_kfs, sampled = args
    from oracle import pyoracle as po
    from srba_b200 import datasets as D
    ds = D.graphslam_se2_tour(n_kfs, seed=seed)
    e = po.OracleEngine(po.default_params(**engine_kwargs(ds)))
    sampled = set(sampled)
    sec, trials = 0.0, 0
    for k, f in enumerate(ds["frames"]):
        s0, t0 = e.opt_timer()
        D.feed(e, f)
        s1, t1 = e.opt_timer()
        if k in sampled:
            sec += s1 - s0
            trials += t1 - t0
    return sec, trials


def fixture_matches(args):
    if not os.path.exists(FIXTURE):
        return False
    fx = np.load(FIXTURE)
    from srba_b200 import datasets as D
    return (int(fx["kfs"]) == args.kfs and int(fx["seed"]) == args.tour_seed and not args.chain_like
            and fx["kf"].tolist() == D.sample_keyframes(args.kfs, args.areas).tolist())


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be built here
    (MRPT + Eigen absent), so this times the oracle port, one independent single-threaded engine per host core
    (BASELINE.md §2.3).  A STEP is a bounded sample of the batch: one local area per core (the areas of all steps are spaced evenly
    over the 256-area batch, loop-closure areas included); value = LM trials / (core-seconds / cores), i.e. the CPU throughput with
    every core busy and perfect load balance."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from srba_b200 import datasets as D
    cores = max(1, len(os.sched_getaffinity(0)))
    reps = args.warmup + args.steps
    if fixture_matches(args):
        res = cpu_sample(args.areas, cores, reps)
        sec = np.zeros(reps)
        tr = np.zeros(reps)
        for wi, w in enumerate(res):
            # a worker solves its areas in keyframe order (the graph grows with the tour); which STEP an area counts for is shuffled, so that
            # the warm-up steps are not systematically the early -- small -- areas of the tour
            label = np.random.RandomState(1234 + wi).permutation(reps)
            for j, (s_, t_, _) in enumerate(w):
                sec[label[j % reps]] += s_
                tr[label[j % reps]] += t_
        sec, tr = sec[args.warmup:], tr[args.warmup:]
        val = float(cores * tr.sum() / sec.sum())
        ms = 1e3 * float(sec.mean() / cores)
        n_p = [q for w in res for (_, _, q) in w]
        sample_txt = ("%d single-threaded oracle engines in parallel (one per host core); each step = one local area per core, the %d areas of all steps spaced "
                      "evenly over the %d-area batch of the %d-keyframe tour (n_p %d..%d, median %d; inputs: tests/golden/g4_tour3000_states.npz), each solved "
                      "once; 'opt' span only; value = cores x LM trials / core-seconds of the timed steps (perfect load balance)"
                      % (cores, len(n_p), args.areas, args.kfs, min(n_p), max(n_p), int(np.median(n_p))))
        used = cores
    elif args.kfs <= 1000:
        sampled = (np.arange(args.kfs - args.areas, args.kfs) if args.chain_like else D.sample_keyframes(args.kfs, args.areas)).tolist()
        per_step = []
        with mp.get_context("spawn").Pool(cores) as pool:
            for step in range(reps):
                res = pool.map(cpu_pass, [(args.tour_seed, args.kfs, sampled[w::cores]) for w in range(cores)])
                if step >= args.warmup:
                    per_step.append((sum(t for _, t in res), max(s for s, _ in res)))
        val = sum(t for t, _ in per_step) / sum(s for _, s in per_step)
        ms = 1e3 * sum(s for _, s in per_step) / max(1, len(per_step))
        sample_txt = ("%d single-threaded oracle engines in parallel (one per core) each run the whole %d-keyframe tour and time their share of the %d sampled "
                      "local areas; 'opt' span only; value = all LM trials / slowest core" % (cores, args.kfs, args.areas))
        used = cores
    else:
        print(json.dumps({"impl": "reference", "unavailable": "no oracle fixture for --kfs %d --areas %d --tour-seed %d (a sequential oracle run of such a tour takes hours): "
                          "run tests/golden/make_g4_states.py" % (args.kfs, args.areas, args.tour_seed)}))
        return
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.areas, args.kfs), "areas_per_gpu": args.areas, "tour_keyframes": args.kfs,
                       "note": "oracle port of the reference CPU path (the reference needs MRPT+Eigen, absent here)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port", "sample": sample_txt},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------- GPU arm
def time_batch(batch, stream, flush, steps, warmup, torch):
    """`steps` timed passes of a resident batch, L2 flushed before each; returns (total device ms, per-step ms list)."""
    sp = stream.cuda_stream
    for _ in range(warmup):
        flush.fill_(1)
        batch.run(sp)
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.fill_(1)                      # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.run(sp)
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in evs]
    return sum(per), per


def size_stats(flats):
    out = {}
    for name, get in (("n_p", lambda p: p.n_unk_edges), ("n_obs_rows", lambda p: p.n_total_obs_rows), ("n_blkA", lambda p: p.n_blkA),
                      ("n_pairs", lambda p: p.n_pairs), ("n_HAp", lambda p: p.n_HAp)):
        v = np.array([get(f.p) for f in flats])
        out[name] = {"min": int(v.min()), "median": int(np.median(v)), "p95": int(np.percentile(v, 95)), "max": int(v.max())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="srba_b200")
    ap.add_argument("--config", type=int, default=4, help="BASELINE config: 4 (default; also #1's dataset) | 2 | 3 | 5 (see tools/bench_configs.py)")
    ap.add_argument("--areas", type=int, default=256, help="independent local areas per GPU (BASELINE config #4)")
    ap.add_argument("--kfs", type=int, default=3000, help="keyframes of the tour the areas are snapshotted from (SURVEY G1: 3 000)")
    ap.add_argument("--chain-like", action="store_true", help="round-1 workload: the LAST `areas` keyframes of the tour (use with --kfs 700)")
    ap.add_argument("--sweep", default="1024,4096", help="further batch sizes (areas per GPU) timed at N=1; '' to skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the chain-like case, the areas sweep and the strong-scaling reading")
    ap.add_argument("--tour-seed", type=int, default=1, help="seed of the tour geometry")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.config != 4:
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the oracle needs minutes to hours per LM iteration on the whole-map / 1 000-area problems of config "
                                  "%d (profiles/r01_v7_whole_map_*.txt, r01_v8_g5_sharded_areas.txt hold one-off CPU timings)" % args.config}))
            return
        return run_reference(args)
    if args.config != 4:
        from tools import bench_configs
        return bench_configs.main(args)

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

    extras = not args.no_extras
    sweep = [int(x) for x in args.sweep.split(",") if x] if (extras and world == 1 and not args.chain_like) else []
    # weak scaling: every rank gets the SAME tour (same shape of every local area, i.e. the same work per GPU) with its own
    # observation-noise realisation; rank 0 is bit-identical to the single-GPU run
    t_build = time.perf_counter()
    ds, eng, sets = build_problems(seed=args.tour_seed, areas=args.areas, n_kfs=args.kfs, noise_seed=None if rank == 0 else 1000 + rank,
                                   last=args.chain_like, extra_samples=sweep)
    t_build = time.perf_counter() - t_build
    flats = sets[args.areas]
    # BASELINE config #1 reading (one sequential srba-slam run): the tour above WAS such a run through the product
    # engine -- every define_new_keyframe a flatten + one-problem srba_b200_optimize call + scatter.  Reported as context only.
    tm = eng.timers()
    seq = {"keyframes": args.kfs, "lm_trials": int(tm["total_trials"]), "gpu_call_seconds": round(tm["gpu_seconds"], 4),
           "flatten_scatter_seconds": round(tm["flatten_seconds"], 4), "wall_seconds_incl_python": round(t_build, 2),
           "lm_iterations_per_s": tm["total_trials"] / max(1e-9, tm["gpu_seconds"] + tm["flatten_seconds"])}
    batch = sb.Batch(flats)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    batch.upload(sp)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    torch.cuda.profiler.start()   # cudaProfilerStart: `ncu --profile-from-start off` then skips the set-up launches
    time_batch(batch, stream, flush, 0, args.warmup, torch)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    dev_ms, per_step = time_batch(batch, stream, flush, args.steps, 0, torch)
    if world > 1:
        dist.barrier()
    batch.download(sp)
    stats = [batch.stats(i) for i in range(batch.n)]
    trials_per_step = sum(s["num_trials"] for s in stats)
    launches_per_step = int(batch.kernel_launches)

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
    # ---- the same call pattern with PERSISTENT problem handles: the batch keeps its symbolic structure on the device; a step ships only
    # the numeric inputs (edge poses, observations) from host buffers, runs, and reads the results back -- no symbolic work, no re-pack
    pers = None
    if extras:
        vals = [(f.array("edge_pose", 3 * f.p.n_edges, np.float64), f.array("obs_z", 3 * f.p.n_obs, np.float64)) for f in flats]
        k = max(2, min(args.steps, 5))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(k):
            flush.fill_(1)
            torch.cuda.synchronize()
            for i, (e_, z_) in enumerate(vals):
                batch.update_values(i, edge_pose=e_, obs_z=z_)
            batch.upload_values(sp)
            batch.run(sp)
            batch.download(sp)
            for i in range(n):
                L.srba_b200_batch_unpack(batch.h, i, rarr[i])
        pers_s = time.perf_counter() - t1
        pers = {"value": sum(b.r.stats.num_trials for b in bufs) * k / pers_s, "unit": UNIT, "ms_per_step": 1e3 * pers_s / k,
                "note": "srba_b200_batch_update_values + upload_values + run + download + unpack per step, host buffers in and out (rank 0)"}

    # ---- BASELINE config #4 read literally: 256 areas IN TOTAL, dealt round-robin to the N ranks (strong scaling)
    strong_ms, strong_trials = dev_ms / max(1, args.steps), float(trials_per_step)
    if extras and world > 1:
        mine = flats[rank::world]
        sb_ = sb.Batch(mine)
        sb_.upload(sp)
        ms_, _ = time_batch(sb_, stream, flush, max(3, args.steps // 2), 2, torch)
        strong_ms = ms_ / max(3, args.steps // 2)
        sb_.download(sp)
        strong_trials = float(sum(sb_.stats(i)["num_trials"] for i in range(sb_.n)))
        del sb_

    from srba_b200 import parallel as PL   # time = MAX over ranks, work = SUM over ranks (the only collectives: no data-path exchange)
    (dev_ms_max, e2e_s_max, strong_ms_max), (trials_all, e2e_trials_all, strong_trials_all) = PL.reduce_max_sum(
        [dev_ms, e2e_s, strong_ms], [trials_per_step, e2e_trials, strong_trials], device="cuda")
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
    # ---- the dominant kernel of the step (profiles/r02_*_launches.csv): the cluster-team forward solver of the wave pipeline, one launch per
    # LM trial.  Its launches are bracketed with CUDA events inside the library (srba_b200_kernel_timing) during two extra, untimed steps.
    dom = None
    wc = []
    for f in flats:
        o8 = (C.c_int64 * 8)()
        K.lib().srba_b200_problem_wave_counts(f.problem, o8)
        wc.append(list(o8))
    if any(w[5] for w in wc):
        K.lib().srba_b200_kernel_timing(1)
        time_batch(batch, stream, flush, 2, 0, torch)
        tk = (C.c_double * 2)()
        K.lib().srba_b200_kernel_timing_get(tk)
        K.lib().srba_b200_kernel_timing(0)
        # unique-touch bytes of the team areas' forward pass per LM trial: gather (reduced block in, tile out: 2 x 72 B per tile), the tiles of the
        # rows below the tail read once as update sources (72 B), every destination entry read-modify-written once (144 B) + its 16-byte record
        # + 8 bytes of pointers; summed over the trials each team area really ran, divided by the number of launches (= lock-step trials)
        fwd_bytes = sum((2 * 72 * w[0] + 72 * w[6] + (144 + 16 + 8) * w[7]) * st["num_trials"] for w, st in zip(wc, stats) if w[5])
        n_launch = max(1.0, tk[1] / 2.0)
        dom = {"kernel": "wave_solve_fwd_kernel<3,0,true> (8-CTA cluster per wide area: gather + level-scheduled block-sparse L D L^t below the dense tail)",
               "areas": int(sum(1 for w in wc if w[5])), "launches_per_step": n_launch, "avg_launch_ms": tk[0] / max(1.0, tk[1]),
               "share_of_step": (tk[0] / 2.0) / (dev_ms / args.steps),
               "algorithmic_bytes_per_launch": fwd_bytes / n_launch, "achieved": fwd_bytes / n_launch / (tk[0] / max(1.0, tk[1]) * 1e-3) / 1e9}
    ncu = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            ncu = json.load(open(tp))
        except Exception:
            ncu = {}
    config = {"workload": workload_name(args.areas, args.kfs) if not args.chain_like else
              "G4 (chain-like case): the last %d keyframes of a %d-keyframe tour" % (args.areas, args.kfs),
              "areas_per_gpu": args.areas, "tour_keyframes": args.kfs, "tour_seed": args.tour_seed, "lm_trials_per_step_per_gpu": trials_per_step,
              "trials_per_area": {"mean": trials_per_step / max(1, n), "max": max(s["num_trials"] for s in stats)},
              "area_sizes": size_stats(flats),
              "ranks": "same tour on every rank, independent observation noise per rank",
              "sequential_run_of_the_tour": seq,
              "l2": "flushed between timed iterations (256 MiB fill)", "timing": "CUDA events around each step on the launching stream",
              "kernel_launches_per_step": launches_per_step, "areas_per_s": n * world * args.steps / (dev_ms_max * 1e-3)}
    if pers:
        config["e2e_persistent_handles"] = pers
    if extras:
        config["strong_scaling_256_areas_total"] = {"areas_total": n, "ms_per_step": strong_ms_max, "value": strong_trials_all / (strong_ms_max * 1e-3), "unit": UNIT,
                                                    "note": "BASELINE config #4 read literally: the SAME %d areas dealt to the %d rank(s)" % (n, world)}
    if extras and world == 1 and not args.chain_like:
        sw = {}
        for a in sweep:
            bb = sb.Batch(sets[a])
            bb.upload(sp)
            k = max(2, min(args.steps, 5))
            ms_, _ = time_batch(bb, stream, flush, k, 1, torch)
            bb.download(sp)
            tr_ = sum(bb.stats(i)["num_trials"] for i in range(bb.n))
            sw[str(a)] = {"ms_per_step": ms_ / k, "value": tr_ * k / (ms_ * 1e-3), "areas_per_s": a * k / (ms_ * 1e-3), "lm_trials_per_step": tr_,
                          "distinct_areas": len({id(f) for f in sets[a]})}
            del bb
        config["areas_sweep"] = sw
        # the round-1 workload, for continuity: chain-like areas only
        _, _, cs = build_problems(seed=1, areas=256, n_kfs=700, last=True)
        bb = sb.Batch(cs[256])
        bb.upload(sp)
        ms_, _ = time_batch(bb, stream, flush, args.steps, 2, torch)
        bb.download(sp)
        tr_ = sum(bb.stats(i)["num_trials"] for i in range(bb.n))
        config["chain_like_case"] = {"workload": "last 256 keyframes of a 700-keyframe tour (round-1 bench: chain-like areas)", "ms_per_step": ms_ / args.steps,
                                     "value": tr_ * args.steps / (ms_ * 1e-3), "lm_trials_per_step": tr_, "area_sizes": size_stats(cs[256])}
        del bb
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(batch.h2d_bytes), "d2h_bytes_per_step": int(batch.d2h_bytes)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"] if dom else "srba_lm_batch_kernel (all instantiations of a step)",
                         "achieved": dom["achieved"] if dom else achieved, "peak": peak, "unit": "GB/s", "frac": (dom["achieved"] if dom else achieved) / peak,
                         "traffic": ncu.get("dominant_kernel_dram_bytes_per_launch"), "peak_source": peak_src,
                         "dominant_kernel": dom,
                         "whole_step": {"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes": ab["total_bytes"],
                                        "note": "all kernels of a step: total algorithmic bytes / step time"},
                         "note": "latency-bound, not bandwidth-bound: the level-scheduled solver is a chain of 80-240 cluster-barrier-separated phases per factorisation "
                                 "(~4 us each: two dependent L2 round trips + barrier); see DESIGN.md 4",
                         "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"] if dom else ab["total_bytes"],
                         "stage_bytes": {k: ab[k] for k in ("chain", "assem_res", "assem_jh", "jacmat", "schur", "solve", "update")},
                         "accounting": "touched bytes only (refreshed pose pairs, tiles of the filled block-sparse pattern, no Schur stage without landmarks); "
                                       "the round-1 figure also charged all pairs per trial and a dense n^2 Cholesky: dense_equiv_chol_bytes",
                         "dense_equiv_chol_bytes": ab["dense_equiv_chol"], "dense_equiv_chol_flop": ab["chol_flop"], "solve_flop_per_launch": ab["solve_flop"],
                         "issue_slot_util_pct": ncu.get("issue_slot_util_pct"), "fp64_pipe_util_pct": ncu.get("fp64_pipe_util_pct"),
                         "ncu_source": ncu.get("source")}}
    if not args.no_cpu_baseline and world == 1:
        t = time.perf_counter()
        cores = max(1, len(os.sched_getaffinity(0)))
        if fixture_matches(args):
            res = cpu_sample(args.areas, cores, 2)
            sec = sum(s_ for w in res for (s_, _, _) in w)
            tr = sum(t_ for w in res for (_, t_, _) in w)
            n_p = [q for w in res for (_, _, q) in w]
            txt = ("oracle (g++ -O2), single-thread rate: %d areas spaced evenly over the %d-area batch (n_p %d..%d, median %d; inputs: tests/golden/g4_tour3000_states.npz), "
                   "each solved once by one of %d concurrent single-threaded engines; 'opt' span only; value = LM trials / summed engine seconds"
                   % (len(n_p), args.areas, min(n_p), max(n_p), int(np.median(n_p)), cores))
            line["cpu_baseline"] = {"value": tr / sec, "unit": UNIT, "cores": 1, "kind": "port", "sample": txt + " (%.1f s wall)" % (time.perf_counter() - t)}
        elif args.kfs <= 1000:
            from srba_b200 import datasets as D
            sampled = (np.arange(args.kfs - args.areas, args.kfs) if args.chain_like else D.sample_keyframes(args.kfs, args.areas)).tolist()
            sec, tr = cpu_pass((args.tour_seed, args.kfs, sampled))
            txt = "oracle (g++ -O2, 1 thread): 'opt' span of the %d sampled local areas during a full sequential run of the %d-keyframe tour" % (len(sampled), args.kfs)
            line["cpu_baseline"] = {"value": tr / sec, "unit": UNIT, "cores": 1, "kind": "port", "sample": txt + " (%.1f s wall)" % (time.perf_counter() - t)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "no oracle fixture for this tour (tests/golden/make_g4_states.py)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
