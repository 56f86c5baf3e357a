"""Generates tests/golden/g4_tour3000_states.npz: the numeric inputs AND the oracle's outputs of the 256 local-area
problems bench.py times (BASELINE config #4 on the G1 tour: 3 000 keyframes, seed 1, areas sampled over the whole tour).

Two stages (a sequential ORACLE run of the whole tour is not feasible: the loop-closure areas of keyframes 2 200-3 000 have
400-640 unknown edges and the oracle needs 10-130 s per keyframe there, hours in total):

  --stage inputs   (needs the GPU: `gpurun -- python tests/golden/make_g4_states.py --stage inputs`)
        The PRODUCT engine runs the sequential SLAM (define_new_keyframe with optimisation at every keyframe, ~1 minute).  At every
        sampled keyframe the optimize_local_area problem of that keyframe -- the problem of the batch -- is recorded: the k2k edges it
        touches and their pre-optimisation poses.  Output: gpurun_out/g4_inputs.npz.
  --stage oracle   (CPU only, parallel: `python tests/golden/make_g4_states.py --stage oracle --procs 8`)
        For every area the ORACLE rebuilds the graph of the tour (no optimisation) up to that keyframe, sets the touched edges to the
        stored input poses and runs optimize_local_area(k, 3) -- the same numeric inputs as the batch problem.  Stored per area: the
        oracle's answer (final poses of the unknown edges, trial count, errors, per-trial trace (lambda, ok, rho, new error, accepted,
        relinearised), first increment, 'opt' seconds).

Any state is a legitimate INPUT for a parity test; the oracle's OUTPUTS never depend on the product.
Consumers: tests/test_gpu_parity.py::test_bench_batch_vs_oracle_fixture (the CUDA path on the same inputs),
bench.py (cpu_baseline / --impl reference: the oracle re-solves a bounded sample of these areas from the stored inputs).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

KW = dict(max_tree_depth=3, max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1, max_error_per_obs_to_stop=1e-8, pose_kind=0, obs_kind=0, noise_kind=1)


def stage_inputs(a):
    import srba_b200 as sb
    from srba_b200 import datasets as D
    ds = D.graphslam_se2_tour(a.kfs, seed=a.seed)
    eng = sb.RbaEngine(sb.default_engine_params(**dict(KW, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))))
    sampled = set(int(k) for k in D.sample_keyframes(a.kfs, a.areas))
    kf, edge_ptr, edge_id, edge_in = [], [0], [], []
    t0 = time.time()
    for k, f in enumerate(ds["frames"]):
        if k in sampled:
            eng.record_problems(True)
        D.feed(eng, f)
        if k in sampled:
            rec = eng.take_recorded()
            eng.record_problems(False)
            assert rec, k
            flat = rec[-1]   # the last optimize_edges call of define_new_keyframe is the local area (define_new_keyframe.h:97)
            te = flat.ids()["edge_ids_touched"].astype(np.int64)
            kf.append(k); edge_id.append(te); edge_in.append(flat.array("edge_pose", 3 * flat.p.n_edges, np.float64).reshape(-1, 3))
            edge_ptr.append(edge_ptr[-1] + len(te))
    np.savez_compressed(a.inputs, kfs=a.kfs, seed=a.seed, kf=np.array(kf, np.int64), edge_ptr=np.array(edge_ptr, np.int64),
                        edge_id=np.concatenate(edge_id), edge_in=np.concatenate(edge_in))
    print("wrote %s: %d areas, %d touched edges, %.0f s" % (a.inputs, len(kf), edge_ptr[-1], time.time() - t0))


def oracle_worker(args):
    inputs, mine = args
    import srba_b200 as sb
    from oracle import pyoracle as po
    from srba_b200 import datasets as D
    fx = np.load(inputs)
    ds = D.graphslam_se2_tour(int(fx["kfs"]), seed=int(fx["seed"]))
    kw = dict(KW, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))
    eo = po.OracleEngine(po.default_params(**kw))
    ep = sb.RbaEngine(sb.default_engine_params(**kw))   # host side only: graph + flattening (ids), never optimises here
    eo.trace_enable(True)
    at = {int(fx["kf"][i]): i for i in mine}
    out = {}
    for k, f in enumerate(ds["frames"]):
        if k > max(at):
            break
        D.feed(eo, f, run_local_optimization=False)
        D.feed(ep, f, run_local_optimization=False)
        if k not in at:
            continue
        i = at[k]
        ids_e = fx["edge_id"][fx["edge_ptr"][i]:fx["edge_ptr"][i + 1]]
        for e, q in zip(ids_e, fx["edge_in"][fx["edge_ptr"][i]:fx["edge_ptr"][i + 1]]):
            eo.set_edge_pose(int(e), q)
            ep.set_edge_pose(int(e), q)
        flat = ep.flatten_local_area(k, 3)
        ids = flat.ids()
        assert ids["edge_ids_touched"].astype(np.int64).tolist() == ids_e.tolist(), k
        eo.trace_clear()
        s0, _ = eo.opt_timer()
        res = eo.optimize_local_area(k, 3)
        s1, _ = eo.opt_timer()
        _, poses_out = eo.edges()
        assert res["num_kf2kf_edges_optimized"] == flat.p.n_unk_edges and res["num_observations"] == flat.p.n_total_obs_rows, (k, res, flat.sizes())
        tr = eo.traces()[-1] if res["num_lm_trials"] > 0 else dict(trials=np.zeros((0, 6)), deltas=np.zeros((0, 3 * flat.p.n_unk_edges)))
        ue = ids["k2k_ids"].astype(np.int64)
        out[i] = dict(unk_id=ue, unk_out=poses_out[ue].copy(), trials=np.asarray(tr["trials"], float).reshape(-1, 6),
                      delta0=np.asarray(tr["deltas"][0], float) if len(tr["deltas"]) else np.zeros(0),
                      num_trials=res["num_lm_trials"], num_iters=res["num_iters"], err_init=res["total_sqr_error_init"], err_final=res["total_sqr_error_final"],
                      opt_seconds=s1 - s0, n_obs=res["num_observations"], n_p=flat.p.n_unk_edges)
        # the oracle's own result must not leak into the next area of this worker: the inputs of every area are set explicitly above,
        # and edges an area does not touch are never read by it
        print("kf %5d  n_p %3d  n_obs %5d  trials %3d  opt %.3f s" % (k, flat.p.n_unk_edges, res["num_observations"], res["num_lm_trials"], s1 - s0), flush=True)
    return out


def stage_oracle(a):
    import multiprocessing as mp
    fx = np.load(a.inputs)
    n = len(fx["kf"])
    # longest areas first inside each worker does not matter (a worker walks the tour once, in keyframe order); deal round-robin
    shares = [list(range(w, n, a.procs)) for w in range(a.procs)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(a.procs) as pool:
        parts = pool.map(oracle_worker, [(a.inputs, s) for s in shares if s])
    res = {}
    for p in parts:
        res.update(p)
    out = dict(kfs=int(fx["kfs"]), seed=int(fx["seed"]), kf=fx["kf"], edge_ptr=fx["edge_ptr"], edge_id=fx["edge_id"], edge_in=fx["edge_in"])
    for key, dt in (("num_trials", np.int64), ("num_iters", np.int64), ("err_init", np.float64), ("err_final", np.float64), ("opt_seconds", np.float64),
                    ("n_obs", np.float64), ("n_p", np.int64)):
        out[key] = np.array([res[i][key] for i in range(n)], dt)
    for key, ptr in (("unk_id", "unk_ptr"), ("trials", "trials_ptr"), ("delta0", "delta0_ptr")):
        out[ptr] = np.concatenate([[0], np.cumsum([len(res[i][key]) for i in range(n)])]).astype(np.int64)
        out[key] = np.concatenate([res[i][key] for i in range(n)])
    out["unk_out"] = np.concatenate([res[i]["unk_out"] for i in range(n)])
    np.savez_compressed(a.out, **out)
    print("wrote %s: %d areas, %d LM trials, %.1f s of oracle 'opt' time on the sampled areas, %.0f s wall on %d processes"
          % (a.out, n, int(out["num_trials"].sum()), out["opt_seconds"].sum(), time.time() - t0, a.procs))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", choices=["inputs", "oracle"], required=True)
    ap.add_argument("--kfs", type=int, default=3000)
    ap.add_argument("--areas", type=int, default=256)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--procs", type=int, default=max(1, len(os.sched_getaffinity(0))))
    ap.add_argument("--inputs", default=os.path.join(ROOT, "gpurun_out", "g4_inputs.npz"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "g4_tour3000_states.npz"))
    a = ap.parse_args()
    stage_inputs(a) if a.stage == "inputs" else stage_oracle(a)
