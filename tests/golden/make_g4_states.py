"""Generates tests/golden/g4_tour3000_states.npz: the numeric inputs AND the oracle's outputs of the 256 local-area
problems bench.py times (BASELINE config #4 on the G1 tour: 3 000 keyframes, seed 1, areas sampled over the whole tour).

The ORACLE runs the whole sequential SLAM (define_new_keyframe with optimisation at every keyframe -- ~20-30 minutes on
one core, which is why this is a committed fixture and not part of the GPU-box tests / bench legs).  At every sampled
keyframe k the call is split the way define_new_keyframe does it (impl/define_new_keyframe.h:60-97): graph update +
stage-1 optimisation of the new edges without initial value, THEN a snapshot of the edge poses, then
optimize_local_area(k, 3) -- the problem of the batch.  Stored per area: the k2k edges its flattened problem touches
(ids from the product's host-side flattening, no GPU needed) with their input poses, and the oracle's answer: final
poses of the unknown edges, trial count, errors, per-trial trace (lambda, ok, rho, new error, accepted, relinearised),
first increment, 'opt' seconds.

Consumers: tests/test_gpu_parity.py::test_bench_batch_vs_oracle_fixture (the CUDA path on the same inputs),
bench.py (cpu_baseline / --impl reference: the oracle re-solves a bounded sample of these areas from the stored
inputs instead of re-running the 20-minute tour).
Usage: python tests/golden/make_g4_states.py [--kfs 3000 --areas 256 --seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import srba_b200 as sb
from oracle import pyoracle as po
from srba_b200 import datasets as D

ap = argparse.ArgumentParser()
ap.add_argument("--kfs", type=int, default=3000)
ap.add_argument("--areas", type=int, default=256)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "g4_tour3000_states.npz"))
a = ap.parse_args()

ds = D.graphslam_se2_tour(a.kfs, seed=a.seed)
kw = dict(max_tree_depth=3, max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1, max_error_per_obs_to_stop=1e-8, pose_kind=0, obs_kind=0,
          noise_kind=1, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))
eo = po.OracleEngine(po.default_params(**kw))
ep = sb.RbaEngine(sb.default_engine_params(**kw))   # host side only: graph + flattening (ids), never optimises here
eo.trace_enable(True)
sampled = set(int(k) for k in D.sample_keyframes(a.kfs, a.areas))
rec = dict(kf=[], edge_ptr=[0], edge_id=[], edge_in=[], unk_ptr=[0], unk_id=[], unk_out=[], trials_ptr=[0], trials=[], delta0_ptr=[0], delta0=[],
           num_trials=[], num_iters=[], err_init=[], err_final=[], opt_seconds=[], n_obs=[])
t_start = time.time()
for k, f in enumerate(ds["frames"]):
    D.feed(ep, f, run_local_optimization=False)
    if k not in sampled:
        eo.trace_clear()
        D.feed(eo, f)
        continue
    r = D.feed(eo, f, run_local_optimization=False)
    for e in r["created_edges"]:                      # optimize_new_edges_alone (define_new_keyframe.h:66-88)
        if not e["has_approx_init_val"]:
            eo.optimize_edges([e["id"]])
    _, poses_in = eo.edges()
    flat = ep.flatten_local_area(k, 3)
    ids = flat.ids()
    eo.trace_clear()
    s0, _ = eo.opt_timer()
    res = eo.optimize_local_area(k, 3)
    s1, _ = eo.opt_timer()
    _, poses_out = eo.edges()
    assert res["num_kf2kf_edges_optimized"] == flat.p.n_unk_edges and res["num_observations"] == flat.p.n_total_obs_rows, (k, res, flat.sizes())
    tr = eo.traces()[-1] if res["num_lm_trials"] > 0 else dict(trials=np.zeros((0, 6)), deltas=np.zeros((0, 3 * flat.p.n_unk_edges)))
    te, ue = ids["edge_ids_touched"].astype(np.int64), ids["k2k_ids"].astype(np.int64)
    rec["kf"].append(k)
    rec["edge_id"].append(te); rec["edge_in"].append(poses_in[te]); rec["edge_ptr"].append(rec["edge_ptr"][-1] + len(te))
    rec["unk_id"].append(ue); rec["unk_out"].append(poses_out[ue]); rec["unk_ptr"].append(rec["unk_ptr"][-1] + len(ue))
    rec["trials"].append(np.asarray(tr["trials"], float).reshape(-1, 6)); rec["trials_ptr"].append(rec["trials_ptr"][-1] + len(tr["trials"]))
    d0 = np.asarray(tr["deltas"][0], float) if len(tr["deltas"]) else np.zeros(0)
    rec["delta0"].append(d0); rec["delta0_ptr"].append(rec["delta0_ptr"][-1] + len(d0))
    for key in ("num_trials", "num_iters", "err_init", "err_final"):
        rec[key].append(res[{"num_trials": "num_lm_trials", "num_iters": "num_iters", "err_init": "total_sqr_error_init", "err_final": "total_sqr_error_final"}[key]])
    rec["opt_seconds"].append(s1 - s0); rec["n_obs"].append(res["num_observations"])
    print("kf %5d  n_p %3d  n_obs %4d  trials %3d  opt %.3f s   [%.0f s elapsed]" % (k, flat.p.n_unk_edges, res["num_observations"], res["num_lm_trials"], s1 - s0, time.time() - t_start), flush=True)

out = dict(kfs=a.kfs, seed=a.seed, kf=np.array(rec["kf"], np.int64))
for key in ("edge_ptr", "unk_ptr", "trials_ptr", "delta0_ptr", "num_trials", "num_iters"):
    out[key] = np.array(rec[key], np.int64)
for key in ("err_init", "err_final", "opt_seconds", "n_obs"):
    out[key] = np.array(rec[key], np.float64)
out["edge_id"] = np.concatenate(rec["edge_id"]); out["edge_in"] = np.concatenate(rec["edge_in"])
out["unk_id"] = np.concatenate(rec["unk_id"]); out["unk_out"] = np.concatenate(rec["unk_out"])
out["trials"] = np.concatenate(rec["trials"]); out["delta0"] = np.concatenate(rec["delta0"])
np.savez_compressed(a.out, **out)
print("wrote %s: %d areas, %d LM trials, %.1f s of oracle 'opt' time on the sampled areas, %.0f s in total" % (a.out, len(rec["kf"]), int(out["num_trials"].sum()), out["opt_seconds"].sum(), time.time() - t_start))
