"""CPU-only checks of the product's host engine (srba_b200/csrc/host_engine.hpp) against the oracle: graph
construction, edge-creation policy, spanning trees, bootstrap values, unknown selection and the FLATTENED problem
(index arrays must be bit-exact).  No GPU work: the product engine runs with
run_local_optimization=False and the oracle's numeric state is injected after each keyframe."""
import math

import numpy as np
import pytest

import srba_b200 as sb
from fixtures import FX, inject_state, se2_noise_lambda
from oracle import pyoracle as po
from srba_b200 import datasets as D

S2 = dict(pose_kind=0, obs_kind=0, noise_kind=1, max_tree_depth=3, max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1,
          max_error_per_obs_to_stop=1e-8)


def make_pair(kw):
    return po.OracleEngine(po.default_params(**kw)), sb.RbaEngine(sb.default_engine_params(**kw))


def flat_pose_chain(flat):
    """numpy re-evaluation of the flattened numeric spanning tree (SE2 only): slot poses from edge poses + paths."""
    p = flat.p
    edge = flat.array("edge_pose", p.n_edges * 3, np.float64).reshape(-1, 3)
    ptr, pe, pd = flat.array("pair_path_ptr", p.n_pairs + 1, np.int32), flat.array("path_edge", p.n_path, np.int32), flat.array("path_dir", p.n_path, np.uint8)
    slots = np.zeros((2 * p.n_pairs, 3))
    for t in range(p.n_pairs):
        acc = np.zeros(3)
        for k in range(ptr[t], ptr[t + 1]):
            acc = D.se2_compose(acc, edge[pe[k]] if pd[k] else D.se2_inverse(edge[pe[k]]))
        slots[2 * t], slots[2 * t + 1] = acc, D.se2_inverse(acc)
    return slots


def flat_total_error(flat, slots):
    p = flat.p
    z = flat.array("obs_z", p.n_obs * 3, np.float64).reshape(-1, 3)
    os_, mult = flat.array("obs_slot", p.n_obs, np.int32), flat.array("obs_mult", p.n_obs, np.int32)
    tot = 0.0
    for i in range(p.n_obs):
        P = slots[os_[i]] if os_[i] >= 0 else np.zeros(3)
        c, s = math.cos(P[2]), math.sin(P[2])
        r = np.array([(z[i, 0] - P[0]) * c + (z[i, 1] - P[1]) * s, -(z[i, 0] - P[0]) * s + (z[i, 1] - P[1]) * c, D.wrap(z[i, 2] - P[2])])
        tot += mult[i] * float(r @ r)
    return tot


def compare_graph(eo, ep, kf):
    to, no, do = eo.next_edge(kf)
    tp, np_, dp = ep.next_edge(kf)
    assert to.tolist() == tp.tolist() and no.tolist() == np_.tolist() and do.tolist() == dp.tolist()
    for t in to:
        a, b = max(kf, int(t)), min(kf, int(t))
        assert eo.all_edges_path(a, b).tolist() == ep.all_edges_path(a, b).tolist()


def run_lockstep(kw, frames, check_every=1):
    eo, ep = make_pair(kw)
    eo.trace_enable(True)
    depth = kw["max_optimize_depth"]
    for k, f in enumerate(frames):
        ro = D.feed(eo, f, run_local_optimization=False)
        rp = D.feed(ep, f, run_local_optimization=False)
        assert ro["kf_id"] == rp["kf_id"] == k
        assert ro["created_edges"] == rp["created_edges"]
        for c in ro["created_edges"]:  # bootstrap values of the new edges
            fo, to_, po_ = eo.get_edge(c["id"])
            fp, tp_, pp_ = ep.get_edge(c["id"])
            assert (fo, to_) == (fp, tp_)
            np.testing.assert_allclose(po_, pp_, rtol=0, atol=1e-12)
        if k % check_every == 0:
            compare_graph(eo, ep, k)
        # stage 1 (define_new_keyframe.h:59-91) then the local area, on the oracle; flatten on the product BEFORE each
        calls = [("edges", [c["id"]]) for c in ro["created_edges"] if not c["has_approx_init_val"]] + [("area", None)]
        for kind, ids in calls:
            if kind == "edges":
                flat = ep.flatten(ids)
            else:
                sel_k2k, sel_lm = ep.select_local_area(k, depth)
                flat = ep.flatten(sel_k2k, sel_lm) if (len(sel_k2k) or len(sel_lm)) else None
            eo.trace_clear()
            info = eo.optimize_edges(ids) if kind == "edges" else eo.optimize_local_area(k, depth)
            tr = eo.traces()
            if flat is None:
                assert len(tr) == 0 or info["num_total_scalar_optimized"] == 0
                continue
            assert len(tr) == 1
            t = tr[0]
            fid = flat.ids()
            assert fid["k2k_ids"].tolist() == t["k2k_ids"].tolist()       # unknown edges, BFS visit order
            assert fid["lm_ids"].tolist() == t["lm_ids"].tolist()         # unknown landmarks, threshold-crossing order
            # observation rows incl. the duplicates of optimize_edges.h:177-193
            mult = flat.array("obs_mult", flat.p.n_obs, np.int32)
            assert flat.p.n_total_obs_rows == len(t["obs_idx"]) == info["num_observations"] == int(mult.sum())
            uniq, cnt = np.unique(t["obs_idx"], return_counts=True)
            mine = dict(zip(fid["obs_global"].tolist(), mult.tolist()))
            assert mine == dict(zip(uniq.tolist(), cnt.tolist()))
            first_touch = list(dict.fromkeys(t["obs_idx"].tolist()))
            assert fid["obs_global"].tolist() == first_touch              # residual order = first-touch order
            assert flat.p.n_blkA + (0 if kw["obs_kind"] == 0 else flat.p.n_blkF) == info["num_jacobians"]
            if kw["pose_kind"] == 0:
                # numeric spanning tree + residuals of the flattened problem reproduce the oracle's initial error
                # (the oracle evaluated it on the same pre-optimisation state)
                slots = flat_pose_chain(flat)
                tot = flat_total_error(flat, slots)
                assert abs(tot - info["total_sqr_error_init"]) <= 1e-9 * max(1.0, info["total_sqr_error_init"])
                npairs = sum(len(eo.next_edge(int(fr))[0][eo.next_edge(int(fr))[0] < fr]) for fr in np.unique(fid["pairs"][:, 0]))
                assert info["num_span_tree_numeric_updates"] >= flat.p.n_pairs > 0 and npairs >= 0
        inject_state(eo, ep)  # the oracle's numeric state becomes the product engine's
    return eo, ep


def test_lockstep_se2_graph_slam_tour():
    ds = D.graphslam_se2_tour(260, seed=3)
    kw = dict(S2, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))
    run_lockstep(kw, ds["frames"])


def test_lockstep_se2_reference_fixtures():
    for key, cols in (("submaps_edge_init_values", (2, 3, 4)), ("tutorial_relgraphslam_se2", (2, 3, 5))):
        rows = FX[key]["dataset"] if key == "submaps_edge_init_values" else FX[key]["dataset"]["rows"]
        nkf = int(max(r[0] for r in rows)) + 1
        frames = []
        for kf in range(nkf):
            ids, obs, fixed = [kf], [[0, 0, 0]], [1]
            for r in rows:
                if int(r[0]) == kf:
                    ids.append(int(r[1])); obs.append([r[cols[0]], r[cols[1]], r[cols[2]]]); fixed.append(0)
            frames.append(dict(feat_id=np.array(ids, np.uint64), obs=np.array(obs, float), is_fixed=np.array(fixed, np.uint8),
                               is_unknown_with_init_val=np.zeros(len(ids), np.uint8), feat_rel_pos=np.zeros((len(ids), 3))))
        kw = dict(S2, submap_size=5, noise_lambda=se2_noise_lambda(0.001, math.radians(0.05)))
        run_lockstep(kw, frames)


@pytest.mark.parametrize("obs_kind,name", [(1, "cartesian3d"), (3, "stereo"), (2, "mono")])
def test_lockstep_se3_landmark_slam(obs_kind, name):
    ds = D.landmark_slam_se3(n_kfs=60, n_lms=400, seed=2, obs_kind=name, lm_init_sigma=0.05 if name == "mono" else None)
    kw = dict(pose_kind=1, obs_kind=obs_kind, noise_kind=0, noise_std=0.5 if obs_kind != 1 else 1e-3, max_tree_depth=3, max_optimize_depth=3,
              submap_size=10, min_obs_to_loop_closure=4)
    if obs_kind != 1:
        kw.update(sensor_pose_kind=1, sensor_pose=ds["sensor_pose_xyzypr"], cam_left=ds["cam"], cam_right=ds["cam"],
                  right_cam_pose=[ds["baseline"], 0, 0, 1, 0, 0, 0])
    run_lockstep(kw, ds["frames"], check_every=5)


def test_symbols_and_no_device_error():
    """The C-ABI library loads, exports every declared symbol, and fails loudly (no CPU fallback) without a GPU."""
    from srba_b200 import _capi as K
    L = K.lib()
    for s in K.EXPORTED_SYMBOLS:
        assert hasattr(L, s), s
    # every function the two public headers declare is exported (and bound in _capi.EXPORTED_SYMBOLS), nothing else is listed
    import os, re
    declared = set()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for h in ("include/srba_b200.h", "include/srba_b200_engine.h"):
        txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, h)).read(), flags=re.S)
        declared |= set(re.findall(r"\b(srba_(?:b200|engine|flat)_[a-z0-9_]+)\s*\(", txt))
    assert len(declared) >= 60
    assert declared == set(K.EXPORTED_SYMBOLS), (sorted(declared - set(K.EXPORTED_SYMBOLS)), sorted(set(K.EXPORTED_SYMBOLS) - declared))
    for s in declared:
        assert hasattr(L, s), s
    if L.srba_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ds = D.graphslam_se2_tour(12, seed=1)
    ep = sb.RbaEngine(sb.default_engine_params(**dict(S2, noise_lambda=D.se2_noise_lambda(1e-3, 1e-3))))
    D.feed(ep, ds["frames"][0])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        D.feed(ep, ds["frames"][1])


def test_solver_kind_reaches_the_flattened_problem_and_is_validated():
    """RBA_OPTIONS::solver_t (srba_options_solver.h:21-77) is a run-time tag here: it travels engine params -> srba_model_t,
    the packer sizes the pristine-gradient array only for the no-Schur variant, and an unknown tag is rejected."""
    import ctypes as C
    from srba_b200 import _capi as K
    ds = D.landmark_slam_se3(n_kfs=12, n_lms=100, seed=2, obs_kind="cartesian3d")
    plans = {}
    for kind in (0, 1, 2, 7):
        ep = sb.RbaEngine(sb.default_engine_params(pose_kind=1, obs_kind=1, noise_kind=0, noise_std=1e-3, max_tree_depth=3, max_optimize_depth=3,
                                                   submap_size=10, min_obs_to_loop_closure=4, solver_kind=kind))
        for f in ds["frames"]:
            D.feed(ep, f, run_local_optimization=False)
        flat = ep.flatten_local_area(11, 3)
        assert flat.p.model.solver_kind == kind
        out = (C.c_int64 * 8)()
        rc = K.lib().srba_b200_problem_plan(flat.problem, out)
        if kind == 7:
            assert rc == -1  # SRBA_ERR_BAD_ARGUMENT
            continue
        assert rc == 0
        plans[kind] = (out[1], flat.p.n_unk_edges * 6 + flat.p.n_unk_lms * 3)
    assert plans[0][0] == plans[1][0]
    extra = plans[2][0] - plans[0][0]
    assert 8 * plans[2][1] <= extra <= 8 * plans[2][1] + 64  # one more gradient-sized array (16-byte aligned)


def test_complete_spanning_tree_bfs_reaches_the_same_keyframes_as_the_oracle():
    """Integer half of create_complete_spanning_tree (impl/spantree_create_complete.h:35-83) on the host engine: same number of
    keyframes reached as the oracle's restatement for several roots and depth limits (no device needed for the size query);
    eval_overall_squared_error needs the device and says so."""
    ds = D.graphslam_se2_tour(150, seed=2)
    kw = dict(S2, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))
    eo = po.OracleEngine(po.default_params(**kw))
    ep = sb.RbaEngine(sb.default_engine_params(**kw))
    for f in ds["frames"]:
        D.feed(eo, f, run_local_optimization=False)
        D.feed(ep, f, run_local_optimization=False)
    for root in (0, 37, 149):
        for depth in (0, 1, 3, 7, 2 ** 64 - 1):
            ids, _ = eo.complete_spanning_tree(root, depth)
            n = ep.L.srba_engine_complete_spanning_tree(ep.h, root, depth, None, None, 0)
            assert n == len(ids), (root, depth)
    from srba_b200 import _capi as K
    if K.lib().srba_b200_device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            ep.eval_overall_squared_error()


def test_eval_overall_c_abi_validates_its_arguments():
    """srba_b200_eval_overall_squared_error (include/srba_b200.h): malformed inputs are rejected on the host before any device work
    (error codes instead of the reference's ASSERT_ exceptions, impl/eval_overall_error.h:106-110); a well-formed call without a
    device reports SRBA_ERR_NO_DEVICE (no CPU fallback)."""
    import ctypes as C
    from srba_b200 import _capi as K
    L = K.lib()

    class EvalProblem(C.Structure):
        _fields_ = [("model", K.Model), ("n_edges", C.c_int32), ("n_pairs", C.c_int32), ("n_path", C.c_int32), ("n_obs", C.c_int32), ("n_lms", C.c_int32),
                    ("reserved_", C.c_int32), ("edge_pose", C.POINTER(C.c_double)), ("pair_path_ptr", C.POINTER(C.c_int32)), ("path_edge", C.POINTER(C.c_int32)),
                    ("path_dir", C.POINTER(C.c_uint8)), ("obs_z", C.POINTER(C.c_double)), ("obs_pair", C.POINTER(C.c_int32)), ("obs_inverse", C.POINTER(C.c_uint8)),
                    ("obs_lm", C.POINTER(C.c_int32)), ("lm_pos", C.POINTER(C.c_double))]

    L.srba_b200_eval_overall_squared_error.restype = C.c_int
    L.srba_b200_eval_overall_squared_error.argtypes = [C.POINTER(EvalProblem), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    edge = np.array([[1.0, 0.0, 0.1], [0.5, 0.2, -0.3]])
    ptr_ = np.array([0, 2], np.int32)
    pe, pd = np.array([0, 1], np.int32), np.array([1, 0], np.uint8)
    z = np.array([[1.4, 0.3, -0.2], [0.0, 0.0, 0.0]])
    op, oi = np.array([0, -1], np.int32), np.array([0, 0], np.uint8)

    def problem(**over):
        q = EvalProblem()
        q.model.pose_kind, q.model.obs_kind = K.POSE_SE2, K.OBS_RELPOSES2D
        q.n_edges, q.n_pairs, q.n_path, q.n_obs, q.n_lms = 2, 1, 2, 2, 0
        q.edge_pose = edge.ctypes.data_as(C.POINTER(C.c_double)); q.pair_path_ptr = ptr_.ctypes.data_as(C.POINTER(C.c_int32))
        q.path_edge = pe.ctypes.data_as(C.POINTER(C.c_int32)); q.path_dir = pd.ctypes.data_as(C.POINTER(C.c_uint8))
        q.obs_z = z.ctypes.data_as(C.POINTER(C.c_double)); q.obs_pair = op.ctypes.data_as(C.POINTER(C.c_int32)); q.obs_inverse = oi.ctypes.data_as(C.POINTER(C.c_uint8))
        for k, v in over.items():
            setattr(q, k, v)
        return q

    out = C.c_double()
    assert L.srba_b200_eval_overall_squared_error(None, C.byref(out), None) == K.ERR_BAD_ARGUMENT
    assert L.srba_b200_eval_overall_squared_error(C.byref(problem(n_path=-1)), C.byref(out), None) == K.ERR_BAD_ARGUMENT
    bad_edge = np.array([0, 7], np.int32)
    assert L.srba_b200_eval_overall_squared_error(C.byref(problem(path_edge=bad_edge.ctypes.data_as(C.POINTER(C.c_int32)))), C.byref(out), None) == K.ERR_BAD_ARGUMENT
    bad_pair = np.array([3, -1], np.int32)
    assert L.srba_b200_eval_overall_squared_error(C.byref(problem(obs_pair=bad_pair.ctypes.data_as(C.POINTER(C.c_int32)))), C.byref(out), None) == K.ERR_BAD_ARGUMENT
    q = problem()
    q.model.obs_kind = K.OBS_STEREO   # SE2 is built with RelativePoses_2D only
    assert L.srba_b200_eval_overall_squared_error(C.byref(q), C.byref(out), None) == K.ERR_UNSUPPORTED_MODEL
    rc = L.srba_b200_eval_overall_squared_error(C.byref(problem()), C.byref(out), None)
    if L.srba_b200_device_count() == 0:
        assert rc == K.ERR_NO_DEVICE and b"no CUDA device" in L.srba_b200_last_error()
    else:
        assert rc == 0 and out.value > 0


def test_observation_indices_to_optimize_flattening_and_oracle_semantics():
    """RbaEngine::optimize_local_area(..., observation_indices_to_optimize) (RbaEngine.h:235-241, impl/optimize_edges.h:218-234): the
    residual rows are the caller's list (repeats counted in nObs), the list must cover every observation of the unknown columns
    (the ASSERT_ of compute_minus_gradient.h:55), Jacobian blocks / Hessian terms are unchanged.  Host flattening vs the oracle."""
    ds = D.graphslam_se2_tour(60, seed=3)
    kw = dict(S2, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]))
    eo, ep = make_pair(kw)
    for f in ds["frames"]:
        D.feed(eo, f)
        D.feed(ep, f, run_local_optimization=False)
        inject_state(eo, ep)
    root = 59
    base = ep.flatten_local_area(root, 3)
    obs = base.ids()["obs_global"]
    lst = np.concatenate([obs[::-1], obs[:3]])            # every involved observation, reversed, three of them twice
    flat = ep.flatten_local_area(root, 3, observation_indices_to_optimize=lst)
    p, q = flat.p, base.p
    assert (p.n_unk_edges, p.n_blkA, p.n_HAp, p.n_HAp_terms, p.n_obs) == (q.n_unk_edges, q.n_blkA, q.n_HAp, q.n_HAp_terms, q.n_obs)
    assert p.n_total_obs_rows == len(lst)
    assert flat.ids()["obs_global"].tolist() == obs[::-1].tolist()      # residual order = the caller's (first touch)
    mult = dict(zip(flat.ids()["obs_global"].tolist(), flat.array("obs_mult", p.n_obs, np.int32).tolist()))
    assert all(mult[int(g)] == (2 if g in obs[:3] else 1) for g in obs)
    # total error over the listed rows == the oracle's initial error for the same list (state is identical: injected)
    io = eo.optimize_local_area(root, 3, observation_indices_to_optimize=lst)
    assert io["num_observations"] == len(lst)
    assert abs(flat_total_error(flat, flat_pose_chain(flat)) - io["total_sqr_error_init"]) <= 1e-9 * io["total_sqr_error_init"]
    # a list that misses an observation of an unknown column: both raise (the reference ASSERT_s)
    with pytest.raises(RuntimeError):
        ep.flatten_local_area(root, 3, observation_indices_to_optimize=obs[1:])
    with pytest.raises(po.OracleError):
        eo.optimize_local_area(root, 3, observation_indices_to_optimize=obs[1:])


def flat_arrays(flat, S, O):
    """Every array of a flattened problem (include/srba_b200.h: srba_problem_t) as numpy copies."""
    p = flat.p
    counts = dict(edge_pose=p.n_edges * S, unk_edge=p.n_unk_edges, lm_pos=p.n_lms * 3, unk_lm=p.n_unk_lms, pair_path_ptr=p.n_pairs + 1, path_edge=p.n_path, path_dir=p.n_path,
                  obs_z=p.n_obs * O, obs_slot=p.n_obs, obs_lm=p.n_obs, obs_mult=p.n_obs, colA_ptr=p.n_unk_edges + 1, blkA_obs=p.n_blkA, blkA_slotA=p.n_blkA, blkA_slotD=p.n_blkA,
                  blkA_normal_dir=p.n_blkA, colF_ptr=p.n_unk_lms + 1, blkF_obs=p.n_blkF, blkF_slot=p.n_blkF, row_ptr=p.n_obs + 1, row_blk=p.n_blkA + p.n_blkF,
                  HAp_row=p.n_HAp, HAp_col=p.n_HAp, HAp_term_ptr=p.n_HAp + 1, HAp_term_a=p.n_HAp_terms, HAp_term_b=p.n_HAp_terms, HApf_row=p.n_HApf, HApf_col=p.n_HApf,
                  HApf_term_ptr=p.n_HApf + 1, HApf_term_a=p.n_HApf_terms, HApf_term_f=p.n_HApf_terms, rowApf_ptr=p.n_unk_edges + 1, colApf_ptr=p.n_unk_lms + 1,
                  colApf_idx=p.n_HApf, S_row=p.n_S, S_col=p.n_S, S_HAp=p.n_S, S_term_ptr=p.n_S + 1, S_term_rowblk=p.n_S_terms, S_term_colblk=p.n_S_terms)
    out = {k: np.zeros(0) if n == 0 else flat.array(k, n, np.float64 if k in ("edge_pose", "lm_pos", "obs_z") else (np.uint8 if k in ("path_dir", "blkA_normal_dir") else np.int32)) for k, n in counts.items()}
    out.update(flat.ids())
    out["sizes"] = np.array(list(flat.sizes().values()))
    return out


def test_flatten_local_areas_in_parallel_equals_one_by_one():
    """srba_engine_flatten_local_areas (snapshot semantics): the local areas of many keyframes of one map state, flattened by
    several host threads that only READ the engine, are the problems flatten_local_area gives one at a time -- every array, bit for bit.
    A root without unknowns (or unknown to the map) gives None."""
    ds = D.landmark_slam_se3(n_kfs=120, n_lms=1200, seed=4, obs_kind="cartesian3d")
    ep = sb.RbaEngine(sb.default_engine_params(pose_kind=1, obs_kind=1, noise_kind=0, noise_std=1e-3, max_tree_depth=4, max_optimize_depth=4, submap_size=10, min_obs_to_loop_closure=4))
    for f in ds["frames"]:
        D.feed(ep, f, run_local_optimization=False)
    roots = list(range(0, 120, 3))
    one = [ep.flatten_local_area(r, 4) for r in roots]
    for nt in (1, 3, 0):
        many = ep.flatten_local_areas(roots, 4, nt)
        assert len(many) == len(one)
        for a, b in zip(one, many):
            assert (a is None) == (b is None)
            if a is None:
                continue
            fa, fb = flat_arrays(a, 12, 3), flat_arrays(b, 12, 3)
            assert fa.keys() == fb.keys()
            for k in fa:
                assert np.array_equal(fa[k], fb[k]), k
    assert ep.flatten_local_areas([], 4) == []
    assert ep.flatten_local_area(10 ** 7, 4) is None                     # an unknown keyframe has no area: None, as one by one
    bad = ep.flatten_local_areas([6, 10 ** 7], 4, 2)
    assert bad[0] is not None and bad[1] is None
    # same through the SE(2) engine (the other instantiation) on a loop-closure tour
    ds2 = D.graphslam_se2_tour(150, seed=2)
    e2 = sb.RbaEngine(sb.default_engine_params(**dict(S2, noise_lambda=D.se2_noise_lambda(ds2["sigma_xy"], ds2["sigma_yaw"]))))
    for f in ds2["frames"]:
        D.feed(e2, f, run_local_optimization=False)
    roots = list(range(10, 150, 7))
    for a, b in zip([e2.flatten_local_area(r, 3) for r in roots], e2.flatten_local_areas(roots, 3, 4)):
        fa, fb = flat_arrays(a, 3, 3), flat_arrays(b, 3, 3)
        for k in fa:
            assert np.array_equal(fa[k], fb[k]), k
