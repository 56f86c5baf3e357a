"""Python mirror of the reference's public class RbaEngine<> (include/srba/RbaEngine.h:66-816) over the native host
engine + CUDA hot path.  Method names, argument meaning and error behaviour follow the reference:
define_new_keyframe (RbaEngine.h:207-211), optimize_local_area (:235-241), optimize_edges (:533-538)."""
import ctypes as C

import numpy as np

from . import _capi as K
from ._capi import lib, ptr


class FlatProblem:
    """A flattened optimize_edges() problem (srba_problem_t, include/srba_b200.h) owned by the native side."""

    def __init__(self, handle):
        self.h = handle
        self.L = lib()
        self.problem = self.L.srba_flat_problem(handle)

    def __del__(self):
        try:
            if self.h:
                self.L.srba_flat_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def p(self):
        return self.problem.contents

    def sizes(self):
        return {n: getattr(self.p, n) for n in K._PROBLEM_INTS if n != "reserved_"}

    def array(self, name, count, dtype):
        a = np.ctypeslib.as_array(getattr(self.p, name), shape=(max(count, 1),))[:count]
        return np.array(a, dtype=dtype)

    def ids(self):
        p = self.p
        k2k, lms = np.zeros(p.n_unk_edges, np.uint64), np.zeros(p.n_unk_lms, np.uint64)
        obs, pairs = np.zeros(p.n_obs, np.uint64), np.zeros((p.n_pairs, 2), np.uint64)
        self.L.srba_flat_ids(self.h, ptr(k2k, C.c_uint64), ptr(lms, C.c_uint64), ptr(obs, C.c_uint64), ptr(pairs, C.c_uint64))
        touched = np.zeros(p.n_edges, np.uint64)
        self.L.srba_flat_touched_edges(self.h, ptr(touched, C.c_uint64), p.n_edges)
        return dict(k2k_ids=k2k, lm_ids=lms, obs_global=obs, pairs=pairs, edge_ids_touched=touched)


class ResultBuffers:
    """Caller-owned output buffers for one problem (srba_result_t)."""

    def __init__(self, p, trace_cap=0, want_dense=False):
        P = 3 if p.model.pose_kind == K.POSE_SE2 else 6
        S = 3 if P == 3 else 12
        self.n_scalars = P * p.n_unk_edges + 3 * p.n_unk_lms
        self.edge_pose = np.zeros((p.n_unk_edges, S))
        self.lm_pos = np.zeros((max(p.n_unk_lms, 1), 3))[:p.n_unk_lms]
        self.slot_pose = np.zeros((2 * p.n_pairs, S))
        self.obs_valid = np.zeros(p.n_obs, np.uint8)
        self.lm_invertible = np.zeros(max(p.n_unk_lms, 1), np.uint8)[:p.n_unk_lms]
        self.Hf_diag = np.zeros((max(p.n_unk_lms, 1), 9))[:p.n_unk_lms]
        self.dense = np.zeros((P * p.n_unk_edges, P * p.n_unk_edges)) if want_dense else None
        self.trace = np.zeros((max(trace_cap, 1), 6))
        self.trace_delta = np.zeros((max(trace_cap, 1), max(self.n_scalars, 1)))
        r = K.Result()
        r.edge_pose = ptr(self.edge_pose, C.c_double)
        r.lm_pos = ptr(self.lm_pos, C.c_double) if p.n_unk_lms else None
        r.slot_pose = ptr(self.slot_pose, C.c_double)
        r.obs_valid = ptr(self.obs_valid, C.c_uint8)
        r.lm_invertible = ptr(self.lm_invertible, C.c_uint8) if p.n_unk_lms else None
        r.Hf_diag = ptr(self.Hf_diag, C.c_double) if p.n_unk_lms else None
        r.dense_hessian = ptr(self.dense, C.c_double) if want_dense else None
        r.trace_cap = trace_cap
        r.trace = ptr(self.trace, C.c_double) if trace_cap else None
        r.trace_delta = ptr(self.trace_delta, C.c_double) if trace_cap else None
        self.r = r

    @property
    def stats(self):
        return self.r.stats.as_dict()

    def trace_rows(self):
        n = self.r.trace_len
        return self.trace[:n].copy(), self.trace_delta[:n, :self.n_scalars].copy()


def optimize(flat, trace_cap=0, want_dense=False):
    """srba_b200_optimize() on one flattened problem; returns the filled ResultBuffers."""
    L = lib()
    res = ResultBuffers(flat.p, trace_cap, want_dense)
    rc = L.srba_b200_optimize(flat.problem, C.byref(res.r))
    if rc != K.SRBA_OK:
        raise K.SrbaError(rc, L.srba_b200_last_error().decode())
    return res


class RbaEngine:
    def __init__(self, params):
        self.L = lib()
        self.params = params
        self.h = self.L.srba_engine_create(C.byref(params))
        if not self.h:
            raise RuntimeError(self.L.srba_engine_last_error().decode())
        self.O = K.OBS_DIMS[params.obs_kind]
        self.nv = self.L.srba_engine_pose_nvals(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.srba_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _err(self):
        return RuntimeError(self.L.srba_engine_last_error().decode())

    def set_params(self, params):
        if self.L.srba_engine_set_params(self.h, C.byref(params)) != 0:
            raise self._err()
        self.params = params

    def define_new_keyframe(self, feat_id, obs_data, is_fixed=None, is_unknown_with_init_val=None, feat_rel_pos=None,
                            run_local_optimization=True):
        feat_id = np.ascontiguousarray(feat_id, dtype=np.uint64)
        n = len(feat_id)
        obs_data = np.ascontiguousarray(obs_data, dtype=np.float64).reshape(n, self.O)
        is_fixed = np.zeros(n, np.uint8) if is_fixed is None else np.ascontiguousarray(is_fixed, dtype=np.uint8)
        iu = np.zeros(n, np.uint8) if is_unknown_with_init_val is None else np.ascontiguousarray(is_unknown_with_init_val, dtype=np.uint8)
        frp = np.zeros((n, 3)) if feat_rel_pos is None else np.ascontiguousarray(feat_rel_pos, dtype=np.float64).reshape(n, 3)
        kf, ne = C.c_uint64(), C.c_uint64()
        cap = n + 2   # the edge-creation policy creates at most one edge per observed base area + the link to the submap centre
        edges = (K.NewEdge * cap)()
        opt, stg1 = K.OptInfo(), K.OptInfo()
        rc = self.L.srba_engine_define_new_keyframe(self.h, n, ptr(feat_id, C.c_uint64), ptr(obs_data, C.c_double), ptr(is_fixed, C.c_uint8),
                                                    ptr(iu, C.c_uint8), ptr(frp, C.c_double), int(run_local_optimization), C.byref(kf), edges, cap,
                                                    C.byref(ne), C.byref(opt), C.byref(stg1))
        if rc != 0:
            raise self._err()
        if ne.value > cap:
            raise RuntimeError("define_new_keyframe created %d edges, more than the %d-entry output buffer" % (ne.value, cap))
        return {"kf_id": kf.value,
                "created_edges": [dict(id=e.id, has_approx_init_val=bool(e.has_approx_init_val),
                                       loopclosure_observer_kf=e.loopclosure_observer_kf, loopclosure_base_kf=e.loopclosure_base_kf)
                                  for e in edges[:ne.value]],
                "optimize_results": opt.as_dict(), "optimize_results_stg1": stg1.as_dict()}

    def optimize_local_area(self, root, win_size, optimize_k2k_edges=True, optimize_landmarks=True, max_visitable_kf_id=K.INVALID,
                            dont_optimize_landmarks_seen_less_than_n_times=2, observation_indices_to_optimize=None):
        o = K.OptInfo()
        if observation_indices_to_optimize is not None and len(observation_indices_to_optimize):   # RbaEngine.h:235-241
            lst = np.ascontiguousarray(observation_indices_to_optimize, dtype=np.uint64)
            if self.L.srba_engine_optimize_local_area_obs(self.h, root, win_size, len(lst), ptr(lst, C.c_uint64), C.byref(o)) != 0:
                raise self._err()
            return o.as_dict()
        rc = self.L.srba_engine_optimize_local_area(self.h, root, win_size, int(optimize_k2k_edges), int(optimize_landmarks), max_visitable_kf_id,
                                                    dont_optimize_landmarks_seen_less_than_n_times, C.byref(o))
        if rc != 0:
            raise self._err()
        return o.as_dict()

    def optimize_edges(self, k2k, lms=()):
        a, b = np.ascontiguousarray(k2k, dtype=np.uint64), np.ascontiguousarray(lms, dtype=np.uint64)
        o = K.OptInfo()
        rc = self.L.srba_engine_optimize_edges(self.h, len(a), ptr(a, C.c_uint64), len(b), ptr(b, C.c_uint64), C.byref(o))
        if rc != 0:
            raise self._err()
        return o.as_dict()

    def select_local_area(self, root, win_size):
        cap = max(16, self.num_edges + 16)
        capl = max(16, self.num_observations + 16)
        a, b = np.zeros(cap, np.uint64), np.zeros(capl, np.uint64)
        na, nb = C.c_uint64(), C.c_uint64()
        rc = self.L.srba_engine_select_local_area(self.h, root, win_size, ptr(a, C.c_uint64), cap, C.byref(na), ptr(b, C.c_uint64), capl, C.byref(nb))
        if rc != 0:
            raise self._err()
        return a[:na.value].copy(), b[:nb.value].copy()

    # flattening ---------------------------------------------------------------
    def flatten(self, k2k, lms=()):
        a, b = np.ascontiguousarray(k2k, dtype=np.uint64), np.ascontiguousarray(lms, dtype=np.uint64)
        h = self.L.srba_engine_flatten(self.h, len(a), ptr(a, C.c_uint64), len(b), ptr(b, C.c_uint64))
        return FlatProblem(h) if h else None

    def flatten_local_area(self, root, win_size, observation_indices_to_optimize=None):
        if observation_indices_to_optimize is not None and len(observation_indices_to_optimize):
            lst = np.ascontiguousarray(observation_indices_to_optimize, dtype=np.uint64)
            h = self.L.srba_engine_flatten_local_area_obs(self.h, root, win_size, len(lst), ptr(lst, C.c_uint64))
            if not h:
                raise self._err()
            return FlatProblem(h)
        h = self.L.srba_engine_flatten_local_area(self.h, root, win_size)
        return FlatProblem(h) if h else None

    def flatten_local_areas(self, roots, win_size, n_threads=0):
        """The local areas of many keyframes of the current map state as independent problems (snapshot semantics), flattened
        by n_threads host threads (0 = all); a list with None where an area has no unknowns."""
        r = np.ascontiguousarray(roots, dtype=np.uint64)
        out = (C.c_void_p * max(1, len(r)))()
        if self.L.srba_engine_flatten_local_areas(self.h, len(r), ptr(r, C.c_uint64), win_size, int(n_threads), out) != 0:
            raise self._err()
        return [FlatProblem(out[i]) if out[i] else None for i in range(len(r))]

    def apply(self, flat, res):
        o = K.OptInfo()
        rc = self.L.srba_flat_apply(self.h, flat.h, C.byref(res.r), C.byref(o))
        if rc != 0:
            raise self._err()
        return o.as_dict()

    def record_problems(self, on=True):
        self.L.srba_engine_record_problems(self.h, int(on))

    def take_recorded(self):
        out = []
        for i in range(self.L.srba_engine_num_recorded(self.h)):
            h = self.L.srba_engine_take_recorded(self.h, i)
            if h:
                out.append(FlatProblem(h))
        return out

    def trace(self, cap):
        self.L.srba_engine_trace(self.h, cap)

    def last_trace(self):
        ns = C.c_uint64()
        n = self.L.srba_engine_last_trace(self.h, None, None, 0, C.byref(ns))
        t, d = np.zeros((max(n, 1), 6)), np.zeros((max(n, 1), max(ns.value, 1)))
        self.L.srba_engine_last_trace(self.h, ptr(t, C.c_double), ptr(d, C.c_double), n, C.byref(ns))
        return t[:n], d[:n, :ns.value]

    # state ----------------------------------------------------------------------
    @property
    def num_keyframes(self):
        return self.L.srba_engine_num_keyframes(self.h)

    @property
    def num_edges(self):
        return self.L.srba_engine_num_edges(self.h)

    @property
    def num_observations(self):
        return self.L.srba_engine_num_observations(self.h)

    def get_edge(self, i):
        f, t = C.c_uint64(), C.c_uint64()
        p = np.zeros(self.nv)
        self.L.srba_engine_get_edge(self.h, i, C.byref(f), C.byref(t), ptr(p, C.c_double))
        return f.value, t.value, p

    def edges(self):
        n = self.num_edges
        ft, poses = np.zeros((n, 2), np.uint64), np.zeros((n, self.nv))
        for i in range(n):
            f, t, p = self.get_edge(i)
            ft[i] = (f, t)
            poses[i] = p
        return ft, poses

    def set_edge_pose(self, i, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        self.L.srba_engine_set_edge_pose(self.h, i, ptr(p, C.c_double))

    def unknown_lms(self):
        n = self.L.srba_engine_num_unknown_lms(self.h)
        ids, base, pos = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros((n, 3))
        self.L.srba_engine_get_unknown_lms(self.h, ptr(ids, C.c_uint64), ptr(base, C.c_uint64), ptr(pos, C.c_double))
        return ids, base, pos

    def set_lm_pos(self, lm_id, pos):
        p = np.ascontiguousarray(pos, dtype=np.float64)
        self.L.srba_engine_set_lm_pos(self.h, lm_id, ptr(p, C.c_double))

    def next_edge(self, kf):
        n = self.L.srba_engine_get_next_edge(self.h, kf, None, None, None, 0)
        t, nx, d = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self.L.srba_engine_get_next_edge(self.h, kf, ptr(t, C.c_uint64), ptr(nx, C.c_uint64), ptr(d, C.c_uint64), n)
        return t, nx, d

    def num(self, kf):
        n = self.L.srba_engine_get_num(self.h, kf, None, None, 0)
        t, poses = np.zeros(n, np.uint64), np.zeros((n, self.nv))
        self.L.srba_engine_get_num(self.h, kf, ptr(t, C.c_uint64), ptr(poses, C.c_double), n)
        return t, poses

    def jacob_validity(self):
        """rba_state.all_observations_Jacob_validity (srba_types.h:722)."""
        n = self.L.srba_engine_get_jacob_validity(self.h, None, 0)
        v = np.zeros(n, np.uint8)
        self.L.srba_engine_get_jacob_validity(self.h, ptr(v, C.c_uint8), n)
        return v

    def eval_overall_squared_error(self):
        """RbaEngine::eval_overall_squared_error (impl/eval_overall_error.h:15-137): every observation of the map, poses from
        complete BFS trees; BFS on the host, pose chains + observation errors + sum on the GPU."""
        out = C.c_double()
        if self.L.srba_engine_eval_overall_squared_error(self.h, C.byref(out)) != 0:
            raise RuntimeError(self.L.srba_engine_last_error().decode())
        return out.value

    def complete_spanning_tree(self, root, max_depth=2 ** 64 - 1):
        """RbaEngine::create_complete_spanning_tree (impl/spantree_create_complete.h:18-127): (ids ascending, poses wrt root)."""
        n = self.L.srba_engine_complete_spanning_tree(self.h, root, max_depth, None, None, 0)
        if n < 0:
            raise RuntimeError(self.L.srba_engine_last_error().decode())
        ids, poses = np.zeros(n, np.uint64), np.zeros((n, self.nv))
        if self.L.srba_engine_complete_spanning_tree(self.h, root, max_depth, ptr(ids, C.c_uint64), ptr(poses, C.c_double), n) < 0:
            raise RuntimeError(self.L.srba_engine_last_error().decode())
        return ids, poses

    def set_num(self, src, dst, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        self.L.srba_engine_set_num(self.h, src, dst, ptr(p, C.c_double))

    def all_edges_path(self, frm, to):
        buf = np.zeros(64, np.uint64)
        n = self.L.srba_engine_get_all_edges_path(self.h, frm, to, ptr(buf, C.c_uint64), len(buf))
        if n > len(buf):   # longer than the first guess: ask again with the reported length
            buf = np.zeros(n, np.uint64)
            n = self.L.srba_engine_get_all_edges_path(self.h, frm, to, ptr(buf, C.c_uint64), len(buf))
        return None if n < 0 else buf[:n].copy()

    def alloc_keyframe(self):
        return self.L.srba_engine_alloc_keyframe(self.h)

    def create_kf2kf_edge(self, new_kf_id, frm, to, init_inv_pose=None):
        ip = None if init_inv_pose is None else np.ascontiguousarray(init_inv_pose, dtype=np.float64)
        r = self.L.srba_engine_create_kf2kf_edge(self.h, new_kf_id, frm, to, ptr(ip, C.c_double))
        if r == K.INVALID:
            raise self._err()
        return r

    def timers(self):
        return dict(flatten_seconds=self.L.srba_engine_flatten_seconds(self.h), gpu_seconds=self.L.srba_engine_gpu_seconds(self.h),
                    total_trials=self.L.srba_engine_total_trials(self.h))
