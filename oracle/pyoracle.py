"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

ctypes view of oracle/_build/libsrba_oracle.so (the CPU restatement of MRPT/srba's RbaEngine<>,
/root/reference/include/srba/RbaEngine.h:66-816).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

POSE_SE2, POSE_SE3 = 0, 1
OBS_RELPOSES2D, OBS_CARTESIAN3D, OBS_MONO, OBS_STEREO = 0, 1, 2, 3
SENSORPOSE_NONE, SENSORPOSE_SE3 = 0, 1
NOISE_IDENTITY, NOISE_CONST_MATRIX = 0, 1
SOLVER_SCHUR_DENSE, SOLVER_SCHUR_SPARSE, SOLVER_NO_SCHUR_SPARSE = 0, 1, 2  # srba_options_solver.h:21-77
OBS_DIMS = {OBS_RELPOSES2D: 3, OBS_CARTESIAN3D: 3, OBS_MONO: 2, OBS_STEREO: 4}
INVALID = 2**64 - 1


class Params(C.Structure):
    _fields_ = [
        ("pose_kind", C.c_int32), ("obs_kind", C.c_int32), ("sensor_pose_kind", C.c_int32), ("noise_kind", C.c_int32),
        ("sensor_pose", C.c_double * 6), ("noise_std", C.c_double), ("noise_lambda", C.c_double * 16),
        ("cam_left", C.c_double * 4), ("cam_right", C.c_double * 4), ("right_cam_pose", C.c_double * 7),
        ("max_tree_depth", C.c_uint64), ("max_optimize_depth", C.c_uint64),
        ("optimize_new_edges_alone", C.c_int32), ("use_robust_kernel", C.c_int32),
        ("use_robust_kernel_stage1", C.c_int32), ("cov_recovery", C.c_int32),
        ("kernel_param", C.c_double), ("max_iters", C.c_uint64),
        ("max_error_per_obs_to_stop", C.c_double), ("max_rho", C.c_double), ("max_lambda", C.c_double),
        ("min_error_reduction_ratio_to_relinearize", C.c_double),
        ("submap_size", C.c_uint64), ("min_obs_to_loop_closure", C.c_uint64),
        ("solver_kind", C.c_int32), ("reserved_", C.c_int32),
    ]


class OptInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "num_observations", "num_jacobians", "num_kf2kf_edges_optimized", "num_kf2lm_edges_optimized",
        "num_total_scalar_optimized", "num_kf_optimized", "num_lm_optimized", "num_span_tree_numeric_updates")] + [
        ("obs_rmse", C.c_double), ("total_sqr_error_init", C.c_double), ("total_sqr_error_final", C.c_double),
        ("num_lm_trials", C.c_uint64), ("num_iters", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class NewEdge(C.Structure):
    _fields_ = [("id", C.c_uint64), ("has_approx_init_val", C.c_int32),
                ("loopclosure_observer_kf", C.c_uint64), ("loopclosure_base_kf", C.c_uint64)]


def build(force=False):
    """Compile the C++ restatement (gcc -O2, the reference's flag, CMakeLists.txt:49)."""
    so = os.path.join(_HERE, "_build", "libsrba_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle_capi.h", "srba_oracle.hpp", "poses.hpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_lib_cache = {}


def lib(variant=""):
    if variant in _lib_cache:
        return _lib_cache[variant]
    so = build()
    if variant:
        so = so.replace("libsrba_oracle.so", "libsrba_oracle_%s.so" % variant)
    L = C.CDLL(so)
    u64, dbl, vp = C.c_uint64, C.c_double, C.c_void_p
    pu64, pdbl, pu8 = C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    L.orc_last_error.restype = C.c_char_p
    L.orc_default_params.argtypes = [C.POINTER(Params)]
    L.orc_create.restype = vp
    L.orc_create.argtypes = [C.POINTER(Params)]
    L.orc_destroy.argtypes = [vp]
    L.orc_set_params.argtypes = [vp, C.POINTER(Params)]
    L.orc_define_new_keyframe.restype = C.c_int
    L.orc_define_new_keyframe.argtypes = [vp, u64, pu64, pdbl, pu8, pu8, pdbl, C.c_int, pu64, C.POINTER(NewEdge), u64, pu64,
                                          C.POINTER(OptInfo), C.POINTER(OptInfo)]
    L.orc_optimize_local_area.restype = C.c_int
    L.orc_optimize_local_area.argtypes = [vp, u64, u64, C.c_int, C.c_int, u64, u64, C.POINTER(OptInfo)]
    L.orc_optimize_local_area_obs.restype = C.c_int
    L.orc_optimize_local_area_obs.argtypes = [vp, u64, u64, u64, pu64, C.POINTER(OptInfo)]
    L.orc_set_num.argtypes = [vp, u64, u64, C.POINTER(C.c_double)]
    L.orc_optimize_edges.restype = C.c_int
    L.orc_optimize_edges.argtypes = [vp, u64, pu64, u64, pu64, C.POINTER(OptInfo)]
    L.orc_alloc_keyframe.restype = u64
    L.orc_alloc_keyframe.argtypes = [vp]
    L.orc_create_kf2kf_edge.restype = u64
    L.orc_create_kf2kf_edge.argtypes = [vp, u64, u64, u64, pdbl]
    L.orc_update_numeric_all.argtypes = [vp]
    L.orc_are_keyframes_connected.restype = C.c_int
    L.orc_are_keyframes_connected.argtypes = [vp, u64, u64]
    L.orc_complete_spanning_tree.restype = u64
    L.orc_complete_spanning_tree.argtypes = [vp, u64, u64, pu64, pdbl, u64]
    L.orc_find_path_bfs_len.restype = C.c_int64
    L.orc_find_path_bfs_len.argtypes = [vp, u64, u64]
    for n in ("orc_num_keyframes", "orc_num_edges", "orc_num_observations", "orc_num_unknown_lms", "orc_num_known_lms",
              "orc_trace_num_calls", "orc_total_opt_trials"):
        getattr(L, n).restype = u64
        getattr(L, n).argtypes = [vp]
    L.orc_pose_nvals.restype = C.c_int
    L.orc_pose_nvals.argtypes = [vp]
    L.orc_get_edge.argtypes = [vp, u64, pu64, pu64, pdbl]
    L.orc_set_edge_pose.argtypes = [vp, u64, pdbl]
    L.orc_get_unknown_lms.argtypes = [vp, pu64, pu64, pdbl]
    L.orc_get_known_lms.argtypes = [vp, pu64, pu64, pdbl]
    L.orc_set_lm_pos.argtypes = [vp, u64, pdbl]
    L.orc_get_next_edge.restype = u64
    L.orc_get_next_edge.argtypes = [vp, u64, pu64, pu64, pu64, u64]
    L.orc_get_num.restype = u64
    L.orc_get_num.argtypes = [vp, u64, pu64, pdbl, pu8, u64]
    L.orc_get_all_edges_path.restype = C.c_int64
    L.orc_get_all_edges_path.argtypes = [vp, u64, u64, pu64, u64]
    L.orc_trace_enable.argtypes = [vp, C.c_int]
    L.orc_trace_clear.argtypes = [vp]
    L.orc_trace_call_sizes.argtypes = [vp, u64, pu64, pu64, pu64, pu64, pu64]
    L.orc_trace_call_data.argtypes = [vp, u64, pu64, pu64, pu64, pdbl, pdbl, pdbl, pdbl]
    L.orc_total_opt_seconds.restype = dbl
    L.orc_total_opt_seconds.argtypes = [vp]
    L.orc_reset_opt_timer.argtypes = [vp]
    L.orc_schur_case.restype = C.c_int
    L.orc_schur_case.argtypes = [u64, u64, pu8, pdbl, pdbl, dbl, pdbl, pdbl, pu64]
    L.orc_get_jacob_validity.restype = u64
    L.orc_get_jacob_validity.argtypes = [vp, pu8, u64]
    L.orc_eval_overall_squared_error.restype = dbl
    L.orc_eval_overall_squared_error.argtypes = [vp]
    L.orc_last_dense_hessian.restype = u64
    L.orc_last_dense_hessian.argtypes = [vp, pdbl, u64, C.POINTER(C.c_int)]
    L.orc_solver_case.restype = C.c_int
    L.orc_solver_case.argtypes = [C.c_int, u64, u64, pu8, pdbl, pdbl, dbl, pdbl, pdbl, C.POINTER(C.c_int)]
    L.orc_eval_dh_dAp.restype = C.c_int
    L.orc_eval_dh_dAp.argtypes = [C.POINTER(Params), pdbl, pdbl, pdbl, C.c_int, pdbl, pdbl]
    L.orc_eval_dh_df.restype = C.c_int
    L.orc_eval_dh_df.argtypes = [C.POINTER(Params), pdbl, pdbl, pdbl]
    L.orc_predict.restype = C.c_int
    L.orc_predict.argtypes = [C.POINTER(Params), pdbl, pdbl, pdbl, pdbl]
    L.orc_pose_compose.argtypes = [C.c_int, pdbl, pdbl, pdbl]
    L.orc_pose_inverse.argtypes = [C.c_int, pdbl, pdbl]
    L.orc_pose_exp.argtypes = [C.c_int, pdbl, pdbl]
    L.orc_pose_from_xyzypr.argtypes = [pdbl, pdbl]
    _lib_cache[variant] = L
    return L


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    set_fields(p, **kw)
    return p


def set_fields(p, **kw):
    for k, v in kw.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            v = np.asarray(v, dtype=np.float64).ravel()
            for i in range(len(v)):
                cur[i] = float(v[i])
        else:
            setattr(p, k, v)
    return p


class OracleError(RuntimeError):
    pass


class OracleEngine:
    """Mirror of RbaEngine<> (RbaEngine.h) over the CPU oracle."""

    def __init__(self, params, variant=""):
        self.L = lib(variant)
        self.params = params
        self.h = self.L.orc_create(C.byref(params))
        if not self.h:
            raise OracleError(self.L.orc_last_error().decode())
        self.O = OBS_DIMS[params.obs_kind]
        self.nv = self.L.orc_pose_nvals(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_params(self, params):
        self.params = params
        self.L.orc_set_params(self.h, C.byref(params))

    def define_new_keyframe(self, feat_id, obs_data, is_fixed=None, is_unknown_with_init_val=None, feat_rel_pos=None,
                            run_local_optimization=True):
        feat_id = np.ascontiguousarray(feat_id, dtype=np.uint64)
        n = len(feat_id)
        obs_data = np.ascontiguousarray(obs_data, dtype=np.float64).reshape(n, self.O)
        is_fixed = np.zeros(n, np.uint8) if is_fixed is None else np.ascontiguousarray(is_fixed, dtype=np.uint8)
        iu = np.zeros(n, np.uint8) if is_unknown_with_init_val is None else np.ascontiguousarray(is_unknown_with_init_val, dtype=np.uint8)
        frp = np.zeros((n, 3)) if feat_rel_pos is None else np.ascontiguousarray(feat_rel_pos, dtype=np.float64).reshape(n, 3)
        kf = C.c_uint64()
        ne = C.c_uint64()
        edges = (NewEdge * 16)()
        opt, stg1 = OptInfo(), OptInfo()
        rc = self.L.orc_define_new_keyframe(self.h, n, _p(feat_id, C.c_uint64), _p(obs_data, C.c_double), _p(is_fixed, C.c_uint8),
                                            _p(iu, C.c_uint8), _p(frp, C.c_double), int(run_local_optimization), C.byref(kf), edges, 16,
                                            C.byref(ne), C.byref(opt), C.byref(stg1))
        if rc != 0:
            raise OracleError(self.L.orc_last_error().decode())
        return {"kf_id": kf.value,
                "created_edges": [dict(id=e.id, has_approx_init_val=bool(e.has_approx_init_val),
                                       loopclosure_observer_kf=e.loopclosure_observer_kf, loopclosure_base_kf=e.loopclosure_base_kf)
                                  for e in edges[:ne.value]],
                "optimize_results": opt.as_dict(), "optimize_results_stg1": stg1.as_dict()}

    def optimize_local_area(self, root, win_size, optimize_k2k_edges=True, optimize_landmarks=True, max_visitable_kf_id=INVALID,
                            dont_optimize_landmarks_seen_less_than_n_times=2, observation_indices_to_optimize=None):
        o = OptInfo()
        if observation_indices_to_optimize is not None and len(observation_indices_to_optimize):
            lst = np.ascontiguousarray(observation_indices_to_optimize, dtype=np.uint64)
            rc = self.L.orc_optimize_local_area_obs(self.h, root, win_size, len(lst), _p(lst, C.c_uint64), C.byref(o))
            if rc != 0:
                raise OracleError(self.L.orc_last_error().decode())
            return o.as_dict()
        rc = self.L.orc_optimize_local_area(self.h, root, win_size, int(optimize_k2k_edges), int(optimize_landmarks), max_visitable_kf_id,
                                            dont_optimize_landmarks_seen_less_than_n_times, C.byref(o))
        if rc != 0:
            raise OracleError(self.L.orc_last_error().decode())
        return o.as_dict()

    def optimize_edges(self, k2k, lms=()):
        a = np.ascontiguousarray(k2k, dtype=np.uint64)
        b = np.ascontiguousarray(lms, dtype=np.uint64)
        o = OptInfo()
        rc = self.L.orc_optimize_edges(self.h, len(a), _p(a, C.c_uint64), len(b), _p(b, C.c_uint64), C.byref(o))
        if rc != 0:
            raise OracleError(self.L.orc_last_error().decode())
        return o.as_dict()

    # low-level graph access -------------------------------------------------
    def alloc_keyframe(self):
        return self.L.orc_alloc_keyframe(self.h)

    def create_kf2kf_edge(self, new_kf_id, frm, to, init_inv_pose=None):
        ip = None if init_inv_pose is None else np.ascontiguousarray(init_inv_pose, dtype=np.float64)
        r = self.L.orc_create_kf2kf_edge(self.h, new_kf_id, frm, to, _p(ip, C.c_double))
        if r == INVALID:
            raise OracleError(self.L.orc_last_error().decode())
        return r

    def update_numeric_all(self):
        self.L.orc_update_numeric_all(self.h)

    def are_keyframes_connected(self, a, b):
        return bool(self.L.orc_are_keyframes_connected(self.h, a, b))

    def complete_spanning_tree(self, root, max_depth):
        n = self.L.orc_complete_spanning_tree(self.h, root, max_depth, None, None, 0)
        ids = np.zeros(n, np.uint64)
        poses = np.zeros((n, self.nv))
        self.L.orc_complete_spanning_tree(self.h, root, max_depth, _p(ids, C.c_uint64), _p(poses, C.c_double), n)
        return ids, poses

    def eval_overall_squared_error(self):
        return self.L.orc_eval_overall_squared_error(self.h)

    def jacob_validity(self):
        n = self.L.orc_get_jacob_validity(self.h, None, 0)
        v = np.zeros(n, np.uint8)
        self.L.orc_get_jacob_validity(self.h, _p(v, C.c_uint8), n)
        return v

    def last_dense_hessian(self):
        valid = C.c_int()
        n2 = self.L.orc_last_dense_hessian(self.h, None, 0, C.byref(valid))
        H = np.zeros(n2)
        self.L.orc_last_dense_hessian(self.h, _p(H, C.c_double), n2, C.byref(valid))
        n = int(round(n2 ** 0.5))
        return H.reshape(n, n), bool(valid.value)

    def find_path_bfs_len(self, a, b):
        return self.L.orc_find_path_bfs_len(self.h, a, b)

    @property
    def num_keyframes(self):
        return self.L.orc_num_keyframes(self.h)

    @property
    def num_edges(self):
        return self.L.orc_num_edges(self.h)

    @property
    def num_observations(self):
        return self.L.orc_num_observations(self.h)

    def get_edge(self, i):
        f, t = C.c_uint64(), C.c_uint64()
        p = np.zeros(self.nv)
        self.L.orc_get_edge(self.h, i, C.byref(f), C.byref(t), _p(p, C.c_double))
        return f.value, t.value, p

    def edges(self):
        n = self.num_edges
        ft = np.zeros((n, 2), np.uint64)
        poses = np.zeros((n, self.nv))
        for i in range(n):
            f, t, p = self.get_edge(i)
            ft[i] = (f, t)
            poses[i] = p
        return ft, poses

    def set_edge_pose(self, i, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        self.L.orc_set_edge_pose(self.h, i, _p(p, C.c_double))

    def unknown_lms(self):
        n = self.L.orc_num_unknown_lms(self.h)
        ids, base, pos = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros((n, 3))
        self.L.orc_get_unknown_lms(self.h, _p(ids, C.c_uint64), _p(base, C.c_uint64), _p(pos, C.c_double))
        return ids, base, pos

    def known_lms(self):
        n = self.L.orc_num_known_lms(self.h)
        ids, base, pos = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros((n, 3))
        self.L.orc_get_known_lms(self.h, _p(ids, C.c_uint64), _p(base, C.c_uint64), _p(pos, C.c_double))
        return ids, base, pos

    def set_lm_pos(self, lm_id, pos):
        p = np.ascontiguousarray(pos, dtype=np.float64)
        self.L.orc_set_lm_pos(self.h, lm_id, _p(p, C.c_double))

    def set_num(self, src, dst, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        self.L.orc_set_num(self.h, src, dst, _p(p, C.c_double))

    def next_edge(self, kf):
        n = self.L.orc_get_next_edge(self.h, kf, None, None, None, 0)
        t, nx, d = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self.L.orc_get_next_edge(self.h, kf, _p(t, C.c_uint64), _p(nx, C.c_uint64), _p(d, C.c_uint64), n)
        return t, nx, d

    def num(self, kf):
        n = self.L.orc_get_num(self.h, kf, None, None, None, 0)
        t, poses, upd = np.zeros(n, np.uint64), np.zeros((n, self.nv)), np.zeros(n, np.uint8)
        self.L.orc_get_num(self.h, kf, _p(t, C.c_uint64), _p(poses, C.c_double), _p(upd, C.c_uint8), n)
        return t, poses, upd

    def all_edges_path(self, frm, to):
        buf = np.zeros(64, np.uint64)
        n = self.L.orc_get_all_edges_path(self.h, frm, to, _p(buf, C.c_uint64), 64)
        return None if n < 0 else buf[:n].copy()

    # traces -----------------------------------------------------------------
    def trace_enable(self, on=True):
        self.L.orc_trace_enable(self.h, int(on))

    def trace_clear(self):
        self.L.orc_trace_clear(self.h)

    def traces(self):
        out = []
        for c in range(self.L.orc_trace_num_calls(self.h)):
            nk, nl, no, nt, ns = (C.c_uint64() for _ in range(5))
            self.L.orc_trace_call_sizes(self.h, c, C.byref(nk), C.byref(nl), C.byref(no), C.byref(nt), C.byref(ns))
            k2k, lms, obs = np.zeros(nk.value, np.uint64), np.zeros(nl.value, np.uint64), np.zeros(no.value, np.uint64)
            ts, ds, g = np.zeros((nt.value, 6)), np.zeros((nt.value, ns.value)), np.zeros(ns.value)
            l0 = C.c_double()
            self.L.orc_trace_call_data(self.h, c, _p(k2k, C.c_uint64), _p(lms, C.c_uint64), _p(obs, C.c_uint64), _p(ts, C.c_double),
                                       _p(ds, C.c_double), _p(g, C.c_double), C.byref(l0))
            out.append(dict(k2k_ids=k2k, lm_ids=lms, obs_idx=obs, trials=ts, deltas=ds, init_grad=g, lambda0=l0.value))
        return out

    def opt_timer(self):
        return self.L.orc_total_opt_seconds(self.h), self.L.orc_total_opt_trials(self.h)

    def reset_opt_timer(self):
        self.L.orc_reset_opt_timer(self.h)


# unit-level helpers -----------------------------------------------------------
def schur_case(n_p, n_f, visible, JA, JF, lam):
    L = lib()
    visible = np.ascontiguousarray(visible, dtype=np.uint8)
    JA = np.ascontiguousarray(JA, dtype=np.float64)
    JF = np.ascontiguousarray(JF, dtype=np.float64)
    n = 6 * n_p
    H = np.zeros((n, n))
    g = np.zeros(n)
    ninv = C.c_uint64()
    rc = L.orc_schur_case(n_p, n_f, _p(visible, C.c_uint8), _p(JA, C.c_double), _p(JF, C.c_double), lam, _p(H, C.c_double), _p(g, C.c_double),
                          C.byref(ninv))
    if rc != 0:
        raise OracleError(L.orc_last_error().decode())
    return H, g, ninv.value


def solver_case(solver_kind, n_p, n_f, visible, JA, JF, lam):
    """One solve(lambda) of solver_engine<> variant `solver_kind` on the schur_case system (gradient = ones)."""
    L = lib()
    visible = np.ascontiguousarray(visible, dtype=np.uint8)
    JA = np.ascontiguousarray(JA, dtype=np.float64)
    JF = np.ascontiguousarray(JF, dtype=np.float64)
    d = np.zeros(6 * n_p + 3 * n_f)
    g = np.zeros(6 * n_p + 3 * n_f)
    ok = C.c_int()
    rc = L.orc_solver_case(solver_kind, n_p, n_f, _p(visible, C.c_uint8), _p(JA, C.c_double), _p(JF, C.c_double), C.c_double(lam), _p(d, C.c_double),
                           _p(g, C.c_double), C.byref(ok))
    if rc != 0:
        raise OracleError(L.orc_last_error().decode())
    return bool(ok.value), d, g


def pose_compose(kind, a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    o = np.zeros_like(a)
    lib().orc_pose_compose(kind, _p(a, C.c_double), _p(b, C.c_double), _p(o, C.c_double))
    return o


def pose_inverse(kind, a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    o = np.zeros_like(a)
    lib().orc_pose_inverse(kind, _p(a, C.c_double), _p(o, C.c_double))
    return o


def pose_exp(kind, v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    o = np.zeros(3 if kind == POSE_SE2 else 12)
    lib().orc_pose_exp(kind, _p(v, C.c_double), _p(o, C.c_double))
    return o


def pose_from_xyzypr(v6):
    v = np.ascontiguousarray(v6, dtype=np.float64)
    o = np.zeros(12)
    lib().orc_pose_from_xyzypr(_p(v, C.c_double), _p(o, C.c_double))
    return o


def eval_dh_dAp(params, A, D, edge_inv_pose, normal_dir, lm_pos):
    o = OBS_DIMS[params.obs_kind]
    p = 3 if params.pose_kind == POSE_SE2 else 6
    J = np.zeros((o, p))
    A = None if A is None else np.ascontiguousarray(A, dtype=np.float64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    e = None if edge_inv_pose is None else np.ascontiguousarray(edge_inv_pose, dtype=np.float64)
    lm = np.ascontiguousarray(lm_pos, dtype=np.float64)
    v = lib().orc_eval_dh_dAp(C.byref(params), _p(A, C.c_double), _p(D, C.c_double), _p(e, C.c_double), int(normal_dir), _p(lm, C.c_double),
                              _p(J, C.c_double))
    return J, v


def eval_dh_df(params, pose, lm_pos):
    o = OBS_DIMS[params.obs_kind]
    J = np.zeros((o, 3))
    pose = None if pose is None else np.ascontiguousarray(pose, dtype=np.float64)
    lm = np.ascontiguousarray(lm_pos, dtype=np.float64)
    v = lib().orc_eval_dh_df(C.byref(params), _p(pose, C.c_double), _p(lm, C.c_double), _p(J, C.c_double))
    return J, v


def predict_err(params, pose, lm_pos, z):
    o = OBS_DIMS[params.obs_kind]
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    lm = np.ascontiguousarray(lm_pos, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    e = np.zeros(o)
    rc = lib().orc_predict(C.byref(params), _p(pose, C.c_double), _p(lm, C.c_double), _p(z, C.c_double), _p(e, C.c_double))
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())
    return e
