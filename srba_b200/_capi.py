"""ctypes binding of the in-tree shared library (include/srba_b200.h + include/srba_b200_engine.h).

There is NO Python/CPU fallback: if libsrba_b200.so is missing this raises, and every compute entry point of the
library fails with SRBA_ERR_NO_DEVICE when no CUDA device is present."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SRBA_B200_LIB") or os.path.join(_HERE, "libsrba_b200.so")   # SRBA_B200_LIB: an experiment build (srba_b200/build.py)

POSE_SE2, POSE_SE3 = 0, 1
OBS_RELPOSES2D, OBS_CARTESIAN3D, OBS_MONO, OBS_STEREO = 0, 1, 2, 3
SENSORPOSE_NONE, SENSORPOSE_SE3 = 0, 1
NOISE_IDENTITY, NOISE_CONST_MATRIX = 0, 1
SOLVER_SCHUR_DENSE, SOLVER_SCHUR_SPARSE, SOLVER_NO_SCHUR_SPARSE = 0, 1, 2  # RBA_OPTIONS::solver_t (srba_options_solver.h:21-77)
OBS_DIMS = {OBS_RELPOSES2D: 3, OBS_CARTESIAN3D: 3, OBS_MONO: 2, OBS_STEREO: 4}
INVALID = 2**64 - 1
SRBA_OK, ERR_BAD_ARGUMENT, ERR_UNSUPPORTED_MODEL, ERR_RANK_DEFICIENT, ERR_CUDA, ERR_NO_DEVICE, ERR_INTERNAL, ERR_NON_FINITE = 0, -1, -2, -3, -4, -5, -6, -7

i32, u64, dbl = C.c_int32, C.c_uint64, C.c_double
pi32, pdbl, pu8, pu64 = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)


class Model(C.Structure):
    _fields_ = [("pose_kind", i32), ("obs_kind", i32), ("sensor_pose_kind", i32), ("noise_kind", i32),
                ("sensor_pose", dbl * 12), ("noise_std", dbl), ("noise_lambda", dbl * 16), ("cam_left", dbl * 4), ("cam_right", dbl * 4),
                ("right_cam_R2L", dbl * 12), ("max_iters", i32), ("use_robust_kernel", i32), ("kernel_param", dbl),
                ("max_error_per_obs_to_stop", dbl), ("max_rho", dbl), ("max_lambda", dbl), ("min_error_reduction_ratio_to_relinearize", dbl),
                ("solver_kind", i32), ("reserved_", i32)]


_PROBLEM_INTS = ["n_edges", "n_unk_edges", "n_lms", "n_unk_lms", "n_pairs", "n_path", "n_obs", "n_blkA", "n_blkF", "n_HAp", "n_HAp_terms",
                 "n_HApf", "n_HApf_terms", "n_S", "n_S_terms", "n_total_obs_rows", "reserved_"]
_PROBLEM_PTRS = [("edge_pose", pdbl), ("unk_edge", pi32), ("lm_pos", pdbl), ("unk_lm", pi32), ("pair_path_ptr", pi32), ("path_edge", pi32),
                 ("path_dir", pu8), ("obs_z", pdbl), ("obs_slot", pi32), ("obs_lm", pi32), ("obs_mult", pi32), ("colA_ptr", pi32),
                 ("blkA_obs", pi32), ("blkA_slotA", pi32), ("blkA_slotD", pi32), ("blkA_normal_dir", pu8), ("colF_ptr", pi32), ("blkF_obs", pi32),
                 ("blkF_slot", pi32), ("row_ptr", pi32), ("row_blk", pi32), ("HAp_row", pi32), ("HAp_col", pi32), ("HAp_term_ptr", pi32),
                 ("HAp_term_a", pi32), ("HAp_term_b", pi32), ("HApf_row", pi32), ("HApf_col", pi32), ("HApf_term_ptr", pi32),
                 ("HApf_term_a", pi32), ("HApf_term_f", pi32), ("rowApf_ptr", pi32), ("colApf_ptr", pi32), ("colApf_idx", pi32),
                 ("S_row", pi32), ("S_col", pi32), ("S_HAp", pi32), ("S_term_ptr", pi32), ("S_term_rowblk", pi32), ("S_term_colblk", pi32)]


class Problem(C.Structure):
    _fields_ = [("model", Model)] + [(n, i32) for n in _PROBLEM_INTS] + _PROBLEM_PTRS


class Stats(C.Structure):
    _fields_ = [(n, i32) for n in ("num_iters", "num_trials", "num_accepted", "num_relinearized", "num_chol_failures", "num_jacobians",
                                   "hessian_valid", "status")] + [(n, dbl) for n in ("total_sqr_error_init", "total_sqr_error_final",
                                                                                     "obs_rmse", "lambda_final")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [("stats", Stats), ("edge_pose", pdbl), ("lm_pos", pdbl), ("slot_pose", pdbl), ("obs_valid", pu8), ("lm_invertible", pu8),
                ("Hf_diag", pdbl), ("dense_hessian", pdbl), ("trace_cap", i32), ("trace_len", i32), ("trace", pdbl), ("trace_delta", pdbl)]


class EngineParams(C.Structure):
    _fields_ = [("pose_kind", i32), ("obs_kind", i32), ("sensor_pose_kind", i32), ("noise_kind", i32), ("sensor_pose", dbl * 6),
                ("noise_std", dbl), ("noise_lambda", dbl * 16), ("cam_left", dbl * 4), ("cam_right", dbl * 4), ("right_cam_pose", dbl * 7),
                ("max_tree_depth", u64), ("max_optimize_depth", u64), ("optimize_new_edges_alone", i32), ("use_robust_kernel", i32),
                ("use_robust_kernel_stage1", i32), ("cov_recovery", i32), ("kernel_param", dbl), ("max_iters", u64),
                ("max_error_per_obs_to_stop", dbl), ("max_rho", dbl), ("max_lambda", dbl), ("min_error_reduction_ratio_to_relinearize", dbl),
                ("submap_size", u64), ("min_obs_to_loop_closure", u64), ("solver_kind", i32), ("reserved_", i32)]


class OptInfo(C.Structure):
    _fields_ = [(n, u64) for n in ("num_observations", "num_jacobians", "num_kf2kf_edges_optimized", "num_kf2lm_edges_optimized",
                                   "num_total_scalar_optimized", "num_kf_optimized", "num_lm_optimized", "num_span_tree_numeric_updates")] + [
        ("obs_rmse", dbl), ("total_sqr_error_init", dbl), ("total_sqr_error_final", dbl), ("num_lm_trials", u64), ("num_iters", u64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class NewEdge(C.Structure):
    _fields_ = [("id", u64), ("has_approx_init_val", i32), ("loopclosure_observer_kf", u64), ("loopclosure_base_kf", u64)]


# every symbol include/srba_b200.h and include/srba_b200_engine.h declare
EXPORTED_SYMBOLS = [
    "srba_b200_optimize", "srba_b200_optimize_batch", "srba_b200_batch_create", "srba_b200_batch_upload", "srba_b200_batch_run",
    "srba_b200_batch_download", "srba_b200_batch_unpack", "srba_b200_batch_destroy", "srba_b200_batch_h2d_bytes", "srba_b200_batch_d2h_bytes",
    "srba_b200_kernel_timing", "srba_b200_kernel_timing_get", "srba_b200_batch_update_values", "srba_b200_batch_upload_values", "srba_b200_batch_kernel_launches", "srba_b200_batch_stage_cycles", "srba_b200_problem_plan", "srba_b200_batch_problem_plan", "srba_b200_problem_ldl_counts", "srba_b200_problem_wave_counts", "srba_b200_debug_ldl_schedule", "srba_b200_debug_level_schedule", "srba_b200_batch_algorithmic_bytes", "srba_b200_eval_overall_squared_error", "srba_b200_last_error", "srba_b200_device_count", "srba_b200_version",
    "srba_b200_release_workspace", "srba_b200_set_option", "srba_b200_get_option",
    "srba_engine_default_params", "srba_engine_create", "srba_engine_destroy", "srba_engine_set_params", "srba_engine_define_new_keyframe",
    "srba_engine_optimize_local_area", "srba_engine_optimize_local_area_obs", "srba_engine_optimize_edges", "srba_engine_select_local_area", "srba_engine_flatten",
    "srba_engine_flatten_local_area", "srba_engine_flatten_local_area_obs", "srba_engine_flatten_local_areas", "srba_flat_problem", "srba_flat_apply", "srba_flat_ids", "srba_flat_touched_edges", "srba_flat_destroy",
    "srba_engine_record_problems", "srba_engine_num_recorded", "srba_engine_take_recorded", "srba_engine_trace", "srba_engine_last_trace",
    "srba_engine_num_keyframes", "srba_engine_num_edges", "srba_engine_num_observations", "srba_engine_pose_nvals", "srba_engine_get_edge",
    "srba_engine_set_edge_pose", "srba_engine_num_unknown_lms", "srba_engine_get_unknown_lms", "srba_engine_set_lm_pos",
    "srba_engine_get_next_edge", "srba_engine_get_num", "srba_engine_set_num", "srba_engine_get_all_edges_path", "srba_engine_alloc_keyframe",
    "srba_engine_create_kf2kf_edge", "srba_engine_flatten_seconds", "srba_engine_gpu_seconds", "srba_engine_total_trials",
    "srba_engine_eval_overall_squared_error", "srba_engine_complete_spanning_tree", "srba_engine_get_jacob_validity", "srba_engine_last_error",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("srba_b200: %s is missing — build it with `python -m srba_b200.build` (there is no CPU fallback)" % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp = C.c_void_p
    pp, pr = C.POINTER(Problem), C.POINTER(Result)
    L.srba_b200_optimize.restype = C.c_int
    L.srba_b200_optimize.argtypes = [pp, pr]
    L.srba_b200_optimize_batch.restype = C.c_int
    L.srba_b200_optimize_batch.argtypes = [i32, C.POINTER(pp), C.POINTER(pr)]
    L.srba_b200_batch_create.restype = vp
    L.srba_b200_batch_create.argtypes = [i32, C.POINTER(pp), i32]
    for n in ("upload", "run", "download"):
        f = getattr(L, "srba_b200_batch_" + n)
        f.restype = C.c_int
        f.argtypes = [vp, vp]
    L.srba_b200_kernel_timing.argtypes = [i32]
    L.srba_b200_kernel_timing_get.argtypes = [pdbl]
    L.srba_b200_batch_update_values.restype = C.c_int
    L.srba_b200_batch_update_values.argtypes = [vp, i32, pdbl, pdbl, pdbl]
    L.srba_b200_batch_upload_values.restype = C.c_int
    L.srba_b200_batch_upload_values.argtypes = [vp, vp]
    L.srba_b200_batch_unpack.restype = C.c_int
    L.srba_b200_batch_unpack.argtypes = [vp, i32, pr]
    L.srba_b200_batch_destroy.argtypes = [vp]
    for n in ("h2d_bytes", "d2h_bytes", "kernel_launches"):
        f = getattr(L, "srba_b200_batch_" + n)
        f.restype = C.c_int64
        f.argtypes = [vp]
    L.srba_b200_batch_algorithmic_bytes.argtypes = [vp, pdbl]
    L.srba_b200_batch_stage_cycles.restype = C.c_int
    L.srba_b200_batch_stage_cycles.argtypes = [vp, i32, pu64]
    pi32 = C.POINTER(C.c_int32)
    L.srba_b200_debug_ldl_schedule.restype = C.c_int
    L.srba_b200_debug_ldl_schedule.argtypes = [pp, pi32, pi32, pi32, pi32, pi32, pi32, pi32, i32, i32, i32]
    L.srba_b200_debug_level_schedule.restype = C.c_int
    L.srba_b200_debug_level_schedule.argtypes = [pp, pi32, pi32, pi32, pi32, pi32, pi32, pi32]
    L.srba_b200_problem_ldl_counts.restype = C.c_int
    L.srba_b200_problem_ldl_counts.argtypes = [pp, C.POINTER(C.c_int64)]
    L.srba_b200_batch_problem_plan.restype = C.c_int
    L.srba_b200_batch_problem_plan.argtypes = [vp, i32, C.POINTER(C.c_int64)]
    L.srba_b200_problem_plan.restype = C.c_int
    L.srba_b200_problem_plan.argtypes = [pp, C.POINTER(C.c_int64)]
    L.srba_b200_last_error.restype = C.c_char_p
    L.srba_b200_version.restype = C.c_char_p
    L.srba_b200_device_count.restype = C.c_int
    L.srba_b200_set_option.restype = C.c_int
    L.srba_b200_set_option.argtypes = [C.c_char_p, i32]
    L.srba_b200_get_option.restype = i32
    L.srba_b200_get_option.argtypes = [C.c_char_p]
    # engine
    pep, poi = C.POINTER(EngineParams), C.POINTER(OptInfo)
    L.srba_engine_default_params.argtypes = [pep]
    L.srba_engine_create.restype = vp
    L.srba_engine_create.argtypes = [pep]
    L.srba_engine_destroy.argtypes = [vp]
    L.srba_engine_set_params.restype = C.c_int
    L.srba_engine_set_params.argtypes = [vp, pep]
    L.srba_engine_define_new_keyframe.restype = C.c_int
    L.srba_engine_define_new_keyframe.argtypes = [vp, u64, pu64, pdbl, pu8, pu8, pdbl, C.c_int, pu64, C.POINTER(NewEdge), u64, pu64, poi, poi]
    L.srba_engine_optimize_local_area.restype = C.c_int
    L.srba_engine_optimize_local_area.argtypes = [vp, u64, u64, C.c_int, C.c_int, u64, u64, poi]
    L.srba_engine_optimize_local_area_obs.restype = C.c_int
    L.srba_engine_optimize_local_area_obs.argtypes = [vp, u64, u64, u64, pu64, poi]
    L.srba_engine_optimize_edges.restype = C.c_int
    L.srba_engine_optimize_edges.argtypes = [vp, u64, pu64, u64, pu64, poi]
    L.srba_engine_select_local_area.restype = C.c_int
    L.srba_engine_select_local_area.argtypes = [vp, u64, u64, pu64, u64, pu64, pu64, u64, pu64]
    L.srba_engine_flatten.restype = vp
    L.srba_engine_flatten.argtypes = [vp, u64, pu64, u64, pu64]
    L.srba_engine_flatten_local_area.restype = vp
    L.srba_engine_flatten_local_area.argtypes = [vp, u64, u64]
    L.srba_engine_flatten_local_areas.restype = C.c_int
    L.srba_engine_flatten_local_areas.argtypes = [vp, u64, pu64, u64, i32, C.POINTER(vp)]
    L.srba_engine_flatten_local_area_obs.restype = vp
    L.srba_engine_flatten_local_area_obs.argtypes = [vp, u64, u64, u64, pu64]
    L.srba_flat_problem.restype = pp
    L.srba_flat_problem.argtypes = [vp]
    L.srba_flat_apply.restype = C.c_int
    L.srba_flat_apply.argtypes = [vp, vp, pr, poi]
    L.srba_flat_ids.argtypes = [vp, pu64, pu64, pu64, pu64]
    L.srba_flat_touched_edges.restype = u64
    L.srba_flat_touched_edges.argtypes = [vp, pu64, u64]
    L.srba_flat_destroy.argtypes = [vp]
    L.srba_engine_record_problems.argtypes = [vp, C.c_int]
    L.srba_engine_num_recorded.restype = u64
    L.srba_engine_num_recorded.argtypes = [vp]
    L.srba_engine_take_recorded.restype = vp
    L.srba_engine_take_recorded.argtypes = [vp, u64]
    L.srba_engine_trace.argtypes = [vp, C.c_int]
    L.srba_engine_last_trace.restype = u64
    L.srba_engine_last_trace.argtypes = [vp, pdbl, pdbl, u64, pu64]
    for n in ("num_keyframes", "num_edges", "num_observations", "num_unknown_lms", "total_trials", "alloc_keyframe"):
        f = getattr(L, "srba_engine_" + n)
        f.restype = u64
        f.argtypes = [vp]
    L.srba_engine_pose_nvals.restype = C.c_int
    L.srba_engine_pose_nvals.argtypes = [vp]
    L.srba_engine_get_edge.argtypes = [vp, u64, pu64, pu64, pdbl]
    L.srba_engine_set_edge_pose.argtypes = [vp, u64, pdbl]
    L.srba_engine_get_unknown_lms.argtypes = [vp, pu64, pu64, pdbl]
    L.srba_engine_set_lm_pos.argtypes = [vp, u64, pdbl]
    L.srba_engine_get_next_edge.restype = u64
    L.srba_engine_get_next_edge.argtypes = [vp, u64, pu64, pu64, pu64, u64]
    L.srba_engine_get_num.restype = u64
    L.srba_engine_get_num.argtypes = [vp, u64, pu64, pdbl, u64]
    L.srba_engine_set_num.argtypes = [vp, u64, u64, pdbl]
    L.srba_engine_get_jacob_validity.restype = u64
    L.srba_engine_get_jacob_validity.argtypes = [vp, pu8, u64]
    L.srba_engine_eval_overall_squared_error.restype = C.c_int
    L.srba_engine_eval_overall_squared_error.argtypes = [vp, pdbl]
    L.srba_engine_complete_spanning_tree.restype = C.c_int64
    L.srba_engine_complete_spanning_tree.argtypes = [vp, u64, u64, pu64, pdbl, u64]
    L.srba_engine_get_all_edges_path.restype = C.c_int64
    L.srba_engine_get_all_edges_path.argtypes = [vp, u64, u64, pu64, u64]
    L.srba_engine_create_kf2kf_edge.restype = u64
    L.srba_engine_create_kf2kf_edge.argtypes = [vp, u64, u64, u64, pdbl]
    L.srba_engine_flatten_seconds.restype = dbl
    L.srba_engine_flatten_seconds.argtypes = [vp]
    L.srba_engine_gpu_seconds.restype = dbl
    L.srba_engine_gpu_seconds.argtypes = [vp]
    L.srba_engine_last_error.restype = C.c_char_p
    _lib = L
    return L


def set_option(name, value):
    """srba_b200_set_option (include/srba_b200.h): execution-mode hooks for tests / experiments; returns the previous value."""
    L = lib()
    old = L.srba_b200_get_option(name.encode())
    if L.srba_b200_set_option(name.encode(), int(value)) != SRBA_OK:
        raise ValueError("unknown srba_b200 option %r" % name)
    return old


def ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


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


def default_engine_params(**kw):
    p = EngineParams()
    lib().srba_engine_default_params(C.byref(p))
    return set_fields(p, **kw)


class SrbaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("srba_b200 error %d: %s" % (code, msg))
        self.code = code
