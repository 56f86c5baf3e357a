/* srba_b200_engine.h — C view of srba_b200::RbaEngine, the host-side mirror of the reference's public class
 * RbaEngine<KF2KF_POSE,LM,OBS,RBA_OPTIONS> (include/srba/RbaEngine.h:66-816) for the hot path only:
 *   define_new_keyframe   RbaEngine.h:207-211  (impl/define_new_keyframe.h:16-116)
 *   optimize_local_area   RbaEngine.h:235-241  (impl/optimize_local_area.h:15-58)
 *   optimize_edges        RbaEngine.h:533-538  (impl/optimize_edges.h:44-793) -> flatten + srba_b200_optimize()
 * The reference's four template arguments become run-time tags.  The CPU keeps the incremental relative-coordinate
 * graph (integer work, bit-exact with the reference); every numeric step of the optimiser runs on the GPU through
 * include/srba_b200.h.  Used by the Python mirror (srba_b200/engine.py) and by C/C++ callers.
 */
#ifndef SRBA_B200_ENGINE_H
#define SRBA_B200_ENGINE_H
#include <stdint.h>

#include "srba_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* RBA_OPTIONS (RbaEngine.h:39-45) + TAllParameters (RbaEngine.h:463-472), defaults impl/rba_problem_common.h:35-56 and
 * ecps/local_areas_fixed_size.h:30-33 (see srba_engine_default_params). */
typedef struct {
	int32_t pose_kind, obs_kind, sensor_pose_kind, noise_kind;
	double sensor_pose[6];      /* x,y,z,yaw,pitch,roll */
	double noise_std;
	double noise_lambda[16];    /* row-major o x o */
	double cam_left[4], cam_right[4]; /* fx,fy,cx,cy */
	double right_cam_pose[7];   /* x,y,z,qr,qx,qy,qz */
	uint64_t max_tree_depth, max_optimize_depth;
	int32_t optimize_new_edges_alone, use_robust_kernel, use_robust_kernel_stage1, cov_recovery;
	double kernel_param;
	uint64_t max_iters;
	double max_error_per_obs_to_stop, max_rho, max_lambda, min_error_reduction_ratio_to_relinearize;
	uint64_t submap_size, min_obs_to_loop_closure;
	int32_t solver_kind;        /* RBA_OPTIONS::solver_t (RbaEngine.h:44, srba_options_solver.h:21-77): SRBA_SOLVER_* of srba_b200.h */
	int32_t reserved_;
} srba_engine_params_t;

/* TOptimizeExtraOutputInfo (RbaEngine.h:125-177), scalar part (+ two counters the reference does not expose). */
typedef struct {
	uint64_t num_observations, num_jacobians, num_kf2kf_edges_optimized, num_kf2lm_edges_optimized, num_total_scalar_optimized,
	    num_kf_optimized, num_lm_optimized, num_span_tree_numeric_updates;
	double obs_rmse, total_sqr_error_init, total_sqr_error_final;
	uint64_t num_lm_trials, num_iters;
} srba_opt_info_t;

typedef struct { /* TNewEdgeInfo srba_types.h:178-200 */
	uint64_t id;
	int32_t has_approx_init_val;
	uint64_t loopclosure_observer_kf, loopclosure_base_kf;
} srba_new_edge_t;

void srba_engine_default_params(srba_engine_params_t *p);
void *srba_engine_create(const srba_engine_params_t *p);
void srba_engine_destroy(void *h);
int srba_engine_set_params(void *h, const srba_engine_params_t *p); /* 0 ok; -1: rejected, engine unchanged (srba_engine_last_error) */

int srba_engine_define_new_keyframe(void *h, uint64_t n_obs, const uint64_t *feat_id, const double *obs_data, const uint8_t *is_fixed,
                                    const uint8_t *is_unknown_with_init_val, const double *feat_rel_pos, int run_local_optimization,
                                    uint64_t *out_kf_id, srba_new_edge_t *out_edges, uint64_t max_edges, uint64_t *out_n_edges,
                                    srba_opt_info_t *out_opt, srba_opt_info_t *out_opt_stg1);
int srba_engine_optimize_local_area(void *h, uint64_t root, uint64_t win_size, int optimize_k2k_edges, int optimize_landmarks,
                                    uint64_t max_visitable_kf_id, uint64_t dont_optimize_landmarks_seen_less_than_n_times, srba_opt_info_t *out);
/* optimize_local_area(root, win_size, out, default TOptimizeLocalAreaParams, observation_indices_to_optimize) (RbaEngine.h:235-241;
 * impl/optimize_edges.h:218-234): the residual rows are the listed observations (indices into all_observations, repeats counted).  As in
 * the reference the list must contain every observation of the selected unknowns (else -1: the ASSERT_ of compute_minus_gradient.h:55). */
int srba_engine_optimize_local_area_obs(void *h, uint64_t root, uint64_t win_size, uint64_t n_obs_idx, const uint64_t *obs_idx, srba_opt_info_t *out);
int srba_engine_optimize_edges(void *h, uint64_t n_k2k, const uint64_t *k2k, uint64_t n_lm, const uint64_t *lms, srba_opt_info_t *out);

/* Unknown selection of optimize_local_area WITHOUT optimising (impl/bfs_visitor.h + VisitorOptimizeLocalArea). Returns counts. */
int srba_engine_select_local_area(void *h, uint64_t root, uint64_t win_size, uint64_t *k2k_ids, uint64_t k2k_cap, uint64_t *n_k2k,
                                  uint64_t *lm_ids, uint64_t lm_cap, uint64_t *n_lm);

/* Flattening (what the shim does before calling srba_b200_optimize): the returned handle owns every array the
 * srba_problem_t points to.  NULL when there is nothing to optimise.  srba_flat_apply() scatters a result back
 * into the engine exactly as optimize_edges would have (edges, landmarks, numeric spanning tree, validity, info matrices). */
void *srba_engine_flatten(void *h, uint64_t n_k2k, const uint64_t *k2k, uint64_t n_lm, const uint64_t *lms);
void *srba_engine_flatten_local_area(void *h, uint64_t root, uint64_t win_size);
/* The local areas of n keyframes of the CURRENT map state as independent problems (snapshot semantics): flattened by up to n_threads host
 * threads (<= 0: one per hardware thread); out[i] = flat handle, or NULL when area i has no unknowns.  0, or -1 with every handle released. */
int srba_engine_flatten_local_areas(void *h, uint64_t n, const uint64_t *roots, uint64_t win_size, int n_threads, void **out);
void *srba_engine_flatten_local_area_obs(void *h, uint64_t root, uint64_t win_size, uint64_t n_obs_idx, const uint64_t *obs_idx); /* with observation_indices_to_optimize */
const srba_problem_t *srba_flat_problem(const void *flat);
int srba_flat_apply(void *h, const void *flat, const srba_result_t *res, srba_opt_info_t *out);
void srba_flat_ids(const void *flat, uint64_t *k2k_ids, uint64_t *lm_ids, uint64_t *obs_global_idx, uint64_t *pair_from_to);
/* global ids of the k2k edges behind srba_problem_t::edge_pose (local edge index -> k2k edge id); returns their number */
uint64_t srba_flat_touched_edges(const void *flat, uint64_t *edge_ids, uint64_t cap);
void srba_flat_destroy(void *flat);
/* When on, every optimize_edges() call of the engine keeps its flattened problem (pre-optimisation snapshot). */
void srba_engine_record_problems(void *h, int on);
uint64_t srba_engine_num_recorded(void *h);
void *srba_engine_take_recorded(void *h, uint64_t i); /* ownership moves to the caller (srba_flat_destroy) */
void srba_engine_trace(void *h, int trace_cap);         /* keep per-trial traces of the next optimize_edges calls */
uint64_t srba_engine_last_trace(void *h, double *trace6, double *delta, uint64_t cap_trials, uint64_t *n_scalars);

/* state getters (same conventions as the flat pose layout) */
uint64_t srba_engine_num_keyframes(void *h);
uint64_t srba_engine_num_edges(void *h);
uint64_t srba_engine_num_observations(void *h);
int srba_engine_pose_nvals(void *h);
void srba_engine_get_edge(void *h, uint64_t id, uint64_t *from, uint64_t *to, double *inv_pose);
void srba_engine_set_edge_pose(void *h, uint64_t id, const double *inv_pose);
uint64_t srba_engine_num_unknown_lms(void *h);
void srba_engine_get_unknown_lms(void *h, uint64_t *ids, uint64_t *base, double *pos);
void srba_engine_set_lm_pos(void *h, uint64_t id, const double *pos);
uint64_t srba_engine_get_next_edge(void *h, uint64_t kf, uint64_t *target, uint64_t *next, uint64_t *dist, uint64_t cap);
/* rba_state.all_observations_Jacob_validity (srba_types.h:722): one byte per observation; returns the number of observations. */
uint64_t srba_engine_get_jacob_validity(void *h, uint8_t *out, uint64_t cap);
/* RbaEngine::eval_overall_squared_error() (RbaEngine.h:381, impl/eval_overall_error.h:15-137): BFS trees on the host, all
 * numerics on the device through srba_b200_eval_overall_squared_error.  0 ok, -1 error (srba_engine_last_error). */
int srba_engine_eval_overall_squared_error(void *h, double *out_sqerr);
/* RbaEngine::create_complete_spanning_tree(root, span_tree, max_depth) (RbaEngine.h:343-348, impl/spantree_create_complete.h):
 * ids ascending + pose of each as seen from root (S doubles each); returns the number of keyframes reached (call with NULL
 * arrays for the size), -1 on error. */
int64_t srba_engine_complete_spanning_tree(void *h, uint64_t root, uint64_t max_depth, uint64_t *ids, double *poses, uint64_t cap);
uint64_t srba_engine_get_num(void *h, uint64_t kf, uint64_t *target, double *poses, uint64_t cap);
/* overwrite one numeric spanning-tree entry num[src][dst] (state injection for tests / checkpoint restore) */
void srba_engine_set_num(void *h, uint64_t src, uint64_t dst, const double *pose);
int64_t srba_engine_get_all_edges_path(void *h, uint64_t from, uint64_t to, uint64_t *edge_ids, uint64_t cap);
/* low-level graph construction (tests/spantree_unittest.cpp style) */
uint64_t srba_engine_alloc_keyframe(void *h);
uint64_t srba_engine_create_kf2kf_edge(void *h, uint64_t new_kf_id, uint64_t from, uint64_t to, const double *init_inv_pose);
double srba_engine_flatten_seconds(void *h);  /* accumulated host time spent flattening + scattering */
double srba_engine_gpu_seconds(void *h);      /* accumulated wall time inside srba_b200_optimize */
uint64_t srba_engine_total_trials(void *h);

const char *srba_engine_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
