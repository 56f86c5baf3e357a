/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Plain-C view of the CPU restatement of MRPT/srba's RbaEngine<> (include/srba/RbaEngine.h:66-816)
 * so that tests/ and bench.py's cpu_baseline leg can drive it through ctypes.
 * The product (srba_b200/) must never include or link this. */
#ifndef SRBA_ORACLE_CAPI_H
#define SRBA_ORACLE_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Model tags: the 4 model combinations named by BASELINE.json configs. */
enum { ORC_POSE_SE2 = 0, ORC_POSE_SE3 = 1 };
enum { ORC_OBS_RELPOSES2D = 0, ORC_OBS_CARTESIAN3D = 1, ORC_OBS_MONO = 2, ORC_OBS_STEREO = 3 };
enum { ORC_SENSORPOSE_NONE = 0, ORC_SENSORPOSE_SE3 = 1 };
enum { ORC_NOISE_IDENTITY = 0, ORC_NOISE_CONST_MATRIX = 1 };

/* Compile-time options (RBA_OPTIONS, RbaEngine.h:39-45) + run-time parameters
 * (TSRBAParameters RbaEngine.h:424-460, defaults impl/rba_problem_common.h:35-56;
 *  ecp params ecps/local_areas_fixed_size.h:27-37). */
typedef struct {
	int32_t pose_kind, obs_kind, sensor_pose_kind, noise_kind;
	double sensor_pose[6];      /* x,y,z,yaw,pitch,roll  (sensor_pose_on_robot_se3::parameters_t) */
	double noise_std;           /* observation_noise_identity::std_noise_observations */
	double noise_lambda[16];    /* observation_noise_constant_matrix::lambda, row-major o x o */
	double cam_left[4];         /* fx,fy,cx,cy */
	double cam_right[4];
	double right_cam_pose[7];   /* x,y,z,qr,qx,qy,qz (TStereoCamera::rightCameraPose) */
	uint64_t max_tree_depth, max_optimize_depth;
	int32_t optimize_new_edges_alone, use_robust_kernel, use_robust_kernel_stage1, cov_recovery;
	double kernel_param;
	uint64_t max_iters;
	double max_error_per_obs_to_stop, max_rho, max_lambda, min_error_reduction_ratio_to_relinearize;
	uint64_t submap_size, min_obs_to_loop_closure;
	int32_t solver_kind;        /* RBA_OPTIONS::solver_t (srba_options_solver.h:21-77): 0 Schur + dense Cholesky (default, RbaEngine.h:44),
	                             * 1 Schur + sparse Cholesky, 2 no Schur + sparse Cholesky */
	int32_t reserved_;
} orc_params_t;

/* TOptimizeExtraOutputInfo (RbaEngine.h:125-177), scalar part. */
typedef struct {
	uint64_t num_observations, num_jacobians, num_kf2kf_edges_optimized, num_kf2lm_edges_optimized,
	    num_total_scalar_optimized, num_kf_optimized, num_lm_optimized, num_span_tree_numeric_updates;
	double obs_rmse, total_sqr_error_init, total_sqr_error_final;
	uint64_t num_lm_trials;   /* NOT in the reference: number of solve(lambda) trials (BASELINE metric) */
	uint64_t num_iters;       /* NOT in the reference: final value of `iter` (optimize_edges.h:452) */
} orc_opt_info_t;

/* TNewEdgeInfo (srba_types.h:178-200) */
typedef struct {
	uint64_t id;
	int32_t has_approx_init_val;
	uint64_t loopclosure_observer_kf, loopclosure_base_kf;
} orc_new_edge_t;

void orc_default_params(orc_params_t *p);

void *orc_create(const orc_params_t *p);
void orc_destroy(void *h);
void orc_set_params(void *h, const orc_params_t *p); /* run-time parameters only */

/* define_new_keyframe (impl/define_new_keyframe.h:16-116).  obs_data is n x OBS_DIMS row-major,
 * feat_rel_pos is n x 3.  Returns 0 or <0 on an ASSERT-style failure (message via orc_last_error). */
int orc_define_new_keyframe(void *h, uint64_t n_obs, const uint64_t *feat_id, const double *obs_data,
                            const uint8_t *is_fixed, const uint8_t *is_unknown_with_init_val,
                            const double *feat_rel_pos, int run_local_optimization, uint64_t *out_kf_id,
                            orc_new_edge_t *out_edges, uint64_t max_edges, uint64_t *out_n_edges,
                            orc_opt_info_t *out_opt, orc_opt_info_t *out_opt_stg1);

/* optimize_local_area (impl/optimize_local_area.h:15-58) with TOptimizeLocalAreaParams. */
int orc_optimize_local_area(void *h, uint64_t root, uint64_t win_size, int optimize_k2k_edges,
                            int optimize_landmarks, uint64_t max_visitable_kf_id,
                            uint64_t dont_optimize_landmarks_seen_less_than_n_times, orc_opt_info_t *out);
/* protected optimize_edges (impl/optimize_edges.h:44-793) */
/* optimize_local_area(root, win_size, out, default params, observation_indices_to_optimize)  (RbaEngine.h:235-241) */
int orc_optimize_local_area_obs(void *h, uint64_t root, uint64_t win_size, uint64_t n_obs_idx, const uint64_t *obs_idx, orc_opt_info_t *out);
void orc_set_num(void *h, uint64_t src, uint64_t dst, const double *pose); /* spanning_tree.num[src][dst].pose (test support) */
int orc_optimize_edges(void *h, uint64_t n_k2k, const uint64_t *k2k, uint64_t n_lm, const uint64_t *lms,
                       orc_opt_info_t *out);

/* Low-level graph access used by the reference's own unit tests (tests/spantree_unittest.cpp). */
uint64_t orc_alloc_keyframe(void *h);
uint64_t orc_create_kf2kf_edge(void *h, uint64_t new_kf_id, uint64_t from, uint64_t to, const double *init_inv_pose);
void orc_update_numeric_all(void *h);
int orc_are_keyframes_connected(void *h, uint64_t a, uint64_t b);
/* create_complete_spanning_tree (impl/spantree_create_complete.h:18-127): returns count; ids/poses may be NULL */
uint64_t orc_complete_spanning_tree(void *h, uint64_t root, uint64_t max_depth, uint64_t *ids, double *poses, uint64_t cap);
/* all_observations_Jacob_validity (srba_types.h:722), one byte per observation; returns the number of observations. */
uint64_t orc_get_jacob_validity(void *h, uint8_t *out, uint64_t cap);
/* RbaEngine::eval_overall_squared_error (impl/eval_overall_error.h:15-137): every observation, poses from complete BFS trees. */
double orc_eval_overall_squared_error(void *h);
/* extra_results.hessian of the last optimize_edges (optimize_edges.h:769): row-major n x n, upper triangle valid; returns n*n. */
uint64_t orc_last_dense_hessian(void *h, double *out, uint64_t cap, int *valid);
int64_t orc_find_path_bfs_len(void *h, uint64_t from, uint64_t to);

/* State getters.  Poses are flattened: SE2 = (x,y,phi); SE3 = R row-major (9) + t (3). */
uint64_t orc_num_keyframes(void *h);
uint64_t orc_num_edges(void *h);
uint64_t orc_num_observations(void *h);
int orc_pose_nvals(void *h);
void orc_get_edge(void *h, uint64_t id, uint64_t *from, uint64_t *to, double *inv_pose);
void orc_set_edge_pose(void *h, uint64_t id, const double *inv_pose);
uint64_t orc_num_unknown_lms(void *h);
void orc_get_unknown_lms(void *h, uint64_t *ids, uint64_t *base, double *pos);
uint64_t orc_num_known_lms(void *h);
void orc_get_known_lms(void *h, uint64_t *ids, uint64_t *base, double *pos);
void orc_set_lm_pos(void *h, uint64_t id, const double *pos);
/* sym.next_edge[kf]: returns count; fills (target,next,distance) */
uint64_t orc_get_next_edge(void *h, uint64_t kf, uint64_t *target, uint64_t *next, uint64_t *dist, uint64_t cap);
/* num[kf]: returns count; fills target ids, poses, updated flags */
uint64_t orc_get_num(void *h, uint64_t kf, uint64_t *target, double *poses, uint8_t *updated, uint64_t cap);
/* sym.all_edges[from][to] as edge ids: returns path length or -1 */
int64_t orc_get_all_edges_path(void *h, uint64_t from, uint64_t to, uint64_t *edge_ids, uint64_t cap);

/* Trace of optimize_edges calls (parity debugging; mirrors what feedback_user_iteration would see).
 * Tracing is off by default; each traced call stores the ordered unknown ids, the involved observation
 * indices (WITH the duplicates of optimize_edges.h:177-193), and per solve(lambda) trial:
 * [lambda, solved_ok, rho, new_total_sq_err, accepted, relinearized] followed by delta_eps. */
void orc_trace_enable(void *h, int on);
void orc_trace_clear(void *h);
uint64_t orc_trace_num_calls(void *h);
void orc_trace_call_sizes(void *h, uint64_t call, uint64_t *n_k2k, uint64_t *n_lm, uint64_t *n_obs,
                          uint64_t *n_trials, uint64_t *n_scalars);
void orc_trace_call_data(void *h, uint64_t call, uint64_t *k2k_ids, uint64_t *lm_ids, uint64_t *obs_idx,
                         double *trial_scalars /* n_trials x 6 */, double *deltas /* n_trials x n_scalars */,
                         double *init_grad /* n_scalars */, double *lambda0);
double orc_total_opt_seconds(void *h);   /* accumulated wall time inside optimize_edges ("opt" section) */
uint64_t orc_total_opt_trials(void *h);
void orc_reset_opt_timer(void *h);

/* Unit-level entry: restatement of tests/schur_unittest.cpp:71-279.  JA: per (kf>=1, lm) visible block 3x6,
 * JF: per visible (kf, lm) block 3x3, both in the test's generation order.  Outputs the reduced dense HAp
 * (6np x 6np, symmetric-filled) and the reduced gradient head. */
int orc_schur_case(uint64_t n_p, uint64_t n_f, const uint8_t *visible, const double *JA, const double *JF,
                   double lambda, double *out_HAp_dense, double *out_grad_Ap, uint64_t *out_n_invertible);

/* Unit-level entry for the three solver_engine<> variants (impl/lev-marq_solvers.h:27-209,213-405,409-591): the system of
 * orc_schur_case with minus_grad = ones, one solve(lambda).  out_delta: [6 n_p | 3 n_f]; out_minus_grad_after: the caller's
 * gradient vector as solve() leaves it (mutated in place by the Schur variants, untouched without Schur). */
int orc_solver_case(int solver_kind, uint64_t n_p, uint64_t n_f, const uint8_t *visible, const double *JA, const double *JF, double lambda,
                    double *out_delta, double *out_minus_grad_after, int *out_solved);

/* Unit-level entries for Jacobian checks (impl/jacobians.h:207-354, 888-1013 and the numeric recipe :90-186).
 * A may be NULL (identity). edge_inv_pose only read when normal_dir==0. */
int orc_eval_dh_dAp(const orc_params_t *p, const double *A, const double *D, const double *edge_inv_pose,
                    int normal_dir, const double *lm_pos, double *out_J /* o x p row-major */);
int orc_eval_dh_df(const orc_params_t *p, const double *pose_base_wrt_obs, const double *lm_pos, double *out_J);
/* h(.) through the reference's composition: returns z_zero - observe_error i.e. the prediction h; used for FD */
int orc_predict(const orc_params_t *p, const double *pose_base_wrt_obs, const double *lm_pos, const double *z, double *out_err);
void orc_pose_compose(int pose_kind, const double *a, const double *b, double *out);
void orc_pose_inverse(int pose_kind, const double *a, double *out);
void orc_pose_exp(int pose_kind, const double *v, double *out);
void orc_pose_from_xyzypr(const double *v6, double *out12);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
