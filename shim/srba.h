// shim/srba.h -- drop-in for the reference's <srba.h>: the public template API surface of
//     srba::RbaEngine<KF2KF_POSE_TYPE, LM_TYPE, OBS_TYPE, RBA_OPTIONS>            (/root/reference/include/srba/RbaEngine.h:66-816)
// for the four model combinations BASELINE.json names, implemented over the C ABI of this repo (include/srba_b200_engine.h ->
// libsrba_b200.so: host engine + CUDA hot path).  User code written against the reference --
//     my_srba_t rba;  rba.parameters.srba.max_tree_depth = 3;  ...
//     rba.define_new_keyframe(list_obs, new_kf_info, true);        // RbaEngine.h:207-211
//     rba.optimize_local_area(root, win_size, out_info);           // RbaEngine.h:235-241 (incl. observation_indices_to_optimize)
//     rba.get_k2k_edges()[0].inv_pose                              // RbaEngine.h:412
// -- compiles unchanged against this header (tests/shim/*.cpp are the reference's tutorials, examples/cpp/tutorial-srba-*.cpp, with
// only the GUI / random-noise / .dot lines removed).  Link: -L<repo>/srba_b200 -lsrba_b200.
//
// Types: same names and members as the reference (file:line cited per item).  MRPT types that appear in the signatures come from
// shim/mrpt_lite.h unless real MRPT headers were included first (SRBA_B200_USE_REAL_MRPT).  Errors: the reference throws
// std::exception from ASSERT_ macros; this header throws std::runtime_error carrying srba_engine_last_error().
// NOT provided (outside SURVEY.md 8): OpenGL/dot export, config-file loading, the four models BASELINE does not name.
#pragma once
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef SRBA_B200_USE_REAL_MRPT
#include "mrpt_lite.h"
#endif
#include "../include/srba_b200_engine.h"

namespace srba {

typedef uint64_t TKeyFrameID;   // srba_types.h:21
typedef uint64_t TLandmarkID;   // srba_types.h:22
typedef uint64_t topo_dist_t;   // srba_types.h:23
#define SRBA_INVALID_KEYFRAMEID static_cast<srba::TKeyFrameID>(-1)
#define SRBA_INVALID_INDEX static_cast<size_t>(-1)

struct TNewEdgeInfo {  // srba_types.h:178-200
	size_t id;
	bool has_approx_init_val;
	TKeyFrameID loopclosure_observer_kf, loopclosure_base_kf;
};
enum TCovarianceRecoveryPolicy { crpNone = 0, crpLandmarksApprox };  // srba_types.h:30-34

/** Small fixed-size matrix standing in for the Eigen type of observation_noise_constant_matrix<>::parameters_t::lambda */
template <int N>
struct fixed_matrix {
	double m[N * N];
	fixed_matrix() { setZero(); }
	void setZero() { for (double &v : m) v = 0; }
	void setIdentity() { setZero(); for (int i = 0; i < N; i++) m[i * N + i] = 1; }
	double &operator()(int r, int c) { return m[r * N + c]; }
	double operator()(int r, int c) const { return m[r * N + c]; }
};

namespace kf2kf_poses {  // models/kf2kf_poses.h
struct SE2 { static const size_t REL_POSE_DIMS = 3; typedef mrpt::poses::CPose2D pose_t; enum { KIND = SRBA_POSE_SE2, NVALS = 3 }; };
struct SE3 { static const size_t REL_POSE_DIMS = 6; typedef mrpt::poses::CPose3D pose_t; enum { KIND = SRBA_POSE_SE3, NVALS = 12 }; };
inline void pose_from_flat(const double *v, mrpt::poses::CPose2D &p) { p = mrpt::poses::CPose2D(v[0], v[1], v[2]); }
inline void pose_from_flat(const double *v, mrpt::poses::CPose3D &p) { p = mrpt::poses::CPose3D::fromRotationAndTranslation(v, v + 9); }
}  // namespace kf2kf_poses

namespace landmarks {  // models/landmarks.h
struct Euclidean3D { static const size_t LM_DIMS = 3; };
struct RelativePoses2D { static const size_t LM_DIMS = 3; };
}  // namespace landmarks

namespace observations {  // models/observations_*.h
struct TPixelCoordf { float x, y; };
struct Cartesian_3D {  // observations_Cartesian_3D.h:21-36
	static const size_t OBS_DIMS = 3;
	enum { KIND = SRBA_OBS_CARTESIAN3D };
	struct obs_data_t {
		struct { double x, y, z; } pt;
		void getAsArray(double *o) const { o[0] = pt.x; o[1] = pt.y; o[2] = pt.z; }
	};
	struct TObservationParams {};
	static void fill(srba_engine_params_t &, const TObservationParams &) {}
};
struct MonocularCamera {  // observations_MonocularCamera.h:22-44
	static const size_t OBS_DIMS = 2;
	enum { KIND = SRBA_OBS_MONO };
	struct obs_data_t {
		TPixelCoordf px;
		void getAsArray(double *o) const { o[0] = px.x; o[1] = px.y; }
	};
	struct TObservationParams { mrpt::utils::TCamera camera_calib; };
	static void fill(srba_engine_params_t &c, const TObservationParams &p) {
		c.cam_left[0] = p.camera_calib.fx(); c.cam_left[1] = p.camera_calib.fy(); c.cam_left[2] = p.camera_calib.cx(); c.cam_left[3] = p.camera_calib.cy();
	}
};
struct StereoCamera {  // observations_StereoCamera.h:30-60
	static const size_t OBS_DIMS = 4;
	enum { KIND = SRBA_OBS_STEREO };
	struct obs_data_t {
		TPixelCoordf left_px, right_px;
		void getAsArray(double *o) const { o[0] = left_px.x; o[1] = left_px.y; o[2] = right_px.x; o[3] = right_px.y; }
	};
	struct TObservationParams { mrpt::utils::TStereoCamera camera_calib; };
	static void fill(srba_engine_params_t &c, const TObservationParams &p) {
		const mrpt::utils::TCamera &l = p.camera_calib.leftCamera, &r = p.camera_calib.rightCamera;
		c.cam_left[0] = l.fx(); c.cam_left[1] = l.fy(); c.cam_left[2] = l.cx(); c.cam_left[3] = l.cy();
		c.cam_right[0] = r.fx(); c.cam_right[1] = r.fy(); c.cam_right[2] = r.cx(); c.cam_right[3] = r.cy();
		for (int i = 0; i < 7; i++) c.right_cam_pose[i] = p.camera_calib.rightCameraPose.v[i];
	}
};
struct RelativePoses_2D {  // observations_RelativePoses_2D.h:21-35
	static const size_t OBS_DIMS = 3;
	enum { KIND = SRBA_OBS_RELPOSES2D };
	struct obs_data_t {
		double x, y, yaw;
		void getAsArray(double *o) const { o[0] = x; o[1] = y; o[2] = yaw; }
	};
	struct TObservationParams {};
	static void fill(srba_engine_params_t &, const TObservationParams &) {}
};
}  // namespace observations

namespace options {
struct sensor_pose_on_robot_none {  // srba_options_sensor_pose.h:28-60
	struct parameters_t {};
	static void fill(srba_engine_params_t &c, const parameters_t &) { c.sensor_pose_kind = SRBA_SENSORPOSE_NONE; }
};
struct sensor_pose_on_robot_se3 {  // srba_options_sensor_pose.h:64-115
	struct parameters_t { mrpt::poses::CPose3D relative_pose; };
	static void fill(srba_engine_params_t &c, const parameters_t &p) {
		c.sensor_pose_kind = SRBA_SENSORPOSE_SE3;
		c.sensor_pose[0] = p.relative_pose.x(); c.sensor_pose[1] = p.relative_pose.y(); c.sensor_pose[2] = p.relative_pose.z();
		p.relative_pose.getYawPitchRoll(c.sensor_pose[3], c.sensor_pose[4], c.sensor_pose[5]);
	}
};
struct observation_noise_identity {  // srba_options_noise.h:27-63
	struct parameters_t { double std_noise_observations; parameters_t() : std_noise_observations(1.) {} };
	static void fill(srba_engine_params_t &c, const parameters_t &p) { c.noise_kind = SRBA_NOISE_IDENTITY; c.noise_std = p.std_noise_observations; }
};
template <class OBS_T>
struct observation_noise_constant_matrix {  // srba_options_noise.h:70-115
	struct parameters_t { fixed_matrix<OBS_T::OBS_DIMS> lambda; parameters_t() { lambda.setIdentity(); } };
	static void fill(srba_engine_params_t &c, const parameters_t &p) {
		c.noise_kind = SRBA_NOISE_CONST_MATRIX;
		const int o = (int)OBS_T::OBS_DIMS;
		for (int i = 0; i < 16; i++) c.noise_lambda[i] = 0;
		for (int i = 0; i < o; i++)
			for (int j = 0; j < o; j++) c.noise_lambda[i * o + j] = p.lambda(i, j);
	}
};
struct solver_extra_results_t { void clear() {} };
struct solver_LM_schur_dense_cholesky { enum { KIND = SRBA_SOLVER_SCHUR_DENSE_CHOLESKY }; typedef solver_extra_results_t extra_results_t; };     // srba_options_solver.h:21-37
struct solver_LM_schur_sparse_cholesky { enum { KIND = SRBA_SOLVER_SCHUR_SPARSE_CHOLESKY }; typedef solver_extra_results_t extra_results_t; };   // :41-57
struct solver_LM_no_schur_sparse_cholesky { enum { KIND = SRBA_SOLVER_NO_SCHUR_SPARSE_CHOLESKY }; typedef solver_extra_results_t extra_results_t; };  // :61-77
}  // namespace options

namespace ecps {
struct local_areas_fixed_size {  // ecps/local_areas_fixed_size.h:22-50
	struct parameters_t { size_t submap_size, min_obs_to_loop_closure; parameters_t() : submap_size(15), min_obs_to_loop_closure(4) {} };
};
}  // namespace ecps

struct RBA_OPTIONS_DEFAULT {  // RbaEngine.h:39-45
	typedef ecps::local_areas_fixed_size edge_creation_policy_t;
	typedef options::sensor_pose_on_robot_none sensor_pose_on_robot_t;
	typedef options::observation_noise_identity obs_noise_matrix_t;
	typedef options::solver_LM_schur_dense_cholesky solver_t;
};

template <class KF2KF_POSE_TYPE, class LM_TYPE, class OBS_TYPE, class RBA_OPTIONS = RBA_OPTIONS_DEFAULT>
class RbaEngine {
public:
	typedef KF2KF_POSE_TYPE kf2kf_pose_t;
	typedef LM_TYPE landmark_t;
	typedef OBS_TYPE obs_t;
	typedef RBA_OPTIONS rba_options_t;
	static const size_t REL_POSE_DIMS = kf2kf_pose_t::REL_POSE_DIMS;
	static const size_t LM_DIMS = landmark_t::LM_DIMS;
	static const size_t OBS_DIMS = obs_t::OBS_DIMS;
	typedef typename kf2kf_pose_t::pose_t pose_t;

	struct k2k_edge_t {  // srba_types.h:82-98
		TKeyFrameID from, to;
		pose_t inv_pose;  //!< pose of "from" as seen from "to"
		size_t id;
	};
	typedef std::deque<k2k_edge_t> k2k_edges_deque_t;
	struct TRelativeLandmarkPos {  // srba_types.h:103-121
		TKeyFrameID id_frame_base;
		double pos[3];
	};
	typedef std::map<TLandmarkID, TRelativeLandmarkPos> TRelativeLandmarkPosMap;
	struct pose_flag_t { pose_t pose; bool updated; };
	typedef std::map<TKeyFrameID, pose_flag_t> frameid2pose_map_t;

	struct observation_t {  // srba_types.h:310-316
		TLandmarkID feat_id;
		typename obs_t::obs_data_t obs_data;
	};
	struct new_kf_observation_t {  // srba_types.h:326-360
		new_kf_observation_t() : is_fixed(false), is_unknown_with_init_val(false) { feat_rel_pos[0] = feat_rel_pos[1] = feat_rel_pos[2] = 0; }
		observation_t obs;
		bool is_fixed, is_unknown_with_init_val;
		double feat_rel_pos[3];
		void setRelPos(double x, double y, double z) { feat_rel_pos[0] = x; feat_rel_pos[1] = y; feat_rel_pos[2] = z; }
	};
	typedef std::deque<new_kf_observation_t> new_kf_observations_t;

	struct TOptimizeExtraOutputInfo {  // RbaEngine.h:125-177
		TOptimizeExtraOutputInfo() { clear(); }
		size_t num_observations, num_jacobians, num_kf2kf_edges_optimized, num_kf2lm_edges_optimized, num_total_scalar_optimized, num_kf_optimized, num_lm_optimized,
		    num_span_tree_numeric_updates;
		double obs_rmse, total_sqr_error_init, total_sqr_error_final, HAp_condition_number;
		std::vector<size_t> optimized_k2k_edge_indices, optimized_landmark_indices;
		typename RBA_OPTIONS::solver_t::extra_results_t extra_results;
		void clear() {
			num_observations = num_jacobians = num_kf2kf_edges_optimized = num_kf2lm_edges_optimized = num_total_scalar_optimized = num_kf_optimized = num_lm_optimized =
			    num_span_tree_numeric_updates = 0;
			obs_rmse = total_sqr_error_init = total_sqr_error_final = HAp_condition_number = 0;
			optimized_k2k_edge_indices.clear(); optimized_landmark_indices.clear(); extra_results.clear();
		}
	};
	struct TNewKeyFrameInfo {  // RbaEngine.h:180-196
		TKeyFrameID kf_id;
		std::vector<TNewEdgeInfo> created_edge_ids;
		TOptimizeExtraOutputInfo optimize_results, optimize_results_stg1;
		void clear() { kf_id = static_cast<TKeyFrameID>(-1); created_edge_ids.clear(); optimize_results.clear(); }
	};
	struct TOptimizeLocalAreaParams {  // RbaEngine.h:214-228
		bool optimize_k2k_edges, optimize_landmarks;
		TKeyFrameID max_visitable_kf_id;
		size_t dont_optimize_landmarks_seen_less_than_n_times;
		TOptimizeLocalAreaParams() : optimize_k2k_edges(true), optimize_landmarks(true), max_visitable_kf_id(static_cast<TKeyFrameID>(-1)), dont_optimize_landmarks_seen_less_than_n_times(2) {}
	};
	struct TSRBAParameters {  // RbaEngine.h:417-460, defaults impl/rba_problem_common.h:35-56
		size_t max_tree_depth, max_optimize_depth;
		bool optimize_new_edges_alone, use_robust_kernel, use_robust_kernel_stage1;
		double kernel_param;
		size_t max_iters;
		double max_error_per_obs_to_stop, max_rho, max_lambda, min_error_reduction_ratio_to_relinearize;
		bool numeric_jacobians;
		void (*feedback_user_iteration)(unsigned int iter, const double total_sq_err, const double mean_sqroot_error);
		bool compute_condition_number, compute_sparsity_stats;
		double max_rmse_show_red_warning;
		TCovarianceRecoveryPolicy cov_recovery;
		TSRBAParameters() {
			srba_engine_params_t d;
			srba_engine_default_params(&d);
			max_tree_depth = d.max_tree_depth; max_optimize_depth = d.max_optimize_depth;
			optimize_new_edges_alone = d.optimize_new_edges_alone != 0; use_robust_kernel = d.use_robust_kernel != 0; use_robust_kernel_stage1 = d.use_robust_kernel_stage1 != 0;
			kernel_param = d.kernel_param; max_iters = d.max_iters; max_error_per_obs_to_stop = d.max_error_per_obs_to_stop; max_rho = d.max_rho; max_lambda = d.max_lambda;
			min_error_reduction_ratio_to_relinearize = d.min_error_reduction_ratio_to_relinearize;
			numeric_jacobians = false; feedback_user_iteration = NULL; compute_condition_number = compute_sparsity_stats = false; max_rmse_show_red_warning = 0.5;
			cov_recovery = crpLandmarksApprox;
		}
	};
	struct TAllParameters {  // RbaEngine.h:463-472
		TSRBAParameters srba;
		typename obs_t::TObservationParams sensor;
		typename RBA_OPTIONS::sensor_pose_on_robot_t::parameters_t sensor_pose;
		typename RBA_OPTIONS::obs_noise_matrix_t::parameters_t obs_noise;
		typename RBA_OPTIONS::edge_creation_policy_t::parameters_t ecp;
	};
	TAllParameters parameters;

	/** A read-only view of what tutorials read from get_rba_state() (srba_types.h:700-760) */
	struct rba_problem_state_t {
		k2k_edges_deque_t k2k_edges;
		TRelativeLandmarkPosMap unknown_lms;
		size_t num_keyframes, num_observations;
	};

	RbaEngine() : m_h(NULL) {}
	~RbaEngine() { if (m_h) srba_engine_destroy(m_h); }
	RbaEngine(const RbaEngine &) = delete;
	RbaEngine &operator=(const RbaEngine &) = delete;

	void setVerbosityLevel(int) {}
	void enable_time_profiler(bool = true) {}

	// ---- main API (RbaEngine.h:198-260)
	void define_new_keyframe(const new_kf_observations_t &obs, TNewKeyFrameInfo &out_new_kf_info, const bool run_local_optimization = true) {
		sync_params();
		const size_t n = obs.size();
		std::vector<uint64_t> ids(n);
		std::vector<double> data(n * OBS_DIMS), rel(n * 3);
		std::vector<uint8_t> fixed(n), init(n);
		for (size_t i = 0; i < n; i++) {
			ids[i] = obs[i].obs.feat_id;
			obs[i].obs.obs_data.getAsArray(&data[i * OBS_DIMS]);
			fixed[i] = obs[i].is_fixed; init[i] = obs[i].is_unknown_with_init_val;
			for (int k = 0; k < 3; k++) rel[3 * i + k] = obs[i].feat_rel_pos[k];
		}
		std::vector<srba_new_edge_t> edges(n + 4);
		uint64_t kf = 0, ne = 0;
		srba_opt_info_t o, o1;
		std::memset(&o, 0, sizeof(o)); std::memset(&o1, 0, sizeof(o1));
		if (srba_engine_define_new_keyframe(m_h, n, ids.data(), data.data(), fixed.data(), init.data(), rel.data(), run_local_optimization ? 1 : 0, &kf, edges.data(),
		                                    edges.size(), &ne, &o, &o1) != 0)
			throw std::runtime_error(srba_engine_last_error());
		out_new_kf_info.clear();
		out_new_kf_info.kf_id = kf;
		for (uint64_t i = 0; i < ne && i < edges.size(); i++) {
			TNewEdgeInfo e = {(size_t)edges[i].id, edges[i].has_approx_init_val != 0, edges[i].loopclosure_observer_kf, edges[i].loopclosure_base_kf};
			out_new_kf_info.created_edge_ids.push_back(e);
		}
		to_info(o, out_new_kf_info.optimize_results);
		to_info(o1, out_new_kf_info.optimize_results_stg1);
		m_state_dirty = true;
	}

	void optimize_local_area(const TKeyFrameID root_id, const unsigned int win_size, TOptimizeExtraOutputInfo &out_info,
	                         const TOptimizeLocalAreaParams &params = TOptimizeLocalAreaParams(),
	                         const std::vector<size_t> &observation_indices_to_optimize = std::vector<size_t>()) {
		sync_params();
		srba_opt_info_t o;
		std::memset(&o, 0, sizeof(o));
		int rc;
		const bool feedback = parameters.srba.feedback_user_iteration != NULL;
		if (feedback) srba_engine_trace(m_h, 8 * (int)parameters.srba.max_iters + 16);
		if (!observation_indices_to_optimize.empty()) {
			std::vector<uint64_t> idx(observation_indices_to_optimize.begin(), observation_indices_to_optimize.end());
			rc = srba_engine_optimize_local_area_obs(m_h, root_id, win_size, idx.size(), idx.data(), &o);
		} else
			rc = srba_engine_optimize_local_area(m_h, root_id, win_size, params.optimize_k2k_edges, params.optimize_landmarks, params.max_visitable_kf_id,
			                                     params.dont_optimize_landmarks_seen_less_than_n_times, &o);
		if (feedback) srba_engine_trace(m_h, 0);
		if (rc != 0) throw std::runtime_error(srba_engine_last_error());
		to_info(o, out_info);
		if (feedback && o.num_observations > 0) replay_feedback(o);
		{  // optimized index lists (RbaEngine.h:146-147): the unknown selection of this very call
			uint64_t nk = 0, nl = 0;
			srba_engine_select_local_area(m_h, root_id, win_size, NULL, 0, &nk, NULL, 0, &nl);
			std::vector<uint64_t> k(nk), l(nl);
			srba_engine_select_local_area(m_h, root_id, win_size, k.data(), nk, &nk, l.data(), nl, &nl);
			if (params.optimize_k2k_edges) out_info.optimized_k2k_edge_indices.assign(k.begin(), k.end());
			if (params.optimize_landmarks) out_info.optimized_landmark_indices.assign(l.begin(), l.end());
		}
		m_state_dirty = true;
	}

private:
	/** parameters.srba.feedback_user_iteration replayed AFTER the call from the per-trial trace of the device loop: (0, initial error, RMSE) as
	  * optimize_edges.h:417-418, then (iter, new error, new RMSE) for every accepted step with the value `iter` had (:595-596).
	  * (define_new_keyframe runs up to three optimize_edges calls internally and does not replay them.) */
	void replay_feedback(const srba_opt_info_t &o) {
		const uint64_t n = srba_engine_last_trace(m_h, NULL, NULL, 0, NULL);
		std::vector<double> tr(6 * n + 6);
		srba_engine_last_trace(m_h, tr.data(), NULL, n, NULL);
		const double nobs = (double)o.num_observations;
		(*parameters.srba.feedback_user_iteration)(0u, o.total_sqr_error_init, std::sqrt(o.total_sqr_error_init / nobs));
		unsigned iter = 0;
		for (uint64_t t = 0; t < n; t++)
			if (tr[6 * t + 1] != 0 && tr[6 * t + 4] != 0) {
				(*parameters.srba.feedback_user_iteration)(iter, tr[6 * t + 3], std::sqrt(tr[6 * t + 3] / nobs));
				iter++;
			}
	}

public:
	double eval_overall_squared_error() const {  // RbaEngine.h:381
		const_cast<RbaEngine *>(this)->sync_params();
		double e = 0;
		if (srba_engine_eval_overall_squared_error(m_h, &e) != 0) throw std::runtime_error(srba_engine_last_error());
		return e;
	}

	void create_complete_spanning_tree(const TKeyFrameID root_id, frameid2pose_map_t &span_tree, const size_t max_depth = std::numeric_limits<size_t>::max()) const {  // RbaEngine.h:343-348
		const_cast<RbaEngine *>(this)->sync_params();
		const int64_t n = srba_engine_complete_spanning_tree(m_h, root_id, max_depth, NULL, NULL, 0);
		if (n < 0) throw std::runtime_error(srba_engine_last_error());
		std::vector<uint64_t> ids(n);
		std::vector<double> poses((size_t)n * kf2kf_pose_t::NVALS);
		srba_engine_complete_spanning_tree(m_h, root_id, max_depth, ids.data(), poses.data(), n);
		span_tree.clear();
		for (int64_t i = 0; i < n; i++) {
			pose_flag_t &f = span_tree[ids[i]];
			kf2kf_poses::pose_from_flat(&poses[(size_t)i * kf2kf_pose_t::NVALS], f.pose);
			f.updated = true;
		}
	}

	const k2k_edges_deque_t &get_k2k_edges() const { refresh(); return m_state.k2k_edges; }                 // RbaEngine.h:412
	const TRelativeLandmarkPosMap &get_unknown_feats() const { refresh(); return m_state.unknown_lms; }     // RbaEngine.h:414
	const rba_problem_state_t &get_rba_state() const { refresh(); return m_state; }                         // RbaEngine.h:409

private:
	void *m_h;
	mutable rba_problem_state_t m_state;
	mutable bool m_state_dirty = true;

	static void to_info(const srba_opt_info_t &o, TOptimizeExtraOutputInfo &out) {
		out.clear();
		out.num_observations = o.num_observations; out.num_jacobians = o.num_jacobians; out.num_kf2kf_edges_optimized = o.num_kf2kf_edges_optimized;
		out.num_kf2lm_edges_optimized = o.num_kf2lm_edges_optimized; out.num_total_scalar_optimized = o.num_total_scalar_optimized;
		out.num_kf_optimized = o.num_kf_optimized; out.num_lm_optimized = o.num_lm_optimized; out.num_span_tree_numeric_updates = o.num_span_tree_numeric_updates;
		out.obs_rmse = o.obs_rmse; out.total_sqr_error_init = o.total_sqr_error_init; out.total_sqr_error_final = o.total_sqr_error_final;
	}
	void to_c_params(srba_engine_params_t &c) const {
		srba_engine_default_params(&c);
		c.pose_kind = kf2kf_pose_t::KIND;
		c.obs_kind = obs_t::KIND;
		obs_t::fill(c, parameters.sensor);
		RBA_OPTIONS::sensor_pose_on_robot_t::fill(c, parameters.sensor_pose);
		RBA_OPTIONS::obs_noise_matrix_t::fill(c, parameters.obs_noise);
		const TSRBAParameters &s = parameters.srba;
		c.max_tree_depth = s.max_tree_depth; c.max_optimize_depth = s.max_optimize_depth;
		c.optimize_new_edges_alone = s.optimize_new_edges_alone; c.use_robust_kernel = s.use_robust_kernel; c.use_robust_kernel_stage1 = s.use_robust_kernel_stage1;
		c.kernel_param = s.kernel_param; c.max_iters = s.max_iters; c.max_error_per_obs_to_stop = s.max_error_per_obs_to_stop; c.max_rho = s.max_rho;
		c.max_lambda = s.max_lambda; c.min_error_reduction_ratio_to_relinearize = s.min_error_reduction_ratio_to_relinearize;
		c.cov_recovery = (int)s.cov_recovery;
		c.submap_size = parameters.ecp.submap_size; c.min_obs_to_loop_closure = parameters.ecp.min_obs_to_loop_closure;
		c.solver_kind = RBA_OPTIONS::solver_t::KIND;
	}
	/** `parameters` is a public, freely writable member in the reference: forward its current value before every call */
	void sync_params() {
		srba_engine_params_t c;
		to_c_params(c);
		if (!m_h) {
			m_h = srba_engine_create(&c);
			if (!m_h) throw std::runtime_error(srba_engine_last_error());
		} else if (srba_engine_set_params(m_h, &c) != 0)
			throw std::runtime_error(srba_engine_last_error());
	}
	void refresh() const {
		if (!m_h) { m_state.k2k_edges.clear(); m_state.unknown_lms.clear(); m_state.num_keyframes = m_state.num_observations = 0; return; }
		if (!m_state_dirty) return;
		const uint64_t ne = srba_engine_num_edges(m_h);
		m_state.k2k_edges.resize(ne);
		double v[12];
		for (uint64_t i = 0; i < ne; i++) {
			uint64_t from, to;
			srba_engine_get_edge(m_h, i, &from, &to, v);
			k2k_edge_t &e = m_state.k2k_edges[i];
			e.from = from; e.to = to; e.id = i;
			kf2kf_poses::pose_from_flat(v, e.inv_pose);
		}
		const uint64_t nl = srba_engine_num_unknown_lms(m_h);
		std::vector<uint64_t> ids(nl), base(nl);
		std::vector<double> pos(3 * nl);
		if (nl) srba_engine_get_unknown_lms(m_h, ids.data(), base.data(), pos.data());
		m_state.unknown_lms.clear();
		for (uint64_t i = 0; i < nl; i++) {
			TRelativeLandmarkPos &p = m_state.unknown_lms[ids[i]];
			p.id_frame_base = base[i];
			for (int k = 0; k < 3; k++) p.pos[k] = pos[3 * i + k];
		}
		m_state.num_keyframes = srba_engine_num_keyframes(m_h);
		m_state.num_observations = srba_engine_num_observations(m_h);
		m_state_dirty = false;
	}
};

}  // namespace srba
