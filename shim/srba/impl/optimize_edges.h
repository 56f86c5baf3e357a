// shim/srba/impl/optimize_edges.h -- REPLACEMENT for the reference's include/srba/impl/optimize_edges.h (:44-793): the member template
//     RbaEngine<KF2KF_POSE_TYPE,LM_TYPE,OBS_TYPE,RBA_OPTIONS>::optimize_edges(run_k2k_edges, run_feat_ids, out_info, obs_indices)
// declared at RbaEngine.h:533-538 and pulled in by the quoted include at RbaEngine.h:824-851.  Put this directory before the reference's on the
// include path (the other headers stay the reference's own) and link libsrba_b200.so: every local optimisation then runs on the GPU.
//   steps A-D, I, O (unknown filtering, observation list, tree roots, Hessian / Schur symbolic)  -> srba_b200_shim::flatten over rba_state
//   steps E-S numeric                                                                            -> srba_b200_optimize()
//   write-back (optimize_edges.h:508-539,550-565,722-751,778-781)                                -> srba_b200_shim::scatter + out_info
// Supported model combinations (BASELINE.json): SE2 + RelativePoses2D/RelativePoses_2D; SE3 + Euclidean3D + Cartesian_3D | MonocularCamera |
// StereoCamera; sensor_pose_on_robot_none / _se3; observation_noise_identity / _constant_matrix; the three solver_t options.
// MRPT (>= 1.3) types are reached only through the few accessors below (CPose2D::x/y/phi, CPose3D::getRotationMatrix/x/y/z/
// setFromValues..., TCamera::fx/fy/cx/cy, CPose3DQuat::operator[]).  NOT compiled against real MRPT in this repository (absent from the
// image): tests/shim/reference_stub/ holds declarations of exactly the reference / MRPT names used here and tests/test_shim.py type-checks
// an explicit instantiation against them; the flattening / scatter algorithm itself (srba_b200_shim.hpp) runs in the tests on the oracle's
// restatement of the same data model.
#pragma once
#include <stdexcept>

#include "../../srba_b200_shim.hpp"

namespace srba {
namespace b200 {

inline void pose_to_flat(const mrpt::poses::CPose2D &p, double *o) { o[0] = p.x(); o[1] = p.y(); o[2] = p.phi(); }
inline void pose_from_flat(const double *o, mrpt::poses::CPose2D &p) { p.x(o[0]); p.y(o[1]); p.phi(o[2]); }
inline void pose_to_flat(const mrpt::poses::CPose3D &p, double *o) {  // R row-major + t
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) o[3 * r + c] = p.getRotationMatrix()(r, c);
	o[9] = p.x(); o[10] = p.y(); o[11] = p.z();
}
inline void pose_from_flat(const double *o, mrpt::poses::CPose3D &p) {
	mrpt::math::CMatrixDouble33 R;
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) R(r, c) = o[3 * r + c];
	p.setRotationMatrix(R);
	p.x(o[9]); p.y(o[10]); p.z(o[11]);
}
template <class POSE> struct pose_nvals;
template <> struct pose_nvals<mrpt::poses::CPose2D> { enum { value = 3 }; };
template <> struct pose_nvals<mrpt::poses::CPose3D> { enum { value = 12 }; };

/** ACCESS policy of srba_b200_shim::flatten / scatter over the reference's TRBA_Problem_state (srba_types.h:548-785) */
template <class RBA>
struct ReferenceAccess {
	typedef typename RBA::rba_problem_state_t ST;
	typedef typename RBA::pose_t pose_t;
	enum { POSE_NVALS = pose_nvals<pose_t>::value, OBS_DIMS = RBA::OBS_DIMS };
	typedef typename RBA::TSparseBlocksJacobians_dh_dAp::col_t colAp_t;
	typedef typename RBA::TSparseBlocksJacobians_dh_df::col_t colF_t;
	typedef std::map<TKeyFrameID, typename ST::k2k_edge_vector_t> edges_row_t;
	static const colAp_t &colAp(const ST &st, size_t e) { return st.lin_system.dh_dAp.getCol(e); }
	static const colF_t &colF(const ST &st, size_t feat_id) {
		const mrpt::utils::map_as_vector<size_t, size_t> &remap = st.lin_system.dh_df.getColInverseRemappedIndices();
		const mrpt::utils::map_as_vector<size_t, size_t>::const_iterator it = remap.find(feat_id);
		if (it == remap.end()) throw std::runtime_error("optimize_edges: the landmark has no dh_df column (optimize_edges.h:104-108)");
		return st.lin_system.dh_df.getCol(it->second);
	}
	static size_t n_k2k(const ST &st) { return st.k2k_edges.size(); }
	static uint64_t edge_from(const ST &st, size_t e) { return st.k2k_edges[e].from; }
	static uint64_t edge_to(const ST &st, size_t e) { return st.k2k_edges[e].to; }
	static void edge_pose_get(const ST &st, size_t e, double *o) { pose_to_flat(st.k2k_edges[e].inv_pose, o); }
	static void edge_pose_set(ST &st, size_t e, const double *o) { pose_from_flat(o, st.k2k_edges[e].inv_pose); }
	static bool lm_known(const ST &st, size_t id) { return st.all_lms[id].has_known_pos; }
	static uint64_t lm_base(const ST &st, size_t id) { return st.all_lms[id].rfp->id_frame_base; }
	static double *lm_pos(const ST &st, size_t id) { return &st.all_lms[id].rfp->pos[0]; }
	static size_t n_obs(const ST &st) { return st.all_observations.size(); }
	static uint64_t obs_kf(const ST &st, size_t g) { return st.all_observations[g].obs.kf_id; }
	static uint64_t obs_feat(const ST &st, size_t g) { return st.all_observations[g].obs.obs.feat_id; }
	static bool obs_first_of_unknown(const ST &st, size_t g) { return st.all_observations[g].is_first_obs_of_unknown; }
	static void obs_data(const ST &st, size_t g, double *o) { for (int k = 0; k < (int)OBS_DIMS; k++) o[k] = st.all_observations[g].obs.obs_arr[k]; }
	static const edges_row_t *all_edges_from(const ST &st, uint64_t kf) {
		const typename ST::TSpanningTree::all_edges_maps_t::const_iterator it = st.spanning_tree.sym.all_edges.find(kf);
		return it == st.spanning_tree.sym.all_edges.end() ? NULL : &it->second;
	}
	static uint64_t edge_id(const typename ST::k2k_edge_t *e) { return e->id; }
	static void set_num(ST &st, uint64_t src, uint64_t dst, const double *o) { pose_from_flat(o, st.spanning_tree.num[src][dst].pose); }
	static void set_validity(ST &st, size_t g, char v) { st.all_observations_Jacob_validity[g] = v; }
};

// ---- srba_model_t from the engine's template arguments and `parameters` (RbaEngine.h:417-472) ----
inline void fill_sensor(srba_model_t &m, const observations::Cartesian_3D::TObservationParams &) { m.obs_kind = SRBA_OBS_CARTESIAN3D; }
inline void fill_sensor(srba_model_t &m, const observations::RelativePoses_2D::TObservationParams &) { m.obs_kind = SRBA_OBS_RELPOSES2D; }
inline void fill_sensor(srba_model_t &m, const observations::MonocularCamera::TObservationParams &p) {
	m.obs_kind = SRBA_OBS_MONO;
	m.cam_left[0] = p.camera_calib.fx(); m.cam_left[1] = p.camera_calib.fy(); m.cam_left[2] = p.camera_calib.cx(); m.cam_left[3] = p.camera_calib.cy();
}
inline void fill_sensor(srba_model_t &m, const observations::StereoCamera::TObservationParams &p) {
	m.obs_kind = SRBA_OBS_STEREO;
	const mrpt::utils::TCamera &l = p.camera_calib.leftCamera, &r = p.camera_calib.rightCamera;
	m.cam_left[0] = l.fx(); m.cam_left[1] = l.fy(); m.cam_left[2] = l.cx(); m.cam_left[3] = l.cy();
	m.cam_right[0] = r.fx(); m.cam_right[1] = r.fy(); m.cam_right[2] = r.cx(); m.cam_right[3] = r.cy();
	// right camera as seen from the left one, inverted once: R2L = (-rightCameraPose) (models/sensors.h:193,259)
	pose_to_flat(-mrpt::poses::CPose3D(p.camera_calib.rightCameraPose), m.right_cam_R2L);
}
inline void fill_sensor_pose(srba_model_t &m, const options::sensor_pose_on_robot_none::parameters_t &) { m.sensor_pose_kind = SRBA_SENSORPOSE_NONE; }
inline void fill_sensor_pose(srba_model_t &m, const options::sensor_pose_on_robot_se3::parameters_t &p) {
	m.sensor_pose_kind = SRBA_SENSORPOSE_SE3;
	pose_to_flat(p.relative_pose, m.sensor_pose);
}
inline void fill_noise(srba_model_t &m, const options::observation_noise_identity::parameters_t &p) { m.noise_kind = SRBA_NOISE_IDENTITY; m.noise_std = p.std_noise_observations; }
template <class PRM>
inline void fill_noise(srba_model_t &m, const PRM &p) {  // observation_noise_constant_matrix<OBS_T>::parameters_t: `lambda`, OBS_DIMS x OBS_DIMS
	m.noise_kind = SRBA_NOISE_CONST_MATRIX;
	const int o = (int)p.lambda.rows();
	for (int i = 0; i < o; i++)
		for (int j = 0; j < o; j++) m.noise_lambda[i * o + j] = p.lambda(i, j);
}
template <class RBA>
srba_model_t make_model(const RBA &rba) {
	srba_model_t m;
	std::memset(&m, 0, sizeof(m));
	m.pose_kind = pose_nvals<typename RBA::pose_t>::value == 3 ? SRBA_POSE_SE2 : SRBA_POSE_SE3;
	fill_sensor(m, rba.parameters.sensor);
	fill_sensor_pose(m, rba.parameters.sensor_pose);
	fill_noise(m, rba.parameters.obs_noise);
	m.max_iters = (int32_t)rba.parameters.srba.max_iters;
	m.use_robust_kernel = rba.parameters.srba.use_robust_kernel ? 1 : 0;
	m.kernel_param = rba.parameters.srba.kernel_param;
	m.max_error_per_obs_to_stop = rba.parameters.srba.max_error_per_obs_to_stop;
	m.max_rho = rba.parameters.srba.max_rho;
	m.max_lambda = rba.parameters.srba.max_lambda;
	m.min_error_reduction_ratio_to_relinearize = rba.parameters.srba.min_error_reduction_ratio_to_relinearize;
	typedef typename RBA::rba_options_t::solver_t solver_t;   // srba_options_solver.h:21-77
	m.solver_kind = solver_t::USE_SCHUR ? (solver_t::DENSE_CHOLESKY ? SRBA_SOLVER_SCHUR_DENSE_CHOLESKY : SRBA_SOLVER_SCHUR_SPARSE_CHOLESKY) : SRBA_SOLVER_NO_SCHUR_SPARSE_CHOLESKY;
	return m;
}

}  // namespace b200

template <class KF2KF_POSE_TYPE, class LM_TYPE, class OBS_TYPE, class RBA_OPTIONS>
void RbaEngine<KF2KF_POSE_TYPE, LM_TYPE, OBS_TYPE, RBA_OPTIONS>::optimize_edges(const std::vector<size_t> &run_k2k_edges_in, const std::vector<size_t> &run_feat_ids_in,
                                                                                TOptimizeExtraOutputInfo &out_info,
                                                                                const std::vector<size_t> &in_observation_indices_to_optimize) {
	typedef b200::ReferenceAccess<rba_engine_t> access_t;
	m_profiler.enter("opt");
	out_info.clear();
	std::unique_ptr<srba_b200_shim::Flat> flat =
	    srba_b200_shim::flatten<access_t>(rba_state, b200::make_model(*this), run_k2k_edges_in, run_feat_ids_in, in_observation_indices_to_optimize);
	if (!flat.get()) {  // "Skipping optimization since no observation depends on any of the given variables" (:132-138)
		m_profiler.leave("opt");
		return;
	}
	// (the per-trial trace is only requested when the user installed parameters.srba.feedback_user_iteration)
	srba_result_t res = flat->make_result_buffers(access_t::POSE_NVALS, parameters.srba.feedback_user_iteration ? 8 * (int)parameters.srba.max_iters + 16 : 0);
	if (srba_b200_optimize(&flat->p, &res) != SRBA_OK) {
		m_profiler.leave("opt");
		throw std::runtime_error(srba_b200_last_error());   // (the reference throws from its ASSERT_ macros)
	}
	srba_b200_shim::scatter<access_t>(rba_state, *flat, res);
	srba_b200_shim::replay_feedback(parameters.srba.feedback_user_iteration, *flat, res);   // optimize_edges.h:417-418,595-596 (after the call, not during)
	// landmark information matrices (optimize_edges.h:722-751, crpLandmarksApprox): the final H_f diagonal blocks of invertible landmarks
	rba_state.unknown_lms_inf_matrices.clear();
	if (parameters.srba.cov_recovery == crpLandmarksApprox)
		for (size_t i = 0; i < flat->lm_ids.size(); i++)
			if (res.lm_invertible[i]) {
				typename hessian_traits_t::TSparseBlocksHessian_f::matrix_t &M = rba_state.unknown_lms_inf_matrices[flat->lm_ids[i]];
				for (int r = 0; r < 3; r++)
					for (int c = 0; c < 3; c++) M(r, c) = res.Hf_diag[i * 9 + 3 * r + c];
			}
	// TOptimizeExtraOutputInfo (RbaEngine.h:125-177; optimize_edges.h:406-412,753-781)
	out_info.num_observations = flat->p.n_total_obs_rows;
	out_info.num_jacobians = res.stats.num_jacobians;
	out_info.num_kf2kf_edges_optimized = flat->k2k_ids.size();
	out_info.num_kf2lm_edges_optimized = flat->lm_ids.size();
	out_info.num_total_scalar_optimized = REL_POSE_DIMS * flat->k2k_ids.size() + LM_DIMS * flat->lm_ids.size();
	out_info.num_kf_optimized = flat->num_kf_optimized;
	out_info.num_lm_optimized = flat->num_lm_optimized;
	out_info.num_span_tree_numeric_updates = flat->num_span_tree_numeric_updates;
	out_info.total_sqr_error_init = res.stats.total_sqr_error_init;
	out_info.total_sqr_error_final = res.stats.total_sqr_error_final;
	out_info.obs_rmse = res.stats.obs_rmse;
	if (parameters.srba.compute_sparsity_stats) {   // optimize_edges.h:315-327: stored blocks vs rows x columns of the active part
		out_info.sparsity_dh_dAp_nnz = flat->p.n_blkA; out_info.sparsity_dh_dAp_max_size = (size_t)flat->p.n_obs * flat->p.n_unk_edges;
		out_info.sparsity_dh_df_nnz = flat->p.n_blkF; out_info.sparsity_dh_df_max_size = (size_t)flat->p.n_obs * flat->p.n_unk_lms;
		out_info.sparsity_HAp_nnz = flat->p.n_HAp; out_info.sparsity_HAp_max_size = (size_t)flat->p.n_unk_edges * flat->p.n_unk_edges;
		out_info.sparsity_Hf_nnz = flat->p.n_unk_lms; out_info.sparsity_Hf_max_size = (size_t)flat->p.n_unk_lms * flat->p.n_unk_lms;
		out_info.sparsity_HApf_nnz = flat->p.n_HApf; out_info.sparsity_HApf_max_size = (size_t)flat->p.n_unk_edges * flat->p.n_unk_lms;
	}
	// (HAp_condition_number, optimize_edges.h:753-765, needs a dense SVD of H_Ap: not provided -- left at 0)
	out_info.optimized_k2k_edge_indices.assign(flat->k2k_ids.begin(), flat->k2k_ids.end());
	out_info.optimized_landmark_indices.assign(flat->lm_ids.begin(), flat->lm_ids.end());
	m_profiler.leave("opt");
}

}  // namespace srba
