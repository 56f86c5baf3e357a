// DECLARATIONS ONLY (test infrastructure): the names of the reference's / MRPT's interface that shim/srba/impl/optimize_edges.h touches,
// restated so that the replacement header can be TYPE-CHECKED in an image without MRPT + Eigen.  Each item cites where the reference
// declares it.  Nothing here computes anything; tests/test_shim.py compiles tests/shim/seam_reference_syntax.cpp (explicit instantiations of
// RbaEngine<>::optimize_edges for the BASELINE model combinations) against it.
#pragma once
#include <stdint.h>

#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "../../../include/srba_b200.h"

namespace mrpt {
namespace math {
struct CMatrixDouble33 { double m[9]; double &operator()(int r, int c) { return m[3 * r + c]; } double operator()(int r, int c) const { return m[3 * r + c]; } };
}
namespace poses {
class CPose3DQuat { public: double operator[](int) const; };
class CPose2D { public: double x() const; double y() const; double phi() const; void x(double); void y(double); void phi(double); };
class CPose3D {
public:
	CPose3D();
	explicit CPose3D(const CPose3DQuat &);
	const math::CMatrixDouble33 &getRotationMatrix() const;
	void setRotationMatrix(const math::CMatrixDouble33 &);
	double x() const; double y() const; double z() const;
	void x(double); void y(double); void z(double);
	CPose3D operator-() const;
};
}  // namespace poses
namespace utils {
template <class K, class V>
class map_as_vector {  // mrpt/utils/map_as_vector.h: find() is O(1), iterators point to std::pair<K,V>
public:
	typedef std::pair<K, V> value_type;
	typedef typename std::vector<value_type>::iterator iterator;
	typedef typename std::vector<value_type>::const_iterator const_iterator;
	iterator find(const K &);
	const_iterator find(const K &) const;
	iterator end();
	const_iterator end() const;
	V &operator[](const K &);
	size_t size() const;
};
class TCamera { public: double fx() const; double fy() const; double cx() const; double cy() const; };
class TStereoCamera { public: TCamera leftCamera, rightCamera; poses::CPose3DQuat rightCameraPose; };
class CTimeLogger { public: void enter(const char *); double leave(const char *); };
}  // namespace utils
}  // namespace mrpt

namespace srba {
typedef uint64_t TKeyFrameID;  // srba_types.h:21
typedef uint64_t TLandmarkID;  // srba_types.h:22
enum TCovarianceRecoveryPolicy { crpNone = 0, crpLandmarksApprox };  // srba_types.h:30-34

namespace kf2kf_poses {  // models/kf2kf_poses.h:24-40
struct SE2 { static const size_t REL_POSE_DIMS = 3; typedef mrpt::poses::CPose2D pose_t; };
struct SE3 { static const size_t REL_POSE_DIMS = 6; typedef mrpt::poses::CPose3D pose_t; };
}
namespace landmarks {  // models/landmarks.h
struct Euclidean3D { static const size_t LM_DIMS = 3; };
struct RelativePoses2D { static const size_t LM_DIMS = 3; };
}
namespace observations {  // models/observations_*.h: OBS_DIMS + TObservationParams
struct Cartesian_3D { static const size_t OBS_DIMS = 3; struct TObservationParams {}; };
struct RelativePoses_2D { static const size_t OBS_DIMS = 3; struct TObservationParams {}; };
struct MonocularCamera { static const size_t OBS_DIMS = 2; struct TObservationParams { mrpt::utils::TCamera camera_calib; }; };
struct StereoCamera { static const size_t OBS_DIMS = 4; struct TObservationParams { mrpt::utils::TStereoCamera camera_calib; }; };
}
namespace options {
struct sensor_pose_on_robot_none { struct parameters_t {}; };                                           // srba_options_sensor_pose.h:28-60
struct sensor_pose_on_robot_se3 { struct parameters_t { mrpt::poses::CPose3D relative_pose; }; };      // :64-115
struct observation_noise_identity { struct parameters_t { double std_noise_observations; }; };         // srba_options_noise.h:27-63
template <int N> struct eigen_like_matrix { double operator()(int, int) const; int rows() const; };
template <class OBS_T> struct observation_noise_constant_matrix { struct parameters_t { eigen_like_matrix<OBS_T::OBS_DIMS> lambda; }; };  // :70-115
struct solver_LM_schur_dense_cholesky { static const bool USE_SCHUR = true; static const bool DENSE_CHOLESKY = true; struct extra_results_t { void clear(); }; };     // srba_options_solver.h:21-37
struct solver_LM_schur_sparse_cholesky { static const bool USE_SCHUR = true; static const bool DENSE_CHOLESKY = false; struct extra_results_t { void clear(); }; };  // :41-57
struct solver_LM_no_schur_sparse_cholesky { static const bool USE_SCHUR = false; static const bool DENSE_CHOLESKY = false; struct extra_results_t { void clear(); }; };  // :61-77
}  // namespace options
struct RBA_OPTIONS_DEFAULT {  // RbaEngine.h:39-45
	typedef options::sensor_pose_on_robot_none sensor_pose_on_robot_t;
	typedef options::observation_noise_identity obs_noise_matrix_t;
	typedef options::solver_LM_schur_dense_cholesky solver_t;
};

template <class ENTRY>
struct block_sparse_cols {  // mrpt::math::MatrixBlockSparseCols<> as used at impl/add-observations.h:179,225,235
	typedef ENTRY TEntry;
	typedef std::map<size_t, ENTRY> col_t;
	const col_t &getCol(size_t) const;
	col_t &getCol(size_t);
	const mrpt::utils::map_as_vector<size_t, size_t> &getColInverseRemappedIndices() const;
};

template <class KF2KF_POSE_TYPE, class LM_TYPE, class OBS_TYPE, class RBA_OPTIONS>
struct TRBA_Problem_state {  // srba_types.h:548-785
	typedef typename KF2KF_POSE_TYPE::pose_t pose_t;
	struct k2k_edge_t { TKeyFrameID from, to; pose_t inv_pose; size_t id; };                         // :82-91
	typedef std::deque<k2k_edge_t *> k2k_edge_vector_t;                                                // :93
	struct pose_flag_t { pose_t pose; mutable bool updated; };                                         // :57-68
	typedef std::map<TKeyFrameID, pose_flag_t> frameid2pose_map_t;
	struct TRelativeLandmarkPos { TKeyFrameID id_frame_base; double pos[LM_TYPE::LM_DIMS]; };          // :106-121
	struct TLandmarkEntry { bool has_known_pos; TRelativeLandmarkPos *rfp; };                          // :127-135
	struct observation_t { TLandmarkID feat_id; };                                                     // :310-316
	struct kf_observation_t { observation_t obs; double obs_arr[OBS_TYPE::OBS_DIMS]; TKeyFrameID kf_id; };  // :501-511
	struct k2f_edge_t { kf_observation_t obs; bool feat_has_known_rel_pos, is_first_obs_of_unknown; TRelativeLandmarkPos *feat_rel_pos; };  // :515-525
	struct TJacobianSymbolicInfo_dh_dAp { bool edge_normal_dir; TKeyFrameID kf_d, kf_base; size_t k2k_edge_id, obs_idx; char *is_valid; };  // :204-240
	struct TJacobianSymbolicInfo_dh_df { size_t obs_idx; char *is_valid; };                                                                 // :246-266
	struct TEntryAp { double num[OBS_TYPE::OBS_DIMS * KF2KF_POSE_TYPE::REL_POSE_DIMS]; TJacobianSymbolicInfo_dh_dAp sym; };
	struct TEntryF { double num[OBS_TYPE::OBS_DIMS * LM_TYPE::LM_DIMS]; TJacobianSymbolicInfo_dh_df sym; };
	struct TSpanningTree {  // :560-660
		typedef mrpt::utils::map_as_vector<TKeyFrameID, std::map<TKeyFrameID, k2k_edge_vector_t> > all_edges_maps_t;
		struct { all_edges_maps_t all_edges; } sym;
		mrpt::utils::map_as_vector<TKeyFrameID, frameid2pose_map_t> num;
	};
	struct TLinearSystem { block_sparse_cols<TEntryAp> dh_dAp; block_sparse_cols<TEntryF> dh_df; };  // :670-690
	std::deque<k2k_edge_t> k2k_edges;
	std::deque<TLandmarkEntry> all_lms;
	TSpanningTree spanning_tree;
	std::deque<k2f_edge_t> all_observations;
	TLinearSystem lin_system;
	std::deque<char> all_observations_Jacob_validity;
	struct matrix33_t { double &operator()(int, int); };
	std::map<TLandmarkID, matrix33_t> unknown_lms_inf_matrices;
};

template <class KF2KF_POSE_TYPE, class LM_TYPE, class OBS_TYPE, class RBA_OPTIONS = RBA_OPTIONS_DEFAULT>
class RbaEngine {  // RbaEngine.h:66-816 (only what optimize_edges touches)
public:
	typedef RbaEngine<KF2KF_POSE_TYPE, LM_TYPE, OBS_TYPE, RBA_OPTIONS> rba_engine_t;
	typedef RBA_OPTIONS rba_options_t;
	typedef typename KF2KF_POSE_TYPE::pose_t pose_t;
	static const size_t REL_POSE_DIMS = KF2KF_POSE_TYPE::REL_POSE_DIMS, LM_DIMS = LM_TYPE::LM_DIMS, OBS_DIMS = OBS_TYPE::OBS_DIMS;
	typedef TRBA_Problem_state<KF2KF_POSE_TYPE, LM_TYPE, OBS_TYPE, RBA_OPTIONS> rba_problem_state_t;
	typedef block_sparse_cols<typename rba_problem_state_t::TEntryAp> TSparseBlocksJacobians_dh_dAp;
	typedef block_sparse_cols<typename rba_problem_state_t::TEntryF> TSparseBlocksJacobians_dh_df;
	struct hessian_traits_t { struct TSparseBlocksHessian_f { typedef typename rba_problem_state_t::matrix33_t matrix_t; }; };
	struct TOptimizeExtraOutputInfo {  // RbaEngine.h:125-177
		size_t num_observations, num_jacobians, num_kf2kf_edges_optimized, num_kf2lm_edges_optimized, num_total_scalar_optimized, num_kf_optimized, num_lm_optimized,
		    num_span_tree_numeric_updates;
		double obs_rmse, total_sqr_error_init, total_sqr_error_final, HAp_condition_number;
		size_t sparsity_dh_dAp_nnz, sparsity_dh_dAp_max_size, sparsity_dh_df_nnz, sparsity_dh_df_max_size, sparsity_HAp_nnz, sparsity_HAp_max_size, sparsity_Hf_nnz,
		    sparsity_Hf_max_size, sparsity_HApf_nnz, sparsity_HApf_max_size;
		std::vector<size_t> optimized_k2k_edge_indices, optimized_landmark_indices;
		void clear();
	};
	struct TSRBAParameters {  // RbaEngine.h:417-460
		size_t max_iters;
		bool use_robust_kernel;
		double kernel_param, max_error_per_obs_to_stop, max_rho, max_lambda, min_error_reduction_ratio_to_relinearize;
		void (*feedback_user_iteration)(unsigned int iter, const double total_sq_err, const double mean_sqroot_error);
		bool compute_condition_number, compute_sparsity_stats;
		TCovarianceRecoveryPolicy cov_recovery;
	};
	struct TAllParameters {  // RbaEngine.h:463-472
		TSRBAParameters srba;
		typename OBS_TYPE::TObservationParams sensor;
		typename RBA_OPTIONS::sensor_pose_on_robot_t::parameters_t sensor_pose;
		typename RBA_OPTIONS::obs_noise_matrix_t::parameters_t obs_noise;
	};
	TAllParameters parameters;

protected:
	void optimize_edges(const std::vector<size_t> &run_k2k_edges, const std::vector<size_t> &run_feat_ids_in, TOptimizeExtraOutputInfo &out_info,
	                    const std::vector<size_t> &in_observation_indices_to_optimize = std::vector<size_t>());  // RbaEngine.h:533-538
	rba_problem_state_t rba_state;           // RbaEngine.h:500
	mrpt::utils::CTimeLogger m_profiler;     // RbaEngine.h:503
};
}  // namespace srba
