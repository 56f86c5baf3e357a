// The reference's examples/cpp/tutorial-srba-relative-graph-slam-se2.cpp against shim/srba.h: same SRBA API calls (options :21-27,
// typedef :29-34, parameters :85-106, per-keyframe observation lists :121-158, define_new_keyframe :162); removed: GUI / .dot / random
// noise; the embedded dataset comes from a text file written by the test (tests/golden/reference_fixtures.json).
#include <srba.h>

#include <cstdio>
#include <fstream>
#include <iostream>

using namespace srba;
using namespace std;

struct RBA_OPTIONS : public RBA_OPTIONS_DEFAULT {
	typedef options::observation_noise_constant_matrix<observations::RelativePoses_2D> obs_noise_matrix_t;  // a given constant information matrix
};

typedef RbaEngine<kf2kf_poses::SE2, landmarks::RelativePoses2D, observations::RelativePoses_2D, RBA_OPTIONS> my_srba_t;

struct basic_graph_slam_dataset_entry_t { unsigned int current_kf, observed_kf; double x, y, z, yaw, pitch, roll; };
static double square(double v) { return v * v; }

int main(int argc, char **argv) {
	const double STD_NOISE_XY = 0.001, STD_NOISE_YAW = DEG2RAD(0.05);
	ifstream f(argv[1]);
	size_t nObs;
	f >> nObs;
	vector<basic_graph_slam_dataset_entry_t> dataset(nObs);
	for (size_t i = 0; i < nObs; i++) f >> dataset[i].current_kf >> dataset[i].observed_kf >> dataset[i].x >> dataset[i].y >> dataset[i].z >> dataset[i].yaw >> dataset[i].pitch >> dataset[i].roll;

	my_srba_t rba;
	rba.setVerbosityLevel(1);
	rba.parameters.srba.use_robust_kernel = false;
	{
		fixed_matrix<3> ObsL;   // (Eigen::Matrix3d in the tutorial)
		ObsL.setZero();
		ObsL(0, 0) = 1 / square(STD_NOISE_XY);
		ObsL(1, 1) = 1 / square(STD_NOISE_XY);
		ObsL(2, 2) = 1 / square(STD_NOISE_YAW);
		rba.parameters.obs_noise.lambda = ObsL;
	}
	rba.parameters.srba.max_tree_depth = 3;
	rba.parameters.srba.max_optimize_depth = 3;
	rba.parameters.ecp.submap_size = 5;
	rba.parameters.ecp.min_obs_to_loop_closure = 1;

	size_t cur_kf = 0;
	double worst_rmse = 0;
	for (size_t obsIdx = 0; obsIdx < nObs; cur_kf++) {
		my_srba_t::new_kf_observations_t list_obs;
		{  // each keyframe MUST have exactly ONE fixed "fake landmark", representing its pose
			my_srba_t::new_kf_observation_t obs_field;
			obs_field.is_fixed = true;
			obs_field.obs.feat_id = cur_kf;
			obs_field.obs.obs_data.x = 0;
			obs_field.obs.obs_data.y = 0;
			obs_field.obs.obs_data.yaw = 0;
			list_obs.push_back(obs_field);
		}
		while (obsIdx < nObs && dataset[obsIdx].current_kf == cur_kf) {
			my_srba_t::new_kf_observation_t obs_field;
			obs_field.is_fixed = false;
			obs_field.is_unknown_with_init_val = false;
			obs_field.obs.feat_id = dataset[obsIdx].observed_kf;
			obs_field.obs.obs_data.x = dataset[obsIdx].x;
			obs_field.obs.obs_data.y = dataset[obsIdx].y;
			obs_field.obs.obs_data.yaw = dataset[obsIdx].yaw;
			list_obs.push_back(obs_field);
			obsIdx++;
		}
		my_srba_t::TNewKeyFrameInfo new_kf_info;
		rba.define_new_keyframe(list_obs, new_kf_info, true);
		if (new_kf_info.optimize_results.num_observations) worst_rmse = max(worst_rmse, new_kf_info.optimize_results.obs_rmse);
		cout << "Created KF #" << new_kf_info.kf_id << " | # kf-to-kf edges created:" << new_kf_info.created_edge_ids.size()
		     << " | error: " << new_kf_info.optimize_results.total_sqr_error_init << " -> " << new_kf_info.optimize_results.total_sqr_error_final << endl;
	}
	printf("RESULT kfs %zu edges %zu worst_rmse %.17g overall %.17g\n", rba.get_rba_state().num_keyframes, rba.get_k2k_edges().size(), worst_rmse, rba.eval_overall_squared_error());
	for (size_t i = 0; i < rba.get_k2k_edges().size(); i++) {
		const my_srba_t::k2k_edge_t &e = rba.get_k2k_edges()[i];
		printf("EDGE %zu %llu %llu %.17g %.17g %.17g\n", e.id, (unsigned long long)e.from, (unsigned long long)e.to, e.inv_pose.x(), e.inv_pose.y(), e.inv_pose.phi());
	}
	return 0;
}
