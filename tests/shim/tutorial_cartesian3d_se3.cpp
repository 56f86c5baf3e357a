// The reference's examples/cpp/tutorial-srba-cartesian3d-se3.cpp against shim/srba.h (the srba_b200 facade): the SRBA API calls are the
// tutorial's, line for line (typedefs :22-34, parameters :117-126, observation lists :146-166 / :183-203, define_new_keyframe :157,196,
// result access :210-211); removed: the GUI / OpenGL / .dot lines and the random noise; the embedded dataset is read from a text file
// written by the test (tests/golden/reference_fixtures.json, extracted from the tutorial) instead of being compiled in.
#include <srba.h>

#include <cstdio>
#include <fstream>
#include <iostream>

using namespace srba;
using namespace std;

struct RBA_OPTIONS : public RBA_OPTIONS_DEFAULT {};

typedef RbaEngine<kf2kf_poses::SE3,            // Parameterization  of KF-to-KF poses
                  landmarks::Euclidean3D,      // Parameterization of landmark positions
                  observations::Cartesian_3D,  // Type of observations
                  RBA_OPTIONS>
    my_srba_t;

struct basic_euclidean_dataset_entry_t { unsigned int landmark_id; double x, y, z; };

static vector<basic_euclidean_dataset_entry_t> read_block(istream &f) {
	size_t n;
	f >> n;
	vector<basic_euclidean_dataset_entry_t> v(n);
	for (size_t i = 0; i < n; i++) f >> v[i].landmark_id >> v[i].x >> v[i].y >> v[i].z;
	return v;
}

static int g_fb_calls = 0;
static unsigned g_fb_first_iter = 99;
static double g_fb_first_err = -1, g_fb_last_err = -1;
static void on_iteration(unsigned int iter, const double total_sq_err, const double mean_sqroot_error) {
	if (!g_fb_calls++) { g_fb_first_iter = iter; g_fb_first_err = total_sq_err; }
	g_fb_last_err = total_sq_err;
	(void)mean_sqroot_error;
}

int main(int argc, char **argv) {
	ifstream f(argv[1]);
	const vector<basic_euclidean_dataset_entry_t> observations_0 = read_block(f), observations_10 = read_block(f);

	my_srba_t rba;  //  Create an empty RBA problem
	rba.setVerbosityLevel(1);
	rba.parameters.srba.use_robust_kernel = true;
	rba.parameters.obs_noise.std_noise_observations = 0.1;
	rba.parameters.srba.max_tree_depth = 3;
	rba.parameters.srba.max_optimize_depth = 3;

	my_srba_t::new_kf_observations_t list_obs;
	my_srba_t::new_kf_observation_t obs_field;
	obs_field.is_fixed = false;
	obs_field.is_unknown_with_init_val = false;
	for (size_t i = 0; i < observations_0.size(); i++) {
		obs_field.obs.feat_id = observations_0[i].landmark_id;
		obs_field.obs.obs_data.pt.x = observations_0[i].x;
		obs_field.obs.obs_data.pt.y = observations_0[i].y;
		obs_field.obs.obs_data.pt.z = observations_0[i].z;
		list_obs.push_back(obs_field);
	}
	my_srba_t::TNewKeyFrameInfo new_kf_info;
	rba.define_new_keyframe(list_obs, new_kf_info, true);
	cout << "Created KF #" << new_kf_info.kf_id << " | # kf-to-kf edges created:" << new_kf_info.created_edge_ids.size() << endl
	     << "Optimization error: " << new_kf_info.optimize_results.total_sqr_error_init << " -> " << new_kf_info.optimize_results.total_sqr_error_final << endl;

	list_obs.clear();
	for (size_t i = 0; i < observations_10.size(); i++) {
		obs_field.obs.feat_id = observations_10[i].landmark_id;
		obs_field.obs.obs_data.pt.x = observations_10[i].x;
		obs_field.obs.obs_data.pt.y = observations_10[i].y;
		obs_field.obs.obs_data.pt.z = observations_10[i].z;
		list_obs.push_back(obs_field);
	}
	rba.define_new_keyframe(list_obs, new_kf_info, true);
	cout << "Created KF #" << new_kf_info.kf_id << " | # kf-to-kf edges created:" << new_kf_info.created_edge_ids.size() << endl
	     << "Optimization error: " << new_kf_info.optimize_results.total_sqr_error_init << " -> " << new_kf_info.optimize_results.total_sqr_error_final << endl;

	cout << "inv_pose of KF-to-KF edge #0 (relative pose of KF#0 wrt KF#1):\n" << rba.get_k2k_edges()[0].inv_pose << endl;
	const mrpt::poses::CPose3D rel = -rba.get_k2k_edges()[0].inv_pose;
	cout << "Relative pose of KF#1 wrt KF#0:\n" << rel << endl;

	// machine-readable line for the test
	printf("RESULT kf %llu edges %zu err_init %.17g err_final %.17g rmse %.17g obs %zu pose", (unsigned long long)new_kf_info.kf_id, rba.get_k2k_edges().size(),
	       new_kf_info.optimize_results.total_sqr_error_init, new_kf_info.optimize_results.total_sqr_error_final, new_kf_info.optimize_results.obs_rmse,
	       new_kf_info.optimize_results.num_observations);
	for (int i = 0; i < 9; i++) printf(" %.17g", rel.rotation()[i]);
	for (int i = 0; i < 3; i++) printf(" %.17g", rel.translation()[i]);
	printf("\n");
	// optimize_local_area + observation_indices_to_optimize through the facade (RbaEngine.h:235-241)
	my_srba_t::TOptimizeExtraOutputInfo info;
	rba.parameters.srba.feedback_user_iteration = &on_iteration;   // RbaEngine.h:307 (replayed after the call)
	rba.optimize_local_area(1, 3, info);
	rba.parameters.srba.feedback_user_iteration = NULL;
	printf("FEEDBACK calls %d first_iter %u first_err %.17g last_err %.17g init %.17g final %.17g\n", g_fb_calls, g_fb_first_iter, g_fb_first_err, g_fb_last_err,
	       info.total_sqr_error_init, info.total_sqr_error_final);
	vector<size_t> all_obs;
	for (size_t i = 0; i < rba.get_rba_state().num_observations; i++) all_obs.push_back(i);
	my_srba_t::TOptimizeExtraOutputInfo info2;
	rba.optimize_local_area(1, 3, info2, my_srba_t::TOptimizeLocalAreaParams(), all_obs);
	printf("LOCALAREA obs %zu k2k %zu lms %zu final %.17g | listed obs %zu final %.17g | overall %.17g\n", info.num_observations, info.optimized_k2k_edge_indices.size(),
	       info.optimized_landmark_indices.size(), info.total_sqr_error_final, info2.num_observations, info2.total_sqr_error_final, rba.eval_overall_squared_error());
	return 0;
}
