// Compile/link check of shim/srba.h for the remaining two model combinations BASELINE.json names (monocular, stereo: options and
// parameters as in examples/cpp/tutorial-srba-monocular-se3.cpp:19-33,116-135 and tutorial-srba-stereo-se3.cpp:17-31,126-147) and the
// sparse-Cholesky solver options (srba_options_solver.h:41-77).  Runs nothing on the device unless called with "run".
#include <srba.h>

#include <cstdio>
#include <cstring>
using namespace srba;

struct OPT_CAM : public RBA_OPTIONS_DEFAULT {
	typedef options::sensor_pose_on_robot_se3 sensor_pose_on_robot_t;
	typedef options::observation_noise_identity obs_noise_matrix_t;
};
struct OPT_SPARSE : public OPT_CAM { typedef options::solver_LM_schur_sparse_cholesky solver_t; };
struct OPT_NOSCHUR : public RBA_OPTIONS_DEFAULT { typedef options::solver_LM_no_schur_sparse_cholesky solver_t; };
typedef RbaEngine<kf2kf_poses::SE3, landmarks::Euclidean3D, observations::MonocularCamera, OPT_CAM> mono_t;
typedef RbaEngine<kf2kf_poses::SE3, landmarks::Euclidean3D, observations::StereoCamera, OPT_SPARSE> stereo_t;
typedef RbaEngine<kf2kf_poses::SE3, landmarks::Euclidean3D, observations::Cartesian_3D, OPT_NOSCHUR> c3d_noschur_t;

template <class RBA>
static void touch(RBA &rba) {
	typename RBA::new_kf_observations_t list_obs;
	typename RBA::TNewKeyFrameInfo info;
	typename RBA::TOptimizeExtraOutputInfo out;
	rba.define_new_keyframe(list_obs, info, true);
	rba.optimize_local_area(0, 3, out);
	(void)rba.get_k2k_edges();
	(void)rba.get_unknown_feats();
	typename RBA::frameid2pose_map_t st;
	rba.create_complete_spanning_tree(0, st, 4);
}

int main(int argc, char **argv) {
	mono_t m;
	mrpt::utils::TCamera &c = m.parameters.sensor.camera_calib;
	c.ncols = 800; c.nrows = 600;
	c.cx(400); c.cy(300); c.fx(200); c.fy(200);
	m.parameters.sensor_pose.relative_pose = mrpt::poses::CPose3D(0, 0, 0, DEG2RAD(-90), DEG2RAD(0), DEG2RAD(-90));
	m.parameters.obs_noise.std_noise_observations = 0.5;
	stereo_t s;
	mrpt::utils::TCamera &lc = s.parameters.sensor.camera_calib.leftCamera;
	lc.ncols = 1024; lc.nrows = 768; lc.cx(512); lc.cy(384); lc.fx(200); lc.fy(150);
	s.parameters.sensor.camera_calib.rightCamera = lc;
	s.parameters.sensor.camera_calib.rightCameraPose.fromString("[0.2 0 0  1 0 0 0]");
	s.parameters.sensor_pose.relative_pose = mrpt::poses::CPose3D(0, 0, 0, DEG2RAD(-90), DEG2RAD(0), DEG2RAD(-90));
	c3d_noschur_t n;
	if (argc > 1 && !std::strcmp(argv[1], "run")) { touch(m); touch(s); touch(n); }
	std::printf("ok %zu %zu %zu\n", (size_t)mono_t::OBS_DIMS, (size_t)stereo_t::OBS_DIMS, (size_t)c3d_noschur_t::REL_POSE_DIMS);
	return 0;
}
