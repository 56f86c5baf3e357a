// Type-check of shim/srba/impl/optimize_edges.h (the replacement member template) against the declarations of the reference / MRPT interface
// it uses (tests/shim/reference_stub/srba_reference_stub.h): explicit instantiations for the BASELINE model / option combinations.
// Compiled with -fsyntax-only by tests/test_shim.py -- MRPT itself is not available in this image.
#include "reference_stub/srba_reference_stub.h"
#include <srba/impl/optimize_edges.h>

namespace srba {
struct OPT_GRAPH : public RBA_OPTIONS_DEFAULT {
	typedef options::observation_noise_constant_matrix<observations::RelativePoses_2D> obs_noise_matrix_t;
	typedef options::solver_LM_no_schur_sparse_cholesky solver_t;
};
struct OPT_CAM : public RBA_OPTIONS_DEFAULT {
	typedef options::sensor_pose_on_robot_se3 sensor_pose_on_robot_t;
	typedef options::solver_LM_schur_sparse_cholesky solver_t;
};
template class RbaEngine<kf2kf_poses::SE2, landmarks::RelativePoses2D, observations::RelativePoses_2D, OPT_GRAPH>;
template class RbaEngine<kf2kf_poses::SE3, landmarks::Euclidean3D, observations::Cartesian_3D, RBA_OPTIONS_DEFAULT>;
template class RbaEngine<kf2kf_poses::SE3, landmarks::Euclidean3D, observations::MonocularCamera, OPT_CAM>;
template class RbaEngine<kf2kf_poses::SE3, landmarks::Euclidean3D, observations::StereoCamera, OPT_CAM>;
}  // namespace srba
