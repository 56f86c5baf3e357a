"""`-m gpu`: the srba-slam front end (srba_b200/slam_runner.py main()) on the CUDA path, configured as the reference app
configures its instances: stereo = sensor_pose_on_robot_se3 at (0,0,0,-90deg,0,-90deg) + camera calibration from --sensor-params-cfg-file
(apps/srba-slam/srba-run-generic-impl.h:72-80, instance_se3_lm3d_stereo.cpp:24-41); relative graph-SLAM = the no-Schur sparse-Cholesky
solver instance (instance_relative_graph_slam_se2.cpp:23).  Results against the oracle driven by the same runner on the same files."""
import numpy as np
import pytest

from oracle import pyoracle as po
from srba_b200 import datasets as D
from srba_b200 import slam_runner as R

pytestmark = pytest.mark.gpu


def test_stereo_dataset_through_the_cli(tmp_path):
    ds = D.landmark_slam_se3(n_kfs=25, n_lms=250, seed=4, obs_kind="stereo")
    data = str(tmp_path / "stereo_SENSOR.txt")
    R.write_dataset(data, ds["frames"], "stereo")
    cam, b = ds["cam"], ds["baseline"]
    cfg = tmp_path / "cam.ini"
    cfg.write_text("[CAMERA_LEFT]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\n[CAMERA_RIGHT]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\n[CAMERA_LEFT2RIGHT_POSE]\npose_quaternion=[%r 0 0 1 0 0 0]\n"
                   % (cam[0], cam[1], cam[2], cam[3], cam[0], cam[1], cam[2], cam[3], b))
    prefix = str(tmp_path / "prof")
    summary = R.main(["--se3", "--lm-3d", "--obs", "StereoCamera", "-d", data, "--sensor-params-cfg-file", str(cfg), "--submap-size", "10",
                      "--max-spanning-tree-depth", "3", "--max-optimize-depth", "3", "--noise", "0.5", "--profile-stats", prefix, "--profile-stats-length", "5"])
    # the oracle through the same runner, same files, same parameters as main() sets them
    frames, npar = R.load_dataset(data, "stereo", noise=0.5)
    kw = dict(pose_kind=1, obs_kind=3, noise_kind=0, noise_std=npar["noise_std"], sensor_pose_kind=1, sensor_pose=[0, 0, 0, np.deg2rad(-90.0), 0.0, np.deg2rad(-90.0)],
              submap_size=10, max_tree_depth=3, max_optimize_depth=3, max_iters=20, min_obs_to_loop_closure=4, max_error_per_obs_to_stop=1e-8,
              **R.load_sensor_params(str(cfg), "stereo"))
    eo = po.OracleEngine(po.default_params(**kw))
    _, so = R.run_slam(eo, frames)
    assert summary["keyframes"] == so["keyframes"] == 25
    # (0.5 px noise on each of the 4 pixel coordinates: the per-observation RMSE sits near 1 px)
    assert summary["final_rmse"] < 1.5 and abs(summary["final_rmse"] - so["final_rmse"]) <= 1e-3 * so["final_rmse"] + 1e-9   # free-running engines, pixels
    means = open(prefix + "_means.csv").read().splitlines()
    assert [l.split("\t")[0] for l in means[1:]] == ["0", "5", "10", "15", "20"] and '"opt.gpu_call"' in means[0]
    print("slam_runner stereo: final RMSE %.4g px (oracle %.4g), %d LM trials" % (summary["final_rmse"], so["final_rmse"], summary["lm_trials"]))


def test_relative_graph_slam_dataset_through_the_cli(tmp_path):
    ds = D.graphslam_se2_tour(80, seed=6, add_noise=False)
    data = str(tmp_path / "rel_SENSOR.txt")
    R.write_dataset(data, ds["frames"], "relposes2d", graph_slam=True)
    summary = R.main(["--se2", "--graph-slam", "-d", data, "--submap-size", "10", "--max-spanning-tree-depth", "3", "--max-optimize-depth", "3",
                      "--noise", "0.001", "--noise-ang", "0.2", "--add-noise", "--seed", "3"])
    assert summary["keyframes"] == 80 and summary["final_rmse"] < 0.01 and summary["lm_trials"] > 80
