"""CPU-only checks of the srba-slam dataset front end (srba_b200/slam_runner.py): the reference's text
format round-trips, frames are grouped exactly like apps/srba-slam/srba-run-generic-impl.h:337-465 (fixed fake landmark
first in --graph-slam mode), and the runner drives an engine keyframe by keyframe (here the CPU oracle mirror)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from srba_b200 import datasets as D
from srba_b200 import slam_runner as R


def test_relgraph2d_text_format_round_trip(tmp_path):
    ds = D.graphslam_se2_tour(40, seed=5, add_noise=False)
    p = str(tmp_path / "ds_SENSOR.txt")
    R.write_dataset(p, ds["frames"], "relposes2d", graph_slam=True)
    frames, par = R.load_dataset(p, "relposes2d", graph_slam=True)
    assert len(frames) == 40 and par["sigma_xy"] == 0.10 and abs(par["sigma_yaw"] - np.radians(4)) < 1e-15   # reference defaults
    for a, b in zip(ds["frames"], frames):
        assert a["feat_id"].tolist() == b["feat_id"].tolist() and a["is_fixed"].tolist() == b["is_fixed"].tolist()
        np.testing.assert_allclose(a["obs"], b["obs"], rtol=0, atol=1e-15)
    # the fake landmark of --graph-slam mode: first entry, fixed, feat_id == keyframe id
    for k, f in enumerate(frames):
        assert f["is_fixed"][0] == 1 and int(f["feat_id"][0]) == k and not f["is_fixed"][1:].any()
    noisy, _ = R.load_dataset(p, "relposes2d", graph_slam=True, add_noise=True, noise=0.001, noise_ang_deg=0.2, seed=1)
    d = np.concatenate([n["obs"][1:] - c["obs"][1:] for n, c in zip(noisy, frames) if len(c["obs"]) > 1])
    assert 0.0005 < d[:, 0].std() < 0.002 and 0.001 < np.degrees(d[:, 2].std()) < 0.4


def test_column_count_is_checked(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_text("% comment\n0 1 1.0 2.0\n0 2 1.0 2.0\n1 1 1.0 2.0\n")
    with pytest.raises(ValueError):
        R.load_dataset(str(p), "cartesian3d")
    frames, par = R.load_dataset(str(p), "mono")
    assert len(frames) == 2 and frames[0]["obs"].shape == (2, 2) and par["noise_std"] == 1e-4


def test_runner_drives_an_engine(tmp_path):
    ds = D.graphslam_se2_tour(30, seed=2, add_noise=False)
    p = str(tmp_path / "ds.txt")
    R.write_dataset(p, ds["frames"], "relposes2d", graph_slam=True)
    frames, par = R.load_dataset(p, "relposes2d", graph_slam=True, noise=0.001, noise_ang_deg=0.2)
    eng = po.OracleEngine(po.default_params(pose_kind=0, obs_kind=0, noise_kind=1, noise_lambda=D.se2_noise_lambda(par["sigma_xy"], par["sigma_yaw"]),
                                            max_tree_depth=3, max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1, max_error_per_obs_to_stop=1e-8))
    rows, summary = R.run_slam(eng, frames, profile_prefix=str(tmp_path / "prof"))
    assert summary["keyframes"] == 30 and summary["final_rmse"] < 1e-6      # noise-free data: the map is recovered exactly
    tab = np.loadtxt(str(tmp_path / "prof_keyframes.csv"), delimiter=",", skiprows=1)
    assert tab.shape == (30, 8)
    # --profile-stats files (srba-run-generic-impl.h:623-647,778-829): one row per segment of 10 keyframes, tab-separated
    means = open(str(tmp_path / "prof_means.csv")).read().splitlines()
    calls = open(str(tmp_path / "prof_calls.csv")).read().splitlines()
    assert means[0].startswith("timestep\t") and '"define_new_keyframe"' in means[0] and '"opt"' in means[0] and '"total_num_kfs"' in means[0]
    assert [l.split("\t")[0] for l in means[1:]] == ["0", "10", "20"] and len(calls) == 4
    cols = [c.strip().strip('"') for c in means[0].split("\t")]
    assert [float(l.split("\t")[cols.index("total_num_kfs")]) for l in means[1:]] == [1, 11, 21]


def test_sensor_params_cfg_file(tmp_path):
    """--sensor-params-cfg-file as the app's camera instances read it (instance_se3_lm3d_monocular.cpp:24-28, instance_se3_lm3d_stereo.cpp:24-31)."""
    mono = tmp_path / "mono.ini"
    mono.write_text("[CAMERA]\nresolution = [800 600]\ncx = 400\ncy = 300\nfx = 200\nfy = 200\ndist = [0 0 0 0 0]\n")
    assert R.load_sensor_params(str(mono), "mono") == dict(cam_left=[200.0, 200.0, 400.0, 300.0])
    st = tmp_path / "stereo.ini"
    st.write_text("[CAMERA_LEFT]\nresolution=[1024 768]\ncx=512\ncy=384\nfx=200\nfy=150\n[CAMERA_RIGHT]\ncx=512\ncy=384\nfx=200\nfy=150\n"
                  "[CAMERA_LEFT2RIGHT_POSE]\npose_quaternion=[0.2 0 0 1 0 0 0]\n")
    p = R.load_sensor_params(str(st), "stereo")
    assert p["cam_left"] == [200.0, 150.0, 512.0, 384.0] and p["cam_right"] == p["cam_left"] and p["right_cam_pose"] == [0.2, 0, 0, 1, 0, 0, 0]
    bad = tmp_path / "bad.ini"
    bad.write_text("[CAMERA_LEFT]\ncx=1\ncy=1\nfx=1\nfy=1\n[CAMERA_RIGHT]\ncx=1\ncy=1\nfx=1\nfy=1\n[CAMERA_LEFT2RIGHT_POSE]\npose_quaternion=[0 0 0 1 0 0 0]\n")
    with pytest.raises(ValueError):
        R.load_sensor_params(str(bad), "stereo")      # ASSERT_(baseline != 0)
