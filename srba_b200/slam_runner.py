"""`srba-slam`-style dataset front end for the B200 hot path.

Reads the reference's dataset TEXT format -- one observation per line, columns `FRAME_ID FEAT_ID [sensor fields]`
(apps/srba-slam/CDatasetParserBase.h:56-112; lines starting with % or # are comments, as mrpt loadFromTextFile) -- and
drives `define_new_keyframe()` once per keyframe exactly like apps/srba-slam/srba-run-generic-impl.h:337-509: in
--graph-slam mode every keyframe first gets its own FIXED fake landmark (feat_id == keyframe id, :370-377), then all the
observations of that frame in file order.  Sensor fields:
  RelativePoses_2D : X Y Z YAW PITCH ROLL QR QX QY QZ -> (x, y, yaw)   CDatasetParser_RelGraphSLAM2D.h:30-43
  Cartesian_3D     : X Y Z                                              CDatasetParser_Cartesian_3D.h:21-37
  MonocularCamera  : px.x px.y                                          CDatasetParser_Monocular.h:27-42
  StereoCamera     : px_l.x px_l.y px_r.x px_r.y                        CDatasetParser_Stereo.h:27-44
`--add-noise/--noise/--noise-ang` follow the reference's CLI (README.md:65-71); the noise stream is a seeded numpy
generator instead of MRPT's.  `--profile-stats PREFIX` writes the per-keyframe table the reference's CSV dump is built
from (srba-run-generic-impl.h:623-647).  The engine is any object with the RbaEngine.define_new_keyframe() signature
(the CUDA product engine `srba_b200.RbaEngine`, or the CPU oracle mirror in tests).
"""
import argparse
import math
import time

import numpy as np

SENSORS = {"relposes2d": (12, 3), "cartesian3d": (5, 3), "mono": (4, 2), "stereo": (6, 4)}


def load_dataset(path, sensor, graph_slam=False, add_noise=False, noise=None, noise_ang_deg=None, seed=0):
    """Returns the list of per-keyframe dicts that datasets.feed() consumes, plus the noise parameters actually used."""
    ncol, o = SENSORS[sensor]
    rows = []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s[0] in "%#":
                continue
            v = [float(x) for x in s.replace(",", " ").split()]
            if len(v) != ncol:
                raise ValueError("%s: expected %d columns for %s, got %d (checkObsProperSize)" % (path, ncol, sensor, len(v)))
            rows.append(v)
    if len(rows) <= 2:
        raise ValueError("%s: dataset too short (ASSERT_ABOVE_(rows,2), CDatasetParserBase.h:96)" % path)
    M = np.array(rows)
    rng = np.random.default_rng(seed)
    if sensor == "relposes2d":
        sxy = 0.10 if noise is None else noise                       # CDatasetParser_RelGraphSLAM2D.h:22-27
        syaw = math.radians(4.0 if noise_ang_deg is None else noise_ang_deg)
        obs = M[:, [2, 3, 5]].copy()
        if add_noise:
            obs += np.c_[rng.normal(0, sxy, len(M)), rng.normal(0, sxy, len(M)), rng.normal(0, syaw, len(M))]
        params = dict(sigma_xy=sxy, sigma_yaw=syaw)
    else:
        std = {"cartesian3d": 1e-3, "mono": 1e-4, "stereo": 1e-4}[sensor] if noise is None else noise
        obs = M[:, 2:2 + o].copy()
        if add_noise:
            obs += rng.normal(0, std, obs.shape)
        params = dict(noise_std=std)
    frames, i, n = [], 0, len(M)
    if np.any(np.diff(M[:, 0]) < 0):
        raise ValueError("%s: frame ids must be non-decreasing" % path)
    for kf in range(int(M[-1, 0]) + 1):   # a keyframe without rows still exists in --graph-slam mode (keyframe 0: only its fake landmark)
        j = i
        while j < n and int(M[j, 0]) == kf:
            j += 1
        ids, ob, fixed = M[i:j, 1].astype(np.uint64).tolist(), obs[i:j].tolist(), [0] * (j - i)
        if graph_slam:
            ids, ob, fixed = [kf] + ids, [[0.0] * o] + ob, [1] + fixed
        if not ids:
            raise ValueError("%s: keyframe %d has no observations (ASSERT_(!new_obs_in_this_frame.empty()), srba-run-generic-impl.h:446)" % (path, kf))
        m = len(ids)
        frames.append(dict(feat_id=np.array(ids, np.uint64), obs=np.array(ob, np.float64).reshape(m, o), is_fixed=np.array(fixed, np.uint8),
                           is_unknown_with_init_val=np.zeros(m, np.uint8), feat_rel_pos=np.zeros((m, 3))))
        i = j
    return frames, params


def write_dataset(path, frames, sensor, graph_slam=False):
    """Inverse of load_dataset (no noise): lets the synthetic generators of datasets.py feed the reference's own srba-slam."""
    with open(path, "w") as f:
        f.write("%% FRAME_ID FEAT_ID %s\n" % {"relposes2d": "X Y Z YAW PITCH ROLL QR QX QY QZ", "cartesian3d": "X Y Z", "mono": "PX PY",
                                               "stereo": "PXL PYL PXR PYR"}[sensor])
        for k, fr in enumerate(frames):
            for i in range(len(fr["feat_id"])):
                if graph_slam and fr["is_fixed"][i]:
                    continue
                o = fr["obs"][i]
                if sensor == "relposes2d":
                    c, s = math.cos(o[2] / 2), math.sin(o[2] / 2)
                    vals = [o[0], o[1], 0.0, o[2], 0.0, 0.0, c, 0.0, 0.0, s]
                else:
                    vals = list(o)
                f.write("%d %d %s\n" % (k, int(fr["feat_id"][i]), " ".join(repr(float(x)) for x in vals)))


def load_sensor_params(path, sensor):
    """--sensor-params-cfg-file (apps/srba-slam/instance_se3_lm3d_monocular.cpp:24-28, instance_se3_lm3d_stereo.cpp:24-31): an MRPT
    .ini file; monocular: section [CAMERA] with cx, cy, fx, fy (mrpt::utils::TCamera::loadFromConfigFile); stereo: sections
    [CAMERA_LEFT], [CAMERA_RIGHT] and [CAMERA_LEFT2RIGHT_POSE] with pose_quaternion=[x y z qr qx qy qz] (TStereoCamera)."""
    import configparser
    cp = configparser.ConfigParser(inline_comment_prefixes=("//", ";", "#"), strict=False)
    cp.read_string(open(path).read())
    sec = {k.upper(): v for k, v in cp.items()}

    def cam(name):
        if name not in sec:
            raise ValueError("section [%s] missing in %s" % (name, path))
        c = sec[name]
        return [float(c["fx"]), float(c["fy"]), float(c["cx"]), float(c["cy"])]

    def vec(txt):
        return [float(x) for x in txt.replace("[", " ").replace("]", " ").replace(",", " ").split()]
    if sensor == "mono":
        return dict(cam_left=cam("CAMERA"))
    pose = vec(sec["CAMERA_LEFT2RIGHT_POSE"]["pose_quaternion"])
    if len(pose) != 7 or pose[0] == 0:
        raise ValueError("ASSERT_(baseline!=0) (instance_se3_lm3d_stereo.cpp:30)")
    return dict(cam_left=cam("CAMERA_LEFT"), cam_right=cam("CAMERA_RIGHT"), right_cam_pose=pose)


def run_slam(engine, frames, profile_prefix=None, verbose=0, profile_stats_length=10):
    """One define_new_keyframe() per frame (srba-run-generic-impl.h:451-465); returns per-keyframe rows and a summary.
    profile_prefix: the two files of --profile-stats (srba-run-generic-impl.h:201-202,623-647,778-829): <prefix>_means.csv and
    <prefix>_calls.csv, one row per segment of `profile_stats_length` keyframes, tab-separated, columns = timed sections (mean seconds
    per call within the segment / number of calls) plus total_num_kfs / total_num_lms / total_num_obs; the timers restart after
    every segment.  Sections of this implementation: define_new_keyframe (wall), opt = the optimize_edges span (flatten + GPU call +
    scatter, impl/optimize_edges.h:56,786), opt.gpu_call = inside srba_b200_optimize.  A per-keyframe table goes to
    <prefix>_keyframes.csv."""
    rows, t_all = [], time.perf_counter()
    seg_ts, seg_means, seg_calls = [], {}, {}
    acc = dict(define_new_keyframe=[0.0, 0], opt=[0.0, 0], **{"opt.gpu_call": [0.0, 0]})
    timers = getattr(engine, "timers", None) or (lambda: dict(gpu_seconds=0.0, flatten_seconds=0.0))   # (the oracle engine of the CPU tests has no GPU timers)
    tm_prev = timers()
    for k, fr in enumerate(frames):
        t0 = time.perf_counter()
        r = engine.define_new_keyframe(fr["feat_id"], fr["obs"], is_fixed=fr["is_fixed"], is_unknown_with_init_val=fr["is_unknown_with_init_val"],
                                       feat_rel_pos=fr["feat_rel_pos"], run_local_optimization=True)
        dt = time.perf_counter() - t0
        o = r["optimize_results"]
        rows.append((k, dt, len(r["created_edges"]), o["num_observations"], o["num_kf2kf_edges_optimized"], o["num_kf2lm_edges_optimized"],
                     o.get("num_lm_trials", 0), o["obs_rmse"]))
        if verbose:
            print("KF %d: %d new edges, %d obs, rmse %.3g, %.2f ms" % (k, rows[-1][2], rows[-1][3], rows[-1][7], 1e3 * dt))
        if profile_prefix:
            tm = timers()
            n_opt = (1 if o["num_observations"] else 0) + (1 if r["optimize_results_stg1"]["num_observations"] else 0)
            acc["define_new_keyframe"][0] += dt; acc["define_new_keyframe"][1] += 1
            acc["opt"][0] += (tm["gpu_seconds"] - tm_prev["gpu_seconds"]) + (tm["flatten_seconds"] - tm_prev["flatten_seconds"]); acc["opt"][1] += n_opt
            acc["opt.gpu_call"][0] += tm["gpu_seconds"] - tm_prev["gpu_seconds"]; acc["opt.gpu_call"][1] += n_opt
            tm_prev = tm
            if k % max(1, profile_stats_length) == 0:   # (curFrameIdx % SAVE_TIMING_SEGMENT_LENGTH) == 0  (:625)
                seg_ts.append(k)
                for name, (sec, n) in acc.items():
                    seg_means.setdefault(name, []).append(sec / n if n else 0.0)
                    seg_calls.setdefault(name, []).append(n)
                    acc[name] = [0.0, 0]
                seg_means.setdefault("total_num_kfs", []).append(engine.num_keyframes)
                seg_means.setdefault("total_num_lms", []).append(len(engine.unknown_lms()[0]))   # landmarks with unknown relative position
                seg_means.setdefault("total_num_obs", []).append(engine.num_observations)
    total = time.perf_counter() - t_all
    a = np.array(rows, float)
    summary = dict(keyframes=len(rows), seconds=total, mean_kf_ms=1e3 * a[:, 1].mean(), lm_trials=int(a[:, 6].sum()), final_rmse=float(a[-1, 7]),
                   mean_rmse=float(np.nanmean(a[:, 7])))
    if profile_prefix:
        np.savetxt(profile_prefix + "_keyframes.csv", a, delimiter=",", header="kf,define_new_keyframe_s,new_edges,obs,k2k_opt,lm_opt,lm_trials,obs_rmse",
                   comments="")
        for suffix, table in (("_means.csv", seg_means), ("_calls.csv", seg_calls)):
            names = sorted(table)
            with open(profile_prefix + suffix, "w") as f:
                f.write("timestep" + "".join('\t "%s"' % n for n in names) + "\n")
                for i, ts in enumerate(seg_ts):
                    f.write(str(ts) + "".join("\t%g" % table[n][i] for n in names) + "\n")
    return rows, summary


def main(argv=None):
    ap = argparse.ArgumentParser(description="srba-slam front end on the B200 hot path")
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument("--se2", action="store_true")
    g.add_argument("--se3", action="store_true")
    ap.add_argument("--graph-slam", action="store_true")
    ap.add_argument("--lm-3d", action="store_true")
    ap.add_argument("--obs", default=None, help="Cartesian_3D | MonocularCamera | StereoCamera (relative graph-SLAM implies RelativePoses_2D)")
    ap.add_argument("-d", "--dataset", required=True)
    ap.add_argument("--submap-size", type=int, default=20)
    ap.add_argument("--max-spanning-tree-depth", type=int, default=4)
    ap.add_argument("--max-optimize-depth", type=int, default=4)
    ap.add_argument("--min-obs-to-loop-closure", type=int, default=None)
    ap.add_argument("--max-iters", type=int, default=20)
    ap.add_argument("--noise", type=float, default=None)
    ap.add_argument("--noise-ang", type=float, default=None)
    ap.add_argument("--add-noise", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--profile-stats", default=None)
    ap.add_argument("--profile-stats-length", type=int, default=10, help="Length in KFs of each saved profiled segment")
    ap.add_argument("--sensor-params-cfg-file", default=None, help="Config file from where to load the sensor parameters (mandatory for camera observations)")
    ap.add_argument("--verbose", type=int, default=0)
    a = ap.parse_args(argv)
    import srba_b200 as sb
    from . import datasets as D
    if a.se2:
        if not a.graph_slam:
            ap.error("--se2 is built with --graph-slam (RelativePoses_2D) only")
        # solver_t of the app's SE(2) instance: solver_LM_no_schur_sparse_cholesky (apps/srba-slam/instance_relative_graph_slam_se2.cpp:23)
        sensor, kw = "relposes2d", dict(pose_kind=0, obs_kind=0, noise_kind=1, solver_kind=sb.SOLVER_NO_SCHUR_SPARSE)
    else:
        sensor = {"cartesian_3d": "cartesian3d", "monocularcamera": "mono", "stereocamera": "stereo"}[(a.obs or "Cartesian_3D").lower()]
        kw = dict(pose_kind=1, obs_kind={"cartesian3d": 1, "mono": 2, "stereo": 3}[sensor], noise_kind=0)
        if sensor in ("mono", "stereo"):
            # the app's camera instances: sensor_pose_on_robot_se3 with the camera pointing forwards (camera +Z = robot +X),
            # apps/srba-slam/srba-run-generic-impl.h:72-80, instance_se3_lm3d_monocular.cpp:38, instance_se3_lm3d_stereo.cpp:41; calibration from
            # the mandatory --sensor-params-cfg-file (instance_se3_lm3d_monocular.cpp:24-28, instance_se3_lm3d_stereo.cpp:24-31)
            if not a.sensor_params_cfg_file:
                ap.error("--sensor-params-cfg-file is mandatory for this type of observations.")
            kw.update(sensor_pose_kind=1, sensor_pose=[0, 0, 0, np.deg2rad(-90.0), 0.0, np.deg2rad(-90.0)], **load_sensor_params(a.sensor_params_cfg_file, sensor))
    frames, npar = load_dataset(a.dataset, sensor, graph_slam=a.graph_slam, add_noise=a.add_noise, noise=a.noise, noise_ang_deg=a.noise_ang, seed=a.seed)
    if sensor == "relposes2d":
        kw["noise_lambda"] = D.se2_noise_lambda(npar["sigma_xy"], npar["sigma_yaw"])
    else:
        kw["noise_std"] = npar["noise_std"]
    kw.update(submap_size=a.submap_size, max_tree_depth=a.max_spanning_tree_depth, max_optimize_depth=a.max_optimize_depth, max_iters=a.max_iters,
              min_obs_to_loop_closure=a.min_obs_to_loop_closure if a.min_obs_to_loop_closure is not None else (1 if a.graph_slam else 4),
              max_error_per_obs_to_stop=1e-8)   # srba-run-generic-impl.h:127,178-182
    eng = sb.RbaEngine(sb.default_engine_params(**kw))
    _, summary = run_slam(eng, frames, a.profile_stats, a.verbose, a.profile_stats_length)
    print("%(keyframes)d keyframes in %(seconds).2f s (%(mean_kf_ms).3f ms per define_new_keyframe), %(lm_trials)d LM trials, final obs RMSE %(final_rmse).4g" % summary)
    return summary


if __name__ == "__main__":
    main()
