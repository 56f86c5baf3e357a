#!/usr/bin/env python
"""Extracts the literal test DATA (numeric arrays + ground-truth poses quoted in comments) that the reference's
own tests/tutorials embed for this path, into tests/golden/reference_fixtures.json.  Data only: no code is copied.
Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/extract_fixtures.py
Sources:
  tests/sensor-pose_unittest.cpp:31-173, tests/fixed-transformations_unittest.cpp:39-44,133-165,
  tests/submaps_edge_init_values.cpp:61-79, examples/cpp/tutorial-srba-{cartesian3d,monocular,stereo}-se3.cpp,
  examples/cpp/tutorial-srba-relative-graph-slam-se2.cpp
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")

NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)"


def arrays(path):
    """All `type name[] = { {..}, {..} };` numeric tables in a C++ file, with the comment line preceding each."""
    src = open(os.path.join(REF, path)).read()
    out = {}
    for m in re.finditer(r"(?:(//[^\n]*)\n)?\s*(?:const\s+)?[\w:]+\s+(\w+)\s*\[\]\s*(?:\[\d*\])?\s*=\s*\{(.*?)\};", src, re.S):
        comment, name, body = m.group(1), m.group(2), m.group(3)
        rows = []
        for r in re.finditer(r"\{([^{}]*)\}", body):
            txt = re.sub(r"DEG2RAD\(\s*(" + NUM + r")\s*\)", lambda g: repr(float(g.group(1)) * 3.141592653589793 / 180.0), r.group(1))
            vals = [float(v) for v in re.findall(NUM, txt)]
            if vals:
                rows.append(vals)
        if rows:
            out[name] = {"rows": rows, "comment": comment or ""}
    return out


def gt_from_comment(comment):
    vals = [float(v) for v in re.findall(NUM, comment.split(":")[-1].split("=")[-1])]
    return vals[-7:]


def main():
    fx = {}
    a = arrays("tests/sensor-pose_unittest.cpp")
    fx["sensor_pose_unittest"] = {k: a[k]["rows"] for k in ("observations_0", "observations_10", "observations_0_displ", "observations_10_displ")}
    fx["sensor_pose_unittest"]["GT0_xyzq"] = [0, 0, 0, 0.995247, 0.003802, -0.043453, 0.087073]
    fx["sensor_pose_unittest"]["GT1_xyzq"] = [1.226072, 0.293638, 0.110099, 0.999103, 0.000160, -0.042322, 0.001049]
    fx["sensor_pose_unittest"]["sensor_pose_xyzypr"] = [0.1, 0.2, 0.3, 0.1745329251994328, 0.3490658503988660, 0.5235987755982987]
    a = arrays("tests/fixed-transformations_unittest.cpp")
    fx["fixed_transformations"] = {"dummy_obs": a["dummy_obs"]["rows"], "test_fixed_transfs": a["test_fixed_transfs"]["rows"]}
    a = arrays("tests/submaps_edge_init_values.cpp")
    fx["submaps_edge_init_values"] = {"dataset": a["dataset"]["rows"]}
    for key, path in (("tutorial_cartesian3d", "examples/cpp/tutorial-srba-cartesian3d-se3.cpp"),
                      ("tutorial_monocular", "examples/cpp/tutorial-srba-monocular-se3.cpp"),
                      ("tutorial_stereo", "examples/cpp/tutorial-srba-stereo-se3.cpp"),
                      ("tutorial_relgraphslam_se2", "examples/cpp/tutorial-srba-relative-graph-slam-se2.cpp")):
        a = arrays(path)
        fx[key] = {k: {"rows": v["rows"], "comment": v["comment"]} for k, v in a.items()}
    # Ground truth + calibration literals quoted by the tutorials (robot-frame GT: tutorial-srba-cartesian3d-se3.cpp:44,68;
    # camera-frame GT: tutorial-srba-stereo-se3.cpp:41,77; calibrations: tutorial-srba-monocular-se3.cpp:125-135,
    # tutorial-srba-stereo-se3.cpp:135-147).
    fx["tutorial_gt"] = {
        "robot_pose0_xyzq": [0, 0, 0, 0.995247, 0.003802, -0.043453, 0.087073],
        "robot_pose10_xyzq": [1.226071, 0.293637, 0.110099, 0.999103, 0.000160, -0.042322, 0.001049],
        "camera_pose0_xyzq": [0, 0, 0, 0.564787, -0.517532, 0.434261, -0.473913],
        "camera_pose10_xyzq": [1.226071, 0.293637, 0.110099, 0.521317, -0.478835, 0.477946, -0.520108],
        "sensor_pose_xyzypr_deg": [0, 0, 0, -90, 0, -90],
        "mono_cam_fx_fy_cx_cy": [200, 200, 400, 320],
        "stereo_cam_fx_fy_cx_cy": [200, 150, 512, 384],
        "stereo_right_cam_pose_xyzq": [0.2, 0, 0, 1, 0, 0, 0],
    }
    json.dump(fx, open(OUT, "w"), indent=0)
    print("wrote", OUT, {k: list(v.keys()) for k, v in fx.items()})


if __name__ == "__main__":
    main()
