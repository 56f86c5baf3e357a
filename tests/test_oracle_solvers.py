"""Pins the oracle's restatement of the three RBA_OPTIONS::solver_t variants (srba_options_solver.h:21-77,
impl/lev-marq_solvers.h:27-209 no Schur + sparse Cholesky, :213-405 Schur + sparse Cholesky, :409-591 Schur + dense Cholesky)
against a naive dense solve of the full system, on the random shapes of the reference's tests/schur_unittest.cpp:285-360,
and checks the one behavioural difference between them: the Schur variants mutate the caller's gradient in place
(impl/schur.h:248-265,283-296), the no-Schur variant leaves it untouched."""
import numpy as np
import pytest

from oracle import pyoracle as po

KINDS = [po.SOLVER_SCHUR_DENSE, po.SOLVER_SCHUR_SPARSE, po.SOLVER_NO_SCHUR_SPARSE]


def make_case(n_p, n_f, prob, seed):
    rng = np.random.default_rng(seed)
    visible = rng.uniform(size=(n_p + 1, n_f)) <= prob
    for lm in range(n_f):
        if not visible[:, lm].any():
            visible[0, lm] = True
    for kf in range(1, n_p + 1):
        if not visible[kf].any():
            visible[kf, 0] = True
    blocksA, blocksF, rowsA, rowsF = [], [], [], []
    obs = 0
    for kf in range(n_p + 1):
        for lm in range(n_f):
            if not visible[kf, lm]:
                continue
            if kf > 0:
                JA = rng.standard_normal((3, 6))
                blocksA.append(JA)
                rowsA.append((obs, kf - 1, JA))
            JF = rng.standard_normal((3, 3))
            blocksF.append(JF)
            rowsF.append((obs, lm, JF))
            obs += 1
    J = np.zeros((3 * obs, 6 * n_p + 3 * n_f))
    for o, c, B in rowsA:
        J[3 * o:3 * o + 3, 6 * c:6 * c + 6] = B
    for o, c, B in rowsF:
        J[3 * o:3 * o + 3, 6 * n_p + 3 * c:6 * n_p + 3 * c + 3] = B
    return visible.astype(np.uint8), np.array(blocksA).reshape(-1), np.array(blocksF).reshape(-1), J


@pytest.mark.parametrize("n_p,n_f,prob", [(1, 7, 1.0), (1, 7, 0.95), (2, 20, 0.9), (5, 30, 0.9), (8, 40, 0.5)])
@pytest.mark.parametrize("lam", [1e3, 1e-2])
def test_three_solvers_match_dense_solve(n_p, n_f, prob, lam):
    for seed in range(1, 4):
        vis, JA, JF, J = make_case(n_p, n_f, prob, seed)
        n = J.shape[1]
        ref = np.linalg.solve(J.T @ J + lam * np.eye(n), np.ones(n))
        for kind in KINDS:
            ok, d, g = po.solver_case(kind, n_p, n_f, vis, JA, JF, lam)
            assert ok
            assert np.abs(d - ref).max() <= 1e-9 * np.abs(ref).max(), (kind, seed)
            if kind == po.SOLVER_NO_SCHUR_SPARSE:
                assert np.array_equal(g, np.ones(n))  # lev-marq_solvers.h:182: backsub(minus_grad, delta_eps), gradient untouched
            else:
                assert not np.array_equal(g[:6 * n_p], np.ones(6 * n_p))  # schur.h:248-265: reduced in place


def test_not_positive_definite_is_reported_by_all_solvers():
    """lambda < 0 beyond the smallest eigenvalue: Eigen::LLT info()!=Success (:530-534) / CExceptionNotDefPos (:171-175,347-351)."""
    vis, JA, JF, J = make_case(3, 12, 0.9, 5)
    lam = -1.05 * np.linalg.eigvalsh(J.T @ J).max()
    for kind in KINDS:
        ok, _, _ = po.solver_case(kind, 3, 12, vis, JA, JF, lam)
        assert not ok, kind


def test_engine_runs_with_every_solver_and_agrees():
    """End to end on a small SE(3) landmark map: the three variants walk the same LM iterations up to rounding while no step is
    rejected (after a rejection the Schur variants re-solve with their mutated gradient) and reach the same optimum."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from srba_b200 import datasets as D
    ds = D.landmark_slam_se3(n_kfs=25, n_lms=250, seed=3, obs_kind="cartesian3d")
    res = []
    for kind in KINDS:
        e = po.OracleEngine(po.default_params(pose_kind=po.POSE_SE3, obs_kind=po.OBS_CARTESIAN3D, noise_std=1e-3, max_tree_depth=3, max_optimize_depth=3,
                                              submap_size=10, min_obs_to_loop_closure=4, solver_kind=kind))
        for f in ds["frames"]:
            D.feed(e, f)
        _, poses = e.edges()
        _, _, lms = e.unknown_lms()
        res.append((np.array(poses), np.array(lms)))
    for poses, lms in res[1:]:
        assert np.abs(poses - res[0][0]).max() <= 1e-6
        assert np.abs(lms - res[0][1]).max() <= 1e-6
