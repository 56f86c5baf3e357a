"""Pins the oracle's Hessian build + Schur complement against the naive dense computation, exactly as the
reference's own unit test does (tests/schur_unittest.cpp:71-279, cases :285-360, tolerance 1e-10 at :258,267)."""
import numpy as np
import pytest

from oracle import pyoracle as po

LAMBDA = 1e3  # schur_unittest.cpp:79


def run_case(n_p, n_f, visible, rng):
    visible = np.asarray(visible, dtype=np.uint8).reshape(n_p + 1, n_f)
    blocksA, blocksF = [], []
    rowsA, rowsF = [], []  # dense assembly
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
    dA = np.zeros((3 * obs, 6 * n_p))
    dF = np.zeros((3 * obs, 3 * n_f))
    for o, c, J in rowsA:
        dA[3 * o:3 * o + 3, 6 * c:6 * c + 6] = J
    for o, c, J in rowsF:
        dF[3 * o:3 * o + 3, 3 * c:3 * c + 3] = J
    HAp, Hf, HApf = dA.T @ dA, dF.T @ dF, dA.T @ dF
    Hf_inv = np.linalg.inv(Hf + LAMBDA * np.eye(3 * n_f))
    dense_schur = HAp - HApf @ Hf_inv @ HApf.T
    g = np.ones(6 * n_p + 3 * n_f)
    dense_g = g[:6 * n_p] - HApf @ Hf_inv @ g[6 * n_p:]
    H, gr, ninv = po.schur_case(n_p, n_f, visible, np.array(blocksA).reshape(-1), np.array(blocksF).reshape(-1), LAMBDA)
    assert ninv == n_f
    assert np.abs(dense_schur - H).max() / np.abs(dense_schur).max() < 1e-10
    assert np.abs(dense_g - gr).max() / np.abs(dense_g).max() < 1e-10


@pytest.mark.parametrize("n_p,n_f,prob,seeds", [
    (1, 6, 1.0, range(1, 5)), (1, 7, 1.0, range(1, 15)), (1, 7, 0.95, range(1, 15)), (1, 30, 1.0, range(1, 15)),
    (2, 20, 0.9, range(1, 5)), (5, 30, 0.9, range(1, 5))])
def test_dense_vs_sparse_random(n_p, n_f, prob, seeds):
    for seed in seeds:
        rng = np.random.default_rng(seed)
        visible = rng.uniform(size=(n_p + 1, n_f)) <= prob
        # every landmark must be seen at least once so that every dh_df column is non-empty (reference asserts this)
        for lm in range(n_f):
            if not visible[:, lm].any():
                visible[0, lm] = True
        for kf in range(1, n_p + 1):
            if not visible[kf].any():
                visible[kf, 0] = True
        run_case(n_p, n_f, visible, rng)


def test_dense_vs_sparse_manual_masks():
    rng = np.random.default_rng(7)
    T, F = True, False
    run_case(2, 5, [T, T, T, T, F, T, T, T, F, T, T, T, F, T, T], rng)        # schur_unittest.cpp:334-341
    run_case(2, 6, [F, T, F, T, F, T, T, F, T, F, T, F, T, T, T, T, T, T], rng)  # :342-350
