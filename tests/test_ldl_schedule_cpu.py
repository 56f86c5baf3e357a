"""CPU-only check of the packer's symbolic factorisation (srba_b200/csrc/capi.cpp: symbolic_tiles), i.e. the schedule the
on-chip block-sparse L D L^t solver of lm_kernel.cu executes instead of the reference's dense fill + Eigen::LLT
(impl/lev-marq_solvers.h:492-549).  The schedule is emulated step by step in numpy on a random SPD system with the
problem's own block pattern and compared with a dense solve: this pins the fill-in, the tile numbering, the update lists,
the implicit pivot update and the source mapping without a GPU."""
import ctypes as C

import numpy as np
import pytest

import srba_b200 as sb
from srba_b200 import _capi as K
from srba_b200 import datasets as D


def schedule(flat, max_on_chip=32767):
    L = K.lib()
    counts = (C.c_int32 * 6)()
    assert L.srba_b200_debug_ldl_schedule(flat.problem, counts, None, None, None, None, None, None, 0, 0, max_on_chip) == 0
    ntr, ntiles, nupd, ok, nph, fits = list(counts)
    assert ok == 1
    row = np.zeros(4 * (ntr + 1), np.int32); tij = np.zeros(ntiles, np.int32); src = np.zeros(ntiles, np.int32); upd = np.zeros(2 * nupd, np.int32)
    orow = np.zeros(ntr, np.int32)
    phase = np.zeros(16 * (nph + 1), np.int32)
    p32 = C.POINTER(C.c_int32)
    assert L.srba_b200_debug_ldl_schedule(flat.problem, counts, row.ctypes.data_as(p32), tij.ctypes.data_as(p32), src.ctypes.data_as(p32),
                                          upd.ctypes.data_as(p32), orow.ctypes.data_as(p32), phase.ctypes.data_as(p32), ntiles, nupd, max_on_chip) == 0
    return ntr, row.reshape(-1, 4), tij, src, upd.reshape(-1, 2), orow, phase.reshape(-1, 16)[:nph]


def emulate(flat, P, rng, max_on_chip=32767):
    """Random SPD blocks with the problem's reduced-system pattern -> solve through the schedule and densely."""
    p = flat.p
    n_p, n = p.n_unk_edges, P * p.n_unk_edges
    S_row, S_col = flat.array("S_row", p.n_S, np.int32), flat.array("S_col", p.n_S, np.int32)
    blocks = rng.normal(size=(p.n_S, P, P))
    A = np.zeros((n, n))
    for b in range(p.n_S):
        r, c = S_row[b], S_col[b]
        if r == c:
            blocks[b] = blocks[b] @ blocks[b].T + 4.0 * P * np.eye(P)   # diagonally dominant: the whole matrix is SPD
        A[P * r:P * r + P, P * c:P * c + P] = blocks[b]
        A[P * c:P * c + P, P * r:P * r + P] = blocks[b].T
    for b in range(p.n_S):
        if S_row[b] == S_col[b]:
            A[P * S_row[b]:P * S_row[b] + P, P * S_row[b]:P * S_row[b] + P] += 0.3 * n_p * np.eye(P)
            blocks[b] += 0.3 * n_p * np.eye(P)
    g = rng.normal(size=n)
    lam = 0.37
    Sflat = blocks.reshape(-1)
    ntr, row, tij, src, upd, orow, phase = schedule(flat, max_on_chip)
    assert sorted(orow.tolist()) == list(range(ntr))
    assert ntr == n // 3
    tiles = np.zeros((len(tij), 3, 3))
    for t in range(len(tij)):
        ti, tj = tij[t] >> 16, tij[t] & 0xffff
        assert ti <= tj <= ntr
        if src[t] >= 0:
            off = src[t] & 0x1fffffff
            for a in range(3):
                tiles[t, a] = Sflat[off + a * P: off + a * P + 3]
            if (src[t] >> 29) & 1:
                tiles[t] = tiles[t].T.copy()
            assert bool(src[t] >> 30) == (ti == tj)
            if ti == tj:
                tiles[t] += lam * np.eye(3)
        elif src[t] <= -2:
            assert tj == ntr and -src[t] - 2 == orow[ti]
            tiles[t, :, 0] = g[3 * orow[ti]:3 * orow[ti] + 3]
    # rows are contiguous, diagonal first, right-hand side last
    for k in range(ntr):
        r0, r1 = row[k, 0], row[k, 1] & 0x3fffffff
        cols = [tij[t] & 0xffff for t in range(r0, r1)]
        assert cols[0] == k and cols[-1] == ntr and cols == sorted(cols) and all(tij[t] >> 16 == k for t in range(r0, r1))
    # phases: 1 or 2 independent rows eliminated together; the pivots of the next phase are updated through their link tile
    inv = [None] * ntr
    assert phase[0, 0] == 0 and sum(int(ph[1]) for ph in phase) == ntr
    for r in range(phase[0, 0], phase[0, 0] + phase[0, 1]):
        inv[r] = np.linalg.inv(tiles[row[r, 0]])
    done_rows = 0
    for m, ph in enumerate(phase):
        row0, nrows, u0, umid, u1, nnext, piv0, piv1, link0, link1, t0, tmid, t1 = [int(x) for x in ph[:13]]
        assert tmid == (row[row0, 1] & 0x3fffffff)
        assert row0 == done_rows and nrows in (1, 2) and u0 == row[row0, 2] and u1 == row[row0 + nrows - 1, 3] and umid == row[row0, 3]
        assert t0 == row[row0, 0] and t1 == (row[row0 + nrows - 1, 1] & 0x3fffffff)
        touched = set()
        for u in range(u0, u1):
            k = row0 if u < umid else row0 + 1
            r0, r1 = row[k, 0], row[k, 1] & 0x3fffffff
            dst, ti_, tj_ = upd[u, 0] & 0xffff, upd[u, 0] >> 16, upd[u, 1]
            assert r0 < ti_ < r1 and r0 < tj_ < r1 and (tij[dst] >> 16) == (tij[ti_] & 0xffff) and (tij[dst] & 0xffff) == (tij[tj_] & 0xffff)
            assert dst not in touched and dst not in (piv0 if nnext > 0 else -1, piv1 if nnext > 1 else -1)   # one writer per tile and phase
            touched.add(dst)
            tiles[dst] -= tiles[ti_].T @ inv[k] @ tiles[tj_]
        for gi in range(nnext):
            q = row0 + nrows + gi
            piv, link = (piv0, link0) if gi == 0 else (piv1, link1)
            assert piv == row[q, 0]
            srcs = [r for r in range(row0, row0 + nrows) if any((tij[t] & 0xffff) == q for t in range(row[r, 0] + 1, row[r, 1] & 0x3fffffff))]
            assert len(srcs) <= 1 and (link >= 0) == (len(srcs) == 1)
            if link >= 0:
                second, link = link >> 30, link & 0x3fffffff
                assert (tij[link] >> 16) == srcs[0] == row0 + second and (tij[link] & 0xffff) == q
                tiles[piv] -= tiles[link].T @ inv[tij[link] >> 16] @ tiles[link]
            inv[q] = np.linalg.inv(tiles[piv])
        done_rows += nrows
    assert done_rows == ntr
    for k in range(ntr):   # W = inv(D) T in place
        for t in range(row[k, 0] + 1, row[k, 1] & 0x3fffffff):
            tiles[t] = inv[k] @ tiles[t]
    x = np.zeros((ntr, 3))
    for k in range(ntr - 1, -1, -1):
        r0, r1 = row[k, 0], row[k, 1] & 0x3fffffff
        acc = tiles[r1 - 1][:, 0].copy()
        for t in range(r0 + 1, r1 - 1):
            acc -= tiles[t] @ x[tij[t] & 0xffff]
        x[k] = acc
    ref = np.linalg.solve(A + lam * np.eye(n), g)
    xo = np.zeros(n)
    for t in range(ntr):
        xo[3 * orow[t]:3 * orow[t] + 3] = x[t]   # back to the caller's unknown order
    return xo, ref, len(tij), p.n_S, orow, len(phase)


def flats_se2():
    ds = D.graphslam_se2_tour(90, seed=3)
    kw = dict(pose_kind=0, obs_kind=0, noise_kind=1, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]), max_tree_depth=3,
              max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1)
    e = sb.RbaEngine(sb.default_engine_params(**kw))
    out = []
    for k, f in enumerate(ds["frames"]):
        D.feed(e, f, run_local_optimization=False)
        if k in (5, 33, 60, 89):
            out.append(e.flatten_local_area(k, 3))
    return out


def flats_se3():
    ds = D.landmark_slam_se3(n_kfs=25, n_lms=150, seed=5, obs_kind="cartesian3d")
    kw = dict(pose_kind=1, obs_kind=1, noise_kind=0, noise_std=1e-3, max_tree_depth=3, max_optimize_depth=3, submap_size=5, min_obs_to_loop_closure=4)
    e = sb.RbaEngine(sb.default_engine_params(**kw))
    out = []
    for k, f in enumerate(ds["frames"]):
        D.feed(e, f, run_local_optimization=False)
        if k in (8, 24):
            out.append(e.flatten_local_area(k, 3))
    return out


@pytest.mark.parametrize("kind", ["se2", "se3"])
def test_schedule_solves_the_reduced_system(kind):
    rng = np.random.default_rng(11)
    flats, P = (flats_se2(), 3) if kind == "se2" else (flats_se3(), 6)
    assert flats
    for fl in flats:
        x, ref, ntiles, n_S, orow, nph = emulate(fl, P, rng)
        assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
        if kind == "se3":
            assert orow.tolist() == list(range(len(orow)))   # the reference's own order is kept when its fill fits on chip
        # a tiny on-chip budget forces the minimum-degree re-ordering: same solution, tiles from transposed blocks included
        if kind == "se2" and len(orow) >= 16:
            assert nph <= 0.8 * len(orow)   # chain-like areas are eliminated from both ends: two block rows per phase until the fronts meet
        x2, ref2, ntiles2, _, orow2, _ = emulate(fl, P, rng, max_on_chip=8)
        assert np.abs(x2 - ref2).max() <= 1e-10 * max(1.0, np.abs(ref2).max())


def test_plan_fits_the_shared_memory_budget():
    L = K.lib()
    for fl in flats_se2():
        out = (C.c_int64 * 8)()
        assert L.srba_b200_problem_plan(fl.problem, out) == 0
        img, ws, res, sm, chol, n_img, n_ws, budget = list(out)
        assert 0 < sm <= budget and chol == 0 and n_ws >= 10 and img > 0 and res > 0


# ------------------------------------------------------------------------------------------------ level schedule (wave pipeline)
def level_schedule(flat):
    L = K.lib()
    counts = (C.c_int32 * 6)()
    assert L.srba_b200_debug_level_schedule(flat.problem, counts, None, None, None, None, None, None) == 0
    ntr, ntiles, nlev, ndst, ncon, ok = list(counts)
    assert ok == 1
    row2 = np.zeros(2 * (ntr + 1), np.int32); tij = np.zeros(ntiles, np.int32); src = np.zeros(ntiles, np.int32); orow = np.zeros(ntr, np.int32)
    lv = np.zeros(2 * (nlev + 1), np.int32); gu = np.zeros(ndst + ndst + 1 + 2 * ncon, np.int32)
    p32 = C.POINTER(C.c_int32)
    assert L.srba_b200_debug_level_schedule(flat.problem, counts, *[a.ctypes.data_as(p32) for a in (row2, tij, src, orow, lv, gu)]) == 0
    return ntr, row2.reshape(-1, 2), tij, src, orow, lv.reshape(-1, 2), gu[:ndst], gu[ndst:2 * ndst + 1], gu[2 * ndst + 1:].reshape(-1, 2)


def emulate_levels(flat, P, rng):
    """The level schedule executed step by step in numpy: per level A (pivot inverse, W = inv(D) T), then B (grouped destination
    updates T(a,b) -= sum_k T(k,a)^t W(k,b)); backward pass by descending level.  Compared with a dense solve."""
    p = flat.p
    n_p, n = p.n_unk_edges, P * p.n_unk_edges
    S_row, S_col = flat.array("S_row", p.n_S, np.int32), flat.array("S_col", p.n_S, np.int32)
    blocks = rng.normal(size=(p.n_S, P, P))
    A = np.zeros((n, n))
    for b in range(p.n_S):
        r, c = S_row[b], S_col[b]
        if r == c:
            blocks[b] = blocks[b] @ blocks[b].T + (4.0 * P + 0.3 * n_p) * np.eye(P)
        A[P * r:P * r + P, P * c:P * c + P] = blocks[b]
        A[P * c:P * c + P, P * r:P * r + P] = blocks[b].T
    g = rng.normal(size=n)
    lam = 0.37
    Sflat = blocks.reshape(-1)
    ntr, row, tij, src, orow, lv, dst, cptr, cab = level_schedule(flat)
    assert sorted(orow.tolist()) == list(range(ntr)) and ntr == n // 3
    T = np.zeros((len(tij), 3, 3))
    for t in range(len(tij)):
        ti, tj = tij[t] >> 16, tij[t] & 0xffff
        assert ti <= tj <= ntr
        if src[t] >= 0:
            off = src[t] & 0x1fffffff
            for a in range(3):
                T[t, a] = Sflat[off + a * P: off + a * P + 3]
            if (src[t] >> 29) & 1:
                T[t] = T[t].T.copy()
            assert bool(src[t] >> 30) == (ti == tj)
            if ti == tj:
                T[t] += lam * np.eye(3)
        elif src[t] <= -2:
            assert tj == ntr and -src[t] - 2 == orow[ti]
            T[t, :, 0] = g[3 * orow[ti]:3 * orow[ti] + 3]
    for k in range(ntr):   # rows contiguous, diagonal first, right-hand side last
        cols = [tij[t] & 0xffff for t in range(row[k, 0], row[k, 1])]
        assert cols[0] == k and cols[-1] == ntr and cols == sorted(cols) and all(tij[t] >> 16 == k for t in range(row[k, 0], row[k, 1]))
    nlev = len(lv) - 1
    assert lv[0, 0] == 0 and lv[-1, 0] == ntr and lv[-1, 1] == len(dst) and np.all(np.diff(lv[:, 0]) >= 1)
    Wt = np.zeros_like(T)
    for L in range(nlev):
        rows_L = range(lv[L, 0], lv[L + 1, 0])
        for k in rows_L:
            Dinv = np.linalg.inv(T[row[k, 0]])
            assert np.all(np.linalg.eigvalsh(T[row[k, 0]]) > 0)
            for t in range(row[k, 0] + 1, row[k, 1]):
                assert lv[L + 1, 0] <= (tij[t] & 0xffff)   # every column of a row lies in a HIGHER level (or is the right-hand side)
                Wt[t] = Dinv @ T[t]
        seen = set()
        for e in range(lv[L, 1], lv[L + 1, 1]):
            d = dst[e]
            assert d not in seen   # one owner per destination and level
            seen.add(d)
            acc = np.zeros((3, 3))
            last_k = -1
            for c in range(cptr[e], cptr[e + 1]):
                ta, tb = cab[c]
                k = tij[ta] >> 16
                assert k in rows_L and k == (tij[tb] >> 16) and k > last_k   # ascending source row, all of this level
                last_k = k
                assert (tij[d] >> 16) == (tij[ta] & 0xffff) and (tij[d] & 0xffff) == (tij[tb] & 0xffff)
                acc += T[ta].T @ Wt[tb]
            T[d] -= acc
        # completeness: every pair of tiles of every row of the level is somebody's contribution
        want = sum((row[k, 1] - row[k, 0] - 2) * (row[k, 1] - row[k, 0] - 1) // 2 + (row[k, 1] - row[k, 0] - 2) for k in rows_L)
        assert want == cptr[lv[L + 1, 1]] - cptr[lv[L, 1]]
    x = np.zeros((ntr, 3))
    for L in range(nlev - 1, -1, -1):
        for k in range(lv[L, 0], lv[L + 1, 0]):
            acc = Wt[row[k, 1] - 1][:, 0].copy()
            for t in range(row[k, 0] + 1, row[k, 1] - 1):
                acc -= Wt[t] @ x[tij[t] & 0xffff]
            x[k] = acc
    ref = np.linalg.solve(A + lam * np.eye(n), g)
    xo = np.zeros(n)
    for t in range(ntr):
        xo[3 * orow[t]:3 * orow[t] + 3] = x[t]
    return xo, ref, nlev, ntr


@pytest.mark.parametrize("kind", ["se2", "se3", "se2_loops"])
def test_level_schedule_solves_the_reduced_system(kind):
    rng = np.random.default_rng(12)
    if kind == "se2_loops":   # a tour that revisits its path: loop closures, fill-in, several rows per level
        ds = D.graphslam_se2_tour(420, seed=2)
        kw = dict(pose_kind=0, obs_kind=0, noise_kind=1, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]), max_tree_depth=3,
                  max_optimize_depth=3, submap_size=10, min_obs_to_loop_closure=1)
        e = sb.RbaEngine(sb.default_engine_params(**kw))
        flats = []
        for k, f in enumerate(ds["frames"]):
            D.feed(e, f, run_local_optimization=False)
            if k in (300, 419):
                flats.append(e.flatten_local_area(k, 3))
        P = 3
    else:
        flats, P = (flats_se2(), 3) if kind == "se2" else (flats_se3(), 6)
    for fl in flats:
        x, ref, nlev, ntr = emulate_levels(fl, P, rng)
        assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
        assert nlev <= ntr
        if kind == "se2" and ntr >= 16:
            assert nlev <= 0.6 * ntr   # independent minimum-degree rounds: a chain is halved per round, not peeled from its ends


def test_wave_counts_and_execution_mode_policy():
    """srba_b200_problem_wave_counts / srba_b200_problem_plan (host-only): which problems "auto" sends to the grid-wide wave pipeline --
    SE(2) areas that overflow one CTA's arena (loop closures), not chain-like SE(2) areas, not SE(3) landmark areas of local-area size --
    and the shape of their level schedule (dense tail of at most SRBA_WAVE_TAIL_ROWS = 70 tile rows, team flag)."""
    import ctypes as C
    from srba_b200 import _capi as K
    L = K.lib()

    def counts(flat):
        o8, o = (C.c_int64 * 8)(), (C.c_int64 * 8)()
        assert L.srba_b200_problem_wave_counts(flat.problem, o8) == 0 and L.srba_b200_problem_plan(flat.problem, o) == 0
        return list(o8), list(o)
    # a tour that keeps revisiting its corridors: areas with hundreds of unknown edges
    ds = D.graphslam_se2_tour(1500, seed=1)
    kw = dict(pose_kind=0, obs_kind=0, noise_kind=1, noise_lambda=D.se2_noise_lambda(ds["sigma_xy"], ds["sigma_yaw"]), max_tree_depth=3, max_optimize_depth=3,
              submap_size=10, min_obs_to_loop_closure=1)
    ep = sb.RbaEngine(sb.default_engine_params(**kw))
    small = big = None
    for k, f in enumerate(ds["frames"]):
        D.feed(ep, f, run_local_optimization=False)
        if k == 60:
            small = ep.flatten_local_area(k, 3)
        if k > 800 and k % 20 == 0:     # keep the largest area met on the way
            fl = ep.flatten_local_area(k, 3)
            if big is None or fl.p.n_unk_edges > big.p.n_unk_edges:
                big = fl
    w, plan = counts(small)
    assert w == [0] * 8 and plan[4] < 4 and plan[6] > 0          # chain-like: the one-CTA program with its arrays on chip
    w, plan = counts(big)
    assert big.p.n_unk_edges > 100 and plan[4] & 4 and plan[6] == 0   # wave pipeline, every array in the global workspace
    tiles, ndst, ncontrib, levels, tail_level, team, tiles_below, dst_below = w
    assert tiles > 0 and ncontrib >= ndst > 0 and 0 <= tail_level <= levels and tiles_below <= tiles and dst_below <= ndst
    assert team == (1 if tail_level > 0 and ndst >= 600 * tail_level else 0)
    old = K.set_option("wave", 0)
    try:
        assert counts(big)[0] == [0] * 8 and not (counts(big)[1][4] & 4)
    finally:
        K.set_option("wave", old)
    # SE(3) landmark areas of local-area size stay on the one-CTA / cluster program unless the wave mode is forced
    d3 = D.landmark_slam_se3(n_kfs=40, n_lms=400, seed=3, obs_kind="cartesian3d")
    e3 = sb.RbaEngine(sb.default_engine_params(pose_kind=1, obs_kind=1, noise_kind=0, noise_std=1e-3, max_tree_depth=4, max_optimize_depth=4, submap_size=10,
                                               min_obs_to_loop_closure=4))
    for f in d3["frames"]:
        D.feed(e3, f, run_local_optimization=False)
    f3 = e3.flatten_local_area(39, 4)
    assert counts(f3)[0] == [0] * 8
    old = K.set_option("wave", 1)
    try:
        assert counts(f3)[0][0] > 0 and counts(f3)[1][4] & 4
    finally:
        K.set_option("wave", old)
