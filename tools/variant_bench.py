"""A/B timing of experiment builds of the library on the bench batch (BASELINE config #4, 256 areas of the 3 000-keyframe tour).

    SRBA_B200_LIB=srba_b200/_variants/x.so SRBA_B200_NVCC_EXTRA="-DWAVE_FWD_UNROLL=2" python -c "from srba_b200 import build; build.build(force=True)"
    python tools/variant_bench.py srba_b200/_variants/*.so          (on the GPU box)

The problems are built ONCE with the in-tree library; every variant .so is then loaded next to it (plain C structs cross the boundary),
packs the same problems, runs the whole batch `--runs` times (L2 flushed in between, CUDA events on the launching stream) and must
reproduce the in-tree library's final edge poses bit for bit (same arithmetic, different schedule)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def declare(L, K):
    vp, i32 = C.c_void_p, C.c_int32
    pp, pr = C.POINTER(K.Problem), C.POINTER(K.Result)
    L.srba_b200_batch_create.restype = vp
    L.srba_b200_batch_create.argtypes = [i32, C.POINTER(pp), i32]
    for n in ("upload", "run", "download"):
        f = getattr(L, "srba_b200_batch_" + n)
        f.restype = C.c_int
        f.argtypes = [vp, vp]
    L.srba_b200_batch_unpack.restype = C.c_int
    L.srba_b200_batch_unpack.argtypes = [vp, i32, pr]
    L.srba_b200_batch_destroy.argtypes = [vp]
    L.srba_b200_last_error.restype = C.c_char_p
    L.srba_b200_kernel_timing.argtypes = [i32]
    L.srba_b200_kernel_timing_get.argtypes = [C.POINTER(C.c_double)]
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--areas", type=int, default=256)
    ap.add_argument("--kfs", type=int, default=3000)
    a = ap.parse_args()
    import torch
    import srba_b200 as sb
    from srba_b200 import _capi as K
    import bench
    ds, eng, sets = bench.build_problems(seed=1, areas=a.areas, n_kfs=a.kfs)
    flats = sets[a.areas]
    n = len(flats)
    parr = (C.POINTER(K.Problem) * n)(*[f.problem for f in flats])
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ref = None
    for path in [K.SO_PATH] + [os.path.abspath(p) for p in a.libs]:
        L = declare(C.CDLL(path), K)
        h = L.srba_b200_batch_create(n, parr, 0)
        if not h:
            print("%-40s batch_create failed: %s" % (os.path.basename(path), L.srba_b200_last_error()))
            continue
        ok = L.srba_b200_batch_upload(h, sp) == 0
        for _ in range(2):
            ok = ok and L.srba_b200_batch_run(h, sp) == 0
        torch.cuda.synchronize()
        if not ok:
            print("%-40s failed: %s" % (os.path.basename(path), L.srba_b200_last_error()))
            continue
        L.srba_b200_kernel_timing(1)
        ms = []
        for _ in range(a.runs):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); L.srba_b200_batch_run(h, sp); e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        kt = (C.c_double * 4)()
        L.srba_b200_kernel_timing_get(kt)
        L.srba_b200_kernel_timing(0)
        L.srba_b200_batch_download(h, sp)
        torch.cuda.synchronize()
        outs, trials = [], 0
        for i, f in enumerate(flats):
            rb = sb.ResultBuffers(f.p)
            L.srba_b200_batch_unpack(h, i, C.byref(rb.r))
            outs.append(np.array(rb.edge_pose, copy=True))
            trials += rb.r.stats.num_trials
        if "timers" in os.path.basename(path):   # a -DSRBA_WAVE_TIMERS build: in-kernel clocks of the solver, accumulated over every run of this batch
            L.srba_b200_batch_stage_cycles.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
            big = np.argsort([-f.p.n_unk_edges for f in flats])[:6]
            print("   n_p trials | us per trial: fwd gather, level work, level barrier | bwd level work, barrier | tail factor, tail back-substitution")
            for i in big:
                sc = (C.c_uint64 * 24)()
                L.srba_b200_batch_stage_cycles(h, int(i), sc)
                rb = sb.ResultBuffers(flats[i].p)
                L.srba_b200_batch_unpack(h, int(i), C.byref(rb.r))
                per = 1e-3 / ((2 + a.runs) * max(1, rb.r.stats.num_trials))
                print("   %4d %4d | %7.1f %7.1f %7.1f | %7.1f %7.1f | %7.1f %7.1f" % ((flats[i].p.n_unk_edges, rb.r.stats.num_trials) + tuple(sc[k] * per for k in (0, 3, 4, 5, 6, 7, 8))))
        L.srba_b200_batch_destroy(h)
        if ref is None:
            ref = outs
        worst = max(float(np.abs(x - y).max()) for x, y in zip(outs, ref))
        print("%-40s step %8.3f ms (min %8.3f)  trials %6d  dominant kernel %8.3f ms in %4d launches  max|edge - in-tree| %.3e"
              % (os.path.basename(path), float(np.median(ms)), min(ms), trials, kt[0] / a.runs, int(kt[1]) // a.runs, worst), flush=True)


if __name__ == "__main__":
    main()
