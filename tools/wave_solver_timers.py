"""In-kernel phase clocks of the wave solver (build with SRBA_B200_NVCC_EXTRA=-DSRBA_WAVE_TIMERS): per area, ns spent by thread 0 of the team in
the forward kernel (gather | level work | level barrier), the backward kernel (level work | barrier) and the dense tail (factorisation |
back-substitution), summed over the LM trials of one batch run.  tools/variant_bench.py prints the same for a `*timers*.so` experiment build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
import srba_b200 as sb
import bench
ds, eng, sets = bench.build_problems(seed=1, areas=256, n_kfs=3000)
flats = sets[256]
b = sb.Batch(flats)
b.upload(); b.run(); b.download()
rows = []
for i, f in enumerate(flats):
    out = np.zeros(24, np.uint64)
    b.L.srba_b200_batch_stage_cycles(b.h, i, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    rows.append((f.p.n_unk_edges, b.stats(i)["num_trials"], out[[0, 3, 4, 5, 6, 7, 8]].astype(np.float64) / 1e3))
rows.sort(key=lambda r: -r[0])
print("n_p trials | us per trial: fwd gather, level work, level barrier | bwd level work, barrier | tail factor, tail back-substitution")
for n_p, tr, t in rows[:10] + rows[100:104] + rows[200:203]:
    print(n_p, tr, " ".join("%8.1f" % (x / max(1, tr)) for x in t))
