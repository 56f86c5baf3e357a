"""In-kernel phase clocks of the wave solver (build with SRBA_B200_NVCC_EXTRA=-DSRBA_WAVE_TIMERS): per area, ns spent by thread 0 of the team in
gather / A / sync / B / sync (forward levels) and row / sync (backward levels), summed over the LM trials of one batch run."""
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
    rows.append((f.p.n_unk_edges, b.stats(i)["num_trials"], out[:7].astype(np.float64) / 1e3))
rows.sort(key=lambda r: -r[0])
print("n_p trials | us per trial: gather - - B syncB | bwd row sync")
for n_p, tr, t in rows[:10] + rows[100:104] + rows[200:203]:
    print(n_p, tr, " ".join("%8.1f" % (x / max(1, tr)) for x in t))
