"""Prints where the LM kernel spends its cycles (per stage, measured by the kernel with clock64) for the bench batch.
Needs a library built with the probes:  SRBA_SOLVER_TIMERS=1 python -m srba_b200.build --force
Usage: python tools/stage_breakdown.py [areas] [tour_keyframes] [tour_seed]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
import srba_b200 as sb

areas = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kfs = int(sys.argv[2]) if len(sys.argv) > 2 else 700
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ds, eng, flats = bench.build_problems(seed, areas, kfs)
b = sb.Batch(flats)
b.upload(); b.run(); b.download()
import torch
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); b.run(); t1.record(); torch.cuda.synchronize()
print("batch of %d: %.3f ms" % (len(flats), t0.elapsed_time(t1)))
b.download()
tot = {}
rows = []
for i in range(b.n):
    sc = b.stage_cycles(i)
    st = b.stats(i)
    rows.append((sum(v for k, v in sc.items() if not k.startswith("chol.")), flats[i].p.n_unk_edges, st["num_trials"], sc))
    for k, v in sc.items():
        tot[k] = tot.get(k, 0) + v
T = sum(v for k, v in tot.items() if not k.startswith("chol."))
print("all problems: total %.1f Mcycles; share per stage:" % (T / 1e6))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("  %-10s %5.1f%%" % (k, 100.0 * v / T))
rows.sort(key=lambda r: -r[0])
for r in rows[:3] + rows[len(rows) // 2:len(rows) // 2 + 2]:
    print("problem n_p=%d trials=%d total %.0f kcycles (%.1f us/trial @1.96GHz):" % (r[1], r[2], r[0] / 1e3, r[0] / max(1, r[2]) / 1960.0),
          {k: int(v / max(1, r[2])) for k, v in r[3].items()})
tr = np.array([r[2] for r in rows]); npv = np.array([r[1] for r in rows]); cyc = np.array([r[0] for r in rows])
print("trials: mean %.1f max %d; n_p: median %d max %d; per-problem ms: median %.3f max %.3f" % (tr.mean(), tr.max(), np.median(npv), npv.max(), np.median(cyc) / 1.96e6, cyc.max() / 1.96e6))
