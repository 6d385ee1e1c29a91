"""solver setup / iterate stage times and batch counts for the two batch-assignment modes on the settled bench scene"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = capi.World(bench.bench_config(capi, side))
scenes.bench_config4_scene(w, *bench.scene_dims(side))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, 250)
w.synchronize()
w.enable_stage_timing(True)
for mode in (0, 1, 0, 1):
    w.set_colouring(mode)
    su, it, tot, nb = [], [], [], []
    for _ in range(20):
        w.step(1 / 60)
        t = w.stage_timings()
        su.append(t[3]); it.append(t[4]); tot.append(t[6]); nb.append(w.counters()[2])
    print("colouring %d: setup %.3f ms  iterate %.3f ms  step %.3f ms  batches %d..%d" % (mode, np.median(su), np.median(it), np.median(tot), min(nb), max(nb)), flush=True)
print("batch sizes (mode 1)", np.diff(w.batches()).tolist())
b = w.bodies()
print("finite", bool(np.isfinite(b["pos"]).all()), "overflow flags", w.counters()[4])
