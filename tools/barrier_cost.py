"""estimate the cost of one solver phase (grid barrier + a small batch) from a small scene: iterate time / (2 * iters * batches)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bullet3_b200 import capi, scenes

for n in (6, 12, 24):
    w = capi.World(capi.default_config(n ** 3 + 16))
    scenes.box_stack(w, n, n, n)
    w.upload()
    for iters in (10, 50):
        w.set_solver(capi.SOLVER_PGS, iters)
        w.step_n(1 / 60, 30)
        w.enable_stage_timing(True)
        t = np.zeros(8)
        for _ in range(10):
            w.step(1 / 60)
            t += w.stage_timings()
        t /= 10
        sizes = np.diff(w.batches())
        nb = len(sizes)
        print("n=%d iters=%d contacts=%d batches=%d sizes=%s iterate=%.1f us -> %.2f us per phase" % (
            n, iters, w.counters()[1], nb, sizes.tolist()[:12], t[4] * 1e3, t[4] * 1e3 / (2 * iters * max(nb, 1))))
        w.enable_stage_timing(False)
    w.close()
