"""solver iteration stage (ms) of the barrier kernel vs the dataflow kernel on the settled bench scene"""
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
for mode in (False, True, False, True):
    w.set_solver_dataflow(mode)
    it, tot = [], []
    for _ in range(20):
        w.step(1 / 60)
        t = w.stage_timings()
        it.append(t[4])
        tot.append(t[6])
    print("dataflow" if mode else "barrier ", "iterate %.3f ms (min %.3f)  step %.3f ms" % (np.median(it), np.min(it), np.median(tot)), flush=True)
b = w.bodies()
print("finite", bool(np.isfinite(b["pos"]).all()), "max speed %.3f" % float(np.abs(b["linVel"][:, :3]).max()))
