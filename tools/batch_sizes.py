import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from bullet3_b200 import capi, scenes
w = capi.World(bench.bench_config(capi, 64))
scenes.bench_config4_scene(w, *bench.scene_dims(64))
w.upload(); w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1/60, 250); w.synchronize()
off = w.batches()
print("batch sizes", np.diff(off).tolist())
