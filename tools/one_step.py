"""build the bench scene, settle, then run a few steps (for ncu launch lists / captures)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bullet3_b200 import capi, scenes
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 60
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w = capi.World(capi.default_config(side ** 3 + 16))
scenes.bench_convex_scene(w, side, side, side)
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, settle)
w.synchronize()
import torch
torch.cuda.profiler.start()
w.step_n(1 / 60, steps)
w.synchronize()
torch.cuda.profiler.stop()
print(w.counters())
