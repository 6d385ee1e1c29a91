"""launch-bound small scenes: ms/step of BASELINE configs[0] (1 000 boxes) with and without the captured step graph"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bullet3_b200 import capi, scenes  # noqa: E402

for graphs in (0, 1):
    w = capi.World(capi.default_config(2048))
    scenes.box_plane_scene(w, 10, 10, 10)
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 4)
    w.set_step_graphs(graphs)
    w.step_n(1 / 60, 100)
    w.synchronize()
    t0 = time.perf_counter()
    w.step_n(1 / 60, 1000)
    w.synchronize()
    t1 = time.perf_counter()
    print("graphs=%d: %.4f ms/step (1001 bodies, %d contacts)" % (graphs, (t1 - t0), w.counters()[1]))
    w.close()
