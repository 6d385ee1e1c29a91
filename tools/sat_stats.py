"""development aid: statistics of the feasibility-cone SAT on the settled bench scene (library built with
   make FLAGS_narrowphase=-DB3B200_SAT_STATS)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 250
w = capi.World(bench.bench_config(capi, side))
scenes.bench_config4_scene(w, *bench.scene_dims(side))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, settle)
w.synchronize()
L = C.CDLL(os.path.join(ROOT, "bullet3_b200", "libb3b200.so"))
out = np.zeros(16, np.uint64)
L.b3b200_debug_sat_stats(None, 1)
w.step(1 / 60)
w.synchronize()
L.b3b200_debug_sat_stats(out.ctypes.data_as(C.c_void_p), 0)
names = ["items", "rounds", "axes", "faceCandAfterCone", "aliveRows", "aliveCols", "pairsEnumerated", "pairsKept", "conesBuilt", "conesUseful", "sumCosT*1000", "separated"]
n = max(int(out[0]), 1)
for k, v in zip(names, out):
    print("%-20s %12d  per item %.2f" % (k, int(v), int(v) / n))
print("mean cosT over built cones: %.4f" % (int(out[10]) / 1000 / max(int(out[8]), 1)))
