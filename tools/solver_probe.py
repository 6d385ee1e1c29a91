"""development aid: settle the bench scene, then print where CTA 0 of the solver iteration kernel spends a pass
(globaltimer stamps: 1 publish, 2 arrive, 3 cross colours done, 4 boundary reloaded, then the interior until the next 1)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
L = capi.lib()
L.b3b200_debug_solver_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
w.enable_stage_timing(True)
for rep in range(2):
    L.b3b200_debug_solver_probe(w.h, None, 0)
    w.step(1 / 60)
    buf = np.zeros(512, np.uint64)
    L.b3b200_debug_solver_probe(w.h, buf.ctypes.data, 512)
    raw_all = buf.copy()
    buf = buf[:256]
    buf = buf[buf != 0]
    t, tag = (buf >> np.uint64(4)).astype(np.int64), (buf & np.uint64(15)).astype(int)
    print("stage ms", w.stage_timings(), "counters", w.counters())
    names = {0: "start", 1: "pass begins", 2: "published", 3: "cross done", 4: "reloaded", 5: "interior done (last)", 6: "stored"}
    seg = {}
    for i in range(1, len(t)):
        seg.setdefault((tag[i - 1], tag[i]), []).append((t[i] - t[i - 1]) / 1e3)
    for k, v in seg.items():
        print("%-22s -> %-22s n=%3d  mean %7.2f us  min %7.2f  max %7.2f" % (names[k[0]], names[k[1]], len(v), np.mean(v), np.min(v), np.max(v)))
    print("kernel total %.1f us" % ((t[-1] - t[0]) / 1e3))
    # interior of pass 1 (normal rows), warp 0 of CTA 0: stamps at every colour barrier (tag = colour) and around its tile solves (200 / 201)
    raw = raw_all[256:]
    raw = raw[raw != 0]
    if len(raw):
        tt, tg = (raw >> np.uint64(8)).astype(np.int64), (raw & np.uint64(255)).astype(int)
        print("interior pass 1, warp 0:", " ".join("%s+%.2f" % (("c%d" % g) if g < 200 else ("S" if g == 200 else "E"), (tt[i] - tt[i - 1]) / 1e3) for i, g in enumerate(tg) if i > 0))
