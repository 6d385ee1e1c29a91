"""time b3b200_cast_rays on the settled bench scene (BASELINE configs[3]): args side settle numRays"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 100
nrays = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
w = capi.World(bench.bench_config(capi, side))
scenes.bench_config4_scene(w, *bench.scene_dims(side))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, settle)
w.synchronize()
b = w.bodies()
lo, hi = b["pos"][1:, :3].min(0), b["pos"][1:, :3].max(0)
rng = np.random.default_rng(0)
# camera-style picking rays: from a point above the pile through random points inside it
frm = np.tile(np.float32([(lo[0] + hi[0]) / 2, hi[1] + 40.0, (lo[2] + hi[2]) / 2]), (nrays, 1))
to = rng.uniform(lo, hi, (nrays, 3)).astype(np.float32)
to = frm + (to - frm) * 1.5
for rep in range(6):
    w.set_ray_accel(0 if rep < 2 else 1)
    t0 = time.perf_counter()
    h = w.cast_rays(frm, to)
    dt = time.perf_counter() - t0
    print("bodies %d rays %d: %.2f ms (%.2f Mrays/s), hits %d" % (len(b), nrays, dt * 1e3, nrays / dt / 1e6, (h["hitBody"] >= 0).sum()))
