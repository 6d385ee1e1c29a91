"""development aid: stage timings of one rank's share of the slab bench scene (32 x 64 x 256 boxes) on one GPU, no neighbours"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bullet3_b200 import capi, scenes  # noqa: E402

nx, ny, nz = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 64, 256
i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
pos = np.stack([(((j + 1) & 1) + 2.2 * i).reshape(-1), (1.0 + 2.0 * j).reshape(-1), (((j + 1) & 1) + 2.2 * k).reshape(-1), 0 * i.reshape(-1)], 1).astype(np.float32)
n = len(pos)
w = capi.World(capi.default_config(n + 64))
ground = w.register_convex_points(scenes.box_points(2000.0))
box = w.register_convex_points(scenes.box_points(1.0))
w.register_instance(0.0, (0.0, -2000.0, 0.0), scenes.IDENT, ground)
w.register_instances(np.ones(n, np.float32), pos, np.tile(np.array(scenes.IDENT, np.float32), (n, 1)), np.full(n, box, np.int32))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, 10)
w.enable_stage_timing(True)
acc = np.zeros(8)
for _ in range(10):
    w.step(1 / 60)
    acc += w.stage_timings()
print("bodies", n, "stage ms [aabb, bp, np, setup, iterate, integrate, total, sat]", np.round(acc / 10, 3), "counters", w.counters())
