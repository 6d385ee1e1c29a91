"""stage timings of the settled bench scene (BASELINE configs[3]) under the current environment (A/B runs of env switches):
   python tools/stage_ab.py [side] [settle] [steps]   ->  one line of per-stage ms (b3b200_stage_timings) + counters"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 250
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
w = capi.World(bench.bench_config(capi, side))
scenes.bench_config4_scene(w, *bench.scene_dims(side))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, settle)
w.synchronize()
w.enable_stage_timing(True)
st = np.zeros(8)
for _ in range(steps):
    w.step(1 / 60)
    st += w.stage_timings()
st /= steps
names = ["aabb", "broadphase", "narrowphase", "setup", "iterate", "integrate", "total?", "sat"]
print(" ".join("%s=%.3f" % (n, v) for n, v in zip(names, st)), "| env", {k: v for k, v in os.environ.items() if k.startswith("B3B200")}, "| ctr", w.counters()[:5].tolist())
b = w.bodies()
print("checksum pos %.6f vel %.6f" % (float(np.abs(b["pos"][:, :3]).sum()), float(np.abs(b["linVel"][:, :3]).sum())))
