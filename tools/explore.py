"""scene statistics + per-stage timings (development aid; run under gpurun)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bullet3_b200 import capi, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
bp = int(sys.argv[4]) if len(sys.argv) > 4 else 1
t0 = time.time()
w = capi.World(capi.default_config(n * n * n + 16))
ny = int(sys.argv[5]) if len(sys.argv) > 5 else n
nx = nz = int(round((n * n * n / ny) ** 0.5))
scenes.bench_convex_scene(w, nx, ny, nz)
w.upload()
w.set_solver(capi.SOLVER_PGS, iters)
w.set_broadphase(bp)
w.set_solver_dataflow(int(os.environ.get('DATAFLOW', '1')))
print("setup %.1fs bodies=%d" % (time.time() - t0, w.num_bodies), flush=True)
w.enable_stage_timing(True)
for s in range(steps):
    w.step(1 / 60)
    if s % 10 == 0 or s == steps - 1:
        c = w.counters()
        ms = w.stage_timings()
        print("step %3d pairs=%d contacts=%d batches=%d rounds=%d ovf=%d | aabb %.3f bp %.3f np %.3f setup %.3f iter %.3f integ %.3f total %.3f ms" %
              (s, c[0], c[1], c[2], c[3], c[4], ms[0], ms[1], ms[2], ms[3], ms[4], ms[5], ms[6]), flush=True)
b = w.bodies()
dyn = b["invMass"] != 0
print("finite", np.isfinite(b["pos"]).all(), "min y", b["pos"][dyn, 1].min(), "max |v|", np.abs(b["linVel"][dyn, :3]).max(), "median |v|", np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)))
