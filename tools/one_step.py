"""build the bench scene (BASELINE configs[3]), settle, then run a few steps inside a cudaProfilerStart/Stop window
(for `ncu --profile-from-start off` launch lists / captures).  args: side settle steps [scene: c4|convex]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 250
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
scene = sys.argv[4] if len(sys.argv) > 4 else "c4"
w = capi.World(bench.bench_config(capi, side))
if scene == "c4":
    scenes.bench_config4_scene(w, *bench.scene_dims(side))
else:
    scenes.bench_convex_scene(w, *bench.scene_dims(side))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, settle)
w.synchronize()
import torch  # noqa: E402

torch.cuda.profiler.start()
w.step_n(1 / 60, steps)
w.synchronize()
torch.cuda.profiler.stop()
print(w.counters())
