"""ms/step of the B200 path on every BASELINE.json config (configs[3] is bench.py's headline; this adds the others).
Writes one line per config; run on a GPU box:  python tools/configs_report.py > gpurun_out/configs.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402


def timed_steps(w, steps, settle):
    w.step_n(1 / 60, settle)
    w.synchronize()
    w.enable_stage_timing(True)
    t = np.zeros(8)
    n = 20
    for _ in range(n):
        w.step(1 / 60)
        t += w.stage_timings()
    w.enable_stage_timing(False)
    w.synchronize()
    t0 = time.perf_counter()
    w.step_n(1 / 60, steps)
    w.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    return t / n, wall


def report(name, w, stage, wall):
    c = w.counters()
    print("%-58s bodies %7d pairs %8d contacts %7d batches %3d | step %.3f ms (aabb %.3f bp %.3f np %.3f setup %.3f solve %.3f integ %.3f) | %.2f M bodies*steps/s" % (
        name, w.num_bodies, c[0], c[1], c[2], wall, stage[0], stage[1], stage[2], stage[3], stage[4], stage[5], w.num_bodies / wall / 1e3))


# config 1: 1 000 unit boxes (10x10x10) on a static ground, 600 steps
w = capi.World(capi.default_config(1100))
scenes.box_stack(w, 10, 10, 10)
w.upload()
w.set_solver(capi.SOLVER_PGS, 4)
st, wall = timed_steps(w, 600, 60)
report("configs[0] 1 000 boxes 10x10x10, PGS 4 it.", w, st, wall)
w.close()

# config 2: broadphase only on 64006GPUAABBs (PairBench recipe), grid and SAP; inflated by 2 so that pairs exist
import pairbench  # noqa: E402

aabbs, small, large = pairbench.load_pairbench_aabbs()
for margin in (0.0, 2.0):
    a = aabbs.copy()
    a["min"][:, :3] -= margin
    a["max"][:, :3] += margin
    for kind, nm in ((capi.BP_GRID, "grid"), (capi.BP_SAP, "SAP")):
        bp = capi.Broadphase(kind, len(a) + 16, 3 << 20)
        for i in range(len(a)):
            pass
        is_large = np.zeros(len(a), bool)
        is_large[large] = True
        for i in range(len(a)):
            (bp.create_large_proxy if is_large[i] else bp.create_proxy)(a["min"][i, :3], a["max"][i, :3], int(a["minIndex"][i]))
        bp.write_aabbs()
        ms = []
        for _ in range(12):
            bp.calculate_pairs(3 << 20)
            ms.append(bp.last_ms())
        print("configs[1] PairBench 64 006 AABBs (+%.0f margin), %-4s                 pairs %8d | %.3f ms per calculateOverlappingPairs" % (
            margin, nm, bp.num_overlap(), float(np.median(ms[2:]))))
        bp.close()

# config 3: GpuBoxPlaneScene, 64x32x64 = 131 072 boxes on a plane, Jacobi 8 iterations, SAP
w = capi.World(capi.default_config(131072 + 16))
w.register_instance(0.0, (0, 0, 0), scenes.IDENT, w.register_plane((0, 1, 0), 0.0))
scenes.box_plane_scene(w, 64, 32, 64, ground=False)
w.upload()
w.set_broadphase(capi.BP_SAP)
w.set_solver(capi.SOLVER_JACOBI, 8)
st, wall = timed_steps(w, 100, 200)
report("configs[2] GpuBoxPlaneScene 131 072 boxes, Jacobi 8 it., SAP", w, st, wall)
w.set_broadphase(capi.BP_GRID)
w.set_solver(capi.SOLVER_PGS, 8)
st, wall = timed_steps(w, 100, 20)
report("           same scene, batched PGS 8 it., grid", w, st, wall)
w.close()

# config 5 (i), one GPU's share: 1 024 independent worlds (8x4x8 cubes + their own static ground box each, config-3 recipe), all at
# the SAME coordinates, batched into one b3b200 world (b3b200_set_current_world)
nw = 1024
w = capi.World(capi.default_config(nw * 257 + 64))
scenes.batched_box_worlds(w, nw)
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
st, wall = timed_steps(w, 100, 200)
p = w.pairs()
world_of = w.body_worlds()
cross = int((world_of[p["x"]] != world_of[p["y"]]).sum())
report("configs[4](i) 1 024 batched worlds x 257 bodies, PGS 10 it.", w, st, wall)
print("           pairs that join two worlds: %d; cross-block batches %d" % (cross, w.counters()[3]))
w.close()
