#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native rigid-body step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bodies-side S]

Workload (BASELINE.json metric: "bodies*steps/s and ms/step (256k convex)", configs[3]): the
GpuConvexScene-style scene at 64^3 = 262 144 dynamic bodies (boxes, tetrahedra, seeded random
8-32-vertex hulls, three-box compounds, seeded random orientations) settled on a concave
heightfield trimesh (256 x 256 quads), batched PGS with 10 iterations, dt = 1/60.
A "step" is one b3GpuRigidBodyPipeline::stepSimulation.
N > 1 (torchrun): every rank steps its own independent copy of the scene (batched
independent worlds, no data-path collective) -> "scaling": "weak".

Prints ONE JSON line (rank 0).  See the module docstring of DESIGN.md section 6 for the
roofline byte model.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = 1.0 / 60.0
ITERS = 10
METRIC = "bodies*steps/s (256k convex scene)"  # one string for both arms: the driver divides lines with equal metrics
SETTLE_STEPS = 250  # untimed: the 16-layer lattice drops and comes to rest (contacts/body plateaus) before anything is measured
LAYERS = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies-side", type=int, default=64, help="scene is side^3 bodies (64 -> 262 144)")
    ap.add_argument("--ref-bodies", type=int, default=8192, help="--impl reference: bodies of the scene region one reference step simulates")
    ap.add_argument("--bt2-bodies", type=int, default=65536, help="bodies of the scene region the Bullet 2 MT baseline steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched independent worlds (BASELINE configs[4](i)) that are timed after the headline scene")
    ap.add_argument("--worlds-per-gpu", type=int, default=1024, help="independent 256-box worlds batched into one b3b200 world per GPU")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the slab-decomposed scene (BASELINE configs[4](ii)) that is timed after the replicas")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region"""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU arms
def settled_state(side):
    """(full-lattice body array, True) after SETTLE_STEPS steps on the GPU when there is one -- state PREPARATION for the
    CPU arms, outside every timed region -- else (None, False): the CPU arms then start from the unsettled lattice"""
    try:
        dev = os.environ.get("B3B200_SETTLED_NPY")  # development aid: a state saved by tools/dump_settled.py
        if dev and os.path.exists(dev):
            return np.load(dev), True
        import torch

        if not torch.cuda.is_available():
            return None, False
        from bullet3_b200 import capi, scenes

        w = capi.World(bench_config(capi, side))
        scenes.bench_config4_scene(w, *scene_dims(side))
        w.upload()
        w.set_solver(capi.SOLVER_PGS, ITERS)
        w.step_n(DT, SETTLE_STEPS)
        bodies = w.bodies()
        w.close()
        return bodies, True
    except Exception:
        return None, False


def region_of(side, bodies, target):
    """indices (into the lattice bodies, 0-based without the mesh) of a square x-z window around the centre of the pile that
    holds about `target` bodies: one connected piece of the real scene, all 16 layers"""
    nx, ny, nz = scene_dims(side)
    n = nx * ny * nz
    if target >= n:
        return np.arange(n)
    if bodies is not None:
        x, z = bodies["pos"][1:, 0], bodies["pos"][1:, 2]
    else:
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        x, z = (2.2 * i).reshape(-1).astype(np.float32), (2.2 * k).reshape(-1).astype(np.float32)
    cx, cz = np.median(x), np.median(z)
    d = np.maximum(np.abs(x - cx), np.abs(z - cz))
    half = np.sort(d)[target - 1]
    return np.nonzero(d <= half)[0]


def build_region_world(world, side, bodies, keep):
    """register the bench scene recipe restricted to `keep` (same shapes, same mesh) with the poses of `bodies` (or the lattice)"""
    from bullet3_b200 import scenes

    poses = None
    if bodies is not None:
        pos = np.ascontiguousarray(bodies["pos"][1:]).copy()
        pos[:, 3] = 0.0
        poses = (pos, np.ascontiguousarray(bodies["quat"][1:]))
    scenes.bench_config4_scene(world, *scene_dims(side), keep=keep, poses=poses)


def run_reference_arm(side, steps, warmup, target_bodies):
    """the reference's own b3GpuRigidBodyPipeline::stepSimulation, UNMODIFIED, with every host-twin switch on (oracle/_ref/
    libb3refcl.so: gCalcWorldSpaceAabbOnCpu, gUseDbvt, CHECK_ON_HOST contact loop + concave host twins, the gCpu* solver stages,
    gIntegrateOnCpu), stepping one connected region of the settled scene.  Returns (value, ms/step, description, cores)."""
    import oracle_api as oa

    if not oa.refcl_available():
        raise SystemExit("oracle/_ref/libb3refcl.so is missing (built by __graft_entry__.build() where /root/reference exists)")
    from bullet3_b200 import capi

    bodies, settled = settled_state(side)
    keep = region_of(side, bodies, target_bodies)
    w = oa.RefPipeline(bench_config(capi, side))
    build_region_world(w, side, bodies, keep)
    w.upload()
    if bodies is not None:
        state = w.bodies()
        for f in ("linVel", "angVel"):
            state[f][1:] = bodies[f][1:][keep]
        w.set_bodies(state)
    n = w.num_bodies
    out = w.step(DT, warmup) if warmup else [0, 0, 0]
    w.profile_zones()
    t0 = time.perf_counter()
    out = w.step(DT, steps)
    el = time.perf_counter() - t0
    zones = w.profile_zones()
    w.close()
    desc = ("UNMODIFIED reference b3GpuRigidBodyPipeline::stepSimulation with every host-twin flag on (oracle/_ref/libb3refcl.so over a host-memory "
            "fake OpenCL): AABBs on CPU, b3DynamicBvhBroadphase, CHECK_ON_HOST contacts (hulls, compounds, trimesh), b3GpuPgsContactSolver host "
            "stages -> b3Solver::solveContactConstraintHost (its hard-coded 4 iterations), integrate on CPU; single-threaded by construction; "
            "one world = a connected %d-body region (all 16 layers, %s) of the %d-body scene on the full mesh, %d warm-up + %d timed steps; "
            "%d DBVT pairs / %d contacts in the last step.  The full 262 145-body world takes 44 s per step through this path even before it "
            "has settled (measured; its host batching scans all bodies per cell and batch), so the region is the bounded sample" %
            (n, "settled state of the GPU run" if settled else "UNSETTLED lattice: no GPU for state preparation", side ** 3 + 1, warmup, steps, out[0], out[1]))
    top = sorted(zones.items(), key=lambda kv: -kv[1])[:8]
    desc += ".  Its own B3_PROFILE zones, ms per step (inclusive): " + ", ".join("%s %.0f" % (k, v / steps * 1e3) for k, v in top)
    return n * steps / el, el / steps * 1e3, desc, 1


def dump_bt2_scene(path, tables, bodies):
    """scene file of oracle/bt2/bt2mt_bench.cpp: the collidables as point clouds / compounds of them / one triangle mesh"""
    col, convex, verts = tables["collidables"], tables["convex"], tables["vertices"]
    faces, indices, children = tables["faces"], tables["indices"], tables["child_shapes"]
    mesh_v, mesh_i = np.zeros((0, 3), np.float32), np.zeros(0, np.int32)
    with open(path, "wb") as f:
        f.write(np.array([0x62743273, len(col), len(bodies), 0, 0], np.int32).tobytes())  # header patched below
        for c in col:
            st, si = int(c["shapeType"]), int(c["shapeIndex"])
            if st == 3:  # SHAPE_CONVEX_HULL
                cv = convex[si]
                pts = np.ascontiguousarray(verts[cv["vertexOffset"]: cv["vertexOffset"] + cv["numVertices"]]).view(np.float32).reshape(-1, 4)[:, :3]
                f.write(np.array([0, len(pts)], np.int32).tobytes())
                f.write(np.ascontiguousarray(pts, np.float32).tobytes())
            elif st == 6:  # SHAPE_COMPOUND_OF_CONVEX_HULLS: numChildShapes / first child ride in the collidable's unions
                nch, first = int(c["numChildShapes"]), si
                f.write(np.array([1, nch], np.int32).tobytes())
                for ch in children[first: first + nch]:
                    f.write(np.array([int(ch["shapeIndex"])], np.int32).tobytes())
                    f.write(np.concatenate([np.asarray(ch["childPosition"], np.float32).reshape(-1)[:3],
                                            np.asarray(ch["childOrientation"], np.float32).reshape(-1)[:4]]).astype(np.float32).tobytes())
            else:
                f.write(np.array([0, 0], np.int32).tobytes())
                if st == 5:  # SHAPE_CONCAVE_TRIMESH: one convex-table entry whose faces are the triangles
                    cv = convex[si]
                    mesh_v = np.ascontiguousarray(verts[cv["vertexOffset"]: cv["vertexOffset"] + cv["numVertices"]]).view(np.float32).reshape(-1, 4)[:, :3]
                    fs = faces[cv["faceOffset"]: cv["faceOffset"] + cv["numFaces"]]
                    mesh_i = np.concatenate([indices[o: o + 3] for o in fs["indexOffset"]]).astype(np.int32) if len(fs) else mesh_i
        f.write(np.ascontiguousarray(mesh_v, np.float32).tobytes())
        f.write(np.ascontiguousarray(mesh_i, np.int32).tobytes())
        shape_type = col["shapeType"][bodies["collidableIdx"]]
        rec = np.zeros(len(bodies), np.dtype([("shape", np.int32), ("v", np.float32, 14)]))
        rec["shape"] = np.where(shape_type == 5, -1, bodies["collidableIdx"])
        rec["v"][:, 0] = np.where(bodies["invMass"] != 0, 1.0 / np.where(bodies["invMass"] != 0, bodies["invMass"], 1.0), 0.0)
        rec["v"][:, 1:4] = bodies["pos"][:, :3]
        rec["v"][:, 4:8] = bodies["quat"]
        rec["v"][:, 8:11] = bodies["linVel"][:, :3]
        rec["v"][:, 11:14] = bodies["angVel"][:, :3]
        f.write(rec.tobytes())
        f.seek(0)
        f.write(np.array([0x62743273, len(col), len(bodies), len(mesh_v), len(mesh_i)], np.int32).tobytes())


def run_bt2mt(tables, bodies, warmup, steps, threads=0):
    """Bullet 2 btDiscreteDynamicsWorldMt + btDbvtBroadphase + btCollisionDispatcherMt + btSequentialImpulseConstraintSolverMt assembled like
    examples/MultiThreadedDemo/CommonRigidBodyMTBase.cpp:518-567, from the unmodified reference sources (oracle/_ref/bt2mt_bench)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "bt2mt_bench")
    if not os.path.exists(exe):
        return None
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "scene.bin")
        dump_bt2_scene(path, tables, bodies)
        r = subprocess.run([exe, path, str(warmup), str(steps), str(ITERS), str(threads)], capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        return None
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_cpu_pipeline_config1(steps=600):
    """BASELINE configs[0]: 1 000 unit boxes (10 x 10 x 10) on a static 400-box through the reference's b3CpuRigidBodyPipeline
    (+ b3CpuNarrowPhase + b3DynamicBvhBroadphase; it has no contact solver), 600 steps at 1/60 s, single thread by construction"""
    import oracle_api as oa
    from bullet3_b200 import capi, scenes

    if not oa.ref_available():
        return None
    w = oa.RefCpuPipeline(capi.default_config(2048))
    scenes.box_stack(w, 10, 10, 10)
    n = w.num_bodies
    t0 = time.perf_counter()
    w.step(DT, steps)
    el = time.perf_counter() - t0
    w.close()
    return {"value": n * steps / el, "unit": "bodies*steps/s", "cores": 1, "kind": "reference", "ms_per_step": el / steps * 1e3,
            "sample": "BASELINE configs[0]: %d bodies (10x10x10 unit boxes + static ground box), b3CpuRigidBodyPipeline::stepSimulation x %d "
                      "(AABBs, b3DynamicBvhBroadphase, b3CpuNarrowPhase SAT contacts, integrate; the reference's CPU pipeline has no solver)" % (n, steps)}


def ncu_traffic(kernel):
    """dram bytes read + written per step by `kernel` (one launch; solverIterateKernel: its two phase launches together) from the
    committed ncu --set full summary of the CURRENT build (profiles/r02s2_ncu_summary.txt, written by tools/ncu_summary.py from
    gpurun_out/r02s2_full.ncu-rep); None when the file has no such line"""
    try:
        for line in open(os.path.join(ROOT, "profiles", "r02s2_ncu_summary.txt")):
            p = line.split()
            if len(p) >= 3 and p[0] == "traffic" and p[1] == kernel:
                return float(p[2])
    except Exception:
        pass
    return None


def ncu_sm_throughput(kernel):
    """sm__throughput (% of the peak issue rate) of `kernel` from the same committed ncu summary; None when absent"""
    try:
        for line in open(os.path.join(ROOT, "profiles", "r02s2_ncu_summary.txt")):
            if line.startswith(kernel + " ") and "sm throughput" in line:
                return float(line.split("sm throughput")[1].split("%")[0])
    except Exception:
        pass
    return None


def slab_leg(world_size, rank, local_rank, steps, warmup, per_rank_x=32, ny=64, nz=256):
    """BASELINE configs[4](ii): ONE box field (config-3 recipe) of 32*N x 64 x 256 bodies -- 4 194 304 at N = 8 -- cut into N slabs along x; every
    rank steps its slab, boundary bodies are mirrored on the neighbours by NCCL send/recv issued by the library itself on the world's stream
    (b3b200_slab_step: device-side counts, fixed-capacity messages, no host synchronisation).  Returns the dict of the "slab" key (rank 0)."""
    import torch
    import torch.distributed as dist
    from bullet3_b200 import capi, scenes, slab

    nx = per_rank_x * world_size
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.stack([(((j + 1) & 1) + 2.2 * i).reshape(-1), (1.0 + 2.0 * j).reshape(-1), (((j + 1) & 1) + 2.2 * k).reshape(-1)], 1).astype(np.float32)
    n = len(pos)
    quat = np.tile(np.array(scenes.IDENT, np.float32), (n, 1))

    def shapes(world):
        return [world.register_convex_points(scenes.box_points(2000.0)), world.register_convex_points(scenes.box_points(1.0))]

    scene = dict(shapes=shapes, static=[((0.0, -2000.0, 0.0), scenes.IDENT, 0)], pos=pos, quat=quat, shape_slot=np.ones(n, np.int64))
    stream = torch.cuda.Stream()
    sw = slab.SlabWorld(scene, rank, world_size, local_rank, stream.cuda_stream, max_ghosts=3 * ny * nz, margin=3.0)
    sw.world.set_solver(capi.SOLVER_PGS, ITERS)
    sw.enable_c_exchange()
    sw.exchange()
    sw.step_n(DT, warmup)
    sw.world.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sw.step_n(DT, steps)
    e1.record(stream)
    sw.world.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    left, right = sw.last_halo_counts()
    halo = torch.tensor([float((left + right) * slab.HALO_RECORD)], dtype=torch.float64, device="cuda")
    dist.all_reduce(halo, op=dist.ReduceOp.SUM)
    ctr = sw.world.counters()
    cnt = torch.tensor([float(ctr[0]), float(ctr[1]), float(ctr[4])], dtype=torch.float64, device="cuda")
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    b = sw.world.bodies()
    finite = bool(np.isfinite(b["pos"][: sw.num_owned]).all())
    sw.world.close()
    return {"workload": "BASELINE configs[4](ii): one %d x %d x %d = %d-body box field (GpuBoxPlaneScene recipe) on a static ground box, %d slabs along x, "
                        "PGS %d iterations; %d warm-up + %d timed steps while the field falls and piles up" % (nx, ny, nz, n, world_size, ITERS, warmup, steps),
            "bodies": n, "ms_per_step": float(ms[0]), "bodies_steps_per_s": n / (float(ms[0]) * 1e-3), "halo_bytes_per_step": float(halo[0]),
            "exchange": "b3b200_slab_step: pack -> ncclSend/ncclRecv (called by the library, world's stream) -> unpack; fixed-capacity messages of %d records per side, "
                        "device-side counts, no host synchronisation" % sw.max_ghosts,
            "pairs": float(cnt[0]), "contacts": float(cnt[1]), "overflow_flags": int(cnt[2]), "finite": finite}


def batched_worlds_leg(world_size, local_rank, steps, warmup, worlds_per_gpu):
    """BASELINE configs[4](i): independent 256-box worlds batched into ONE b3b200 world per GPU (b3b200_set_current_world): one
    broadphase / narrowphase / solve for all of them, no communication between ranks"""
    import torch
    import torch.distributed as dist
    from bullet3_b200 import capi, scenes

    per_world = 257
    cfg = capi.default_config(worlds_per_gpu * per_world + 64)
    stream = torch.cuda.Stream()
    w = capi.World(cfg, device=local_rank, stream=stream.cuda_stream)
    scenes.batched_box_worlds(w, worlds_per_gpu)
    w.upload()
    w.set_solver(capi.SOLVER_PGS, ITERS)
    settle = 180
    w.step_n(DT, settle + warmup)
    w.synchronize()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    w.step_n(DT, steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ctr = w.counters()
    b = w.bodies()
    dyn = b["invMass"] != 0
    ok = bool(np.isfinite(b["pos"]).all() and (b["pos"][dyn, 1] > 0.5).all())
    n = worlds_per_gpu * per_world
    w.close()
    return {"workload": "BASELINE configs[4](i): %d independent worlds per GPU (8 x 4 x 8 cubes, GpuBoxPlaneScene recipe, + 1 static ground box each, all at the same "
                        "coordinates) batched into one b3b200 world per GPU; PGS %d iterations; settled %d steps" % (worlds_per_gpu, ITERS, settle),
            "worlds": worlds_per_gpu * world_size, "bodies": n * world_size, "ms_per_step": float(ms[0]), "bodies_steps_per_s": n * world_size / (float(ms[0]) * 1e-3),
            "world_steps_per_s": worlds_per_gpu * world_size / (float(ms[0]) * 1e-3), "pairs_per_gpu": int(ctr[0]), "contacts_per_gpu": int(ctr[1]),
            "batches": int(ctr[2]), "cross_block_batches": int(ctr[3]), "overflow_flags": int(ctr[4]), "resting_on_their_grounds": ok}


# ---------------------------------------------------------------------------------- main
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return
        value, ms, desc, used = run_reference_arm(a.bodies_side, a.steps, a.warmup, a.ref_bodies)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "bodies*steps/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world_size),
            "cpu_baseline": {"value": value, "unit": "bodies*steps/s", "cores": used, "kind": "reference", "sample": desc, "host_cores": cores},
            "e2e": {"value": value, "unit": "bodies*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from bullet3_b200 import capi, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    side = a.bodies_side
    stream = torch.cuda.Stream()
    w = capi.World(bench_config(capi, side), device=local_rank, stream=stream.cuda_stream)
    scenes.bench_config4_scene(w, *scene_dims(side))
    w.upload()
    w.set_solver(capi.SOLVER_PGS, ITERS)
    nbodies = w.num_bodies
    w.step_n(DT, SETTLE_STEPS)
    w.synchronize()

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    w.step_n(DT, a.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = capi.lib().b3b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    w.step_n(DT, a.steps)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = capi.lib().b3b200_launch_count() - launches0
    clocks = sampler.stop()

    # ---- per-stage timing + counters (separate pass, CUDA events between stages on the same stream)
    w.enable_stage_timing(True)
    stage = np.zeros(8)
    nst = 16
    for _ in range(nst):
        w.step(DT)
        stage += w.stage_timings()
    stage /= nst
    ctr = w.counters()
    w.enable_stage_timing(False)
    P, Cn, nb = int(ctr[0]), int(ctr[1]), int(ctr[2])

    # ---- end to end through the C ABI with HOST buffers: upload body state, step, read it back
    host_bodies = w.bodies()
    pinned = torch.empty(host_bodies.nbytes, dtype=torch.uint8).pin_memory()
    hb = np.frombuffer(pinned.numpy(), dtype=capi.rigid_body_t)
    hb[:] = host_bodies
    e2e_steps = max(10, a.steps // 2)  # (long enough that the pipe's fill and drain -- one upload, one download -- do not show)
    for _ in range(3):
        w.write_bodies(hb)
        w.step(DT)
        check = capi.lib().b3b200_readback_bodies(w.h, capi.ptr(hb), len(hb))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        w.write_bodies(hb)
        w.step(DT)
        check = capi.lib().b3b200_readback_bodies(w.h, capi.ptr(hb), len(hb))
        assert check == 0
    barrier()
    e2e_blocking_s = time.perf_counter() - t0
    # the same three operations per step through the pipelined entry point: the upload of step c overlaps the compute of step
    # c - 1, the download overlaps step c + 1 (two page-locked input and two output buffers; every step uploads a full body
    # state and downloads the stepped one)
    pins = [torch.empty(host_bodies.nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    hin = [np.frombuffer(p.numpy(), dtype=capi.rigid_body_t) for p in pins[:2]]
    hout = [np.frombuffer(p.numpy(), dtype=capi.rigid_body_t) for p in pins[2:]]
    for h in hin:
        h[:] = hb
    for i in range(4):
        w.step_host_async(DT, hin[i & 1], hout[i & 1])
    w.step_host_wait()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        w.step_host_async(DT, hin[i & 1], hout[i & 1])
    w.step_host_wait()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.isfinite(hout[0]["pos"]).all() and np.isfinite(hout[1]["pos"]).all()

    # ---- max over ranks
    ms_t = torch.tensor([ms_total, e2e_s, e2e_blocking_s], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, e2e_blocking_s = float(ms_t[0]), float(ms_t[1]), float(ms_t[2])
    ms_per_step = ms_total / a.steps
    value = nbodies * world_size * a.steps / (ms_total * 1e-3)
    e2e_value = nbodies * world_size * e2e_steps / e2e_s

    if rank == 0:
        peak, peak_src = peaks()
        I = ITERS
        # SURVEY 8(d) algorithmic bytes per stage
        stage_bytes = {
            "broadphase": 168.0 * nbodies + 16.0 * P,
            "narrowphase": 176.0 * P + 112.0 * Cn,
            "solver_setup": 608.0 * Cn,
            "solver_iterate": 2.0 * I * 192.0 * Cn + 96.0 * nbodies,
            "integrate_aabb": 160.0 * nbodies + 32.0 * nbodies,
        }
        stage_ms = {"broadphase": stage[1], "narrowphase": stage[2], "solver_setup": stage[3], "solver_iterate": stage[4], "integrate_aabb": stage[5]}
        stages = {k: {"ms": float(stage_ms[k]), "alg_bytes": stage_bytes[k], "gbs": stage_bytes[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else 0.0,
                      "frac_of_hbm_peak": stage_bytes[k] / (stage_ms[k] * 1e-3) / 1e9 / peak if stage_ms[k] > 0 else 0.0} for k in stage_ms}
        # dominant KERNEL: the SAT kernel (timed alone by its own pair of events) or the solver iteration kernel (a
        # single-kernel stage).  Algorithmic bytes per launch: SAT = 112 B per work item (16 B item +
        # two 32-B pose records + 32 B appended item and axis); iterations = 2*I*192*C + 96*N.
        # `traffic` = dram bytes read + written per launch of that kernel, parsed from the committed ncu --set full summary
        # of this build (profiles/r02s2_ncu_summary.txt); null when the summary has no line for the kernel.
        sat_items = int(ctr[7])
        kern = {"satKernel": {"ms": float(stage[7]), "alg_bytes": 112.0 * sat_items,
                              "note": "FP32-issue bound, not HBM bound (ncu: sm__throughput 72 % of peak issue rate, L1 hit 97 %, 36 MB of DRAM traffic)"},
                "solverIterateKernel": {"ms": float(stage[4]), "alg_bytes": stage_bytes["solver_iterate"],
                                        "note": "latency bound: per pass %d grid barriers (batches of contacts between blocks) + the per-block batch loop in shared memory" % int(ctr[3])}}
        domk = max(kern, key=lambda k: kern[k]["ms"])
        dk = kern[domk]
        dk_gbs = dk["alg_bytes"] / (dk["ms"] * 1e-3) / 1e9 if dk["ms"] > 0 else 0.0
        out = {
            "metric": METRIC, "value": value, "unit": "bodies*steps/s", "n_gpus": world_size, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a, world_size),
            "counts": {"bodies": nbodies, "pairs": P, "contacts": Cn, "batches": nb, "cross_block_batches": int(ctr[3]), "overflow_flags": int(ctr[4])},
            "stages": stages,
            "roofline": {"bound": "hbm", "kernel": domk, "achieved": dk_gbs, "peak": peak, "unit": "GB/s", "frac": dk_gbs / peak,
                         "traffic": ncu_traffic(domk), "sm_throughput_pct_ncu": ncu_sm_throughput(domk), "peak_source": peak_src, "kernel_ms": dk["ms"], "alg_bytes_per_launch": dk["alg_bytes"],
                         "note": "dominant kernel, timed live with CUDA events on the world's stream; " + dk["note"],
                         "other_kernel": {k: {"ms": v["ms"], "gbs": v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0, "traffic": ncu_traffic(k)}
                                          for k, v in kern.items() if k != domk}},
            "e2e": {"value": e2e_value, "unit": "bodies*steps/s", "h2d_bytes_per_step": int(host_bodies.nbytes) * world_size,
                    "d2h_bytes_per_step": int(host_bodies.nbytes) * world_size, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "path": "b3b200_step_host_async(pinned host AoS in, pinned host AoS out) x steps, b3b200_step_host_wait: every step uploads all bodies, steps, "
                            "downloads all bodies; copies of neighbouring steps overlap the compute (two staging slots, own copy streams)",
                    "blocking": {"value": nbodies * world_size * e2e_steps / e2e_blocking_s, "ms_per_step": e2e_blocking_s / e2e_steps * 1e3,
                                 "path": "b3b200_write_bodies (pinned host AoS) -> b3b200_step -> b3b200_readback_bodies, each blocking (the reference's host loop)"}},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not a.no_cpu_baseline and world_size == 1:
            # the CPU baselines north_star names, on this box's host cores, bounded samples (reported, not a target)
            tables = w.tables()
            keep = region_of(side, host_bodies, a.bt2_bodies)
            sel = np.concatenate([[0], keep + 1])
            bt = run_bt2mt(tables, host_bodies[sel], 2, 5, 0)
            extra = []
            if bt:
                out["cpu_baseline"] = {"value": bt["bodies"] * 1e3 / bt["ms_per_step"], "unit": "bodies*steps/s", "cores": bt["threads"], "kind": "reference",
                                       "ms_per_step_sample": bt["ms_per_step"],
                                       "sample": "Bullet 2 btDiscreteDynamicsWorldMt + btDbvtBroadphase + btCollisionDispatcherMt(40) + btConstraintSolverPoolMt + "
                                                 "btSequentialImpulseConstraintSolverMt (CommonRigidBodyMTBase.cpp:518-567; unmodified reference sources, -DBT_THREADSAFE=1, "
                                                 "oracle/_ref/bt2mt_bench), %d threads, %d iterations, no sleeping: a connected %d-body region (all layers) of the settled "
                                                 "scene on the full mesh, 2 warm-up + %d timed steps, %d manifolds" % (bt["threads"], bt["iterations"], bt["bodies"], bt["steps"], bt["manifolds"])}
            c1 = run_cpu_pipeline_config1(600)
            if c1:
                extra.append(dict(c1, name="b3CpuRigidBodyPipeline, BASELINE configs[0]"))
                if "cpu_baseline" not in out:
                    out["cpu_baseline"] = c1
            out["cpu_baselines_other"] = extra
    batched_out = None
    if not a.no_batched:
        try:
            batched_out = batched_worlds_leg(world_size, local_rank, max(5, a.steps), max(3, a.warmup), a.worlds_per_gpu)
        except Exception as e:  # the headline line must still be printed
            batched_out = {"error": repr(e)[:300]}
    slab_out = None
    if world_size > 1 and not a.no_slab:
        # the multi-GPU mode WITH communication (the independent replicas above have none)
        w.close()
        del w
        torch.cuda.empty_cache()
        try:
            slab_out = slab_leg(world_size, rank, local_rank, max(5, a.steps), max(3, a.warmup))
        except Exception as e:  # the headline line must still be printed
            slab_out = {"error": repr(e)[:300]}
    if rank == 0:
        if batched_out is not None:
            out["batched_worlds"] = batched_out
        if slab_out is not None:
            out["slab"] = slab_out
        print(json.dumps(out))
    if world_size > 1:
        dist.destroy_process_group()


def bench_config(capi, side):
    """b3Config for the bench scene: 16 pairs / contacts per body like the reference, child-pair and triangle-pair capacities sized for it"""
    n = side ** 3 + 16
    cfg = capi.default_config(n)
    cfg["compoundPairCapacity"] = max(1 << 20, 24 * n)
    cfg["maxTriConvexPairCapacity"] = max(1 << 18, 4 * n)
    cfg["maxConvexVertices"] = 1 << 17  # the 257 x 257 heightfield vertices share the vertex / index tables with the hulls
    cfg["maxConvexIndices"] = 1 << 20
    return cfg


def scene_dims(side):
    """side^3 bodies laid out as a flat pile of LAYERS layers (64 -> 128 x 16 x 128)"""
    ny = min(LAYERS, side)
    nxz = int(round((side ** 3 / ny) ** 0.5))
    return nxz, ny, nxz


def workload_config(a, world_size):
    nx, ny, nz = scene_dims(a.bodies_side)
    return {"workload": "BASELINE configs[3] (GpuConvexScene): pile of %d x %d x %d = %d bodies (1/4 boxes, 1/4 tetrahedra, 1/4 seeded 8-32-vertex "
                        "hulls, 1/4 three-box compounds, random orientations) settled on a concave heightfield trimesh (256 x 256 quads at full "
                        "size); batched PGS %d iterations; dt 1/60; grid broadphase" % (nx, ny, nz, nx * ny * nz, ITERS),
            "worlds_per_gpu": 1, "parallelism": "independent worlds x%d" % world_size,
            "l2": "per-step working set (pairs+contacts+constraints+bodies) exceeds the 126 MB L2; no explicit flush",
            "settle_steps": SETTLE_STEPS}


if __name__ == "__main__":
    main()
