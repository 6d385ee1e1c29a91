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
SETTLE_STEPS = 250  # untimed: the 16-layer lattice drops and comes to rest (contacts/body plateaus) before anything is measured
LAYERS = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies-side", type=int, default=64, help="scene is side^3 bodies (64 -> 262 144)")
    ap.add_argument("--cpu-sample-side", type=int, default=16, help="side of the bounded CPU-baseline sample scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
def cpu_pipeline_step(oa, lib, prefix, bodies, sh, inertias, iters):
    """one step of the path on the CPU: AABBs -> pairs -> SAT/clip contacts -> PGS -> integrate.
    `lib/prefix` select the compiled reference (ref_) or the oracle port (orc_) where both exist."""
    aabbs = oa.update_aabbs(lib, prefix, bodies, sh)
    small = np.nonzero(bodies["invMass"] != 0)[0].astype(np.int32)
    large = np.nonzero(bodies["invMass"] == 0)[0].astype(np.int32)
    # pair finding: sort-and-sweep port (the reference's host twin is O(N^2) brute force)
    _, pairs = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 16 * len(bodies), fn="sweep_pairs")
    cap = 24 * len(bodies)
    if prefix == "ref_":
        # convex x convex through the compiled reference header path; compound children and the trimesh through the port
        hull = sh.collidables["shapeType"][bodies["collidableIdx"]] == 3
        both = hull[pairs["x"]] & hull[pairs["y"]]
        c1, _ = oa.convex_contacts_ref(pairs[both], bodies, sh, cap)
        c2 = oa.contacts_oracle(pairs[~both], bodies, sh, -1e30, 0.02, cap)
    else:
        c1 = oa.contacts_oracle(pairs, bodies, sh, -1e30, 0.02, cap)
        c2 = c1[:0]
    c3, _ = oa.concave_contacts_oracle(pairs, bodies, sh, aabbs, cap)
    contacts = np.concatenate([c1, c2, c3])
    solved, _, _, _ = oa.oracle_pgs_step_velocities(contacts, bodies, inertias, 0, iters)
    return oa.integrate(lib, prefix, solved, DT, 0.99, (0.0, -9.8, 0.0)), len(pairs), len(contacts)


def make_cpu_sample(side, settle_on_gpu):
    """the bench scene recipe at side^3 bodies; state after SETTLE_STEPS steps when a GPU is there"""
    from bullet3_b200 import capi, scenes
    import oracle_api as oa

    dev = 0 if settle_on_gpu else -1
    w = capi.World(bench_config(capi, side), device=dev)
    scenes.bench_config4_scene(w, *scene_dims(side))
    t = w.tables()
    bodies = t["bodies"]
    if settle_on_gpu:
        w.upload()
        w.set_solver(capi.SOLVER_PGS, ITERS)
        w.step_n(DT, SETTLE_STEPS)
        bodies = w.bodies()
    w.close()
    return bodies, oa.Shapes(t), t["inertias"]


def run_cpu_arm(side, threads, steps, warmup, use_ref, settle_on_gpu):
    """`threads` independent sample worlds stepped concurrently (the path shards by world);
    returns (bodies*steps/s, ms per step of one sample world, description)"""
    import oracle_api as oa

    use_ref = use_ref and oa.ref_available()
    lib, prefix = (oa.ref(), "ref_") if use_ref else (oa.oracle(), "orc_")
    bodies0, sh, inertias = make_cpu_sample(side, settle_on_gpu)
    n = len(bodies0)
    state = [bodies0.copy() for _ in range(threads)]
    stats = [None] * threads

    def work(i, k):
        b = state[i]
        for _ in range(k):
            b, npairs, ncontacts = cpu_pipeline_step(oa, lib, prefix, b, sh, inertias, ITERS)
            stats[i] = (npairs, ncontacts)
        state[i] = b

    def run(k):
        ts = [threading.Thread(target=work, args=(i, k)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    run(warmup)
    el = run(steps)
    value = n * steps * threads / el
    desc = ("%s: %d independent %d-body sample worlds of the bench recipe (one per thread), %d steps each: "
            "AABBs, sweep pair finding, SAT+clip contacts (hulls, compound children, trimesh), coloured PGS %d it., integrate; %d pairs / %d contacts per world" %
            ("compiled reference (oracle/_ref: b3UpdateAabbs/b3ContactConvexConvexSAT/b3IntegrateTransforms) + oracle-port compound/trimesh contacts and solver" if use_ref
             else "oracle port", threads, n, steps, ITERS, stats[0][0], stats[0][1]))
    return value, el / steps * 1e3, desc, ("reference" if use_ref else "port")


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
        import torch

        settle = torch.cuda.is_available()
        value, ms, desc, kind = run_cpu_arm(a.cpu_sample_side, cores, max(1, a.steps // 10), max(1, a.warmup // 5), True, settle)
        print(json.dumps({
            "impl": "reference", "metric": "bodies*steps/s (256k convex scene recipe)", "value": value, "unit": "bodies*steps/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world_size),
            "cpu_baseline": {"value": value, "unit": "bodies*steps/s", "cores": cores, "kind": kind, "sample": desc},
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
    nst = 10
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
    e2e_steps = max(5, a.steps // 5)
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
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    ms_t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(ms_t[0]), float(ms_t[1])
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
        # `traffic` = dram bytes read + written by that kernel in the ncu --set full capture of this scene
        # (profiles/r01_np_config4_ncu_summary.txt), per launch.
        sat_items = int(ctr[7])
        kern = {"satKernel": {"ms": float(stage[7]), "alg_bytes": 112.0 * sat_items, "traffic": 64.1e6,
                              "note": "FP32-issue bound, not HBM bound: ncu sm__throughput 78 % of peak issue rate, L1 hit 94 %"},
                "solverIterateKernel": {"ms": float(stage[4]), "alg_bytes": stage_bytes["solver_iterate"], "traffic": 397.4e6,
                                        "note": "grid-barrier latency bound: 2*I*batches phases"}}
        domk = max(kern, key=lambda k: kern[k]["ms"])
        dk = kern[domk]
        dk_gbs = dk["alg_bytes"] / (dk["ms"] * 1e-3) / 1e9 if dk["ms"] > 0 else 0.0
        out = {
            "metric": "bodies*steps/s (256k convex scene)", "value": value, "unit": "bodies*steps/s", "n_gpus": world_size, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a, world_size),
            "counts": {"bodies": nbodies, "pairs": P, "contacts": Cn, "batches": nb, "colour_rounds": int(ctr[3]), "overflow_flags": int(ctr[4])},
            "stages": stages,
            "roofline": {"bound": "hbm", "kernel": domk, "achieved": dk_gbs, "peak": peak, "unit": "GB/s", "frac": dk_gbs / peak,
                         "traffic": dk["traffic"], "peak_source": peak_src, "kernel_ms": dk["ms"], "alg_bytes_per_launch": dk["alg_bytes"],
                         "note": "dominant kernel, timed live with CUDA events on the world's stream; " + dk["note"],
                         "other_kernel": {k: {"ms": v["ms"], "gbs": v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0}
                                          for k, v in kern.items() if k != domk}},
            "e2e": {"value": e2e_value, "unit": "bodies*steps/s", "h2d_bytes_per_step": int(host_bodies.nbytes) * world_size,
                    "d2h_bytes_per_step": int(host_bodies.nbytes) * world_size, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "path": "b3b200_write_bodies (pinned host AoS) -> b3b200_step -> b3b200_readback_bodies"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not a.no_cpu_baseline and world_size == 1:
            v, ms, desc, kind = run_cpu_arm(a.cpu_sample_side, 1, 2, 1, False, True)
            out["cpu_baseline"] = {"value": v, "unit": "bodies*steps/s", "cores": 1, "kind": kind, "sample": desc, "ms_per_step_sample": ms}
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
