"""2+-GPU check of body MIGRATION between slabs (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/slab_migrate_check.py
A cloud of boxes and tetrahedra drifts without gravity with random velocities along x, so bodies keep crossing the slab
boundaries (and bump into each other).  After K steps rank 0 checks
  * every global id is owned by exactly one rank (nothing lost, nothing duplicated) and bodies did change owner
  * every owned body's centre lies inside its owner's slab (+ hysteresis + one step of travel)
  * the trajectories follow the single-GPU run of the same scene (they differ by solver order at collisions only)
Prints MIGRATE OK on success; exit code != 0 otherwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bullet3_b200 import capi, scenes, slab  # noqa: E402

VMAX = 6.0


def make_scene(nx, ny, nz, seed=5):
    rng = np.random.default_rng(seed)
    pos = np.array([[3.0 * i, 2.0 + 3.0 * j, 3.0 * k] for i in range(nx) for j in range(ny) for k in range(nz)], np.float32)
    n = len(pos)
    quat = np.stack([scenes.random_quat(rng) for _ in range(n)]).astype(np.float32)
    slot = 1 + rng.integers(0, 2, n)
    vel = np.zeros((n, 3), np.float32)
    vel[:, 0] = rng.uniform(-VMAX, VMAX, n)
    vel[:, 1:] = rng.uniform(-0.3, 0.3, (n, 2))

    def shapes(world):
        return [world.register_convex_points(scenes.box_points(400.0)), world.register_convex_points(scenes.box_points(1.0)),
                world.register_convex_points(scenes.tetra_points(1.0))]

    return dict(shapes=shapes, static=[((0.0, -400.0, 0.0), scenes.IDENT, 0)], pos=pos, quat=quat, shape_slot=slot, lin_vel=vel)


def main():
    rank = int(os.environ.get("RANK", 0))
    ws = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, nz = int(os.environ.get("SLAB_NX", 40)), int(os.environ.get("SLAB_NY", 4)), int(os.environ.get("SLAB_NZ", 12))
    steps = int(os.environ.get("SLAB_STEPS", 240))
    scene = make_scene(nx, ny, nz)
    n = len(scene["pos"])
    stream = torch.cuda.current_stream().cuda_stream
    sw = slab.SlabWorld(scene, rank, ws, local, stream, max_ghosts=max(1024, 8 * ny * nz), margin=3.0, spare_slots=max(256, n // (2 * ws)), hysteresis=0.25)
    if os.environ.get("SLAB_C", "1") == "1":  # the exchange driven from C (NCCL called by the library); SLAB_C=0: torch.distributed orchestration
        sw.enable_c_exchange()
    sw.world.set_gravity((0.0, 0.0, 0.0))
    sw.world.set_solver(capi.SOLVER_PGS, 4)
    sw.exchange()
    for _ in range(steps):
        sw.step()
    ids = sw.global_ids()
    b = sw.world.bodies()
    own = np.nonzero(ids[sw.n_static: sw.num_owned] >= 0)[0] + sw.n_static
    state = (ids[own].copy(), b["pos"][own, :3].copy(), float(sw.boundaries[rank]), float(sw.boundaries[rank + 1]), sw.migrated_out, sw.migrated_in)
    states = [None] * ws
    dist.all_gather_object(states, state)
    ok = True
    if rank == 0:
        all_ids = np.concatenate([s[0] for s in states])
        moved_out, moved_in = sum(s[4] for s in states), sum(s[5] for s in states)
        uniq = len(np.unique(all_ids)) == n and len(all_ids) == n and all_ids.min() == 0 and all_ids.max() == n - 1
        print("owned ids: %d of %d, unique %s; bodies handed over %d (received %d); owners per rank %s" % (len(all_ids), n, uniq, moved_out, moved_in, [len(s[0]) for s in states]))
        ok &= uniq and moved_out == moved_in and moved_out > n // 50
        slack = sw.hysteresis + VMAX * 2 / 60.0 + 1e-3
        for r, (gid, pos, lo, hi, _, _) in enumerate(states):
            inside = (pos[:, 0] >= lo - slack) & (pos[:, 0] <= hi + slack)
            print("rank %d: slab [%.2f, %.2f], %d owned, %d outside the slab + slack" % (r, lo, hi, len(gid), int((~inside).sum())))
            ok &= bool(inside.all())
        # the same scene on one GPU
        w = capi.World(capi.default_config(n + 64), device=local)
        cols = scene["shapes"](w)
        for p, q, s in scene["static"]:
            w.register_instance(0.0, p, q, cols[s])
        order = sw.global_order
        w.register_instances(np.ones(n, np.float32), scene["pos"][order], scene["quat"][order], np.asarray(cols, np.int32)[scene["shape_slot"][order]])
        w.upload()
        w.set_gravity((0.0, 0.0, 0.0))
        w.set_solver(capi.SOLVER_PGS, 4)
        sb = w.bodies()
        sb["linVel"][len(scene["static"]):, :3] = scene["lin_vel"][order]
        w.write_bodies(sb)
        for _ in range(steps):
            w.step()
        single = w.bodies()["pos"][len(scene["static"]):, :3]
        pos = np.zeros((n, 3), np.float32)
        for gid, p, _, _, _, _ in states:
            pos[gid] = p
        d = np.linalg.norm(pos - single, axis=1)
        print("after %d steps: |dpos| vs single GPU median %.5f p90 %.4f max %.3f (bodies travelled up to %.1f)" % (
            steps, np.median(d), np.percentile(d, 90), d.max(), np.abs(single[:, 0] - scene["pos"][order][:, 0]).max()))
        ok &= np.median(d) < 0.05  # collisions amplify the solver-order differences between the two runs; the hard checks are the ownership invariants above
        print("MIGRATE OK" if ok else "MIGRATE FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
