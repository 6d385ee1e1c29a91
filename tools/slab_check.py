"""2+-GPU check of the slab decomposition (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py
Every rank steps its slab with halo exchange; rank 0 also steps the whole scene on one GPU and checks
  * step 1: the union over ranks of the broadphase pairs / contact body pairs (global ids) == the single-GPU sets
  * after K steps: the pile settles to the same heights (the two solutions differ only by batch order and by the
    boundary contacts being solved on both ranks)
Prints SLAB OK on success; exit code != 0 otherwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bullet3_b200 import capi, scenes, slab  # noqa: E402


def make_scene(nx, ny, nz, seed=3):
    rng = np.random.default_rng(seed)
    pos = np.array([[((j + 1) & 1) + 2.2 * i, 1.0 + 2.0 * j, ((j + 1) & 1) + 2.2 * k] for i in range(nx) for j in range(ny) for k in range(nz)], np.float32)
    n = len(pos)
    quat = np.tile(np.array(scenes.IDENT, np.float32), (n, 1))
    slot = 1 + rng.integers(0, 2, n)  # box or tetrahedron

    def shapes(world):
        return [world.register_convex_points(scenes.box_points(400.0)), world.register_convex_points(scenes.box_points(1.0)),
                world.register_convex_points(scenes.tetra_points(1.0))]

    return dict(shapes=shapes, static=[((0.0, -400.0, 0.0), scenes.IDENT, 0)], pos=pos, quat=quat, shape_slot=slot)


def xy(p):
    return np.stack([p["x"], p["y"]], 1)


def pair_set(pairs, ids):
    a, b = ids[pairs[:, 0]], ids[pairs[:, 1]]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    return set(zip(lo.tolist(), hi.tolist()))


def main():
    rank = int(os.environ.get("RANK", 0))
    ws = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, nz = int(os.environ.get("SLAB_NX", 48)), int(os.environ.get("SLAB_NY", 6)), int(os.environ.get("SLAB_NZ", 24))
    compare = os.environ.get("SLAB_COMPARE", "1") == "1"  # also step the whole scene on rank 0 (needs the memory for it)
    steps = int(os.environ.get("SLAB_STEPS", 120))
    scene = make_scene(nx, ny, nz)
    n = len(scene["pos"])
    stream = torch.cuda.current_stream().cuda_stream
    sw = slab.SlabWorld(scene, rank, ws, local, stream, max_ghosts=max(1024, 4 * ny * nz), margin=3.0)
    if os.environ.get("SLAB_C", "1") == "1":  # the exchange driven from C (NCCL called by the library); SLAB_C=0: torch.distributed orchestration
        sw.enable_c_exchange()
    sw.exchange()
    sw.world.step()
    ids = sw.global_ids()
    ids[: sw.n_static] = -2 - np.arange(sw.n_static)  # static bodies: the same negative id on every rank
    mine = pair_set(xy(sw.world.pairs()), ids)
    c = sw.world.contacts()
    mine_c = pair_set(np.stack([np.abs(c["bodyA"]), np.abs(c["bodyB"])], 1), ids) if len(c) else set()
    assert not any(-1 in p for p in mine), "a parked ghost slot produced a pair"
    gathered = [None] * ws
    dist.all_gather_object(gathered, (mine, mine_c, sw.halo_bytes))
    sw.exchange()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    sw.step_n(1.0 / 60.0, steps - 1)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / max(1, steps - 1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    b = sw.world.bodies()
    own = slice(sw.n_static, sw.num_owned)
    state = (sw.global_first, b["pos"][own].copy(), b["linVel"][own].copy())
    states = [None] * ws
    dist.all_gather_object(states, state)
    ok = True
    if rank == 0 and not compare:
        print("slab step %.3f ms (max over ranks, %d bodies on %d GPUs, halo %s bytes/step/rank); single-GPU comparison skipped" % (ms.item(), n, ws, [g[2] for g in gathered]))
        print("SLAB OK")
    if rank == 0 and compare:
        cfg = capi.default_config(n + 64)
        w = capi.World(cfg, device=local)
        cols = scene["shapes"](w)
        for p, q, s in scene["static"]:
            w.register_instance(0.0, p, q, cols[s])
        order = sw.global_order
        w.register_instances(np.ones(n, np.float32), scene["pos"][order], scene["quat"][order], np.asarray(cols, np.int32)[scene["shape_slot"][order]])
        w.upload()
        w.step()
        sid = np.arange(w.num_bodies) - len(scene["static"])
        sid[: len(scene["static"])] = -2 - np.arange(len(scene["static"]))
        single = pair_set(xy(w.pairs()), sid)
        c1 = w.contacts()
        single_c = pair_set(np.stack([np.abs(c1["bodyA"]), np.abs(c1["bodyB"])], 1), sid)
        union, union_c = set(), set()
        for m, mc, _ in gathered:
            union |= m
            union_c |= mc
        print("pairs: single %d, union over %d ranks %d, missing %d, extra %d" % (len(single), ws, len(union), len(single - union), len(union - single)))
        print("contact body pairs: single %d, union %d, missing %d, extra %d" % (len(single_c), len(union_c), len(single_c - union_c), len(union_c - single_c)))
        # pairs are exact AABB overlaps: identical.  Contacts: a ghost takes the other role (A/B) when its local index order
        # differs from the global order, which may flip grazing (depth ~ 0) pairs exactly like reordering bodies does in
        # the reference; nothing may be missing and the extras stay below 0.2 %
        ok &= union == single and not (single_c - union_c) and len(union_c - single_c) <= max(2, len(single_c) // 500)
        for _ in range(steps - 1):
            w.step()
        sb = w.bodies()[len(scene["static"]):]
        pos = np.zeros((n, 3), np.float32)
        vel = np.zeros((n, 3), np.float32)
        for first, p, v in states:
            pos[first: first + len(p)] = p[:, :3]
            vel[first: first + len(p)] = v[:, :3]
        d = np.linalg.norm(pos - sb["pos"][:, :3], axis=1)
        print("after %d steps: |dpos| median %.4f p99 %.4f max %.4f; mean height slab %.4f single %.4f; max speed slab %.3f single %.3f" % (
            steps, np.median(d), np.percentile(d, 99), d.max(), pos[:, 1].mean(), sb["pos"][:, 1].mean(), np.linalg.norm(vel, axis=1).max(),
            np.linalg.norm(sb["linVel"][:, :3], axis=1).max()))
        print("lowest body: slab %.4f single %.4f" % (pos[:, 1].min(), sb["pos"][:, 1].min()))
        ok &= abs(pos[:, 1].mean() - sb["pos"][:, 1].mean()) < 0.05 and np.median(d) < 0.05 and pos[:, 1].min() > sb["pos"][:, 1].min() - 0.2  # (the lowest body of the single-GPU run itself varies by 0.1 - 0.2 between runs)
        print("slab step %.3f ms (max over ranks, %d bodies on %d GPUs, halo %s bytes/step/rank)" % (ms.item(), n, ws, [g[2] for g in gathered]))
        print("SLAB OK" if ok else "SLAB FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
