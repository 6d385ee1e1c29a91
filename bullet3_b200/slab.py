"""Multi-GPU partitioning of the rigid-body step; one process per GPU, torch.distributed for plumbing.

Two modes, as BASELINE.json's north_star names them:

* batched independent worlds -- `shard_worlds`: world w goes to rank w // (W / G); no data-path collective.
* spatial slab decomposition of ONE scene -- `SlabWorld`: 1-D slabs along x; every rank owns the bodies whose
  initial centre lies in its slab (ownership is static), replicates the static bodies, and mirrors the
  neighbours' boundary bodies in ghost slots.  After every step the boundary bands are packed on the device
  (b3b200_halo_pack), exchanged with the left/right neighbour by NCCL send/recv over NVLink
  (torch.distributed.batch_isend_irecv on CUDA staging tensors) and scattered into the ghost slots
  (b3b200_halo_unpack).  Contacts between an owned body and a ghost are solved on both ranks; the owner's
  state wins at the next exchange.  With `spare_slots` > 0 ownership follows the bodies: a body whose centre has left
  the slab by more than `hysteresis` is handed to the neighbour (b3b200_halo_emigrate / b3b200_halo_adopt, same
  records and the same count-then-payload exchange) before the halo exchange of that step.

The reference has no multi-device code; none of this has a reference counterpart.
"""
import ctypes as C

import numpy as np

from . import capi

HALO_RECORD = 176


class slab_config_t(C.Structure):
    """b3b200_slab_config (include/b3b200_types.h)"""
    _fields_ = [("axis", C.c_int), ("lo", C.c_float), ("hi", C.c_float), ("margin", C.c_float), ("numOwned", C.c_int), ("firstGhostSlot", C.c_int),
                ("maxGhosts", C.c_int), ("globalIdBase", C.c_int), ("rank", C.c_int), ("numRanks", C.c_int)]


# ---------------------------------------------------------------------------------- pure partition logic (CPU-testable)
def shard_worlds(num_worlds, world_size):
    """contiguous world ranges per rank: [(first, count)] * world_size, sizes differ by at most one"""
    base, extra = divmod(int(num_worlds), int(world_size))
    out, first = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def slab_boundaries(x, world_size):
    """world_size+1 boundaries along x with (nearly) equal body counts per slab; outer ones are +-inf"""
    xs = np.sort(np.asarray(x, np.float64))
    b = [-np.inf]
    for r in range(1, world_size):
        k = (len(xs) * r) // world_size
        b.append(0.5 * (xs[k - 1] + xs[k]) if 0 < k < len(xs) else xs[min(k, len(xs) - 1)])
    b.append(np.inf)
    return np.asarray(b)


def slab_of(x, boundaries):
    """slab index of every body (by centre x)"""
    return np.clip(np.searchsorted(boundaries, np.asarray(x, np.float64), side="right") - 1, 0, len(boundaries) - 2)


def band_masks(aabb_min_x, aabb_max_x, lo, hi, margin):
    """which owned bodies must be mirrored on the left / right neighbour (numpy twin of haloPackKernel's test)"""
    left = aabb_min_x <= lo + margin
    right = aabb_max_x >= hi - margin
    return left, right


def exchange_buffers(sides, peer, counts, send, recv, cnt_send, cnt_recv, record_bytes):
    """The neighbour exchange protocol, backend-agnostic (NCCL on CUDA tensors, gloo on CPU tensors in the tests):
    first the record counts, then exactly count * record_bytes payload bytes per side.  Returns the received counts."""
    import torch.distributed as dist

    for s in sides:
        cnt_send[s].fill_(counts[s])
    ops = []
    for s in sides:
        ops.append(dist.P2POp(dist.isend, cnt_send[s], peer[s]))
        ops.append(dist.P2POp(dist.irecv, cnt_recv[s], peer[s]))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    rc = {s: int(cnt_recv[s].item()) for s in sides}
    ops = []
    for s in sides:
        if counts[s]:
            ops.append(dist.P2POp(dist.isend, send[s][: counts[s] * record_bytes], peer[s]))
        if rc[s]:
            ops.append(dist.P2POp(dist.irecv, recv[s][: rc[s] * record_bytes], peer[s]))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return rc


# ---------------------------------------------------------------------------------- GPU side
class SlabWorld:
    """One rank's share of a slab-decomposed scene.

    scene: dict with
        shapes:     callable(world) -> list of collidable indices (same order on every rank)
        static:     list of (position, orientation, shape_slot) replicated on every rank
        pos, quat:  (N,3|4) / (N,4) arrays of the dynamic bodies (identical on every rank)
        lin_vel:    optional (N,3) initial linear velocities
        shape_slot: (N,) index into the list returned by `shapes`
    """

    def __init__(self, scene, rank, world_size, device, stream, max_ghosts, margin=3.0, pairs_per_body=16, spare_slots=0, hysteresis=0.25,
                 max_migrants=4096):
        import torch

        self.torch = torch
        self.rank, self.world_size = rank, world_size
        self.margin = float(margin)
        pos = np.asarray(scene["pos"], np.float32)
        self.boundaries = slab_boundaries(pos[:, 0], world_size)
        slab = slab_of(pos[:, 0], self.boundaries)
        order = np.argsort(slab, kind="stable")  # global id = position in this order: slabs own contiguous id ranges
        self.global_order = order
        counts = np.bincount(slab, minlength=world_size)
        starts = np.concatenate([[0], np.cumsum(counts)])
        mine = order[starts[rank]: starts[rank + 1]]
        self.n_static = len(scene["static"])
        self.n_owned_dyn = len(mine)
        self.global_first = int(starts[rank])
        self.has_left, self.has_right = rank > 0, rank < world_size - 1
        self.max_ghosts = int(max_ghosts)
        n_ghost_slots = self.max_ghosts * (int(self.has_left) + int(self.has_right))
        self.spare_slots = int(spare_slots) if world_size > 1 else 0
        self.hysteresis = float(hysteresis)
        self.max_migrants = int(max_migrants)
        n_bodies = self.n_static + self.n_owned_dyn + self.spare_slots + n_ghost_slots
        cfg = capi.default_config(n_bodies + 16, pairs_per_body)
        self.world = capi.World(cfg, device=device, stream=stream)
        cols = scene["shapes"](self.world)
        for p, q, s in scene["static"]:
            self.world.register_instance(0.0, p, q, cols[s])
        quat = np.asarray(scene["quat"], np.float32)
        slot = np.asarray(scene["shape_slot"])
        self.world.register_instances(np.ones(len(mine), np.float32), pos[mine], quat[mine], np.asarray(cols, np.int32)[slot[mine]])
        # spare owned slots (for bodies that migrate in) and ghost slots: dynamic bodies parked far away until they are filled
        n_parked = self.spare_slots + n_ghost_slots
        first_parked = self.n_static + self.n_owned_dyn
        if n_parked:
            park = np.zeros((n_parked, 3), np.float32)
            park[:, 0] = 1.0e6 + 1024.0 * (first_parked + np.arange(n_parked))
            park[:, 1] = -1.0e6
            park[:, 2] = 1.0e6
            q0 = np.tile(np.array([0, 0, 0, 1], np.float32), (n_parked, 1))
            self.world.register_instances(np.ones(n_parked, np.float32), park, q0, np.full(n_parked, cols[int(slot[0])], np.int32))
        self.world.upload()
        self.num_owned = self.n_static + self.n_owned_dyn + self.spare_slots  # the owned REGION of the slot array
        self.first_ghost = self.num_owned
        self.free_slots = list(range(first_parked, first_parked + self.spare_slots))[::-1]
        if scene.get("lin_vel") is not None and self.n_owned_dyn:
            b = self.world.bodies()
            b["linVel"][self.n_static: first_parked, :3] = np.asarray(scene["lin_vel"], np.float32)[mine]
            self.world.write_bodies(b)
        if self.spare_slots:
            # spare slots start out static (parked) so that no stage treats them as bodies; every slot gets its global id
            b = self.world.bodies()
            b["invMass"][first_parked: first_parked + self.spare_slots] = 0.0
            self.world.write_bodies(b)
            ids = np.full(n_bodies, -1, np.int32)
            ids[self.n_static: first_parked] = self.global_first + np.arange(self.n_owned_dyn)
            capi.check(capi.lib().b3b200_halo_set_ids(self.world.h, capi.ptr(ids), n_bodies), "halo_set_ids")
        dev = torch.device("cuda", device)
        self.send = {s: torch.zeros(self.max_ghosts * HALO_RECORD, dtype=torch.uint8, device=dev) for s in ("left", "right")}
        self.recv = {s: torch.zeros(self.max_ghosts * HALO_RECORD, dtype=torch.uint8, device=dev) for s in ("left", "right")}
        self.cnt_send = {s: torch.zeros(1, dtype=torch.int32, device=dev) for s in ("left", "right")}
        self.cnt_recv = {s: torch.zeros(1, dtype=torch.int32, device=dev) for s in ("left", "right")}
        if self.spare_slots:
            self.mig_send = {s: torch.zeros(self.max_migrants * HALO_RECORD, dtype=torch.uint8, device=dev) for s in ("left", "right")}
            self.mig_recv = {s: torch.zeros(self.max_migrants * HALO_RECORD, dtype=torch.uint8, device=dev) for s in ("left", "right")}
        self.halo_bytes = 0
        self.migrated_out = self.migrated_in = 0
        self.c_driven = False
        assert capi.lib().b3b200_halo_record_size() == HALO_RECORD

    def _pack(self, side):
        lo, hi = float(self.boundaries[self.rank]), float(self.boundaries[self.rank + 1])
        big = 3.0e38
        if side == "left":
            a, b = -big, lo + self.margin  # AABB reaches below lo + margin
        else:
            a, b = hi - self.margin, big
        n = C.c_int(0)
        # pack's global id = globalIdBase + local index; local index of the first owned dynamic body is n_static
        capi.check(capi.lib().b3b200_halo_pack(self.world.h, 0, C.c_float(a), C.c_float(b), int(self.num_owned), int(self.global_first - self.n_static),
                                               int(self.rank), C.c_void_p(self.send[side].data_ptr()), self.max_ghosts, C.byref(n)), "halo_pack")
        return n.value

    def exchange(self):
        """boundary bands -> neighbours -> ghost slots (call after every step, and once before the first)"""
        if self.c_driven:
            capi.check(capi.lib().b3b200_slab_exchange(self.world.h), "slab_exchange")
            a, b = self.last_halo_counts()
            self.halo_bytes = (a + b) * HALO_RECORD
            return {"left": a, "right": b}, None
        torch = self.torch
        sides = [s for s, has in (("left", self.has_left), ("right", self.has_right)) if has]
        peer = {"left": self.rank - 1, "right": self.rank + 1}
        counts = {s: self._pack(s) for s in sides}
        rc = exchange_buffers(sides, peer, counts, self.send, self.recv, self.cnt_send, self.cnt_recv, HALO_RECORD)
        torch.cuda.current_stream().synchronize()
        slot = self.first_ghost
        for s in sides:
            capi.check(capi.lib().b3b200_halo_unpack(self.world.h, C.c_void_p(self.recv[s].data_ptr()), rc[s], slot, self.max_ghosts), "halo_unpack")
            slot += self.max_ghosts
        self.halo_bytes = sum(counts.values()) * HALO_RECORD
        return counts, rc

    def migrate(self):
        """hand the bodies whose centre left this slab (by more than the hysteresis) to the neighbour on that side"""
        if not self.spare_slots:
            return {}, {}
        torch = self.torch
        sides = [s for s, has in (("left", self.has_left), ("right", self.has_right)) if has]
        peer = {"left": self.rank - 1, "right": self.rank + 1}
        lo, hi = float(self.boundaries[self.rank]), float(self.boundaries[self.rank + 1])
        big = 3.0e38
        counts = {}
        for s in sides:
            a, b = (-big, lo - self.hysteresis) if s == "left" else (hi + self.hysteresis, big)
            n = C.c_int(0)
            slots = np.zeros(self.max_migrants, np.int32)
            capi.check(capi.lib().b3b200_halo_emigrate(self.world.h, 0, C.c_float(a), C.c_float(b), int(self.num_owned), int(self.rank),
                                                       C.c_void_p(self.mig_send[s].data_ptr()), self.max_migrants, capi.ptr(slots), C.byref(n)), "halo_emigrate")
            counts[s] = n.value
            self.free_slots.extend(int(x) for x in slots[: n.value])
        rc = exchange_buffers(sides, peer, counts, self.mig_send, self.mig_recv, self.cnt_send, self.cnt_recv, HALO_RECORD)
        torch.cuda.current_stream().synchronize()
        for s in sides:
            if rc[s] > len(self.free_slots):
                raise capi.B3Error("slab migration: %d bodies arrive, %d spare slots left" % (rc[s], len(self.free_slots)))
            slots = np.array([self.free_slots.pop() for _ in range(rc[s])], np.int32)
            if rc[s]:
                capi.check(capi.lib().b3b200_halo_adopt(self.world.h, C.c_void_p(self.mig_recv[s].data_ptr()), rc[s], capi.ptr(slots)), "halo_adopt")
        self.migrated_out += sum(counts.values())
        self.migrated_in += sum(rc.values())
        return counts, rc

    def step(self, dt=1.0 / 60.0):
        if self.c_driven:
            if self.spare_slots:
                self.world.step(dt)
                self.migrate()  # (hand-overs keep their host-side free-slot lists)
                capi.check(capi.lib().b3b200_slab_exchange(self.world.h), "slab_exchange")
            else:
                capi.check(capi.lib().b3b200_slab_step(self.world.h, C.c_float(dt)), "slab_step")
            return
        self.world.step(dt)
        self.migrate()
        self.exchange()

    # ---- the exchange driven from C: b3b200_slab_* call NCCL themselves on the world's stream (no torch tensors, no host sync)
    def enable_c_exchange(self):
        """every rank calls this once after construction: rank 0 makes the NCCL id, torch.distributed carries its 128 bytes to the
        others (any transport would do), b3b200_slab_init builds the communicator and the fixed-capacity message buffers"""
        import torch.distributed as dist

        ident = (C.c_char * 128)()
        if self.rank == 0:
            capi.check(capi.lib().b3b200_slab_unique_id(ident), "slab_unique_id")
        box = [bytes(ident.raw)]
        if self.world_size > 1:
            dist.broadcast_object_list(box, src=0)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        lo, hi = float(self.boundaries[self.rank]), float(self.boundaries[self.rank + 1])
        big = 3.0e38
        cfg = slab_config_t(axis=0, lo=max(lo, -big), hi=min(hi, big), margin=self.margin, numOwned=int(self.num_owned), firstGhostSlot=int(self.first_ghost),
                            maxGhosts=int(self.max_ghosts), globalIdBase=int(self.global_first - self.n_static), rank=int(self.rank), numRanks=int(self.world_size))
        capi.check(capi.lib().b3b200_slab_init(self.world.h, C.byref(cfg), ident), "slab_init")
        self.c_driven = True

    def step_n(self, dt, n):
        """n steps back to back; in the C-driven mode without hand-overs nothing returns to the host in between"""
        if self.c_driven and not self.spare_slots:
            capi.check(capi.lib().b3b200_slab_step_n(self.world.h, C.c_float(dt), int(n)), "slab_step_n")
        else:
            for _ in range(n):
                self.step(dt)

    def last_halo_counts(self):
        a, b = C.c_int(0), C.c_int(0)
        capi.check(capi.lib().b3b200_slab_last_counts(self.world.h, C.byref(a), C.byref(b)), "slab_last_counts")
        return a.value, b.value

    def global_ids(self):
        """global id of every local body (-1 for static and parked ghost slots)"""
        n = self.world.num_bodies
        g = np.zeros(n, np.int32)
        capi.check(capi.lib().b3b200_halo_ghost_ids(self.world.h, capi.ptr(g), n), "halo_ghost_ids")
        if self.spare_slots:
            return g.astype(np.int64)  # every slot carries its id (b3b200_halo_set_ids)
        ids = np.full(n, -1, np.int64)
        ids[self.n_static: self.num_owned] = self.global_first + np.arange(self.n_owned_dyn)
        ids[self.first_ghost:] = g[self.first_ghost:]
        return ids

    def owned_ids(self):
        """global ids of the dynamic bodies this rank owns right now"""
        g = self.global_ids()[self.n_static: self.num_owned]
        return g[g >= 0]
