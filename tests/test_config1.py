"""BASELINE configs[0] / SURVEY 8(d) "Config 1": 1 000 unit boxes (10 x 10 x 10) on a static ground box, stepped by the
reference's own b3CpuRigidBodyPipeline (instantiated unmodified through oracle/_ref/libb3ref.so) for 600 steps at 1/60 s.
The CPU pipeline has no contact solver, so the state is re-seeded from it every step and what is compared per step is what
the survey lists: world AABBs, the (sorted) pair set, per-pair contact counts and the contacts themselves."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_api as oa  # noqa: E402
from bullet3_b200 import capi, scenes  # noqa: E402

needs_ref = pytest.mark.skipif(not oa.ref_available(), reason="compiled reference (oracle/_ref) not built")
DT = 1.0 / 60.0


def contact_table(c):
    """canonical order: by body pair (a box pair has one manifold)"""
    a, b = np.abs(c["bodyA"]), np.abs(c["bodyB"])
    return c[np.lexsort((b, a))]


def twin_world(ref, device):
    """a b3b200 world with the reference pipeline's own hull tables and bodies (config 1: collidable 0 = ground box, 1 = unit box)"""
    w = capi.World(capi.default_config(2048), device=device)
    cols = ref.register_shapes_into(w)
    assert cols == [0, 1]
    for b in ref.bodies():
        w.register_instance(0.0 if b["invMass"] == 0 else 1.0 / b["invMass"], tuple(b["pos"][:3]), tuple(b["quat"]), int(b["collidableIdx"]))
    return w


def exact_pairs(aabbs, bodies):
    small = np.nonzero(bodies["invMass"] != 0)[0].astype(np.int32)
    large = np.nonzero(bodies["invMass"] == 0)[0].astype(np.int32)
    _, pairs = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 1 << 20)
    return pairs


@needs_ref
def test_cpu_pipeline_stages_equal_the_oracle_on_config1():
    """pins the oracle's AABB / contact restatements against the reference pipeline OBJECT (not only its leaf functions), and
    shows its DBVT pair set is a superset of the exact one with the same contacts"""
    ref = oa.RefCpuPipeline(capi.default_config(2048))
    scenes.box_stack(ref, 10, 10, 10)
    host = twin_world(ref, -1)
    sh = oa.Shapes(host.tables())
    assert ref.num_bodies == 1001
    for step in range(40):
        bodies = ref.bodies()
        ref.stage(0)
        ref.stage(1)
        ref.stage(2)
        aabbs = oa.update_aabbs(oa.oracle(), "orc_", bodies, sh)
        ra = ref.aabbs()
        assert np.array_equal(aabbs["min"][:, :3].view(np.uint32), ra["min"][:, :3].view(np.uint32))
        assert np.array_equal(aabbs["max"][:, :3].view(np.uint32), ra["max"][:, :3].view(np.uint32))
        pairs = exact_pairs(aabbs, bodies)
        ex, fat = oa.sorted_pair_set(pairs), oa.sorted_pair_set(ref.pairs())
        assert len(np.setdiff1d(ex.view([("a", ex.dtype), ("b", ex.dtype)]), fat.view([("a", fat.dtype), ("b", fat.dtype)]))) == 0
        oc, _ = oa.convex_contacts_oracle(pairs, bodies, sh, -1.0, 0.0, 1 << 16)
        rc = ref.contacts()
        o, r = contact_table(oc), contact_table(rc)
        assert len(o) == len(r)
        assert np.array_equal(np.abs(o["bodyA"]), np.abs(r["bodyA"])) and np.array_equal(np.abs(o["bodyB"]), np.abs(r["bodyB"]))
        assert np.array_equal(o["worldNormalOnB"].view(np.uint32), r["worldNormalOnB"].view(np.uint32))
        npts = o["worldNormalOnB"][:, 3].astype(int)
        for k in range(4):
            m = npts > k
            assert np.array_equal(o["worldPosB"][m, k].view(np.uint32), r["worldPosB"][m, k].view(np.uint32))
        ref.stage(3, DT)
    assert len(rc) > 2000  # the pile is in contact (and, without a solver, sinking)


@pytest.mark.gpu
@needs_ref
@pytest.mark.timeout(600)
def test_config1_600_steps_against_the_reference_cpu_pipeline():
    ref = oa.RefCpuPipeline(capi.default_config(2048))
    scenes.box_stack(ref, 10, 10, 10)
    g = twin_world(ref, 0)
    g.upload()
    g.set_contact_clip(-1.0, 0.0)  # the shared CPU header's clip window (b3ContactConvexConvexSAT.h:320-321)
    total_contacts = 0
    for step in range(600):
        state = ref.bodies()
        g.write_bodies(state)  # identical input state every step
        ref.stage(0)
        ref.stage(1)
        ref.stage(2)
        g.update_aabbs()
        g.find_pairs()
        g.compute_contacts()
        ga, ra = g.aabbs(), ref.aabbs()
        assert np.array_equal(ga["min"][:, :3].view(np.uint32), ra["min"][:, :3].view(np.uint32)), step
        assert np.array_equal(ga["max"][:, :3].view(np.uint32), ra["max"][:, :3].view(np.uint32)), step
        gp = oa.sorted_pair_set(g.pairs())
        if step % 10 == 0:  # the exact pair set (brute force over the reference's AABBs), bit-exact
            assert np.array_equal(gp, oa.sorted_pair_set(exact_pairs(ga, state))), step  # (ga == ra bit for bit, and carries the body ids in w)
        fat = oa.sorted_pair_set(ref.pairs())  # the DBVT tests fat leaf volumes: a superset
        assert len(np.setdiff1d(gp.view([("a", gp.dtype), ("b", gp.dtype)]), fat.view([("a", fat.dtype), ("b", fat.dtype)]))) == 0, step
        gc, rc = contact_table(g.contacts()), contact_table(ref.contacts())
        assert len(gc) == len(rc), step
        assert np.array_equal(np.abs(gc["bodyA"]), np.abs(rc["bodyA"])) and np.array_equal(np.abs(gc["bodyB"]), np.abs(rc["bodyB"])), step
        # per-pair contact counts bit-exact; normals / points / depths 1e-5 relative (north_star)
        assert np.array_equal(gc["worldNormalOnB"][:, 3], rc["worldNormalOnB"][:, 3]), step
        assert np.allclose(gc["worldNormalOnB"][:, :3], rc["worldNormalOnB"][:, :3], rtol=1e-5, atol=1e-6), step
        npts = rc["worldNormalOnB"][:, 3].astype(int)
        for k in range(4):
            m = npts > k
            assert np.allclose(gc["worldPosB"][m, k], rc["worldPosB"][m, k], rtol=1e-5, atol=1e-5), (step, k)
        total_contacts += len(rc)
        ref.stage(3, DT)
    assert total_contacts > 600 * 1000


# ------------------------------------------------------------------ long trajectories: statistics against the reference's own run
def pile_statistics(bodies, contacts, ground_top=0.0):
    """what north_star asks to compare for chaotic long runs: kinetic + potential energy per body, rest height of the pile,
    speed distribution and penetration-depth statistics of the contact points"""
    dyn = bodies["invMass"] != 0
    v = bodies["linVel"][dyn, :3].astype(np.float64)
    w = bodies["angVel"][dyn, :3].astype(np.float64)
    y = bodies["pos"][dyn, 1].astype(np.float64)
    depth = np.concatenate([contacts["worldPosB"][contacts["worldNormalOnB"][:, 3] > k, k, 3] for k in range(4)]).astype(np.float64) if len(contacts) else np.zeros(1)
    return {"kinetic": float(0.5 * (v * v).sum(1).mean() + 0.5 * 0.2 * (w * w).sum(1).mean()), "potential": float(9.8 * y.mean()), "mean_y": float(y.mean()),
            "min_y": float(y.min()), "max_y": float(y.max()), "median_speed": float(np.median(np.linalg.norm(v, axis=1))),
            "depth_mean": float(depth.mean()), "depth_p01": float(np.percentile(depth, 1)), "depth_min": float(depth.min()), "contacts": len(contacts)}


@pytest.mark.gpu
@pytest.mark.timeout(900)
@pytest.mark.skipif(not oa.refcl_available(), reason="compiled reference host twins (oracle/_ref/libb3refcl.so) not built")
def test_trajectory_statistics_match_the_reference_run():
    """the same 8 x 6 x 8 box pile stepped 240 times by this library and by the reference's own b3GpuRigidBodyPipeline (unmodified,
    host twins): the dynamics are chaotic (and the batch orders differ), so energies, rest heights and penetration statistics
    are compared, not states.  Both solve with the reference's 4 PGS iterations."""
    steps = 240
    g = capi.World(capi.default_config(4096))
    scenes.box_plane_scene(g, 8, 6, 8)
    g.upload()
    g.set_solver(capi.SOLVER_PGS, 4)
    ref = oa.RefPipeline(capi.default_config(4096))
    scenes.box_plane_scene(ref, 8, 6, 8)
    ref.upload()
    for _ in range(steps):
        g.step(1 / 60)
    ref.step(1 / 60, steps)
    gb, rb = g.bodies(), ref.bodies()
    assert np.isfinite(gb["pos"]).all() and np.isfinite(rb["pos"]).all()
    # contacts of the final states through one narrowphase (this library's, checked bit-exact elsewhere): same measuring stick
    g.compute_contacts()
    gs = pile_statistics(gb, g.contacts())
    g.write_bodies(rb)
    g.update_aabbs()
    g.find_pairs()
    g.compute_contacts()
    rs = pile_statistics(rb, g.contacts())
    print("b3b200   ", gs)
    print("reference", rs)
    # both piles came to rest on the ground at the same height, nothing fell through or flew off
    assert gs["min_y"] > 0.9 and rs["min_y"] > 0.9
    assert abs(gs["mean_y"] - rs["mean_y"]) < 0.02 * rs["mean_y"]
    assert abs(gs["max_y"] - rs["max_y"]) < 0.05 * rs["max_y"]
    assert abs(gs["potential"] - rs["potential"]) < 0.02 * rs["potential"]
    # residual motion: the same jitter level (4 PGS iterations do not bring a 6-high pile to rest in either implementation)
    assert abs(gs["kinetic"] - rs["kinetic"]) < 0.3 * rs["kinetic"] + 0.02
    assert abs(gs["median_speed"] - rs["median_speed"]) < 0.1
    # penetration: same contact population and the same depth distribution (positional drift 0.005 allowed by the solver's ERP)
    assert abs(gs["contacts"] - rs["contacts"]) < 0.1 * rs["contacts"]
    assert abs(gs["depth_mean"] - rs["depth_mean"]) < 0.01
    assert abs(gs["depth_p01"] - rs["depth_p01"]) < 0.03
    assert gs["depth_min"] > -0.25 and rs["depth_min"] > -0.25
