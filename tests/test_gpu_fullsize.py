"""BASELINE.json's full-size workload (configs[3]: 262 144 bodies -- hulls, compounds, trimesh) checked through
size-independent properties, since the CPU oracle cannot finish it in seconds:
  * the pair list is duplicate-free, every pair's world AABBs overlap, and no pair has two static bodies
  * every contact has 1..4 points, a unit normal, depths inside the clip window, valid body indices
  * the batches satisfy the reference's invariant: no two constraints of a batch share a dynamic body
    (b3GpuPgsContactSolver.cpp:1385-1529) and cover every contact exactly once
  * a sample of the contacts is re-derived by the oracle from the same body state, bit for bit
  * nothing overflows and the state stays finite over further steps"""
import os
import sys

import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi, scenes

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_world():
    w = capi.World(bench.bench_config(capi, 64))
    scenes.bench_config4_scene(w, *bench.scene_dims(64))
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 10)
    w.step_n(1 / 60, 120)
    w.synchronize()
    yield w
    w.close()


@pytest.mark.timeout(600)
def test_full_size_pairs_contacts_and_batches(full_world):
    w = full_world
    t = w.tables()
    sh = oa.Shapes(t)
    bodies = w.bodies()
    w.write_bodies(bodies)  # the step below starts from exactly this AoS state
    w.update_aabbs()
    w.find_pairs()
    pairs = w.pairs()
    aabbs = w.aabbs()
    w.compute_contacts()
    contacts = w.contacts()
    ctr = w.counters()
    assert ctr[4] == 0, "overflow flags %d" % ctr[4]
    n = len(bodies)
    assert n == 262145 and len(pairs) > 1_000_000 and len(contacts) > 200_000

    # ---- pairs
    a, b = pairs["x"].astype(np.int64), pairs["y"].astype(np.int64)
    assert a.min() >= 0 and b.max() < n and np.all(a != b)
    key = np.minimum(a, b) * n + np.maximum(a, b)
    assert len(np.unique(key)) == len(key), "duplicate pairs"
    mn, mx = aabbs["min"][:, :3], aabbs["max"][:, :3]
    assert np.all((mn[a] <= mx[b]).all(1) & (mn[b] <= mx[a]).all(1)), "a pair whose AABBs do not overlap"
    inv = bodies["invMass"]
    assert not np.any((inv[a] == 0) & (inv[b] == 0))

    # ---- contacts
    npts = contacts["worldNormalOnB"][:, 3]
    assert np.all((npts >= 1) & (npts <= 4) & (npts == np.round(npts)))
    nrm = np.linalg.norm(contacts["worldNormalOnB"][:, :3], axis=1)
    assert np.abs(nrm - 1).max() < 1e-4
    ca, cb = np.abs(contacts["bodyA"]), np.abs(contacts["bodyB"])
    assert ca.max() < n and cb.max() < n
    assert np.all(inv[ca[contacts["bodyA"] < 0]] == 0) and np.all(inv[cb[contacts["bodyB"] < 0]] == 0)  # the sign bit marks static bodies
    assert np.all(contacts["bodyA"][(inv[ca] == 0) & (ca != 0)] < 0)
    for k in range(4):
        m = npts > k
        assert np.all(contacts["worldPosB"][m, k, 3] <= 0.02 + 1e-6)
    types = sh.collidables["shapeType"][bodies["collidableIdx"]]
    assert (types[ca] == capi.SHAPE_CONCAVE_TRIMESH).sum() > 10_000 and (types[cb] == capi.SHAPE_COMPOUND).sum() > 10_000

    # ---- a sample of pairs re-derived by the oracle (all pair types of the scene), bit for bit
    rng = np.random.default_rng(0)
    body_sample = rng.choice(np.arange(1, n), 300, replace=False)
    sel = np.isin(a, body_sample) | np.isin(b, body_sample)
    sub = pairs[sel]
    o_rest = oa.contacts_oracle(sub, bodies, sh, -1e30, 0.02, 1 << 18)
    o_mesh, _ = oa.concave_contacts_oracle(sub, bodies, sh, aabbs, 1 << 18)
    o = np.concatenate([o_rest, o_mesh])
    pk = set(zip(sub["x"].tolist(), sub["y"].tolist()))
    gsel = np.array([(int(x), int(y)) in pk for x, y in zip(ca, cb)])
    g = contacts[gsel]

    def canon(c):
        keys = tuple(c["worldPosB"][:, k, j] for k in range(4) for j in range(4)) + tuple(c["worldNormalOnB"][:, j] for j in range(4))
        return c[np.lexsort(keys + (np.abs(c["bodyB"]), np.abs(c["bodyA"])))]

    g, o = canon(g), canon(o)
    assert len(g) == len(o) and len(o) > 500
    assert np.array_equal(np.abs(g["bodyA"]), np.abs(o["bodyA"])) and np.array_equal(np.abs(g["bodyB"]), np.abs(o["bodyB"]))
    assert np.array_equal(g["worldNormalOnB"].view(np.uint32), o["worldNormalOnB"].view(np.uint32))
    np_o = o["worldNormalOnB"][:, 3].astype(int)
    for k in range(4):
        m = np_o > k
        assert np.array_equal(g["worldPosB"][m, k].view(np.uint32), o["worldPosB"][m, k].view(np.uint32)), k

    # ---- batches
    w.solver_setup()
    off = w.batches()
    cs = w.constraints()
    nb = len(off) - 1
    assert nb == w.counters()[2] and 1 <= nb <= 128
    real = cs["batchIdx"] >= 0
    assert real.sum() == len(contacts)
    for bi in range(nb):
        rows = cs[off[bi]: off[bi + 1]]
        rows = rows[rows["batchIdx"] >= 0]
        assert np.all(rows["batchIdx"] == bi)
        ids = np.concatenate([rows["bodyA"].astype(np.int64), rows["bodyB"].astype(np.int64)])
        ids = ids[inv[ids] != 0]
        assert len(np.unique(ids)) == len(ids), "batch %d uses a dynamic body twice" % bi


@pytest.mark.timeout(600)
def test_full_size_sap_equals_grid_and_the_narrowphase_is_idempotent(full_world):
    """two independent pair finders on the full scene give the same sorted pair set (uniform grid vs 1-axis sweep), and the
    narrowphase run twice on the same state gives the same contacts bit for bit (sorted: the append order is not fixed)"""
    w = full_world
    bodies = w.bodies()
    w.write_bodies(bodies)
    w.update_aabbs()

    def pair_keys():
        w.find_pairs()
        p = w.pairs()
        a, b = p["x"].astype(np.int64), p["y"].astype(np.int64)
        return np.sort(np.minimum(a, b) * len(bodies) + np.maximum(a, b))

    w.set_broadphase(capi.BP_GRID)
    grid = pair_keys()
    w.set_broadphase(capi.BP_SAP)
    sap = pair_keys()
    w.set_broadphase(capi.BP_GRID)
    assert len(grid) > 1_000_000 and np.array_equal(grid, sap)

    def contact_rows():
        w.find_pairs()
        w.compute_contacts()
        c = w.contacts()
        rows = np.concatenate([c["worldPosB"].view(np.uint32).reshape(len(c), -1), c["worldNormalOnB"].view(np.uint32).reshape(len(c), -1),
                               np.abs(c["bodyA"]).astype(np.uint32)[:, None], np.abs(c["bodyB"]).astype(np.uint32)[:, None],
                               c["childA"].astype(np.uint32)[:, None], c["childB"].astype(np.uint32)[:, None]], axis=1)
        return rows[np.lexsort(rows.T[::-1])]

    first, second = contact_rows(), contact_rows()
    assert len(first) > 200_000 and np.array_equal(first, second)


@pytest.mark.timeout(600)
def test_full_size_keeps_stepping(full_world):
    w = full_world
    w.step_n(1 / 60, 60)
    b = w.bodies()
    assert np.isfinite(b["pos"]).all() and np.isfinite(b["linVel"]).all() and np.isfinite(b["angVel"]).all()
    assert w.counters()[4] == 0
    dyn = b["invMass"] != 0
    assert np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)) < 2.0
    # the pile rests on the heightfield: hardly any body below the lowest point of the mesh (h >= -2)
    assert (b["pos"][dyn, 1] < -3.0).mean() < 0.01


@pytest.mark.timeout(600)
def test_full_size_batched_worlds_are_isolated_and_identical_at_the_start():
    """BASELINE configs[4](i) at full per-GPU size: 1 024 identical 257-body worlds at the same coordinates in one b3b200 world.
    Before any step every world must see exactly world 0's pairs and contacts (same local indices, same bits: the
    narrowphase is a pure function of a pair's two bodies); after stepping, no pair joins two worlds, the solver has no cross-block
    batch (blocks hold whole worlds) and every pile rests on its own ground."""
    nw = 1024
    w = capi.World(capi.default_config(nw * 257 + 64))
    per = scenes.batched_box_worlds(w, nw)
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 10)
    # a few steps first so that the piles touch their grounds (all worlds still evolve identically only in exact arithmetic of
    # the same ORDER, so compare at the start, where no solve has happened yet)
    w.update_aabbs()
    w.find_pairs()
    p = w.pairs()
    w.compute_contacts()
    c = w.contacts()
    wid = w.body_worlds()
    assert np.array_equal(wid[p["x"]], wid[p["y"]])
    lx, ly = p["x"] - wid[p["x"]] * per, p["y"] - wid[p["y"]] * per
    key = np.minimum(lx, ly).astype(np.int64) * per + np.maximum(lx, ly)
    order = np.lexsort((key, wid[p["x"]]))
    kw, ww = key[order], wid[p["x"]][order]
    n0 = int((ww == 0).sum())
    assert n0 > 200 and len(kw) == n0 * nw
    assert np.array_equal(kw.reshape(nw, n0), np.tile(kw[:n0], (nw, 1)))
    ca, cb = np.abs(c["bodyA"]), np.abs(c["bodyB"])
    cw = wid[ca]
    assert np.array_equal(cw, wid[cb])
    ckey = (ca - cw * per).astype(np.int64) * per + (cb - cw * per)
    order = np.lexsort((ckey, cw))
    cs = c[order]
    m0 = int((cw == 0).sum())
    assert m0 > 50 and len(cs) == m0 * nw
    for f in ("worldPosB", "worldNormalOnB"):
        v = cs[f].view(np.uint32).reshape(nw, -1)
        assert np.array_equal(v, np.tile(v[0], (nw, 1))), f
    w.step_n(1 / 60, 120)
    p = w.pairs()
    assert np.array_equal(wid[p["x"]], wid[p["y"]])
    ctr = w.counters()
    assert ctr[4] == 0 and ctr[3] == 0 and ctr[1] > 500 * nw
    b = w.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all() and (b["pos"][dyn, 1] > 0.5).all()
    w.close()
