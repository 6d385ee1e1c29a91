"""Pin the CPU restatement (oracle/oracle.cpp) against the UNMODIFIED reference
sources compiled into oracle/_ref/libb3ref.so -- bit for bit, on seeded inputs.
Runs without a GPU.  Skipped when oracle/_ref has not been built (it is built by
__graft_entry__.build() wherever /root/reference exists and travels to the GPU box)."""
import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi, scenes

pytestmark = pytest.mark.skipif(not oa.ref_available(), reason="oracle/_ref/libb3ref.so not built")


def make_world(n_side=5, seed=0, rotate=True, spacing=1.6, shapes="mixed"):
    """host-only world: ground + a jittered grid of overlapping convex bodies"""
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(4096), device=-1)
    scenes.add_ground_box(w, 50.0)
    cols = [w.register_convex_points(scenes.box_points(1.0))]
    if shapes == "mixed":
        cols.append(w.register_convex_points(scenes.tetra_points(1.0)))
        for nv in (6, 9, 12):
            cols.append(w.register_convex_points(scenes.random_hull_points(rng, nv, 0.8, 1.3)))
    for i in range(n_side):
        for j in range(n_side):
            for k in range(n_side):
                p = np.array([i, j, k], np.float64) * spacing + (rng.uniform(-0.2, 0.2, 3) if rotate else 0.0)
                p[1] += 0.9
                q = scenes.random_quat(rng) if rotate else scenes.IDENT
                w.register_instance(1.0, tuple(p), q, cols[int(rng.integers(0, len(cols)))])
    t = w.tables()
    bodies = t["bodies"].copy()
    rngv = np.random.default_rng(seed + 1)
    dyn = bodies["invMass"] != 0
    bodies["linVel"][dyn, :3] = rngv.normal(size=(dyn.sum(), 3)).astype(np.float32)
    bodies["angVel"][dyn, :3] = rngv.normal(size=(dyn.sum(), 3)).astype(np.float32) * 2
    return w, oa.Shapes(t), bodies, t["inertias"]


def all_pairs(bodies, shapes):
    aabbs = oa.update_aabbs(oa.oracle(), "orc_", bodies, shapes)
    small = np.nonzero(bodies["invMass"] != 0)[0].astype(np.int32)
    large = np.nonzero(bodies["invMass"] == 0)[0].astype(np.int32)
    n, pairs = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 1 << 20)
    return aabbs, small, large, pairs


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_update_aabbs_bit_exact(seed):
    w, shapes, bodies, _ = make_world(seed=seed)
    a = oa.update_aabbs(oa.oracle(), "orc_", bodies, shapes)
    b = oa.update_aabbs(oa.ref(), "ref_", bodies, shapes)
    assert np.array_equal(a["min"].view(np.uint32), b["min"].view(np.uint32))
    assert np.array_equal(a["max"].view(np.uint32), b["max"].view(np.uint32))
    assert np.array_equal(a["minIndex"], b["minIndex"])


@pytest.mark.parametrize("seed", [0, 3])
def test_brute_force_pairs_identical(seed):
    w, shapes, bodies, _ = make_world(seed=seed)
    aabbs, small, large, pairs = all_pairs(bodies, shapes)
    n2, pairs2 = oa.brute_force_pairs(oa.ref(), "ref_", aabbs, small, large, 1 << 20)
    assert len(pairs) == n2 and n2 > 100
    assert np.array_equal(pairs["x"], pairs2["x"]) and np.array_equal(pairs["y"], pairs2["y"])
    # the sweep variant used for full-size scenes finds the same set
    n3, pairs3 = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 1 << 20, fn="sweep_pairs")
    assert np.array_equal(oa.sorted_pair_set(pairs), oa.sorted_pair_set(pairs3))


def test_pairs_capacity_clamp():
    w, shapes, bodies, _ = make_world(seed=5)
    aabbs, small, large, pairs = all_pairs(bodies, shapes)
    n, clipped = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 10)
    assert n == len(pairs) and len(clipped) == 10


@pytest.mark.parametrize("seed", [0, 1])
def test_integrate_bit_exact(seed):
    w, shapes, bodies, _ = make_world(seed=seed)
    bodies["angVel"][3, :3] = (1e-5, 0, 0)      # Taylor branch
    bodies["angVel"][4, :3] = (300.0, 10.0, 0)  # clamped branch
    a = oa.integrate(oa.oracle(), "orc_", bodies, 1 / 60, 0.99, (0, -9.8, 0))
    b = oa.integrate(oa.ref(), "ref_", bodies, 1 / 60, 0.99, (0, -9.8, 0))
    for f in ("pos", "quat", "linVel", "angVel"):
        assert np.array_equal(a[f][:, :3].view(np.uint32), b[f][:, :3].view(np.uint32)), f
    assert np.array_equal(a["quat"].view(np.uint32), b["quat"].view(np.uint32))
    assert not np.array_equal(a["pos"], bodies["pos"])


def compare_contacts(a, b):
    assert len(a) == len(b)
    assert np.array_equal(a["bodyA"], b["bodyA"]) and np.array_equal(a["bodyB"], b["bodyB"])
    assert np.array_equal(a["worldNormalOnB"].view(np.uint32), b["worldNormalOnB"].view(np.uint32))
    npts = a["worldNormalOnB"][:, 3].astype(int)
    for k in range(4):
        m = npts > k
        assert np.array_equal(a["worldPosB"][m, k].view(np.uint32), b["worldPosB"][m, k].view(np.uint32)), k
    assert np.array_equal(a["frictionCmp"], b["frictionCmp"])
    return npts


@pytest.mark.parametrize("seed,rotate,shapes", [(0, True, "mixed"), (1, True, "mixed"), (2, False, "box"), (3, True, "box")])
def test_convex_contacts_bit_exact(seed, rotate, shapes):
    w, sh, bodies, _ = make_world(seed=seed, rotate=rotate, shapes=shapes, spacing=1.6 if rotate else 1.999)
    _, _, _, pairs = all_pairs(bodies, sh)
    ca, ia = oa.convex_contacts_oracle(pairs, bodies, sh, -1.0, 0.0, 1 << 16)
    cb, ib = oa.convex_contacts_ref(pairs, bodies, sh, 1 << 16)
    assert np.array_equal(ia, ib)
    npts = compare_contacts(ca, cb)
    assert len(ca) > 50 and npts.max() == 4 and npts.min() >= 1


def test_resting_stack_ties():
    """axis-aligned unit boxes exactly touching: every SAT axis ties, clipping is degenerate"""
    w = capi.World(capi.default_config(512), device=-1)
    scenes.box_stack(w, 4, 4, 4)
    t = w.tables()
    sh, bodies = oa.Shapes(t), t["bodies"]
    _, _, _, pairs = all_pairs(bodies, sh)
    ca, ia = oa.convex_contacts_oracle(pairs, bodies, sh, -1.0, 0.0, 1 << 16)
    cb, ib = oa.convex_contacts_ref(pairs, bodies, sh, 1 << 16)
    assert np.array_equal(ia, ib)
    compare_contacts(ca, cb)
    assert len(ca) > 0


@pytest.mark.parametrize("seed", [0, 4])
def test_build_constraints_bit_exact(seed):
    w, sh, bodies, inertias = make_world(seed=seed)
    _, _, _, pairs = all_pairs(bodies, sh)
    contacts, _ = oa.convex_contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 16)
    contacts["batchIdx"] = np.arange(len(contacts)) % 7
    a = oa.build_constraints(oa.oracle(), "orc_", contacts, bodies, inertias)
    b = oa.build_constraints(oa.ref(), "ref_", contacts, bodies, inertias)
    assert len(a) > 50
    npts = contacts["worldNormalOnB"][:, 3].astype(int)
    for f in ("linear", "worldPos", "center", "jacCoeffInv", "fJacCoeffInv", "appliedRambdaDt", "fAppliedRambdaDt"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
    for k in range(4):
        m = npts > k  # the reference leaves m_b[k] unset for unused rows
        assert np.array_equal(a["b"][m, k].view(np.uint32), b["b"][m, k].view(np.uint32))
    for f in ("bodyA", "bodyB", "batchIdx"):
        assert np.array_equal(a[f], b[f]), f


def canon_faces(verts, faces, indices, pts):
    """{cyclically-canonical tuple of INPUT point ids: plane}.  b3ConvexHullComputer re-derives
    vertices (shift/scale), so they are matched to the nearest input point."""
    out = {}
    for f in faces:
        idx = indices[f["indexOffset"]: f["indexOffset"] + f["numIndices"]]
        ids = [int(np.argmin(np.abs(pts - verts[i][:3]).sum(1))) for i in idx]
        k = ids.index(min(ids))
        out[tuple(ids[k:] + ids[:k])] = f["plane"].astype(np.float64)
    return out


@pytest.mark.parametrize("kind", ["box", "tetra", "hull8", "hull16", "prism"])
def test_hull_builder_matches_b3ConvexUtility(kind):
    import ctypes as C

    rng = np.random.default_rng(11)
    pts = {"box": scenes.box_points(1.0, 0.5, 2.0), "tetra": scenes.tetra_points(), "hull8": scenes.random_hull_points(rng, 8),
           "hull16": scenes.random_hull_points(rng, 16),
           "prism": np.array([[np.cos(a), y, np.sin(a)] for a in np.arange(6) * np.pi / 3 for y in (-1, 1)], np.float32)}[kind]
    w = capi.World(capi.default_config(16), device=-1)
    w.register_convex_points(pts)
    t = w.tables()
    cap = 256
    v = np.zeros((cap, 4), np.float32)
    e = np.zeros((cap, 4), np.float32)
    f = np.zeros(cap, capi.face_t)
    ix = np.zeros(cap * 8, np.int32)
    nv, nf, ni, ne = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    pts32 = np.ascontiguousarray(pts, np.float32)
    rc = oa.ref().ref_build_hull(capi.ptr(pts32), len(pts32), capi.ptr(v), C.byref(nv), capi.ptr(f), C.byref(nf), capi.ptr(ix), C.byref(ni),
                                 capi.ptr(e), C.byref(ne), cap)
    assert rc == 0
    assert nv.value == len(t["vertices"]) and nf.value == len(t["faces"]) and ne.value == len(t["unique_edges"])
    mine = canon_faces(t["vertices"], t["faces"], t["indices"], pts32)
    theirs = canon_faces(v[: nv.value], f[: nf.value], ix[: ni.value], pts32)
    assert set(mine.keys()) == set(theirs.keys())
    for k in mine:
        assert np.allclose(mine[k], theirs[k], atol=2e-3), (mine[k], theirs[k])  # the reference normal comes from float edge cross products

    def canon_edges(ed):
        out = []
        for d in ed:
            d = d[:3].astype(np.float64)
            k = np.argmax(np.abs(d) > 1e-3)
            out.append(-d if d[k] < 0 else d)
        return np.array(sorted(out, key=lambda d: tuple(np.round(d, 2))))

    a, b = canon_edges(t["unique_edges"]), canon_edges(e[: ne.value])
    for d in a:
        assert np.min(np.abs(b - d).sum(1)) < 2e-3
