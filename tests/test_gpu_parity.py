"""GPU parity tests: the sm_100a path, called through the C ABI (libb3b200.so),
against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star):
  * broadphase pair set (sorted) and per-pair contact counts: bit-exact
  * contact points / normals / depths: 1e-5 relative (they are in fact bit-exact)
  * post-solve velocities after one step, same batching + iteration count: 1e-4 relative
  * integration: 1e-6 relative (device sinf/cosf vs libm)
"""
import numpy as np
import pytest

import oracle_api as oa
import pairbench
from bullet3_b200 import capi, scenes

pytestmark = pytest.mark.gpu

G = (0.0, -9.8, 0.0)


# ------------------------------------------------------------------ primitives
@pytest.mark.parametrize("n", [0, 1, 31, 256, 257, 2048, 2049, 65537, 1 << 20])
def test_radix_sort_kv_matches_host_twin(n):
    rng = np.random.default_rng(n)
    d = np.zeros(n, capi.sort_data_t)
    d["key"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    if n > 100:
        d["key"][: n // 2] &= 0xFF  # many duplicates -> stability matters
    d["value"] = np.arange(n, dtype=np.uint32)
    g = d.copy()
    capi.check(capi.lib().b3b200_radix_sort_kv(0, capi.ptr(g), n), "radix_sort_kv")
    h = d.copy()
    oa.oracle().orc_radix_sort_kv(capi.ptr(h), n)
    assert np.array_equal(g["key"], h["key"]) and np.array_equal(g["value"], h["value"])


@pytest.mark.parametrize("n", [1, 1000, 300000])
def test_radix_sort_keys(n):
    rng = np.random.default_rng(n)
    k = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    g = k.copy()
    capi.check(capi.lib().b3b200_radix_sort_keys(0, capi.ptr(g), n), "radix_sort_keys")
    assert np.array_equal(g, np.sort(k))


@pytest.mark.parametrize("n", [1, 2, 1023, 4096, 4097, 262144 + 3])
def test_prefix_scan_matches_host_twin(n):
    import ctypes as C

    rng = np.random.default_rng(n)
    src = rng.integers(0, 1000, n).astype(np.uint32)
    dst = np.zeros(n, np.uint32)
    s = C.c_uint(0)
    capi.check(capi.lib().b3b200_prefix_scan_u32(0, capi.ptr(src), capi.ptr(dst), n, C.byref(s)), "scan")
    ref = np.zeros(n, np.uint32)
    s2 = C.c_uint(0)
    oa.oracle().orc_prefix_scan(capi.ptr(src), capi.ptr(ref), n, C.byref(s2))
    assert np.array_equal(dst, ref) and s.value == s2.value


def test_bound_search_count_and_fill():
    rng = np.random.default_rng(3)
    n, buckets = 100000, 256
    d = np.zeros(n, capi.sort_data_t)
    d["key"] = np.sort(rng.integers(0, buckets, n).astype(np.uint32))
    d["value"] = np.arange(n, dtype=np.uint32)
    c = np.zeros(buckets, np.uint32)
    capi.check(capi.lib().b3b200_bound_search_count(0, capi.ptr(d), n, capi.ptr(c), buckets), "bound_search")
    r = np.zeros(buckets, np.uint32)
    oa.oracle().orc_bound_search_count(capi.ptr(d), n, capi.ptr(r), buckets)
    assert np.array_equal(c, r) and c.sum() == n
    a = np.arange(1000, dtype=np.uint32)
    capi.check(capi.lib().b3b200_fill_u32(0, capi.ptr(a), 7, 100, 50), "fill")
    assert (a[50:150] == 7).all() and a[49] == 49 and a[150] == 150


@pytest.mark.parametrize("option", [0, 1, 2])
def test_bound_search_lower_upper_count_match_host_twin(option):
    """b3BoundSearchCL BOUND_LOWER / BOUND_UPPER / COUNT against its executeHost twin (b3BoundSearchCL.cpp:139-203)"""
    rng = np.random.default_rng(5 + option)
    nb = 300
    keys = np.sort(rng.choice(np.arange(nb), size=4000, p=None)).astype(np.uint32)
    keys = keys[(keys % 7) != 3]  # some buckets stay empty: their entries must be left untouched
    data = np.zeros(len(keys), capi.sort_data_t)
    data["key"] = keys
    data["value"] = rng.integers(0, 1 << 30, len(keys))
    init = rng.integers(0, 1000, nb).astype(np.uint32) if option < 2 else np.zeros(nb, np.uint32)
    g, r = init.copy(), init.copy()
    capi.check(capi.lib().b3b200_bound_search(0, capi.ptr(data), len(data), capi.ptr(g), nb, option), "bound_search")
    oa.refcl().refcl_bound_search(capi.ptr(data), len(data), capi.ptr(r), nb, option)
    assert np.array_equal(g, r)


@pytest.mark.parametrize("n", [1, 33, 1024, 1025, 5000])
def test_prefix_scan_float4_matches_host_twin(n):
    """b3PrefixScanFloat4CL (exclusive, xyz) against executeHost (b3PrefixScanFloat4CL.cpp:95-120); the host twin adds left to
    right, the device scans in a tree: equal to FP32 summation-order tolerance"""
    rng = np.random.default_rng(n)
    src = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    g, r = np.zeros_like(src), np.zeros_like(src)
    gs, rs = np.zeros(4, np.float32), np.zeros(4, np.float32)
    capi.check(capi.lib().b3b200_prefix_scan_float4(0, capi.ptr(src), capi.ptr(g), n, capi.ptr(gs)), "scan4")
    oa.refcl().refcl_prefix_scan_float4(capi.ptr(src), capi.ptr(r), n, capi.ptr(rs))
    assert np.allclose(g[:, :3], r[:, :3], rtol=1e-5, atol=2e-5 * np.sqrt(n))
    assert np.allclose(gs[:3], rs[:3], rtol=1e-5, atol=2e-5 * np.sqrt(n))
    assert (g[:, 3] == 0).all()


# ------------------------------------------------------------------ scene helpers
def gpu_world(n_side=6, seed=0, rotate=True, spacing=1.6, shapes="mixed", max_bodies=8192):
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(max_bodies))
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
    w.upload()
    t = w.tables()
    bodies = t["bodies"].copy()
    rngv = np.random.default_rng(seed + 1)
    dyn = bodies["invMass"] != 0
    bodies["linVel"][dyn, :3] = rngv.normal(size=(dyn.sum(), 3)).astype(np.float32)
    bodies["angVel"][dyn, :3] = rngv.normal(size=(dyn.sum(), 3)).astype(np.float32) * 2
    w.write_bodies(bodies)
    return w, oa.Shapes(t), bodies, t["inertias"]


def small_large(bodies):
    return (np.nonzero(bodies["invMass"] != 0)[0].astype(np.int32), np.nonzero(bodies["invMass"] == 0)[0].astype(np.int32))


def rel_close(a, b, tol):
    scale = np.maximum(np.abs(b), 1.0)
    return np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / scale) <= tol


# ------------------------------------------------------------------ AABBs + integrate
@pytest.mark.parametrize("seed", [0, 1])
def test_update_aabbs_bit_exact(seed):
    w, sh, bodies, _ = gpu_world(seed=seed)
    w.update_aabbs()
    g = w.aabbs()
    o = oa.update_aabbs(oa.oracle(), "orc_", bodies, sh)
    assert np.array_equal(g["min"].view(np.uint32), o["min"].view(np.uint32))
    assert np.array_equal(g["max"].view(np.uint32), o["max"].view(np.uint32))
    assert np.array_equal(g["minIndex"], o["minIndex"]) and np.array_equal(g["maxIndex"], o["maxIndex"])


def test_integrate_matches_oracle():
    w, sh, bodies, _ = gpu_world(seed=2)
    bodies["angVel"][3, :3] = (1e-5, 0, 0)
    bodies["angVel"][4, :3] = (300.0, 10.0, 0)
    w.write_bodies(bodies)
    w.integrate(1 / 60)
    g = w.bodies()
    o = oa.integrate(oa.oracle(), "orc_", bodies, 1 / 60, 0.99, G)
    for f in ("pos", "linVel", "angVel"):
        assert np.array_equal(g[f][:, :3].view(np.uint32), o[f][:, :3].view(np.uint32)), f
    assert rel_close(g["quat"], o["quat"], 1e-6)
    assert np.array_equal(g["collidableIdx"], o["collidableIdx"]) and np.array_equal(g["invMass"], o["invMass"])


# ------------------------------------------------------------------ broadphase
@pytest.mark.parametrize("kind", [capi.BP_GRID, capi.BP_SAP])
@pytest.mark.parametrize("seed,n_side", [(0, 6), (1, 10)])
def test_scene_pair_set_bit_exact(kind, seed, n_side):
    w, sh, bodies, _ = gpu_world(n_side=n_side, seed=seed)
    w.set_broadphase(kind)
    w.update_aabbs()
    w.find_pairs()
    g = oa.sorted_pair_set(w.pairs())
    aabbs = oa.update_aabbs(oa.oracle(), "orc_", bodies, sh)
    small, large = small_large(bodies)
    n, o = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 1 << 22)
    o = oa.sorted_pair_set(o)
    assert len(o) > 200
    assert np.array_equal(g, o)
    assert w.counters()[4] == 0  # no overflow


def run_standalone_bp(kind, aabbs, small, large, max_pairs):
    bp = capi.Broadphase(kind, len(aabbs), max_pairs)
    is_large = np.zeros(len(aabbs), bool)
    is_large[large] = True
    for i in range(len(aabbs)):
        (bp.create_large_proxy if is_large[i] else bp.create_proxy)(aabbs["min"][i], aabbs["max"][i], int(aabbs["minIndex"][i]))
    bp.write_aabbs()
    bp.calculate_pairs(max_pairs)
    return bp


@pytest.mark.parametrize("kind", [capi.BP_GRID, capi.BP_SAP])
@pytest.mark.parametrize("margin", [0.0, 2.0, 6.0])
def test_pairbench_64006_pair_set_bit_exact(kind, margin):
    """config 2: data/64006GPUAABBs.txt parsed as PairBench does.  The raw dump has no overlapping
    pair at all, so it is also run with every AABB inflated by `margin` per side."""
    aabbs, small, large = pairbench.load_pairbench_aabbs()
    aabbs["min"] -= np.float32(margin)
    aabbs["max"] += np.float32(margin)
    max_pairs = min(3 * 1024 * 1024, 16 * len(aabbs))  # PairBench.cpp:554-565
    bp = run_standalone_bp(kind, aabbs, small, large, max_pairs)
    g = oa.sorted_pair_set(bp.pairs())
    n, o = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, max_pairs, fn="sweep_pairs")
    assert n <= max_pairs
    o = oa.sorted_pair_set(o)
    assert bp.num_overlap() == n
    assert np.array_equal(g, o)
    if margin > 0:
        assert n > 1000


def test_broadphase_edge_cases():
    # empty
    bp = capi.Broadphase(capi.BP_GRID, 16, 64)
    bp.write_aabbs()
    bp.calculate_pairs(64)
    assert bp.num_overlap() == 0
    # two touching unit AABBs => 1 pair (test/b3DynamicBvhBroadphase/main.cpp), inclusive test
    for kind in (capi.BP_GRID, capi.BP_SAP):
        bp = capi.Broadphase(kind, 16, 64)
        bp.create_proxy((0, 0, 0), (1, 1, 1), 5)
        bp.create_proxy((1, 0, 0), (2, 1, 1), 9)
        bp.create_proxy((2.5, 0, 0), (3, 1, 1), 11)
        bp.write_aabbs()
        bp.calculate_pairs(64)
        p = bp.pairs()
        assert len(p) == 1 and (p["x"][0], p["y"][0]) == (5, 9) and p["z"][0] == -1
    # overflow clamps and keeps going (b3GpuGridBroadphase.cpp:164-168)
    bp = capi.Broadphase(capi.BP_GRID, 256, 10)
    for i in range(64):
        bp.create_proxy((0, 0, 0), (1, 1, 1), i)
    bp.write_aabbs()
    bp.calculate_pairs(10)
    assert bp.num_overlap() == 10
    # one huge dynamic AABB among small ones stays exact
    rng = np.random.default_rng(0)
    a = np.zeros(500, capi.aabb_t)
    c = rng.uniform(-20, 20, (500, 3)).astype(np.float32)
    a["min"], a["max"] = c - 0.5, c + 0.5
    a["min"][7], a["max"][7] = (-30, -1, -30), (30, 1, 30)
    a["minIndex"] = np.arange(500)
    small, large = np.arange(500, dtype=np.int32), np.zeros(0, np.int32)
    n, o = oa.brute_force_pairs(oa.oracle(), "orc_", a, small, large, 1 << 16)
    for kind in (capi.BP_GRID, capi.BP_SAP):
        bp = run_standalone_bp(kind, a, small, large, 1 << 16)
        assert np.array_equal(oa.sorted_pair_set(bp.pairs()), oa.sorted_pair_set(o))


# ------------------------------------------------------------------ narrowphase
def contact_table(contacts):
    """sort contacts by (|bodyA|, |bodyB|) for order-independent comparison"""
    order = np.lexsort((np.abs(contacts["bodyB"]), np.abs(contacts["bodyA"])))
    return contacts[order]


def assert_contacts_match(g, o):
    assert len(g) == len(o)
    g, o = contact_table(g), contact_table(o)
    assert np.array_equal(g["bodyA"], o["bodyA"]) and np.array_equal(g["bodyB"], o["bodyB"])
    # per-pair contact counts: bit-exact
    assert np.array_equal(g["worldNormalOnB"][:, 3], o["worldNormalOnB"][:, 3])
    assert rel_close(g["worldNormalOnB"][:, :3], o["worldNormalOnB"][:, :3], 1e-5)
    npts = o["worldNormalOnB"][:, 3].astype(int)
    for k in range(4):
        m = npts > k
        assert rel_close(g["worldPosB"][m, k], o["worldPosB"][m, k], 1e-5), k
    assert np.array_equal(g["frictionCmp"], o["frictionCmp"])
    assert np.array_equal(g["childA"], o["childA"]) and np.array_equal(g["childB"], o["childB"])
    exact = all(np.array_equal(g["worldPosB"][npts > k, k].view(np.uint32), o["worldPosB"][npts > k, k].view(np.uint32)) for k in range(4))
    return exact


@pytest.mark.parametrize("clip", [(-1e30, 0.02), (-1.0, 0.0)])
@pytest.mark.parametrize("seed,rotate,shapes,n_side", [(0, True, "mixed", 6), (1, True, "mixed", 8), (2, False, "box", 6), (3, True, "box", 6)])
def test_convex_contacts_match_oracle(clip, seed, rotate, shapes, n_side):
    w, sh, bodies, _ = gpu_world(n_side=n_side, seed=seed, rotate=rotate, shapes=shapes, spacing=1.6 if rotate else 1.999)
    w.set_contact_clip(*clip)
    w.update_aabbs()
    w.find_pairs()
    pairs = w.pairs()
    w.compute_contacts()
    g = w.contacts()
    o, pci = oa.convex_contacts_oracle(pairs, bodies, sh, clip[0], clip[1], 1 << 18)
    assert len(o) > 100
    exact = assert_contacts_match(g, o)
    assert exact, "contact points are expected to be bit-exact (no FMA contraction on either side)"
    # pairs[i].z = contact index or -1, same pairs
    pz = w.pairs()["z"]
    assert np.array_equal(pz >= 0, pci >= 0)


def _large_hull_pile(seed, n_side, verts, settle):
    """a pile of larger seeded hulls (and a few boxes / tetrahedra) settled on a ground box: resting face / edge / vertex
    contacts between hulls with up to 32 vertices -- the items whose edge x edge axes satKernel prunes hardest"""
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(8192))
    scenes.add_ground_box(w, 60.0)
    cols = [w.register_convex_points(scenes.box_points(0.6)), w.register_convex_points(scenes.tetra_points(1.0))]
    for nv in verts:
        cols.append(w.register_convex_points(scenes.random_hull_points(rng, nv, 0.6, 1.1)))
    for i in range(n_side):
        for j in range(n_side):
            for k in range(n_side):
                p = np.array([i, j, k], np.float64) * 1.7 + rng.uniform(-0.2, 0.2, 3)
                p[1] += 1.2
                pick = int(rng.integers(2, len(cols))) if rng.uniform() < 0.8 else int(rng.integers(0, 2))
                w.register_instance(1.0, tuple(p), scenes.random_quat(rng), cols[pick])
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 6)
    w.step_n(1 / 60, settle)
    return w


@pytest.mark.parametrize("seed,n_side,verts,settle", [(0, 8, (16, 24, 32, 32), 90), (1, 7, (32, 32, 28), 200), (2, 9, (10, 20, 32), 30)])
def test_large_hull_pile_contacts_bit_exact(seed, n_side, verts, settle):
    w = _large_hull_pile(seed, n_side, verts, settle)
    t = w.tables()
    sh = oa.Shapes(t)
    for _ in range(3):  # three consecutive states of the same pile
        bodies = w.bodies()
        w.write_bodies(bodies)
        w.update_aabbs()
        w.find_pairs()
        pairs = w.pairs()
        w.compute_contacts()
        g = w.contacts()
        o, pci = oa.convex_contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 18)
        assert len(o) > 300
        assert assert_contacts_match(g, o), "contact points are expected to be bit-exact"
        assert np.array_equal(contact_table(g)["worldNormalOnB"].view(np.uint32), contact_table(o)["worldNormalOnB"].view(np.uint32))
        assert np.array_equal(w.pairs()["z"] >= 0, pci >= 0)
        w.step_n(1 / 60, 7)
    w.close()


def test_resting_stack_contacts_ties():
    w = capi.World(capi.default_config(2048))
    scenes.box_stack(w, 6, 6, 6)
    w.upload()
    t = w.tables()
    sh, bodies = oa.Shapes(t), t["bodies"]
    w.update_aabbs()
    w.find_pairs()
    pairs = w.pairs()
    w.compute_contacts()
    o, _ = oa.convex_contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 18)
    assert len(o) > 300
    assert assert_contacts_match(w.contacts(), o)


def test_contact_capacity_clamp():
    cfg = capi.default_config(2048)
    cfg["maxContactCapacity"] = 50
    w = capi.World(cfg)
    scenes.box_stack(w, 6, 6, 6)
    w.upload()
    w.update_aabbs()
    w.find_pairs()
    w.compute_contacts()
    c = w.counters()
    assert c[1] == 50 and (c[4] & 2)
    # a plain step never reads its counters back; b3b200_synchronize is where the overrun is reported (the call itself succeeds)
    w.step(1 / 60)
    w.synchronize()
    assert "contacts" in capi.last_error() and "overrun" in capi.last_error()


# ------------------------------------------------------------------ solver
def check_batches(contacts, colours, bodies):
    """the reference's batch invariant: no two contacts of a batch share a dynamic body"""
    for b in np.unique(colours[colours >= 0]):
        seg = contacts[colours == b]
        ids = np.concatenate([np.abs(seg["bodyA"]), np.abs(seg["bodyB"])])
        ids = ids[bodies["invMass"][ids] != 0]
        assert len(ids) == len(np.unique(ids)), b


@pytest.mark.parametrize("colouring", [0, 1])
@pytest.mark.parametrize("seed,iters,n_side", [(0, 4, 7), (1, 10, 7), (2, 6, 14)])
def test_pgs_solver_matches_oracle(seed, iters, n_side, colouring):
    """two-level batching (blocks of bodies in shared memory + coloured contacts between blocks); n_side 14 = 2 745 bodies
    -> several blocks, i.e. the cross-block colours and their grid barriers run too"""
    w, sh, bodies, inertias = gpu_world(n_side=n_side, seed=seed)
    w.set_solver(capi.SOLVER_PGS, iters)
    w.set_colouring(colouring)
    w.update_aabbs()
    w.find_pairs()
    w.compute_contacts()
    contacts = w.contacts()
    assert len(contacts) > 200
    w.solver_setup()
    g_contacts = w.contacts()
    # --- "same batching": the device's batch assignment is taken as it is (its validity is checked here) and given to the oracle
    colours = g_contacts["batchIdx"].astype(np.int32)
    nb = int(colours.max()) + 1
    ctr = w.counters()
    assert colours.min() >= 0 and nb == ctr[2]
    if n_side >= 14:
        assert ctr[3] > 0  # there are contacts between blocks
    assert nb <= oa.colour_contacts(contacts, len(bodies), 0)[0] + 12
    check_batches(g_contacts, colours, bodies)
    off = w.batches()
    cs_all = w.constraints()
    assert len(off) - 1 == nb and off[-1] == len(cs_all) == len(contacts)
    for b in range(nb):
        assert (cs_all[off[b]: off[b + 1]]["batchIdx"] == b).all()
    assert np.array_equal(np.diff(off), np.bincount(colours, minlength=nb))
    # --- constraint rows equal the oracle's (order inside a batch is free -> sort by body ids)
    o_cs = oa.build_constraints(oa.oracle(), "orc_", g_contacts, bodies, inertias)

    def key(c):
        return np.lexsort((c["bodyB"], c["bodyA"], c["batchIdx"]))

    gs, os_ = cs_all[key(cs_all)], o_cs[key(o_cs)]
    for f in ("linear", "worldPos", "center", "jacCoeffInv", "b", "fJacCoeffInv"):
        assert rel_close(gs[f][..., :3] if f in ("worldPos", "center") else gs[f], os_[f][..., :3] if f in ("worldPos", "center") else os_[f], 1e-5), f
    # --- velocities after the iterations
    w.solver_iterate()
    g_bodies = w.bodies()
    o_bodies, _, _, _ = oa.oracle_pgs_step_velocities(contacts, bodies, inertias, 0, iters, colours=colours)
    assert rel_close(g_bodies["linVel"][:, :3], o_bodies["linVel"][:, :3], 1e-4)
    assert rel_close(g_bodies["angVel"][:, :3], o_bodies["angVel"][:, :3], 1e-4)
    moved = np.abs(g_bodies["linVel"][:, :3] - bodies["linVel"][:, :3]).max()
    assert moved > 0.1  # the solver did something
    # the applied impulses come back with the rows
    assert np.abs(w.constraints()["appliedRambdaDt"]).max() > 0


def test_pgs_reproducible_colouring_is_bit_identical_between_worlds():
    """colouring mode 0: the batch assignment (and with it every bit of the result) is a function of the contact array"""
    out = []
    for _ in range(2):
        w, sh, bodies, inertias = gpu_world(n_side=14, seed=7)
        w.set_solver(capi.SOLVER_PGS, 10)
        w.set_colouring(0)
        w.update_aabbs()
        w.find_pairs()
        w.compute_contacts()
        if out:
            w.set_contacts(out[0][1])  # the narrowphase appends with atomics: give both worlds the same contact ARRAY
        contacts = w.contacts()
        w.solve_contacts()
        out.append((w.bodies(), contacts, w.contacts()["batchIdx"].copy()))
    assert np.array_equal(out[0][2], out[1][2])
    for f in ("linVel", "angVel"):
        assert np.array_equal(out[0][0][f].view(np.uint32), out[1][0][f].view(np.uint32)), f


@pytest.mark.parametrize("seed,iters", [(0, 7), (3, 8)])
def test_jacobi_solver_matches_oracle(seed, iters):
    """config 3 solver: mass-splitting Jacobi, GPU kernel order (solverUtils.cl)"""
    w, sh, bodies, inertias = gpu_world(n_side=7, seed=seed)
    w.set_solver(capi.SOLVER_JACOBI, iters)
    w.update_aabbs()
    w.find_pairs()
    w.compute_contacts()
    contacts = w.contacts()
    assert len(contacts) > 200
    w.solve_contacts()
    g = w.bodies()
    o = oa.jacobi_solve(contacts, bodies, inertias, 0, iters)
    assert rel_close(g["linVel"][:, :3], o["linVel"][:, :3], 1e-4)
    assert rel_close(g["angVel"][:, :3], o["angVel"][:, :3], 1e-4)
    assert np.abs(g["linVel"][:, :3] - bodies["linVel"][:, :3]).max() > 0.1
    # static bodies untouched
    st = bodies["invMass"] == 0
    assert np.array_equal(g["linVel"][st], bodies["linVel"][st])


def test_jacobi_box_plane_scene_settles():
    """config 3 recipe (GpuBoxPlaneScene), reduced: boxes on the ground, Jacobi 8 iterations, SAP broadphase"""
    w = capi.World(capi.default_config(4096))
    scenes.box_plane_scene(w, 8, 4, 8)
    w.upload()
    w.set_solver(capi.SOLVER_JACOBI, 8)
    w.set_broadphase(capi.BP_SAP)
    for _ in range(200):
        w.step(1 / 60)
    b = w.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all()
    assert b["pos"][dyn, 1].min() > 0.8
    assert np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)) < 0.5


def transfer_batches(device_contacts, oracle_contacts):
    """batch index of every oracle contact = that of the device contact with the same bytes (everything but batchIdx)"""
    def canon(c):
        c = c.copy()
        c["batchIdx"] = 0
        return c.view(np.uint8).reshape(len(c), -1)

    d, o = canon(device_contacts), canon(oracle_contacts)
    do = np.lexsort(d.T[::-1])
    oo = np.lexsort(o.T[::-1])
    assert np.array_equal(d[do], o[oo]), "contact sets differ"
    out = np.zeros(len(o), np.int32)
    out[oo] = device_contacts["batchIdx"][do]
    return out


@pytest.mark.parametrize("colouring,n_side", [(1, 6), (1, 14), (0, 6)])
def test_full_step_matches_oracle_pipeline(colouring, n_side):
    """one whole b3b200_step == oracle stages chained on the CPU, in the DEFAULT configuration too: the oracle computes its own
    AABBs, pairs and contacts and solves them with the device's batch assignment ("same batching")"""
    w, sh, bodies, inertias = gpu_world(n_side=n_side, seed=5)
    w.set_solver(capi.SOLVER_PGS, 4)
    w.set_colouring(colouring)
    w.step(1 / 60)
    g = w.bodies()
    aabbs = oa.update_aabbs(oa.oracle(), "orc_", bodies, sh)
    small, large = small_large(bodies)
    _, pairs = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 1 << 22)
    contacts, _ = oa.convex_contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 18)
    colours = transfer_batches(w.contacts(), contacts)
    check_batches(contacts, colours, bodies)
    solved, _, _, _ = oa.oracle_pgs_step_velocities(contacts, bodies, inertias, 0, 4, colours=colours)
    o = oa.integrate(oa.oracle(), "orc_", solved, 1 / 60, 0.99, G)
    for f in ("pos", "quat", "linVel", "angVel"):
        assert rel_close(g[f][:, :3], o[f][:, :3], 1e-4), f
    # the fused integrate+AABB kernel left valid AABBs for the next step
    assert rel_close(w.aabbs()["min"], oa.update_aabbs(oa.oracle(), "orc_", g, sh)["min"], 1e-6)


def test_trajectory_energy_and_penetration():
    """chaotic long run: compare statistics, not states (north_star)"""
    w = capi.World(capi.default_config(4096))
    scenes.box_plane_scene(w, 8, 6, 8)
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 10)
    for _ in range(240):
        w.step(1 / 60)
    b = w.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all() and np.isfinite(b["linVel"]).all()
    # nothing fell through the ground (top at y=0) and the pile came to rest
    assert b["pos"][dyn, 1].min() > 0.5
    speed = np.linalg.norm(b["linVel"][dyn, :3], axis=1)
    assert np.median(speed) < 0.5
    w.compute_contacts()
    c = w.contacts()
    depth = np.concatenate([c["worldPosB"][c["worldNormalOnB"][:, 3] > k, k, 3] for k in range(4)])
    assert depth.min() > -0.25  # penetration stays small


# ------------------------------------------------------------------ planes and compounds
def shapes_world(seed=0, n=400, plane=True):
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(4096))
    box = w.register_convex_points(scenes.box_points(0.5))
    hull = w.register_convex_points(scenes.random_hull_points(rng, 10, 0.6, 0.9))
    ell = w.register_compound(scenes.compound_children(box, scenes.L_OFFSETS))
    if plane:
        pl = w.register_plane((0, 1, 0), 0.0)
        w.register_instance(0.0, (0, 0, 0), scenes.IDENT, pl)
    else:
        scenes.add_ground_box(w, 50.0)
    kinds = [box, hull, ell]
    side = 6.0
    for i in range(n):
        p = (rng.uniform(-side, side), rng.uniform(0.1, 4.0), rng.uniform(-side, side))
        w.register_instance(1.0, p, scenes.random_quat(rng), kinds[int(rng.integers(0, 3))])
    w.upload()
    t = w.tables()
    return w, oa.Shapes(t), t["bodies"], t["inertias"]


def contact_sort(c):
    return c[np.lexsort((c["worldPosB"][:, 0, 0], c["childB"], c["childA"], np.abs(c["bodyB"]), np.abs(c["bodyA"])))]


@pytest.mark.parametrize("seed,plane", [(0, True), (1, True), (2, False)])
def test_plane_and_compound_contacts_match_oracle(seed, plane):
    w, sh, bodies, _ = shapes_world(seed, plane=plane)
    w.update_aabbs()
    w.find_pairs()
    pairs = w.pairs()
    w.compute_contacts()
    g = contact_sort(w.contacts())
    o = contact_sort(oa.contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 18))
    assert len(o) > 200 and len(g) == len(o)
    for f in ("bodyA", "bodyB", "childA", "childB", "frictionCmp"):
        assert np.array_equal(g[f], o[f]), f
    assert np.array_equal(g["worldNormalOnB"].view(np.uint32), o["worldNormalOnB"].view(np.uint32))
    npts = o["worldNormalOnB"][:, 3].astype(int)
    for k in range(4):
        m = npts > k
        assert np.array_equal(g["worldPosB"][m, k].view(np.uint32), o["worldPosB"][m, k].view(np.uint32)), k
    types = sh.collidables["shapeType"][bodies["collidableIdx"]]
    ta, tb = types[np.abs(o["bodyA"])], types[np.abs(o["bodyB"])]
    assert ((ta == capi.SHAPE_COMPOUND) & (tb == capi.SHAPE_COMPOUND)).any() and ((ta == capi.SHAPE_COMPOUND) ^ (tb == capi.SHAPE_COMPOUND)).any()
    if plane:
        assert (ta == capi.SHAPE_PLANE).sum() > 10


def test_compound_scene_steps_and_rests_on_plane():
    w, sh, bodies, _ = shapes_world(3, n=120)  # random, initially interpenetrating pile
    w.set_solver(capi.SOLVER_PGS, 10)
    for _ in range(480):
        w.step(1 / 60)
    b = w.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all()
    assert b["pos"][dyn, 1].min() > -0.3  # nothing fell through the plane
    assert np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)) < 1.0


# ------------------------------------------------------------------ concave trimesh
def concave_world(seed=0, n=600, nq=24, amplitude=1.5, freq=0.5, drop=0.9, lo=-0.2):
    rng = np.random.default_rng(seed)
    cfg = capi.default_config(4096)
    cfg["maxTriConvexPairCapacity"] = 1 << 18
    w = capi.World(cfg)
    verts, tris = scenes.heightfield_mesh(nq, nq, cell=1.0, amplitude=amplitude, freq=freq)
    mesh = w.register_concave(verts, tris)
    w.register_instance(0.0, (0, 0, 0), scenes.IDENT, mesh)  # the mesh must be the lower body index (b3BvhTraversal.h:35)
    box = w.register_convex_points(scenes.box_points(0.5))
    hull = w.register_convex_points(scenes.random_hull_points(rng, 12, 0.5, 0.8))
    tet = w.register_convex_points(scenes.tetra_points(0.6))
    ell = w.register_compound(scenes.compound_children(box, scenes.L_OFFSETS))
    ball = w.register_sphere(0.5)
    kinds = [box, hull, tet, ell, ball]
    half = 0.5 * nq - 1.5
    for i in range(n):
        x, z = rng.uniform(-half, half, 2)
        h = amplitude * np.sin(freq * x) * np.cos(freq * z)
        w.register_instance(1.0, (x, h + rng.uniform(lo, drop), z), scenes.random_quat(rng), kinds[int(rng.integers(0, 5))])
    w.upload()
    t = w.tables()
    return w, oa.Shapes(t), t["bodies"]


def full_sort(c):
    keys = tuple(c["worldPosB"][:, k, j] for k in range(4) for j in range(4)) + tuple(c["worldNormalOnB"][:, j] for j in range(4))
    return c[np.lexsort(keys + (c["childB"], c["childA"], np.abs(c["bodyB"]), np.abs(c["bodyA"])))]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_concave_contacts_match_oracle(seed):
    w, sh, bodies = concave_world(seed)
    w.update_aabbs()
    w.find_pairs()
    pairs = w.pairs()
    aabbs = w.aabbs()
    w.compute_contacts()
    g = w.contacts()
    assert w.counters()[4] == 0
    o_mesh, ncand = oa.concave_contacts_oracle(pairs, bodies, sh, aabbs, 1 << 18)
    o_rest = oa.contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 18)
    assert len(o_mesh) > 300 and ncand > len(o_mesh)
    is_mesh = (np.abs(g["bodyA"]) == 0) | (np.abs(g["bodyB"]) == 0)  # (sphere x trimesh contacts carry the sphere as A)
    gm, gr = full_sort(g[is_mesh]), full_sort(g[~is_mesh])
    om, orr = full_sort(o_mesh), full_sort(o_rest)
    assert len(gm) == len(om) and len(gr) == len(orr)
    for a, b in ((gm, om), (gr, orr)):
        for f in ("bodyA", "bodyB", "childA", "childB", "frictionCmp"):
            assert np.array_equal(a[f], b[f]), f
        assert np.array_equal(a["worldNormalOnB"].view(np.uint32), b["worldNormalOnB"].view(np.uint32))
        npts = b["worldNormalOnB"][:, 3].astype(int)
        for k in range(4):
            m = npts > k
            assert np.array_equal(a["worldPosB"][m, k].view(np.uint32), b["worldPosB"][m, k].view(np.uint32)), k
    types = sh.collidables["shapeType"][bodies["collidableIdx"]]
    assert (types[np.abs(om["bodyB"])] == capi.SHAPE_COMPOUND).sum() > 20  # compound children against triangles are covered
    assert (types[np.abs(om["bodyA"])] == capi.SHAPE_SPHERE).sum() > 10  # and sphere x triangle


def test_concave_scene_settles_on_the_heightfield():
    w, sh, bodies = concave_world(5, n=400, drop=4.0, lo=1.0)  # every body starts above the (zero-thickness) surface
    w.set_solver(capi.SOLVER_PGS, 10)
    for _ in range(420):
        w.step(1 / 60)
    b = w.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all()
    x, z = b["pos"][dyn, 0], b["pos"][dyn, 2]
    inside = (np.abs(x) < 11) & (np.abs(z) < 11)
    ground = 1.5 * np.sin(0.5 * x) * np.cos(0.5 * z)
    assert inside.sum() > 200
    above = (b["pos"][dyn, 1][inside] - ground[inside]) > -0.35
    assert above.mean() > 0.97, above.mean()  # a zero-thickness mesh cannot recover a body squeezed through by the pile; nearly all must rest on it
    assert np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)[inside]) < 1.0


# ------------------------------------------------------------------ spheres
def sphere_world(seed=0, n=500, plane=True):
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(4096))
    box = w.register_convex_points(scenes.box_points(0.5))
    hull = w.register_convex_points(scenes.random_hull_points(rng, 12, 0.5, 0.8))
    s1 = w.register_sphere(0.45)
    s2 = w.register_sphere(0.8)
    ell = w.register_compound(scenes.compound_children(box, scenes.L_OFFSETS))
    if plane:
        w.register_instance(0.0, (0, 0, 0), scenes.IDENT, w.register_plane((0, 1, 0), 0.0))
    else:
        scenes.add_ground_box(w, 50.0)
    kinds = [box, hull, s1, s2, ell]
    for i in range(n):
        p = (rng.uniform(-5, 5), rng.uniform(0.0, 3.5), rng.uniform(-5, 5))
        w.register_instance(1.0, p, scenes.random_quat(rng), kinds[int(rng.integers(0, 5))])
    w.upload()
    t = w.tables()
    return w, oa.Shapes(t), t["bodies"]


@pytest.mark.parametrize("seed,plane", [(0, True), (1, False)])
def test_sphere_contacts_match_oracle(seed, plane):
    """sphere x convex follows the reference's host twin bit for bit; sphere x sphere and plane x sphere restate
    primitiveContacts.cl (no host twin exists) identically in the oracle and on the device"""
    w, sh, bodies = sphere_world(seed, plane=plane)
    w.update_aabbs()
    w.find_pairs()
    pairs = w.pairs()
    w.compute_contacts()
    g = full_sort(w.contacts())
    o = full_sort(oa.contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 18))
    assert len(g) == len(o) and len(o) > 300
    for f in ("bodyA", "bodyB", "childA", "childB", "frictionCmp"):
        assert np.array_equal(g[f], o[f]), f
    assert np.array_equal(g["worldNormalOnB"].view(np.uint32), o["worldNormalOnB"].view(np.uint32))
    npts = o["worldNormalOnB"][:, 3].astype(int)
    for k in range(4):
        m = npts > k
        assert np.array_equal(g["worldPosB"][m, k].view(np.uint32), o["worldPosB"][m, k].view(np.uint32)), k
    types = sh.collidables["shapeType"][bodies["collidableIdx"]]
    ta, tb = types[np.abs(o["bodyA"])], types[np.abs(o["bodyB"])]
    sp = capi.SHAPE_SPHERE
    assert ((ta == sp) & (tb == sp)).sum() > 20 and ((ta == sp) ^ (tb == sp)).sum() > 50
    assert (((ta == sp) & (tb == capi.SHAPE_COMPOUND)) | ((ta == capi.SHAPE_COMPOUND) & (tb == sp))).sum() > 10  # sphere x compound child
    if plane:
        assert ((ta == capi.SHAPE_PLANE) & (tb == sp)).sum() > 5


def test_spheres_and_boxes_settle_on_plane():
    w, sh, bodies = sphere_world(4, n=300)
    w.set_solver(capi.SOLVER_PGS, 10)
    for _ in range(420):
        w.step(1 / 60)
    b = w.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all()
    assert b["pos"][dyn, 1].min() > 0.1  # nothing sank into the plane (spheres rest at radius - drift; the random hull is flat)
    assert np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)) < 1.0


# ------------------------------------------------------------------ halo records (slab mode), two worlds on ONE device
def test_halo_pack_unpack_between_two_worlds():
    """b3b200_halo_pack on one world -> device buffer -> b3b200_halo_unpack into the ghost slots of another world:
    the ghosts carry the owner's state bit for bit, produce the same pairs as the original bodies would, and unused
    ghost slots stay parked."""
    import ctypes as C
    import torch

    rng = np.random.default_rng(0)
    n, n_ghost = 300, 256
    pos = np.stack([rng.uniform(-6, 6, n), rng.uniform(0.6, 3.0, n), rng.uniform(-3, 3, n)], 1).astype(np.float32)
    quat = np.array([scenes.random_quat(rng) for _ in range(n)], np.float32)
    left = pos[:, 0] < 0

    def make(owned_mask):
        w = capi.World(capi.default_config(2048))
        scenes.add_ground_box(w, 50.0)
        box = w.register_convex_points(scenes.box_points(0.5))
        hull = w.register_convex_points(scenes.random_hull_points(np.random.default_rng(5), 10, 0.5, 0.7))
        cols = np.where(np.arange(n) % 2 == 0, box, hull).astype(np.int32)
        idx = np.nonzero(owned_mask)[0]
        w.register_instances(np.ones(len(idx), np.float32), pos[idx], quat[idx], cols[idx])
        park = np.zeros((n_ghost, 3), np.float32)
        park[:, 0] = 1.0e6 + 1024.0 * np.arange(n_ghost)
        park[:, 1] = -1.0e6
        w.register_instances(np.ones(n_ghost, np.float32), park, np.tile(np.array(scenes.IDENT, np.float32), (n_ghost, 1)), np.full(n_ghost, box, np.int32))
        w.upload()
        return w, idx

    wl, idl = make(left)
    wr, idr = make(~left)
    L = capi.lib()
    rec = L.b3b200_halo_record_size()
    buf = torch.zeros(n_ghost * rec, dtype=torch.uint8, device="cuda")
    cnt = C.c_int(0)
    margin = 1.5
    # bodies of the left world whose AABB reaches x >= -margin go to the right world's ghost slots
    capi.check(L.b3b200_halo_pack(wl.h, 0, C.c_float(-margin), C.c_float(3.0e38), 1 + len(idl), 1000 - 1, 0, C.c_void_p(buf.data_ptr()), n_ghost, C.byref(cnt)), "pack")
    aabbs_l = wl.aabbs()
    want = np.nonzero(aabbs_l["max"][1: 1 + len(idl), 0] >= -margin)[0]
    assert cnt.value == len(want) and 5 < cnt.value < n_ghost
    first_ghost = 1 + len(idr)
    capi.check(L.b3b200_halo_unpack(wr.h, C.c_void_p(buf.data_ptr()), cnt.value, first_ghost, n_ghost), "unpack")
    gids = np.zeros(wr.num_bodies, np.int32)
    capi.check(L.b3b200_halo_ghost_ids(wr.h, capi.ptr(gids), len(gids)), "ghost ids")
    got = gids[first_ghost: first_ghost + cnt.value]
    assert sorted(got.tolist()) == sorted((1000 + want).tolist())
    assert np.all(gids[first_ghost + cnt.value: first_ghost + n_ghost] == -1)
    br, bl = wr.bodies(), wl.bodies()
    for slot, gid in enumerate(got):
        src = bl[1 + (gid - 1000)]
        dst = br[first_ghost + slot]
        for f in ("pos", "quat", "linVel", "angVel"):
            assert np.array_equal(np.asarray(src[f])[:3].view(np.uint32), np.asarray(dst[f])[:3].view(np.uint32)), f
        assert src["collidableIdx"] == dst["collidableIdx"] and src["invMass"] == dst["invMass"]
    parked = br[first_ghost + cnt.value:]
    assert np.all(parked["invMass"] == 0) and np.all(parked["pos"][:, 0] >= 1.0e6)
    # pairs of the right world = pairs of a world holding the right bodies + the mirrored left bodies
    wr.update_aabbs()
    wr.find_pairs()
    pr = wr.pairs()
    ids = np.full(wr.num_bodies, -5, np.int64)
    ids[0] = -1
    ids[1: 1 + len(idr)] = idr
    ids[first_ghost: first_ghost + cnt.value] = idl[got - 1000]
    have = set(tuple(sorted((int(ids[a]), int(ids[b])))) for a, b in zip(pr["x"], pr["y"]))
    assert all(-5 not in p for p in have)
    keep = np.concatenate([idr, idl[want]])
    wf = capi.World(capi.default_config(2048))
    scenes.add_ground_box(wf, 50.0)
    box = wf.register_convex_points(scenes.box_points(0.5))
    hull = wf.register_convex_points(scenes.random_hull_points(np.random.default_rng(5), 10, 0.5, 0.7))
    cols = np.where(np.arange(n) % 2 == 0, box, hull).astype(np.int32)
    wf.register_instances(np.ones(len(keep), np.float32), pos[keep], quat[keep], cols[keep])
    wf.upload()
    wf.update_aabbs()
    wf.find_pairs()
    pf = wf.pairs()
    idf = np.concatenate([[-1], keep]).astype(np.int64)
    ref = set(tuple(sorted((int(idf[a]), int(idf[b])))) for a, b in zip(pf["x"], pf["y"]))
    assert have == ref and len(ref) > 100


# ------------------------------------------------------------------ joints
def joint_world(seed=0, n=40):
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(1024))
    box = w.register_convex_points(scenes.box_points(0.5))
    w.register_instance(0.0, (0, 10, 0), scenes.IDENT, box)  # a static anchor high above
    for i in range(n):
        w.register_instance(1.0 + 0.05 * i, (rng.uniform(-8, 8), rng.uniform(6, 14), rng.uniform(-8, 8)), scenes.random_quat(rng), box)
    w.upload()
    b = w.bodies()
    b["linVel"][1:, :3] = rng.uniform(-1, 1, (n, 3))
    b["angVel"][1:, :3] = rng.uniform(-1, 1, (n, 3))
    w.write_bodies(b)
    return w, rng


@pytest.mark.parametrize("seed", [0, 1])
def test_joint_solver_matches_oracle(seed):
    """P2P and fixed joints in chains and stars (several batches): velocities after solve_joints against the oracle's
    restatement of b3GpuPgsConstraintSolver::solveJoints with the same batches"""
    w, rng = joint_world(seed)
    n = w.num_bodies - 1
    for i in range(1, n):  # a chain
        w.create_p2p_constraint(i, i + 1, rng.uniform(-0.5, 0.5, 3), rng.uniform(-0.5, 0.5, 3))
    for i in range(1, n, 5):  # some hang on the static anchor
        w.create_p2p_constraint(0, i, (0, 0, 0), (0, 0.5, 0))
    for i in range(2, n, 7):  # fixed joints across the chain
        w.create_fixed_constraint(i, (i + 3) % n + 1, (0.5, 0, 0), (-0.5, 0, 0), scenes.random_quat(rng))
    assert w.num_constraints > 40
    bodies, inertias, joints = w.bodies(), w.inertias(), w.joints()
    w.solve_joints()
    g = w.bodies()
    ob, oj = oa.solve_joints_oracle(bodies, inertias, joints)
    for f in ("linVel", "angVel"):
        assert rel_close(g[f][:, :3], ob[f][:, :3], 1e-4), f
    assert np.array_equal(w.joints()["flags"], oj["flags"])
    assert not np.allclose(g["linVel"], bodies["linVel"])


def test_joint_breaking_and_removal():
    w, rng = joint_world(3, n=6)
    weak = w.create_p2p_constraint(1, 2, (0.5, 0, 0), (-2.5, 0, 0), breaking_threshold=0.01)  # far apart: breaks at once
    strong = w.create_p2p_constraint(3, 4, (0.5, 0, 0), (-0.5, 0, 0))
    assert (weak, strong) == (0, 1) and w.num_constraints == 2
    w.solve_joints()
    j = w.joints()
    assert j["flags"][j["uid"] == weak][0] == 0 and j["flags"][j["uid"] == strong][0] == 1
    w.remove_constraint(weak)
    assert w.num_constraints == 1 and w.joints()["uid"][0] == strong
    third = w.create_p2p_constraint(5, 6, (0, 0, 0), (0, 0, 0))
    assert third == 2  # uids keep counting (m_constraintUid)
    w.solve_joints()
    assert np.all(w.joints()["flags"] == 1)


def test_pendulum_chain_holds_under_gravity():
    """a chain of boxes hanging from a static anchor through P2P joints, full steps: the pivots stay together"""
    w = capi.World(capi.default_config(256))
    box = w.register_convex_points(scenes.box_points(0.25))
    w.register_instance(0.0, (0, 20, 0), scenes.IDENT, box)
    n = 10
    for i in range(n):
        w.register_instance(1.0, (0.8 * (i + 1), 20, 0), scenes.IDENT, box)
    w.upload()
    for i in range(n):
        w.create_p2p_constraint(i, i + 1, (0.4, 0, 0), (-0.4, 0, 0))
    w.set_solver(capi.SOLVER_PGS, 4)
    for _ in range(240):
        w.step(1 / 60)
    b = w.bodies()
    assert np.isfinite(b["pos"]).all()
    gaps = []
    for i in range(n):
        pa = b["pos"][i, :3] + oa_rotate(b["quat"][i], np.array([0.4, 0, 0]))
        pb = b["pos"][i + 1, :3] + oa_rotate(b["quat"][i + 1], np.array([-0.4, 0, 0]))
        gaps.append(np.linalg.norm(pa - pb))
    assert max(gaps) < 0.25, gaps  # Baumgarte-stabilised joints: small drift only
    assert b["pos"][1:, 1].min() < 15  # and the chain swung down


def oa_rotate(q, v):
    x, y, z, w_ = [float(t) for t in q]
    u = np.array([x, y, z])
    return v + 2 * np.cross(u, np.cross(u, v) + w_ * v)


# ------------------------------------------------------------------ raycast (b3GpuRigidBodyPipeline::castRays)
def ray_bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_hits_equal(g, o):
    assert np.array_equal(g["hitBody"], o["hitBody"])
    assert np.array_equal(ray_bits(g["hitFraction"]), ray_bits(o["hitFraction"]))
    hit = o["hitBody"] >= 0
    assert np.array_equal(ray_bits(g["hitPoint"][hit, :3]), ray_bits(o["hitPoint"][hit, :3]))
    assert np.array_equal(ray_bits(g["hitNormal"][hit, :3]), ray_bits(o["hitNormal"][hit, :3]))


@pytest.mark.parametrize("seed,plane", [(0, True), (1, False)])
def test_cast_rays_match_oracle(seed, plane):
    """hulls + spheres are hit, planes / compounds are skipped like the reference's `default:` case"""
    w, sh, bodies = sphere_world(seed, n=700, plane=plane)
    rng = np.random.default_rng(50 + seed)
    n = 5000
    frm = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    to = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    frm[:, 1] = rng.uniform(0.2, 6, n)
    to[:, 1] = rng.uniform(0.2, 4, n)
    o = oa.cast_rays_oracle(frm, to, bodies, sh)
    assert (o["hitBody"] >= 0).sum() > n // 4 and (o["hitBody"] < 0).sum() > n // 20
    o2 = oa.cast_rays_oracle(frm, to, bodies, sh, max_fraction=0.3)
    o3 = oa.cast_rays_oracle(frm, to, bodies, sh, max_fraction=1.7)
    for accel in (0, 1):  # brute force over the world AABBs / linear BVH: same answers
        w.set_ray_accel(accel)
        assert_hits_equal(w.cast_rays(frm, to), o)
        # capped rays: hits beyond the cap leave the record untouched
        g2 = w.cast_rays(frm, to, max_fraction=0.3)
        assert_hits_equal(g2, o2)
        assert np.all(g2["hitFraction"][g2["hitBody"] < 0] == np.float32(0.3))
        # a cap beyond the end point extends the ray like in the reference (t is only compared with the cap)
        assert_hits_equal(w.cast_rays(frm, to, max_fraction=1.7), o3)


def test_cast_rays_after_stepping_and_edge_cases():
    w, sh, bodies = sphere_world(2, n=400, plane=True)
    assert len(w.cast_rays(np.zeros((0, 3)), np.zeros((0, 3)))) == 0
    for _ in range(20):
        w.step(1 / 60)
    b = w.bodies()
    rng = np.random.default_rng(9)
    # vertical picking rays from above: every ray over a hull / sphere body must report the topmost one
    frm = np.stack([rng.uniform(-5, 5, 3000), np.full(3000, 30.0), rng.uniform(-5, 5, 3000)], 1).astype(np.float32)
    to = frm.copy()
    to[:, 1] = -1.0
    o = oa.cast_rays_oracle(frm, to, b, sh)
    assert (o["hitBody"] >= 0).sum() > 500
    for accel in (0, 1):
        w.set_ray_accel(accel)
        assert_hits_equal(w.cast_rays(frm, to), o)
    # degenerate ray (from == to) hits nothing
    z = w.cast_rays(frm[:4], frm[:4])
    assert np.all(z["hitBody"] == -1)


def test_cast_rays_many_bodies_chunked():
    """more bodies than one candidate chunk (2048) and ties between coincident bodies -> lowest index wins"""
    rng = np.random.default_rng(3)
    w = capi.World(capi.default_config(8192))
    box = w.register_convex_points(scenes.box_points(0.4))
    sph = w.register_sphere(0.4)
    for i in range(5000):
        p = (rng.uniform(-20, 20), rng.uniform(0, 10), rng.uniform(-20, 20))
        w.register_instance(1.0, p, scenes.random_quat(rng), box if i % 3 else sph)
    # exact duplicates of the first 50 bodies, placed after them
    t0 = w.tables()["bodies"]
    for i in range(50):
        w.register_instance(1.0, tuple(t0["pos"][i, :3]), tuple(t0["quat"][i]), box if i % 3 else sph)
    w.upload()
    t = w.tables()
    sh, bodies = oa.Shapes(t), t["bodies"]
    n = 4000
    frm = rng.uniform(-22, 22, (n, 3)).astype(np.float32)
    to = rng.uniform(-22, 22, (n, 3)).astype(np.float32)
    # aim a block of rays straight at the duplicated bodies
    to[:50] = bodies["pos"][:50, :3]
    frm[:50] = bodies["pos"][:50, :3] + np.float32([0, 15, 0])
    o = oa.cast_rays_oracle(frm, to, bodies, sh)
    assert (o["hitBody"][:50] >= 0).all() and (o["hitBody"][:50] < 5000).all()
    for accel in (-1, 0, 1):
        w.set_ray_accel(accel)
        assert_hits_equal(w.cast_rays(frm, to), o)


# ------------------------------------------------------------------ checkpoint / render interop
def jointed_world(n):
    w, _ = joint_world(0, n=n)
    for i in range(1, n, 3):
        w.create_p2p_constraint(i, i + 1 if i + 1 <= n else 0, (0.4, 0, 0), (-0.4, 0, 0), 50.0 if i % 2 else 1e30)
    w.create_fixed_constraint(0, 1, (0, -1, 0), (0, 1, 0), (0, 0, 0, 1))
    return w


def test_checkpoint_roundtrip_restores_bodies_inertias_and_joints(tmp_path):
    w = jointed_world(40)
    w.set_solver(capi.SOLVER_PGS, 4)
    for _ in range(30):
        w.step(1 / 60)
    path = tmp_path / "state.b3cp"
    w.checkpoint_save(path)
    saved_b, saved_j = w.bodies().copy(), w.joints().copy()
    saved_i = w.inertias().copy()
    for _ in range(20):
        w.step(1 / 60)
    assert not np.array_equal(w.bodies()["pos"], saved_b["pos"])
    # restore into the same world and into a freshly built twin
    w2 = jointed_world(40)
    w2.remove_constraint(0)  # the file carries the joint set, whatever the target had
    for target in (w, w2):
        target.checkpoint_load(path)
        assert np.array_equal(target.bodies().view(np.uint8), saved_b.view(np.uint8))
        assert np.array_equal(target.inertias().view(np.uint8), saved_i.view(np.uint8))
        assert np.array_equal(target.joints().view(np.uint8), saved_j.view(np.uint8))
    # both continue from the same state: one more step of the (joint + contact) pipeline gives the same bodies
    w.set_solver(capi.SOLVER_PGS, 4)
    w2.set_solver(capi.SOLVER_PGS, 4)
    w.step(1 / 60)
    w2.step(1 / 60)
    a, b = w.bodies(), w2.bodies()
    assert rel_close(a["pos"], b["pos"], 1e-5) and rel_close(a["linVel"], b["linVel"], 1e-4)
    # a world of another size refuses the file
    w3 = jointed_world(12)
    with pytest.raises(capi.B3Error):
        w3.checkpoint_load(path)
    with pytest.raises(capi.B3Error):
        w.checkpoint_load(tmp_path / "missing.b3cp")


def test_copy_transforms_matches_body_buffer():
    import torch

    w, sh, bodies, inert = gpu_world(6, 1)
    for _ in range(5):
        w.step(1 / 60)
    n = len(bodies)
    out = torch.full((2 * n, 4), -7.0, dtype=torch.float32, device="cuda:0")
    w.copy_transforms(out.data_ptr(), n)
    w.synchronize()
    o = out.cpu().numpy()
    b = w.bodies()
    assert np.array_equal(o[:n, :3].view(np.uint32), b["pos"][:, :3].view(np.uint32))
    assert np.all(o[:n, 3] == 1.0)
    assert np.array_equal(o[n:].view(np.uint32), b["quat"].view(np.uint32))
    # a prefix only: the rest of the buffer is untouched
    out.fill_(-7.0)
    w.copy_transforms(out.data_ptr(), 10)
    w.synchronize()
    o = out.cpu().numpy()
    assert np.array_equal(o[10:20].view(np.uint32), b["quat"][:10].view(np.uint32)) and np.all(o[20:] == -7.0)


def test_standalone_solver_entry_on_device_buffers_matches_world_solve():
    """b3GpuPgsContactSolver / b3GpuJacobiContactSolver::solveContacts as a stand-alone entry: a scratch world solves another
    world's device buffers in place and gives the velocities that world's own solve gives"""
    for kind, iters in ((capi.SOLVER_PGS, 6), (capi.SOLVER_JACOBI, 7)):
        w, sh, bodies, inertias = gpu_world(n_side=8, seed=11)
        w.set_solver(kind, iters)
        w.set_colouring(0)  # two worlds solve the same contacts: the reproducible batch assignment
        w.update_aabbs()
        w.find_pairs()
        w.compute_contacts()
        start = w.bodies()
        ncontacts = len(w.contacts())
        assert ncontacts > 300
        w.solve_contacts()
        own = w.bodies()
        assert not np.array_equal(own["linVel"], start["linVel"])
        w.write_bodies(start)  # back to the pre-solve state; the contact buffer is untouched
        scratch = capi.World(capi.default_config(2048))
        sphere = scratch.register_sphere(0.5)
        for i in range(len(start) + 5):
            scratch.register_instance(1.0, (4.0 * i, 0, 0), scenes.IDENT, sphere)
        scratch.upload()
        scratch.set_solver(kind, iters)
        scratch.set_colouring(0)
        scratch.solve_contacts_device(len(start), w.device_buffer(0), w.device_buffer(4), ncontacts, w.device_buffer(3), 0)
        got = w.bodies()
        for f in ("linVel", "angVel"):
            if kind == capi.SOLVER_PGS:
                assert np.array_equal(got[f].view(np.uint32), own[f].view(np.uint32)), f
            else:  # a body's split slots are handed out with an atomic counter (like CountBodiesKernel): equal up to summation order
                assert rel_close(got[f], own[f], 1e-4), f
        assert np.array_equal(got["pos"].view(np.uint32), start["pos"].view(np.uint32))
        # host pointers work too (cudaMemcpyDefault)
        hb, hi, hc = start.copy(), w.inertias(), w.contacts()
        scratch.solve_contacts_device(len(hb), hb.ctypes.data, hi.ctypes.data, len(hc), hc.ctypes.data, 0)
        assert rel_close(hb["linVel"], own["linVel"], 1e-4 if kind == capi.SOLVER_JACOBI else 0.0)


def test_halo_emigrate_adopt_between_two_worlds():
    """migration on one GPU: the bodies of world A whose centre lies beyond x = 0 are handed to free slots of world B;
    B then holds their exact state under the same global ids, A's slots are parked and id-less, nothing else moves"""
    import ctypes as C
    import torch

    rng = np.random.default_rng(1)
    n, spare = 200, 160

    def make(num_real):
        w = capi.World(capi.default_config(1024))
        scenes.add_ground_box(w, 50.0)
        box = w.register_convex_points(scenes.box_points(0.5))
        tet = w.register_convex_points(scenes.tetra_points(0.6))
        for i in range(num_real):
            p = (rng.uniform(-6, 6), rng.uniform(0.6, 3.0), rng.uniform(-3, 3))
            w.register_instance(1.0 + 0.01 * i, p, scenes.random_quat(rng), box if i % 2 else tet)
        for k in range(spare):
            w.register_instance(1.0, (1.0e6 + 1024.0 * (1 + num_real + k), -1.0e6, 1.0e6), scenes.IDENT, box)
        w.upload()
        b = w.bodies()
        b["invMass"][1 + num_real:] = 0.0
        b["linVel"][1: 1 + num_real, :3] = rng.normal(size=(num_real, 3))
        b["angVel"][1: 1 + num_real, :3] = rng.normal(size=(num_real, 3))
        w.write_bodies(b)
        return w

    wa, wb = make(n), make(40)
    L = capi.lib()
    ids_a = np.full(wa.num_bodies, -1, np.int32)
    ids_a[1: 1 + n] = 5000 + np.arange(n)
    ids_b = np.full(wb.num_bodies, -1, np.int32)
    ids_b[1: 41] = 9000 + np.arange(40)
    capi.check(L.b3b200_halo_set_ids(wa.h, capi.ptr(ids_a), len(ids_a)), "set_ids")
    capi.check(L.b3b200_halo_set_ids(wb.h, capi.ptr(ids_b), len(ids_b)), "set_ids")
    before_a, before_b, inert_a = wa.bodies(), wb.bodies(), wa.inertias()
    want = np.nonzero(before_a["pos"][1: 1 + n, 0] >= 0.0)[0] + 1
    rec = L.b3b200_halo_record_size()
    buf = torch.zeros(spare * rec, dtype=torch.uint8, device="cuda")
    slots = np.zeros(spare, np.int32)
    cnt = C.c_int(0)
    capi.check(L.b3b200_halo_emigrate(wa.h, 0, C.c_float(0.0), C.c_float(3.0e38), wa.num_bodies, 0, C.c_void_p(buf.data_ptr()), spare, capi.ptr(slots), C.byref(cnt)), "emigrate")
    assert cnt.value == len(want) and 40 < cnt.value < spare
    assert sorted(slots[: cnt.value].tolist()) == want.tolist()
    after_a = wa.bodies()
    gone = np.zeros(wa.num_bodies, bool)
    gone[want] = True
    assert np.all(after_a["invMass"][gone] == 0) and np.all(after_a["pos"][gone, 0] >= 1.0e6)
    assert np.array_equal(after_a[~gone].view(np.uint8), before_a[~gone].view(np.uint8))
    ga = np.zeros(wa.num_bodies, np.int32)
    capi.check(L.b3b200_halo_ghost_ids(wa.h, capi.ptr(ga), len(ga)), "ids")
    assert np.all(ga[gone] == -1) and np.array_equal(ga[~gone], ids_a[~gone])
    # a second call finds nothing left to move
    capi.check(L.b3b200_halo_emigrate(wa.h, 0, C.c_float(0.0), C.c_float(3.0e38), wa.num_bodies, 0, C.c_void_p(buf.data_ptr() + 0), spare, capi.ptr(np.zeros(spare, np.int32)), C.byref(C.c_int(0))), "emigrate")
    # adopt into B's free slots (any order the caller likes)
    free = np.arange(41, 41 + cnt.value, dtype=np.int32)[::-1].copy()
    capi.check(L.b3b200_halo_adopt(wb.h, C.c_void_p(buf.data_ptr()), cnt.value, capi.ptr(free)), "adopt")
    after_b, inert_b = wb.bodies(), wb.inertias()
    gb = np.zeros(wb.num_bodies, np.int32)
    capi.check(L.b3b200_halo_ghost_ids(wb.h, capi.ptr(gb), len(gb)), "ids")
    for k in range(cnt.value):
        src, dst = int(slots[k]), int(free[k])
        assert gb[dst] == ids_a[src]
        for f in ("pos", "quat", "linVel", "angVel"):
            assert np.array_equal(np.asarray(before_a[src][f])[:3].view(np.uint32), np.asarray(after_b[dst][f])[:3].view(np.uint32)), f
        assert before_a[src]["invMass"] == after_b[dst]["invMass"] and before_a[src]["collidableIdx"] == after_b[dst]["collidableIdx"]
        assert np.array_equal(inert_a[src: src + 1].view(np.uint8), inert_b[dst: dst + 1].view(np.uint8))
    untouched = np.ones(wb.num_bodies, bool)
    untouched[free] = False
    assert np.array_equal(after_b[untouched].view(np.uint8), before_b[untouched].view(np.uint8))
    # the adopted bodies take part in B's step like any other body
    wb.step(1 / 60)
    assert np.isfinite(wb.bodies()["pos"]).all() and wb.counters()[0] > 0
    # capacity overflow is reported, the bodies that did not fit stay
    wc = make(n)
    capi.check(L.b3b200_halo_set_ids(wc.h, capi.ptr(ids_a), len(ids_a)), "set_ids")
    small = 8
    rc = L.b3b200_halo_emigrate(wc.h, 0, C.c_float(-3.0e38), C.c_float(3.0e38), wc.num_bodies, 0, C.c_void_p(buf.data_ptr()), small, capi.ptr(slots), C.byref(cnt))
    assert rc < 0 and cnt.value == small
    assert int((wc.bodies()["invMass"][1: 1 + n] != 0).sum()) == n - small


@pytest.mark.timeout(120)
@pytest.mark.parametrize("colouring", [0, 1])
@pytest.mark.parametrize("side,overflow", [(10, False), (12, True)])
def test_high_degree_body_colouring(side, overflow, colouring):
    """one dynamic plate carrying side^2 boxes: its contacts all need different batches.  Up to B3_MAX_NUM_BATCHES = 128
    colours that works; beyond, the extra contacts are left out of the solve and the overflow flag is raised (the reference
    errors out, b3GpuPgsContactSolver.cpp:1497-1502) -- in both batch-assignment modes, without hanging"""
    w = capi.World(capi.default_config(1024))
    scenes.add_ground_box(w, 50.0)
    plate = w.register_convex_points(scenes.box_points(side * 0.6, 0.25, side * 0.6))
    box = w.register_convex_points(scenes.box_points(0.5))
    w.register_instance(50.0, (0, 0.25, 0), scenes.IDENT, plate)
    for i in range(side):
        for k in range(side):
            w.register_instance(1.0, ((i - (side - 1) / 2) * 1.1, 0.5 + 0.5 - 0.005, (k - (side - 1) / 2) * 1.1), scenes.IDENT, box)
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 4)
    w.set_colouring(colouring)
    w.update_aabbs()
    w.find_pairs()
    w.compute_contacts()
    contacts = w.contacts()
    on_plate = int(((np.abs(contacts["bodyA"]) == 1) | (np.abs(contacts["bodyB"]) == 1)).sum())
    assert on_plate == side * side + 1  # every box + the ground
    w.solver_setup()
    flags = int(w.counters()[4])
    cols = w.contacts()["batchIdx"]
    if overflow:
        assert flags & 4 and (cols < 0).sum() == on_plate - 128 and cols.max() == 127
    else:
        assert not (flags & 4) and cols.min() >= 0 and cols.max() + 1 >= on_plate
    # valid batches: no dynamic body twice in a batch
    b = w.bodies()
    for c in np.unique(cols[cols >= 0]):
        seg = contacts[cols == c]
        ids = np.concatenate([np.abs(seg["bodyA"]), np.abs(seg["bodyB"])])
        ids = ids[b["invMass"][ids] != 0]
        assert len(ids) == len(np.unique(ids))
    w.solver_iterate()
    w.integrate(1 / 60)
    assert np.isfinite(w.bodies()["pos"]).all()
    for _ in range(5):
        w.step(1 / 60)
    assert np.isfinite(w.bodies()["pos"]).all()


# ------------------------------------------------------------------ body edits and uploads after stepping
def test_upload_after_stepping_keeps_the_device_state():
    """registering one more body mid-simulation and uploading must not rewind the bodies that are already on the device
    (the reference copies only the new body: copyFromHostPointer(&body, 1, bodyIndex), b3GpuNarrowPhase.cpp:903-907)"""
    w = capi.World(capi.default_config(1024))
    scenes.add_ground_box(w, 50.0)
    box = w.register_convex_points(scenes.box_points(0.5))
    for i in range(50):
        w.register_instance(1.0, (1.5 * (i % 10), 0.5 + 1.2 * (i // 10), 0.0), scenes.IDENT, box)
    w.upload()
    w.set_solver(capi.SOLVER_PGS, 4)
    for _ in range(20):
        w.step(1 / 60)
    before = w.bodies()
    assert np.abs(before["linVel"][1:, 1]).max() > 0.01 or np.abs(before["pos"][1:, 1] - 0.5).max() > 1e-3  # it did move
    new = w.register_instance(1.0, (0.0, 30.0, 5.0), scenes.IDENT, box)
    w.upload()
    after = w.bodies()
    assert len(after) == len(before) + 1 and new == len(before)
    for f in ("pos", "quat", "linVel", "angVel"):
        assert np.array_equal(after[f][:-1].view(np.uint32), before[f].view(np.uint32)), f
    assert tuple(after["pos"][-1][:3]) == (0.0, 30.0, 5.0)
    w.step(1 / 60)
    assert np.isfinite(w.bodies()["pos"]).all()


def test_write_body_and_read_body_touch_one_body():
    w, sh, bodies, inertias = gpu_world(n_side=4, seed=3)
    L = capi.lib()
    w.step(1 / 60)
    before = w.bodies()
    one = np.zeros(1, capi.rigid_body_t)
    capi.check(L.b3b200_read_body(w.h, 7, capi.ptr(one)), "read_body")
    assert one.tobytes() == before[7:8].tobytes()
    one["pos"][0][:3] = (1.0, 9.0, -2.0)
    one["linVel"][0][:3] = (0.0, 0.0, 3.0)
    capi.check(L.b3b200_write_body(w.h, 7, capi.ptr(one)), "write_body")
    after = w.bodies()
    assert after[7:8].tobytes() == one.tobytes()
    keep = np.arange(len(after)) != 7
    assert after[keep].tobytes() == before[keep].tobytes()
    assert L.b3b200_write_body(w.h, len(after), capi.ptr(one)) != 0 and L.b3b200_read_body(w.h, -1, capi.ptr(one)) != 0


# ------------------------------------------------------------------ whole-step CUDA graphs
def test_step_graph_replays_match_kernel_by_kernel_steps():
    """b3b200_step / step_n through captured graphs (one per (AABBs valid, partition due) key) against the same steps launched
    kernel by kernel.  The scene is order independent (every box touches only the ground, so the order in which the narrowphase's
    atomics append the contacts cannot change a bit), which lets the two worlds be compared bit for bit over 60 steps that
    include re-partitions, a settings call that drops the graphs and a body upload that invalidates the AABBs."""
    def build(graphs):
        w = capi.World(capi.default_config(4096))
        scenes.add_ground_box(w, 80.0)
        col = w.register_convex_points(scenes.box_points(0.5))
        rng = np.random.default_rng(5)
        for i in range(20):
            for k in range(20):
                w.register_instance(1.0, (i * 3.0 - 30, 0.7 + 0.3 * rng.uniform(), k * 3.0 - 30), scenes.random_quat(rng), col)
        w.upload()
        w.set_solver(capi.SOLVER_PGS, 6)
        w.set_step_graphs(graphs)
        return w

    a, b = build(False), build(True)
    l0 = capi.lib().b3b200_launch_count()
    a.step_n(1 / 60, 25)
    la = capi.lib().b3b200_launch_count() - l0
    b.step_n(1 / 60, 25)
    lb = capi.lib().b3b200_launch_count() - l0 - la
    assert la == lb and la > 25 * 20, (la, lb)  # replays count the kernels they contain
    for w in (a, b):
        w.set_gravity((0.0, -9.8, 0.5))  # any settings call drops the graphs
        w.step_n(1 / 60, 10)
        st = w.bodies()
        st["linVel"][1:, 1] += 1.0
        w.write_bodies(st)  # AABBs invalid -> the next step is a different graph
        for _ in range(25):
            w.step(1 / 60)
    ba, bb = a.bodies(), b.bodies()
    assert a.counters()[1] == b.counters()[1] > 300
    for f in ("pos", "quat", "linVel", "angVel"):
        assert np.array_equal(ba[f].view(np.uint32), bb[f].view(np.uint32)), f
    a.close()
    b.close()


# ------------------------------------------------------------------ batched independent worlds
def _pile_world(w, rng, nx, ny, nz, cols, ground_col):
    w.register_instance(0.0, (0.0, -50.0, 0.0), scenes.IDENT, ground_col)
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                p = np.array([i, j, k], np.float64) * 1.5 + rng.uniform(-0.15, 0.15, 3)
                p[1] += 0.9
                w.register_instance(1.0, tuple(p), scenes.random_quat(rng), cols[int(rng.integers(0, len(cols)))])


@pytest.mark.parametrize("kind", [capi.BP_GRID, capi.BP_SAP])
def test_batched_worlds_see_exactly_their_own_pairs_and_contacts(kind):
    """b3b200_set_current_world: 7 different piles registered at the SAME coordinates as 7 worlds of one b3b200 world, against
    each pile alone in a world of its own: per world the sorted pair set and the contacts (points, normals, depths) are
    bit-identical, no pair joins two worlds, and after stepping the batch every world still rests on its own ground."""
    n_worlds = 7
    def shapes(w):
        cols = [w.register_convex_points(scenes.box_points(0.5)), w.register_convex_points(scenes.tetra_points(0.9))]
        r = np.random.default_rng(99)
        cols.append(w.register_convex_points(scenes.random_hull_points(r, 12, 0.5, 0.8)))
        ground = w.register_convex_points(scenes.box_points(50.0))
        return cols, ground

    batch = capi.World(capi.default_config(8192))
    cols, ground = shapes(batch)
    for k in range(n_worlds):
        batch.set_current_world(k)
        _pile_world(batch, np.random.default_rng(100 + k), 4, 3 + k % 3, 4, cols, ground)
    batch.upload()
    assert batch.num_worlds() == n_worlds
    batch.set_broadphase(kind)
    world_of = batch.body_worlds()
    first = [int(np.nonzero(world_of == k)[0][0]) for k in range(n_worlds)]
    batch.update_aabbs()
    batch.find_pairs()
    bp = batch.pairs()
    batch.compute_contacts()
    bc = batch.contacts()
    assert np.array_equal(world_of[bp["x"]], world_of[bp["y"]]), "a pair between two worlds"
    for k in range(n_worlds):
        single = capi.World(capi.default_config(1024))
        cs, gs = shapes(single)
        _pile_world(single, np.random.default_rng(100 + k), 4, 3 + k % 3, 4, cs, gs)
        single.upload()
        single.set_broadphase(kind)
        single.update_aabbs()
        single.find_pairs()
        sp = single.pairs()
        single.compute_contacts()
        sc = single.contacts()
        mine = bp[world_of[bp["x"]] == k].copy()
        mine["x"] -= first[k]
        mine["y"] -= first[k]
        got, want = oa.sorted_pair_set(mine), oa.sorted_pair_set(sp)
        assert len(want) > 50 and np.array_equal(got, want), k
        cm = bc[world_of[np.abs(bc["bodyA"])] == k].copy()
        cm["bodyA"] = np.sign(cm["bodyA"]) * (np.abs(cm["bodyA"]) - first[k])
        cm["bodyB"] = np.sign(cm["bodyB"]) * (np.abs(cm["bodyB"]) - first[k])
        # (body 0 of a stand-alone world is its static ground: -0 == 0, compare magnitudes + the static flag through invMass)
        a, b = contact_table(cm), contact_table(sc)
        assert len(a) == len(b) > 10
        assert np.array_equal(np.abs(a["bodyA"]), np.abs(b["bodyA"])) and np.array_equal(np.abs(a["bodyB"]), np.abs(b["bodyB"]))
        assert np.array_equal(a["worldNormalOnB"].view(np.uint32), b["worldNormalOnB"].view(np.uint32))
        assert np.array_equal(a["worldPosB"].view(np.uint32), b["worldPosB"].view(np.uint32))
        single.close()
    batch.set_solver(capi.SOLVER_PGS, 6)
    batch.step_n(1 / 60, 150)
    b = batch.bodies()
    dyn = b["invMass"] != 0
    assert np.isfinite(b["pos"]).all() and (b["pos"][dyn, 1] > 0.1).all() and (b["pos"][dyn, 1] < 12).all()
    assert np.median(np.linalg.norm(b["linVel"][dyn, :3], axis=1)) < 0.5
    batch.close()


def test_batched_identical_worlds_have_no_cross_block_contacts_and_agree_with_a_single_world():
    """config 5(i) in small: 40 copies of one 8 x 4 x 8 box pile as 40 worlds.  Blocks hold whole worlds (equal dynamic body
    counts), so the solver has no cross-block batch; after 120 steps every world's kinetic / potential energy agrees with the
    same pile stepped alone (the solve ORDER differs between the two, so the comparison is statistical)."""
    def pile(w):
        col = w.register_convex_points(scenes.box_points(0.5))
        ground = w.register_convex_points(scenes.box_points(50.0))
        return col, ground

    def add(w, col, ground):
        w.register_instance(0.0, (0.0, -50.0, 0.0), scenes.IDENT, ground)
        for i in range(8):
            for j in range(4):
                for k in range(8):
                    w.register_instance(1.0, (((j + 1) & 1) * 0.3 + 1.1 * i, 0.6 + 1.05 * j, ((j + 1) & 1) * 0.3 + 1.1 * k), scenes.IDENT, col)

    n_worlds = 40
    batch = capi.World(capi.default_config(16384))
    col, ground = pile(batch)
    for k in range(n_worlds):
        batch.set_current_world(k)
        add(batch, col, ground)
    batch.upload()
    batch.set_solver(capi.SOLVER_PGS, 10)
    single = capi.World(capi.default_config(1024))
    c1, g1 = pile(single)
    add(single, c1, g1)
    single.upload()
    single.set_solver(capi.SOLVER_PGS, 10)
    batch.step_n(1 / 60, 120)
    single.step_n(1 / 60, 120)
    b, s = batch.bodies(), single.bodies()
    world_of = batch.body_worlds()
    assert batch.counters()[1] > 30 * n_worlds * 10

    def energy(x):
        d = x["invMass"] != 0
        return 0.5 * (x["linVel"][d, :3] ** 2).sum(), 9.8 * x["pos"][d, 1].sum()

    ks, ps = energy(s)
    for k in range(n_worlds):
        kb, pb = energy(b[world_of == k])
        assert abs(pb - ps) < 0.02 * ps, (k, pb, ps)
        assert kb < max(4 * ks, 2.0), (k, kb, ks)
    # the batch keeps the reference's invariant and has no cross-block colour
    batch.update_aabbs()
    batch.find_pairs()
    batch.compute_contacts()
    batch.solver_setup()
    cs = batch.constraints()
    assert len(cs) > 0
    assert batch.counters()[3] == 0, "cross-block colours in a batch of equal worlds"
    batch.close()
    single.close()


def test_pipelined_host_stepping_matches_blocking_write_step_readback():
    """b3b200_step_host_async (upload / step / download overlapped across calls) against b3b200_write_bodies -> b3b200_step ->
    b3b200_readback_bodies on the same inputs, bit for bit (order-independent scene: every box touches only the ground)"""
    import torch

    def build():
        w = capi.World(capi.default_config(4096))
        scenes.add_ground_box(w, 80.0)
        col = w.register_convex_points(scenes.box_points(0.5))
        rng = np.random.default_rng(11)
        for i in range(24):
            for k in range(24):
                w.register_instance(1.0, (i * 3.0 - 36, 0.6 + 0.2 * rng.uniform(), k * 3.0 - 36), scenes.random_quat(rng), col)
        w.upload()
        w.set_solver(capi.SOLVER_PGS, 5)
        return w

    a, b = build(), build()
    a.step_n(1 / 60, 30)
    base = a.bodies()
    n = len(base)
    rng = np.random.default_rng(3)
    inputs = []
    for k in range(7):
        s = base.copy()
        s["linVel"][1:, :3] += rng.normal(size=(n - 1, 3)).astype(np.float32) * 0.3
        s["pos"][1:, 1] += 0.01 * k
        inputs.append(s)
    want = []
    for s in inputs:
        a.write_bodies(s)
        a.step(1 / 60)
        want.append(a.bodies())
    pins = [torch.empty(base.nbytes, dtype=torch.uint8).pin_memory() for _ in range(2 * len(inputs))]
    hin = [np.frombuffer(p.numpy(), dtype=capi.rigid_body_t) for p in pins[: len(inputs)]]
    hout = [np.frombuffer(p.numpy(), dtype=capi.rigid_body_t) for p in pins[len(inputs):]]
    for h, s in zip(hin, inputs):
        h[:] = s
    for h, o in zip(hin, hout):
        b.step_host_async(1 / 60, h, o)
    b.step_host_wait()
    for k in range(len(inputs)):
        for f in ("pos", "quat", "linVel", "angVel", "invMass", "collidableIdx"):
            assert np.array_equal(hout[k][f].view(np.uint32), want[k][f].view(np.uint32)), (k, f)
    # chained: no upload, the device state keeps stepping, every state comes back
    b.write_bodies(inputs[0])
    a.write_bodies(inputs[0])
    for k in range(3):
        b.step_host_async(1 / 60, None, hout[k])
        a.step(1 / 60)
        want[k] = a.bodies()
    b.step_host_wait()
    for k in range(3):
        assert np.array_equal(hout[k]["pos"].view(np.uint32), want[k]["pos"].view(np.uint32)), k
    assert np.array_equal(b.bodies()["linVel"].view(np.uint32), a.bodies()["linVel"].view(np.uint32))
    a.close()
    b.close()


def test_two_worlds_stepped_from_two_threads():
    """SURVEY 8(b) threading: worlds are independent objects; two of them stepped concurrently from two host threads (ctypes drops
    the GIL; each world has its own stream, graphs and error slot) end bit-identical to a world stepped alone"""
    import threading

    def build():
        w = capi.World(capi.default_config(4096))
        scenes.add_ground_box(w, 80.0)
        col = w.register_convex_points(scenes.box_points(0.5))
        rng = np.random.default_rng(21)
        for i in range(16):
            for k in range(16):
                w.register_instance(1.0, (i * 3.0 - 24, 0.6 + 0.2 * rng.uniform(), k * 3.0 - 24), scenes.random_quat(rng), col)
        w.upload()
        w.set_solver(capi.SOLVER_PGS, 5)
        return w

    alone = build()
    alone.step_n(1 / 60, 80)
    want = alone.bodies()
    ws = [build(), build()]
    errs = []

    def run(w):
        try:
            for _ in range(40):
                w.step_n(1 / 60, 2)
            w.synchronize()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=run, args=(w,)) for w in ws]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs
    for w in ws:
        b = w.bodies()
        for f in ("pos", "quat", "linVel", "angVel"):
            assert np.array_equal(b[f].view(np.uint32), want[f].view(np.uint32)), f
        w.close()
    alone.close()


def test_batched_worlds_with_gaps_and_unequal_sizes():
    """world ids need not be dense or equally filled: worlds 0, 3 and 9 (1-8 otherwise empty) with 40 / 1 / 150 dynamic bodies, one of
    them without any static body -- pairs stay inside their worlds, the lone body of world 3 falls freely through the others"""
    w = capi.World(capi.default_config(1024))
    col = w.register_convex_points(scenes.box_points(0.5))
    ground = w.register_convex_points(scenes.box_points(30.0))
    rng = np.random.default_rng(2)

    def pile(n, with_ground):
        if with_ground:
            w.register_instance(0.0, (0.0, -30.0, 0.0), scenes.IDENT, ground)
        for _ in range(n):
            w.register_instance(1.0, tuple(rng.uniform((-3, 0.6, -3), (3, 6, 3))), scenes.random_quat(rng), col)

    w.set_current_world(0)
    pile(40, True)
    w.set_current_world(3)
    pile(1, False)
    w.set_current_world(9)
    pile(150, True)
    w.upload()
    assert w.num_worlds() == 10
    wid = w.body_worlds()
    w.set_solver(capi.SOLVER_PGS, 6)
    lone = int(np.nonzero(wid == 3)[0][0])
    y0 = w.bodies()["pos"][lone, 1]
    for _ in range(30):
        w.step(1 / 60)
        p = w.pairs()
        assert np.array_equal(wid[p["x"]], wid[p["y"]])
    b = w.bodies()
    assert w.counters()[1] > 20 and w.counters()[4] == 0
    fall = 0.5 * 9.8 * (30 / 60) ** 2
    assert abs((y0 - b["pos"][lone, 1]) - fall) < 0.15 * fall  # free fall: nothing of the other worlds touched it
    dyn = (b["invMass"] != 0) & (wid != 3)
    assert (b["pos"][dyn, 1] > 0.2).all()
    w.close()


def test_grid_broadphase_keeps_a_few_long_bodies_out_of_the_cells():
    """a few long dynamic AABBs among 30 000 small ones: the grid cell is sized for the small ones, the long ones are listed and tested
    against everything (the cell used to be the widest AABB: every cell then held hundreds of bodies).  Exact pair set, and the
    time of calculateOverlappingPairs stays near that of the scene without the long bodies"""
    rng = np.random.default_rng(4)
    n = 30000
    a = np.zeros(n, capi.aabb_t)
    c = rng.uniform(-60, 60, (n, 3)).astype(np.float32)
    h = rng.uniform(0.3, 0.6, (n, 3)).astype(np.float32)
    a["min"][:, :3], a["max"][:, :3] = c - h, c + h
    a["minIndex"] = np.arange(n)
    small, large = np.arange(n, dtype=np.int32), np.zeros(0, np.int32)
    bp = run_standalone_bp(capi.BP_GRID, a, small, large, 1 << 20)
    base_ms = bp.last_ms()
    base_pairs = bp.num_overlap()
    for k, (lo, hi) in enumerate([((-55, -1, -1), (55, 1, 1)), ((-1, -58, -1), (1, 58, 1)), ((-50, -50, 3), (50, 50, 4))]):
        a["min"][7 + k, :3], a["max"][7 + k, :3] = lo, hi
    _, o = oa.brute_force_pairs(oa.oracle(), "orc_", a, small, large, 1 << 20)
    bp = run_standalone_bp(capi.BP_GRID, a, small, large, 1 << 20)
    assert bp.num_overlap() > base_pairs + 200
    assert np.array_equal(oa.sorted_pair_set(bp.pairs()), oa.sorted_pair_set(o))
    ms = []
    for _ in range(5):
        bp.calculate_pairs(1 << 20)
        ms.append(bp.last_ms())
    assert min(ms) < 5 * max(base_ms, 0.05), (ms, base_ms)
