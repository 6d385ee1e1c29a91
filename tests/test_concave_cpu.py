"""Concave trimesh contacts: the oracle (orc_concave_contacts) against the reference's own host twins
(b3BvhTraversal, b3FindConcaveSeparatingAxisKernel, clipFacesAndFindContactsKernel, b3NewContactReductionKernel run
by GpuSatCollision::computeConvexConvexContactsGPUSAT with the file-scope GPU switches off,
b3ConvexHullContact.cpp:20-24, 3513-3569, 3700-3770, 3850-3885, 3951-4003).  Runs without a GPU."""
import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi, scenes

pytestmark = pytest.mark.skipif(not oa.refcl_available(), reason="oracle/_ref/libb3refcl.so not built")


def build_both(seed=0, n=80, with_compounds=True, amplitude=1.5, freq=0.6):
    """a heightfield trimesh as body 0 plus a cloud of hulls / compounds around its surface, in a host-only B200
    world and in the reference narrowphase"""
    rng = np.random.default_rng(seed)
    cfg = capi.default_config(1024)
    cfg["maxTriConvexPairCapacity"] = 1 << 16
    w = capi.World(cfg, device=-1)
    r = oa.RefNarrowphase(cfg)
    verts, tris = scenes.heightfield_mesh(12, 12, cell=1.0, amplitude=amplitude, freq=freq)
    mesh = (w.register_concave(verts, tris), r.register_concave(verts, tris))
    assert mesh[0] == mesh[1]
    cols = {}

    def reg_convex(pts):
        rc = r.register_convex_points(pts)
        cv = r.table(2, capi.convex_t)[-1]
        vv = r.table(3, np.dtype(("f4", 4)))[cv["vertexOffset"]: cv["vertexOffset"] + cv["numVertices"]]
        faces = r.table(5, capi.face_t)[cv["faceOffset"]: cv["faceOffset"] + cv["numFaces"]].copy()
        idx_all = r.table(6, np.dtype("i4"))
        lo = int(faces["indexOffset"].min())
        hi = int((faces["indexOffset"] + faces["numIndices"]).max())
        faces["indexOffset"] -= lo
        edges = r.table(4, np.dtype(("f4", 4)))[cv["uniqueEdgesOffset"]: cv["uniqueEdgesOffset"] + cv["numUniqueEdges"]]
        poly = np.zeros(1, capi.convex_t)
        poly[0] = cv
        return w.register_convex(vv, faces, idx_all[lo:hi], edges, poly), rc

    cols["box"] = reg_convex(scenes.box_points(0.5))
    cols["hull"] = reg_convex(scenes.random_hull_points(rng, 12, 0.5, 0.8))
    cols["tetra"] = reg_convex(scenes.tetra_points(0.6))
    if with_compounds:
        cols["L"] = (w.register_compound(scenes.compound_children(cols["box"][0], scenes.L_OFFSETS)),
                     r.register_compound(scenes.compound_children(cols["box"][1], scenes.L_OFFSETS)))
    for k, (a, b) in cols.items():
        assert a == b, k
    kinds = list(cols)
    bodies = [(0.0, (0.0, 0.0, 0.0), scenes.IDENT, mesh)]
    for i in range(n):
        x, z = rng.uniform(-5.0, 5.0, 2)
        h = amplitude * np.sin(freq * x) * np.cos(freq * z)
        p = (x, h + rng.uniform(-0.2, 0.9), z)
        bodies.append((1.0, p, scenes.random_quat(rng), cols[kinds[int(rng.integers(0, len(kinds)))]]))
    for mass, p, q, col in bodies:
        w.register_instance(mass, p, q, col[0])
        r.register_body(col[1], mass, p, q, (-1, -1, -1), (1, 1, 1))
    t = w.tables()
    return w, r, oa.Shapes(t), t["bodies"]


def mesh_pairs(bodies, sh):
    aabbs = oa.update_aabbs(oa.oracle(), "orc_", bodies, sh)
    small = np.nonzero(bodies["invMass"] != 0)[0].astype(np.int32)
    large = np.nonzero(bodies["invMass"] == 0)[0].astype(np.int32)
    n, pairs = oa.brute_force_pairs(oa.oracle(), "orc_", aabbs, small, large, 1 << 18)
    keep = (pairs["x"] == 0) | (pairs["y"] == 0)
    return aabbs, pairs[keep]


def key_sort(c):
    keys = tuple(c["worldPosB"][:, k, j] for k in range(4) for j in range(4)) + tuple(c["worldNormalOnB"][:, j] for j in range(4)) + (np.abs(c["bodyB"]),)
    return c[np.lexsort(keys)]


@pytest.mark.parametrize("seed", [0, 1])
def test_mesh_tables_match_reference(seed):
    w, r, sh, bodies = build_both(seed, n=4)
    cv_r, cv_m = r.table(2, capi.convex_t)[0], sh.convex[0]
    for f in ("faceOffset", "numFaces", "numVertices", "vertexOffset", "uniqueEdgesOffset", "numUniqueEdges"):
        assert cv_r[f] == cv_m[f], f
    nf, nv = int(cv_r["numFaces"]), int(cv_r["numVertices"])
    fr, fm = r.table(5, capi.face_t)[:nf], sh.faces[:nf]
    assert np.array_equal(fr["plane"].view(np.uint32), fm["plane"].view(np.uint32))
    assert np.array_equal(fr["indexOffset"], fm["indexOffset"]) and np.array_equal(fr["numIndices"], fm["numIndices"])
    assert np.array_equal(r.table(3, np.dtype(("f4", 4)))[:nv, :3].view(np.uint32), sh.vertices[:nv, :3].view(np.uint32))
    assert np.array_equal(r.table(6, np.dtype("i4"))[: 3 * nf], sh.indices[: 3 * nf])
    cr, cm = r.table(0, capi.collidable_t)[0], sh.collidables[0]
    assert cr["shapeType"] == cm["shapeType"] == capi.SHAPE_CONCAVE_TRIMESH and cr["shapeIndex"] == cm["shapeIndex"]
    ar, am = r.table(1, capi.aabb_t)[0], sh.local_aabbs[0]
    assert np.array_equal(ar["min"][:3].view(np.uint32), am["min"][:3].view(np.uint32)) and np.array_equal(ar["max"][:3].view(np.uint32), am["max"][:3].view(np.uint32))
    # the quantized BVH tables of every shape registered so far (the trimesh's b3OptimizedBvh, the compounds' b3QuantizedBvh):
    # headers, 16-byte nodes and subtree headers equal the reference's byte for byte
    assert_bvh_tables_equal(w, r)


def assert_bvh_tables_equal(w, r):
    ir, im = r.table(8, capi.bvh_info_t), w.table("bvh_infos")
    assert len(ir) == len(im) >= 1
    for f in ("aabbMin", "aabbMax", "quantization"):
        assert np.array_equal(np.asarray(ir[f])[:, :3].view(np.uint32), np.asarray(im[f])[:, :3].view(np.uint32)), f
    for f in ("numNodes", "numSubTrees", "nodeOffset", "subTreeOffset"):
        assert np.array_equal(ir[f], im[f]), f
    nr, nm = r.table(9, capi.bvh_node_t), w.table("bvh_nodes")
    assert len(nr) == len(nm) >= int(ir["numNodes"].sum())  # (a compound's header counts its children, its table has 2 n slots)
    for f in ("qmin", "qmax", "escapeIndexOrTriangleIndex"):
        assert np.array_equal(nr[f], nm[f]), f
    sr, sm = r.table(10, capi.bvh_subtree_t), w.table("bvh_subtrees")
    assert len(sr) == len(sm) == int(ir["numSubTrees"].sum())
    for f in ("qmin", "qmax", "rootNodeIndex", "subtreeSize"):
        assert np.array_equal(sr[f], sm[f]), f


@pytest.mark.parametrize("nx,nz,amp", [(1, 1, 0.0), (3, 2, 0.5), (40, 40, 2.0), (64, 33, 1.0)])
def test_quantized_bvh_tables_of_trimeshes_equal_the_reference(nx, nz, amp):
    """b3OptimizedBvh::build through registerConcaveMesh (b3GpuNarrowPhase.cpp:521-605) for meshes from 2 triangles (one leaf
    pair, one subtree header) to 5 120 triangles (many subtree headers), and two meshes in one world (offsets)"""
    cfg = capi.default_config(64)
    w = capi.World(cfg, device=-1)
    r = oa.RefNarrowphase(cfg)
    verts, tris = scenes.heightfield_mesh(nx, nz, cell=0.7, amplitude=amp, freq=0.9)
    assert w.register_concave(verts, tris) == r.register_concave(verts, tris)
    verts2, tris2 = scenes.heightfield_mesh(5, 7, cell=1.3, amplitude=0.3, freq=0.4)
    assert w.register_concave(verts2, tris2) == r.register_concave(verts2, tris2)
    assert_bvh_tables_equal(w, r)
    n = w.table("bvh_infos")
    assert n["numNodes"][0] == 2 * (len(tris) // 3)
    if len(tris) // 3 > 200:
        assert n["numSubTrees"][0] > 4


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_concave_contacts_bit_exact_vs_host_twins(seed):
    w, r, sh, bodies = build_both(seed)
    aabbs, pairs = mesh_pairs(bodies, sh)
    assert len(pairs) > 20
    r.concave_host_twins(True)
    try:
        ref = r.compute_contacts(bodies, pairs, aabbs, 1 << 16)
    finally:
        r.concave_host_twins(False)
    mine, ncand = oa.concave_contacts_oracle(pairs, bodies, sh, aabbs, 1 << 16)
    assert ncand > len(mine) > 30
    assert len(ref) == len(mine)
    a, b = key_sort(mine), key_sort(ref)
    assert np.array_equal(a["bodyA"], b["bodyA"]) and np.array_equal(a["bodyB"], b["bodyB"])
    assert np.array_equal(a["worldNormalOnB"].view(np.uint32), b["worldNormalOnB"].view(np.uint32))
    npts = a["worldNormalOnB"][:, 3].astype(int)
    assert npts.min() >= 1 and npts.max() <= 4
    for k in range(4):
        m = npts > k
        assert np.array_equal(a["worldPosB"][m, k].view(np.uint32), b["worldPosB"][m, k].view(np.uint32)), k
    assert np.array_equal(a["frictionCmp"], b["frictionCmp"])
    assert np.all(a["childA"] == -1) and np.all(a["childB"] == -1) and np.all(b["childB"] == -1)


@pytest.mark.parametrize("counts", [(1,), (2, 3), (7, 90, 5), (200,)])
def test_quantized_bvh_tables_of_compounds_equal_the_reference(counts):
    """b3QuantizedBvh::buildInternal through registerCompoundShape (b3GpuNarrowPhase.cpp:370-515): compounds from one child (a
    single leaf) to 200 children (400 node slots: subtree headers), several per world"""
    cfg = capi.default_config(64)
    cfg["maxCompoundChildShapes"] = 4096
    w = capi.World(cfg, device=-1)
    r = oa.RefNarrowphase(cfg)
    rc = r.register_convex_points(scenes.box_points(0.4))
    cv = r.table(2, capi.convex_t)[-1]
    vv = r.table(3, np.dtype(("f4", 4)))[cv["vertexOffset"]: cv["vertexOffset"] + cv["numVertices"]]
    faces = r.table(5, capi.face_t)[cv["faceOffset"]: cv["faceOffset"] + cv["numFaces"]].copy()
    idx = r.table(6, np.dtype("i4"))
    edges = r.table(4, np.dtype(("f4", 4)))[cv["uniqueEdgesOffset"]: cv["uniqueEdgesOffset"] + cv["numUniqueEdges"]]
    poly = np.zeros(1, capi.convex_t)
    poly[0] = cv
    wc = w.register_convex(vv, faces, idx[: int((faces["indexOffset"] + faces["numIndices"]).max())], edges, poly)
    rng = np.random.default_rng(len(counts))
    for k in counts:
        offs = rng.uniform(-6, 6, (k, 3))
        orns = [scenes.random_quat(rng) for _ in range(k)]
        a = w.register_compound(scenes.compound_children(wc, offs, orns))
        b = r.register_compound(scenes.compound_children(rc, offs, orns))
        assert a == b
    assert_bvh_tables_equal(w, r)
    if max(counts) >= 90:
        assert w.table("bvh_infos")["numSubTrees"].max() >= 2
