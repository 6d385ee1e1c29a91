"""orc_cast_rays against the reference's own b3GpuRaycast::castRaysHost (b3GpuRaycast.cpp:173-246), compiled
unmodified in oracle/_ref/libb3refcl.so.  Hull-only scenes are compared bit for bit; sphere scenes are compared where
the reference's missing `break` after SHAPE_SPHERE (b3GpuRaycast.cpp:205) cannot change the answer."""
import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi, scenes
from test_shapes_cpu import build_both

pytestmark = pytest.mark.skipif(not oa.refcl_available(), reason="oracle/_ref/libb3refcl.so not built")


def random_rays(rng, n, extent=4.0):
    frm = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    to = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    return frm, to


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_hull_rays_bit_exact_vs_reference_host(seed):
    w, r, sh, bodies = build_both(seed, n=80, with_compounds=False, plane=False)
    rng = np.random.default_rng(100 + seed)
    frm, to = random_rays(rng, 600)
    ref = r.cast_rays(frm, to, bodies)
    mine = oa.cast_rays_oracle(frm, to, bodies, sh)
    assert (ref["hitBody"] >= 0).sum() > 100 and (ref["hitBody"] < 0).sum() > 50
    assert np.array_equal(mine["hitBody"], ref["hitBody"])
    assert np.array_equal(bits(mine["hitFraction"]), bits(ref["hitFraction"]))
    hit = ref["hitBody"] >= 0
    assert np.array_equal(bits(mine["hitPoint"][hit, :3]), bits(ref["hitPoint"][hit, :3]))
    assert np.array_equal(bits(mine["hitNormal"][hit, :3]), bits(ref["hitNormal"][hit, :3]))


def test_rays_respect_the_callers_max_fraction():
    w, r, sh, bodies = build_both(3, n=80, with_compounds=False, plane=False)
    rng = np.random.default_rng(7)
    frm, to = random_rays(rng, 400)
    full = oa.cast_rays_oracle(frm, to, bodies, sh)
    ref = r.cast_rays(frm, to, bodies, max_fraction=0.5)
    mine = oa.cast_rays_oracle(frm, to, bodies, sh, max_fraction=0.5)
    assert np.array_equal(mine["hitBody"], ref["hitBody"])
    assert np.array_equal(bits(mine["hitFraction"]), bits(ref["hitFraction"]))
    # a hit beyond the cap is no hit at all, and the record is left as the caller initialised it
    late = (full["hitBody"] >= 0) & (full["hitFraction"] >= 0.5)
    assert late.sum() > 10 and np.all(mine["hitBody"][late] == -1) and np.all(mine["hitFraction"][late] == 0.5)


def test_ray_starting_inside_a_hull_misses_it():
    """rayConvex wants an entering plane (enterFraction stays -0.1 otherwise): b3GpuRaycast.cpp:166-167"""
    w, r, sh, bodies = build_both(4, n=20, with_compounds=False, plane=False)
    centre = bodies["pos"][:, :3].copy()
    to = centre + np.float32([0.05, 0.02, 0.01])
    ref = r.cast_rays(centre, to, bodies)
    mine = oa.cast_rays_oracle(centre, to, bodies, sh)
    assert np.array_equal(mine["hitBody"], ref["hitBody"])
    assert np.all(mine["hitBody"] != np.arange(len(bodies)))


def test_sphere_rays_match_reference_where_the_fallthrough_is_inert():
    """Spheres: the reference falls through into the convex test with the sphere collidable's m_shapeIndex (0), i.e. it also
    tests convex shape 0 at the sphere's transform.  Shape 0 here is a 1 cm box, far inside every sphere, so the
    fall-through can never produce the closer hit and both must agree bit for bit."""
    rng = np.random.default_rng(5)
    cfg = capi.default_config(1024)
    w = capi.World(cfg, device=-1)
    r = oa.RefNarrowphase(cfg)
    rc = r.register_convex_points(scenes.box_points(0.005))
    cv = r.table(2, capi.convex_t)[-1]
    verts = r.table(3, np.dtype(("f4", 4)))[cv["vertexOffset"]: cv["vertexOffset"] + cv["numVertices"]]
    faces = r.table(5, capi.face_t)[cv["faceOffset"]: cv["faceOffset"] + cv["numFaces"]].copy()
    idx_all = r.table(6, np.dtype("i4"))
    edges = r.table(4, np.dtype(("f4", 4)))[cv["uniqueEdgesOffset"]: cv["uniqueEdgesOffset"] + cv["numUniqueEdges"]]
    poly = np.zeros(1, capi.convex_t)
    poly[0] = cv
    assert w.register_convex(verts, faces, idx_all, edges, poly) == rc
    kinds = [(w.register_sphere(rad), r.register_sphere(rad)) for rad in (0.3, 0.6)]
    for i in range(60):
        p = rng.uniform(-3, 3, 3)
        k = kinds[i % 2]
        q = scenes.random_quat(rng)
        w.register_instance(1.0, p, q, k[0])
        r.register_body(k[1], 1.0, p, q, (-1, -1, -1), (1, 1, 1))
    t = w.tables()
    sh, bodies = oa.Shapes(t), t["bodies"]
    frm, to = random_rays(rng, 500, 4.0)
    ref = r.cast_rays(frm, to, bodies)
    mine = oa.cast_rays_oracle(frm, to, bodies, sh)
    assert (ref["hitBody"] >= 0).sum() > 80
    assert np.array_equal(mine["hitBody"], ref["hitBody"])
    assert np.array_equal(bits(mine["hitFraction"]), bits(ref["hitFraction"]))
    hit = ref["hitBody"] >= 0
    assert np.array_equal(bits(mine["hitNormal"][hit, :3]), bits(ref["hitNormal"][hit, :3]))
    assert np.array_equal(bits(mine["hitPoint"][hit, :3]), bits(ref["hitPoint"][hit, :3]))
