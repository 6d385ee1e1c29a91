"""The reference's own golden vectors for the narrowphase: four serialized launches of mprPenetrationKernel
(data/unittest_data.zip, expected contact totals 0 / 1 / 46 / 98,
test/OpenCL/AllBullet3Kernels/testExecuteBullet3NarrowphaseKernels.cpp:397-413), trimmed to tests/golden/mpr_*.npz by
tests/golden/make_mpr_golden.py.  CPU: the reference's b3MprPenetration (shared/b3MprPenetration.h, compiled unmodified into
oracle/_ref/libb3ref.so) reproduces the totals here.  GPU: b3b200_mpr_penetration (csrc/mpr.cu) reproduces the totals and agrees
with the reference header pair by pair, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS = ["60", "61", "70", "128"]


def load(tag):
    return np.load(os.path.join(GOLD, "mpr_%s.npz" % tag))


def ref_mpr(g):
    pairs = g["pairs"].copy()
    sep = g["sep_normals"].copy()
    has = g["has_sep_axis"].copy()
    cap = int(g["capacity"])
    contacts = np.zeros(4096, capi.contact4_t)
    n = C.c_int(int(g["count0"]))
    res = np.zeros(len(pairs), capi.mpr_result_t)
    bodies, coll, convex, verts = (np.ascontiguousarray(g[k]) for k in ("bodies", "collidables", "convex", "vertices"))
    oa.ref().ref_mpr_kernel(capi.ptr(pairs), len(pairs), capi.ptr(bodies), capi.ptr(coll), capi.ptr(convex), capi.ptr(verts), capi.ptr(sep), capi.ptr(has),
                            capi.ptr(contacts), min(cap, len(contacts)), C.byref(n), capi.ptr(res))
    return pairs, sep, has, contacts[: n.value], n.value, res


@pytest.mark.skipif(not oa.ref_available(), reason="oracle/_ref/libb3ref.so not built")
@pytest.mark.parametrize("tag", TAGS)
def test_reference_header_reproduces_the_golden_totals(tag):
    g = load(tag)
    _, _, has, contacts, total, res = ref_mpr(g)
    assert total == int(g["expected_total"])
    assert (res["result"] == 0).sum() == total - int(g["count0"])
    assert np.all(contacts["worldNormalOnB"][:, 3] == 1)


def test_fixtures_are_what_the_generator_describes():
    for tag in TAGS:
        g = load(tag)
        assert g["pairs"]["x"].max() < len(g["bodies"]) and g["pairs"]["y"].max() < len(g["bodies"])
        assert g["bodies"]["collidableIdx"].max() < len(g["collidables"])
        last = g["convex"][-1]
        assert int(last["vertexOffset"]) + int(last["numVertices"]) == len(g["vertices"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_cuda_mpr_kernel_reproduces_the_golden_totals_and_the_reference_pair_by_pair(tag):
    g = load(tag)
    pairs, sep, has, contacts, total, res = capi.mpr_penetration(g["pairs"], g["bodies"], g["collidables"], g["convex"], g["vertices"], g["sep_normals"],
                                                                 g["has_sep_axis"], min(int(g["capacity"]), 4096), int(g["count0"]))
    assert total == int(g["expected_total"])
    if not oa.ref_available():
        return
    rp, rsep, rhas, rcontacts, rtotal, rres = ref_mpr(g)
    assert np.array_equal(res["result"], rres["result"])
    assert np.array_equal(has, rhas)
    hit = rres["result"] == 0
    for f in ("depth", "dir", "pos"):
        assert np.array_equal(res[f][hit].view(np.uint32), rres[f][hit].view(np.uint32)), f
    assert np.array_equal(sep[rhas == 1].view(np.uint32), rsep[rhas == 1].view(np.uint32))
    assert np.array_equal(pairs["z"] >= 0, rp["z"] >= 0)
    # contacts are appended with an atomic: compare through the pairs' contact indices
    for i in np.nonzero(hit)[0]:
        a, b = contacts[pairs["z"][i]], rcontacts[rp["z"][i]]
        assert np.array_equal(a["worldPosB"][0].view(np.uint32), b["worldPosB"][0].view(np.uint32))
        assert np.array_equal(a["worldNormalOnB"].view(np.uint32), b["worldNormalOnB"].view(np.uint32))
        assert a["bodyA"] == b["bodyA"] and a["bodyB"] == b["bodyB"] and a["batchIdx"] == b["batchIdx"] and a["frictionCmp"] == b["frictionCmp"]
