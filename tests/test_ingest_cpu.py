"""Wavefront ingestion (b3b200_register_concave_obj) on a host-only world: same tables as registering the triangle soup
that ConcaveScene::createConcaveMesh would build (examples/OpenCL/rigidbody/ConcaveScene.cpp:28-158)."""
import numpy as np
import pytest

from bullet3_b200 import capi

OBJ = """# a 2 x 2 heightfield patch, mixed face syntax
v 0 0 0
v 1 0.2 0
v 2 0 0
v 0 0.1 1
v 1 0.5 1
v 2 0.1 1
v 0 0 2
v 1 0.2 2
v 2 0 2
vn 0 1 0
vt 0.5 0.5
f 1 4 5 2
f 2/1 5/1 6/1 3/1
f 4//1 7//1 8//1
f 4/1/1 8/1/1 5/1/1
f -5 -2 -1 -4
"""


def expected_soup(shift, scaling):
    v = np.array([[0, 0, 0], [1, 0.2, 0], [2, 0, 0], [0, 0.1, 1], [1, 0.5, 1], [2, 0.1, 1], [0, 0, 2], [1, 0.2, 2], [2, 0, 2]], np.float32)
    polys = [[0, 3, 4, 1], [1, 4, 5, 2], [3, 6, 7], [3, 7, 4], [4, 7, 8, 5]]
    tris = []
    for p in polys:
        for k in range(2, len(p)):
            tris.append([p[0], p[k - 1], p[k]])
    soup = (v[np.array(tris).reshape(-1)] + np.float32(shift)) * np.float32(scaling)
    return soup.astype(np.float32), np.arange(len(soup), dtype=np.int32)


@pytest.mark.parametrize("shift,scaling", [((0, 0, 0), (1, 1, 1)), ((0.5, -1.0, 2.0), (2.0, 1.5, 0.5))])
def test_obj_ingestion_equals_registering_the_soup(tmp_path, shift, scaling):
    path = tmp_path / "patch.obj"
    path.write_text(OBJ)
    a = capi.World(capi.default_config(64), device=-1)
    b = capi.World(capi.default_config(64), device=-1)
    ca = a.register_concave_obj(path, shift, scaling)
    soup, idx = expected_soup(shift, scaling)
    assert len(idx) == 3 * 8
    cb = b.register_concave(soup, idx)
    assert ca == cb
    ta, tb = a.tables(), b.tables()
    for k in ("collidables", "local_aabbs", "convex", "vertices", "faces", "indices"):
        assert np.array_equal(np.asarray(ta[k]).view(np.uint8), np.asarray(tb[k]).view(np.uint8)), k


def test_obj_ingestion_errors(tmp_path):
    w = capi.World(capi.default_config(64), device=-1)
    with pytest.raises(capi.B3Error):
        w.register_concave_obj(tmp_path / "missing.obj")
    p = tmp_path / "nofaces.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\n")
    with pytest.raises(capi.B3Error):
        w.register_concave_obj(p)
    p = tmp_path / "badindex.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(capi.B3Error):
        w.register_concave_obj(p)
