"""The C-ABI library loads and exports every symbol include/b3b200.h declares;
POD layouts match the reference.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bullet3_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b3b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*|long long)\s+(b3b200_[a-z0-9_]+)\s*\(", src, re.M)))


def test_header_symbols_exported():
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) > 50
    for n in names:
        assert hasattr(lib, n), "libb3b200.so does not export %s" % n
    assert sorted(capi.SYMBOLS) == names


def test_pod_sizes_match_reference_abi():
    for name, (dt, size) in capi.ABI_SIZES.items():
        assert dt.itemsize == size, name
    assert capi.rigid_body_t.fields["collidableIdx"][1] == 64
    assert capi.rigid_body_t.fields["invMass"][1] == 68
    assert capi.contact4_t.fields["worldNormalOnB"][1] == 64
    assert capi.contact4_t.fields["frictionCmp"][1] == 82
    assert capi.contact4_t.fields["bodyA"][1] == 88
    assert capi.contact4_t.fields["childA"][1] == 96
    assert capi.constraint4_t.fields["jacCoeffInv"][1] == 96
    assert capi.constraint4_t.fields["fJacCoeffInv"][1] == 144
    assert capi.constraint4_t.fields["bodyA"][1] == 160
    assert capi.convex_t.fields["faceOffset"][1] == 68
    assert capi.convex_t.fields["numUniqueEdges"][1] == 88


def test_pod_sizes_match_compiled_reference():
    import oracle_api

    if not oracle_api.ref_available():
        pytest.skip("oracle/_ref not built")
    out = np.zeros(16, np.int32)
    n = oracle_api.ref().ref_sizes(capi.ptr(out), 16)
    assert list(out[:n]) == [80, 96, 16, 48, 32, 96, 32, 16, 112, 176]


def test_config_default_matches_b3Config():
    cfg = capi.default_config()
    assert cfg["maxConvexBodies"][0] == 128 * 1024
    assert cfg["maxBroadphasePairs"][0] == 16 * 128 * 1024
    assert cfg["maxContactCapacity"][0] == 16 * 128 * 1024
    assert cfg["compoundPairCapacity"][0] == 1024 * 1024
    assert cfg["maxVerticesPerFace"][0] == 64
    assert cfg["maxTriConvexPairCapacity"][0] == 256 * 1024


def test_host_only_world_registers_and_refuses_gpu_work():
    from bullet3_b200 import scenes

    w = capi.World(capi.default_config(64), device=-1)
    col = w.register_convex_points(scenes.box_points(0.5))
    assert col == 0
    b = w.register_instance(1.0, (0, 1, 0), scenes.IDENT, col)
    assert b == 0
    with pytest.raises(capi.B3Error):
        w.register_instance(1.0, (0, 1, 0), scenes.IDENT, 7)  # bad collidable -> -1 like the reference
    with pytest.raises(capi.B3Error):
        w.upload()
    t = w.tables()
    assert len(t["vertices"]) == 8 and len(t["faces"]) == 6 and len(t["unique_edges"]) == 3
    assert t["convex"]["numVertices"][0] == 8
    w.close()


def test_capacity_errors_return_minus_one():
    from bullet3_b200 import scenes

    w = capi.World(capi.default_config(2), device=-1)
    col = w.register_convex_points(scenes.box_points(0.5))
    w.register_instance(1.0, (0, 0, 0), scenes.IDENT, col)
    w.register_instance(1.0, (0, 2, 0), scenes.IDENT, col)
    with pytest.raises(capi.B3Error) as e:
        w.register_instance(1.0, (0, 4, 0), scenes.IDENT, col)
    assert "exceeding" in str(e.value)
    w.close()


def test_settings_and_state_errors_without_a_device(tmp_path):
    """argument checking of the newer entry points on a host-only world: settings reject bad modes, everything that needs
    device state reports a state error instead of touching a device (there is no CPU fallback)"""
    from bullet3_b200 import scenes

    L = capi.lib()
    w = capi.World(capi.default_config(64), device=-1)
    col = w.register_convex_points(scenes.box_points(0.5))
    w.register_instance(1.0, (0, 1, 0), scenes.IDENT, col)
    w.register_instance(1.0, (0, 3, 0), scenes.IDENT, col)
    # settings: valid modes accepted, others refused
    for mode, ok in ((0, True), (1, True), (2, False), (-1, False)):
        assert (L.b3b200_set_colouring(w.h, mode) == 0) == ok
    for mode, ok in ((-1, True), (0, True), (1, True), (2, False), (-2, False)):
        assert (L.b3b200_set_ray_accel(w.h, mode) == 0) == ok
    # joints live on the host until the first solve: creation / removal / query work without a device
    uid = w.create_p2p_constraint(0, 1, (0, 1, 0), (0, -1, 0))
    assert uid == 0 and w.num_constraints == 1
    assert len(w.joints()) == 1 and w.joints()["rbB"][0] == 1
    w.remove_constraint(uid)
    assert w.num_constraints == 0
    with pytest.raises(capi.B3Error):
        w.create_p2p_constraint(0, 9, (0, 0, 0), (0, 0, 0))  # no such body
    # device work is refused
    with pytest.raises(capi.B3Error):
        w.cast_rays(np.zeros((1, 3)), np.ones((1, 3)))
    with pytest.raises(capi.B3Error):
        w.checkpoint_save(tmp_path / "x.b3cp")
    with pytest.raises(capi.B3Error):
        w.checkpoint_load(tmp_path / "x.b3cp")
    with pytest.raises(capi.B3Error):
        w.solve_joints()
    assert L.b3b200_copy_transforms(w.h, None, 0) != 0
    assert L.b3b200_solve_contacts_device(w.h, 1, None, None, 0, None, -1) != 0
    ids = np.zeros(2, np.int32)
    assert L.b3b200_halo_set_ids(w.h, capi.ptr(ids), 2) != 0
    assert L.b3b200_halo_adopt(w.h, None, 0, None) != 0
    # null handles never crash
    assert L.b3b200_set_colouring(None, 0) != 0 and L.b3b200_cast_rays(None, None, 0, None) != 0
    assert L.b3b200_checkpoint_save(None, b"x") != 0 and L.b3b200_register_concave_obj(None, b"x", None, None) == -1
    w.close()


def test_round2_entry_points_without_a_device():
    """argument / state checking of the entry points added in round 2, on a host-only world (no GPU work happens)"""
    from bullet3_b200 import scenes

    L = capi.lib()
    w = capi.World(capi.default_config(64), device=-1)
    col = w.register_convex_points(scenes.box_points(0.5))
    # batched worlds: ids are recorded at registration
    assert w.num_worlds() == 1
    w.register_instance(1.0, (0, 1, 0), scenes.IDENT, col)
    w.set_current_world(3)
    w.register_instance(1.0, (0, 1, 0), scenes.IDENT, col)
    w.register_instance(0.0, (0, -1, 0), scenes.IDENT, col)
    assert w.num_worlds() == 4 and w.body_worlds().tolist() == [0, 3, 3]
    assert L.b3b200_set_current_world(w.h, -1) != 0 and L.b3b200_set_current_world(None, 0) != 0 and L.b3b200_num_worlds(None) < 0
    # step graphs are a setting; stepping itself needs a device
    assert L.b3b200_set_step_graphs(w.h, 0) == 0 and L.b3b200_set_step_graphs(w.h, 1) == 0 and L.b3b200_set_step_graphs(None, 1) != 0
    buf = np.zeros(3, capi.rigid_body_t)
    assert L.b3b200_step_host_async(w.h, C.c_float(1 / 60), capi.ptr(buf), capi.ptr(buf), 3) != 0  # host-only world
    assert L.b3b200_step_host_async(None, C.c_float(1 / 60), capi.ptr(buf), capi.ptr(buf), 3) != 0
    assert L.b3b200_step_host_wait(w.h) != 0
    # the MPR stage checks its arguments before touching a device
    n = C.c_int(0)
    assert L.b3b200_mpr_penetration(0, None, 5, None, 0, None, 0, None, 0, None, 0, None, None, None, 0, C.byref(n), None) != 0
    assert L.b3b200_mpr_penetration(0, None, -1, None, 0, None, 0, None, 0, None, 0, None, None, None, 0, C.byref(n), None) != 0
    assert L.b3b200_mpr_penetration(0, None, 0, None, 0, None, 0, None, 0, None, 0, None, None, None, 0, None, None) != 0
    # quantized BVH tables are host tables: available without a device
    ch = scenes.compound_children(col, [(0, 0, 0), (1, 0, 0), (0, 1, 0)])
    w.register_compound(ch)
    info = w.table("bvh_infos")
    assert len(info) == 1 and info["numNodes"][0] == 3 and info["numSubTrees"][0] == 1 and len(w.table("bvh_nodes")) == 6
    w.close()


def test_batched_box_worlds_scene_on_a_host_only_world():
    """BASELINE configs[4](i) recipe: every world is 8 x 4 x 8 cubes + its own static ground box, all worlds at the same coordinates"""
    from bullet3_b200 import scenes

    w = capi.World(capi.default_config(5 * 257 + 8), device=-1)
    per_world = scenes.batched_box_worlds(w, 5)
    assert per_world == 257 and w.num_bodies == 5 * 257 and w.num_worlds() == 5
    wid = w.body_worlds()
    assert np.array_equal(wid, np.repeat(np.arange(5), 257))
    b = w.table("bodies")
    assert (b["invMass"][wid == 3] == 0).sum() == 1 and b["invMass"][257 * 3] == 0  # the ground box comes first in its world
    assert np.array_equal(b["pos"][:257], b["pos"][257 * 4:])  # same coordinates in every world
    w.close()
