"""ctypes binding of the C ABI in include/b3b200.h (libb3b200.so).

This is plumbing for tests and bench.py: the product is the shared library and
the C++ drop-in classes under csrc/host/.  There is NO CPU fallback -- loading
fails loudly when the CUDA library has not been built.

POD layouts are numpy structured dtypes that mirror include/b3b200_types.h
(sizes are asserted at import; reference layouts: SURVEY.md Appendix A).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb3b200.so")

# ---------------------------------------------------------------- dtypes
F4 = ("f4", 4)
rigid_body_t = np.dtype([("pos", *F4), ("quat", *F4), ("linVel", *F4), ("angVel", *F4),
                         ("collidableIdx", "i4"), ("invMass", "f4"), ("restitution", "f4"), ("friction", "f4")])
inertia_t = np.dtype([("invInertiaWorld", "f4", (3, 4)), ("initInvInertia", "f4", (3, 4))])
collidable_t = np.dtype([("numChildShapes", "i4"), ("radius", "f4"), ("shapeType", "i4"), ("shapeIndex", "i4")])
child_shape_t = np.dtype([("childPosition", *F4), ("childOrientation", *F4), ("shapeIndex", "i4"),
                          ("numChildShapes", "i4"), ("collidableShapeIndex", "i4"), ("shapeType", "i4")])
face_t = np.dtype([("plane", *F4), ("indexOffset", "i4"), ("numIndices", "i4"), ("pad1", "i4"), ("pad2", "i4")])
convex_t = np.dtype([("localCenter", *F4), ("extents", *F4), ("mC", *F4), ("mE", *F4), ("radius", "f4"),
                     ("faceOffset", "i4"), ("numFaces", "i4"), ("numVertices", "i4"), ("vertexOffset", "i4"),
                     ("uniqueEdgesOffset", "i4"), ("numUniqueEdges", "i4"), ("unused", "i4")])
aabb_t = np.dtype([("min", "f4", 3), ("minIndex", "i4"), ("max", "f4", 3), ("maxIndex", "i4")])
int4_t = np.dtype([("x", "i4"), ("y", "i4"), ("z", "i4"), ("w", "i4")])
mpr_result_t = np.dtype([("result", "i4"), ("depth", "f4"), ("dir", "f4", 3), ("pos", "f4", 3)])
contact4_t = np.dtype([("worldPosB", "f4", (4, 4)), ("worldNormalOnB", *F4), ("restitutionCmp", "u2"), ("frictionCmp", "u2"),
                       ("batchIdx", "i4"), ("bodyA", "i4"), ("bodyB", "i4"), ("childA", "i4"), ("childB", "i4"),
                       ("unused1", "i4"), ("unused2", "i4")])
constraint4_t = np.dtype([("linear", *F4), ("worldPos", "f4", (4, 4)), ("center", *F4), ("jacCoeffInv", *F4), ("b", *F4),
                          ("appliedRambdaDt", *F4), ("fJacCoeffInv", "f4", 2), ("fAppliedRambdaDt", "f4", 2),
                          ("bodyA", "u4"), ("bodyB", "u4"), ("batchIdx", "i4"), ("paddings", "u4")])
config_t = np.dtype([(n, "i4") for n in (
    "maxConvexBodies", "maxConvexShapes", "maxBroadphasePairs", "maxContactCapacity", "compoundPairCapacity",
    "maxVerticesPerFace", "maxFacesPerShape", "maxConvexVertices", "maxConvexIndices", "maxConvexUniqueEdges",
    "maxCompoundChildShapes", "maxTriConvexPairCapacity")])
ray_info_t = np.dtype([("from", "f4", 4), ("to", "f4", 4)])
ray_hit_t = np.dtype([("hitFraction", "f4"), ("hitBody", "i4"), ("hitResult1", "i4"), ("hitResult2", "i4"), ("hitPoint", "f4", 4), ("hitNormal", "f4", 4)])
joint_t = np.dtype([("constraintType", "i4"), ("rbA", "i4"), ("rbB", "i4"), ("breakingImpulseThreshold", "f4"), ("pivotInA", "f4", 4), ("pivotInB", "f4", 4),
                    ("relTargetAB", "f4", 4), ("flags", "i4"), ("uid", "i4"), ("padding", "i4", 2)])
sort_data_t = np.dtype([("key", "u4"), ("value", "u4")])
bvh_node_t = np.dtype([("qmin", "u2", 3), ("qmax", "u2", 3), ("escapeIndexOrTriangleIndex", "i4")])
bvh_subtree_t = np.dtype([("qmin", "u2", 3), ("qmax", "u2", 3), ("rootNodeIndex", "i4"), ("subtreeSize", "i4"), ("padding", "i4", 3)])
bvh_info_t = np.dtype([("aabbMin", *F4), ("aabbMax", *F4), ("quantization", *F4), ("numNodes", "i4"), ("numSubTrees", "i4"),
                       ("nodeOffset", "i4"), ("subTreeOffset", "i4")])

ABI_SIZES = {"rigid_body": (rigid_body_t, 80), "inertia": (inertia_t, 96), "collidable": (collidable_t, 16),
             "child_shape": (child_shape_t, 48), "face": (face_t, 32), "convex": (convex_t, 96), "aabb": (aabb_t, 32),
             "int4": (int4_t, 16), "contact4": (contact4_t, 112), "constraint4": (constraint4_t, 176), "config": (config_t, 48),
             "sort_data": (sort_data_t, 8), "bvh_node": (bvh_node_t, 16), "bvh_subtree": (bvh_subtree_t, 32), "bvh_info": (bvh_info_t, 64)}
for _n, (_t, _s) in ABI_SIZES.items():
    assert _t.itemsize == _s, (_n, _t.itemsize, _s)

SHAPE_CONVEX_HULL, SHAPE_PLANE, SHAPE_CONCAVE_TRIMESH, SHAPE_COMPOUND, SHAPE_SPHERE = 3, 4, 5, 6, 7
BP_SAP, BP_GRID = 0, 1
SOLVER_PGS, SOLVER_JACOBI = 0, 1

# every symbol include/b3b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "b3b200_last_error", "b3b200_version", "b3b200_launch_count", "b3b200_config_default", "b3b200_create", "b3b200_destroy",
    "b3b200_reset", "b3b200_register_convex", "b3b200_register_convex_points", "b3b200_register_plane", "b3b200_register_sphere",
    "b3b200_register_compound", "b3b200_register_concave", "b3b200_register_instance", "b3b200_register_body", "b3b200_register_instances", "b3b200_upload", "b3b200_set_gravity",
    "b3b200_set_solver", "b3b200_set_broadphase", "b3b200_set_colouring", "b3b200_set_step_graphs", "b3b200_set_current_world", "b3b200_mpr_penetration", "b3b200_step_host_async", "b3b200_step_host_wait", "b3b200_num_worlds", "b3b200_get_body_worlds", "b3b200_set_contact_clip", "b3b200_set_angular_damping", "b3b200_write_bodies",
    "b3b200_readback_bodies", "b3b200_write_body", "b3b200_read_body", "b3b200_readback_inertias", "b3b200_num_bodies", "b3b200_step", "b3b200_step_n", "b3b200_synchronize",
    "b3b200_update_aabbs", "b3b200_find_pairs", "b3b200_compute_contacts", "b3b200_solve_contacts", "b3b200_solve_joints", "b3b200_create_p2p_constraint", "b3b200_create_fixed_constraint", "b3b200_remove_constraint",
    "b3b200_num_constraints", "b3b200_get_joints", "b3b200_cast_rays", "b3b200_set_ray_accel", "b3b200_solver_setup",
    "b3b200_solver_iterate", "b3b200_integrate", "b3b200_get_aabbs", "b3b200_get_pairs", "b3b200_get_contacts", "b3b200_set_contacts",
    "b3b200_get_constraints", "b3b200_get_batches", "b3b200_get_counters", "b3b200_get_work_counters", "b3b200_enable_stage_timing", "b3b200_stage_timings",
    "b3b200_device_buffer", "b3b200_get_table", "b3b200_device_to_host", "b3b200_halo_record_size", "b3b200_halo_pack", "b3b200_halo_unpack", "b3b200_halo_ghost_ids", "b3b200_halo_set_ids", "b3b200_halo_emigrate", "b3b200_halo_adopt", "b3b200_bp_create", "b3b200_bp_destroy", "b3b200_bp_create_proxy", "b3b200_bp_create_large_proxy",
    "b3b200_bp_write_aabbs", "b3b200_bp_set_aabbs", "b3b200_bp_calculate_pairs", "b3b200_bp_num_overlap", "b3b200_bp_get_pairs",
    "b3b200_bp_device_pairs", "b3b200_bp_device_aabbs", "b3b200_bp_last_ms", "b3b200_radix_sort_kv", "b3b200_radix_sort_keys",
    "b3b200_prefix_scan_u32", "b3b200_bound_search_count", "b3b200_fill_u32", "b3b200_bound_search", "b3b200_prefix_scan_float4", "b3b200_slab_unique_id", "b3b200_slab_init", "b3b200_slab_exchange", "b3b200_slab_step",
    "b3b200_slab_step_n", "b3b200_slab_last_counts", "b3b200_slab_shutdown",
    "b3b200_solve_contacts_device", "b3b200_register_concave_obj", "b3b200_checkpoint_save", "b3b200_checkpoint_load", "b3b200_copy_transforms",
]

_lib = None


class B3Error(RuntimeError):
    pass


def lib():
    """Load libb3b200.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B3Error("%s not built: run `make -C bullet3_b200/csrc` or __graft_entry__.build(); "
                          "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.b3b200_last_error.restype = C.c_char_p
        _lib.b3b200_launch_count.restype = C.c_longlong
    return _lib


def last_error():
    return lib().b3b200_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc < 0:
        raise B3Error("%s failed (%d): %s" % (what, rc, last_error()))
    return rc


def ptr(a):
    """pointer to a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def f3(v):
    return (C.c_float * 4)(float(v[0]), float(v[1]), float(v[2]), float(v[3]) if len(v) > 3 else 0.0)


def default_config(max_bodies=None, pairs_per_body=16):
    cfg = np.zeros(1, config_t)
    check(lib().b3b200_config_default(ptr(cfg)), "config_default")
    if max_bodies is not None:
        cfg["maxConvexBodies"] = max_bodies
        cfg["maxConvexShapes"] = max_bodies
        cfg["maxBroadphasePairs"] = pairs_per_body * max_bodies
        cfg["maxContactCapacity"] = pairs_per_body * max_bodies
    return cfg


class World:
    """Thin handle around b3b200_world (the b3GpuRigidBodyPipeline + b3GpuNarrowPhase + broadphase trio)."""

    def __init__(self, cfg=None, device=0, stream=None):
        self.L = lib()
        self.cfg = default_config() if cfg is None else cfg
        h = C.c_void_p()
        check(self.L.b3b200_create(ptr(self.cfg), int(device), C.c_void_p(stream or 0), C.byref(h)), "create")
        self.h = h

    def close(self):
        if self.h:
            self.L.b3b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shapes
    def register_convex(self, vertices, faces, indices, unique_edges, poly=None):
        vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 4)
        unique_edges = np.ascontiguousarray(unique_edges, np.float32).reshape(-1, 4)
        faces = np.ascontiguousarray(faces, face_t)
        indices = np.ascontiguousarray(indices, np.int32)
        if poly is None:
            poly = np.zeros(1, convex_t)
        r = self.L.b3b200_register_convex(self.h, ptr(vertices), len(vertices), ptr(faces), len(faces), ptr(indices), len(indices),
                                          ptr(unique_edges), len(unique_edges), ptr(poly))
        if r < 0:
            raise B3Error("register_convex: " + last_error())
        return r

    def register_convex_points(self, points, scaling=(1.0, 1.0, 1.0)):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        sc = (C.c_float * 3)(*[float(x) for x in scaling])
        r = self.L.b3b200_register_convex_points(self.h, ptr(pts), 12, len(pts), sc)
        if r < 0:
            raise B3Error("register_convex_points: " + last_error())
        return r

    def register_plane(self, normal, constant):
        r = self.L.b3b200_register_plane(self.h, f3(normal), C.c_float(constant))
        if r < 0:
            raise B3Error("register_plane: " + last_error())
        return r

    def register_sphere(self, radius):
        r = self.L.b3b200_register_sphere(self.h, C.c_float(radius))
        if r < 0:
            raise B3Error("register_sphere: " + last_error())
        return r

    def register_compound(self, children):
        children = np.ascontiguousarray(children, child_shape_t)
        r = self.L.b3b200_register_compound(self.h, ptr(children), len(children))
        if r < 0:
            raise B3Error("register_compound: " + last_error())
        return r

    def register_concave(self, vertices, tri_indices, scaling=(1.0, 1.0, 1.0)):
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(tri_indices, np.int32).reshape(-1)
        sc = (C.c_float * 3)(*[float(x) for x in scaling])
        r = self.L.b3b200_register_concave(self.h, ptr(v), len(v), ptr(i), len(i), sc)
        if r < 0:
            raise B3Error("register_concave: " + last_error())
        return r

    def register_concave_obj(self, path, shift=(0.0, 0.0, 0.0), scaling=(1.0, 1.0, 1.0)):
        """Wavefront .obj -> trimesh collidable (ConcaveScene::createConcaveMesh recipe)"""
        sh = (C.c_float * 3)(*[float(x) for x in shift])
        sc = (C.c_float * 3)(*[float(x) for x in scaling])
        r = self.L.b3b200_register_concave_obj(self.h, str(path).encode(), sh, sc)
        if r < 0:
            raise B3Error("register_concave_obj: " + last_error())
        return r

    # ---- bodies
    def register_instance(self, mass, position, orientation, collidable, user_index=0):
        r = self.L.b3b200_register_instance(self.h, C.c_float(mass), f3(position), f3(orientation), int(collidable), int(user_index))
        if r < 0:
            raise B3Error("register_instance: " + last_error())
        return r

    def register_instances(self, masses, positions, orientations, collidables):
        m = np.ascontiguousarray(masses, np.float32)
        n = len(m)
        p = np.zeros((n, 4), np.float32)
        p[:, :3] = np.asarray(positions, np.float32).reshape(n, -1)[:, :3]
        q = np.ascontiguousarray(np.asarray(orientations, np.float32).reshape(n, 4))
        c = np.ascontiguousarray(collidables, np.int32)
        r = self.L.b3b200_register_instances(self.h, n, ptr(m), ptr(p), ptr(q), ptr(c))
        if r < 0:
            raise B3Error("register_instances: " + last_error())
        return r

    def upload(self):
        check(self.L.b3b200_upload(self.h), "upload")

    def set_gravity(self, g):
        check(self.L.b3b200_set_gravity(self.h, f3(g)), "set_gravity")

    def set_solver(self, kind, iterations):
        check(self.L.b3b200_set_solver(self.h, int(kind), int(iterations)), "set_solver")

    def set_broadphase(self, kind):
        check(self.L.b3b200_set_broadphase(self.h, int(kind)), "set_broadphase")

    def step_host_async(self, dt, host_in, host_out):
        """host_in / host_out: rigid_body_t arrays in page-locked memory (or None)"""
        n = len(host_in) if host_in is not None else len(host_out)
        check(self.L.b3b200_step_host_async(self.h, C.c_float(dt), ptr(host_in) if host_in is not None else None, ptr(host_out) if host_out is not None else None, n), "step_host_async")

    def step_host_wait(self):
        check(self.L.b3b200_step_host_wait(self.h), "step_host_wait")

    def set_current_world(self, k):
        check(self.L.b3b200_set_current_world(self.h, int(k)), "set_current_world")

    def num_worlds(self):
        return int(self.L.b3b200_num_worlds(self.h))

    def body_worlds(self):
        out = np.zeros(int(self.L.b3b200_num_bodies(self.h)), np.int32)
        check(self.L.b3b200_get_body_worlds(self.h, ptr(out), len(out)), "get_body_worlds")
        return out

    def set_step_graphs(self, on):
        check(self.L.b3b200_set_step_graphs(self.h, int(bool(on))), "set_step_graphs")

    def set_colouring(self, mode):
        check(self.L.b3b200_set_colouring(self.h, int(mode)), "set_colouring")

    def set_contact_clip(self, min_dist, max_dist):
        check(self.L.b3b200_set_contact_clip(self.h, C.c_float(min_dist), C.c_float(max_dist)), "set_contact_clip")

    def set_angular_damping(self, d):
        check(self.L.b3b200_set_angular_damping(self.h, C.c_float(d)), "set_angular_damping")

    @property
    def num_bodies(self):
        return check(self.L.b3b200_num_bodies(self.h), "num_bodies")

    def write_bodies(self, bodies):
        bodies = np.ascontiguousarray(bodies, rigid_body_t)
        check(self.L.b3b200_write_bodies(self.h, ptr(bodies), len(bodies)), "write_bodies")

    def bodies(self):
        out = np.zeros(self.num_bodies, rigid_body_t)
        check(self.L.b3b200_readback_bodies(self.h, ptr(out), len(out)), "readback_bodies")
        return out

    def inertias(self):
        out = np.zeros(self.num_bodies, inertia_t)
        check(self.L.b3b200_readback_inertias(self.h, ptr(out), len(out)), "readback_inertias")
        return out

    # ---- step and stages
    def step(self, dt=1.0 / 60.0):
        check(self.L.b3b200_step(self.h, C.c_float(dt)), "step")

    def step_n(self, dt, n):
        check(self.L.b3b200_step_n(self.h, C.c_float(dt), int(n)), "step_n")

    def synchronize(self):
        check(self.L.b3b200_synchronize(self.h), "synchronize")

    def update_aabbs(self):
        check(self.L.b3b200_update_aabbs(self.h), "update_aabbs")

    def find_pairs(self):
        check(self.L.b3b200_find_pairs(self.h), "find_pairs")

    def compute_contacts(self):
        check(self.L.b3b200_compute_contacts(self.h), "compute_contacts")

    def solve_contacts(self):
        check(self.L.b3b200_solve_contacts(self.h), "solve_contacts")

    def solver_setup(self):
        check(self.L.b3b200_solver_setup(self.h), "solver_setup")

    def solver_iterate(self):
        check(self.L.b3b200_solver_iterate(self.h), "solver_iterate")

    def integrate(self, dt=1.0 / 60.0):
        check(self.L.b3b200_integrate(self.h, C.c_float(dt)), "integrate")

    # ---- results
    def aabbs(self):
        out = np.zeros(self.num_bodies, aabb_t)
        check(self.L.b3b200_get_aabbs(self.h, ptr(out), len(out)), "get_aabbs")
        return out

    def counters(self):
        out = np.zeros(8, np.int32)
        check(self.L.b3b200_get_counters(self.h, ptr(out)), "get_counters")
        return out

    # ---- joints
    def create_p2p_constraint(self, body_a, body_b, pivot_a, pivot_b, breaking_threshold=1e30):
        fa = (C.c_float * 3)(*[float(x) for x in pivot_a[:3]])
        fb = (C.c_float * 3)(*[float(x) for x in pivot_b[:3]])
        r = self.L.b3b200_create_p2p_constraint(self.h, int(body_a), int(body_b), fa, fb, C.c_float(breaking_threshold))
        if r < 0:
            raise B3Error("create_p2p_constraint: " + last_error())
        return r

    def create_fixed_constraint(self, body_a, body_b, pivot_a, pivot_b, rel_target_ab, breaking_threshold=1e30):
        fa = (C.c_float * 3)(*[float(x) for x in pivot_a[:3]])
        fb = (C.c_float * 3)(*[float(x) for x in pivot_b[:3]])
        fq = (C.c_float * 4)(*[float(x) for x in rel_target_ab[:4]])
        r = self.L.b3b200_create_fixed_constraint(self.h, int(body_a), int(body_b), fa, fb, fq, C.c_float(breaking_threshold))
        if r < 0:
            raise B3Error("create_fixed_constraint: " + last_error())
        return r

    def remove_constraint(self, uid):
        check(self.L.b3b200_remove_constraint(self.h, int(uid)), "remove_constraint")

    @property
    def num_constraints(self):
        return check(self.L.b3b200_num_constraints(self.h), "num_constraints")

    def joints(self):
        n = C.c_int(0)
        check(self.L.b3b200_get_joints(self.h, None, 0, C.byref(n)), "get_joints")
        out = np.zeros(n.value, joint_t)
        if n.value:
            check(self.L.b3b200_get_joints(self.h, ptr(out), n.value, C.byref(n)), "get_joints")
        return out

    def solve_joints(self):
        check(self.L.b3b200_solve_joints(self.h), "solve_joints")

    def cast_rays(self, ray_from, ray_to, max_fraction=1.0):
        """returns ray_hit_t array; hitBody = -1 where nothing was hit"""
        f = np.asarray(ray_from, np.float32).reshape(-1, 3)
        t = np.asarray(ray_to, np.float32).reshape(-1, 3)
        rays = np.zeros(len(f), ray_info_t)
        rays["from"][:, :3] = f
        rays["to"][:, :3] = t
        hits = np.zeros(len(f), ray_hit_t)
        hits["hitFraction"] = max_fraction
        hits["hitBody"] = -1
        check(self.L.b3b200_cast_rays(self.h, ptr(rays), len(rays), ptr(hits)), "cast_rays")
        return hits

    def set_ray_accel(self, mode):
        check(self.L.b3b200_set_ray_accel(self.h, int(mode)), "set_ray_accel")

    def work_counters(self):
        out = np.zeros(24, np.int32)
        check(self.L.b3b200_get_work_counters(self.h, ptr(out), 24), "get_work_counters")
        return out

    def pairs(self):
        n = C.c_int(0)
        check(self.L.b3b200_get_pairs(self.h, None, 0, C.byref(n)), "get_pairs")
        out = np.zeros(n.value, int4_t)
        if n.value:
            check(self.L.b3b200_get_pairs(self.h, ptr(out), n.value, C.byref(n)), "get_pairs")
        return out

    def contacts(self):
        n = C.c_int(0)
        check(self.L.b3b200_get_contacts(self.h, None, 0, C.byref(n)), "get_contacts")
        out = np.zeros(n.value, contact4_t)
        if n.value:
            check(self.L.b3b200_get_contacts(self.h, ptr(out), n.value, C.byref(n)), "get_contacts")
        return out

    def set_contacts(self, contacts):
        contacts = np.ascontiguousarray(contacts, contact4_t)
        check(self.L.b3b200_set_contacts(self.h, ptr(contacts), len(contacts)), "set_contacts")

    def constraints(self):
        n = C.c_int(0)
        check(self.L.b3b200_get_constraints(self.h, None, 0, C.byref(n)), "get_constraints")
        out = np.zeros(n.value, constraint4_t)
        if n.value:
            check(self.L.b3b200_get_constraints(self.h, ptr(out), n.value, C.byref(n)), "get_constraints")
        return out

    def batches(self):
        n = C.c_int(0)
        off = np.zeros(130, np.int32)
        check(self.L.b3b200_get_batches(self.h, ptr(off), len(off), C.byref(n)), "get_batches")
        return off[: n.value + 1].copy()

    def enable_stage_timing(self, on=True):
        check(self.L.b3b200_enable_stage_timing(self.h, int(bool(on))), "enable_stage_timing")

    def stage_timings(self):
        out = np.zeros(8, np.float32)
        check(self.L.b3b200_stage_timings(self.h, ptr(out)), "stage_timings")
        return out

    TABLES = {"collidables": (0, collidable_t), "local_aabbs": (1, aabb_t), "convex": (2, convex_t), "vertices": (3, np.dtype(("f4", 4))),
              "unique_edges": (4, np.dtype(("f4", 4))), "faces": (5, face_t), "indices": (6, np.dtype("i4")), "child_shapes": (7, child_shape_t),
              "bvh_infos": (8, bvh_info_t), "bvh_nodes": (9, bvh_node_t), "bvh_subtrees": (10, bvh_subtree_t), "bodies": (11, rigid_body_t),
              "inertias": (12, inertia_t)}

    def table(self, name):
        which, dt = self.TABLES[name]
        n = C.c_int(0)
        check(self.L.b3b200_get_table(self.h, which, None, 0, C.byref(n)), "get_table")
        if dt.subdtype:
            out = np.zeros((n.value,) + dt.subdtype[1], dt.subdtype[0])
        else:
            out = np.zeros(n.value, dt)
        if n.value:
            check(self.L.b3b200_get_table(self.h, which, ptr(out), n.value, C.byref(n)), "get_table")
        return out

    def tables(self):
        return {k: self.table(k) for k in self.TABLES}

    def solve_contacts_device(self, num_bodies, bodies_ptr, inertias_ptr, num_contacts, contacts_ptr, static0_index):
        """the stand-alone solver entry (b3GpuPgsContactSolver / b3GpuJacobiContactSolver::solveContacts) on caller-owned buffers"""
        check(self.L.b3b200_solve_contacts_device(self.h, int(num_bodies), C.c_void_p(int(bodies_ptr)), C.c_void_p(int(inertias_ptr)), int(num_contacts),
                                                  C.c_void_p(int(contacts_ptr)), int(static0_index)), "solve_contacts_device")

    def checkpoint_save(self, path):
        check(self.L.b3b200_checkpoint_save(self.h, str(path).encode()), "checkpoint_save")

    def checkpoint_load(self, path):
        check(self.L.b3b200_checkpoint_load(self.h, str(path).encode()), "checkpoint_load")

    def copy_transforms(self, dst_device_ptr, num_nodes):
        """copyTransformsToVBOKernel: (pos.xyz, 1) then orientations into a device buffer of 2 * num_nodes float4"""
        check(self.L.b3b200_copy_transforms(self.h, C.c_void_p(int(dst_device_ptr)), int(num_nodes)), "copy_transforms")

    def device_buffer(self, which):
        p = C.c_void_p()
        check(self.L.b3b200_device_buffer(self.h, int(which), C.byref(p)), "device_buffer")
        return p.value


class Broadphase:
    """b3GpuBroadphaseInterface handle (stand-alone, as PairBench uses it)."""

    def __init__(self, kind, max_proxies, max_pairs, device=0):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.b3b200_bp_create(int(kind), int(device), None, int(max_proxies), int(max_pairs), C.byref(h)), "bp_create")
        self.h = h

    def close(self):
        if self.h:
            self.L.b3b200_bp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create_proxy(self, mn, mx, user_ptr):
        check(self.L.b3b200_bp_create_proxy(self.h, f3(mn), f3(mx), int(user_ptr)), "bp_create_proxy")

    def create_large_proxy(self, mn, mx, user_ptr):
        check(self.L.b3b200_bp_create_large_proxy(self.h, f3(mn), f3(mx), int(user_ptr)), "bp_create_large_proxy")

    def write_aabbs(self):
        check(self.L.b3b200_bp_write_aabbs(self.h), "bp_write_aabbs")

    def set_aabbs(self, aabbs):
        aabbs = np.ascontiguousarray(aabbs, aabb_t)
        check(self.L.b3b200_bp_set_aabbs(self.h, ptr(aabbs), len(aabbs)), "bp_set_aabbs")

    def calculate_pairs(self, max_pairs):
        check(self.L.b3b200_bp_calculate_pairs(self.h, int(max_pairs)), "bp_calculate_pairs")

    def num_overlap(self):
        return check(self.L.b3b200_bp_num_overlap(self.h), "bp_num_overlap")

    def pairs(self):
        n = C.c_int(0)
        check(self.L.b3b200_bp_get_pairs(self.h, None, 0, C.byref(n)), "bp_get_pairs")
        out = np.zeros(n.value, int4_t)
        if n.value:
            check(self.L.b3b200_bp_get_pairs(self.h, ptr(out), n.value, C.byref(n)), "bp_get_pairs")
        return out

    def last_ms(self):
        ms = C.c_float(0)
        check(self.L.b3b200_bp_last_ms(self.h, C.byref(ms)), "bp_last_ms")
        return ms.value


def mpr_penetration(pairs, bodies, collidables, convex, vertices, sep_normals, has_sep_axis, capacity, count0=0, device=0):
    """mprPenetrationKernel on host arrays (b3b200_mpr_penetration).  Returns (pairs, sep_normals, has_sep_axis, contacts[:total], total, results)"""
    L = lib()
    pairs = np.ascontiguousarray(pairs.copy())
    sep = np.ascontiguousarray(sep_normals, np.float32).copy()
    has = np.ascontiguousarray(has_sep_axis, np.int32).copy()
    contacts = np.zeros(max(capacity, 1), contact4_t)
    n = C.c_int(int(count0))
    res = np.zeros(len(pairs), mpr_result_t)
    bodies, collidables, convex = np.ascontiguousarray(bodies), np.ascontiguousarray(collidables), np.ascontiguousarray(convex)
    vertices = np.ascontiguousarray(vertices, np.float32)
    check(L.b3b200_mpr_penetration(int(device), ptr(pairs), len(pairs), ptr(bodies), len(bodies), ptr(collidables), len(collidables), ptr(convex), len(convex),
                                   ptr(vertices), len(vertices), ptr(sep), ptr(has), ptr(contacts), int(capacity), C.byref(n), ptr(res)), "mpr_penetration")
    return pairs, sep, has, contacts[: min(n.value, capacity)], n.value, res
