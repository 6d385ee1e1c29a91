"""ctypes access to the CPU oracle (oracle/_build/liboracle.so) and to the compiled
reference (oracle/_ref/libb3ref.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from bullet3_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libb3ref.so")

_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"])
        _oracle = C.CDLL(ORACLE_SO)
    return _oracle


def ref_available():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


P = capi.ptr


def _arr(a, dt):
    return np.ascontiguousarray(a, dt)


class Shapes:
    """flat shape tables + bodies, as both oracle and reference shim consume them"""

    def __init__(self, tables):
        self.collidables = _arr(tables["collidables"], capi.collidable_t)
        self.local_aabbs = _arr(tables["local_aabbs"], capi.aabb_t)
        self.convex = _arr(tables["convex"], capi.convex_t)
        self.vertices = _arr(tables["vertices"], np.float32).reshape(-1, 4)
        self.unique_edges = _arr(tables["unique_edges"], np.float32).reshape(-1, 4)
        self.faces = _arr(tables["faces"], capi.face_t)
        self.indices = _arr(tables["indices"], np.int32)
        self.child_shapes = _arr(tables.get("child_shapes", np.zeros(0, capi.child_shape_t)), capi.child_shape_t)


def update_aabbs(lib, prefix, bodies, shapes):
    bodies = _arr(bodies, capi.rigid_body_t)
    out = np.zeros(len(bodies), capi.aabb_t)
    getattr(lib, prefix + "update_aabbs")(P(bodies), len(bodies), P(shapes.collidables), P(shapes.local_aabbs), P(out))
    return out


def brute_force_pairs(lib, prefix, aabbs, small_idx, large_idx, max_pairs, fn="brute_force_pairs"):
    aabbs = _arr(aabbs, capi.aabb_t)
    small_idx = _arr(small_idx, np.int32)
    large_idx = _arr(large_idx, np.int32)
    out = np.zeros(max(max_pairs, 1), capi.int4_t)
    f = getattr(lib, prefix + fn)
    n = f(P(aabbs), P(small_idx), len(small_idx), P(large_idx), len(large_idx), P(out), int(max_pairs))
    return n, out[: min(n, max_pairs)]


def integrate(lib, prefix, bodies, dt, damping, gravity):
    bodies = _arr(bodies, capi.rigid_body_t).copy()
    g = (C.c_float * 3)(*[float(x) for x in gravity])
    getattr(lib, prefix + "integrate")(P(bodies), len(bodies), C.c_float(dt), C.c_float(damping), g)
    return bodies


def convex_contacts_oracle(pairs, bodies, shapes, clip_min, clip_max, max_contacts):
    pairs = _arr(pairs, capi.int4_t)
    bodies = _arr(bodies, capi.rigid_body_t)
    out = np.zeros(max(max_contacts, 1), capi.contact4_t)
    pci = np.full(max(len(pairs), 1), -1, np.int32)
    n = oracle().orc_convex_contacts(P(pairs), len(pairs), P(bodies), P(shapes.collidables), P(shapes.convex), P(shapes.vertices),
                                     P(shapes.unique_edges), P(shapes.faces), P(shapes.indices), C.c_float(clip_min), C.c_float(clip_max),
                                     P(out), int(max_contacts), P(pci))
    return out[:n], pci[: len(pairs)]


def convex_contacts_ref(pairs, bodies, shapes, max_contacts):
    pairs = _arr(pairs, capi.int4_t)
    bodies = _arr(bodies, capi.rigid_body_t)
    out = np.zeros(max(max_contacts, 1), capi.contact4_t)
    pci = np.full(max(len(pairs), 1), -1, np.int32)
    n = ref().ref_convex_contacts(P(pairs), len(pairs), P(bodies), len(bodies), P(shapes.collidables), len(shapes.collidables),
                                  P(shapes.convex), len(shapes.convex), P(shapes.vertices), len(shapes.vertices),
                                  P(shapes.unique_edges), len(shapes.unique_edges), P(shapes.faces), len(shapes.faces),
                                  P(shapes.indices), len(shapes.indices), P(out), int(max_contacts), P(pci))
    return out[:n], pci[: len(pairs)]


def colour_contacts(contacts, num_bodies, static_idx):
    contacts = _arr(contacts, capi.contact4_t)
    colours = np.zeros(max(len(contacts), 1), np.int32)
    nb = oracle().orc_colour_contacts(P(contacts), len(contacts), int(num_bodies), int(static_idx), P(colours))
    return nb, colours[: len(contacts)]


def build_constraints(lib, prefix, contacts, bodies, inertias, dt=1.0 / 60.0, drift=0.005, coeff=0.2):
    contacts = _arr(contacts, capi.contact4_t)
    bodies = _arr(bodies, capi.rigid_body_t)
    inertias = _arr(inertias, capi.inertia_t)
    out = np.zeros(max(len(contacts), 1), capi.constraint4_t)
    getattr(lib, prefix + "build_constraints")(P(contacts), len(contacts), P(bodies), P(inertias), C.c_float(dt), C.c_float(drift),
                                               C.c_float(coeff), P(out))
    return out[: len(contacts)]


def solve(constraints, batch_offsets, bodies, inertias, iterations):
    cs = _arr(constraints, capi.constraint4_t).copy()
    bodies = _arr(bodies, capi.rigid_body_t).copy()
    inertias = _arr(inertias, capi.inertia_t)
    off = _arr(batch_offsets, np.int32)
    oracle().orc_solve(P(cs), P(off), len(off) - 1, P(bodies), P(inertias), int(iterations))
    return bodies, cs


def oracle_pgs_step_velocities(contacts, bodies, inertias, static_idx, iterations, colours=None):
    """colour -> sort by batch -> build rows -> solve, all on the CPU oracle (colours: take this batch assignment instead)"""
    contacts = _arr(contacts, capi.contact4_t).copy()
    if colours is None:
        nb, colours = colour_contacts(contacts, len(bodies), static_idx)
    else:
        colours = np.asarray(colours, np.int32)
        nb = int(colours.max()) + 1 if len(colours) else 0
    contacts["batchIdx"] = colours
    order = np.argsort(colours, kind="stable")
    sorted_contacts = contacts[order]
    counts = np.bincount(colours, minlength=nb)[:nb] if len(colours) else np.zeros(0, np.int64)
    off = np.zeros(nb + 1, np.int32)
    off[1:] = np.cumsum(counts)
    cs = build_constraints(oracle(), "orc_", sorted_contacts, bodies, inertias)
    out_bodies, cs = solve(cs, off, bodies, inertias, iterations)
    return out_bodies, cs, off, colours


def sorted_pair_set(pairs):
    """normalise (min,max), sort lexicographically -> (n,2) int array"""
    if len(pairs) == 0:
        return np.zeros((0, 2), np.int32)
    x = np.minimum(pairs["x"], pairs["y"])
    y = np.maximum(pairs["x"], pairs["y"])
    xy = np.stack([x, y], axis=1)
    order = np.lexsort((xy[:, 1], xy[:, 0]))
    return xy[order]


def jacobi_solve(contacts, bodies, inertias, static_idx, iterations, dt=1.0 / 60.0, drift=0.005, coeff=0.99, host_order=False):
    """host_order=False: the GPU path's order (contacts, average, friction, average per iteration); True: the order of the reference's
    host twin solveGroupHost (all contact iterations, then all friction iterations) -- same code, one switch"""
    contacts = _arr(contacts, capi.contact4_t)
    bodies = _arr(bodies, capi.rigid_body_t).copy()
    inertias = _arr(inertias, capi.inertia_t)
    oracle().orc_jacobi_solve_ordered(P(contacts), len(contacts), P(bodies), len(bodies), P(inertias), int(static_idx), int(iterations),
                                      C.c_float(dt), C.c_float(drift), C.c_float(coeff), int(bool(host_order)))
    return bodies


def refcl_jacobi_solve_host(contacts, bodies, inertias, static_idx, iterations, dt=1.0 / 60.0, drift=0.005, coeff=0.99):
    """b3GpuJacobiContactSolver::solveGroupHost of the compiled reference"""
    contacts = _arr(contacts, capi.contact4_t).copy()
    bodies = _arr(bodies, capi.rigid_body_t).copy()
    inertias = _arr(inertias, capi.inertia_t).copy()
    refcl().refcl_jacobi_solve_host(P(contacts), len(contacts), P(bodies), len(bodies), P(inertias), int(static_idx), int(iterations),
                                    C.c_float(dt), C.c_float(drift), C.c_float(coeff))
    return bodies


# ---------------------------------------------------------------- reference host twins of src/Bullet3OpenCL (fake OpenCL build)
REFCL_SO = os.path.join(ROOT, "oracle", "_ref", "libb3refcl.so")
_refcl = None


def refcl_available():
    return os.path.exists(REFCL_SO)


def refcl():
    global _refcl
    if _refcl is None:
        _refcl = C.CDLL(REFCL_SO, mode=C.RTLD_GLOBAL)
        _refcl.refcl_np_create.restype = C.c_void_p
    return _refcl


def refcl_pairs_host(aabbs, large_idx, max_pairs):
    aabbs = _arr(aabbs, capi.aabb_t)
    is_large = np.zeros(len(aabbs), np.uint8)
    is_large[np.asarray(large_idx, np.int64)] = 1
    out = np.zeros(max(max_pairs, 1), capi.int4_t)
    n = refcl().refcl_pairs_host(P(aabbs), len(aabbs), P(is_large), P(out), int(max_pairs))
    return n, out[: min(n, max_pairs)]


def refcl_pgs_solve(sorted_contacts, batch_sizes, bodies, inertias, iterations, dt=1.0 / 60.0):
    contacts = _arr(sorted_contacts, capi.contact4_t)
    bs = _arr(batch_sizes, np.int32)
    bodies = _arr(bodies, capi.rigid_body_t).copy()
    inertias = _arr(inertias, capi.inertia_t)
    cs = np.zeros(max(len(contacts), 1), capi.constraint4_t)
    rc = refcl().refcl_pgs_solve(P(contacts), len(contacts), P(bs), len(bs), P(bodies), len(bodies), P(inertias), int(iterations), C.c_float(dt), P(cs))
    assert rc == 0
    return bodies, cs[: len(contacts)]


def contacts_oracle(pairs, bodies, shapes, clip_min, clip_max, max_contacts):
    """general contact loop: convex, compound children, planes"""
    pairs = _arr(pairs, capi.int4_t)
    bodies = _arr(bodies, capi.rigid_body_t)
    out = np.zeros(max(max_contacts, 1), capi.contact4_t)
    ch = shapes.child_shapes if len(shapes.child_shapes) else np.zeros(1, capi.child_shape_t)
    n = oracle().orc_contacts(P(pairs), len(pairs), P(bodies), P(shapes.collidables), P(shapes.convex), P(shapes.vertices), P(shapes.unique_edges),
                              P(shapes.faces), P(shapes.indices), P(ch), C.c_float(clip_min), C.c_float(clip_max), P(out), int(max_contacts))
    return out[:n]


class RefNarrowphase:
    """the reference b3GpuNarrowPhase (CHECK_ON_HOST build over the fake OpenCL)"""

    def __init__(self, cfg):
        self.L = refcl()
        self.h = C.c_void_p(self.L.refcl_np_create(P(cfg)))

    def close(self):
        if self.h:
            self.L.refcl_np_destroy(self.h)
            self.h = None

    def register_convex_points(self, pts, scaling=(1.0, 1.0, 1.0)):
        pts = _arr(pts, np.float32).reshape(-1, 3)
        sc = (C.c_float * 4)(*[float(x) for x in scaling], 1.0)
        return self.L.refcl_np_register_convex_points(self.h, P(pts), len(pts), sc)

    def register_plane(self, normal, constant):
        n = (C.c_float * 3)(*[float(x) for x in normal])
        return self.L.refcl_np_register_plane(self.h, n, C.c_float(constant))

    def register_compound(self, children):
        children = _arr(children, capi.child_shape_t)
        return self.L.refcl_np_register_compound(self.h, P(children), len(children))

    def register_body(self, collidable, mass, pos, orn, aabb_min, aabb_max):
        f = lambda v, n: (C.c_float * n)(*[float(x) for x in list(v)[:n]])
        return self.L.refcl_np_register_body(self.h, int(collidable), C.c_float(mass), f(list(pos) + [0], 4), f(orn, 4), f(aabb_min, 3), f(aabb_max, 3))

    def table(self, which, dt):
        n = C.c_int(0)
        self.L.refcl_np_get_table(self.h, which, None, 0, C.byref(n))
        out = np.zeros(n.value, dt)
        if n.value:
            self.L.refcl_np_get_table(self.h, which, P(out), n.value, C.byref(n))
        return out

    def compute_contacts(self, bodies, pairs, aabbs, max_contacts):
        bodies = _arr(bodies, capi.rigid_body_t)
        pairs = _arr(pairs, capi.int4_t)
        aabbs = _arr(aabbs, capi.aabb_t)
        out = np.zeros(max(max_contacts, 1), capi.contact4_t)
        n = self.L.refcl_np_compute_contacts(self.h, P(bodies), len(bodies), P(pairs), len(pairs), P(aabbs), P(out), int(max_contacts), None)
        return out[:n]


def concave_contacts_oracle(pairs, bodies, shapes, aabbs, max_contacts):
    """trimesh x convex / compound-child contacts (host twins of the concave leg); returns (contacts, num_candidates)"""
    pairs = _arr(pairs, capi.int4_t)
    bodies = _arr(bodies, capi.rigid_body_t)
    aabbs = _arr(aabbs, capi.aabb_t)
    out = np.zeros(max(max_contacts, 1), capi.contact4_t)
    ch = shapes.child_shapes if len(shapes.child_shapes) else np.zeros(1, capi.child_shape_t)
    ncand = C.c_int(0)
    n = oracle().orc_concave_contacts(P(pairs), len(pairs), P(bodies), P(shapes.collidables), P(shapes.convex), P(shapes.vertices), P(shapes.unique_edges),
                                      P(shapes.faces), P(shapes.indices), P(ch), P(aabbs), P(out), int(max_contacts), C.byref(ncand))
    return out[:n], ncand.value


def _ref_register_concave(self, vertices, tri_indices, scaling=(1.0, 1.0, 1.0)):
    v = _arr(vertices, np.float32).reshape(-1, 3)
    i = _arr(tri_indices, np.int32).reshape(-1)
    sc = (C.c_float * 4)(*[float(x) for x in scaling], 1.0)
    return self.L.refcl_np_register_concave(self.h, P(v), len(v), P(i), len(i), sc)


def _ref_concave_host_twins(self, on=True):
    self.L.refcl_concave_host_twins(int(bool(on)))


RefNarrowphase.register_concave = _ref_register_concave
RefNarrowphase.concave_host_twins = _ref_concave_host_twins


def _ref_register_sphere(self, radius):
    return self.L.refcl_np_register_sphere(self.h, C.c_float(radius))


RefNarrowphase.register_sphere = _ref_register_sphere


def solve_joints_oracle(bodies, inertias, joints):
    """b3GpuPgsConstraintSolver::solveJoints restated (orc_solve_joints); returns (bodies, joints) after the solve"""
    b = _arr(bodies, capi.rigid_body_t).copy()
    j = _arr(joints, capi.joint_t).copy()
    inert = _arr(inertias, capi.inertia_t)
    oracle().orc_solve_joints(P(b), len(b), P(inert), P(j), len(j))
    return b, j


def solve_joints_ref(bodies, inertias, joints):
    """the reference's CPU joint path (b3PgsJacobiSolver + b3Point2PointConstraint, sequential in index order)"""
    b = _arr(bodies, capi.rigid_body_t).copy()
    j = _arr(joints, capi.joint_t)
    inert = _arr(inertias, capi.inertia_t).copy()
    ref().ref_solve_joints_p2p(P(b), P(inert), len(b), P(j), len(j))
    return b


def make_rays(ray_from, ray_to, max_fraction=1.0):
    f = np.asarray(ray_from, np.float32).reshape(-1, 3)
    t = np.asarray(ray_to, np.float32).reshape(-1, 3)
    rays = np.zeros(len(f), capi.ray_info_t)
    rays["from"][:, :3] = f
    rays["to"][:, :3] = t
    hits = np.zeros(len(f), capi.ray_hit_t)
    hits["hitFraction"] = max_fraction
    hits["hitBody"] = -1
    return rays, hits


def cast_rays_oracle(ray_from, ray_to, bodies, shapes, max_fraction=1.0):
    """b3GpuRaycast::castRaysHost restated (orc_cast_rays): closest hit over all bodies, first body wins ties"""
    rays, hits = make_rays(ray_from, ray_to, max_fraction)
    b = _arr(bodies, capi.rigid_body_t)
    oracle().orc_cast_rays(P(rays), len(rays), P(hits), P(b), len(b), P(shapes.collidables), P(shapes.convex), P(shapes.faces))
    return hits


def _ref_cast_rays(self, ray_from, ray_to, bodies, max_fraction=1.0):
    """the reference's own b3GpuRaycast::castRaysHost on this narrowphase's shapes"""
    rays, hits = make_rays(ray_from, ray_to, max_fraction)
    b = _arr(bodies, capi.rigid_body_t)
    self.L.refcl_cast_rays_host(self.h, P(b), len(b), P(rays), len(rays), P(hits))
    return hits


RefNarrowphase.cast_rays = _ref_cast_rays


class RefPipeline(RefNarrowphase):
    """the reference's own b3GpuRigidBodyPipeline (unmodified, every host-twin switch on, fake OpenCL) behind the method names
    bullet3_b200.scenes uses, so that one scene recipe builds both worlds.  bench.py --impl reference times its step()."""

    def __init__(self, cfg):
        self.L = refcl()
        self.L.refcl_pipeline_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.refcl_pipeline_create(P(cfg)))
        self.num_bodies = 0

    def close(self):
        if self.h:
            self.L.refcl_pipeline_destroy(self.h)
            self.h = None

    def register_instance(self, mass, pos, orn, collidable, user_index=0):
        f = lambda v, n: (C.c_float * n)(*[float(x) for x in list(v)[:n]])
        r = self.L.refcl_pipeline_register_instance(self.h, C.c_float(mass), f(list(pos) + [0.0], 4), f(orn, 4), int(collidable), int(user_index))
        self.num_bodies += 1
        return r

    def register_instances(self, masses, positions4, orientations4, collidables):
        positions4 = _arr(positions4, np.float32).reshape(-1, 4)
        orientations4 = _arr(orientations4, np.float32).reshape(-1, 4)
        first = self.num_bodies
        for i in range(len(masses)):
            self.L.refcl_pipeline_register_instance(self.h, C.c_float(float(masses[i])), P(positions4[i]), P(orientations4[i]), int(collidables[i]), 0)
        self.num_bodies += len(masses)
        return first

    def upload(self):
        self.L.refcl_pipeline_upload(self.h)

    def set_bodies(self, bodies):
        bodies = _arr(bodies, capi.rigid_body_t)
        self.L.refcl_pipeline_set_bodies(self.h, P(bodies), len(bodies))

    def bodies(self):
        out = np.zeros(self.num_bodies, capi.rigid_body_t)
        self.L.refcl_pipeline_get_bodies(self.h, P(out), len(out))
        return out

    def step(self, dt, steps=1):
        out = (C.c_int * 3)()
        self.L.refcl_pipeline_step(self.h, C.c_float(dt), int(steps), 4, out)
        return list(out)

    def profile_zones(self):
        """seconds per B3_PROFILE zone of the reference (inclusive) since the last call"""
        buf = C.create_string_buffer(1 << 16)
        self.L.refcl_profile_zones(buf, len(buf))
        return {k: float(v) for k, v in (kv.split("=") for kv in buf.value.decode().split(";") if "=" in kv)}


class RefCpuPipeline:
    """the reference's b3CpuRigidBodyPipeline + b3CpuNarrowPhase + b3DynamicBvhBroadphase, instantiated as they are
    (BASELINE configs[0]); per-stage access for the per-step parity test of SURVEY 8(d) config 1"""

    def __init__(self, cfg):
        self.L = ref()
        self.L.ref_cpu_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.ref_cpu_create(P(cfg)))

    def close(self):
        if self.h:
            self.L.ref_cpu_destroy(self.h)
            self.h = None

    def register_convex_points(self, pts, scaling=(1.0, 1.0, 1.0)):
        pts = _arr(pts, np.float32).reshape(-1, 3)
        sc = (C.c_float * 4)(*[float(x) for x in scaling], 1.0)
        return self.L.ref_cpu_register_convex_points(self.h, P(pts), len(pts), sc)

    def register_instance(self, mass, pos, orn, collidable, user_index=0):
        f = lambda v, n: (C.c_float * n)(*[float(x) for x in list(v)[:n]])
        return self.L.ref_cpu_register_instance(self.h, C.c_float(mass), f(list(pos) + [0.0], 4), f(orn, 4), int(collidable), int(user_index))

    @property
    def num_bodies(self):
        return self.L.ref_cpu_num_bodies(self.h)

    def step(self, dt, steps=1):
        self.L.ref_cpu_step(self.h, C.c_float(dt), int(steps))

    def stage(self, which, dt=0.0):
        self.L.ref_cpu_stage(self.h, int(which), C.c_float(dt))

    def bodies(self):
        out = np.zeros(self.num_bodies, capi.rigid_body_t)
        self.L.ref_cpu_get_bodies(self.h, P(out), len(out))
        return out

    def set_bodies(self, bodies):
        bodies = _arr(bodies, capi.rigid_body_t)
        self.L.ref_cpu_set_bodies(self.h, P(bodies), len(bodies))

    def aabbs(self):
        out = np.zeros(self.num_bodies, capi.aabb_t)
        self.L.ref_cpu_get_aabbs(self.h, P(out), len(out))
        return out

    def pairs(self):
        n = self.L.ref_cpu_get_pairs(self.h, None, 0)
        out = np.zeros(max(n, 1), capi.int4_t)
        self.L.ref_cpu_get_pairs(self.h, P(out), n)
        return out[:n]

    def contacts(self):
        n = self.L.ref_cpu_get_contacts(self.h, None, 0)
        out = np.zeros(max(n, 1), capi.contact4_t)
        self.L.ref_cpu_get_contacts(self.h, P(out), n)
        return out[:n]


def _refcpu_table(self, which, dt):
    n = C.c_int(0)
    self.L.ref_cpu_get_table(self.h, which, None, 0, C.byref(n))
    out = np.zeros(n.value, dt)
    if n.value:
        self.L.ref_cpu_get_table(self.h, which, P(out), n.value, C.byref(n))
    return out


def _refcpu_register_shapes_into(self, world):
    """register the reference's own hull tables (b3ConvexUtility output) in a b3b200 world through b3b200_register_convex --
    the registerConvexHullShape(b3ConvexUtility*) route of the drop-in boundary -- so both sides work on identical shape data"""
    col = _refcpu_table(self, 0, capi.collidable_t)
    convex = _refcpu_table(self, 2, capi.convex_t)
    verts = _refcpu_table(self, 3, np.dtype(("f4", 4)))
    edges = _refcpu_table(self, 4, np.dtype(("f4", 4)))
    faces = _refcpu_table(self, 5, capi.face_t)
    idx = _refcpu_table(self, 6, np.dtype("i4"))
    out = []
    for c in col:
        cv = convex[int(c["shapeIndex"])]
        f = faces[cv["faceOffset"]: cv["faceOffset"] + cv["numFaces"]].copy()
        lo = int(f["indexOffset"].min())
        hi = int((f["indexOffset"] + f["numIndices"]).max())
        f["indexOffset"] -= lo
        out.append(world.register_convex(verts[cv["vertexOffset"]: cv["vertexOffset"] + cv["numVertices"]], f, idx[lo:hi],
                                         edges[cv["uniqueEdgesOffset"]: cv["uniqueEdgesOffset"] + cv["numUniqueEdges"]], np.array([cv])))
    return out


RefCpuPipeline.table = _refcpu_table
RefCpuPipeline.register_shapes_into = _refcpu_register_shapes_into
