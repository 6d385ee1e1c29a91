"""Headless restatements of the reference's demo scene recipes (BASELINE.json configs).

Only the *recipes* (positions, shapes, counts) are restated; bodies are created
through the C ABI exactly like examples/OpenCL/rigidbody/GpuConvexScene.cpp does
through b3GpuNarrowPhase / b3GpuRigidBodyPipeline.
"""
import numpy as np

IDENT = (0.0, 0.0, 0.0, 1.0)


def box_points(hx, hy=None, hz=None):
    hy = hx if hy is None else hy
    hz = hx if hz is None else hz
    return np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32)


def tetra_points(scale=1.0):
    # tetra_vertices of examples/OpenGLWindow/ShapeData.h:1034 (unit tetrahedron around the origin)
    return np.array([[0.0, 1.0, 0.0], [1.0, -1.0, 1.0], [-1.0, -1.0, 1.0], [0.0, -1.0, -1.0]], np.float32) * np.float32(scale)


def random_hull_points(rng, n, radius_lo=0.5, radius_hi=1.5):
    """n points on a sphere of random radius (seeded hulls)."""
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = rng.uniform(radius_lo, radius_hi)
    return (v * r).astype(np.float32)


def random_quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(np.float32(q))


def add_ground_box(world, half=400.0, y_top=0.0):
    """static box like GpuConvexScene::createStaticEnvironment (GpuConvexScene.cpp:280-304)"""
    col = world.register_convex_points(box_points(half))
    return world.register_instance(0.0, (0.0, y_top - half, 0.0), IDENT, col)


def box_stack(world, nx, ny, nz, half=0.5, spacing=1.0, y0=0.5, mass=1.0, ground=True):
    """config 1: nx*ny*nz unit boxes at (i, 0.5+j, k) on a static 400-box"""
    if ground:
        add_ground_box(world)
    col = world.register_convex_points(box_points(half))
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                world.register_instance(mass, (i * spacing, y0 + j * spacing, k * spacing), IDENT, col)
    return col


def batched_box_worlds(world, num_worlds, nx=8, ny=4, nz=8):
    """BASELINE configs[4](i) / SURVEY 8(d) config 5(i): `num_worlds` independent worlds, each an nx x ny x nz pile of cubes
    (GpuBoxPlaneScene recipe) on its own static ground box, all at the SAME coordinates (b3b200_set_current_world)"""
    ground = world.register_convex_points(box_points(400.0))
    col = world.register_convex_points(box_points(1.0))
    n = nx * ny * nz + 1
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    pos = np.zeros((n, 3), np.float32)
    pos[0] = (0.0, -400.0, 0.0)
    pos[1:, 0] = ((j + 1) & 1) + 2.2 * i
    pos[1:, 1] = 1.0 + 2.0 * j
    pos[1:, 2] = ((j + 1) & 1) + 2.2 * k
    masses = np.ones(n, np.float32)
    masses[0] = 0.0
    quats = np.tile(np.asarray(IDENT, np.float32), (n, 1))
    cols = np.full(n, col, np.int32)
    cols[0] = ground
    for wd in range(num_worlds):
        world.set_current_world(wd)
        world.register_instances(masses, pos, quats, cols)
    return n


def box_plane_scene(world, nx, ny, nz, ground=True):
    """config 3: GpuBoxPlaneScene recipe (GpuConvexScene.cpp:163-278): cubes of half-extent 1,
    pos = (((j+1)&1) + 2.2 i, 1 + 2 j, ((j+1)&1) + 2.2 k)"""
    if ground:
        add_ground_box(world)
    col = world.register_convex_points(box_points(1.0))
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                world.register_instance(1.0, (((j + 1) & 1) + 2.2 * i, 1.0 + 2.0 * j, ((j + 1) & 1) + 2.2 * k), IDENT, col)
    return col


def mixed_convex_scene(world, nx, ny, nz, seed=1234, num_hull_shapes=8, rotate=True, ground=True, hull_verts=(8, 16)):
    """config 4 (convex part): mix of tetrahedra, boxes and seeded random hulls dropped on a
    static ground.  Shapes are instanced (a handful of collidables), bodies get seeded
    random orientations.  Layout follows the config-3 grid with 2.8 spacing so that the
    larger hulls start just short of touching."""
    rng = np.random.default_rng(seed)
    if ground:
        add_ground_box(world)
    shapes = [world.register_convex_points(box_points(1.0)), world.register_convex_points(tetra_points(1.0))]
    for _ in range(num_hull_shapes):
        n = int(rng.integers(hull_verts[0], hull_verts[1] + 1))
        shapes.append(world.register_convex_points(random_hull_points(rng, n, 0.8, 1.3)))
    kinds = rng.integers(0, len(shapes), size=nx * ny * nz)
    t = 0
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                q = random_quat(rng) if rotate else IDENT
                world.register_instance(1.0, (2.8 * i, 1.5 + 2.8 * j, 2.8 * k), q, shapes[int(kinds[t])])
                t += 1
    return shapes


def bench_convex_scene(world, nx, ny, nz, seed=1234, num_hull_shapes=8, spacing=(2.2, 2.0, 2.2)):
    """BASELINE.json config 4, convex part, at bench size (64x64x64 = 262 144 bodies): an equal mix of
    boxes, tetrahedra and `num_hull_shapes` seeded random hulls (8-16 vertices), seeded random
    orientations, packed on the config-3 lattice so that neighbours touch from the first step,
    resting on the static 400-box.  Built with one bulk registration call."""
    rng = np.random.default_rng(seed)
    add_ground_box(world)
    shapes = [world.register_convex_points(box_points(0.8)), world.register_convex_points(tetra_points(0.9))]
    for _ in range(num_hull_shapes):
        n = int(rng.integers(8, 17))
        shapes.append(world.register_convex_points(random_hull_points(rng, n, 0.8, 1.1)))
    n = nx * ny * nz
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.zeros((n, 4), np.float32)
    pos[:, 0] = (((j + 1) & 1) * 0.5 + spacing[0] * i).reshape(-1)
    pos[:, 1] = (1.2 + spacing[1] * j).reshape(-1)
    pos[:, 2] = (((j + 1) & 1) * 0.5 + spacing[2] * k).reshape(-1)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    # kind: 1/3 boxes, 1/3 tetrahedra, 1/3 hulls
    kind = rng.integers(0, 3, n)
    hull = rng.integers(0, num_hull_shapes, n) + 2
    col = np.where(kind == 2, hull, kind)
    col = np.asarray(shapes, np.int32)[col]
    world.register_instances(np.ones(n, np.float32), pos, q.astype(np.float32), col)
    return shapes


def compound_children(child_collidable, offsets, orientations=None):
    """b3GpuChildShape records for registerCompoundShape (children must be convex-hull collidables)"""
    from . import capi

    ch = np.zeros(len(offsets), capi.child_shape_t)
    for i, o in enumerate(offsets):
        ch["childPosition"][i, :3] = o
        ch["childOrientation"][i] = IDENT if orientations is None else orientations[i]
        ch["shapeIndex"][i] = child_collidable
        ch["shapeType"][i] = capi.SHAPE_CONVEX_HULL
    return ch


L_OFFSETS = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)]  # 3-box "L" 


def heightfield_mesh(nx, nz, cell=1.0, amplitude=2.0, freq=0.1, x0=None, z0=None):
    """config 4's synthetic concave ground: an nx x nz quad grid, two triangles per quad,
    h = amplitude * sin(freq x) * cos(freq z).  Returns (vertices (V,3) f32, triangle indices (T*3,) i32)."""
    x0 = -0.5 * nx * cell if x0 is None else x0
    z0 = -0.5 * nz * cell if z0 is None else z0
    xs = x0 + cell * np.arange(nx + 1, dtype=np.float64)
    zs = z0 + cell * np.arange(nz + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = amplitude * np.sin(freq * X) * np.cos(freq * Z)
    verts = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    i, k = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    v00 = (i * (nz + 1) + k).reshape(-1)
    v01 = v00 + 1
    v10 = v00 + (nz + 1)
    v11 = v10 + 1
    # counter-clockwise seen from +y: normals point up
    tris = np.stack([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)], 1).reshape(-1, 3)
    return verts, tris.astype(np.int32).reshape(-1)


def bench_config4_scene(world, nx, ny, nz, seed=1234, num_hull_shapes=8, hull_verts=(8, 32), spacing=(2.2, 2.0, 2.2), mesh_quads=None, keep=None, poses=None):
    """BASELINE.json config 4: a pile of nx*ny*nz bodies -- 1/4 boxes, 1/4 tetrahedra, 1/4 seeded random hulls
    (`hull_verts` vertices on a sphere), 1/4 three-box "L" compounds -- with seeded random orientations on the config-3
    lattice, dropped on a synthetic concave heightfield trimesh h = 2 sin(0.1 x) cos(0.1 z) that extends 10 % beyond the
    pile (256 x 256 quads = 131 072 triangles at bench size).  Body 0 is the mesh (it has to be the lower body index of
    its pairs, b3BvhTraversal.h:35).  Returns the list of collidables.
    keep: optional index array into the nx*ny*nz lattice bodies -- only those are registered (a sub-region of the scene with the
    same shapes and the same mesh); poses: optional (positions n x 4, orientations n x 4) replacing the lattice poses (e.g. a
    settled state), indexed like the full lattice."""
    rng = np.random.default_rng(seed)
    wx, wz = spacing[0] * nx, spacing[2] * nz
    if mesh_quads is None:
        mesh_quads = int(min(256, max(8, 2 * max(nx, nz))))
    cell = 1.2 * max(wx, wz) / mesh_quads
    verts, tris = heightfield_mesh(mesh_quads, mesh_quads, cell=cell, amplitude=2.0, freq=0.1, x0=0.5 * wx - 0.5 * mesh_quads * cell,
                                   z0=0.5 * wz - 0.5 * mesh_quads * cell)
    mesh = world.register_concave(verts, tris)
    world.register_instance(0.0, (0.0, 0.0, 0.0), IDENT, mesh)
    box = world.register_convex_points(box_points(0.8))
    tet = world.register_convex_points(tetra_points(0.9))
    hulls = []
    for _ in range(num_hull_shapes):
        n = int(rng.integers(hull_verts[0], hull_verts[1] + 1))
        hulls.append(world.register_convex_points(random_hull_points(rng, n, 0.8, 1.1)))
    small_box = world.register_convex_points(box_points(0.4))
    ell = world.register_compound(compound_children(small_box, [(-0.4, -0.4, 0.0), (0.4, -0.4, 0.0), (-0.4, 0.4, 0.0)]))
    n = nx * ny * nz
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.zeros((n, 4), np.float32)
    pos[:, 0] = (((j + 1) & 1) * 0.5 + spacing[0] * i).reshape(-1)
    pos[:, 1] = (3.4 + spacing[1] * j).reshape(-1)
    pos[:, 2] = (((j + 1) & 1) * 0.5 + spacing[2] * k).reshape(-1)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    kind = rng.integers(0, 4, n)
    hull = np.asarray(hulls, np.int32)[rng.integers(0, num_hull_shapes, n)]
    col = np.where(kind == 0, box, np.where(kind == 1, tet, np.where(kind == 2, hull, ell))).astype(np.int32)
    q = q.astype(np.float32)
    if poses is not None:
        pos, q = np.ascontiguousarray(poses[0], np.float32).reshape(n, 4), np.ascontiguousarray(poses[1], np.float32).reshape(n, 4)
    if keep is not None:
        keep = np.asarray(keep, np.int64)
        pos, q, col = np.ascontiguousarray(pos[keep]), np.ascontiguousarray(q[keep]), np.ascontiguousarray(col[keep])
    world.register_instances(np.ones(len(col), np.float32), pos, q, col)
    return [mesh, box, tet] + hulls + [small_box, ell]
