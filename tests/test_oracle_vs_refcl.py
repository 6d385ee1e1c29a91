"""Pin the oracle against the reference's own HOST TWINS inside src/Bullet3OpenCL, compiled
unmodified against a host-memory fake OpenCL (oracle/ref_cl/, `make -C oracle refcl`):
calculateOverlappingPairsHost, convertToConstraints + solveContactConstraintHost, and the
executeHost twins of the parallel primitives.  Runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi
from test_oracle_vs_ref import all_pairs, make_world

pytestmark = pytest.mark.skipif(not oa.refcl_available(), reason="oracle/_ref/libb3refcl.so not built")


@pytest.mark.parametrize("seed", [0, 2])
def test_pairs_match_calculateOverlappingPairsHost(seed):
    w, sh, bodies, _ = make_world(seed=seed)
    aabbs, small, large, pairs = all_pairs(bodies, sh)
    n, ref_pairs = oa.refcl_pairs_host(aabbs, large, 1 << 20)
    assert n == len(pairs) and n > 100
    # same pairs in the same order (small x small first, then small x large)
    assert np.array_equal(pairs["x"], ref_pairs["x"]) and np.array_equal(pairs["y"], ref_pairs["y"])


@pytest.mark.parametrize("seed,iters", [(0, 4), (1, 10)])
def test_pgs_rows_and_solve_match_b3Solver_host(seed, iters):
    w, sh, bodies, inertias = make_world(seed=seed, n_side=6)
    _, _, _, pairs = all_pairs(bodies, sh)
    contacts, _ = oa.convex_contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 16)
    nb, colours = oa.colour_contacts(contacts, len(bodies), 0)
    contacts["batchIdx"] = colours
    order = np.argsort(colours, kind="stable")
    sorted_contacts = contacts[order]
    sizes = np.bincount(colours, minlength=nb).astype(np.int32)
    ref_bodies, ref_rows = oa.refcl_pgs_solve(sorted_contacts, sizes, bodies, inertias, iters)
    o_bodies, o_rows, off, _ = oa.oracle_pgs_step_velocities(contacts, bodies, inertias, 0, iters)
    assert len(contacts) > 150 and nb >= 4
    # constraint rows from b3Solver::convertToConstraints (host branch)
    pre = oa.build_constraints(oa.oracle(), "orc_", sorted_contacts, bodies, inertias)
    for f in ("linear", "worldPos", "center", "jacCoeffInv", "fJacCoeffInv"):
        assert np.array_equal(pre[f].view(np.uint32), ref_rows[f].view(np.uint32)), f
    # velocities after b3Solver::solveContactConstraintHost: bit for bit
    for f in ("linVel", "angVel"):
        assert np.array_equal(o_bodies[f][:, :3].view(np.uint32), ref_bodies[f][:, :3].view(np.uint32)), f
    assert np.abs(o_bodies["linVel"][:, :3] - bodies["linVel"][:, :3]).max() > 0.1


@pytest.mark.parametrize("n", [1, 255, 4096, 100003])
def test_primitives_match_executeHost(n):
    rng = np.random.default_rng(n)
    d = np.zeros(n, capi.sort_data_t)
    d["key"] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    d["key"][: n // 2] &= 0x3F
    d["value"] = np.arange(n, dtype=np.uint32)
    a, b = d.copy(), d.copy()
    oa.oracle().orc_radix_sort_kv(capi.ptr(a), n)
    oa.refcl().refcl_radix_sort(capi.ptr(b), n)
    assert np.array_equal(a["key"], b["key"]) and np.array_equal(a["value"], b["value"])
    src = rng.integers(0, 100, n).astype(np.uint32)
    x, y = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    oa.oracle().orc_prefix_scan(capi.ptr(src), capi.ptr(x), n, None)
    oa.refcl().refcl_prefix_scan(capi.ptr(src), capi.ptr(y), n, None)
    assert np.array_equal(x, y)
    buckets = 256
    s = np.zeros(n, capi.sort_data_t)
    s["key"] = np.sort(rng.integers(0, buckets, n).astype(np.uint32))
    c1, c2 = np.zeros(buckets, np.uint32), np.zeros(buckets, np.uint32)
    oa.oracle().orc_bound_search_count(capi.ptr(s), n, capi.ptr(c1), buckets)
    oa.refcl().refcl_bound_search_count(capi.ptr(s), n, capi.ptr(c2), buckets)
    assert np.array_equal(c1, c2)


@pytest.mark.parametrize("seed,iters", [(0, 7), (3, 8)])
def test_jacobi_matches_solveGroupHost(seed, iters):
    """pins the oracle's mass-splitting Jacobi solver: run in the loop order of the reference's host twin
    (b3GpuJacobiContactSolver::solveGroupHost, b3GpuJacobiContactSolver.cpp:462-697) it must equal that twin; the GPU path's
    variant (what the CUDA kernels are compared with, `host_order=False`) is the same code with the kernels' loop nest and the
    kernels' form of the friction step (solverUtils.cl:654-790 re-evaluates velocity + delta per tangent direction)"""
    w, sh, bodies, inertias = make_world(seed=seed, n_side=6)
    _, _, _, pairs = all_pairs(bodies, sh)
    contacts, _ = oa.convex_contacts_oracle(pairs, bodies, sh, -1e30, 0.02, 1 << 16)
    assert len(contacts) > 150
    ref_bodies = oa.refcl_jacobi_solve_host(contacts, bodies, inertias, 0, iters)
    o_bodies = oa.jacobi_solve(contacts, bodies, inertias, 0, iters, host_order=True)
    moved = np.abs(ref_bodies["linVel"][:, :3] - bodies["linVel"][:, :3]).max()
    assert moved > 0.1
    for f in ("linVel", "angVel"):
        a, b = o_bodies[f][:, :3], ref_bodies[f][:, :3]
        err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))
        print(f, "max rel err", err, "bit-equal", np.array_equal(a.view(np.uint32), b.view(np.uint32)))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (f, err)  # bit for bit
    # the two loop orders really differ (so the switch is not a no-op)
    g_bodies = oa.jacobi_solve(contacts, bodies, inertias, 0, iters, host_order=False)
    assert np.abs(g_bodies["linVel"][:, :3] - o_bodies["linVel"][:, :3]).max() > 1e-4
