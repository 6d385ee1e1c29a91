"""Joint solver: the oracle (orc_solve_joints, a restatement of b3GpuPgsConstraintSolver::solveJoints) against the
reference's CPU joint path -- b3PgsJacobiSolver::solveContacts with b3Point2PointConstraint objects, the code
b3GpuRigidBodyPipeline::stepSimulation itself calls for b3TypedConstraint joints (b3GpuRigidBodyPipeline.cpp:375-385) and
that the GPU joint solver was ported from.  (The host twins inside b3GpuPgsConstraintSolver.cpp cannot run: with
useGpuInfo1 off they index m_cpuConstraintRowOffsets before anything fills it.)  The CPU solver walks the joints in index
order, the GPU solver in batch order: identical for one joint and for joints that share no dynamic body.
Runs without a GPU."""
import numpy as np
import pytest

import oracle_api as oa
from bullet3_b200 import capi, scenes

pytestmark = pytest.mark.skipif(not oa.ref_available(), reason="oracle/_ref/libb3ref.so not built")


def two_bodies(seed, static_a=False):
    rng = np.random.default_rng(seed)
    w = capi.World(capi.default_config(64), device=-1)
    box = w.register_convex_points(scenes.box_points(0.5))
    w.register_instance(0.0 if static_a else 1.5, rng.uniform(-1, 1, 3), scenes.random_quat(rng), box)
    w.register_instance(2.0, rng.uniform(-1, 1, 3) + (1.5, 0, 0), scenes.random_quat(rng), box)
    t = w.tables()
    bodies, inertias = t["bodies"].copy(), t["inertias"]
    bodies["linVel"][:, :3] = rng.uniform(-2, 2, (2, 3))
    bodies["angVel"][:, :3] = rng.uniform(-2, 2, (2, 3))
    if static_a:
        bodies["linVel"][0] = 0
        bodies["angVel"][0] = 0
    return bodies, inertias, rng


def p2p(rng, thr=1e30):
    j = np.zeros(1, capi.joint_t)
    j["constraintType"] = 3
    j["rbA"], j["rbB"] = 0, 1
    j["breakingImpulseThreshold"] = thr
    j["pivotInA"][0, :3] = rng.uniform(-0.5, 0.5, 3)
    j["pivotInB"][0, :3] = rng.uniform(-0.5, 0.5, 3)
    j["flags"] = 1
    return j


@pytest.mark.parametrize("seed,static_a", [(0, False), (1, False), (2, True), (3, True)])
def test_p2p_single_joint_matches_cpu_solver(seed, static_a):
    bodies, inertias, rng = two_bodies(seed, static_a)
    j = p2p(rng)
    rb = oa.solve_joints_ref(bodies, inertias, j)
    ob, oj = oa.solve_joints_oracle(bodies, inertias, j)
    assert not np.array_equal(rb["linVel"], bodies["linVel"])  # the joint did something
    for f in ("linVel", "angVel"):
        # b3PgsJacobiSolver and the GPU solver's host code order a few operations differently: 1-2 ulp
        assert np.allclose(rb[f][:, :3], ob[f][:, :3], rtol=2e-6, atol=2e-6), f
    assert oj["flags"][0] == 1


def test_p2p_breaking_threshold_disables_joint():
    bodies, inertias, rng = two_bodies(5)
    j = p2p(rng, thr=0.05)
    rb = oa.solve_joints_ref(bodies, inertias, j)
    ob, oj = oa.solve_joints_oracle(bodies, inertias, j)
    assert oj["flags"][0] == 0  # the impulse limits are +-threshold, a row that reaches them breaks the joint
    for f in ("linVel", "angVel"):
        # b3PgsJacobiSolver and the GPU solver's host code order a few operations differently: 1-2 ulp
        assert np.allclose(rb[f][:, :3], ob[f][:, :3], rtol=2e-6, atol=2e-6), f
    # a disabled joint is skipped
    ob2, oj2 = oa.solve_joints_oracle(ob, inertias, oj)
    assert np.array_equal(ob2["linVel"], ob["linVel"])


def test_chain_batches_and_convergence():
    """oracle-only: a chain of P2P joints pulls the pivots together over repeated solves + integration-free updates"""
    rng = np.random.default_rng(1)
    w = capi.World(capi.default_config(64), device=-1)
    box = w.register_convex_points(scenes.box_points(0.5))
    n = 8
    for i in range(n):
        w.register_instance(0.0 if i == 0 else 1.0, (1.2 * i, 0, 0), scenes.IDENT, box)
    t = w.tables()
    bodies, inertias = t["bodies"].copy(), t["inertias"]
    j = np.zeros(n - 1, capi.joint_t)
    j["constraintType"] = 3
    j["rbA"] = np.arange(n - 1)
    j["rbB"] = np.arange(1, n)
    j["breakingImpulseThreshold"] = 1e30
    j["pivotInA"][:, 0] = 0.5
    j["pivotInB"][:, 0] = -0.5  # pivots 0.2 apart along x: the joints pull the boxes together
    j["flags"] = 1
    ob, oj = oa.solve_joints_oracle(bodies, inertias, j)
    assert np.all(oj["flags"] == 1)
    assert np.all(ob["linVel"][0] == 0)  # the static anchor does not move
    assert ob["linVel"][1:, 0].mean() < 0 and ob["linVel"][-1, 0] < 0  # the chain is pulled towards the anchor
    assert np.isfinite(ob["linVel"]).all() and np.isfinite(ob["angVel"]).all()


def test_independent_joints_match_cpu_solver():
    """several joints that share no dynamic body: batch order == index order in effect"""
    rng = np.random.default_rng(9)
    w = capi.World(capi.default_config(64), device=-1)
    box = w.register_convex_points(scenes.box_points(0.5))
    w.register_instance(0.0, (0, 0, 0), scenes.IDENT, box)  # a shared static anchor
    m = 6
    for i in range(2 * m):
        w.register_instance(1.0 + 0.1 * i, rng.uniform(-3, 3, 3), scenes.random_quat(rng), box)
    t = w.tables()
    bodies, inertias = t["bodies"].copy(), t["inertias"]
    bodies["linVel"][1:, :3] = rng.uniform(-1, 1, (2 * m, 3))
    bodies["angVel"][1:, :3] = rng.uniform(-1, 1, (2 * m, 3))
    j = np.zeros(m + 3, capi.joint_t)
    j["constraintType"] = 3
    j["breakingImpulseThreshold"] = 1e30
    j["flags"] = 1
    for i in range(m):  # body 1+2i with body 2+2i ... except the last three, which hang on the static anchor
        j["rbA"][i], j["rbB"][i] = 1 + 2 * i, 2 + 2 * i
    for i in range(3):
        j["rbA"][m + i], j["rbB"][m + i] = 0, 0  # placeholders, fixed below
    j = j[:m]
    j["pivotInA"][:, :3] = rng.uniform(-0.5, 0.5, (m, 3))
    j["pivotInB"][:, :3] = rng.uniform(-0.5, 0.5, (m, 3))
    rb = oa.solve_joints_ref(bodies, inertias, j)
    ob, oj = oa.solve_joints_oracle(bodies, inertias, j)
    for f in ("linVel", "angVel"):
        # b3PgsJacobiSolver and the GPU solver's host code order a few operations differently: 1-2 ulp
        assert np.allclose(rb[f][:, :3], ob[f][:, :3], rtol=2e-6, atol=2e-6), f
