"""The C++ drop-in class layer (bullet3_b200/csrc/host): caller code in the style of
GpuRigidBodyDemo / PairBench, compiled against the reference's own headers, runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "bullet3_b200", "dropin_demo")
LIB = os.path.join(ROOT, "bullet3_b200", "libBullet3OpenCL_b200.so")


def test_dropin_library_exports_reference_classes():
    if not os.path.exists(LIB):
        pytest.skip("drop-in layer not built (needs the reference headers)")
    out = subprocess.run(["nm", "-DC", LIB], capture_output=True, text=True).stdout
    for sym in ("b3GpuRigidBodyPipeline::stepSimulation(float)", "b3GpuRigidBodyPipeline::registerPhysicsInstance(",
                "b3GpuNarrowPhase::registerConvexHullShape(float const*, int, int, float const*)", "b3GpuNarrowPhase::readbackAllBodiesToCpu()",
                "b3B200BroadphaseBase::calculateOverlappingPairs(int)", "b3GpuRigidBodyPipeline::getBodyBuffer()",
                "b3GpuRigidBodyPipeline::castRays(", "b3GpuRigidBodyPipeline::createPoint2PointConstraint(", "b3GpuPgsContactSolver::solveContacts(",
                "b3GpuJacobiContactSolver::solveContacts("):
        assert sym in out, sym


@pytest.mark.gpu
@pytest.mark.parametrize("bp", ["sap", "grid"])
def test_dropin_demo_runs(bp):
    if not os.path.exists(DEMO):
        pytest.skip("dropin_demo not built")
    r = subprocess.run([DEMO, bp[0]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DROPIN OK" in r.stdout, r.stdout + r.stderr
