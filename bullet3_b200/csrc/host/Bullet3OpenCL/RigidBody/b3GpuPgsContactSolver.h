// b3GpuPgsContactSolver with the reference's public surface (src/Bullet3OpenCL/RigidBody/b3GpuPgsContactSolver.h:31-34)
// on top of b3b200_solve_contacts_device (include/b3b200.h): batched PGS, 4 iterations like the reference
// (b3GpuPgsContactSolver.cpp:1049), on caller-owned device buffers with the reference AoS layouts.
#ifndef B3_GPU_BATCHING_PGS_SOLVER_H
#define B3_GPU_BATCHING_PGS_SOLVER_H
#include "Bullet3OpenCL/Initialize/b3OpenCLInclude.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"

class b3B200ContactSolverBase
{
protected:
	struct b3b200_world* m_scratch;
	int m_device, m_pairCapacity, m_bodyCapacity, m_contactCapacity, m_kind, m_iterations;
	void* m_queue;
	bool ensureScratch(int numBodies, int numContacts);
	void solve(int numBodies, cl_mem bodyBuf, cl_mem inertiaBuf, int numContacts, cl_mem contactBuf, int static0Index);

public:
	b3B200ContactSolverBase(cl_device_id device, cl_command_queue q, int pairCapacity, int kind, int iterations);
	virtual ~b3B200ContactSolverBase();
	// B200 addition: the reference hard-codes the iteration counts
	void setNumIterations(int n) { m_iterations = n; }
};

class b3GpuPgsContactSolver : public b3B200ContactSolverBase
{
public:
	b3GpuPgsContactSolver(cl_context ctx, cl_device_id device, cl_command_queue q, int pairCapacity);
	virtual ~b3GpuPgsContactSolver() {}
	void solveContacts(int numBodies, cl_mem bodyBuf, cl_mem inertiaBuf, int numContacts, cl_mem contactBuf, const struct b3Config& config, int static0Index);
};
#endif
