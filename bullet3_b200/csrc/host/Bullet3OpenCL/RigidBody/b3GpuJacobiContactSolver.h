// b3GpuJacobiContactSolver with the reference's public surface (src/Bullet3OpenCL/RigidBody/b3GpuJacobiContactSolver.h:16-47)
// on top of b3b200_solve_contacts_device: mass-splitting Jacobi, b3JacobiSolverInfo defaults (7 iterations).
#ifndef B3_GPU_JACOBI_CONTACT_SOLVER_H
#define B3_GPU_JACOBI_CONTACT_SOLVER_H
#include "Bullet3OpenCL/RigidBody/b3GpuPgsContactSolver.h"

struct b3JacobiSolverInfo
{
	int m_fixedBodyIndex;
	float m_deltaTime;
	float m_positionDrift;
	float m_positionConstraintCoeff;
	int m_numIterations;
	b3JacobiSolverInfo() : m_fixedBodyIndex(0), m_deltaTime(1. / 60.f), m_positionDrift(0.005f), m_positionConstraintCoeff(0.99f), m_numIterations(7) {}
};

class b3GpuJacobiContactSolver : public b3B200ContactSolverBase
{
public:
	b3GpuJacobiContactSolver(cl_context ctx, cl_device_id device, cl_command_queue queue, int pairCapacity);
	virtual ~b3GpuJacobiContactSolver() {}
	void solveContacts(int numBodies, cl_mem bodyBuf, cl_mem inertiaBuf, int numContacts, cl_mem contactBuf, const struct b3Config& config, int static0Index);
};
#endif
