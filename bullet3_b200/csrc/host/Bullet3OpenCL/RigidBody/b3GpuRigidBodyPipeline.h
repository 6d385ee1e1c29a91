// b3GpuRigidBodyPipeline with the reference's public surface
// (src/Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.h:25-68) on top of the C ABI (include/b3b200.h).
#ifndef B3_GPU_RIGIDBODY_PIPELINE_H
#define B3_GPU_RIGIDBODY_PIPELINE_H
#include "Bullet3OpenCL/Initialize/b3OpenCLInclude.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"
#include "Bullet3Common/b3AlignedObjectArray.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3RaycastInfo.h"

class b3GpuRigidBodyPipeline
{
protected:
	class b3GpuNarrowPhase* m_np;
	class b3B200BroadphaseBase* m_bp;
	b3Config m_config;

public:
	b3GpuRigidBodyPipeline(cl_context ctx, cl_device_id device, cl_command_queue q, class b3GpuNarrowPhase* narrowphase,
						   class b3GpuBroadphaseInterface* broadphaseSap, struct b3DynamicBvhBroadphase* broadphaseDbvt, const b3Config& config);
	virtual ~b3GpuRigidBodyPipeline();

	void stepSimulation(float deltaTime);
	void integrate(float timeStep);
	void setupGpuAabbsFull();
	// declared but never defined in the reference; here it forwards to the narrowphase
	int registerConvexPolyhedron(class b3ConvexUtility* convex);
	int registerPhysicsInstance(float mass, const float* position, const float* orientation, int collisionShapeIndex, int userData, bool writeInstanceToGpu);
	void writeAllInstancesToGpu();
	// joints (b3GpuRigidBodyPipeline.h:28-34 of the reference); solved before the contacts, broken joints get flags = 0
	int createPoint2PointConstraint(int bodyA, int bodyB, const float* pivotInA, const float* pivotInB, float breakingThreshold);
	int createFixedConstraint(int bodyA, int bodyB, const float* pivotInA, const float* pivotInB, const float* relTargetAB, float breakingThreshold);
	void removeConstraintByUid(int uid);
	void copyConstraintsToHost();
	int getNumConstraints() const;
	// b3GpuRigidBodyPipeline.h:63 of the reference; hitResults[i].m_hitFraction is the caller's cap on entry (b3GpuRaycast.cpp:182)
	void castRays(const b3AlignedObjectArray<b3RayInfo>& rays, b3AlignedObjectArray<b3RayHit>& hitResults);
	void setGravity(const float* grav);
	void reset();
	// B200 additions: solver selection (the reference uses the global gUseJacobi) and iteration count
	void setSolver(bool jacobi, int iterations);
	cl_mem getBodyBuffer();
	int getNumBodies() const;
};
#endif
