// b3GpuNarrowPhase with the reference's public surface (src/Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.h:9-99).
// It owns the B200 world handle (bodies, inertias, shapes, contacts live there), exactly like the reference
// narrowphase owns those buffers (b3GpuNarrowPhaseInternalData.h:24-86).
#ifndef B3_GPU_NARROWPHASE_H
#define B3_GPU_NARROWPHASE_H
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3Collidable.h"
#include "Bullet3OpenCL/Initialize/b3OpenCLInclude.h"
#include "Bullet3Common/b3AlignedObjectArray.h"
#include "Bullet3Common/b3Vector3.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3RigidBodyData.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3Contact4Data.h"
#include "Bullet3OpenCL/BroadphaseCollision/b3SapAabb.h"

struct b3b200_world;
class b3ConvexUtility;

class b3GpuNarrowPhase
{
protected:
	b3b200_world* m_world;
	int m_device;
	int m_static0Index;
	mutable b3AlignedObjectArray<b3RigidBodyData> m_bodiesCPU;
	mutable bool m_cpuEdited = false;  // m_bodiesCPU holds setObject*Cpu edits that writeAllBodiesToGpu has to send
	mutable b3AlignedObjectArray<b3Collidable> m_collidablesCPU;
	mutable b3AlignedObjectArray<b3SapAabb> m_localAabbsCPU;
	mutable b3AlignedObjectArray<b3Contact4Data> m_contactsCPU;
	void refreshTables() const;
	friend class b3GpuRigidBodyPipeline;

public:
	b3GpuNarrowPhase(cl_context vtx, cl_device_id dev, cl_command_queue q, const struct b3Config& config);
	virtual ~b3GpuNarrowPhase();

	int registerSphereShape(float radius);
	int registerPlaneShape(const b3Vector3& planeNormal, float planeConstant);
	int registerCompoundShape(b3AlignedObjectArray<b3GpuChildShape>* childShapes);
	int registerConcaveMesh(b3AlignedObjectArray<b3Vector3>* vertices, b3AlignedObjectArray<int>* indices, const float* scaling);
	// b3ConvexUtility* overload: the polyhedron tables are taken as they are (m_vertices, m_faces, m_uniqueEdges)
	int registerConvexHullShape(b3ConvexUtility* utilPtr);
	int registerConvexHullShape(const float* vertices, int strideInBytes, int numVertices, const float* scaling);

	// NB: also creates the broadphase proxy (the B200 world keeps AABB index == body index)
	int registerRigidBody(int collidableIndex, float mass, const float* position, const float* orientation, const float* aabbMin, const float* aabbMax, bool writeToGpu);
	void setObjectTransform(const float* position, const float* orientation, int bodyIndex);
	void writeAllBodiesToGpu();
	void reset();
	void readbackAllBodiesToCpu();
	bool getObjectTransformFromCpu(float* position, float* orientation, int bodyIndex) const;
	void setObjectTransformCpu(float* position, float* orientation, int bodyIndex);
	void setObjectVelocityCpu(float* linVel, float* angVel, int bodyIndex);

	virtual void computeContacts(cl_mem broadphasePairs, int numBroadphasePairs, cl_mem aabbsWorldSpace, int numObjects);

	cl_mem getBodiesGpu();
	const struct b3RigidBodyData* getBodiesCpu() const;
	int getNumBodiesGpu() const;
	cl_mem getBodyInertiasGpu();
	int getNumBodyInertiasGpu() const;
	const struct b3Collidable* getCollidablesCpu() const;
	int getNumCollidablesGpu() const;
	const struct b3SapAabb* getLocalSpaceAabbsCpu() const;
	const struct b3Contact4Data* getContactsCPU() const;
	cl_mem getContactsGpu();
	int getNumContactsGpu() const;
	int getNumRigidBodies() const;
	int getStatic0Index() const { return m_static0Index; }
	const struct b3SapAabb& getLocalSpaceAabb(int collidableIndex) const;
	b3b200_world* getB200World() { return m_world; }
};
#endif
