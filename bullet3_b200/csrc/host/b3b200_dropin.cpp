// b3b200_dropin.cpp -- the reference's C++ class surface (b3GpuRigidBodyPipeline, b3GpuNarrowPhase,
// b3GpuSapBroadphase / b3GpuGridBroadphase) implemented as thin forwards to the C ABI of
// include/b3b200.h.  Compiled against the reference's own Bullet3Common / shared POD headers
// (-I bullet3_b200/csrc/host -I <bullet3>/src), so caller code written for src/Bullet3OpenCL
// (examples/OpenCL/rigidbody/GpuRigidBodyDemo.cpp:129-150, broadphase/PairBench.cpp:208-379)
// compiles unchanged.  No physics is computed in this file.
#include <string.h>
#include <stdio.h>
#include "Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.h"
#include "Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.h"
#include "Bullet3OpenCL/RigidBody/b3GpuPgsContactSolver.h"
#include "Bullet3OpenCL/RigidBody/b3GpuJacobiContactSolver.h"
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuSapBroadphase.h"
#include <vector>
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuGridBroadphase.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3ConvexUtility.h"
#include "Bullet3Common/b3Logging.h"
#include "../../../include/b3b200.h"

static_assert(sizeof(b3RigidBodyData) == sizeof(b3b200_rigid_body), "abi");
static_assert(sizeof(b3Collidable) == sizeof(b3b200_collidable), "abi");
static_assert(sizeof(b3SapAabb) == sizeof(b3b200_aabb), "abi");
static_assert(sizeof(b3Contact4Data) == sizeof(b3b200_contact4), "abi");
static_assert(sizeof(b3GpuChildShape) == sizeof(b3b200_child_shape), "abi");
static_assert(sizeof(b3Config) == sizeof(b3b200_config), "abi");
static_assert(sizeof(b3Int4) == sizeof(b3b200_int4), "abi");

static void reportError(const char* what) { b3Error("%s: %s\n", what, b3b200_last_error()); }

// ---------------------------------------------------------------------------- broadphase
b3B200BroadphaseBase::b3B200BroadphaseBase(int kind, cl_device_id device, cl_command_queue q)
	: m_kind(kind), m_device(b3b200DeviceOrdinal(device)), m_stream(q), m_bp(0), m_world(0), m_numOverlap(0)
{
}
b3B200BroadphaseBase::~b3B200BroadphaseBase()
{
	if (m_bp) b3b200_bp_destroy(m_bp);
}
void b3B200BroadphaseBase::reset()
{
	m_pending.clear();
	m_allAabbsCPU.clear();
	if (m_bp) b3b200_bp_destroy(m_bp);
	m_bp = 0;
}
static void addProxy(b3AlignedObjectArray<b3SapAabb>& all, const b3Vector3& mn, const b3Vector3& mx, int userPtr)
{
	b3SapAabb a;
	a.m_minVec = mn;
	a.m_maxVec = mx;
	a.m_minIndices[3] = userPtr;
	a.m_signedMaxIndices[3] = all.size();
	all.push_back(a);
}
void b3B200BroadphaseBase::createProxy(const b3Vector3& aabbMin, const b3Vector3& aabbMax, int userPtr, int, int)
{
	if (m_world) return;  // the pipeline's world already holds one proxy per body
	Proxy p = {{aabbMin.x, aabbMin.y, aabbMin.z}, {aabbMax.x, aabbMax.y, aabbMax.z}, userPtr, false};
	m_pending.push_back(p);
	addProxy(m_allAabbsCPU, aabbMin, aabbMax, userPtr);
}
void b3B200BroadphaseBase::createLargeProxy(const b3Vector3& aabbMin, const b3Vector3& aabbMax, int userPtr, int, int)
{
	if (m_world) return;
	Proxy p = {{aabbMin.x, aabbMin.y, aabbMin.z}, {aabbMax.x, aabbMax.y, aabbMax.z}, userPtr, true};
	m_pending.push_back(p);
	addProxy(m_allAabbsCPU, aabbMin, aabbMax, userPtr);
}
void b3B200BroadphaseBase::ensureStandalone(int maxPairs)
{
	if (m_bp || m_world) return;
	int n = m_pending.size() > 0 ? m_pending.size() : 1;
	if (b3b200_bp_create(m_kind, m_device, m_stream, n, maxPairs > 0 ? maxPairs : 16 * n, &m_bp) < 0)
	{
		reportError("b3GpuBroadphase");
		m_bp = 0;
		return;
	}
	for (int i = 0; i < m_pending.size(); i++)
	{
		const Proxy& p = m_pending[i];
		if (p.large)
			b3b200_bp_create_large_proxy(m_bp, p.mn, p.mx, p.userPtr);
		else
			b3b200_bp_create_proxy(m_bp, p.mn, p.mx, p.userPtr);
	}
	b3b200_bp_write_aabbs(m_bp);
}
void b3B200BroadphaseBase::writeAabbsToGpu()
{
	if (m_world) return;  // done by b3GpuRigidBodyPipeline::writeAllInstancesToGpu
	if (m_bp)
	{
		// AABBs edited on the host through getAllAabbsCPU() are pushed again
		if (b3b200_bp_set_aabbs(m_bp, (const b3b200_aabb*)&m_allAabbsCPU[0], m_allAabbsCPU.size()) < 0) reportError("writeAabbsToGpu");
	}
	else
		ensureStandalone(0);
}
void b3B200BroadphaseBase::calculateOverlappingPairs(int maxPairs)
{
	if (m_world)
	{
		if (b3b200_find_pairs(m_world) < 0) reportError("calculateOverlappingPairs");
		int c[8];
		b3b200_get_counters(m_world, c);
		m_numOverlap = c[0];
		return;
	}
	ensureStandalone(maxPairs);
	if (!m_bp) return;
	if (b3b200_bp_calculate_pairs(m_bp, maxPairs) < 0) reportError("calculateOverlappingPairs");
	m_numOverlap = b3b200_bp_num_overlap(m_bp);
}
int b3B200BroadphaseBase::getNumOverlap()
{
	if (m_world)
	{
		// the pipeline's step finds the pairs; the count lives in the world's device counters
		int c[8];
		if (b3b200_get_counters(m_world, c) == 0) m_numOverlap = c[0];
	}
	return m_numOverlap;
}
cl_mem b3B200BroadphaseBase::getAabbBufferWS()
{
	void* p = 0;
	if (m_world)
		b3b200_device_buffer(m_world, B3B200_BUF_AABBS, &p);
	else if (m_bp)
		b3b200_bp_device_aabbs(m_bp, &p);
	return p;
}
cl_mem b3B200BroadphaseBase::getOverlappingPairBuffer()
{
	void* p = 0;
	if (m_world)
		b3b200_device_buffer(m_world, B3B200_BUF_PAIRS, &p);
	else if (m_bp)
		b3b200_bp_device_pairs(m_bp, &p);
	return p;
}
b3OpenCLArray<b3SapAabb>& b3B200BroadphaseBase::getAllAabbsGPU()
{
	int n = m_world ? b3b200_num_bodies(m_world) : m_allAabbsCPU.size();
	m_allAabbsGPU.setView(getAabbBufferWS(), n, m_device);
	return m_allAabbsGPU;
}
b3AlignedObjectArray<b3SapAabb>& b3B200BroadphaseBase::getAllAabbsCPU()
{
	if (m_world)
	{
		int n = b3b200_num_bodies(m_world);
		m_allAabbsCPU.resize(n);
		if (n) b3b200_get_aabbs(m_world, (b3b200_aabb*)&m_allAabbsCPU[0], n);
	}
	return m_allAabbsCPU;
}
b3OpenCLArray<b3Int4>& b3B200BroadphaseBase::getOverlappingPairsGPU()
{
	m_pairsGPU.setView(getOverlappingPairBuffer(), getNumOverlap(), m_device);
	return m_pairsGPU;
}

// ---------------------------------------------------------------------------- narrowphase
b3GpuNarrowPhase::b3GpuNarrowPhase(cl_context, cl_device_id dev, cl_command_queue q, const b3Config& config) : m_world(0), m_device(b3b200DeviceOrdinal(dev)), m_static0Index(-1)
{
	if (b3b200_create((const b3b200_config*)&config, m_device, q, &m_world) < 0)
	{
		reportError("b3GpuNarrowPhase");
		m_world = 0;
	}
}
b3GpuNarrowPhase::~b3GpuNarrowPhase()
{
	if (m_world) b3b200_destroy(m_world);
}
void b3GpuNarrowPhase::refreshTables() const
{
	int n = 0;
	b3b200_get_table(m_world, B3B200_TBL_COLLIDABLES, 0, 0, &n);
	m_collidablesCPU.resize(n);
	m_localAabbsCPU.resize(n);
	if (n)
	{
		b3b200_get_table(m_world, B3B200_TBL_COLLIDABLES, &m_collidablesCPU[0], n, &n);
		b3b200_get_table(m_world, B3B200_TBL_LOCAL_AABBS, &m_localAabbsCPU[0], n, &n);
	}
}
static int checked(int r, const char* what)
{
	if (r < 0) reportError(what);
	return r;
}
int b3GpuNarrowPhase::registerSphereShape(float radius) { return checked(b3b200_register_sphere(m_world, radius), "registerSphereShape"); }
int b3GpuNarrowPhase::registerPlaneShape(const b3Vector3& n, float c)
{
	float nn[3] = {n.x, n.y, n.z};
	return checked(b3b200_register_plane(m_world, nn, c), "registerPlaneShape");
}
int b3GpuNarrowPhase::registerCompoundShape(b3AlignedObjectArray<b3GpuChildShape>* ch)
{
	return checked(b3b200_register_compound(m_world, (const b3b200_child_shape*)&(*ch)[0], ch->size()), "registerCompoundShape");
}
int b3GpuNarrowPhase::registerConcaveMesh(b3AlignedObjectArray<b3Vector3>* v, b3AlignedObjectArray<int>* idx, const float* scaling)
{
	b3AlignedObjectArray<float> flat;
	flat.resize(3 * v->size());
	for (int i = 0; i < v->size(); i++)
	{
		flat[3 * i] = (*v)[i].x;
		flat[3 * i + 1] = (*v)[i].y;
		flat[3 * i + 2] = (*v)[i].z;
	}
	return checked(b3b200_register_concave(m_world, &flat[0], v->size(), &(*idx)[0], idx->size(), scaling), "registerConcaveMesh");
}
int b3GpuNarrowPhase::registerConvexHullShape(const float* vertices, int strideInBytes, int numVertices, const float* scaling)
{
	return checked(b3b200_register_convex_points(m_world, vertices, strideInBytes, numVertices, scaling), "registerConvexHullShape");
}
int b3GpuNarrowPhase::registerConvexHullShape(b3ConvexUtility* u)
{
	// flat tables as in registerConvexHullShapeInternal (b3GpuNarrowPhase.cpp:234-296)
	b3AlignedObjectArray<b3b200_face> faces;
	b3AlignedObjectArray<int> indices;
	for (int i = 0; i < u->m_faces.size(); i++)
	{
		b3b200_face f;
		f.plane.x = u->m_faces[i].m_plane[0];
		f.plane.y = u->m_faces[i].m_plane[1];
		f.plane.z = u->m_faces[i].m_plane[2];
		f.plane.w = u->m_faces[i].m_plane[3];
		f.indexOffset = indices.size();
		f.numIndices = u->m_faces[i].m_indices.size();
		f.pad1 = f.pad2 = 0;
		for (int p = 0; p < f.numIndices; p++) indices.push_back(u->m_faces[i].m_indices[p]);
		faces.push_back(f);
	}
	b3b200_convex_polyhedron poly;
	memset(&poly, 0, sizeof(poly));
	memcpy(&poly.extents, &u->m_extents, 16);
	memcpy(&poly.mC, &u->mC, 16);
	memcpy(&poly.mE, &u->mE, 16);
	if (!u->m_vertices.size() || !faces.size()) return -1;
	return checked(b3b200_register_convex(m_world, (const b3b200_float4*)&u->m_vertices[0], u->m_vertices.size(), &faces[0], faces.size(), &indices[0],
										  indices.size(), u->m_uniqueEdges.size() ? (const b3b200_float4*)&u->m_uniqueEdges[0] : 0, u->m_uniqueEdges.size(), &poly),
				   "registerConvexHullShape");
}
int b3GpuNarrowPhase::registerRigidBody(int collidableIndex, float mass, const float* position, const float* orientation, const float* aabbMin, const float* aabbMax, bool)
{
	int r = checked(b3b200_register_body(m_world, collidableIndex, mass, position, orientation, aabbMin, aabbMax), "registerRigidBody");
	if (r == 0 && mass == 0.f) m_static0Index = 0;
	return r;
}
void b3GpuNarrowPhase::writeAllBodiesToGpu()
{
	// b3GpuNarrowPhase::writeAllBodiesToGpu (b3GpuNarrowPhase.cpp:970-979) sends m_bodyBufferCPU: the wrapper's CPU copy is the
	// authority for the bodies it holds (setObjectTransformCpu / setObjectVelocityCpu edit it), the world's upload only adds
	// what was registered since (b3b200_upload keeps the device state of the bodies that are already there)
	checked(b3b200_upload(m_world), "writeAllBodiesToGpu");
	const int n = b3b200_num_bodies(m_world);
	if (m_cpuEdited && m_bodiesCPU.size() == n && n > 0)
		checked(b3b200_write_bodies(m_world, (const b3b200_rigid_body*)&m_bodiesCPU[0], n), "writeAllBodiesToGpu");
	m_cpuEdited = false;
}
void b3GpuNarrowPhase::reset()
{
	b3b200_reset(m_world);
	m_static0Index = -1;
}
void b3GpuNarrowPhase::readbackAllBodiesToCpu()
{
	int n = b3b200_num_bodies(m_world);
	m_bodiesCPU.resize(n);
	if (n) checked(b3b200_readback_bodies(m_world, (b3b200_rigid_body*)&m_bodiesCPU[0], n), "readbackAllBodiesToCpu");
	m_cpuEdited = false;
}
const b3RigidBodyData* b3GpuNarrowPhase::getBodiesCpu() const
{
	if (m_bodiesCPU.size() != b3b200_num_bodies(m_world))
	{
		// first use, or bodies were registered since: the world's table = the current device state of the uploaded bodies +
		// the registration state of the new ones (b3b200_get_table refreshes it from the device)
		int n = b3b200_num_bodies(m_world);
		b3AlignedObjectArray<b3RigidBodyData> fresh;
		fresh.resize(n);
		if (n) b3b200_get_table(m_world, B3B200_TBL_BODIES, &fresh[0], n, &n);
		for (int i = 0; i < m_bodiesCPU.size() && i < n && m_cpuEdited; i++) fresh[i] = m_bodiesCPU[i];  // keep pending edits
		m_bodiesCPU = fresh;
	}
	return m_bodiesCPU.size() ? &m_bodiesCPU[0] : 0;
}
bool b3GpuNarrowPhase::getObjectTransformFromCpu(float* position, float* orientation, int bodyIndex) const
{
	const b3RigidBodyData* b = getBodiesCpu();
	if (!b || bodyIndex < 0 || bodyIndex >= m_bodiesCPU.size())
	{
		b3Warning("getObjectTransformFromCpu out of range.\n");
		return false;
	}
	position[0] = b[bodyIndex].m_pos.x;
	position[1] = b[bodyIndex].m_pos.y;
	position[2] = b[bodyIndex].m_pos.z;
	position[3] = 1.f;
	orientation[0] = b[bodyIndex].m_quat.x;
	orientation[1] = b[bodyIndex].m_quat.y;
	orientation[2] = b[bodyIndex].m_quat.z;
	orientation[3] = b[bodyIndex].m_quat.w;
	return true;
}
void b3GpuNarrowPhase::setObjectTransformCpu(float* position, float* orientation, int bodyIndex)
{
	getBodiesCpu();
	if (bodyIndex < 0 || bodyIndex >= m_bodiesCPU.size())
	{
		b3Warning("setObjectVelocityCpu out of range.\n");
		return;
	}
	m_bodiesCPU[bodyIndex].m_pos = b3MakeVector3(position[0], position[1], position[2]);
	m_bodiesCPU[bodyIndex].m_quat.setValue(orientation[0], orientation[1], orientation[2], orientation[3]);
	m_cpuEdited = true;
}
void b3GpuNarrowPhase::setObjectVelocityCpu(float* linVel, float* angVel, int bodyIndex)
{
	getBodiesCpu();
	if (bodyIndex < 0 || bodyIndex >= m_bodiesCPU.size())
	{
		b3Warning("setObjectVelocityCpu out of range.\n");
		return;
	}
	m_bodiesCPU[bodyIndex].m_linVel = b3MakeVector3(linVel[0], linVel[1], linVel[2]);
	m_bodiesCPU[bodyIndex].m_angVel = b3MakeVector3(angVel[0], angVel[1], angVel[2]);
	m_cpuEdited = true;
}
void b3GpuNarrowPhase::setObjectTransform(const float* position, const float* orientation, int bodyIndex)
{
	// b3GpuNarrowPhase::setObjectTransform: the CPU copy of THAT body + an immediate write of that one body to the device
	// (the other bodies keep their device state: the CPU copy may be older than the simulation)
	if (bodyIndex < 0 || bodyIndex >= b3b200_num_bodies(m_world))
	{
		b3Warning("setObjectTransform out of range.\n");
		return;
	}
	b3RigidBodyData one;
	checked(b3b200_read_body(m_world, bodyIndex, (b3b200_rigid_body*)&one), "setObjectTransform");
	one.m_pos = b3MakeVector3(position[0], position[1], position[2]);
	one.m_quat.setValue(orientation[0], orientation[1], orientation[2], orientation[3]);
	checked(b3b200_write_body(m_world, bodyIndex, (const b3b200_rigid_body*)&one), "setObjectTransform");
	if (bodyIndex < m_bodiesCPU.size())
	{
		m_bodiesCPU[bodyIndex].m_pos = one.m_pos;
		m_bodiesCPU[bodyIndex].m_quat = one.m_quat;
	}
}
void b3GpuNarrowPhase::computeContacts(cl_mem, int, cl_mem, int) { checked(b3b200_compute_contacts(m_world), "computeContacts"); }
cl_mem b3GpuNarrowPhase::getBodiesGpu()
{
	void* p = 0;
	b3b200_device_buffer(m_world, B3B200_BUF_BODIES, &p);
	return p;
}
int b3GpuNarrowPhase::getNumBodiesGpu() const { return b3b200_num_bodies(m_world); }
cl_mem b3GpuNarrowPhase::getBodyInertiasGpu()
{
	void* p = 0;
	b3b200_device_buffer(m_world, B3B200_BUF_INERTIAS, &p);
	return p;
}
int b3GpuNarrowPhase::getNumBodyInertiasGpu() const { return b3b200_num_bodies(m_world); }
const b3Collidable* b3GpuNarrowPhase::getCollidablesCpu() const
{
	refreshTables();
	return m_collidablesCPU.size() ? &m_collidablesCPU[0] : 0;
}
int b3GpuNarrowPhase::getNumCollidablesGpu() const
{
	int n = 0;
	b3b200_get_table(m_world, B3B200_TBL_COLLIDABLES, 0, 0, &n);
	return n;
}
const b3SapAabb* b3GpuNarrowPhase::getLocalSpaceAabbsCpu() const
{
	refreshTables();
	return m_localAabbsCPU.size() ? &m_localAabbsCPU[0] : 0;
}
const b3SapAabb& b3GpuNarrowPhase::getLocalSpaceAabb(int collidableIndex) const
{
	refreshTables();
	return m_localAabbsCPU[collidableIndex];
}
const b3Contact4Data* b3GpuNarrowPhase::getContactsCPU() const
{
	int n = 0;
	b3b200_get_contacts(m_world, 0, 0, &n);
	m_contactsCPU.resize(n);
	if (n) b3b200_get_contacts(m_world, (b3b200_contact4*)&m_contactsCPU[0], n, &n);
	return n ? &m_contactsCPU[0] : 0;
}
cl_mem b3GpuNarrowPhase::getContactsGpu()
{
	void* p = 0;
	b3b200_device_buffer(m_world, B3B200_BUF_CONTACTS, &p);
	return p;
}
int b3GpuNarrowPhase::getNumContactsGpu() const
{
	int n = 0;
	b3b200_get_contacts(m_world, 0, 0, &n);
	return n;
}
int b3GpuNarrowPhase::getNumRigidBodies() const { return b3b200_num_bodies(m_world); }

// ---------------------------------------------------------------------------- pipeline
b3GpuRigidBodyPipeline::b3GpuRigidBodyPipeline(cl_context, cl_device_id, cl_command_queue, b3GpuNarrowPhase* narrowphase, b3GpuBroadphaseInterface* broadphaseSap,
											   struct b3DynamicBvhBroadphase*, const b3Config& config)
	: m_np(narrowphase), m_bp(static_cast<b3B200BroadphaseBase*>(broadphaseSap)), m_config(config)
{
	if (m_bp && m_np)
	{
		m_bp->m_world = m_np->m_world;  // the caller's broadphase object becomes a view of the world's broadphase
		b3b200_set_broadphase(m_np->m_world, m_bp->m_kind);
	}
}
b3GpuRigidBodyPipeline::~b3GpuRigidBodyPipeline()
{
	if (m_bp) m_bp->m_world = 0;  // np and bp are owned by the caller (b3GpuRigidBodyPipeline.cpp:116-139)
}
void b3GpuRigidBodyPipeline::stepSimulation(float dt) { checked(b3b200_step(m_np->m_world, dt), "stepSimulation"); }
void b3GpuRigidBodyPipeline::integrate(float dt) { checked(b3b200_integrate(m_np->m_world, dt), "integrate"); }
void b3GpuRigidBodyPipeline::setupGpuAabbsFull() { checked(b3b200_update_aabbs(m_np->m_world), "setupGpuAabbsFull"); }
int b3GpuRigidBodyPipeline::registerConvexPolyhedron(b3ConvexUtility* convex) { return m_np->registerConvexHullShape(convex); }
int b3GpuRigidBodyPipeline::registerPhysicsInstance(float mass, const float* position, const float* orientation, int collidableIndex, int userIndex, bool writeInstanceToGpu)
{
	int r = b3b200_register_instance(m_np->m_world, mass, position, orientation, collidableIndex, userIndex);
	if (r < 0)
	{
		reportError("registerPhysicsInstance");
		return -1;
	}
	if (r == 0 && mass == 0.f) m_np->m_static0Index = 0;
	if (writeInstanceToGpu) writeAllInstancesToGpu();
	return r;
}
void b3GpuRigidBodyPipeline::writeAllInstancesToGpu() { checked(b3b200_upload(m_np->m_world), "writeAllInstancesToGpu"); }
int b3GpuRigidBodyPipeline::createPoint2PointConstraint(int bodyA, int bodyB, const float* pivotInA, const float* pivotInB, float breakingThreshold)
{
	return checked(b3b200_create_p2p_constraint(m_np->m_world, bodyA, bodyB, pivotInA, pivotInB, breakingThreshold), "createPoint2PointConstraint");
}
int b3GpuRigidBodyPipeline::createFixedConstraint(int bodyA, int bodyB, const float* pivotInA, const float* pivotInB, const float* relTargetAB, float breakingThreshold)
{
	return checked(b3b200_create_fixed_constraint(m_np->m_world, bodyA, bodyB, pivotInA, pivotInB, relTargetAB, breakingThreshold), "createFixedConstraint");
}
void b3GpuRigidBodyPipeline::removeConstraintByUid(int uid) { checked(b3b200_remove_constraint(m_np->m_world, uid), "removeConstraintByUid"); }
void b3GpuRigidBodyPipeline::copyConstraintsToHost() { checked(b3b200_get_joints(m_np->m_world, 0, 0, 0), "copyConstraintsToHost"); }
int b3GpuRigidBodyPipeline::getNumConstraints() const { return b3b200_num_constraints(m_np->m_world); }
void b3GpuRigidBodyPipeline::castRays(const b3AlignedObjectArray<b3RayInfo>& rays, b3AlignedObjectArray<b3RayHit>& hitResults)
{
	static_assert(sizeof(b3RayInfo) == sizeof(b3b200_ray_info) && sizeof(b3RayHit) == sizeof(b3b200_ray_hit), "ray records are the reference's PODs");
	if (!rays.size()) return;
	checked(b3b200_cast_rays(m_np->m_world, (const b3b200_ray_info*)&rays[0], rays.size(), (b3b200_ray_hit*)&hitResults[0]), "castRays");
}
void b3GpuRigidBodyPipeline::setGravity(const float* g) { b3b200_set_gravity(m_np->m_world, g); }
void b3GpuRigidBodyPipeline::reset() { m_np->reset(); }
void b3GpuRigidBodyPipeline::setSolver(bool jacobi, int iterations) { b3b200_set_solver(m_np->m_world, jacobi ? B3B200_SOLVER_JACOBI : B3B200_SOLVER_PGS, iterations); }
cl_mem b3GpuRigidBodyPipeline::getBodyBuffer() { return m_np->getBodiesGpu(); }
int b3GpuRigidBodyPipeline::getNumBodies() const { return m_np->getNumRigidBodies(); }

// ------------------------------------------------------------------ stand-alone contact solvers
// The solver kernels work on a world's buffers, so each solver object keeps a private scratch world with enough body and
// contact slots (placeholder bodies; grown on demand) and runs b3b200_solve_contacts_device on the caller's buffers.
b3B200ContactSolverBase::b3B200ContactSolverBase(cl_device_id device, cl_command_queue q, int pairCapacity, int kind, int iterations)
	: m_scratch(0), m_device(b3b200DeviceOrdinal(device)), m_pairCapacity(pairCapacity), m_bodyCapacity(0), m_contactCapacity(0), m_kind(kind), m_iterations(iterations), m_queue(q)
{
}
b3B200ContactSolverBase::~b3B200ContactSolverBase()
{
	if (m_scratch) b3b200_destroy(m_scratch);
}
bool b3B200ContactSolverBase::ensureScratch(int numBodies, int numContacts)
{
	if (m_scratch && numBodies <= m_bodyCapacity && numContacts <= m_contactCapacity) return true;
	if (m_scratch) b3b200_destroy(m_scratch);
	m_scratch = 0;
	b3b200_config cfg;
	b3b200_config_default(&cfg);
	m_bodyCapacity = numBodies + numBodies / 4 + 16;
	m_contactCapacity = numContacts + numContacts / 4 + 16;
	if (m_contactCapacity < m_pairCapacity) m_contactCapacity = m_pairCapacity;
	cfg.maxConvexBodies = m_bodyCapacity;
	cfg.maxContactCapacity = m_contactCapacity;
	cfg.maxBroadphasePairs = 1024;  // the scratch world never finds pairs
	cfg.compoundPairCapacity = 0;
	cfg.maxTriConvexPairCapacity = 0;
	if (b3b200_create(&cfg, m_device, m_queue, &m_scratch) < 0) return reportError("contact solver scratch world"), false;
	const int shape = b3b200_register_sphere(m_scratch, 0.5f);
	std::vector<float> mass((size_t)m_bodyCapacity, 1.f), pos((size_t)m_bodyCapacity * 4, 0.f), orn((size_t)m_bodyCapacity * 4, 0.f);
	std::vector<int> col((size_t)m_bodyCapacity, shape);
	for (int i = 0; i < m_bodyCapacity; i++)
	{
		pos[4 * (size_t)i] = 4.f * (float)i;
		orn[4 * (size_t)i + 3] = 1.f;
	}
	if (shape < 0 || b3b200_register_instances(m_scratch, m_bodyCapacity, &mass[0], &pos[0], &orn[0], &col[0]) < 0 || b3b200_upload(m_scratch) < 0)
	{
		reportError("contact solver scratch world");
		b3b200_destroy(m_scratch);
		m_scratch = 0;
		return false;
	}
	return true;
}
void b3B200ContactSolverBase::solve(int numBodies, cl_mem bodyBuf, cl_mem inertiaBuf, int numContacts, cl_mem contactBuf, int static0Index)
{
	if (numBodies <= 0 || numContacts <= 0) return;  // b3GpuPgsContactSolver.cpp:1105 / b3GpuJacobiContactSolver.cpp:701: nothing to do
	if (!ensureScratch(numBodies, numContacts)) return;
	b3b200_set_solver(m_scratch, m_kind, m_iterations);
	checked(b3b200_solve_contacts_device(m_scratch, numBodies, bodyBuf, inertiaBuf, numContacts, contactBuf, static0Index), "solveContacts");
}
b3GpuPgsContactSolver::b3GpuPgsContactSolver(cl_context, cl_device_id device, cl_command_queue q, int pairCapacity)
	: b3B200ContactSolverBase(device, q, pairCapacity, B3B200_SOLVER_PGS, 4)
{
}
void b3GpuPgsContactSolver::solveContacts(int numBodies, cl_mem bodyBuf, cl_mem inertiaBuf, int numContacts, cl_mem contactBuf, const b3Config&, int static0Index)
{
	solve(numBodies, bodyBuf, inertiaBuf, numContacts, contactBuf, static0Index);
}
b3GpuJacobiContactSolver::b3GpuJacobiContactSolver(cl_context, cl_device_id device, cl_command_queue q, int pairCapacity)
	: b3B200ContactSolverBase(device, q, pairCapacity, B3B200_SOLVER_JACOBI, b3JacobiSolverInfo().m_numIterations)
{
}
void b3GpuJacobiContactSolver::solveContacts(int numBodies, cl_mem bodyBuf, cl_mem inertiaBuf, int numContacts, cl_mem contactBuf, const b3Config&, int static0Index)
{
	solve(numBodies, bodyBuf, inertiaBuf, numContacts, contactBuf, static0Index);
}
