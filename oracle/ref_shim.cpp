// ref_shim.cpp -- thin extern "C" wrappers around the UNMODIFIED reference
// sources under /root/reference/src, compiled where they lie (see Makefile) into
// oracle/_ref/libb3ref.so.  TEST INFRASTRUCTURE ONLY.  Nothing here restates an
// algorithm: every function calls straight into a reference header/function.
// Used to pin oracle/oracle.cpp (tests/test_oracle_vs_ref.py) and as the
// "reference" CPU baseline of bench.py.
#include <string.h>
#include <stdio.h>
#include <vector>
#include "Bullet3Common/b3AlignedObjectArray.h"
#include "Bullet3Common/b3Vector3.h"
#include "Bullet3Common/b3Quaternion.h"
#include "Bullet3Common/b3Logging.h"
#include "Bullet3Common/b3Scalar.h"
#define B3_PROFILE(x)
#include "Bullet3Common/shared/b3Int4.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3RigidBodyData.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3Collidable.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3ConvexPolyhedronData.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3ContactConvexConvexSAT.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3UpdateAabbs.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3ConvexUtility.h"
#include "Bullet3Collision/BroadPhaseCollision/shared/b3Aabb.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3MprPenetration.h"
#include "Bullet3Dynamics/shared/b3IntegrateTransforms.h"
#include "Bullet3Dynamics/shared/b3ConvertConstraint4.h"

#include "Bullet3Dynamics/ConstraintSolver/b3PgsJacobiSolver.h"
#include "Bullet3Dynamics/ConstraintSolver/b3Point2PointConstraint.h"
#include "../include/b3b200_types.h"

static_assert(sizeof(b3RigidBodyData) == sizeof(b3b200_rigid_body), "abi");
static_assert(sizeof(b3InertiaData) == sizeof(b3b200_inertia), "abi");
static_assert(sizeof(b3Collidable) == sizeof(b3b200_collidable), "abi");
static_assert(sizeof(b3ConvexPolyhedronData) == sizeof(b3b200_convex_polyhedron), "abi");
static_assert(sizeof(b3GpuFace) == sizeof(b3b200_face), "abi");
static_assert(sizeof(b3Aabb) == sizeof(b3b200_aabb), "abi");
static_assert(sizeof(b3Contact4Data) == sizeof(b3b200_contact4), "abi");
static_assert(sizeof(b3ContactConstraint4) == sizeof(b3b200_constraint4), "abi");
static_assert(sizeof(b3Int4) == sizeof(b3b200_int4), "abi");
static_assert(sizeof(b3Vector3) == sizeof(b3b200_float4), "abi");

template <typename T, typename S>
static void fill(b3AlignedObjectArray<T>& dst, const S* src, int n)
{
	static_assert(sizeof(T) == sizeof(S), "abi");
	dst.resize(n);
	if (n) memcpy(&dst[0], src, sizeof(T) * (size_t)n);
}

extern "C" {

int ref_sizes(int* out, int n)
{
	int s[] = {(int)sizeof(b3RigidBodyData), (int)sizeof(b3InertiaData), (int)sizeof(b3Collidable), (int)sizeof(b3GpuChildShape),
			   (int)sizeof(b3GpuFace), (int)sizeof(b3ConvexPolyhedronData), (int)sizeof(b3Aabb), (int)sizeof(b3Int4), (int)sizeof(b3Contact4Data),
			   (int)sizeof(b3ContactConstraint4)};
	int m = (int)(sizeof(s) / sizeof(s[0]));
	for (int i = 0; i < m && i < n; i++) out[i] = s[i];
	return m;
}

// b3ComputeWorldAabb (shared/b3UpdateAabbs.h:8-33).  The reference reads
// body[bodyId].m_invMass == bodies[2*bodyId] for max.w, so the body array is
// padded to 2n here; max.w is not compared by the tests.
void ref_update_aabbs(const b3b200_rigid_body* bodies, int n, const b3b200_collidable* collidables, const b3b200_aabb* localAabbs, b3b200_aabb* out)
{
	std::vector<b3RigidBodyData> padded(2 * (size_t)n + 1);
	memset(padded.data(), 0, padded.size() * sizeof(b3RigidBodyData));
	memcpy(padded.data(), bodies, sizeof(b3RigidBodyData) * (size_t)n);
	for (int i = 0; i < n; i++)
		b3ComputeWorldAabb(i, padded.data(), (const b3Collidable*)collidables, (const b3Aabb*)localAabbs, (b3Aabb*)out);
}

// predicate b3TestAabbAgainstAabb (shared/b3Aabb.h:45-53) in the loop structure of
// b3GpuSapBroadphase::calculateOverlappingPairsHost (b3GpuSapBroadphase.cpp:862-981)
int ref_brute_force_pairs(const b3b200_aabb* aabbs, const int* smallIdx, int nSmall, const int* largeIdx, int nLarge, b3b200_int4* pairsOut, int maxPairs)
{
	const b3Aabb* A = (const b3Aabb*)aabbs;
	int count = 0;
	for (int i = 0; i < nSmall; i++)
		for (int j = i + 1; j < nSmall; j++)
		{
			const b3Aabb& a = A[smallIdx[i]];
			const b3Aabb& b = A[smallIdx[j]];
			if (b3TestAabbAgainstAabb(a.m_minVec, a.m_maxVec, b.m_minVec, b.m_maxVec))
			{
				int x = a.m_minIndices[3], y = b.m_minIndices[3];
				if (count < maxPairs)
				{
					pairsOut[count].x = x <= y ? x : y;
					pairsOut[count].y = x <= y ? y : x;
					pairsOut[count].z = -1;
					pairsOut[count].w = -1;
				}
				count++;
			}
		}
	for (int i = 0; i < nSmall; i++)
		for (int j = 0; j < nLarge; j++)
		{
			const b3Aabb& a = A[smallIdx[i]];
			const b3Aabb& b = A[largeIdx[j]];
			if (b3TestAabbAgainstAabb(a.m_minVec, a.m_maxVec, b.m_minVec, b.m_maxVec))
			{
				int x = b.m_minIndices[3], y = a.m_minIndices[3];
				if (count < maxPairs)
				{
					pairsOut[count].x = x <= y ? x : y;
					pairsOut[count].y = x <= y ? y : x;
					pairsOut[count].z = -1;
					pairsOut[count].w = -1;
				}
				count++;
			}
		}
	return count;
}

// integrateSingleTransform (Bullet3Dynamics/shared/b3IntegrateTransforms.h:5-55)
void ref_integrate(b3b200_rigid_body* bodies, int n, float dt, float angularDamping, const float* gravity)
{
	b3Vector3 g = b3MakeVector3(gravity[0], gravity[1], gravity[2]);
	for (int i = 0; i < n; i++) integrateSingleTransform((b3RigidBodyData*)bodies, i, dt, angularDamping, g);
}

// b3ContactConvexConvexSAT (shared/b3ContactConvexConvexSAT.h:407-484), the loop of
// b3CpuNarrowPhase::computeContacts (b3CpuNarrowPhase.cpp:53-141).  Clip window is the
// header's own (-1, 0).
int ref_convex_contacts(const b3b200_int4* pairs, int nPairs, const b3b200_rigid_body* bodies, int nBodies, const b3b200_collidable* collidables,
						int nCollidables, const b3b200_convex_polyhedron* convex, int nConvex, const b3b200_float4* vertices, int nVerts,
						const b3b200_float4* uniqueEdges, int nEdges, const b3b200_face* faces, int nFaces, const int* indices, int nIndices,
						b3b200_contact4* out, int maxContacts, int* pairContactIndex)
{
	b3AlignedObjectArray<b3RigidBodyData> rb;
	b3AlignedObjectArray<b3Collidable> col;
	b3AlignedObjectArray<b3ConvexPolyhedronData> cv;
	b3AlignedObjectArray<b3Vector3> vt, ue;
	b3AlignedObjectArray<b3GpuFace> fc;
	b3AlignedObjectArray<int> ix;
	fill(rb, bodies, nBodies);
	fill(col, collidables, nCollidables);
	fill(cv, convex, nConvex);
	fill(vt, vertices, nVerts);
	fill(ue, uniqueEdges, nEdges);
	fill(fc, faces, nFaces);
	fill(ix, indices, nIndices);
	b3AlignedObjectArray<b3Contact4Data> contacts;
	contacts.reserve(maxContacts);
	int numContacts = 0;
	for (int i = 0; i < nPairs; i++)
	{
		int bodyIndexA = pairs[i].x, bodyIndexB = pairs[i].y;
		int collidableIndexA = rb[bodyIndexA].m_collidableIdx, collidableIndexB = rb[bodyIndexB].m_collidableIdx;
		int idx = -1;
		if (col[collidableIndexA].m_shapeType == SHAPE_CONVEX_HULL && col[collidableIndexB].m_shapeType == SHAPE_CONVEX_HULL)
			idx = b3ContactConvexConvexSAT(i, bodyIndexA, bodyIndexB, collidableIndexA, collidableIndexB, rb, col, cv, vt, ue, ix, fc, contacts,
										   numContacts, maxContacts);
		if (pairContactIndex) pairContactIndex[i] = idx;
	}
	for (int i = 0; i < numContacts; i++) memcpy(&out[i], &contacts[i], sizeof(b3Contact4Data));
	return numContacts;
}

// setConstraint4 (Bullet3Dynamics/shared/b3ConvertConstraint4.h:62-148) driven like
// b3Solver::convertToConstraints' host branch (b3Solver.cpp:889-933)
void ref_build_constraints(const b3b200_contact4* contacts, int n, const b3b200_rigid_body* bodies, const b3b200_inertia* inertias, float dt,
						   float positionDrift, float positionConstraintCoeff, b3b200_constraint4* out)
{
	const b3RigidBodyData* gBodies = (const b3RigidBodyData*)bodies;
	const b3InertiaData* gShapes = (const b3InertiaData*)inertias;
	for (int gIdx = 0; gIdx < n; gIdx++)
	{
		b3Contact4Data c;
		memcpy(&c, &contacts[gIdx], sizeof(c));
		int aIdx = abs(c.m_bodyAPtrAndSignBit);
		int bIdx = abs(c.m_bodyBPtrAndSignBit);
		b3ContactConstraint4_t cs;
		memset(&cs, 0, sizeof(cs));
		setConstraint4(gBodies[aIdx].m_pos, gBodies[aIdx].m_linVel, gBodies[aIdx].m_angVel, gBodies[aIdx].m_invMass, gShapes[aIdx].m_initInvInertia,
					   gBodies[bIdx].m_pos, gBodies[bIdx].m_linVel, gBodies[bIdx].m_angVel, gBodies[bIdx].m_invMass, gShapes[bIdx].m_initInvInertia, &c, dt,
					   positionDrift, positionConstraintCoeff, &cs);
		cs.m_batchIdx = c.m_batchIdx;
		memcpy(&out[gIdx], &cs, sizeof(cs));
	}
}

// b3ConvexUtility::initializePolyhedralFeatures (b3ConvexUtility.cpp:26) -> flat tables,
// the way b3GpuNarrowPhase::registerConvexHullShapeInternal lays them out
// (Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.cpp:234-296).  Returns 0 on success.
int ref_build_hull(const float* points, int n, b3b200_float4* vertsOut, int* nVerts, b3b200_face* facesOut, int* nFaces, int* indicesOut, int* nIndices,
				   b3b200_float4* edgesOut, int* nEdges, int cap)
{
	b3AlignedObjectArray<b3Vector3> verts;
	for (int i = 0; i < n; i++) verts.push_back(b3MakeVector3(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
	b3ConvexUtility util;
	if (!util.initializePolyhedralFeatures(&verts[0], verts.size(), true)) return -1;
	if (util.m_vertices.size() > cap || util.m_faces.size() > cap || util.m_uniqueEdges.size() > cap) return -2;
	*nVerts = util.m_vertices.size();
	for (int i = 0; i < util.m_vertices.size(); i++) memcpy(&vertsOut[i], &util.m_vertices[i], 16);
	*nEdges = util.m_uniqueEdges.size();
	for (int i = 0; i < util.m_uniqueEdges.size(); i++) memcpy(&edgesOut[i], &util.m_uniqueEdges[i], 16);
	*nFaces = util.m_faces.size();
	int ni = 0;
	for (int i = 0; i < util.m_faces.size(); i++)
	{
		facesOut[i].plane.x = util.m_faces[i].m_plane[0];
		facesOut[i].plane.y = util.m_faces[i].m_plane[1];
		facesOut[i].plane.z = util.m_faces[i].m_plane[2];
		facesOut[i].plane.w = util.m_faces[i].m_plane[3];
		facesOut[i].indexOffset = ni;
		facesOut[i].numIndices = util.m_faces[i].m_indices.size();
		facesOut[i].pad1 = facesOut[i].pad2 = 0;
		for (int p = 0; p < util.m_faces[i].m_indices.size(); p++)
		{
			if (ni >= cap * 8) return -2;
			indicesOut[ni++] = util.m_faces[i].m_indices[p];
		}
	}
	*nIndices = ni;
	return 0;
}


// The reference's CPU joint path: b3PgsJacobiSolver::solveContacts with b3Point2PointConstraint objects -- what
// b3GpuRigidBodyPipeline::stepSimulation itself calls for b3TypedConstraint joints (b3GpuRigidBodyPipeline.cpp:375-385) and
// what b3GpuPgsConstraintSolver::solveJoints was ported from (same 4 iterations, dt 1/60, ERP 0.2).  Joints are solved
// sequentially in index order.  Only P2P joints (type 3) are taken; returns how many were solved.
int ref_solve_joints_p2p(b3b200_rigid_body* bodies, b3b200_inertia* inertias, int numBodies, const b3b200_generic_constraint* cs, int n)
{
	std::vector<b3TypedConstraint*> joints;
	for (int i = 0; i < n; i++)
	{
		if (cs[i].constraintType != 3 || !(cs[i].flags & 1)) continue;
		b3Point2PointConstraint* p = new b3Point2PointConstraint(cs[i].rbA, cs[i].rbB, b3MakeVector3(cs[i].pivotInA.x, cs[i].pivotInA.y, cs[i].pivotInA.z),
																 b3MakeVector3(cs[i].pivotInB.x, cs[i].pivotInB.y, cs[i].pivotInB.z));
		p->setBreakingImpulseThreshold(cs[i].breakingImpulseThreshold);
		joints.push_back(p);
	}
	b3PgsJacobiSolver solver(true);
	if (!joints.empty()) solver.solveContacts(numBodies, (b3RigidBodyData*)bodies, (b3InertiaData*)inertias, 0, 0, (int)joints.size(), &joints[0]);
	int solved = (int)joints.size();
	for (size_t i = 0; i < joints.size(); i++) delete joints[i];
	return solved;
}

}  // extern "C"

// ------------------------------------------------------------------ the reference's CPU pipeline itself
// b3CpuRigidBodyPipeline (src/Bullet3Dynamics/b3CpuRigidBodyPipeline.cpp) + b3CpuNarrowPhase + b3DynamicBvhBroadphase,
// instantiated and stepped as they are: BASELINE configs[0] ("1,000 unit boxes ... b3CpuRigidBodyPipeline on CPU, 600 steps")
// and the CPU baseline north_star names.  stepSimulation = AABBs + DBVT pairs + contacts + integrate (it has no solver,
// b3CpuRigidBodyPipeline.cpp:75-90).  The per-stage getters exist for the per-step parity test of SURVEY 8(d) config 1.
#include <unistd.h>
#include <fcntl.h>
#include "Bullet3Dynamics/b3CpuRigidBodyPipeline.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3CpuNarrowPhase.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"
#include "Bullet3Collision/BroadPhaseCollision/b3DynamicBvhBroadphase.h"
#include "Bullet3Dynamics/shared/b3Inertia.h"
// layout mirror of the pipeline's private data (b3CpuRigidBodyPipeline.cpp:14-23), to read the world AABBs and to re-seed the bodies
struct b3CpuRigidBodyPipelineInternalData
{
	b3AlignedObjectArray<b3RigidBodyData> m_rigidBodies;
	b3AlignedObjectArray<b3Inertia> m_inertias;
	b3AlignedObjectArray<b3Aabb> m_aabbWorldSpace;
	b3DynamicBvhBroadphase* m_bp;
	b3CpuNarrowPhase* m_np;
	b3Config m_config;
};
struct RefCpuPipe : public b3CpuRigidBodyPipeline
{
	RefCpuPipe(b3CpuNarrowPhase* np, b3DynamicBvhBroadphase* bp, const b3Config& c) : b3CpuRigidBodyPipeline(np, bp, c) {}
	b3CpuRigidBodyPipelineInternalData* data() { return m_data; }
};
struct RefCpu
{
	b3Config cfg;
	b3CpuNarrowPhase* np;
	b3DynamicBvhBroadphase* bp;
	RefCpuPipe* pipe;
};
struct Quiet  // the pipeline printf()s "numPairs=..." every step
{
	int saved, devnull;
	Quiet()
	{
		fflush(stdout);
		saved = dup(1);
		devnull = open("/dev/null", O_WRONLY);
		if (saved >= 0 && devnull >= 0) dup2(devnull, 1);
	}
	~Quiet()
	{
		fflush(stdout);
		if (saved >= 0 && devnull >= 0) dup2(saved, 1);
		if (saved >= 0) close(saved);
		if (devnull >= 0) close(devnull);
	}
};

extern "C" {
void* ref_cpu_create(const b3b200_config* cfg)
{
	RefCpu* r = new RefCpu;
	memcpy(&r->cfg, cfg, sizeof(b3Config));
	r->np = new b3CpuNarrowPhase(r->cfg);
	r->bp = new b3DynamicBvhBroadphase(r->cfg.m_maxConvexBodies);
	r->pipe = new RefCpuPipe(r->np, r->bp, r->cfg);
	return r;
}
void ref_cpu_destroy(void* h)
{
	RefCpu* r = (RefCpu*)h;
	delete r->pipe;
	delete r->np;
	delete r->bp;
	delete r;
}
int ref_cpu_register_convex_points(void* h, const float* pts, int n, const float* scaling) { return ((RefCpu*)h)->np->registerConvexHullShape(pts, 12, n, scaling); }
int ref_cpu_register_instance(void* h, float mass, const float* pos4, const float* orn4, int collidable, int userIndex)
{
	return ((RefCpu*)h)->pipe->registerPhysicsInstance(mass, pos4, orn4, collidable, userIndex);
}
// (b3CpuRigidBodyPipeline::setGravity is declared but never defined: the pipeline integrates with its built-in (0,-9,0), b3CpuRigidBodyPipeline.cpp:365)
int ref_cpu_num_bodies(void* h) { return ((RefCpu*)h)->pipe->getNumBodies(); }
void ref_cpu_step(void* h, float dt, int steps)
{
	Quiet q;
	for (int i = 0; i < steps; i++) ((RefCpu*)h)->pipe->stepSimulation(dt);
}
// one stage: 0 updateAabbWorldSpace, 1 computeOverlappingPairs, 2 computeContactPoints, 3 integrate
void ref_cpu_stage(void* h, int which, float dt)
{
	Quiet q;
	RefCpuPipe* p = ((RefCpu*)h)->pipe;
	if (which == 0) p->updateAabbWorldSpace();
	if (which == 1) p->computeOverlappingPairs();
	if (which == 2) p->computeContactPoints();
	if (which == 3) p->integrate(dt);
}
void ref_cpu_get_bodies(void* h, b3b200_rigid_body* out, int n)
{
	RefCpu* r = (RefCpu*)h;
	const b3RigidBodyData* b = r->pipe->getBodyBuffer();
	for (int i = 0; i < n && i < r->pipe->getNumBodies(); i++) memcpy(&out[i], &b[i], sizeof(b3RigidBodyData));
}
void ref_cpu_set_bodies(void* h, const b3b200_rigid_body* in, int n)
{
	b3CpuRigidBodyPipelineInternalData* d = ((RefCpu*)h)->pipe->data();
	for (int i = 0; i < n && i < d->m_rigidBodies.size(); i++) memcpy(&d->m_rigidBodies[i], &in[i], sizeof(b3RigidBodyData));
}
int ref_cpu_get_aabbs(void* h, b3b200_aabb* out, int n)
{
	b3CpuRigidBodyPipelineInternalData* d = ((RefCpu*)h)->pipe->data();
	for (int i = 0; i < n && i < d->m_aabbWorldSpace.size(); i++) memcpy(&out[i], &d->m_aabbWorldSpace[i], sizeof(b3Aabb));
	return d->m_aabbWorldSpace.size();
}
int ref_cpu_get_pairs(void* h, b3b200_int4* out, int cap)
{
	b3AlignedObjectArray<b3Int4>& p = ((RefCpu*)h)->bp->getOverlappingPairCache()->getOverlappingPairArray();
	for (int i = 0; i < p.size() && i < cap; i++) memcpy(&out[i], &p[i], sizeof(b3Int4));
	return p.size();
}
int ref_cpu_get_contacts(void* h, b3b200_contact4* out, int cap)
{
	const b3AlignedObjectArray<b3Contact4Data>& c = ((RefCpu*)h)->np->getContacts();
	for (int i = 0; i < c.size() && i < cap; i++) memcpy(&out[i], &c[i], sizeof(b3Contact4Data));
	return c.size();
}
// the shape tables b3CpuNarrowPhase built (b3ConvexUtility hulls): 0 collidables, 2 convex polyhedra, 3 vertices, 4 unique edges, 5 faces, 6 indices
int ref_cpu_get_table(void* h, int which, void* dst, int capacity, int* count)
{
	// layout mirror of b3CpuNarrowPhaseInternalData (b3CpuNarrowPhase.cpp:8-24)
	struct NpData
	{
		b3AlignedObjectArray<b3Aabb> m_localShapeAABBCPU;
		b3AlignedObjectArray<b3Collidable> m_collidablesCPU;
		b3AlignedObjectArray<b3ConvexUtility*> m_convexData;
		b3Config m_config;
		b3AlignedObjectArray<b3ConvexPolyhedronData> m_convexPolyhedra;
		b3AlignedObjectArray<b3Vector3> m_uniqueEdges;
		b3AlignedObjectArray<b3Vector3> m_convexVertices;
		b3AlignedObjectArray<int> m_convexIndices;
		b3AlignedObjectArray<b3GpuFace> m_convexFaces;
	};
	const NpData* d = (const NpData*)((RefCpu*)h)->np->getInternalData();
	const void* src = 0;
	int n = 0, sz = 0;
	switch (which)
	{
		case 0: n = d->m_collidablesCPU.size(); sz = sizeof(b3Collidable); src = n ? &d->m_collidablesCPU[0] : 0; break;
		case 1: n = d->m_localShapeAABBCPU.size(); sz = sizeof(b3Aabb); src = n ? &d->m_localShapeAABBCPU[0] : 0; break;
		case 2: n = d->m_convexPolyhedra.size(); sz = sizeof(b3ConvexPolyhedronData); src = n ? &d->m_convexPolyhedra[0] : 0; break;
		case 3: n = d->m_convexVertices.size(); sz = 16; src = n ? &d->m_convexVertices[0] : 0; break;
		case 4: n = d->m_uniqueEdges.size(); sz = 16; src = n ? &d->m_uniqueEdges[0] : 0; break;
		case 5: n = d->m_convexFaces.size(); sz = sizeof(b3GpuFace); src = n ? &d->m_convexFaces[0] : 0; break;
		case 6: n = d->m_convexIndices.size(); sz = 4; src = n ? &d->m_convexIndices[0] : 0; break;
		default: return -1;
	}
	*count = n;
	const int m = n < capacity ? n : capacity;
	if (dst && m > 0 && src) memcpy(dst, src, (size_t)sz * m);
	return 0;
}

// mprPenetrationKernel (src/Bullet3OpenCL/NarrowphaseCollision/kernels/mpr.cl:14-89) as a host loop around the reference's own
// b3MprPenetration (shared/b3MprPenetration.h:825-888): the kernel body is 30 lines of bookkeeping, restated here because the .cl
// file cannot be compiled by g++; everything numeric is the header's.  res[i] = what b3MprPenetration returned (-2: skipped).
int ref_mpr_kernel(b3b200_int4* pairs, int numPairs, const b3b200_rigid_body* bodies, const b3b200_collidable* collidables, const b3b200_convex_polyhedron* convex,
				   const b3b200_float4* vertices, b3b200_float4* sepNormals, int* hasSepAxis, b3b200_contact4* contactsOut, int contactCapacity, int* numContacts,
				   b3b200_mpr_result* res)
{
	const b3RigidBodyData* rb = (const b3RigidBodyData*)bodies;
	const b3Collidable* col = (const b3Collidable*)collidables;
	const b3ConvexPolyhedronData* cv = (const b3ConvexPolyhedronData*)convex;
	const b3Vector3* vtx = (const b3Vector3*)vertices;
	b3Vector3* sep = (b3Vector3*)sepNormals;
	b3Contact4Data* out = (b3Contact4Data*)contactsOut;
	for (int i = 0; i < numPairs; i++)
	{
		const int bodyIndexA = pairs[i].x, bodyIndexB = pairs[i].y;
		if (res)
		{
			res[i].result = -2;
			res[i].depth = 0.f;
		}
		if (rb[bodyIndexA].m_invMass == 0 && rb[bodyIndexB].m_invMass == 0) continue;
		if (col[rb[bodyIndexA].m_collidableIdx].m_shapeType != SHAPE_CONVEX_HULL || col[rb[bodyIndexB].m_collidableIdx].m_shapeType != SHAPE_CONVEX_HULL) continue;
		float depthOut = 0.f;
		b3Float4 dirOut = b3MakeFloat4(0, 0, 0, 0), posOut = b3MakeFloat4(0, 0, 0, 0);
		int r = b3MprPenetration(i, bodyIndexA, bodyIndexB, rb, cv, col, vtx, sep, hasSepAxis, &depthOut, &dirOut, &posOut);
		if (res)
		{
			res[i].result = r;
			res[i].depth = depthOut;
			for (int k = 0; k < 3; k++)
			{
				res[i].dir[k] = dirOut[k];
				res[i].pos[k] = posOut[k];
			}
		}
		if (r == 0)
		{
			int dstIdx = (*numContacts)++;
			if (dstIdx < contactCapacity)
			{
				pairs[i].z = dstIdx;
				b3Contact4Data* c = out + dstIdx;
				c->m_worldNormalOnB = -dirOut;
				c->m_restituitionCoeffCmp = (0.f * 0xffff);
				c->m_frictionCoeffCmp = (0.7f * 0xffff);
				c->m_batchIdx = i;
				c->m_bodyAPtrAndSignBit = rb[bodyIndexA].m_invMass == 0 ? -bodyIndexA : bodyIndexA;
				c->m_bodyBPtrAndSignBit = rb[bodyIndexB].m_invMass == 0 ? -bodyIndexB : bodyIndexB;
				c->m_childIndexA = -1;
				c->m_childIndexB = -1;
				posOut.w = -depthOut;
				c->m_worldPosB[0] = posOut;
				c->m_worldNormalOnB.w = 1.f;  // GET_NPOINTS(*c) = 1
			}
		}
	}
	return 0;
}
}
