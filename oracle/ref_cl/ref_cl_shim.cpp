// ref_cl_shim.cpp -- extern "C" access to the reference's own HOST TWINS inside src/Bullet3OpenCL
// (compiled unmodified, -DB3_USE_CLEW, against the host-memory fake OpenCL of fake_cl.cpp).
// TEST INFRASTRUCTURE ONLY; nothing here restates an algorithm.
//   refcl_pairs_host        b3GpuSapBroadphase::calculateOverlappingPairsHost   (b3GpuSapBroadphase.cpp:862-981)
//   refcl_pgs_solve         b3Solver::convertToConstraints (gConvertConstraintOnCpu) + solveContactConstraintHost
//                           (b3Solver.cpp:889-933, 468-637)
//   refcl_radix_sort / scan / bound_search   the executeHost twins of ParallelPrimitives
//   refcl_cast_rays_host    b3GpuRaycast::castRaysHost (b3GpuRaycast.cpp:173-246) on the shapes/bodies of a refcl_np handle
//   refcl_np_*              b3GpuNarrowPhase + GpuSatCollision built with -DCHECK_ON_HOST: shape registration
//                           (incl. the compound / mesh BVH builders) and the host contact loop
//                           (b3ConvexHullContact.cpp:2595-2748)
#include <string.h>
#include <stdio.h>
#include "Bullet3OpenCL/Initialize/b3OpenCLInclude.h"
#include "Bullet3OpenCL/ParallelPrimitives/b3OpenCLArray.h"
#include "Bullet3OpenCL/ParallelPrimitives/b3RadixSort32CL.h"
#include "Bullet3OpenCL/ParallelPrimitives/b3PrefixScanCL.h"
#include "Bullet3OpenCL/ParallelPrimitives/b3BoundSearchCL.h"
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuSapBroadphase.h"
#include "Bullet3OpenCL/RigidBody/b3Solver.h"
#include "Bullet3OpenCL/RigidBody/b3GpuJacobiContactSolver.h"
#include <unistd.h>
#include <fcntl.h>
#include "Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.h"
#include "Bullet3OpenCL/RigidBody/b3GpuNarrowPhaseInternalData.h"
#include "Bullet3OpenCL/Raycast/b3GpuRaycast.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Contact4.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3RigidBodyData.h"
#include "Bullet3Dynamics/shared/b3ContactConstraint4.h"
#include "../../include/b3b200_types.h"

extern bool bvhTraversalKernelGPU, findConcaveSeparatingAxisKernelGPU, clipConcaveFacesAndFindContactsCPU, reduceConcaveContactsOnGPU;

extern "C" void b3ref_cl_init();
extern bool gConvertConstraintOnCpu;  // b3Solver.cpp:20
extern int g_fakeClLaunchIsNoop, g_fakeClLaunches;

static cl_context CTX = 0;
static cl_device_id DEV = 0;
static cl_command_queue Q = 0;
static void init()
{
	static bool done = false;
	if (!done) b3ref_cl_init();
	done = true;
}

extern "C" {

int refcl_pairs_host(const b3b200_aabb* aabbs, int n, const unsigned char* isLarge, b3b200_int4* pairsOut, int maxPairs)
{
	init();
	b3GpuSapBroadphase bp(CTX, DEV, Q);
	for (int i = 0; i < n; i++)
	{
		b3Vector3 mn = b3MakeVector3(aabbs[i].min[0], aabbs[i].min[1], aabbs[i].min[2]);
		b3Vector3 mx = b3MakeVector3(aabbs[i].max[0], aabbs[i].max[1], aabbs[i].max[2]);
		if (isLarge[i])
			bp.createLargeProxy(mn, mx, aabbs[i].minIndices[3], 1, 1);
		else
			bp.createProxy(mn, mx, aabbs[i].minIndices[3], 1, 1);
	}
	bp.writeAabbsToGpu();
	bp.calculateOverlappingPairsHost(maxPairs);
	b3AlignedObjectArray<b3Int4> host;
	bp.getOverlappingPairsGPU().copyToHost(host);
	for (int i = 0; i < host.size() && i < maxPairs; i++) memcpy(&pairsOut[i], &host[i], 16);
	return host.size();
}

// contacts must be sorted by batch; batchSizes[b] = number of contacts in batch b
int refcl_pgs_solve(const b3b200_contact4* contacts, int n, const int* batchSizes, int numBatches, b3b200_rigid_body* bodies, int numBodies,
					const b3b200_inertia* inertias, int iterations, float dt, b3b200_constraint4* constraintsOut)
{
	init();
	if (numBatches >= B3_MAX_NUM_BATCHES) return -1;
	b3Solver solver(CTX, DEV, Q, n > 512 ? n : 512);
	b3OpenCLArray<b3RigidBodyData> bodyBuf(CTX, Q);
	b3OpenCLArray<b3InertiaData> shapeBuf(CTX, Q);
	b3OpenCLArray<b3Contact4> contactBuf(CTX, Q);
	b3OpenCLArray<b3GpuConstraint4> constraintBuf(CTX, Q);
	bodyBuf.resize(numBodies);
	shapeBuf.resize(numBodies);
	contactBuf.resize(n);
	bodyBuf.copyFromHostPointer((const b3RigidBodyData*)bodies, numBodies, 0, true);
	shapeBuf.copyFromHostPointer((const b3InertiaData*)inertias, numBodies, 0, true);
	contactBuf.copyFromHostPointer((const b3Contact4*)contacts, n, 0, true);
	b3SolverBase::ConstraintCfg cfg(dt);
	gConvertConstraintOnCpu = true;
	solver.convertToConstraints(&bodyBuf, &shapeBuf, &contactBuf, &constraintBuf, 0, n, cfg);
	if (constraintsOut) constraintBuf.copyToHostPointer((b3GpuConstraint4*)constraintsOut, n, 0, true);
	// one spatial cell (index 0) holding every constraint; its batches are the global batches
	b3AlignedObjectArray<unsigned int> counts, offsets;
	counts.resize(B3_SOLVER_N_CELLS);
	offsets.resize(B3_SOLVER_N_CELLS);
	for (int i = 0; i < B3_SOLVER_N_CELLS; i++)
	{
		counts[i] = 0;
		offsets[i] = 0;
	}
	counts[0] = n;
	solver.m_numConstraints->copyFromHost(counts);
	solver.m_offsets->copyFromHost(offsets);
	b3AlignedObjectArray<int> bs;
	bs.resize(B3_SOLVER_N_CELLS * B3_MAX_NUM_BATCHES);
	for (int i = 0; i < bs.size(); i++) bs[i] = 0;
	for (int b = 0; b < numBatches; b++) bs[b] = batchSizes[b];
	solver.m_nIterations = iterations;
	solver.solveContactConstraintHost(&bodyBuf, &shapeBuf, &constraintBuf, 0, n, numBatches, &bs);
	bodyBuf.copyToHostPointer((b3RigidBodyData*)bodies, numBodies, 0, true);
	return 0;
}

void refcl_radix_sort(b3b200_sort_data* data, int n)
{
	init();
	b3RadixSort32CL sorter(CTX, DEV, Q);
	b3AlignedObjectArray<b3SortData> a;
	a.resize(n);
	if (n) memcpy(&a[0], data, sizeof(b3SortData) * (size_t)n);
	sorter.executeHost(a);
	if (n) memcpy(data, &a[0], sizeof(b3SortData) * (size_t)n);
}

void refcl_prefix_scan(const unsigned int* src, unsigned int* dst, int n, unsigned int* sum)
{
	init();
	b3PrefixScanCL scan(CTX, DEV, Q, n + 16);
	b3AlignedObjectArray<unsigned int> a, b;
	a.resize(n);
	b.resize(n);
	if (n) memcpy(&a[0], src, 4 * (size_t)n);
	scan.executeHost(a, b, n, sum);
	if (n) memcpy(dst, &b[0], 4 * (size_t)n);
}

void refcl_bound_search_count(const b3b200_sort_data* sorted, int n, unsigned int* counts, int numBuckets)
{
	init();
	b3BoundSearchCL search(CTX, DEV, Q, numBuckets);
	b3AlignedObjectArray<b3SortData> a;
	b3AlignedObjectArray<unsigned int> c;
	a.resize(n);
	c.resize(numBuckets);
	if (n) memcpy(&a[0], sorted, sizeof(b3SortData) * (size_t)n);
	for (int i = 0; i < numBuckets; i++) c[i] = 0;
	search.executeHost(a, n, c, numBuckets, b3BoundSearchCL::COUNT);
	memcpy(counts, &c[0], 4 * (size_t)numBuckets);
}

// b3BoundSearchCL::executeHost with BOUND_LOWER (0) / BOUND_UPPER (1) / COUNT (2); dst is in/out
void refcl_bound_search(const b3b200_sort_data* sorted, int n, unsigned int* dst, int numBuckets, int option)
{
	init();
	b3BoundSearchCL search(CTX, DEV, Q, numBuckets);
	b3AlignedObjectArray<b3SortData> a;
	b3AlignedObjectArray<unsigned int> c;
	a.resize(n);
	c.resize(numBuckets);
	if (n) memcpy(&a[0], sorted, sizeof(b3SortData) * (size_t)n);
	memcpy(&c[0], dst, 4 * (size_t)numBuckets);
	search.executeHost(a, n, c, numBuckets, option == 0 ? b3BoundSearchCL::BOUND_LOWER : (option == 1 ? b3BoundSearchCL::BOUND_UPPER : b3BoundSearchCL::COUNT));
	memcpy(dst, &c[0], 4 * (size_t)numBuckets);
}

// b3GpuJacobiContactSolver::solveGroupHost (b3GpuJacobiContactSolver.cpp:462-697): the reference's host twin of the mass-splitting
// Jacobi solver; bodies are updated in place
void refcl_jacobi_solve_host(b3b200_contact4* contacts, int n, b3b200_rigid_body* bodies, int numBodies, b3b200_inertia* inertias, int staticIdx, int iterations,
							 float dt, float positionDrift, float positionConstraintCoeff)
{
	init();
	b3GpuJacobiContactSolver solver(CTX, DEV, Q, n > 512 ? n : 512);
	b3JacobiSolverInfo info;
	info.m_fixedBodyIndex = staticIdx;
	info.m_deltaTime = dt;
	info.m_positionDrift = positionDrift;
	info.m_positionConstraintCoeff = positionConstraintCoeff;
	info.m_numIterations = iterations;
	fflush(stdout);
	const int saved = dup(1), devnull = open("/dev/null", O_WRONLY);  // it printf()s totalNumSplitBodies
	if (saved >= 0 && devnull >= 0) dup2(devnull, 1);
	solver.solveGroupHost((b3RigidBodyData*)bodies, (b3InertiaData*)inertias, numBodies, (b3Contact4*)contacts, n, info);
	fflush(stdout);
	if (saved >= 0 && devnull >= 0) dup2(saved, 1);
	if (saved >= 0) close(saved);
	if (devnull >= 0) close(devnull);
}

// ------------------------------------------------------------------ narrowphase (CHECK_ON_HOST build)
struct RefNp
{
	b3GpuNarrowPhase* np;
	b3Config cfg;
};

void* refcl_np_create(const b3b200_config* cfg)
{
	init();
	RefNp* r = new RefNp;
	memcpy(&r->cfg, cfg, sizeof(b3Config));
	r->np = new b3GpuNarrowPhase(CTX, DEV, Q, r->cfg);
	return r;
}
void refcl_np_destroy(void* h)
{
	RefNp* r = (RefNp*)h;
	delete r->np;
	delete r;
}
int refcl_np_register_convex_points(void* h, const float* pts, int n, const float* scaling)
{
	return ((RefNp*)h)->np->registerConvexHullShape(pts, 12, n, scaling);
}
int refcl_np_register_plane(void* h, const float* normal, float c) { return ((RefNp*)h)->np->registerPlaneShape(b3MakeVector3(normal[0], normal[1], normal[2]), c); }
int refcl_np_register_sphere(void* h, float radius) { return ((RefNp*)h)->np->registerSphereShape(radius); }
int refcl_np_register_compound(void* h, const b3b200_child_shape* children, int n)
{
	b3AlignedObjectArray<b3GpuChildShape> ch;
	ch.resize(n);
	memcpy(&ch[0], children, sizeof(b3GpuChildShape) * (size_t)n);
	return ((RefNp*)h)->np->registerCompoundShape(&ch);
}
// route the concave stages through their host twins (file-scope switches of b3ConvexHullContact.cpp:20-24)
void refcl_concave_host_twins(int on)
{
	bvhTraversalKernelGPU = !on;
	findConcaveSeparatingAxisKernelGPU = !on;
	clipConcaveFacesAndFindContactsCPU = on != 0;
	reduceConcaveContactsOnGPU = !on;
}
int refcl_np_register_concave(void* h, const float* verts, int nv, const int* idx, int ni, const float* scaling)
{
	b3AlignedObjectArray<b3Vector3> v;
	b3AlignedObjectArray<int> i;
	for (int k = 0; k < nv; k++) v.push_back(b3MakeVector3(verts[3 * k], verts[3 * k + 1], verts[3 * k + 2]));
	for (int k = 0; k < ni; k++) i.push_back(idx[k]);
	return ((RefNp*)h)->np->registerConcaveMesh(&v, &i, scaling);
}
int refcl_np_register_body(void* h, int collidable, float mass, const float* pos, const float* orn, const float* aabbMin, const float* aabbMax)
{
	return ((RefNp*)h)->np->registerRigidBody(collidable, mass, pos, orn, aabbMin, aabbMax, false);
}
// host contact loop on caller-supplied pairs and world AABBs; returns the contact count
int refcl_np_compute_contacts(void* h, const b3b200_rigid_body* bodies, int numBodies, const b3b200_int4* pairs, int numPairs, const b3b200_aabb* aabbsWS,
							  b3b200_contact4* out, int maxContacts, b3b200_int4* pairsOut)
{
	RefNp* r = (RefNp*)h;
	b3GpuNarrowPhaseInternalData* d = r->np->getInternalData();
	// take the caller's body state
	for (int i = 0; i < numBodies && i < d->m_bodyBufferCPU->size(); i++) memcpy(&d->m_bodyBufferCPU->at(i), &bodies[i], sizeof(b3RigidBodyData));
	r->np->writeAllBodiesToGpu();
	b3OpenCLArray<b3Int4> pairBuf(CTX, Q);
	b3OpenCLArray<b3SapAabb> aabbBuf(CTX, Q);
	pairBuf.resize(numPairs);
	aabbBuf.resize(numBodies);
	pairBuf.copyFromHostPointer((const b3Int4*)pairs, numPairs, 0, true);
	aabbBuf.copyFromHostPointer((const b3SapAabb*)aabbsWS, numBodies, 0, true);
	g_fakeClLaunchIsNoop = 1;
	r->np->computeContacts(pairBuf.getBufferCL(), numPairs, aabbBuf.getBufferCL(), numBodies);
	g_fakeClLaunchIsNoop = 0;
	int n = r->np->getNumContactsGpu();
	b3AlignedObjectArray<b3Contact4> host;
	d->m_pBufContactBuffersGPU[d->m_currentContactBuffer]->copyToHost(host);
	for (int i = 0; i < n && i < maxContacts; i++) memcpy(&out[i], &host[i], sizeof(b3Contact4));
	if (pairsOut) pairBuf.copyToHostPointer((b3Int4*)pairsOut, numPairs, 0, true);
	return n;
}
// flat shape tables as the reference built them (b3GpuNarrowPhaseInternalData.h:24-86)
int refcl_np_get_table(void* h, int which, void* dst, int capacity, int* count)
{
	b3GpuNarrowPhaseInternalData* d = ((RefNp*)h)->np->getInternalData();
	const void* src = 0;
	int n = 0, sz = 0;
	switch (which)
	{
		case 0: src = d->m_collidablesCPU.size() ? &d->m_collidablesCPU[0] : 0; n = d->m_collidablesCPU.size(); sz = sizeof(b3Collidable); break;
		case 1: src = d->m_localShapeAABBCPU->size() ? &d->m_localShapeAABBCPU->at(0) : 0; n = d->m_localShapeAABBCPU->size(); sz = sizeof(b3SapAabb); break;
		case 2: src = d->m_convexPolyhedra.size() ? &d->m_convexPolyhedra[0] : 0; n = d->m_convexPolyhedra.size(); sz = sizeof(b3ConvexPolyhedronData); break;
		case 3: src = d->m_convexVertices.size() ? &d->m_convexVertices[0] : 0; n = d->m_convexVertices.size(); sz = 16; break;
		case 4: src = d->m_uniqueEdges.size() ? &d->m_uniqueEdges[0] : 0; n = d->m_uniqueEdges.size(); sz = 16; break;
		case 5: src = d->m_convexFaces.size() ? &d->m_convexFaces[0] : 0; n = d->m_convexFaces.size(); sz = sizeof(b3GpuFace); break;
		case 6: src = d->m_convexIndices.size() ? &d->m_convexIndices[0] : 0; n = d->m_convexIndices.size(); sz = 4; break;
		case 7: src = d->m_cpuChildShapes.size() ? &d->m_cpuChildShapes[0] : 0; n = d->m_cpuChildShapes.size(); sz = sizeof(b3GpuChildShape); break;
		case 8: src = d->m_bvhInfoCPU.size() ? &d->m_bvhInfoCPU[0] : 0; n = d->m_bvhInfoCPU.size(); sz = sizeof(b3BvhInfo); break;
		case 9: src = d->m_treeNodesCPU.size() ? &d->m_treeNodesCPU[0] : 0; n = d->m_treeNodesCPU.size(); sz = sizeof(b3QuantizedBvhNode); break;
		case 10: src = d->m_subTreesCPU.size() ? &d->m_subTreesCPU[0] : 0; n = d->m_subTreesCPU.size(); sz = sizeof(b3BvhSubtreeInfo); break;
		case 11: src = d->m_bodyBufferCPU->size() ? &d->m_bodyBufferCPU->at(0) : 0; n = ((RefNp*)h)->np->getNumRigidBodies(); sz = sizeof(b3RigidBodyData); break;
		case 12: src = d->m_inertiaBufferCPU->size() ? &d->m_inertiaBufferCPU->at(0) : 0; n = ((RefNp*)h)->np->getNumRigidBodies(); sz = sizeof(b3InertiaData); break;
		default: return -1;
	}
	*count = n;
	int m = n < capacity ? n : capacity;
	if (dst && m > 0 && src) memcpy(dst, src, (size_t)sz * m);
	return 0;
}
// b3GpuRaycast::castRaysHost over the bodies and shape tables of this narrowphase (the caller's body state is taken first)
void refcl_cast_rays_host(void* h, const b3b200_rigid_body* bodies, int numBodies, const b3b200_ray_info* rays, int numRays, b3b200_ray_hit* hits)
{
	RefNp* r = (RefNp*)h;
	b3GpuNarrowPhaseInternalData* d = r->np->getInternalData();
	for (int i = 0; i < numBodies && i < d->m_bodyBufferCPU->size(); i++) memcpy(&d->m_bodyBufferCPU->at(i), &bodies[i], sizeof(b3RigidBodyData));
	b3AlignedObjectArray<b3RayInfo> in;
	b3AlignedObjectArray<b3RayHit> out;
	in.resize(numRays);
	out.resize(numRays);
	if (numRays) memcpy(&in[0], rays, sizeof(b3RayInfo) * (size_t)numRays);
	if (numRays) memcpy(&out[0], hits, sizeof(b3RayHit) * (size_t)numRays);
	static b3GpuRaycast* rc = 0;
	if (!rc) rc = new b3GpuRaycast(CTX, DEV, Q);
	rc->castRaysHost(in, out, numBodies, &d->m_bodyBufferCPU->at(0), d->m_collidablesCPU.size(), &d->m_collidablesCPU[0], d);
	if (numRays) memcpy(hits, &out[0], sizeof(b3RayHit) * (size_t)numRays);
}
}

// ------------------------------------------------------------------ the whole step: b3GpuRigidBodyPipeline::stepSimulation
// (b3GpuRigidBodyPipeline.cpp:221-463), UNMODIFIED, with every host-twin switch of the reference turned on, so that each
// stage runs the reference's own CPU code under the fake OpenCL:
//   world AABBs   gCalcWorldSpaceAabbOnCpu            (b3GpuRigidBodyPipeline.cpp:508-533)
//   pairs         gUseDbvt -> b3DynamicBvhBroadphase  (:231-251; the brute-force calculateOverlappingPairsHost is O(N^2))
//   contacts      -DCHECK_ON_HOST contact loop + the concave host twins (b3ConvexHullContact.cpp:2595-2748, 20-24)
//   solver        gCpuSortContactsDeterminism, gCpuSetSortData, gCpuRadixSort, gUseScanHost, gReorderContactsOnCpu,
//                 gUseCpuCopyConstraints, gCpuBatchContacts, gConvertConstraintOnCpu, gCpuSolveConstraint
//                 (b3GpuPgsContactSolver.cpp:568-1103 -> b3Solver::solveContactConstraintHost, b3Solver.cpp:468-637)
//   integrate     gIntegrateOnCpu                     (b3GpuRigidBodyPipeline.cpp:471-486)
// This is what bench.py --impl reference times.  Single-threaded by construction.
#include <unistd.h>
#include <fcntl.h>
#include "Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.h"
#include "Bullet3OpenCL/RigidBody/b3GpuPgsContactSolver.h"
#include "Bullet3Collision/BroadPhaseCollision/b3DynamicBvhBroadphase.h"
extern bool gUseDbvt, gCalcWorldSpaceAabbOnCpu, gIntegrateOnCpu, gClearPairsOnGpu, gUseJacobi;
extern bool gCpuBatchContacts, gCpuSolveConstraint, gCpuRadixSort, gCpuSetSortData, gCpuSortContactsDeterminism, gUseCpuCopyConstraints, gUseScanHost,
	gReorderContactsOnCpu, gUseLargeBatches;

// the reference's own B3_PROFILE zones (b3Logging.h:21-63), summed per zone name: where its step spends the time
#include <chrono>
#include <map>
#include <string>
#include <vector>
#include "Bullet3Common/b3Logging.h"
static std::map<std::string, double> g_zoneSeconds;
static std::vector<std::pair<const char*, std::chrono::steady_clock::time_point> > g_zoneStack;
static void zoneEnter(const char* name) { g_zoneStack.push_back(std::make_pair(name, std::chrono::steady_clock::now())); }
static void zoneLeave()
{
	if (g_zoneStack.empty()) return;
	g_zoneSeconds[g_zoneStack.back().first] += std::chrono::duration<double>(std::chrono::steady_clock::now() - g_zoneStack.back().second).count();
	g_zoneStack.pop_back();
}

struct RefPipeline
{
	RefNp np;  // first member: a RefPipeline* is also a valid refcl_np_* handle for the shape registration calls
	b3GpuSapBroadphase* sap;
	b3DynamicBvhBroadphase* dbvt;
	b3GpuRigidBodyPipeline* pipe;
};

extern "C" {
void* refcl_pipeline_create(const b3b200_config* cfg)
{
	init();
	RefPipeline* r = new RefPipeline;
	memcpy(&r->np.cfg, cfg, sizeof(b3Config));
	r->np.np = new b3GpuNarrowPhase(CTX, DEV, Q, r->np.cfg);
	r->sap = new b3GpuSapBroadphase(CTX, DEV, Q);
	r->dbvt = new b3DynamicBvhBroadphase(r->np.cfg.m_maxConvexBodies);
	r->pipe = new b3GpuRigidBodyPipeline(CTX, DEV, Q, r->np.np, r->sap, r->dbvt, r->np.cfg);
	gUseDbvt = true;
	gCalcWorldSpaceAabbOnCpu = true;
	gIntegrateOnCpu = true;
	gClearPairsOnGpu = false;
	gUseJacobi = false;
	gCpuBatchContacts = gCpuSolveConstraint = gCpuRadixSort = gCpuSetSortData = gCpuSortContactsDeterminism = true;
	gUseCpuCopyConstraints = gUseScanHost = gReorderContactsOnCpu = true;
	gConvertConstraintOnCpu = true;
	gUseLargeBatches = false;
	refcl_concave_host_twins(1);
	b3SetCustomEnterProfileZoneFunc(zoneEnter);
	b3SetCustomLeaveProfileZoneFunc(zoneLeave);
	return r;
}
void refcl_pipeline_destroy(void* h)
{
	RefPipeline* r = (RefPipeline*)h;
	delete r->pipe;
	delete r->np.np;
	delete r->sap;
	delete r->dbvt;
	delete r;
}
int refcl_pipeline_register_instance(void* h, float mass, const float* pos4, const float* orn4, int collidable, int userIndex)
{
	return ((RefPipeline*)h)->pipe->registerPhysicsInstance(mass, pos4, orn4, collidable, userIndex, false);
}
// writeAllInstancesToGpu (GpuRigidBodyDemo.cpp:148-150)
void refcl_pipeline_upload(void* h)
{
	RefPipeline* r = (RefPipeline*)h;
	r->np.np->writeAllBodiesToGpu();
	r->sap->writeAabbsToGpu();
	r->pipe->writeAllInstancesToGpu();
}
// overwrite the body state (poses, velocities) with the caller's
void refcl_pipeline_set_bodies(void* h, const b3b200_rigid_body* bodies, int n)
{
	RefPipeline* r = (RefPipeline*)h;
	b3GpuNarrowPhaseInternalData* d = r->np.np->getInternalData();
	for (int i = 0; i < n && i < d->m_bodyBufferCPU->size(); i++) memcpy(&d->m_bodyBufferCPU->at(i), &bodies[i], sizeof(b3RigidBodyData));
	r->np.np->writeAllBodiesToGpu();
}
void refcl_pipeline_get_bodies(void* h, b3b200_rigid_body* bodies, int n)
{
	RefPipeline* r = (RefPipeline*)h;
	r->np.np->readbackAllBodiesToCpu();
	b3GpuNarrowPhaseInternalData* d = r->np.np->getInternalData();
	for (int i = 0; i < n && i < d->m_bodyBufferCPU->size(); i++) memcpy(&bodies[i], &d->m_bodyBufferCPU->at(i), sizeof(b3RigidBodyData));
}
// `steps` x stepSimulation(dt).  The contact solver runs the 4 iterations b3GpuPgsContactSolver::solveContacts hard-codes
// (b3GpuPgsContactSolver.cpp:1049-1051: `int numIter = 4`; b3Config has no iteration count) -- `iterations` is ignored.
// out3 = {broadphase pairs, contacts, device kernel launches that were skipped} of the last step.
int refcl_pipeline_step(void* h, float dt, int steps, int iterations, int* out3)
{
	RefPipeline* r = (RefPipeline*)h;
	// (the narrowphase still enqueues its device kernels after the CHECK_ON_HOST loop, see fake_cl.cpp)
	g_fakeClLaunchIsNoop = 1;
	g_fakeClLaunches = 0;
	(void)iterations;
	// the host twins printf diagnostics every step ("maxNumAabbChecks=..."): keep them off the caller's stdout
	fflush(stdout);
	const int saved = dup(1), devnull = open("/dev/null", O_WRONLY);
	if (saved >= 0 && devnull >= 0) dup2(devnull, 1);
	for (int i = 0; i < steps; i++) r->pipe->stepSimulation(dt);
	fflush(stdout);
	if (saved >= 0 && devnull >= 0) dup2(saved, 1);
	if (saved >= 0) close(saved);
	if (devnull >= 0) close(devnull);
	g_fakeClLaunchIsNoop = 0;
	if (out3)
	{
		out3[0] = r->dbvt->getOverlappingPairCache()->getNumOverlappingPairs();
		out3[1] = r->np.np->getNumContactsGpu();
		out3[2] = g_fakeClLaunches;
	}
	return 0;
}
// seconds per B3_PROFILE zone since the last call, as "name=seconds;..." (inclusive times; zones nest)
int refcl_profile_zones(char* dst, int cap)
{
	std::string out;
	for (std::map<std::string, double>::iterator it = g_zoneSeconds.begin(); it != g_zoneSeconds.end(); ++it)
	{
		char buf[256];
		snprintf(buf, sizeof(buf), "%s=%.6f;", it->first.c_str(), it->second);
		out += buf;
	}
	g_zoneSeconds.clear();
	if (dst && cap > 0)
	{
		strncpy(dst, out.c_str(), (size_t)cap - 1);
		dst[cap - 1] = 0;
	}
	return (int)out.size();
}
}
