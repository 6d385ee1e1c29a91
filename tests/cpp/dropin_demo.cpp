// dropin_demo.cpp -- caller code written against the reference's class API
// (examples/OpenCL/rigidbody/GpuRigidBodyDemo.cpp:110-154, 225 and examples/OpenCL/broadphase/PairBench.cpp:208-379),
// compiled unchanged against the B200 drop-in headers.  Prints "DROPIN OK" on success.
#include <stdio.h>
#include <math.h>
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuSapBroadphase.h"
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuGridBroadphase.h"
#include "Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.h"
#include "Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.h"
#include "Bullet3OpenCL/RigidBody/b3GpuPgsContactSolver.h"
#include "Bullet3OpenCL/RigidBody/b3GpuJacobiContactSolver.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3RigidBodyData.h"

#include "Bullet3Collision/NarrowPhaseCollision/shared/b3Collidable.h"

static const float cube[8 * 3] = {-1, -1, -1, -1, -1, 1, -1, 1, -1, -1, 1, 1, 1, -1, -1, 1, -1, 1, 1, 1, -1, 1, 1, 1};

// GpuConcaveScene / GpuCompoundScene-style caller code (examples/OpenCL/rigidbody/ConcaveScene.cpp:150-260,
// GpuCompoundScene.cpp:60-160): a trimesh ground (body 0), three-box compounds, hulls and spheres dropped on it.
static bool meshScene(bool useUniformGrid)
{
	cl_context ctx = 0;
	cl_device_id dev = 0;
	cl_command_queue q = 0;
	b3Config config;
	b3GpuNarrowPhase* np = new b3GpuNarrowPhase(ctx, dev, q, config);
	b3GpuBroadphaseInterface* bp = useUniformGrid ? (b3GpuBroadphaseInterface*)new b3GpuGridBroadphase(ctx, dev, q) : (b3GpuBroadphaseInterface*)new b3GpuSapBroadphase(ctx, dev, q);
	b3GpuRigidBodyPipeline* pipe = new b3GpuRigidBodyPipeline(ctx, dev, q, np, bp, 0, config);
	const int Q = 24;
	b3AlignedObjectArray<b3Vector3> verts;
	b3AlignedObjectArray<int> idx;
	for (int i = 0; i <= Q; i++)
		for (int k = 0; k <= Q; k++)
		{
			float x = (float)i - 0.5f * Q, z = (float)k - 0.5f * Q;
			verts.push_back(b3MakeVector3(x, 0.8f * sinf(0.4f * x) * cosf(0.4f * z), z));
		}
	for (int i = 0; i < Q; i++)
		for (int k = 0; k < Q; k++)
		{
			int v00 = i * (Q + 1) + k, v01 = v00 + 1, v10 = v00 + Q + 1, v11 = v10 + 1;
			int t[6] = {v00, v01, v11, v00, v11, v10};
			for (int j = 0; j < 6; j++) idx.push_back(t[j]);
		}
	float one[4] = {1, 1, 1, 1}, orn[4] = {0, 0, 0, 1}, origin[4] = {0, 0, 0, 0};
	int meshShape = np->registerConcaveMesh(&verts, &idx, one);
	int meshBody = pipe->registerPhysicsInstance(0.f, origin, orn, meshShape, 0, false);
	float half[4] = {0.4f, 0.4f, 0.4f, 1};
	int smallBox = np->registerConvexHullShape(cube, 3 * sizeof(float), 8, half);
	b3AlignedObjectArray<b3GpuChildShape> children;
	const float offs[3][3] = {{-0.4f, -0.4f, 0}, {0.4f, -0.4f, 0}, {-0.4f, 0.4f, 0}};
	for (int c = 0; c < 3; c++)
	{
		b3GpuChildShape ch;
		ch.m_childPosition = b3MakeVector3(offs[c][0], offs[c][1], offs[c][2]);
		ch.m_childOrientation = b3Quaternion(0, 0, 0, 1);
		ch.m_shapeIndex = smallBox;
		ch.m_shapeType = SHAPE_CONVEX_HULL;
		children.push_back(ch);
	}
	int compound = np->registerCompoundShape(&children);
	int sphere = np->registerSphereShape(0.5f);
	int shapes[2] = {smallBox, compound};  // (sphere x trimesh is not built: spheres would fall through)
	int n = 0;
	for (int i = 0; i < 6; i++)
		for (int k = 0; k < 6; k++)
			for (int j = 0; j < 2; j++)
			{
				float pos[4] = {-7.5f + 3.f * i, 2.5f + 1.6f * j, -7.5f + 3.f * k, 0};
				if (pipe->registerPhysicsInstance(1.f, pos, orn, shapes[(i + k + j) % 2], n, false) >= 0) n++;
			}
	pipe->writeAllInstancesToGpu();
	np->writeAllBodiesToGpu();
	bp->writeAabbsToGpu();
	// GpuTetraScene-style joints (examples/OpenCL/rigidbody/GpuConvexScene.cpp:411-583): tie two neighbours together,
	// and one deliberately weak joint that has to break
	float pivA[3] = {0.6f, 0, 0}, pivB[3] = {-0.6f, 0, 0};
	int uid0 = pipe->createPoint2PointConstraint(1, 2, pivA, pivB, 1e30f);
	float far[3] = {-8.f, 0, 0};
	int uid1 = pipe->createPoint2PointConstraint(3, 4, pivA, far, 0.001f);
	for (int s = 0; s < 240; s++) pipe->stepSimulation(1.f / 60.f);
	pipe->copyConstraintsToHost();
	pipe->removeConstraintByUid(uid1);
	bool jointsOk = uid0 == 0 && uid1 == 1 && pipe->getNumConstraints() == 1;
	np->readbackAllBodiesToCpu();
	const b3RigidBodyData* b = np->getBodiesCpu();
	int resting = 0;
	for (int i = 1; i < pipe->getNumBodies(); i++)
	{
		float ground = 0.8f * sinf(0.4f * b[i].m_pos.x) * cosf(0.4f * b[i].m_pos.z);
		if (fabsf(b[i].m_pos.x) < 12 && fabsf(b[i].m_pos.z) < 12 && b[i].m_pos.y > ground - 0.2f && b[i].m_pos.y < ground + 3.f) resting++;
	}
	// picking rays straight down on every box body (GpuRigidBodyDemo-style mouse pick, b3GpuRigidBodyPipeline::castRays): the box
	// under the ray, or one resting on top of it, must be reported; compounds and the mesh are skipped like in the reference
	b3AlignedObjectArray<b3RayInfo> rays;
	b3AlignedObjectArray<b3RayHit> hits;
	b3AlignedObjectArray<int> target;
	for (int i = 1; i < pipe->getNumBodies(); i++)
	{
		if (np->getCollidablesCpu()[b[i].m_collidableIdx].m_shapeType != SHAPE_CONVEX_HULL) continue;
		b3RayInfo r;
		r.m_from = b3MakeVector3(b[i].m_pos.x, 50.f, b[i].m_pos.z);
		r.m_to = b3MakeVector3(b[i].m_pos.x, -50.f, b[i].m_pos.z);
		b3RayHit h;
		h.m_hitFraction = 1.f;
		h.m_hitBody = -1;
		rays.push_back(r);
		hits.push_back(h);
		target.push_back(i);
	}
	pipe->castRays(rays, hits);
	int picked = 0;
	for (int i = 0; i < rays.size(); i++)
		if (hits[i].m_hitBody > 0 && hits[i].m_hitPoint.y >= b[target[i]].m_pos.y && hits[i].m_hitFraction < 0.5f) picked++;
	printf("mesh scene: %d picking rays, %d hit a box at or above their target\n", rays.size(), picked);
	jointsOk = jointsOk && rays.size() > 0 && picked == rays.size();
	printf("mesh scene: shapes mesh=%d box=%d compound=%d sphere=%d, bodies=%d, contacts=%d, resting on the mesh=%d\n", meshShape, smallBox, compound, sphere,
		   pipe->getNumBodies(), np->getNumContactsGpu(), resting);
	bool ok = meshShape >= 0 && meshBody == 0 && compound > smallBox && sphere > compound && pipe->getNumBodies() == n + 1 && np->getNumContactsGpu() >= n / 2 &&
			  resting >= (9 * n) / 10 && jointsOk;
	delete pipe;
	delete bp;
	delete np;
	return ok;
}

int main(int argc, char** argv)
{
	bool useUniformGrid = argc > 1 && argv[1][0] == 'g';
	cl_context ctx = 0;
	cl_device_id dev = 0;
	cl_command_queue q = 0;
	b3Config config;
	const int X = 10, Y = 10, Z = 10;
	config.m_maxConvexBodies = b3Max(config.m_maxConvexBodies, X * Y * Z + 10);
	config.m_maxConvexShapes = config.m_maxConvexBodies;
	config.m_maxBroadphasePairs = 16 * config.m_maxConvexBodies;
	config.m_maxContactCapacity = config.m_maxBroadphasePairs;

	b3GpuNarrowPhase* np = new b3GpuNarrowPhase(ctx, dev, q, config);
	b3GpuBroadphaseInterface* bp = useUniformGrid ? (b3GpuBroadphaseInterface*)new b3GpuGridBroadphase(ctx, dev, q) : (b3GpuBroadphaseInterface*)new b3GpuSapBroadphase(ctx, dev, q);
	b3GpuRigidBodyPipeline* pipe = new b3GpuRigidBodyPipeline(ctx, dev, q, np, bp, 0, config);

	// static environment (GpuConvexScene::createStaticEnvironment) + dynamic boxes (GpuBoxPlaneScene recipe)
	float scaling[4] = {400, 400, 400, 1};
	int groundShape = np->registerConvexHullShape(cube, 3 * sizeof(float), 8, scaling);
	float gpos[4] = {0, -400, 0, 0}, orn[4] = {0, 0, 0, 1};
	int ground = pipe->registerPhysicsInstance(0.f, gpos, orn, groundShape, 0, false);
	float one[4] = {1, 1, 1, 1};
	int boxShape = np->registerConvexHullShape(cube, 3 * sizeof(float), 8, one);
	int n = 0;
	for (int i = 0; i < X; i++)
		for (int j = 0; j < Y; j++)
			for (int k = 0; k < Z; k++)
			{
				float pos[4] = {((j + 1) & 1) + 2.2f * i, 1.f + 2.f * j, ((j + 1) & 1) + 2.2f * k, 0};
				if (pipe->registerPhysicsInstance(1.f, pos, orn, boxShape, n, false) >= 0) n++;
			}
	int bad = pipe->registerPhysicsInstance(1.f, gpos, orn, 12345, 0, false);  // invalid collidable -> -1
	pipe->writeAllInstancesToGpu();
	np->writeAllBodiesToGpu();
	bp->writeAabbsToGpu();

	for (int s = 0; s < 120; s++) pipe->stepSimulation(1.f / 60.f);

	np->readbackAllBodiesToCpu();
	const b3RigidBodyData* b = np->getBodiesCpu();
	float minY = 1e30f, maxSpeed = 0.f;
	for (int i = 1; i < pipe->getNumBodies(); i++)
	{
		minY = b3Min(minY, b[i].m_pos.y);
		maxSpeed = b3Max(maxSpeed, b[i].m_linVel.length());
	}
	printf("bodies=%d ground=%d bad=%d pairs=%d contacts=%d minY=%f maxSpeed=%f bodyBuffer=%p\n", pipe->getNumBodies(), ground, bad, bp->getNumOverlap(),
		   np->getNumContactsGpu(), minY, maxSpeed, pipe->getBodyBuffer());
	bool ok = pipe->getNumBodies() == n + 1 && ground == 0 && bad == -1 && bp->getNumOverlap() > n && np->getNumContactsGpu() > n / 2 && minY > 0.9f && minY < 1.1f &&
			  maxSpeed < 2.f && np->getStatic0Index() == 0;

	// PairBench-style stand-alone broadphase use
	b3GpuBroadphaseInterface* bp2 = b3GpuGridBroadphase::CreateFunc(ctx, dev, q);
	bp2->createProxy(b3MakeVector3(0, 0, 0), b3MakeVector3(1, 1, 1), 1024, 1, 1);
	bp2->createProxy(b3MakeVector3(0.5f, 0, 0), b3MakeVector3(1.5f, 1, 1), 1025, 1, 1);
	bp2->createProxy(b3MakeVector3(5, 0, 0), b3MakeVector3(6, 1, 1), 1026, 1, 1);
	bp2->createLargeProxy(b3MakeVector3(-100, -1, -100), b3MakeVector3(100, 0.1f, 100), 1027, 1, 1);
	bp2->writeAabbsToGpu();
	bp2->calculateOverlappingPairs(64);
	b3AlignedObjectArray<b3Int4> pairs;
	bp2->getOverlappingPairsGPU().copyToHost(pairs);
	printf("standalone pairs=%d\n", bp2->getNumOverlap());
	ok = ok && bp2->getNumOverlap() == 4 && pairs.size() == 4;
	delete bp2;

	// the stand-alone solver classes on the pipeline's own device buffers, the way b3GpuRigidBodyPipeline::stepSimulation
	// calls them (b3GpuRigidBodyPipeline.cpp:396-421): every box is pushed down at 1 m/s, the solver has to stop the pile
	for (int pass = 0; pass < 2; pass++)
	{
		float down[3] = {0, -1, 0}, zero[3] = {0, 0, 0};
		np->readbackAllBodiesToCpu();  // the CPU copy = the stepped state
		b3AlignedObjectArray<b3RigidBodyData> before;
		before.resize(pipe->getNumBodies());
		for (int i = 0; i < pipe->getNumBodies(); i++) before[i] = np->getBodiesCpu()[i];
		for (int i = 1; i < pipe->getNumBodies(); i++) np->setObjectVelocityCpu(down, zero, i);
		np->writeAllBodiesToGpu();
		{
			// the edit arrived on the device and nothing else moved (the pick flow of GpuRigidBodyDemo.cpp:457-463)
			np->readbackAllBodiesToCpu();
			const b3RigidBodyData* now = np->getBodiesCpu();
			bool arrived = true;
			for (int i = 1; i < pipe->getNumBodies(); i++)
				arrived = arrived && now[i].m_linVel.y == -1.f && now[i].m_pos.x == before[i].m_pos.x && now[i].m_pos.y == before[i].m_pos.y &&
						  now[i].m_pos.z == before[i].m_pos.z && now[i].m_quat.w == before[i].m_quat.w;
			printf("edit on CPU -> writeAllBodiesToGpu -> device: %s\n", arrived ? "arrived" : "LOST");
			ok = ok && arrived;
		}
		if (pass == 0)
		{
			b3GpuPgsContactSolver pgs(ctx, dev, q, config.m_maxBroadphasePairs);
			pgs.setNumIterations(20);
			pgs.solveContacts(np->getNumRigidBodies(), np->getBodiesGpu(), np->getBodyInertiasGpu(), np->getNumContactsGpu(), np->getContactsGpu(), config, np->getStatic0Index());
		}
		else
		{
			b3GpuJacobiContactSolver jac(ctx, dev, q, config.m_maxBroadphasePairs);
			jac.setNumIterations(40);
			jac.solveContacts(np->getNumRigidBodies(), np->getBodiesGpu(), np->getBodyInertiasGpu(), np->getNumContactsGpu(), np->getContactsGpu(), config, np->getStatic0Index());
		}
		np->readbackAllBodiesToCpu();
		const b3RigidBodyData* bb = np->getBodiesCpu();
		double sumVy = 0;
		bool finite = true;
		for (int i = 1; i < pipe->getNumBodies(); i++)
		{
			sumVy += bb[i].m_linVel.y;
			finite = finite && bb[i].m_linVel.y == bb[i].m_linVel.y;
		}
		const double meanVy = sumVy / (pipe->getNumBodies() - 1);
		printf("standalone %s solver: mean vertical velocity %.3f after the solve (was -1)\n", pass == 0 ? "PGS" : "Jacobi", meanVy);
		// the solver acted on the velocities that were written (-1): PGS stops the pile (and adds its penetration recovery), the
		// mass-splitting Jacobi converges more slowly through a 10-high pile
		ok = ok && finite && meanVy > (pass == 0 ? -0.5 : -0.9);
	}

	delete pipe;
	delete bp;
	delete np;
	ok = meshScene(useUniformGrid) && ok;
	printf(ok ? "DROPIN OK\n" : "DROPIN FAILED\n");
	return ok ? 0 : 1;
}
