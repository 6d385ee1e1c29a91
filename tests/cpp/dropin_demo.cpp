// dropin_demo.cpp -- caller code written against the reference's class API
// (examples/OpenCL/rigidbody/GpuRigidBodyDemo.cpp:110-154, 225 and examples/OpenCL/broadphase/PairBench.cpp:208-379),
// compiled unchanged against the B200 drop-in headers.  Prints "DROPIN OK" on success.
#include <stdio.h>
#include <math.h>
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuSapBroadphase.h"
#include "Bullet3OpenCL/BroadphaseCollision/b3GpuGridBroadphase.h"
#include "Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.h"
#include "Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.h"
#include "Bullet3Collision/NarrowPhaseCollision/b3Config.h"
#include "Bullet3Collision/NarrowPhaseCollision/shared/b3RigidBodyData.h"

static const float cube[8 * 3] = {-1, -1, -1, -1, -1, 1, -1, 1, -1, -1, 1, 1, 1, -1, -1, 1, -1, 1, 1, 1, -1, 1, 1, 1};

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

	delete pipe;
	delete bp;
	delete np;
	printf(ok ? "DROPIN OK\n" : "DROPIN FAILED\n");
	return ok ? 0 : 1;
}
