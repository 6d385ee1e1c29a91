// bt2mt_bench.cpp -- CPU baseline north_star names: Bullet 2's multithreaded world, assembled exactly like
// examples/MultiThreadedDemo/CommonRigidBodyMTBase.cpp:518-567 (btDiscreteDynamicsWorldMt + btDbvtBroadphase +
// btCollisionDispatcherMt(grain 40) + btConstraintSolverPoolMt + btSequentialImpulseConstraintSolverMt), built from the
// UNMODIFIED reference sources with -DBT_THREADSAFE=1 (oracle/Makefile target `bt2mt`), stepping a scene file that
// bench.py / tools dump from the same state the GPU world is in.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
//   bt2mt_bench <scene.bin> <warmup> <steps> <iterations> <threads (0 = all)>
// prints one JSON line {"ms_per_step":..,"bodies":..,"threads":..,"manifolds":..}
//
// scene file (little endian): int32 magic 0x62743273, numShapes, numBodies, numMeshVerts, numMeshIdx;
//   per shape: int32 kind (0 = convex points, 1 = compound of convex), int32 n;
//       kind 0: n x float3 points;   kind 1: n x {int32 childShape, float3 pos, float4 orn}
//   mesh: numMeshVerts x float3, numMeshIdx x int32   (static triangle mesh at the origin; may be empty)
//   per body: int32 shape (-1 = the mesh), float mass, float3 pos, float4 orn, float3 linVel, float3 angVel
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "btBulletDynamicsCommon.h"
#include "BulletCollision/CollisionDispatch/btCollisionDispatcherMt.h"
#include "BulletDynamics/Dynamics/btDiscreteDynamicsWorldMt.h"
#include "BulletDynamics/ConstraintSolver/btSequentialImpulseConstraintSolverMt.h"
#include "LinearMath/btThreads.h"

template <typename T>
static bool rd(FILE* f, T* dst, size_t n) { return fread(dst, sizeof(T), n, f) == n; }

int main(int argc, char** argv)
{
	if (argc < 6)
	{
		fprintf(stderr, "usage: %s scene.bin warmup steps iterations threads\n", argv[0]);
		return 2;
	}
	const int warmup = atoi(argv[2]), steps = atoi(argv[3]), iterations = atoi(argv[4]);
	int threads = atoi(argv[5]);
	FILE* f = fopen(argv[1], "rb");
	if (!f) return 3;
	int hdr[5];
	if (!rd(f, hdr, 5) || hdr[0] != 0x62743273) return 4;
	const int numShapes = hdr[1], numBodies = hdr[2], numMeshVerts = hdr[3], numMeshIdx = hdr[4];

	// ---- task scheduler + world (CommonRigidBodyMTBase.cpp:518-567)
	btITaskScheduler* sched = btCreateDefaultTaskScheduler();  // LinearMath/btThreads.h:158
	if (!sched) sched = btGetSequentialTaskScheduler();
	if (threads <= 0 || threads > sched->getMaxNumThreads()) threads = sched->getMaxNumThreads();
	sched->setNumThreads(threads);
	btSetTaskScheduler(sched);
	btDefaultCollisionConstructionInfo cci;
	cci.m_defaultMaxPersistentManifoldPoolSize = 80000;
	cci.m_defaultMaxCollisionAlgorithmPoolSize = 80000;
	btDefaultCollisionConfiguration* config = new btDefaultCollisionConfiguration(cci);
	btCollisionDispatcherMt* dispatcher = new btCollisionDispatcherMt(config, 40);
	btDbvtBroadphase* broadphase = new btDbvtBroadphase();
	btConstraintSolver* solvers[BT_MAX_THREAD_COUNT];
	for (int i = 0; i < BT_MAX_THREAD_COUNT; i++) solvers[i] = new btSequentialImpulseConstraintSolver();
	btConstraintSolverPoolMt* pool = new btConstraintSolverPoolMt(solvers, BT_MAX_THREAD_COUNT);
	btSequentialImpulseConstraintSolverMt* solverMt = new btSequentialImpulseConstraintSolverMt();
	btDiscreteDynamicsWorldMt* world = new btDiscreteDynamicsWorldMt(dispatcher, broadphase, pool, solverMt, config);
	world->setGravity(btVector3(0, -9.8f, 0));
	world->getSolverInfo().m_numIterations = iterations;

	// ---- shapes
	std::vector<btCollisionShape*> shapes(numShapes, nullptr);
	for (int s = 0; s < numShapes; s++)
	{
		int kn[2];
		if (!rd(f, kn, 2)) return 5;
		if (kn[0] == 0)
		{
			std::vector<float> p(3 * (size_t)kn[1]);
			if (!rd(f, p.data(), p.size())) return 5;
			btConvexHullShape* h = new btConvexHullShape(p.data(), kn[1], 12);
			h->setMargin(0.01f);
			shapes[s] = h;
		}
		else
		{
			btCompoundShape* c = new btCompoundShape();
			for (int k = 0; k < kn[1]; k++)
			{
				int child;
				float v[7];
				if (!rd(f, &child, 1) || !rd(f, v, 7)) return 5;
				btTransform t(btQuaternion(v[3], v[4], v[5], v[6]), btVector3(v[0], v[1], v[2]));
				c->addChildShape(t, shapes[child]);
			}
			shapes[s] = c;
		}
	}
	btTriangleIndexVertexArray* meshData = nullptr;
	btBvhTriangleMeshShape* mesh = nullptr;
	std::vector<float> mv(3 * (size_t)numMeshVerts);
	std::vector<int> mi((size_t)numMeshIdx);
	if (numMeshVerts && (!rd(f, mv.data(), mv.size()) || !rd(f, mi.data(), mi.size()))) return 6;
	if (numMeshVerts)
	{
		meshData = new btTriangleIndexVertexArray(numMeshIdx / 3, mi.data(), 12, numMeshVerts, mv.data(), 12);
		mesh = new btBvhTriangleMeshShape(meshData, true);
	}
	// ---- bodies
	int dynamic = 0;
	for (int b = 0; b < numBodies; b++)
	{
		int shape;
		float v[14];
		if (!rd(f, &shape, 1) || !rd(f, v, 14)) return 7;
		btCollisionShape* cs = shape < 0 ? (btCollisionShape*)mesh : shapes[shape];
		if (!cs) continue;
		const float mass = v[0];
		btVector3 inertia(0, 0, 0);
		if (mass != 0.f) cs->calculateLocalInertia(mass, inertia), dynamic++;
		btRigidBody::btRigidBodyConstructionInfo ci(mass, nullptr, cs, inertia);
		ci.m_startWorldTransform = btTransform(btQuaternion(v[4], v[5], v[6], v[7]), btVector3(v[1], v[2], v[3]));
		ci.m_friction = 0.7f;
		btRigidBody* body = new btRigidBody(ci);
		body->setLinearVelocity(btVector3(v[8], v[9], v[10]));
		body->setAngularVelocity(btVector3(v[11], v[12], v[13]));
		body->setActivationState(DISABLE_DEACTIVATION);  // the GPU pipeline has no sleeping: every body is stepped every step
		world->addRigidBody(body);
	}
	fclose(f);
	for (int i = 0; i < warmup; i++) world->stepSimulation(1.f / 60.f, 0, 1.f / 60.f);
	const auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < steps; i++) world->stepSimulation(1.f / 60.f, 0, 1.f / 60.f);
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	printf("{\"ms_per_step\": %.4f, \"bodies\": %d, \"dynamic\": %d, \"threads\": %d, \"steps\": %d, \"manifolds\": %d, \"iterations\": %d}\n",
		   sec / (steps > 0 ? steps : 1) * 1e3, numBodies, dynamic, threads, steps, dispatcher->getNumManifolds(), iterations);
	return 0;
}
