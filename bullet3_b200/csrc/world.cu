// world.cu -- world handle, registration, upload, the step and read-backs:
// the C ABI of include/b3b200.h.
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include "internal.h"
#include <algorithm>

namespace b3b200
{
static thread_local char g_lastError[512] = "";
std::atomic<long long> g_launchCount{0};
thread_local long long g_allocEpoch = 0;

void setLastError(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
	va_end(ap);
}

template <typename T, typename H>
static int uploadVec(DevBuf<T>& d, const std::vector<H>& h, size_t minCap, cudaStream_t s)
{
	static_assert(sizeof(T) == sizeof(H), "layout");
	size_t n = h.size();
	B3_TRY(d.reserve(std::max(std::max(n, minCap), (size_t)1)));
	if (n) B3_CUDA_CHECK(cudaMemcpyAsync(d.ptr, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, s));
	return 0;
}

int World::init(const b3b200_config* c, int dev, cudaStream_t st)
{
	cfg = *c;
	device = dev;
	B3_CUDA_CHECK(cudaSetDevice(device));
	if (st)
	{
		stream = st;
		ownStream = false;
	}
	else
	{
		B3_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		ownStream = true;
	}
	cudaDeviceProp prop;
	B3_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
	smCount = prop.multiProcessorCount;
	B3_TRY(bp.init(B3B200_BP_GRID, device, stream, cfg.maxConvexBodies, cfg.maxBroadphasePairs));
	B3_TRY(dCounters.reserve(CTR_COUNT));
	B3_CUDA_CHECK(cudaMemsetAsync(dCounters.ptr, 0, sizeof(unsigned int) * CTR_COUNT, stream));
	bp.ctr = dCounters.ptr;
	B3_TRY(dGridBarrier.reserve(4));
	B3_CUDA_CHECK(cudaMemsetAsync(dGridBarrier.ptr, 0, sizeof(unsigned int) * 4, stream));
	for (int i = 0; i < 8; i++) B3_CUDA_CHECK(cudaEventCreate(&ev[i]));
	for (int i = 0; i < 2; i++) B3_CUDA_CHECK(cudaEventCreate(&evSat[i]));
	for (int i = 0; i < 3; i++) B3_CUDA_CHECK(cudaStreamCreateWithFlags(&npStream[i], cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) B3_CUDA_CHECK(cudaEventCreateWithFlags(&evNpFork[i], cudaEventDisableTiming));
	for (int i = 0; i < 3; i++) B3_CUDA_CHECK(cudaEventCreateWithFlags(&evNpJoin[i], cudaEventDisableTiming));
	if (const char* e = getenv("B3B200_NP_OVERLAP")) npOverlap = atoi(e) != 0;
	if (const char* e = getenv("B3B200_GRAPHS")) useGraphs = atoi(e);
	return 0;
}

void World::dropStepGraphs()
{
	for (StepGraph& g : stepGraphs)
	{
		if (g.exec) cudaGraphExecDestroy(g.exec);
		g = StepGraph();
	}
}

void World::destroy()
{
	if (device < 0) return;  // host-only world: nothing on a device
	cudaSetDevice(device);
	slabDestroy(this);
	if (stream) cudaStreamSynchronize(stream);
	dropStepGraphs();
	if (pipe.h2d)
	{
		cudaStreamSynchronize(pipe.h2d);
		cudaStreamSynchronize(pipe.d2h);
		cudaStreamDestroy(pipe.h2d);
		cudaStreamDestroy(pipe.d2h);
		for (int i = 0; i < 2; i++)
		{
			cudaEventDestroy(pipe.inReady[i]);
			cudaEventDestroy(pipe.stepDone[i]);
			cudaEventDestroy(pipe.outDone[i]);
		}
		pipe.h2d = pipe.d2h = nullptr;
	}
	for (int i = 0; i < 8; i++)
		if (ev[i]) cudaEventDestroy(ev[i]);
	for (int i = 0; i < 2; i++)
		if (evSat[i]) cudaEventDestroy(evSat[i]);
	for (int i = 0; i < 3; i++)
	{
		if (npStream[i]) cudaStreamSynchronize(npStream[i]), cudaStreamDestroy(npStream[i]);
		if (evNpJoin[i]) cudaEventDestroy(evNpJoin[i]);
		npStream[i] = nullptr;
		evNpJoin[i] = nullptr;
	}
	for (int i = 0; i < 2; i++)
	{
		if (evNpFork[i]) cudaEventDestroy(evNpFork[i]);
		evNpFork[i] = nullptr;
	}
	cudaStream_t s = stream;
	bool own = ownStream;
	bp.stream = 0;  // shared with the world
	bp.ownStream = false;
	bp.destroy();
	if (own && s) cudaStreamDestroy(s);
	stream = 0;
}

// b3TransformAabb (src/Bullet3Geometry/b3AabbUtil.h:182-197), float, scalar path
void transformAabbHost(const float* lmn, const float* lmx, float margin, const float* pos, const float* orn, float* outMin, float* outMax)
{
	float4 half = mk4(0.5f * (lmx[0] - lmn[0]), 0.5f * (lmx[1] - lmn[1]), 0.5f * (lmx[2] - lmn[2]));
	half = mk4(half.x + margin, half.y + margin, half.z + margin);
	float4 lc = mk4(0.5f * (lmx[0] + lmn[0]), 0.5f * (lmx[1] + lmn[1]), 0.5f * (lmx[2] + lmn[2]));
	float4 q = mk4(orn[0], orn[1], orn[2], orn[3]);
	Mat3 m = matFromQuat(q);
	float4 a0 = mk4(fabsf(m.r0.x), fabsf(m.r0.y), fabsf(m.r0.z));
	float4 a1 = mk4(fabsf(m.r1.x), fabsf(m.r1.y), fabsf(m.r1.z));
	float4 a2 = mk4(fabsf(m.r2.x), fabsf(m.r2.y), fabsf(m.r2.z));
	float4 r = matMulVec(m, lc);
	float c[3] = {r.x + pos[0], r.y + pos[1], r.z + pos[2]};
	float e[3] = {dot3(half, a0), dot3(half, a1), dot3(half, a2)};
	for (int i = 0; i < 3; i++)
	{
		outMin[i] = c[i] - e[i];
		outMax[i] = c[i] + e[i];
	}
}

int allocateCollidable(World* w)
{
	// b3GpuNarrowPhase::allocateCollidable (b3GpuNarrowPhase.cpp:144-157)
	if ((int)w->collidables.size() >= w->cfg.maxConvexShapes)
	{
		setLastError("allocateCollidable out-of-range %d", w->cfg.maxConvexShapes);
		return -1;
	}
	b3b200_collidable c;
	memset(&c, 0, sizeof(c));
	w->collidables.push_back(c);
	b3b200_aabb a;
	memset(&a, 0, sizeof(a));
	w->localAabbs.push_back(a);
	return (int)w->collidables.size() - 1;
}

static int registerConvexInternal(World* w, const b3b200_float4* verts, int nV, const b3b200_face* faces, int nF, const int* idx, int nI,
								  const b3b200_float4* edges, int nE, const b3b200_convex_polyhedron* poly)
{
	// b3GpuNarrowPhase::registerConvexHullShapeInternal (b3GpuNarrowPhase.cpp:234-296)
	if ((int)w->vertices.size() + nV > w->cfg.maxConvexVertices || (int)w->indices.size() + nI > w->cfg.maxConvexIndices ||
		(int)w->uniqueEdges.size() + nE > w->cfg.maxConvexUniqueEdges || (int)w->convex.size() + 1 > w->cfg.maxConvexShapes)
	{
		setLastError("registerConvexHullShape: exceeding shape table capacity");
		return -1;
	}
	b3b200_convex_polyhedron c = *poly;
	c.numUniqueEdges = nE;
	c.uniqueEdgesOffset = (int)w->uniqueEdges.size();
	for (int i = 0; i < nE; i++) w->uniqueEdges.push_back(edges[i]);
	c.faceOffset = (int)w->faces.size();
	c.numFaces = nF;
	for (int i = 0; i < nF; i++)
	{
		b3b200_face f = faces[i];
		int off = (int)w->indices.size();
		if (f.indexOffset < 0 || f.numIndices < 0 || f.indexOffset + f.numIndices > nI)
		{
			setLastError("registerConvexHullShape: face %d indexes outside the index array", i);
			return -1;
		}
		for (int p = 0; p < f.numIndices; p++) w->indices.push_back(idx[f.indexOffset + p]);
		f.indexOffset = off;
		// pad1 = 1 when an earlier face of this hull has the bitwise same or exactly negated normal.
		// Such a face yields the identical SAT depth as the earlier one, so with the reference's strict
		// "d < dmin" it can never win: the narrowphase skips it (see narrowphase.cu).
		f.pad1 = 0;
		for (int j = 0; j < i; j++)
		{
			const b3b200_float4& a = faces[j].plane;
			const b3b200_float4& b = f.plane;
			bool same = a.x == b.x && a.y == b.y && a.z == b.z;
			bool opp = a.x == -b.x && a.y == -b.y && a.z == -b.z;
			if (same || opp)
			{
				f.pad1 = 1;
				break;
			}
		}
		w->faces.push_back(f);
	}
	{
		// radius = inscribed-sphere radius about localCenter (the reference leaves m_radius unset
		// unless TEST_INTERNAL_OBJECTS is defined, b3ConvexUtility.cpp:427-436).  Rounded down; used
		// only for the exact-safe lower bound that lets the SAT skip hopeless edge-edge axes.
		double rin = 1e300;
		for (int i = 0; i < nF; i++)
		{
			const b3b200_float4& pl = faces[i].plane;
			double nl = sqrt((double)pl.x * pl.x + (double)pl.y * pl.y + (double)pl.z * pl.z);
			if (nl < 1e-12) continue;
			double dist = -((double)pl.x * c.localCenter.x + (double)pl.y * c.localCenter.y + (double)pl.z * c.localCenter.z + (double)pl.w) / nl;
			if (dist < rin) rin = dist;
		}
		if (rin > 1e299 || rin < 0) rin = 0;
		c.radius = (float)(rin * (1.0 - 1e-5));
		// m_unused (b3ConvexPolyhedronData.h:33) carries the bits of the circumscribed radius about localCenter, rounded
		// up: the conservative bounding sphere the child-pair cull of compounds uses (narrowphase.cu)
		double rout = 0.0;
		for (int i = 0; i < nV; i++)
		{
			const double dx = (double)verts[i].x - c.localCenter.x, dy = (double)verts[i].y - c.localCenter.y, dz = (double)verts[i].z - c.localCenter.z;
			rout = std::max(rout, sqrt(dx * dx + dy * dy + dz * dz));
		}
		const float routF = (float)(rout * (1.0 + 1e-5)) + 1e-6f;
		memcpy(&c.unused, &routF, sizeof(float));
	}
	c.numVertices = nV;
	c.vertexOffset = (int)w->vertices.size();
	for (int i = 0; i < nV; i++) w->vertices.push_back(verts[i]);
	w->convex.push_back(c);
	return (int)w->convex.size() - 1;
}

static void initInertia(World* w, int bodyIndex, float mass, const float* aabbMin, const float* aabbMax)
{
	// b3GpuNarrowPhase::registerRigidBody (b3GpuNarrowPhase.cpp:859-903)
	b3b200_inertia& I = w->inertias[bodyIndex];
	memset(&I, 0, sizeof(I));
	if (mass == 0.f) return;
	float he[3] = {aabbMax[0] - aabbMin[0], aabbMax[1] - aabbMin[1], aabbMax[2] - aabbMin[2]};
	float lx = 2.f * he[0], ly = 2.f * he[1], lz = 2.f * he[2];
	float li[3] = {(mass / 12.0f) * (ly * ly + lz * lz), (mass / 12.0f) * (lx * lx + lz * lz), (mass / 12.0f) * (lx * lx + ly * ly)};
	float inv[3] = {1.f / li[0], 1.f / li[1], 1.f / li[2]};
	I.initInvInertia.row[0].x = inv[0];
	I.initInvInertia.row[1].y = inv[1];
	I.initInvInertia.row[2].z = inv[2];
	const b3b200_rigid_body& b = w->bodies[bodyIndex];
	Mat3 m = matFromQuat(mk4(b.quat.x, b.quat.y, b.quat.z, b.quat.w));
	// m.scaled(inv) * m.transpose()  (b3Matrix3x3.h operator*, tdotx/y/z)
	float M[3][3] = {{m.r0.x, m.r0.y, m.r0.z}, {m.r1.x, m.r1.y, m.r1.z}, {m.r2.x, m.r2.y, m.r2.z}};
	float A[3][3];
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) A[r][c] = M[r][c] * inv[c];
	float R[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) R[i][j] = M[j][0] * A[i][0] + M[j][1] * A[i][1] + M[j][2] * A[i][2];
	for (int i = 0; i < 3; i++)
	{
		I.invInertiaWorld.row[i].x = R[i][0];
		I.invInertiaWorld.row[i].y = R[i][1];
		I.invInertiaWorld.row[i].z = R[i][2];
		I.invInertiaWorld.row[i].w = 0.f;
	}
}

static int syncAoS(World* w)
{
	if (w->soaDirty) B3_TRY(launchUnpackSoA(w));
	return 0;
}

static int recordStage(World* w, int i)
{
	if (w->timing) B3_CUDA_CHECK(cudaEventRecord(w->ev[i], w->stream));
	return 0;
}

static int stepOnce(World* w, float dt)
{
	B3_TRY(recordStage(w, 0));
	if (!w->aabbsValid) B3_TRY(launchUpdateAabbs(w));
	B3_TRY(recordStage(w, 1));
	B3_TRY(w->bp.calculatePairs(w->cfg.maxBroadphasePairs));
	B3_TRY(recordStage(w, 2));
	B3_TRY(launchNarrowphase(w));
	B3_TRY(recordStage(w, 3));
	// joints first, then contacts (b3GpuRigidBodyPipeline.cpp:361-374 before :389-460)
	if (!w->joints.empty()) B3_TRY(launchSolveJoints(w));
	if (w->solverKind == B3B200_SOLVER_JACOBI)
	{
		B3_TRY(launchJacobi(w));
		B3_TRY(recordStage(w, 4));
	}
	else
	{
		B3_TRY(launchSolverSetup(w));
		B3_TRY(recordStage(w, 4));
		B3_TRY(launchSolverIterate(w));
	}
	B3_TRY(recordStage(w, 5));
	B3_TRY(launchIntegrate(w, dt, true));
	B3_TRY(recordStage(w, 6));
	return 0;
}

// One step through a captured CUDA graph (one cudaGraphLaunch instead of ~35 kernel launches + memsets: what a small world's
// step costs is launch latency).  The reference has nothing like it (7..15 clFinish per step).
static int stepGraphed(World* w, float dt)
{
	const bool partDue = !(w->partValid && w->partAge < w->partInterval && w->partBodies == w->numBodies);
	const bool eligible = w->useGraphs && !w->timing && w->joints.empty() && !w->npOverlap && w->solverKind == B3B200_SOLVER_PGS && w->dSolverProbe.ptr == nullptr;
	if (!eligible) return stepOnce(w, dt);
	World::StepGraph& g = w->stepGraphs[(w->aabbsValid ? 1 : 0) | (partDue ? 2 : 0)];
	if (g.state == 2 && g.dt != dt)
	{
		cudaGraphExecDestroy(g.exec);
		g = World::StepGraph();
		g.state = 1;
	}
	if (g.state == 0)
	{
		g.state = 1;
		return stepOnce(w, dt);
	}
	if (g.state == 1)
	{
		const long long epoch = g_allocEpoch, launches0 = g_launchCount;
		if (cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess)
		{
			// e.g. a caller-owned stream that is being captured by the caller already: step kernel by kernel from now on
			cudaGetLastError();
			w->useGraphs = 0;
			return stepOnce(w, dt);
		}
		const int rc = stepOnce(w, dt);  // also advances the host-side state exactly like an eager step
		cudaGraph_t graph = nullptr;
		const cudaError_t e = cudaStreamEndCapture(w->stream, &graph);
		if (rc < 0 || e != cudaSuccess || !graph)
		{
			if (graph) cudaGraphDestroy(graph);
			cudaGetLastError();
			w->useGraphs = 0;  // capture is not possible here (e.g. a caller-owned stream that is already capturing): stay eager
			if (rc < 0) return rc;
			setLastError("step graph capture failed: %s", cudaGetErrorString(e));
			return B3B200_ERR_CUDA;
		}
		cudaGraphExec_t exec = nullptr;
		const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
		cudaGraphDestroy(graph);
		if (ei != cudaSuccess)
		{
			cudaGetLastError();
			w->useGraphs = 0;
			setLastError("cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
			return B3B200_ERR_CUDA;
		}
		const long long launches = g_launchCount - launches0;
		B3_CUDA_CHECK(cudaGraphLaunch(exec, w->stream));
		if (g_allocEpoch != epoch)
		{
			// a buffer grew while capturing: this graph may hold a freed pointer after the next growth; use it once, then re-capture
			B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
			cudaGraphExecDestroy(exec);
			return 0;
		}
		g.exec = exec;
		g.launches = launches;
		g.dt = dt;
		g.state = 2;
		return 0;
	}
	B3_CUDA_CHECK(cudaGraphLaunch(g.exec, w->stream));
	g_launchCount += g.launches;
	// the host-side state an eager step leaves behind (solver.cu ensurePartition, launchSolverIterate, dynamics.cu launchIntegrate)
	if (partDue)
	{
		w->partValid = true;
		w->partAge = 1;
		w->partBodies = w->numBodies;
	}
	else
		w->partAge++;
	w->soaDirty = true;
	w->aabbsValid = true;
	return 0;
}

}  // namespace b3b200

using namespace b3b200;

// every API entry point that is not a plain step / copy drops the captured step graphs (it may change what a step launches)
#define W_CHECK_KEEP(w)                                                         \
	if (!(w)) return B3B200_ERR_INVALID;                                        \
	if ((w)->device < 0)                                                        \
	{                                                                           \
		setLastError("host-only world (device -1): no GPU work is possible");   \
		return B3B200_ERR_STATE;                                                \
	}                                                                           \
	B3_CUDA_CHECK(cudaSetDevice((w)->device))
#define W_CHECK(w)   \
	W_CHECK_KEEP(w); \
	(w)->dropStepGraphs()
#define W_UPLOADED_KEEP(w)                                     \
	W_CHECK_KEEP(w);                                           \
	if (!(w)->uploaded)                                        \
	{                                                          \
		setLastError("world not uploaded (call b3b200_upload)"); \
		return B3B200_ERR_STATE;                               \
	}
#define W_UPLOADED(w)   \
	W_UPLOADED_KEEP(w); \
	(w)->dropStepGraphs()

extern "C" const char* b3b200_last_error(void) { return g_lastError; }
extern "C" int b3b200_version(void) { return 100; }
extern "C" long long b3b200_launch_count(void) { return g_launchCount.load(); }

extern "C" int b3b200_config_default(b3b200_config* c)
{
	if (!c) return B3B200_ERR_INVALID;
	c->maxConvexBodies = 128 * 1024;
	c->maxVerticesPerFace = 64;
	c->maxFacesPerShape = 12;
	c->maxConvexVertices = 8192;
	c->maxConvexIndices = 81920;
	c->maxConvexUniqueEdges = 8192;
	c->maxCompoundChildShapes = 8192;
	c->maxTriConvexPairCapacity = 256 * 1024;
	c->maxConvexShapes = c->maxConvexBodies;
	c->maxBroadphasePairs = 16 * c->maxConvexBodies;
	c->maxContactCapacity = c->maxBroadphasePairs;
	c->compoundPairCapacity = 1024 * 1024;
	return 0;
}

extern "C" int b3b200_create(const b3b200_config* cfg, int device, void* stream, b3b200_world** out)
{
	if (!cfg || !out) return B3B200_ERR_INVALID;
	if (cfg->maxConvexBodies <= 0 || cfg->maxBroadphasePairs < 0 || cfg->maxContactCapacity < 0) return B3B200_ERR_INVALID;
	if (device == -1)
	{
		// host-only world: shape/body registration and table queries work, nothing else does
		b3b200_world* hw = new b3b200_world();
		hw->cfg = *cfg;
		hw->device = -1;
		hw->bp.maxProxies = cfg->maxConvexBodies;
		*out = hw;
		return 0;
	}
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || device < 0 || device >= count)
	{
		setLastError("no CUDA device %d (%s); this library has no CPU fallback", device, e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
		return B3B200_ERR_CUDA;
	}
	b3b200_world* w = new b3b200_world();
	int r = w->init(cfg, device, (cudaStream_t)stream);
	if (r < 0)
	{
		w->destroy();
		delete w;
		return r;
	}
	*out = w;
	return 0;
}

extern "C" int b3b200_destroy(b3b200_world* w)
{
	if (!w) return B3B200_ERR_INVALID;
	w->destroy();
	delete w;
	return 0;
}

extern "C" int b3b200_reset(b3b200_world* w)
{
	if (!w) return B3B200_ERR_INVALID;
	w->collidables.clear();
	w->localAabbs.clear();
	w->convex.clear();
	w->vertices.clear();
	w->uniqueEdges.clear();
	w->faces.clear();
	w->indices.clear();
	w->childShapes.clear();
	w->bvhInfos.clear();
	w->bvhNodes.clear();
	w->meshNodes.clear();
	w->meshTris.clear();
	w->meshInfos.clear();
	w->bvhSubtrees.clear();
	w->bodies.clear();
	w->inertias.clear();
	w->joints.clear();
	w->jointUid = 0;
	w->jointsDirty = false;
	w->jointBatchesDirty = false;
	w->haloIdsSet = false;
	w->numBodies = 0;
	w->static0Index = -1;
	w->uploaded = false;
	w->everUploaded = false;
	w->aabbsValid = false;
	w->partValid = false;
	w->solverMisc = nullptr;
	w->bp.reset();
	w->bp.worldOf = nullptr;
	w->bodyWorld.clear();
	w->currentWorld = 0;
	w->numWorlds = 1;
	w->worldDynBodies = 0;
	return 0;
}

extern "C" int b3b200_register_convex(b3b200_world* w, const b3b200_float4* vertices, int numVertices, const b3b200_face* faces, int numFaces,
									  const int* indices, int numIndices, const b3b200_float4* uniqueEdges, int numUniqueEdges,
									  const b3b200_convex_polyhedron* poly)
{
	if (!w || !vertices || !faces || !indices || !poly || numVertices <= 0 || numFaces <= 0 || numIndices <= 0 || numUniqueEdges < 0 ||
		(numUniqueEdges > 0 && !uniqueEdges))
	{
		setLastError("registerConvexHullShape: invalid argument");
		return -1;
	}
	int ci = allocateCollidable(w);
	if (ci < 0) return -1;
	b3b200_collidable& col = w->collidables[ci];
	col.shapeType = B3B200_SHAPE_CONVEX_HULL;
	col.shapeIndex = -1;
	b3b200_convex_polyhedron p = *poly;
	{
		// localCenter = vertex average (b3GpuNarrowPhase.cpp:332-337)
		float cx = 0.f, cy = 0.f, cz = 0.f;
		for (int i = 0; i < numVertices; i++)
		{
			cx += vertices[i].x;
			cy += vertices[i].y;
			cz += vertices[i].z;
		}
		float s = 1.f / numVertices;
		p.localCenter.x = cx * s;
		p.localCenter.y = cy * s;
		p.localCenter.z = cz * s;
		p.localCenter.w = 0.f;
	}
	int si = registerConvexInternal(w, vertices, numVertices, faces, numFaces, indices, numIndices, uniqueEdges, numUniqueEdges, &p);
	if (si < 0)
	{
		w->collidables.pop_back();
		w->localAabbs.pop_back();
		return -1;
	}
	w->collidables[ci].shapeIndex = si;
	// local AABB from the vertices (b3GpuNarrowPhase.cpp:343-365)
	b3b200_aabb& a = w->localAabbs[ci];
	float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
	for (int i = 0; i < numVertices; i++)
	{
		const float v[3] = {vertices[i].x, vertices[i].y, vertices[i].z};
		for (int k = 0; k < 3; k++)
		{
			if (v[k] < mn[k]) mn[k] = v[k];
			if (v[k] > mx[k]) mx[k] = v[k];
		}
	}
	for (int k = 0; k < 3; k++)
	{
		a.min[k] = mn[k];
		a.max[k] = mx[k];
	}
	a.minIndices[3] = 0;
	a.signedMaxIndices[3] = 0;
	return ci;
}

extern "C" int b3b200_register_convex_points(b3b200_world* w, const float* vertices, int strideInBytes, int numVertices, const float* scaling)
{
	if (!w || !vertices || numVertices <= 0 || strideInBytes < 12)
	{
		setLastError("registerConvexHullShape: invalid argument");
		return -1;
	}
	static const float one[3] = {1.f, 1.f, 1.f};
	if (!scaling) scaling = one;
	std::vector<b3b200_float4> pts(numVertices);
	const unsigned char* base = (const unsigned char*)vertices;
	for (int i = 0; i < numVertices; i++)
	{
		const float* v = (const float*)(base + (size_t)i * strideInBytes);
		pts[i].x = v[0] * scaling[0];
		pts[i].y = v[1] * scaling[1];
		pts[i].z = v[2] * scaling[2];
		pts[i].w = 0.f;
	}
	HullOut h;
	if (!buildConvexHull(pts, h))
	{
		setLastError("registerConvexHullShape: degenerate point set (need 4 non-coplanar points)");
		return -1;
	}
	return b3b200_register_convex(w, h.vertices.data(), (int)h.vertices.size(), h.faces.data(), (int)h.faces.size(), h.indices.data(),
								  (int)h.indices.size(), h.uniqueEdges.data(), (int)h.uniqueEdges.size(), &h.poly);
}

static int registerBodyCommon(b3b200_world* w, float mass, const float* position, const float* orientation, int collidableIndex,
							  const float* aabbMin, const float* aabbMax)
{
	if ((int)w->bodies.size() >= w->cfg.maxConvexBodies)
	{
		setLastError("registerRigidBody: exceeding the number of rigid bodies, %d > %d", (int)w->bodies.size(), w->cfg.maxConvexBodies);
		return -1;
	}
	// b3GpuNarrowPhase::registerRigidBody (b3GpuNarrowPhase.cpp:816-908)
	b3b200_rigid_body b;
	memset(&b, 0, sizeof(b));
	b.friction = 1.f;
	b.restitution = 0.f;
	b.pos.x = position[0];
	b.pos.y = position[1];
	b.pos.z = position[2];
	b.quat.x = orientation[0];
	b.quat.y = orientation[1];
	b.quat.z = orientation[2];
	b.quat.w = orientation[3];
	b.collidableIdx = collidableIndex;
	b.invMass = mass ? 1.f / mass : 0.f;
	int bodyIndex = (int)w->bodies.size();
	w->bodies.push_back(b);
	w->inertias.push_back(b3b200_inertia());
	if (mass == 0.f && bodyIndex == 0) w->static0Index = 0;
	initInertia(w, bodyIndex, mass, aabbMin, aabbMax);
	// createProxy / createLargeProxy (b3GpuRigidBodyPipeline.cpp:648-657)
	int r = w->bp.createProxy(aabbMin, aabbMax, bodyIndex, mass == 0.f);
	if (r < 0)
	{
		w->bodies.pop_back();
		w->inertias.pop_back();
		return -1;
	}
	w->uploaded = false;
	w->bodyWorld.push_back(w->currentWorld);
	if (w->currentWorld + 1 > w->numWorlds) w->numWorlds = w->currentWorld + 1;
	return bodyIndex;
}

extern "C" int b3b200_set_current_world(b3b200_world* w, int worldIndex)
{
	if (!w || worldIndex < 0 || worldIndex >= (1 << 20)) return B3B200_ERR_INVALID;
	w->currentWorld = worldIndex;
	return 0;
}
extern "C" int b3b200_num_worlds(b3b200_world* w) { return w ? w->numWorlds : B3B200_ERR_INVALID; }
extern "C" int b3b200_get_body_worlds(b3b200_world* w, int* dst, int n)
{
	if (!w || !dst || n < 0 || n > (int)w->bodyWorld.size()) return B3B200_ERR_INVALID;
	for (int i = 0; i < n; i++) dst[i] = w->bodyWorld[i];
	return 0;
}

extern "C" int b3b200_register_instance(b3b200_world* w, float mass, const float* position, const float* orientation, int collidableIndex, int userIndex)
{
	(void)userIndex;
	if (!w || !position || !orientation) return -1;
	if (collidableIndex < 0 || collidableIndex >= (int)w->collidables.size())
	{
		setLastError("registerPhysicsInstance using invalid collidableIndex");
		return -1;
	}
	const b3b200_aabb& la = w->localAabbs[collidableIndex];
	float aabbMin[3], aabbMax[3];
	transformAabbHost(la.min, la.max, 0.01f, position, orientation, aabbMin, aabbMax);
	return registerBodyCommon(w, mass, position, orientation, collidableIndex, aabbMin, aabbMax);
}

extern "C" int b3b200_register_body(b3b200_world* w, int collidableIndex, float mass, const float* position, const float* orientation,
									const float* aabbMin3, const float* aabbMax3)
{
	if (!w || !position || !orientation || !aabbMin3 || !aabbMax3) return -1;
	if (collidableIndex < 0 || collidableIndex >= (int)w->collidables.size())
	{
		setLastError("registerRigidBody using invalid collidableIndex");
		return -1;
	}
	return registerBodyCommon(w, mass, position, orientation, collidableIndex, aabbMin3, aabbMax3);
}

extern "C" int b3b200_register_instances(b3b200_world* w, int n, const float* masses, const float* positions4, const float* orientations4,
										 const int* collidableIndices)
{
	if (!w || n < 0 || (n > 0 && (!masses || !positions4 || !orientations4 || !collidableIndices))) return -1;
	int first = (int)w->bodies.size();
	w->bodies.reserve(w->bodies.size() + n);
	w->inertias.reserve(w->inertias.size() + n);
	for (int i = 0; i < n; i++)
		if (b3b200_register_instance(w, masses[i], positions4 + 4 * i, orientations4 + 4 * i, collidableIndices[i], 0) < 0) return -1;
	return first;
}

extern "C" int b3b200_upload(b3b200_world* w)
{
	W_CHECK(w);
	cudaStream_t s = w->stream;
	B3_CUDA_CHECK(cudaStreamSynchronize(s));
	if (w->everUploaded && w->numBodies > 0 && w->dBodiesAoS.ptr)
	{
		// the bodies that are already on the device have been stepped / written / solved there: their state comes back into the
		// host mirror before the tables are re-sent, so that an upload after registering one more body does not rewind them
		if (w->soaDirty) B3_TRY(launchUnpackSoA(w));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->bodies.data(), w->dBodiesAoS.ptr, sizeof(b3b200_rigid_body) * (size_t)w->numBodies, cudaMemcpyDeviceToHost, s));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->inertias.data(), w->dInertias.ptr, sizeof(b3b200_inertia) * (size_t)w->numBodies, cudaMemcpyDeviceToHost, s));  // (checkpoint / halo paths write them on the device)
		B3_CUDA_CHECK(cudaStreamSynchronize(s));
	}
	w->hostBodiesStale = false;
	w->numBodies = (int)w->bodies.size();
	const size_t nb = std::max(w->numBodies, 1);
	B3_TRY(uploadVec(w->dCollidables, w->collidables, 0, s));
	B3_TRY(uploadVec(w->dLocalAabbs, w->localAabbs, 0, s));
	B3_TRY(uploadVec(w->dConvex, w->convex, 0, s));
	B3_TRY(uploadVec(w->dVertices, w->vertices, 0, s));
	B3_TRY(uploadVec(w->dUniqueEdges, w->uniqueEdges, 0, s));
	B3_TRY(uploadVec(w->dFaces, w->faces, 0, s));
	B3_TRY(uploadVec(w->dIndices, w->indices, 0, s));
	B3_TRY(uploadVec(w->dChildShapes, w->childShapes, 0, s));
	{
		// conservative bounding sphere per child hull, in the compound's frame: centre = childPos + R(childOrn) * hull centre,
		// radius = the hull's circumscribed radius (convex entry's unused word), padded; w < 0 marks a child that is not a hull
		std::vector<float4> spheres(w->childShapes.size());
		for (size_t k = 0; k < w->childShapes.size(); k++)
		{
			const b3b200_child_shape& ch = w->childShapes[k];
			float4 sp = make_float4(0.f, 0.f, 0.f, -1.f);
			const int ci = ch.shapeIndex;
			if (ci >= 0 && ci < (int)w->collidables.size() && w->collidables[ci].shapeType == B3B200_SHAPE_CONVEX_HULL)
			{
				const b3b200_convex_polyhedron& cv = w->convex[w->collidables[ci].shapeIndex];
				float r;
				memcpy(&r, &cv.unused, sizeof(float));
				const float4 lc = quatRotate(mk4(ch.childOrientation.x, ch.childOrientation.y, ch.childOrientation.z, ch.childOrientation.w),
											 mk4(cv.localCenter.x, cv.localCenter.y, cv.localCenter.z));
				sp = make_float4(lc.x + ch.childPosition.x, lc.y + ch.childPosition.y, lc.z + ch.childPosition.z, r * 1.0001f + 1e-5f);
			}
			spheres[k] = sp;
		}
		B3_TRY(uploadVec(w->dChildSpheres, spheres, 0, s));
	}
	B3_TRY(uploadVec(w->dBvhInfos, w->bvhInfos, 0, s));
	B3_TRY(uploadVec(w->dBvhNodes, w->bvhNodes, 0, s));
	B3_TRY(uploadVec(w->dBvhSubtrees, w->bvhSubtrees, 0, s));
	B3_TRY(uploadVec(w->dMeshNodes, w->meshNodes, 0, s));
	B3_TRY(uploadVec(w->dMeshTris, w->meshTris, 0, s));
	B3_TRY(uploadVec(w->dMeshInfos, w->meshInfos, 0, s));
	B3_TRY(uploadVec(w->dBodiesAoS, w->bodies, 0, s));
	B3_TRY(uploadVec(w->dInertias, w->inertias, 0, s));
	B3_TRY(w->dPose.reserve(2 * nb));
	B3_TRY(w->dVel.reserve(2 * nb));
	B3_TRY(w->dCollidableIdx.reserve(nb));
	const size_t nc = std::max(w->cfg.maxContactCapacity, 1);
	B3_TRY(w->dContacts.reserve(nc));
	// work items: one per convex pair + child pairs of compounds (b3Config::m_compoundPairCapacity)
	const size_t nItems = (size_t)std::max(w->cfg.maxBroadphasePairs, 1) + (w->childShapes.empty() ? 0 : (size_t)std::max(w->cfg.compoundPairCapacity, 0));
	B3_TRY(w->dSurvivors.reserve(nItems));
	B3_TRY(w->dSmallItems.reserve(nItems));
	B3_TRY(w->dOverlapPairs.reserve(nItems));
	B3_TRY(w->dOverlapSep.reserve(nItems));
	w->hasPlanes = false;
	w->hasConcave = false;
	for (size_t i = 0; i < w->collidables.size(); i++)
	{
		if (w->collidables[i].shapeType == B3B200_SHAPE_PLANE || w->collidables[i].shapeType == B3B200_SHAPE_SPHERE) w->hasPlanes = true;  // primitives kernel
		if (w->collidables[i].shapeType == B3B200_SHAPE_CONCAVE_TRIMESH) w->hasConcave = true;
	}
	if (w->hasConcave)
	{
		B3_TRY(w->dConcavePairs.reserve((size_t)std::max(w->cfg.maxTriConvexPairCapacity, 1)));
		B3_TRY(w->dConcaveSurvivors.reserve((size_t)std::max(w->cfg.maxTriConvexPairCapacity, 1)));
	}
	B3_TRY(w->dBodyCount.reserve(std::max(nb, (size_t)1024)));
	w->bp.worldOf = nullptr;
	w->worldDynBodies = 0;
	if (w->numWorlds > 1)
	{
		// batched independent worlds: world id per body (= per proxy: one proxy per body, in creation order), the large proxies
		// grouped by world, and -- when every world has the same number of dynamic bodies -- that number (solver.cu blockSizeFor)
		B3_TRY(uploadVec(w->dBodyWorld, w->bodyWorld, 0, s));
		w->bp.worldOf = w->dBodyWorld.ptr;
		std::stable_sort(w->bp.largeIdx.begin(), w->bp.largeIdx.end(), [&](int a, int b) { return w->bodyWorld[a] < w->bodyWorld[b]; });
		std::vector<int> start((size_t)w->numWorlds + 1, 0), dyn((size_t)w->numWorlds, 0);
		for (int idx : w->bp.largeIdx) start[(size_t)w->bodyWorld[idx] + 1]++;
		for (int k = 0; k < w->numWorlds; k++) start[k + 1] += start[k];
		B3_TRY(uploadVec(w->bp.largeStart, start, 0, s));
		for (int i = 0; i < w->numBodies; i++)
			if (w->bodies[i].invMass != 0.f) dyn[w->bodyWorld[i]]++;
		bool uniform = true;
		for (int k = 1; k < w->numWorlds; k++) uniform = uniform && dyn[k] == dyn[0];
		if (uniform && dyn[0] > 0) w->worldDynBodies = dyn[0];
	}
	B3_TRY(w->bp.writeAabbs());
	B3_TRY(launchPackSoA(w));
	B3_CUDA_CHECK(cudaStreamSynchronize(s));
	w->uploaded = true;
	w->everUploaded = true;
	w->aabbsValid = false;
	w->soaDirty = false;
	return 0;
}

extern "C" int b3b200_set_gravity(b3b200_world* w, const float* g)
{
	if (!w || !g) return B3B200_ERR_INVALID;
	w->dropStepGraphs();  // a captured step holds the old value
	w->gravity[0] = g[0];
	w->gravity[1] = g[1];
	w->gravity[2] = g[2];
	return 0;
}
extern "C" int b3b200_set_solver(b3b200_world* w, int kind, int iterations)
{
	if (!w || iterations < 0 || (kind != B3B200_SOLVER_PGS && kind != B3B200_SOLVER_JACOBI)) return B3B200_ERR_INVALID;
	w->dropStepGraphs();  // a captured step holds the old value
	w->solverKind = kind;
	w->solverIterations = iterations;
	return 0;
}
extern "C" int b3b200_set_broadphase(b3b200_world* w, int kind)
{
	if (!w || (kind != B3B200_BP_SAP && kind != B3B200_BP_GRID)) return B3B200_ERR_INVALID;
	w->dropStepGraphs();  // a captured step holds the old value
	w->bp.kind = kind;
	return 0;
}
extern "C" int b3b200_set_step_graphs(b3b200_world* w, int on)
{
	if (!w) return B3B200_ERR_INVALID;
	w->dropStepGraphs();
	w->useGraphs = on ? 1 : 0;
	return 0;
}
extern "C" int b3b200_set_colouring(b3b200_world* w, int mode)
{
	if (!w || mode < 0 || mode > 1) return B3B200_ERR_INVALID;
	w->dropStepGraphs();  // a captured step holds the old value
	w->solverColouring = mode;
	return 0;
}
extern "C" int b3b200_set_contact_clip(b3b200_world* w, float minDist, float maxDist)
{
	if (!w) return B3B200_ERR_INVALID;
	w->dropStepGraphs();  // a captured step holds the old value
	w->clipMinDist = minDist;
	w->clipMaxDist = maxDist;
	return 0;
}
extern "C" int b3b200_set_angular_damping(b3b200_world* w, float d)
{
	if (!w) return B3B200_ERR_INVALID;
	w->dropStepGraphs();  // a captured step holds the old value
	w->angularDamping = d;
	return 0;
}

extern "C" int b3b200_write_bodies(b3b200_world* w, const b3b200_rigid_body* src, int n)
{
	W_UPLOADED_KEEP(w);
	if (!src || n != w->numBodies) return B3B200_ERR_INVALID;
	// the host mirror (get_table) is refreshed lazily: copying 80 B x N on the host here cost more than the transfer
	w->hostBodiesStale = true;
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dBodiesAoS.ptr, src, sizeof(b3b200_rigid_body) * n, cudaMemcpyHostToDevice, w->stream));
	B3_TRY(launchPackSoA(w));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	w->aabbsValid = false;
	return 0;
}

extern "C" int b3b200_readback_bodies(b3b200_world* w, b3b200_rigid_body* dst, int n)
{
	W_UPLOADED_KEEP(w);
	if (!dst || n < 0 || n > w->numBodies) return B3B200_ERR_INVALID;
	B3_TRY(syncAoS(w));
	if (n) B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->dBodiesAoS.ptr, sizeof(b3b200_rigid_body) * n, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}
extern "C" int b3b200_write_body(b3b200_world* w, int bodyIndex, const b3b200_rigid_body* src)
{
	W_UPLOADED_KEEP(w);
	if (!src || bodyIndex < 0 || bodyIndex >= w->numBodies) return B3B200_ERR_INVALID;
	B3_TRY(syncAoS(w));
	B3_CUDA_CHECK(cudaMemcpyAsync(&w->dBodiesAoS.ptr[bodyIndex], src, sizeof(b3b200_rigid_body), cudaMemcpyHostToDevice, w->stream));
	B3_TRY(launchPackSoA(w));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	w->hostBodiesStale = true;
	w->aabbsValid = false;
	return 0;
}
extern "C" int b3b200_read_body(b3b200_world* w, int bodyIndex, b3b200_rigid_body* dst)
{
	W_UPLOADED_KEEP(w);
	if (!dst || bodyIndex < 0 || bodyIndex >= w->numBodies) return B3B200_ERR_INVALID;
	B3_TRY(syncAoS(w));
	B3_CUDA_CHECK(cudaMemcpyAsync(dst, &w->dBodiesAoS.ptr[bodyIndex], sizeof(b3b200_rigid_body), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}
extern "C" int b3b200_readback_inertias(b3b200_world* w, b3b200_inertia* dst, int n)
{
	W_UPLOADED(w);
	if (!dst || n < 0 || n > w->numBodies) return B3B200_ERR_INVALID;
	if (n) B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->dInertias.ptr, sizeof(b3b200_inertia) * n, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}
extern "C" int b3b200_num_bodies(b3b200_world* w) { return w ? (int)w->bodies.size() : B3B200_ERR_INVALID; }

extern "C" int b3b200_step(b3b200_world* w, float dt)
{
	W_UPLOADED_KEEP(w);
	B3_TRY(stepGraphed(w, dt));
	if (w->timing)
	{
		B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
		for (int i = 0; i < 6; i++) cudaEventElapsedTime(&w->stageMs[i], w->ev[i], w->ev[i + 1]);
		cudaEventElapsedTime(&w->stageMs[6], w->ev[0], w->ev[6]);
		if (cudaEventElapsedTime(&w->stageMs[7], w->evSat[0], w->evSat[1]) != cudaSuccess)
		{
			w->stageMs[7] = 0.f;  // no SAT launch was recorded in this step
			cudaGetLastError();
		}
	}
	return 0;
}
extern "C" int b3b200_step_n(b3b200_world* w, float dt, int n)
{
	W_UPLOADED_KEEP(w);
	for (int i = 0; i < n; i++) B3_TRY(stepGraphed(w, dt));
	return 0;
}
// ---------------------------------------------------------------- pipelined stepping from / to host memory
// The reference's host loop is writeAllBodiesToGpu -> stepSimulation -> readbackAllBodiesToCpu, each blocking
// (b3GpuRigidBodyPipeline.cpp:221-463, b3GpuNarrowPhase.cpp:1020-1040).  Here the three overlap ACROSS calls: call c uploads into
// staging slot c & 1 on an upload stream while call c - 1 still computes, and its result leaves on a download stream while
// call c + 1 computes.  Nothing blocks the host; b3b200_step_host_wait drains the pipe.
static int ensureHostPipe(World* w)
{
	World::HostPipe& p = w->pipe;
	if (!p.h2d)
	{
		B3_CUDA_CHECK(cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking));
		B3_CUDA_CHECK(cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking));
		for (int i = 0; i < 2; i++)
		{
			B3_CUDA_CHECK(cudaEventCreateWithFlags(&p.inReady[i], cudaEventDisableTiming));
			B3_CUDA_CHECK(cudaEventCreateWithFlags(&p.stepDone[i], cudaEventDisableTiming));
			B3_CUDA_CHECK(cudaEventCreateWithFlags(&p.outDone[i], cudaEventDisableTiming));
		}
	}
	for (int i = 0; i < 2; i++) B3_TRY(p.stage[i].reserve((size_t)std::max(w->numBodies, 1)));
	return 0;
}

extern "C" int b3b200_step_host_async(b3b200_world* w, float dt, const b3b200_rigid_body* hostIn, b3b200_rigid_body* hostOut, int n)
{
	W_UPLOADED_KEEP(w);
	if (n != w->numBodies || (!hostIn && !hostOut)) return B3B200_ERR_INVALID;
	B3_TRY(ensureHostPipe(w));
	World::HostPipe& p = w->pipe;
	const int slot = (int)(p.calls++ & 1);
	const size_t bytes = sizeof(b3b200_rigid_body) * (size_t)n;
	if (hostIn)
	{
		// the slot is free once the step that used it two calls ago has consumed it and its result has left
		if (p.stepRecorded[slot]) B3_CUDA_CHECK(cudaStreamWaitEvent(p.h2d, p.stepDone[slot], 0));
		if (p.outPending[slot]) B3_CUDA_CHECK(cudaStreamWaitEvent(p.h2d, p.outDone[slot], 0));
		B3_CUDA_CHECK(cudaMemcpyAsync(p.stage[slot].ptr, hostIn, bytes, cudaMemcpyHostToDevice, p.h2d));
		B3_CUDA_CHECK(cudaEventRecord(p.inReady[slot], p.h2d));
		B3_CUDA_CHECK(cudaStreamWaitEvent(w->stream, p.inReady[slot], 0));
		B3_TRY(launchPackSoAFrom(w, p.stage[slot].ptr));
		w->hostBodiesStale = true;
	}
	B3_TRY(stepGraphed(w, dt));
	if (hostOut)
	{
		if (!hostIn)
		{
			// chained stepping: the records' other fields come from the world's own array
			if (p.outPending[slot]) B3_CUDA_CHECK(cudaStreamWaitEvent(w->stream, p.outDone[slot], 0));
			B3_TRY(syncAoS(w));
			B3_CUDA_CHECK(cudaMemcpyAsync(p.stage[slot].ptr, w->dBodiesAoS.ptr, bytes, cudaMemcpyDeviceToDevice, w->stream));
		}
		else
			B3_TRY(launchUnpackSoATo(w, p.stage[slot].ptr));
	}
	B3_CUDA_CHECK(cudaEventRecord(p.stepDone[slot], w->stream));
	p.stepRecorded[slot] = true;
	p.outPending[slot] = false;
	if (hostOut)
	{
		B3_CUDA_CHECK(cudaStreamWaitEvent(p.d2h, p.stepDone[slot], 0));
		B3_CUDA_CHECK(cudaMemcpyAsync(hostOut, p.stage[slot].ptr, bytes, cudaMemcpyDeviceToHost, p.d2h));
		B3_CUDA_CHECK(cudaEventRecord(p.outDone[slot], p.d2h));
		p.outPending[slot] = true;
	}
	return 0;
}
extern "C" int b3b200_step_host_wait(b3b200_world* w)
{
	W_CHECK_KEEP(w);
	if (w->pipe.d2h) B3_CUDA_CHECK(cudaStreamSynchronize(w->pipe.d2h));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}

extern "C" int b3b200_slab_step_n(b3b200_world* w, float dt, int n)
{
	W_UPLOADED(w);
	if (!w->slab.active)
	{
		setLastError("slab_step without slab_init");
		return B3B200_ERR_STATE;
	}
	for (int i = 0; i < n; i++)
	{
		B3_TRY(stepOnce(w, dt));
		B3_TRY(slabExchange(w));
	}
	return 0;
}
extern "C" int b3b200_slab_step(b3b200_world* w, float dt) { return b3b200_slab_step_n(w, dt, 1); }
extern "C" int b3b200_synchronize(b3b200_world* w)
{
	W_CHECK_KEEP(w);
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	// the steps themselves never look at their counters (no host round trip); this is where a capacity overrun becomes visible
	// without asking for the counters: the call still succeeds (the buffers were clamped, like the reference's b3Error + clamp),
	// b3b200_last_error() says which capacity was hit
	if (w->dCounters.ptr)
	{
		unsigned int ovf = 0;
		if (cudaMemcpy(&ovf, &w->dCounters.ptr[CTR_OVERFLOW], sizeof(ovf), cudaMemcpyDeviceToHost) == cudaSuccess && ovf)
			setLastError("capacity overrun in the last steps (flags 0x%x:%s%s%s%s%s%s): results were clamped", ovf, (ovf & OVF_PAIRS) ? " pairs" : "",
						 (ovf & OVF_CONTACTS) ? " contacts" : "", (ovf & OVF_BATCHES) ? " batches" : "", (ovf & OVF_COMPOUND) ? " work-items" : "",
						 (ovf & OVF_CONCAVE) ? " triangle-pairs" : "", (ovf & OVF_HALO) ? " halo" : "");
	}
	return 0;
}

extern "C" int b3b200_update_aabbs(b3b200_world* w)
{
	W_UPLOADED(w);
	return launchUpdateAabbs(w);
}
extern "C" int b3b200_find_pairs(b3b200_world* w)
{
	W_UPLOADED(w);
	return w->bp.calculatePairs(w->cfg.maxBroadphasePairs);
}
extern "C" int b3b200_compute_contacts(b3b200_world* w)
{
	W_UPLOADED(w);
	return launchNarrowphase(w);
}
extern "C" int b3b200_solver_setup(b3b200_world* w)
{
	W_UPLOADED(w);
	return launchSolverSetup(w);
}
extern "C" int b3b200_solver_iterate(b3b200_world* w)
{
	W_UPLOADED(w);
	return launchSolverIterate(w);
}
extern "C" int b3b200_solve_contacts(b3b200_world* w)
{
	W_UPLOADED(w);
	if (w->solverKind == B3B200_SOLVER_JACOBI) return launchJacobi(w);
	B3_TRY(launchSolverSetup(w));
	return launchSolverIterate(w);
}
extern "C" int b3b200_integrate(b3b200_world* w, float dt)
{
	W_UPLOADED(w);
	return launchIntegrate(w, dt, false);
}

static int readCounters(World* w, unsigned int* c)
{
	B3_CUDA_CHECK(cudaMemcpyAsync(c, w->dCounters.ptr, sizeof(unsigned int) * CTR_COUNT, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}

extern "C" int b3b200_get_aabbs(b3b200_world* w, b3b200_aabb* dst, int n)
{
	W_UPLOADED(w);
	if (!dst || n < 0 || n > w->numBodies) return B3B200_ERR_INVALID;
	if (n) B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->bp.aabbs.ptr, sizeof(b3b200_aabb) * n, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}
extern "C" int b3b200_get_pairs(b3b200_world* w, b3b200_int4* dst, int capacity, int* numPairs)
{
	W_UPLOADED(w);
	if (!numPairs || capacity < 0) return B3B200_ERR_INVALID;
	unsigned int c[CTR_COUNT];
	B3_TRY(readCounters(w, c));
	*numPairs = (int)c[CTR_PAIRS];
	int m = std::min((int)c[CTR_PAIRS], capacity);
	if (m > 0 && dst)
	{
		B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->bp.pairs.ptr, sizeof(b3b200_int4) * m, cudaMemcpyDeviceToHost, w->stream));
		B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	}
	return 0;
}
extern "C" int b3b200_get_contacts(b3b200_world* w, b3b200_contact4* dst, int capacity, int* numContacts)
{
	W_UPLOADED(w);
	if (!numContacts || capacity < 0) return B3B200_ERR_INVALID;
	unsigned int c[CTR_COUNT];
	B3_TRY(readCounters(w, c));
	*numContacts = (int)c[CTR_CONTACTS];
	int m = std::min((int)c[CTR_CONTACTS], capacity);
	if (m > 0 && dst)
	{
		B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->dContacts.ptr, sizeof(b3b200_contact4) * m, cudaMemcpyDeviceToHost, w->stream));
		B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	}
	return 0;
}
// b3GpuPgsContactSolver::solveContacts / b3GpuJacobiContactSolver::solveContacts (b3GpuPgsContactSolver.cpp:568-1103,
// b3GpuJacobiContactSolver.cpp:699-869) on CALLER-OWNED buffers with the reference AoS layouts.  `w` is only the scratch
// context (solver kind, iteration count, work buffers): the first numBodies body slots, the inertias and the contact
// buffer are overwritten, the solve runs, and the bodies (velocities updated) are copied back.  The pointers may be
// device or host pointers (cudaMemcpyDefault).
extern "C" int b3b200_solve_contacts_device(b3b200_world* w, int numBodies, void* bodies, const void* inertias, int numContacts, const void* contacts,
											int static0Index)
{
	W_UPLOADED(w);
	if (numBodies < 0 || numBodies > w->numBodies || numContacts < 0 || numContacts > w->cfg.maxContactCapacity) return B3B200_ERR_INVALID;
	if ((numBodies > 0 && (!bodies || !inertias)) || (numContacts > 0 && !contacts)) return B3B200_ERR_INVALID;
	if (numBodies == 0) return 0;
	cudaStream_t s = w->stream;
	B3_TRY(syncAoS(w));
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dBodiesAoS.ptr, bodies, sizeof(b3b200_rigid_body) * (size_t)numBodies, cudaMemcpyDefault, s));
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dInertias.ptr, inertias, sizeof(b3b200_inertia) * (size_t)numBodies, cudaMemcpyDefault, s));
	if (numContacts) B3_CUDA_CHECK(cudaMemcpyAsync(w->dContacts.ptr, contacts, sizeof(b3b200_contact4) * (size_t)numContacts, cudaMemcpyDefault, s));
	const unsigned int n = (unsigned int)numContacts;
	B3_CUDA_CHECK(cudaMemcpyAsync(&w->dCounters.ptr[CTR_CONTACTS], &n, sizeof(n), cudaMemcpyHostToDevice, s));
	w->hostBodiesStale = true;
	w->aabbsValid = false;
	B3_TRY(launchPackSoA(w));
	const int saved = w->static0Index;
	w->static0Index = static0Index;
	int rc = 0;
	if (w->solverKind == B3B200_SOLVER_JACOBI)
		rc = launchJacobi(w);
	else
	{
		rc = launchSolverSetup(w);
		if (rc == 0) rc = launchSolverIterate(w);
	}
	w->static0Index = saved;
	if (rc) return rc;
	B3_TRY(syncAoS(w));
	B3_CUDA_CHECK(cudaMemcpyAsync(bodies, w->dBodiesAoS.ptr, sizeof(b3b200_rigid_body) * (size_t)numBodies, cudaMemcpyDefault, s));
	B3_CUDA_CHECK(cudaStreamSynchronize(s));
	return 0;
}

extern "C" int b3b200_set_contacts(b3b200_world* w, const b3b200_contact4* src, int numContacts)
{
	W_UPLOADED(w);
	if (numContacts < 0 || numContacts > w->cfg.maxContactCapacity || (numContacts > 0 && !src)) return B3B200_ERR_INVALID;
	if (numContacts) B3_CUDA_CHECK(cudaMemcpyAsync(w->dContacts.ptr, src, sizeof(b3b200_contact4) * numContacts, cudaMemcpyHostToDevice, w->stream));
	unsigned int n = (unsigned int)numContacts;
	B3_CUDA_CHECK(cudaMemcpyAsync(&w->dCounters.ptr[CTR_CONTACTS], &n, sizeof(n), cudaMemcpyHostToDevice, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}
extern "C" int b3b200_get_constraints(b3b200_world* w, b3b200_constraint4* dst, int capacity, int* numConstraints)
{
	W_UPLOADED(w);
	if (!numConstraints || capacity < 0) return B3B200_ERR_INVALID;
	std::vector<b3b200_constraint4> rows;
	std::vector<int> off;
	B3_TRY(exportConstraints(w, rows, off));
	*numConstraints = (int)rows.size();
	const int m = std::min((int)rows.size(), capacity);
	if (m > 0 && dst) memcpy(dst, rows.data(), sizeof(b3b200_constraint4) * (size_t)m);
	return 0;
}
extern "C" int b3b200_get_batches(b3b200_world* w, int* batchOffsets, int capacity, int* numBatches)
{
	W_UPLOADED(w);
	if (!numBatches || capacity < 0) return B3B200_ERR_INVALID;
	std::vector<b3b200_constraint4> rows;
	std::vector<int> off;
	B3_TRY(exportConstraints(w, rows, off));
	const int nb = (int)off.size() - 1;
	*numBatches = nb;
	if (batchOffsets)
		for (int i = 0; i <= nb && i < capacity; i++) batchOffsets[i] = off[i];
	return 0;
}
extern "C" int b3b200_get_counters(b3b200_world* w, int* dst8)
{
	W_CHECK_KEEP(w);
	if (!dst8) return B3B200_ERR_INVALID;
	unsigned int c[CTR_COUNT];
	B3_TRY(readCounters(w, c));
	for (int i = 0; i < 8; i++) dst8[i] = (int)c[i];
	dst8[7] = (int)c[CTR_SURVIVORS];
	return 0;
}
extern "C" int b3b200_get_work_counters(b3b200_world* w, int* dst, int n)
{
	W_CHECK_KEEP(w);
	if (!dst || n < 0 || n > CTR_COUNT) return B3B200_ERR_INVALID;
	unsigned int c[CTR_COUNT];
	B3_TRY(readCounters(w, c));
	for (int i = 0; i < n; i++) dst[i] = (int)c[i];
	return 0;
}
// development aid (not part of include/b3b200.h): arm (dst == nullptr) or read back the solver kernel's globaltimer stamps
extern "C" int b3b200_debug_solver_probe(b3b200_world* w, unsigned long long* dst, int n)
{
	W_CHECK(w);
	if (!dst)
	{
		B3_TRY(w->dSolverProbe.reserve(512));
		B3_CUDA_CHECK(cudaMemsetAsync(w->dSolverProbe.ptr, 0, 512 * sizeof(unsigned long long), w->stream));
		return 0;
	}
	if (!w->dSolverProbe.ptr || n > 512) return B3B200_ERR_STATE;
	B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->dSolverProbe.ptr, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}
extern "C" int b3b200_enable_stage_timing(b3b200_world* w, int enable)
{
	if (!w) return B3B200_ERR_INVALID;
	w->timing = enable != 0;
	return 0;
}
extern "C" int b3b200_stage_timings(b3b200_world* w, float* ms8)
{
	if (!w || !ms8) return B3B200_ERR_INVALID;
	for (int i = 0; i < 8; i++) ms8[i] = w->stageMs[i];
	return 0;
}
extern "C" int b3b200_device_buffer(b3b200_world* w, int which, void** p)
{
	W_UPLOADED(w);
	if (!p) return B3B200_ERR_INVALID;
	switch (which)
	{
		case B3B200_BUF_BODIES:
			B3_TRY(syncAoS(w));
			*p = w->dBodiesAoS.ptr;
			return 0;
		case B3B200_BUF_AABBS:
			*p = w->bp.aabbs.ptr;
			return 0;
		case B3B200_BUF_PAIRS:
			*p = w->bp.pairs.ptr;
			return 0;
		case B3B200_BUF_CONTACTS:
			*p = w->dContacts.ptr;
			return 0;
		case B3B200_BUF_INERTIAS:
			*p = w->dInertias.ptr;
			return 0;
	}
	return B3B200_ERR_INVALID;
}

template <typename T>
static int copyTable(const std::vector<T>& v, void* dst, int capacity, int* count)
{
	*count = (int)v.size();
	int m = std::min((int)v.size(), capacity);
	if (dst && m > 0) memcpy(dst, v.data(), sizeof(T) * (size_t)m);
	return 0;
}
extern "C" int b3b200_get_table(b3b200_world* w, int which, void* dst, int capacity, int* count)
{
	if (!w || !count || capacity < 0) return B3B200_ERR_INVALID;
	switch (which)
	{
		case B3B200_TBL_COLLIDABLES:
			return copyTable(w->collidables, dst, capacity, count);
		case B3B200_TBL_LOCAL_AABBS:
			return copyTable(w->localAabbs, dst, capacity, count);
		case B3B200_TBL_CONVEX:
			return copyTable(w->convex, dst, capacity, count);
		case B3B200_TBL_VERTICES:
			return copyTable(w->vertices, dst, capacity, count);
		case B3B200_TBL_UNIQUE_EDGES:
			return copyTable(w->uniqueEdges, dst, capacity, count);
		case B3B200_TBL_FACES:
			return copyTable(w->faces, dst, capacity, count);
		case B3B200_TBL_INDICES:
			return copyTable(w->indices, dst, capacity, count);
		case B3B200_TBL_CHILD_SHAPES:
			return copyTable(w->childShapes, dst, capacity, count);
		case B3B200_TBL_BVH_INFOS:
			return copyTable(w->bvhInfos, dst, capacity, count);
		case B3B200_TBL_BVH_NODES:
			return copyTable(w->bvhNodes, dst, capacity, count);
		case B3B200_TBL_BVH_SUBTREES:
			return copyTable(w->bvhSubtrees, dst, capacity, count);
		case B3B200_TBL_BODIES:
			if (w->device >= 0 && w->uploaded && w->numBodies > 0)
			{
				// "last written": the AoS buffer as b3b200_write_bodies left it is not kept on the host any more
				B3_CUDA_CHECK(cudaSetDevice(w->device));
				B3_TRY(syncAoS(w));
				B3_CUDA_CHECK(cudaMemcpyAsync(w->bodies.data(), w->dBodiesAoS.ptr, sizeof(b3b200_rigid_body) * (size_t)w->numBodies, cudaMemcpyDeviceToHost, w->stream));
				B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
				w->hostBodiesStale = false;
			}
			return copyTable(w->bodies, dst, capacity, count);
		case B3B200_TBL_INERTIAS:
			return copyTable(w->inertias, dst, capacity, count);
	}
	return B3B200_ERR_INVALID;
}

extern "C" int b3b200_device_to_host(void* dstHost, const void* srcDevice, unsigned long long bytes, int device)
{
	if (!dstHost || !srcDevice) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(device));
	B3_CUDA_CHECK(cudaDeviceSynchronize());
	B3_CUDA_CHECK(cudaMemcpy(dstHost, srcDevice, (size_t)bytes, cudaMemcpyDeviceToHost));
	return 0;
}
