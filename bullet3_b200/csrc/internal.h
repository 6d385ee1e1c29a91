// internal.h -- host-side state behind the C ABI handles.
#pragma once
#include <vector>
#include "common.cuh"
#include "primitives.cuh"
#include "hull.h"

namespace b3b200
{
// device-side counters shared by the stages of one step
enum Counter
{
	CTR_PAIRS = 0,
	CTR_CONTACTS = 1,
	CTR_BATCHES = 2,
	CTR_COLOUR_ROUNDS = 3,
	CTR_OVERFLOW = 4,
	CTR_COMPOUND_PAIRS = 5,
	CTR_CONCAVE_PAIRS = 6,
	CTR_UNCOLOURED = 7,
	CTR_SURVIVORS = 8,
	CTR_OVERLAPS = 9,
	CTR_HALO = 10,
	CTR_CONCAVE_SURVIVORS = 11,
	CTR_CURSOR_SAT = 12,  // dynamic work distribution cursors of the warp-per-item kernels (12..14 are cleared together)
	CTR_CURSOR_CLIP = 13,
	CTR_CURSOR_CONCAVE = 14,
	CTR_MESH_PAIRS = 15,  // broadphase pairs with a trimesh as A (listed by npCullKernel)
	CTR_SMALL_ITEMS = 16,  // small x small hull items (thread-per-item kernel): box-like pairs, filled from the front of the list
	CTR_SMALL_ITEMS_BACK = 17,  // the other small pairs, filled from the back of the same list (cleared together with 16)
	CTR_CONCAVE_SURVIVORS_BACK = 19,  // trimesh items with a larger hull B (warp-per-item kernel), filled from the back of the survivor list
	CTR_CLIP_FALLBACK = 18,  // overlapping items whose faces are too large for the thread-per-item clip (cleared together with 16)
	CTR_HALO_RIGHT = 20,  // halo records packed for the right neighbour (CTR_HALO: left) in the C-driven slab step
	CTR_COUNT = 24
};
enum OverflowBits
{
	OVF_PAIRS = 1,
	OVF_CONTACTS = 2,
	OVF_BATCHES = 4,
	OVF_COMPOUND = 8,
	OVF_CONCAVE = 16,
	OVF_HALO = 32
};

// ------------------------------------------------------------------ broadphase
// Owns the world-space AABB array (creation order, min.w = user handle) and the
// pair buffer, like b3GpuSapBroadphase / b3GpuGridBroadphase
// (b3GpuSapBroadphase.h:14-141, b3GpuGridBroadphase.h:7-78).
struct Broadphase
{
	int kind = B3B200_BP_GRID;
	int device = 0;
	cudaStream_t stream = 0;
	bool ownStream = false;
	int maxProxies = 0;
	int maxPairs = 0;

	std::vector<b3b200_aabb> aabbsCPU;  // all proxies, creation order
	std::vector<int> smallIdx, largeIdx;  // indices into aabbsCPU

	DevBuf<b3b200_aabb> aabbs;
	DevBuf<int> smallMap, largeMap;
	int numSmall = 0, numLarge = 0, numAabbs = 0;
	// batched independent worlds (set by the owning World at upload; nullptr = one world): world id per proxy, and the large
	// proxies grouped by world (largeMap) with their first index per world
	const int* worldOf = nullptr;
	DevBuf<int> largeStart;

	DevBuf<b3b200_int4> pairs;
	DevBuf<unsigned int> counters;  // CTR_COUNT (own copy when stand-alone; world shares its own)
	unsigned int* ctr = nullptr;    // -> counters actually used

	// scratch
	DevBuf<unsigned int> keys, vals;
	DevBuf<b3b200_aabb> sortedAabbs;
	DevBuf<int> wideList;          // small AABBs much wider than the rest: kept out of the grid, tested against everything (broadphase.cu)
	DevBuf<int> cellStart;         // 128^3 + 4: first sorted index of every cell (dense exclusive scan of the counts)
	DevBuf<int> cellCnt;           // bodies per cell
	DevBuf<unsigned int> scanTotals;  // scratch of the multi-CTA scan
	DevBuf<float> scalars;         // [0]=maxExtent bits/cellSize ... see broadphase.cu
	RadixSortTemp sortTmp;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	float lastMs = 0.f;

	int init(int kind, int device, cudaStream_t stream, int maxProxies, int maxPairs);
	void destroy();
	int createProxy(const float* mn, const float* mx, int userPtr, bool large);
	int writeAabbs();           // host -> device, (re)allocates scratch
	int calculatePairs(int maxPairsNow);  // async on stream
	int reset();
};

// slab decomposition driven through the C ABI (halo.cu): NCCL communicator + fixed-capacity message buffers of one rank
struct SlabState
{
	bool active = false;
	void* comm = nullptr;  // ncclComm_t
	int axis = 0, rank = 0, numRanks = 1, numOwned = 0, firstGhostSlot = 0, maxGhosts = 0, globalIdBase = 0;
	float lo = 0.f, hi = 0.f, margin = 0.f;
	bool hasLeft = false, hasRight = false;
	DevBuf<unsigned char> sendBuf[2], recvBuf[2];  // [left, right]: header record + maxGhosts records
};

// ------------------------------------------------------------------ world
struct World
{
	b3b200_config cfg;
	int device = 0;
	cudaStream_t stream = 0;
	bool ownStream = false;
	bool uploaded = false;
	bool everUploaded = false;  // bodies [0, numBodies) have device state that an upload must not rewind
	bool aabbsValid = false;  // world AABBs on device match the current poses

	// batched independent worlds (b3b200_set_current_world): bodies only collide inside their own world
	int currentWorld = 0, numWorlds = 1;
	std::vector<int> bodyWorld;  // per body
	DevBuf<int> dBodyWorld;
	int worldDynBodies = 0;  // > 0: every world has exactly this many dynamic bodies (solver blocks then hold whole worlds)

	// settings
	float gravity[3] = {0.f, -9.8f, 0.f};  // b3GpuRigidBodyPipeline.cpp:92
	float angularDamping = 0.99f;          // b3GpuRigidBodyPipeline.cpp:469
	int solverKind = B3B200_SOLVER_PGS;
	int solverIterations = 4;  // b3GpuPgsContactSolver.cpp:1049
	int solverColouring = 1;  // batch assignment: 0 = priority rounds (reproducible for a given contact array), 1 = single-pass first fit with atomics (solver.cu)
	float clipMinDist = -1e30f, clipMaxDist = 0.02f;  // satClipHullContacts.cl:916-917
	int static0Index = -1;     // b3GpuNarrowPhase.cpp:861-864

	// host-side shape tables (b3GpuNarrowPhaseInternalData.h:24-86)
	std::vector<b3b200_collidable> collidables;
	std::vector<b3b200_aabb> localAabbs;  // per collidable
	std::vector<b3b200_convex_polyhedron> convex;
	std::vector<b3b200_float4> vertices, uniqueEdges;
	std::vector<b3b200_face> faces;
	std::vector<int> indices;
	std::vector<b3b200_child_shape> childShapes;
	std::vector<b3b200_bvh_info> bvhInfos;
	std::vector<b3b200_bvh_node> bvhNodes;
	std::vector<b3b200_bvh_subtree> bvhSubtrees;
	// concave trimeshes: this build's own acceleration structure (shapes.cu): a binary AABB tree per mesh with
	// exact float boxes.  Node = 2 x float4 {min.xyz, bits(left child | first triangle slot)}, {max.xyz, bits(0 |
	// triangle count)}; children of a node are adjacent; leaves index meshTris (original triangle numbers).
	std::vector<b3b200_float4> meshNodes;
	std::vector<int> meshTris;
	std::vector<b3b200_int4> meshInfos;  // per bvhInfos entry: {first node (in nodes), node count, first slot in meshTris, triangle count}
	// joints (b3GpuRigidBodyPipeline: m_cpuConstraints / m_gpuConstraints, m_constraintUid)
	std::vector<b3b200_generic_constraint> joints;
	int jointUid = 0;
	bool jointsDirty = false;         // host copy newer than the device copy
	bool jointBatchesDirty = false;   // constraint set changed: rebuild the batches (b3GpuPgsConstraintSolver::recomputeBatches)
	std::vector<int> jointOrder, jointBatchOffset;
	// host-side bodies
	std::vector<b3b200_rigid_body> bodies;
	std::vector<b3b200_inertia> inertias;

	// device shape tables
	DevBuf<b3b200_collidable> dCollidables;
	DevBuf<b3b200_aabb> dLocalAabbs;
	DevBuf<b3b200_convex_polyhedron> dConvex;
	DevBuf<float4> dVertices, dUniqueEdges;
	DevBuf<b3b200_face> dFaces;
	DevBuf<int> dIndices;
	DevBuf<b3b200_child_shape> dChildShapes;
	DevBuf<float4> dChildSpheres;  // bounding sphere of every child hull in its compound's frame (built at upload)
	DevBuf<b3b200_bvh_info> dBvhInfos;
	DevBuf<b3b200_bvh_node> dBvhNodes;
	DevBuf<b3b200_bvh_subtree> dBvhSubtrees;
	DevBuf<float4> dMeshNodes;
	DevBuf<int> dMeshTris;
	DevBuf<int4> dMeshInfos;

	// device body state.  AoS (reference layout) is the boundary format; the
	// step runs on the SoA split below: pose = {pos.xyz, invMass | quat} and
	// vel = {linVel | angVel}, one 32-byte sector each.
	int numBodies = 0;
	DevBuf<b3b200_rigid_body> dBodiesAoS;
	DevBuf<b3b200_inertia> dInertias;
	DevBuf<float4> dPose;  // 2 float4 per body
	DevBuf<float4> dVel;   // 2 float4 per body
	DevBuf<int> dCollidableIdx;
	DevBuf<int> dGhostGlobalId;  // slab mode: global id mirrored by each ghost slot (-1 = parked / owned)
	DevBuf<int> dHaloSlots;  // slot lists of emigrate / adopt
	bool haloIdsSet = false;  // dGhostGlobalId holds the global id of every slot (b3b200_halo_set_ids)
	SlabState slab;
	bool soaDirty = false;  // SoA is newer than AoS
	bool hostBodiesStale = false;  // the host mirror `bodies` is older than the device AoS (after b3b200_write_bodies)
	bool hasConcave = false;  // any SHAPE_CONCAVE_TRIMESH collidable registered (enables the concave kernels)
	bool hasPlanes = false;  // any SHAPE_PLANE collidable registered (enables the primitive-contact kernel)

	Broadphase bp;

	// narrowphase output
	DevBuf<b3b200_contact4> dContacts;
	DevBuf<unsigned int> dCounters;  // CTR_COUNT
	DevBuf<b3b200_int4> dCompoundPairs;
	DevBuf<int4> dConcavePairs;  // (pair, triangle, child shape of B or -1, 0) work items of the concave path
	DevBuf<int4> dConcaveSurvivors;  // ... that passed the quick reject
	DevBuf<int4> dSurvivors;     // work items (pair, childA, childB, 0) that passed the quick SAT reject
	DevBuf<int4> dSmallItems;    // ... of which both hulls are small (boxes, tetrahedra): thread-per-item kernel
	DevBuf<int4> dOverlapPairs;  // work items with a penetrating SAT result
	DevBuf<float4> dOverlapSep;  // their minimum-penetration axes

	// solver (solver.cu): Morton-order partition of the dynamic bodies into blocks, per-step contact classification and
	// colouring, rows as 32-row structure-of-arrays tiles
	DevBuf<b3b200_constraint4> dConstraints;  // b3ContactConstraint4 rows in contact order (Jacobi solver only, allocated on first use)
	DevBuf<int> dBodyLoc;                  // per body: (block << 12) | slot, -1 = static
	DevBuf<unsigned int> dPartKeys, dPartVals, dPartBounds;  // Morton keys | bodies in Morton order | {min[3], max[3], numDynamic}
	RadixSortTemp partSortTmp;
	int partS = 0, partBlocksMax = 0, partAge = 0, partInterval = 8, partBodies = -1;
	bool partValid = false;
	DevBuf<unsigned long long> dBodyMask;  // colours of a body's cross contacts (2 words / body)
	DevBuf<unsigned int> dCtaBase;        // per setup CTA and block: start of the CTA's contacts inside the block's list (classify -> scatter)
	DevBuf<unsigned int> dBodyPrio;       // max pending priority per body (reproducible colouring; 2 words / body)
	DevBuf<int> dContactBlock, dContactColour;
	DevBuf<unsigned int> dContactSlots, dBlockList, dCrossList;
	DevBuf<int2> dContactPair;
	DevBuf<int> dTileSrc;
	DevBuf<unsigned int> dSolverScratch, dBlockStart, dBlockTileBase, dBlockTileOff, dCrossTileOff;
	DevBuf<int> dBlockStatics;
	DevBuf<float4> dTilesN, dTilesF;
	DevBuf<float4> dSortedState;          // body state in block order: lin | ang | inertia (upper triangle, 2 float4) | pos
	DevBuf<unsigned char> dSortedBoundary;
	unsigned int* solverMisc = nullptr;    // -> misc words of dSolverScratch once a setup has run
	bool solverAttrSet = false;
	DevBuf<unsigned long long> dSolverProbe;  // development aid (b3b200_debug_solver_probe)
	DevBuf<unsigned int> dGridBarrier;    // software grid barrier state
	// joints
	DevBuf<b3b200_generic_constraint> dJoints;
	DevBuf<int> dJointOrder, dJointBatchOffset, dJointNumRows;
	DevBuf<float4> dJointRows, dJointDelta;
	// raycast
	DevBuf<float4> dRays, dRayHits;
	DevBuf<unsigned long long> dRayBest;
	// per-call linear BVH of the ray path (raycast.cu): Morton keys / order, leaf + chunk + super boxes, centre bounds
	DevBuf<unsigned int> dRayKeys, dRayOrder, dRayBounds;
	DevBuf<float4> dRayTree;
	RadixSortTemp raySortTmp;
	int rayAccel = -1;
	// jacobi
	DevBuf<unsigned int> dBodyCount, dBodyOffset;
	DevBuf<float4> dDeltaLin, dDeltaAng;
	DevBuf<unsigned int> dContactSlot;  // 2 per contact

	int smCount = 148;

	// CUDA graphs of one whole step (world.cu, stepGraphed): the launch sequence of a step depends on the host only through
	// (world AABBs valid?, partition due?), so there is one captured graph per such key.  Any API call that can change launch
	// parameters (registration, upload, settings, stand-alone stage calls ...) drops them (graphEpoch); a step whose capture saw
	// a device (re)allocation is not kept.  state: 0 = this key has not run yet (run eagerly: sizes the buffers), 1 = ran
	// eagerly (capture next), 2 = graph ready
	struct StepGraph
	{
		int state = 0;
		cudaGraphExec_t exec = nullptr;
		long long launches = 0;
		float dt = 0.f;
	};
	StepGraph stepGraphs[4];

	// pipelined host stepping (b3b200_step_host_async): two staging slots, an upload and a download stream of their own
	struct HostPipe
	{
		cudaStream_t h2d = nullptr, d2h = nullptr;
		DevBuf<b3b200_rigid_body> stage[2];
		cudaEvent_t inReady[2] = {nullptr, nullptr}, stepDone[2] = {nullptr, nullptr}, outDone[2] = {nullptr, nullptr};
		bool stepRecorded[2] = {false, false}, outPending[2] = {false, false};
		unsigned long long calls = 0;
	} pipe;
	int useGraphs = 1;  // B3B200_GRAPHS=0 / b3b200_set_step_graphs(w, 0): every step launches its kernels one by one
	void dropStepGraphs();

	// timing
	bool timing = false;
	cudaEvent_t ev[8] = {nullptr};
	cudaEvent_t evSat[2] = {nullptr, nullptr};  // around satKernel -> stageMs[7]
	// narrowphase branches that share no buffer (small pairs | SAT -> clip | trimesh chain | primitives) run on side streams
	// forked from / joined into `stream` with events, so one branch's tail overlaps the others' bulk
	cudaStream_t npStream[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t evNpFork[2] = {nullptr, nullptr}, evNpJoin[3] = {nullptr, nullptr, nullptr};
	bool npOverlap = false;  // measured: 8.10 vs 8.12-8.16 ms per step on the bench scene (each branch already fills the GPU); B3B200_NP_OVERLAP=1 turns it on
	float stageMs[8] = {0.f};

	int init(const b3b200_config* cfg, int device, cudaStream_t stream);
	void destroy();
};

constexpr int MAX_BATCHES = 128;  // B3_MAX_NUM_BATCHES (b3Solver.h:33-41)

// host helpers shared by world.cu / shapes.cu
void transformAabbHost(const float* lmn, const float* lmx, float margin, const float* pos, const float* orn, float* outMin, float* outMax);
int allocateCollidable(World* w);

// stage launchers (each async on w->stream)
int launchPackSoA(World* w);    // AoS -> SoA
int launchUnpackSoA(World* w);  // SoA -> AoS
int launchPackSoAFrom(World* w, const b3b200_rigid_body* src);
int launchUnpackSoATo(World* w, b3b200_rigid_body* dst);
int launchUpdateAabbs(World* w);
int launchIntegrate(World* w, float dt, bool alsoAabbs);
int launchNarrowphase(World* w);
int launchConcave(World* w, cudaStream_t s);  // concave.cu; called by launchNarrowphase when a trimesh is registered
int launchSolveJoints(World* w);  // joints.cu
int launchSolverSetup(World* w);
int launchSolverIterate(World* w);
int launchJacobi(World* w);
int slabExchange(World* w);  // halo.cu
void slabDestroy(World* w);
int exportConstraints(World* w, std::vector<b3b200_constraint4>& out, std::vector<int>& batchOffsets);  // solver.cu

}  // namespace b3b200

struct b3b200_world : b3b200::World
{
};
struct b3b200_broadphase : b3b200::Broadphase
{
};
