/*
 * b3b200_types.h -- plain-C POD mirror of the Bullet3 GPU rigid-body ABI.
 *
 * Every struct here has the same size and field offsets as the reference
 * struct it names (verified by the static asserts at the bottom and by
 * tests/test_abi.py).  The C++ drop-in headers under
 * bullet3_b200/csrc/host/ typedef these to the reference names.
 *
 * Reference layouts (paths relative to the bullet3 tree):
 *   b3RigidBodyData / b3InertiaData  src/Bullet3Collision/NarrowPhaseCollision/shared/b3RigidBodyData.h:10-29
 *   b3Collidable / b3GpuChildShape   src/Bullet3Collision/NarrowPhaseCollision/shared/b3Collidable.h:8-66
 *   b3ConvexPolyhedronData/b3GpuFace src/Bullet3Collision/NarrowPhaseCollision/shared/b3ConvexPolyhedronData.h:9-34
 *   b3Aabb / b3SapAabb               src/Bullet3Collision/BroadPhaseCollision/shared/b3Aabb.h:10-22
 *   b3Int4 (b3BroadphasePair)        src/Bullet3Common/shared/b3Int4.h
 *   b3Contact4Data                   src/Bullet3Collision/NarrowPhaseCollision/shared/b3Contact4Data.h:8-34
 *   b3ContactConstraint4             src/Bullet3Dynamics/shared/b3ContactConstraint4.h:8-29
 *   b3Config                         src/Bullet3Collision/NarrowPhaseCollision/b3Config.h:4-37
 *   b3QuantizedBvhNodeData           src/Bullet3Collision/NarrowPhaseCollision/shared/b3QuantizedBvhNodeData.h:14-40
 *   b3BvhSubtreeInfoData             src/Bullet3Collision/NarrowPhaseCollision/shared/b3BvhSubtreeInfoData.h
 *   b3BvhInfo                        src/Bullet3OpenCL/NarrowphaseCollision/b3BvhInfo.h:6-15
 */
#ifndef B3B200_TYPES_H
#define B3B200_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
#define B3B200_ALIGN16 alignas(16)
#else
#define B3B200_ALIGN16 _Alignas(16)
#endif

typedef struct B3B200_ALIGN16 b3b200_float4
{
	float x, y, z, w;
} b3b200_float4;

typedef struct B3B200_ALIGN16 b3b200_int4
{
	int x, y, z, w;
} b3b200_int4;

typedef struct B3B200_ALIGN16 b3b200_mat3x3
{
	b3b200_float4 row[3];
} b3b200_mat3x3;

enum b3b200_shape_type
{
	B3B200_SHAPE_HEIGHT_FIELD = 1,
	B3B200_SHAPE_CONVEX_HULL = 3,
	B3B200_SHAPE_PLANE = 4,
	B3B200_SHAPE_CONCAVE_TRIMESH = 5,
	B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS = 6,
	B3B200_SHAPE_SPHERE = 7
};

typedef struct B3B200_ALIGN16 b3b200_rigid_body
{
	b3b200_float4 pos;
	b3b200_float4 quat;
	b3b200_float4 linVel;
	b3b200_float4 angVel;
	int collidableIdx;
	float invMass;
	float restitution;
	float friction;
} b3b200_rigid_body;

typedef struct B3B200_ALIGN16 b3b200_inertia
{
	b3b200_mat3x3 invInertiaWorld;
	b3b200_mat3x3 initInvInertia;
} b3b200_inertia;

typedef struct b3b200_collidable
{
	union {
		int numChildShapes;
		int bvhIndex;
	};
	union {
		float radius;
		int compoundBvhIndex;
	};
	int shapeType;
	union {
		int shapeIndex;
		float height;
	};
} b3b200_collidable;

typedef struct B3B200_ALIGN16 b3b200_child_shape
{
	b3b200_float4 childPosition;
	b3b200_float4 childOrientation;
	union {
		int shapeIndex;
		int capsuleAxis;
	};
	union {
		float radius;
		int numChildShapes;
	};
	union {
		float height;
		int collidableShapeIndex;
	};
	int shapeType;
} b3b200_child_shape;

typedef struct B3B200_ALIGN16 b3b200_face
{
	b3b200_float4 plane;
	int indexOffset;
	int numIndices;
	int pad1;
	int pad2;
} b3b200_face;

typedef struct B3B200_ALIGN16 b3b200_convex_polyhedron
{
	b3b200_float4 localCenter;
	b3b200_float4 extents;
	b3b200_float4 mC;
	b3b200_float4 mE;
	float radius;
	int faceOffset;
	int numFaces;
	int numVertices;
	int vertexOffset;
	int uniqueEdgesOffset;
	int numUniqueEdges;
	int unused;
} b3b200_convex_polyhedron;

typedef struct B3B200_ALIGN16 b3b200_aabb
{
	union {
		float min[4];
		int minIndices[4];
	};
	union {
		float max[4];
		int signedMaxIndices[4];
	};
} b3b200_aabb;

/* per-pair output of b3b200_mpr_penetration: what b3MprPenetration (shared/b3MprPenetration.h:825-888) returned for the pair */
typedef struct b3b200_mpr_result
{
	int result; /* 0 = penetrating (depth / dir / pos valid), -1 = not, -2 = pair skipped (not two hulls, or both static) */
	float depth;
	float dir[3];
	float pos[3];
} b3b200_mpr_result;

typedef struct B3B200_ALIGN16 b3b200_contact4
{
	b3b200_float4 worldPosB[4]; /* xyz = point on B, w = depth */
	b3b200_float4 worldNormalOnB; /* w = number of points (as float) */
	unsigned short restitutionCmp;
	unsigned short frictionCmp;
	int batchIdx;
	int bodyAPtrAndSignBit;
	int bodyBPtrAndSignBit;
	int childIndexA;
	int childIndexB;
	int unused1;
	int unused2;
} b3b200_contact4;

typedef struct B3B200_ALIGN16 b3b200_constraint4
{
	b3b200_float4 linear; /* normal, w = friction coefficient */
	b3b200_float4 worldPos[4];
	b3b200_float4 center;
	float jacCoeffInv[4];
	float b[4];
	float appliedRambdaDt[4];
	float fJacCoeffInv[2];
	float fAppliedRambdaDt[2];
	unsigned int bodyA;
	unsigned int bodyB;
	int batchIdx;
	unsigned int paddings;
} b3b200_constraint4;

typedef struct b3b200_config
{
	int maxConvexBodies;
	int maxConvexShapes;
	int maxBroadphasePairs;
	int maxContactCapacity;
	int compoundPairCapacity;
	int maxVerticesPerFace;
	int maxFacesPerShape;
	int maxConvexVertices;
	int maxConvexIndices;
	int maxConvexUniqueEdges;
	int maxCompoundChildShapes;
	int maxTriConvexPairCapacity;
} b3b200_config;

typedef struct b3b200_bvh_node
{
	unsigned short quantizedAabbMin[3];
	unsigned short quantizedAabbMax[3];
	int escapeIndexOrTriangleIndex;
} b3b200_bvh_node;

typedef struct b3b200_bvh_subtree
{
	unsigned short quantizedAabbMin[3];
	unsigned short quantizedAabbMax[3];
	int rootNodeIndex;
	int subtreeSize;
	int padding[3];
} b3b200_bvh_subtree;

typedef struct B3B200_ALIGN16 b3b200_bvh_info
{
	b3b200_float4 aabbMin;
	b3b200_float4 aabbMax;
	b3b200_float4 quantization;
	int numNodes;
	int numSubTrees;
	int nodeOffset;
	int subTreeOffset;
} b3b200_bvh_info;

/* b3GpuGenericConstraint (src/Bullet3OpenCL/RigidBody/b3GpuGenericConstraint.h:73-127), 80 bytes */
#define B3B200_CONSTRAINT_P2P 3   /* B3_GPU_POINT2POINT_CONSTRAINT_TYPE */
#define B3B200_CONSTRAINT_FIXED 4 /* B3_GPU_FIXED_CONSTRAINT_TYPE */
#define B3B200_CONSTRAINT_FLAG_ENABLED 1
typedef struct B3B200_ALIGN16 b3b200_generic_constraint
{
	int constraintType;
	int rbA;
	int rbB;
	float breakingImpulseThreshold;
	b3b200_float4 pivotInA;
	b3b200_float4 pivotInB;
	b3b200_float4 relTargetAB;
	int flags;
	int uid;
	int padding[2];
} b3b200_generic_constraint;

/* b3RayInfo / b3RayHit (src/Bullet3Collision/NarrowPhaseCollision/b3RaycastInfo.h:7-23) */
typedef struct B3B200_ALIGN16 b3b200_ray_info
{
	b3b200_float4 from;
	b3b200_float4 to;
} b3b200_ray_info;

typedef struct B3B200_ALIGN16 b3b200_ray_hit
{
	float hitFraction;
	int hitBody;
	int hitResult1;
	int hitResult2;
	b3b200_float4 hitPoint;
	b3b200_float4 hitNormal;
} b3b200_ray_hit;

typedef struct b3b200_sort_data
{
	unsigned int key;
	unsigned int value;
} b3b200_sort_data;

#ifdef __cplusplus
static_assert(sizeof(b3b200_float4) == 16, "abi");
static_assert(sizeof(b3b200_rigid_body) == 80, "abi");
static_assert(sizeof(b3b200_inertia) == 96, "abi");
static_assert(sizeof(b3b200_collidable) == 16, "abi");
static_assert(sizeof(b3b200_child_shape) == 48, "abi");
static_assert(sizeof(b3b200_face) == 32, "abi");
static_assert(sizeof(b3b200_convex_polyhedron) == 96, "abi");
static_assert(sizeof(b3b200_aabb) == 32, "abi");
static_assert(sizeof(b3b200_int4) == 16, "abi");
static_assert(sizeof(b3b200_contact4) == 112, "abi");
static_assert(sizeof(b3b200_constraint4) == 176, "abi");
static_assert(sizeof(b3b200_config) == 48, "abi");
static_assert(sizeof(b3b200_bvh_node) == 16, "abi");
static_assert(sizeof(b3b200_bvh_subtree) == 32, "abi");
static_assert(sizeof(b3b200_bvh_info) == 64, "abi");
static_assert(sizeof(b3b200_sort_data) == 8, "abi");
static_assert(sizeof(b3b200_generic_constraint) == 80, "abi");
static_assert(sizeof(b3b200_ray_info) == 32, "abi");
static_assert(sizeof(b3b200_ray_hit) == 48, "abi");
#endif

/* slab decomposition of one scene over several GPUs, driven from C / C++ (no reference counterpart: Bullet3OpenCL is single-device) */
typedef struct b3b200_nccl_id
{
	char internal[128]; /* an ncclUniqueId */
} b3b200_nccl_id;
typedef struct b3b200_slab_config
{
	int axis;           /* slab axis: 0 x, 1 y, 2 z */
	float lo, hi;       /* this rank's slab along the axis (+-3e38 at the outer ranks) */
	float margin;       /* a body whose AABB reaches within `margin` of a slab face is mirrored on that neighbour */
	int numOwned;       /* bodies [0, numOwned) of the world are owned (static + dynamic + spare); the rest are ghost slots */
	int firstGhostSlot; /* first ghost slot: maxGhosts for the left neighbour (if any), then maxGhosts for the right one */
	int maxGhosts;      /* capacity per side = size of the fixed-size message */
	int globalIdBase;   /* global id of local body i = globalIdBase + i (unless b3b200_halo_set_ids gave every slot its id) */
	int rank, numRanks;
} b3b200_slab_config;

#endif /* B3B200_TYPES_H */
