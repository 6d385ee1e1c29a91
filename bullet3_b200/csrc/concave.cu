// concave.cu -- concave trimesh x convex hull / compound child contacts.
//
// Replaces the concave leg of GpuSatCollision::computeConvexConvexContactsGPUSAT
// (b3ConvexHullContact.cpp:3481-4040: bvhTraversalKernel, findConcaveSeparatingAxis*Kernel, clipFacesAndFindContacts,
// newContactReductionKernel + 6 host round trips) with two launches and no host synchronisation:
//   concaveCullKernel     one thread per broadphase pair whose A is a trimesh (listed by npCullKernel): walks this build's float AABB tree
//                         (shapes.cu) with B's world AABB and emits (pair, triangle, child) work items for every
//                         triangle whose exact AABB overlaps -- the same active set as the reference's quantized
//                         b3BvhTraversal (shared/b3BvhTraversal.h:11-122) followed by the exact triangle-AABB test at
//                         the top of b3FindConcaveSeparatingAxisKernel (shared/b3FindConcaveSatAxis.h:606-610).
//                         Like the reference, the tree is queried in mesh-local space with B's WORLD box.
//   concaveContactKernel  one warp per work item.  Arithmetic follows the reference's shared headers operation by
//                         operation: SAT over the 5-face triangle prism, the faces of B and all edge pairs
//                         (b3FindConcaveSatAxis.h:551-795), b3FindClippingFaces (:417-505),
//                         clipFacesAndFindContactsKernel (shared/b3ClipFaces.h:66-169) and
//                         b3NewContactReductionKernel (shared/b3NewContactReduction.h:10-173).
// Lanes own separating-axis candidates; ties break towards the lower axis index (= the reference's sequential
// "first strict minimum").
#include "internal.h"

namespace b3b200
{
namespace
{
constexpr int CC_THREADS = 128;
constexpr int CC_WARPS = CC_THREADS / 32;
constexpr int CC_MAX_POLY = 64;  // vertexFaceCapacity (b3ConvexHullContact.cpp:3481)
// triangle x SMALL hull (boxes, tetrahedra, box children of compounds): one thread per item (concaveSmallKernel)
constexpr int CCT_VERTS = 8, CCT_FACES = 6, CCT_EDGES = 6;
constexpr int CCT_POLY = 16;  // a face of <= 8 vertices clipped by <= 3 planes
#define FULL 0xffffffffu

struct CcArgs
{
	const b3b200_int4* pairs;
	unsigned int* ctr;
	const float4* pose;
	const int* coll;
	const b3b200_collidable* collidables;
	const b3b200_convex_polyhedron* convex;
	const float4* vertices;
	const float4* uniqueEdges;
	const b3b200_face* faces;
	const int* indices;
	const b3b200_child_shape* childShapes;
	const b3b200_aabb* aabbs;
	const float4* meshNodes;
	const int* meshTris;
	const int4* meshInfos;
	b3b200_contact4* contacts;
	int maxContacts;
	int maxItems;
};

struct HullRef
{
	float4 localCenter;
	float radius;  // inscribed radius about localCenter (set at registration, world.cu)
	int faceOffset, numFaces, numVertices, vertexOffset, uniqueEdgesOffset, numUniqueEdges;
};
B3_D HullRef loadHull(const b3b200_convex_polyhedron* __restrict__ convex, int shapeIndex)
{
	const b3b200_convex_polyhedron* h = &convex[shapeIndex];
	HullRef r;
	r.localCenter = __ldg(reinterpret_cast<const float4*>(&h->localCenter));
	const int4* t = reinterpret_cast<const int4*>(&h->radius);  // radius, faceOffset, numFaces, numVertices
	const int4 a = __ldg(t), b = __ldg(t + 1);                   // vertexOffset, uniqueEdgesOffset, numUniqueEdges, unused
	r.radius = __int_as_float(a.x);
	r.faceOffset = a.y;
	r.numFaces = a.z;
	r.numVertices = a.w;
	r.vertexOffset = b.x;
	r.uniqueEdgesOffset = b.y;
	r.numUniqueEdges = b.z;
	return r;
}
B3_D void warpArgMin(float& d, int& k)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const float od = __shfl_xor_sync(FULL, d, o);
		const int ok = __shfl_xor_sync(FULL, k, o);
		if (ok >= 0 && (k < 0 || od < d || (od == d && ok < k)))
		{
			d = od;
			k = ok;
		}
	}
}
B3_D void warpArgMax(float& d, int& k)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const float od = __shfl_xor_sync(FULL, d, o);
		const int ok = __shfl_xor_sync(FULL, k, o);
		if (ok >= 0 && (k < 0 || od > d || (od == d && ok < k)))
		{
			d = od;
			k = ok;
		}
	}
}
B3_D bool almostZero(const float4& v) { return !(fabsf(v.x) > 1e-6f || fabsf(v.y) > 1e-6f || fabsf(v.z) > 1e-6f); }
B3_D float4 lerp3(const float4& a, const float4& b, float t) { return mk4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, 0.f); }

// b3Project (shared/b3FindConcaveSatAxis.h:14-39) for the hull in the global tables
B3_D void projectHull(const HullRef& hull, const float4& pos, const float4& orn, const float4& dir, const float4* __restrict__ vertices, float& mn, float& mx)
{
	mn = FLT_MAX;
	mx = -FLT_MAX;
	const float4 localDir = quatRotate(quatInverse(orn), dir);
	const float offset = dot3(pos, dir);
	const float4* v = vertices + hull.vertexOffset;
	for (int i = 0; i < hull.numVertices; i++)
	{
		const float dp = dot3(__ldg(&v[i]), localDir);
		mn = fminf(mn, dp);  // == "if (dp < mn) mn = dp" up to the sign of a zero, which nothing downstream observes
		mx = fmaxf(mx, dp);
	}
	if (mn > mx)
	{
		const float t = mn;
		mn = mx;
		mx = t;
	}
	mn += offset;
	mx += offset;
}
// ... and for the triangle (3 private vertices)
B3_D void projectTri(const float4 (&vA)[3], const float4& pos, const float4& orn, const float4& dir, float& mn, float& mx)
{
	mn = FLT_MAX;
	mx = -FLT_MAX;
	const float4 localDir = quatRotate(quatInverse(orn), dir);
	const float offset = dot3(pos, dir);
#pragma unroll
	for (int i = 0; i < 3; i++)
	{
		const float dp = dot3(vA[i], localDir);
		mn = fminf(mn, dp);
		mx = fmaxf(mx, dp);
	}
	mn += offset;
	mx += offset;
}

// clipFaceGlobal (shared/b3ClipFaces.h:21-64), one polygon edge per lane, outputs compacted in order
B3_D int clipFaceGlobalWarp(const float4* in, int numIn, const float4& n, float eq, float4* out, int lane)
{
	int numOut = 0;
	for (int base = 0; base < numIn; base += 32)
	{
		const int ve = base + lane;
		int c = 0;
		float4 o0 = mk4(0, 0, 0), o1 = mk4(0, 0, 0);
		if (ve < numIn)
		{
			const float4 first = in[ve == 0 ? numIn - 1 : ve - 1];
			const float4 end = in[ve];
			const float ds = dot3(n, first) + eq;
			const float de = dot3(n, end) + eq;
			if (ds < 0)
			{
				if (de < 0)
					o0 = end;
				else
					o0 = lerp3(first, end, (ds * 1.f / (ds - de)));
				c = 1;
			}
			else if (de < 0)
			{
				o0 = lerp3(first, end, (ds * 1.f / (ds - de)));
				o1 = end;
				c = 2;
			}
		}
		const unsigned int m1 = __ballot_sync(FULL, c >= 1);
		const unsigned int m2 = __ballot_sync(FULL, c == 2);
		const unsigned int lt = (1u << lane) - 1u;
		const int pos = numOut + __popc(m1 & lt) + __popc(m2 & lt);
		if (c >= 1 && pos < CC_MAX_POLY) out[pos] = o0;
		if (c == 2 && pos + 1 < CC_MAX_POLY) out[pos + 1] = o1;
		numOut += __popc(m1) + __popc(m2);
	}
	__syncwarp();
	return numOut < CC_MAX_POLY ? numOut : CC_MAX_POLY;
}

// ---------------------------------------------------------------------------------------------------------
constexpr int CULL_WARPS = 8;    // warps per CTA of concaveCullKernel
constexpr int CULL_QUEUE = 256;  // frontier capacity per warp

// serial walk of one subtree by one lane (only used when a warp's frontier queue overflows): appends the overlapping
// leaves' triangles through `emit`
template <typename Emit>
B3_D void walkSubtree(const CcArgs& a, const int4& mesh, int root, const float4& qmn, const float4& qmx, Emit emit)
{
	int stack[48];
	int sp = 0;
	stack[sp++] = root;
	while (sp)
	{
		const int node = stack[--sp];
		const float4 n0 = __ldg(&a.meshNodes[2 * (mesh.x + node)]), n1 = __ldg(&a.meshNodes[2 * (mesh.x + node) + 1]);
		if (n0.x > qmx.x || n1.x < qmn.x || n0.y > qmx.y || n1.y < qmn.y || n0.z > qmx.z || n1.z < qmn.z) continue;
		const int count = __float_as_int(n1.w), first = __float_as_int(n0.w);
		if (count == 0)
		{
			if (sp + 2 <= 48)
			{
				stack[sp++] = first + 1;
				stack[sp++] = first;
			}
			continue;
		}
		for (int t = 0; t < count; t++) emit(first + t);
	}
}

// One WARP per trimesh pair (listed by npCullKernel, narrowphase.cu): level-synchronous walk of the mesh tree with a
// per-warp frontier queue in shared memory -- 32 nodes are tested per step instead of one dependent load chain per
// thread -- then the triangles of the overlapping leaves are tested one per lane.
__global__ void __launch_bounds__(CULL_WARPS * 32) concaveCullKernel(CcArgs a, const int* __restrict__ meshPairs, int maxMeshPairs, int4* __restrict__ items)
{
	__shared__ int queueAll[CULL_WARPS][2][CULL_QUEUE];
	__shared__ int leafAll[CULL_WARPS][CULL_QUEUE];  // first triangle slot | (count << 28)
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned int lt = (1u << lane) - 1u;
	int numMeshPairs = (int)a.ctr[CTR_MESH_PAIRS];
	if (numMeshPairs > maxMeshPairs) numMeshPairs = maxMeshPairs;
	for (int q = blockIdx.x * CULL_WARPS + warp; q < numMeshPairs; q += gridDim.x * CULL_WARPS)
	{
		__syncwarp();
		const int p = meshPairs[q];
		const int bodyA = a.pairs[p].x, bodyB = a.pairs[p].y;
		const int cA = a.coll[bodyA], cB = a.coll[bodyB];
		if (cA < 0 || cB < 0) continue;
		if (__ldg(&a.collidables[cA].shapeType) != B3B200_SHAPE_CONCAVE_TRIMESH) continue;  // only with the mesh as A (b3BvhTraversal.h:35)
		const int typeB = __ldg(&a.collidables[cB].shapeType);
		if (typeB != B3B200_SHAPE_CONVEX_HULL && typeB != B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS && typeB != B3B200_SHAPE_SPHERE) continue;
		if (a.pose[2 * bodyA].w == 0.f && a.pose[2 * bodyB].w == 0.f) continue;
		const int4 mesh = __ldg(&a.meshInfos[__ldg(&a.collidables[cA].bvhIndex)]);
		const b3b200_convex_polyhedron* cv = &a.convex[__ldg(&a.collidables[cA].shapeIndex)];
		const int faceOffset = __ldg(&cv->faceOffset), vertexOffset = __ldg(&cv->vertexOffset);
		const float4 qmn = *reinterpret_cast<const float4*>(a.aabbs[bodyB].min), qmx = *reinterpret_cast<const float4*>(a.aabbs[bodyB].max);
		const int firstChild = typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS ? __ldg(&a.collidables[cB].shapeIndex) : -1;
		const int numChildren = typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS ? __ldg(&a.collidables[cB].numChildShapes) : 1;

		// exact triangle-AABB test + emission of one triangle slot (b3FindConcaveSatAxis.h:606-610)
		auto testAndEmit = [&](int triSlot, bool valid) {
			bool hit = false;
			int tri = 0;
			if (valid)
			{
				tri = __ldg(&a.meshTris[mesh.z + triSlot]);
				const int idxOff = __ldg(&a.faces[faceOffset + tri].indexOffset);
				float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
				for (int i = 0; i < 3; i++)
				{
					const float4 v = __ldg(&a.vertices[vertexOffset + __ldg(&a.indices[idxOff + i])]);
					mn[0] = fminf(mn[0], v.x);
					mn[1] = fminf(mn[1], v.y);
					mn[2] = fminf(mn[2], v.z);
					mx[0] = fmaxf(mx[0], v.x);
					mx[1] = fmaxf(mx[1], v.y);
					mx[2] = fmaxf(mx[2], v.z);
				}
				hit = !(mn[0] > qmx.x || mx[0] < qmn.x || mn[2] > qmx.z || mx[2] < qmn.z || mn[1] > qmx.y || mx[1] < qmn.y);
			}
			const unsigned int m = __ballot_sync(FULL, hit);
			if (m)
			{
				unsigned int slot = 0;
				if (lane == 0) slot = atomicAdd(&a.ctr[CTR_CONCAVE_PAIRS], (unsigned int)(__popc(m) * numChildren));
				slot = __shfl_sync(FULL, slot, 0) + (unsigned int)(__popc(m & lt) * numChildren);
				if (hit)
					for (int c = 0; c < numChildren; c++)
						if (slot + c < (unsigned int)a.maxItems) items[slot + c] = make_int4(p, tri, firstChild < 0 ? -1 : firstChild + c, 0);
			}
		};

		int* cur = queueAll[warp][0];
		int* nxt = queueAll[warp][1];
		int* leaves = leafAll[warp];
		int n = 1, numLeaves = 0;
		if (lane == 0) cur[0] = 0;
		__syncwarp();
		while (n > 0)
		{
			int nn = 0;
			for (int base = 0; base < n; base += 32)
			{
				const int i = base + lane;
				int kind = 0;  // 0 = nothing, 1 = internal (push 2 children), 2 = leaf
				int first = 0, count = 0;
				if (i < n)
				{
					const int node = cur[i];
					const float4 n0 = __ldg(&a.meshNodes[2 * (mesh.x + node)]), n1 = __ldg(&a.meshNodes[2 * (mesh.x + node) + 1]);
					if (!(n0.x > qmx.x || n1.x < qmn.x || n0.y > qmx.y || n1.y < qmn.y || n0.z > qmx.z || n1.z < qmn.z))
					{
						count = __float_as_int(n1.w);
						first = __float_as_int(n0.w);
						kind = count == 0 ? 1 : 2;
					}
				}
				const unsigned int mi = __ballot_sync(FULL, kind == 1), ml = __ballot_sync(FULL, kind == 2);
				const int pi = nn + 2 * __popc(mi & lt), pl = numLeaves + __popc(ml & lt);
				bool spilled = false;
				if (kind == 1)
				{
					if (pi + 2 <= CULL_QUEUE)
					{
						nxt[pi] = first;
						nxt[pi + 1] = first + 1;
					}
					else
						spilled = true;
				}
				else if (kind == 2)
				{
					if (pl < CULL_QUEUE)
						leaves[pl] = first | (count << 28);
					else
						spilled = true;
				}
				nn = min(nn + 2 * __popc(mi), CULL_QUEUE);
				numLeaves = min(numLeaves + __popc(ml), CULL_QUEUE);
				// a full queue (a very large box over a fine mesh): the lanes that could not push walk their subtree alone
				if (__any_sync(FULL, spilled))
				{
					if (spilled)
					{
						auto emitSerial = [&](int triSlot) {
							const int tri = __ldg(&a.meshTris[mesh.z + triSlot]);
							const int idxOff = __ldg(&a.faces[faceOffset + tri].indexOffset);
							float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
							for (int v3 = 0; v3 < 3; v3++)
							{
								const float4 v = __ldg(&a.vertices[vertexOffset + __ldg(&a.indices[idxOff + v3])]);
								mn[0] = fminf(mn[0], v.x);
								mn[1] = fminf(mn[1], v.y);
								mn[2] = fminf(mn[2], v.z);
								mx[0] = fmaxf(mx[0], v.x);
								mx[1] = fmaxf(mx[1], v.y);
								mx[2] = fmaxf(mx[2], v.z);
							}
							if (mn[0] > qmx.x || mx[0] < qmn.x || mn[2] > qmx.z || mx[2] < qmn.z || mn[1] > qmx.y || mx[1] < qmn.y) return;
							const unsigned int slot = atomicAdd(&a.ctr[CTR_CONCAVE_PAIRS], (unsigned int)numChildren);
							for (int c = 0; c < numChildren; c++)
								if (slot + c < (unsigned int)a.maxItems) items[slot + c] = make_int4(p, tri, firstChild < 0 ? -1 : firstChild + c, 0);
						};
						if (kind == 1)
						{
							walkSubtree(a, mesh, first, qmn, qmx, emitSerial);
							walkSubtree(a, mesh, first + 1, qmn, qmx, emitSerial);
						}
						else
							for (int t = 0; t < count; t++) emitSerial(first + t);
					}
					__syncwarp();
				}
			}
			__syncwarp();
			int* t = cur;
			cur = nxt;
			nxt = t;
			n = nn;
			// the leaf list is drained whenever it may not take another level's worth
			if (numLeaves > CULL_QUEUE / 2 || n == 0)
			{
				for (int base = 0; base < numLeaves * 4; base += 32)
				{
					const int idx = base + lane;
					bool valid = false;
					int triSlot = 0;
					if (idx < numLeaves * 4)
					{
						const int e = leaves[idx >> 2];
						const int cnt = (int)((unsigned int)e >> 28), f0 = e & 0x0fffffff;
						valid = (idx & 3) < cnt;
						triSlot = f0 + (idx & 3);
					}
					testAndEmit(triSlot, valid);
				}
				numLeaves = 0;
				__syncwarp();
			}
		}
	}
}

// computeContactSphereTriangle (kernels/primitiveContacts.cl:1162-1300, called by findConcaveSphereContactsKernel
// :1305-1373; the reference has no host twin of it).  The sphere is "A" of the contact, the mesh "B".
B3_D void sphereTriangleThread(const CcArgs& a, const int4& it, int meshBody, int sphereBody, int cMesh, int cSphere)
{
	const float radius = __ldg(&a.collidables[cSphere].radius);
	float4 pos = a.pose[2 * meshBody];
	const float4 quat = a.pose[2 * meshBody + 1];
	const float invMassMesh = pos.w;
	pos.w = 0.f;
	float4 spherePos2 = a.pose[2 * sphereBody];
	const float invMassSphere = spherePos2.w;
	spherePos2.w = 0.f;
	const b3b200_convex_polyhedron* cv = &a.convex[__ldg(&a.collidables[cMesh].shapeIndex)];
	const int idxOff = __ldg(&a.faces[__ldg(&cv->faceOffset) + it.y].indexOffset), vOff = __ldg(&cv->vertexOffset);
	const float4 v0 = __ldg(&a.vertices[vOff + __ldg(&a.indices[idxOff])]);
	const float4 v1 = __ldg(&a.vertices[vOff + __ldg(&a.indices[idxOff + 1])]);
	const float4 v2 = __ldg(&a.vertices[vOff + __ldg(&a.indices[idxOff + 2])]);
	const float4 invOrn = quatInverse(quat), invPos = quatRotate(invOrn, neg3(pos));
	const float4 sphereCenter = add3(quatRotate(invOrn, spherePos2), invPos);
	float4 normal = normalized3(cross3(sub3(v1, v0), sub3(v2, v0)));
	float distanceFromPlane = dot3(sub3(sphereCenter, v0), normal);
	if (distanceFromPlane < 0.f)
	{
		distanceFromPlane *= -1.f;
		normal = scale3(normal, -1.f);
	}
	bool hasContact = false;
	float4 contactPoint = mk4(0, 0, 0);
	if (distanceFromPlane < radius)
	{
		const float4 e1 = sub3(v1, v0), e2 = sub3(v2, v1), e3 = sub3(v0, v2);
		const float r1 = dot3(cross3(e1, normal), sub3(sphereCenter, v0));
		const float r2 = dot3(cross3(e2, normal), sub3(sphereCenter, v1));
		const float r3 = dot3(cross3(e3, normal), sub3(sphereCenter, v2));
		if ((r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0))
		{
			hasContact = true;
			contactPoint = sub3(sphereCenter, scale3(normal, distanceFromPlane));
		}
		else
		{
#pragma unroll
			for (int i = 0; i < 3; i++)
			{
				const float4 from = i == 0 ? v0 : (i == 1 ? v1 : v2), to = i == 0 ? v1 : (i == 1 ? v2 : v0);
				float4 diff = sub3(sphereCenter, from);
				const float4 v = sub3(to, from);
				float t = dot3(v, diff);
				if (t > 0)
				{
					const float dotVV = dot3(v, v);
					if (t < dotVV)
					{
						t /= dotVV;
						diff = sub3(diff, scale3(v, t));
					}
					else
					{
						t = 1;
						diff = sub3(diff, v);
					}
				}
				else
					t = 0;
				const float4 nearest = add3(from, scale3(v, t));
				if (dot3(diff, diff) < radius * radius)
				{
					hasContact = true;
					contactPoint = nearest;
				}
			}
		}
	}
	if (!hasContact) return;
	const float4 contactToCenter = sub3(sphereCenter, contactPoint);
	const float minDist = sqrtf(dot3(contactToCenter, contactToCenter));
	if (!(minDist > FLT_EPSILON)) return;
	const float4 hitNormal = normalized3(contactToCenter);
	const float4 normalOnSurfaceB1 = quatRotate(quat, neg3(hitNormal));
	float4 pOnB1 = add3(quatRotate(quat, contactPoint), pos);
	const float actualDepth = minDist - radius;
	if (!(actualDepth <= 0.f) || !(dot3(normalOnSurfaceB1, normalOnSurfaceB1) > FLT_EPSILON)) return;
	const unsigned int slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
	if (slot >= (unsigned int)a.maxContacts) return;
	b3b200_contact4* c = &a.contacts[slot];
	float4* cw = reinterpret_cast<float4*>(c);
	pOnB1.w = actualDepth;
	cw[0] = pOnB1;
	cw[1] = cw[2] = cw[3] = mk4(0, 0, 0, 0);
	cw[4] = mk4(-normalOnSurfaceB1.x, -normalOnSurfaceB1.y, -normalOnSurfaceB1.z, 1.f);
	int4 t5;
	t5.x = (int)(0u | (45874u << 16));
	t5.y = it.y;
	t5.z = invMassSphere == 0.f ? -sphereBody : sphereBody;
	t5.w = invMassMesh == 0.f ? -meshBody : meshBody;
	reinterpret_cast<int4*>(c)[5] = t5;
	reinterpret_cast<int4*>(c)[6] = make_int4(-1, it.y, 0, 0);  // m_childIndexB = faceIndex (:1290)
}

// stage 1b: exact quick reject, one THREAD per (pair, triangle, child) item.  Tests three members of the reference's
// own axis list with its own arithmetic -- the triangle normal, the edge plane of the triangle that faces B most, and
// the face of B that faces the triangle most -- so an item dropped here is one the full SAT would drop too.
__global__ void __launch_bounds__(256, 4) concaveQuickKernel(CcArgs a, const int4* __restrict__ rawItems, int4* __restrict__ items)
{
	int numRaw = (int)a.ctr[CTR_CONCAVE_PAIRS];
	if (numRaw > a.maxItems) numRaw = a.maxItems;
	const int lane = threadIdx.x & 31;
	for (int base = blockIdx.x * blockDim.x; base < numRaw; base += gridDim.x * blockDim.x)
	{
		const int r = base + threadIdx.x;
		bool keep = false, smallB = false;
		int4 it = make_int4(0, 0, 0, 0);
		if (r < numRaw)
		{
			it = rawItems[r];
			const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
			const int cA = a.coll[bodyA], cB = a.coll[bodyB];
			if (__ldg(&a.collidables[cB].shapeType) == B3B200_SHAPE_SPHERE)
			{
				sphereTriangleThread(a, it, bodyA, bodyB, cA, cB);  // one-point contact, finished here
				it = make_int4(0, 0, 0, 0);
			}
			else
			{
			float4 posA = a.pose[2 * bodyA], posB = a.pose[2 * bodyB];
			const float4 ornA = a.pose[2 * bodyA + 1];
			float4 ornB = a.pose[2 * bodyB + 1];
			posA.w = 0.f;
			posB.w = 0.f;
			const b3b200_convex_polyhedron* cvA = &a.convex[__ldg(&a.collidables[cA].shapeIndex)];
			const b3b200_face* faceA = &a.faces[__ldg(&cvA->faceOffset) + it.y];
			const float4 plane = __ldg(reinterpret_cast<const float4*>(&faceA->plane));
			const int idxOffA = __ldg(&faceA->indexOffset), vOffA = __ldg(&cvA->vertexOffset);
			float4 vA[3];
#pragma unroll
			for (int i = 0; i < 3; i++) vA[i] = __ldg(&a.vertices[vOffA + __ldg(&a.indices[idxOffA + i])]);
			const float4 normal = mk4(plane.x, plane.y, plane.z);
			int shapeB;
			if (it.z >= 0)
			{
				const b3b200_child_shape* ch = &a.childShapes[it.z];
				const float4 np = transformPoint(__ldg(reinterpret_cast<const float4*>(&ch->childPosition)), posB, ornB);
				ornB = quatMul(ornB, __ldg(reinterpret_cast<const float4*>(&ch->childOrientation)));
				posB = np;
				shapeB = __ldg(&a.collidables[__ldg(&ch->shapeIndex)].shapeIndex);
			}
			else
				shapeB = __ldg(&a.collidables[cB].shapeIndex);
			const HullRef hB = loadHull(a.convex, shapeB);
			smallB = hB.numVertices <= CCT_VERTS && hB.numFaces <= CCT_FACES && hB.numUniqueEdges <= CCT_EDGES;
			float4 localCenter = add3(add3(vA[0], vA[1]), vA[2]);
			localCenter = scale3(localCenter, 1.f / 3.f);
			const float4 deltaC2 = sub3(transformPoint(localCenter, posA, ornA), transformPoint(hB.localCenter, posB, ornB));
			keep = true;
#pragma unroll 1
			for (int which = 0; which < 3 && keep; which++)
			{
				float4 axis;
				if (which == 0)
					axis = quatRotate(ornA, normal);
				else if (which == 1)
				{
					// edge plane i: normalized(cross(normal, v[prev] - v[i])); pick the one pointing at B most
					const float4 toB = quatRotate(quatInverse(ornA), neg3(deltaC2));
					float best = -FLT_MAX;
					float4 bestN = mk4(0, 0, 0);
					int prev = 2;
#pragma unroll
					for (int i = 0; i < 3; i++)
					{
						const float4 en = normalized3(cross3(normal, sub3(vA[prev], vA[i])));
						const float d = dot3(en, toB);
						if (d > best)
						{
							best = d;
							bestN = en;
						}
						prev = i;
					}
					axis = quatRotate(ornA, bestN);
				}
				else
				{
					const float4 toA = quatRotate(quatInverse(ornB), deltaC2);
					int bf = 0;
					float best = -FLT_MAX;
					for (int f = 0; f < hB.numFaces; f++)
					{
						const float4 n = __ldg(reinterpret_cast<const float4*>(&a.faces[hB.faceOffset + f].plane));
						const float d = dot3(n, toA);
						if (d > best)
						{
							best = d;
							bf = f;
						}
					}
					axis = quatRotate(ornB, __ldg(reinterpret_cast<const float4*>(&a.faces[hB.faceOffset + bf].plane)));
				}
				if (dot3(deltaC2, axis) < 0) axis = mk4(axis.x * -1.f, axis.y * -1.f, axis.z * -1.f);
				float minT, maxT, minH, maxH;
				projectTri(vA, posA, ornA, axis, minT, maxT);
				projectHull(hB, posB, ornB, axis, a.vertices, minH, maxH);
				if (maxT < minH || maxH < minT) keep = false;
			}
			}
		}
		// survivors <= raw items <= capacity, so the two ends of the list cannot meet: items with a small hull B (thread-per-item
		// kernel) from the front, the others (warp-per-item kernel) from the back
		const unsigned int mf = __ballot_sync(FULL, keep && smallB), mb = __ballot_sync(FULL, keep && !smallB);
		if (mf | mb)
		{
			unsigned int sf = 0, sb = 0;
			if (lane == 0)
			{
				if (mf) sf = atomicAdd(&a.ctr[CTR_CONCAVE_SURVIVORS], (unsigned int)__popc(mf));
				if (mb) sb = atomicAdd(&a.ctr[CTR_CONCAVE_SURVIVORS_BACK], (unsigned int)__popc(mb));
			}
			const unsigned int lt = (1u << lane) - 1u;
			sf = __shfl_sync(FULL, sf, 0) + __popc(mf & lt);
			sb = __shfl_sync(FULL, sb, 0) + __popc(mb & lt);
			if (keep && smallB) items[sf] = it;
			if (keep && !smallB) items[(unsigned int)a.maxItems - 1u - sb] = it;
		}
	}
}

// clipFaceGlobal (shared/b3ClipFaces.h:21-64), serial
B3_D int clipFaceGlobalSerial(const float4* in, int numIn, const float4& n, float eq, float4* out)
{
	int numOut = 0;
	for (int ve = 0; ve < numIn; ve++)
	{
		const float4 first = in[ve == 0 ? numIn - 1 : ve - 1];
		const float4 end = in[ve];
		const float ds = dot3(n, first) + eq;
		const float de = dot3(n, end) + eq;
		if (ds < 0)
		{
			if (de < 0)
			{
				if (numOut < CCT_POLY) out[numOut++] = end;
			}
			else if (numOut < CCT_POLY)
				out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
		}
		else if (de < 0)
		{
			if (numOut < CCT_POLY) out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
			if (numOut < CCT_POLY) out[numOut++] = end;
		}
	}
	return numOut;
}

// Triangle x small hull, ONE THREAD per item: the same steps as concaveContactKernel (SAT over the 5-face triangle prism, the
// faces of B and the 3 x E_B edge axes in the reference's order with its first-strict-minimum rule, b3FindClippingFaces,
// clipFaceGlobal, b3ExtractManifoldSequentialGlobal, append), serially.  With at most 5 + 6 + 18 axes over 3 + 8 vertices a
// warp per item keeps about 8 lanes busy; here 32 items share a warp.
B3_D void concaveSmallThread(const CcArgs& a, const int4 it)
{
	const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
	const int cA = a.coll[bodyA], cB = a.coll[bodyB];
	float4 posA = a.pose[2 * bodyA], posB = a.pose[2 * bodyB];
	const float4 ornA = a.pose[2 * bodyA + 1];
	float4 ornB = a.pose[2 * bodyB + 1];
	const float invMassA = posA.w, invMassB = posB.w;
	posA.w = 0.f;
	posB.w = 0.f;
	const b3b200_convex_polyhedron* cvA = &a.convex[__ldg(&a.collidables[cA].shapeIndex)];
	const b3b200_face* faceA = &a.faces[__ldg(&cvA->faceOffset) + it.y];
	const float4 plane = __ldg(reinterpret_cast<const float4*>(&faceA->plane));
	const int idxOffA = __ldg(&faceA->indexOffset), vOffA = __ldg(&cvA->vertexOffset);
	float4 vA[3];
#pragma unroll
	for (int i = 0; i < 3; i++) vA[i] = __ldg(&a.vertices[vOffA + __ldg(&a.indices[idxOffA + i])]);
	const float4 normal = mk4(plane.x, plane.y, plane.z);
	float4 triN[5];
	triN[0] = normal;
	triN[1] = neg3(normal);
	{
		int prev = 2;
#pragma unroll
		for (int i = 0; i < 3; i++)
		{
			triN[2 + i] = normalized3(cross3(normal, sub3(vA[prev], vA[i])));
			prev = i;
		}
	}
	float4 localCenter = add3(add3(vA[0], vA[1]), vA[2]);
	localCenter = scale3(localCenter, 1.f / 3.f);
	int shapeB;
	if (it.z >= 0)
	{
		const b3b200_child_shape* ch = &a.childShapes[it.z];
		const float4 cp = __ldg(reinterpret_cast<const float4*>(&ch->childPosition));
		const float4 co = __ldg(reinterpret_cast<const float4*>(&ch->childOrientation));
		const float4 np = transformPoint(cp, posB, ornB);
		ornB = quatMul(ornB, co);
		posB = np;
		shapeB = __ldg(&a.collidables[__ldg(&ch->shapeIndex)].shapeIndex);
	}
	else
		shapeB = __ldg(&a.collidables[cB].shapeIndex);
	const HullRef hB = loadHull(a.convex, shapeB);
	const float4 c0 = transformPoint(localCenter, posA, ornA);
	const float4 c1 = transformPoint(hB.localCenter, posB, ornB);
	const float4 deltaC2 = sub3(c0, c1);

	// ---- SAT, axes in the reference's order; "d < dmin" keeps the first strict minimum
	const int nFB = hB.numFaces, nEB = hB.numUniqueEdges;
	const int total = 5 + nFB + 3 * nEB;
	float bestD = FLT_MAX;
	float4 sep = mk4(0, 0, 0);
	bool any = false;
#pragma unroll 1
	for (int k = 0; k < total; k++)
	{
		float4 axis;
		if (k < 5)
		{
			if (k == 1) continue;  // -normal: same oriented axis and depth as face 0, never a strict minimum
			axis = quatRotate(ornA, triN[k]);
		}
		else if (k < 5 + nFB)
		{
			const b3b200_face* f = &a.faces[hB.faceOffset + (k - 5)];
			if (__ldg(&f->pad1) != 0) continue;  // bitwise +-duplicate of an earlier face normal
			axis = quatRotate(ornB, __ldg(reinterpret_cast<const float4*>(&f->plane)));
		}
		else
		{
			const int e = k - 5 - nFB;
			const int e0 = e / nEB, e1 = e - e0 * nEB;
			const float4 edgeA = e0 == 0 ? sub3(vA[1], vA[0]) : (e0 == 1 ? sub3(vA[2], vA[1]) : sub3(vA[0], vA[2]));
			const float4 edge0World = quatRotate(ornA, edgeA);
			const float4 edge1World = quatRotate(ornB, __ldg(&a.uniqueEdges[hB.uniqueEdgesOffset + e1]));
			const float4 cr = cross3(edge0World, edge1World);
			if (almostZero(cr)) continue;
			axis = normalized3(cr);
		}
		if (dot3(deltaC2, axis) < 0) axis = mk4(axis.x * -1.f, axis.y * -1.f, axis.z * -1.f);
		float minT, maxT, minH, maxH;
		projectTri(vA, posA, ornA, axis, minT, maxT);
		projectHull(hB, posB, ornB, axis, a.vertices, minH, maxH);
		if (maxT < minH || maxH < minT) return;
		const float d0 = maxT - minH, d1 = maxH - minT;
		const float d = d0 < d1 ? d0 : d1;
		if (d < bestD)
		{
			bestD = d;
			sep = axis;
			any = true;
		}
	}
	if (!any) return;
	if (dot3(neg3(deltaC2), sep) > 0.0f) sep = neg3(sep);

	// ---- b3FindClippingFaces: incident face of B (most aligned), reference face of the triangle prism (least aligned)
	int closestFaceB = -1;
	{
		float dmax = -FLT_MAX;
		for (int f = 0; f < nFB; f++)
		{
			const float4 n = __ldg(reinterpret_cast<const float4*>(&a.faces[hB.faceOffset + f].plane));
			const float d = dot3(quatRotate(ornB, mk4(n.x, n.y, n.z)), sep);
			if (d > dmax)
			{
				dmax = d;
				closestFaceB = f;
			}
		}
	}
	if (closestFaceB < 0) return;
	float4 bufA[CCT_POLY], bufB[CCT_POLY];
	int numVertsIn;
	{
		const b3b200_face* polyB = &a.faces[hB.faceOffset + closestFaceB];
		const int idxOff = __ldg(&polyB->indexOffset);
		numVertsIn = __ldg(&polyB->numIndices);
		if (numVertsIn > CCT_POLY) numVertsIn = CCT_POLY;  // (a small hull's face has at most CCT_VERTS vertices)
		for (int e = 0; e < numVertsIn; e++) bufA[e] = transformPoint(__ldg(&a.vertices[hB.vertexOffset + __ldg(&a.indices[idxOff + e])]), posB, ornB);
	}
	int closestFaceA = 0;
	float4 worldNormalA = mk4(0, 0, 0);
	{
		float dmin = FLT_MAX;
#pragma unroll
		for (int f = 0; f < 5; f++)
		{
			const float4 n = quatRotate(ornA, triN[f]);
			const float d = dot3(n, sep);
			if (d < dmin)
			{
				dmin = d;
				closestFaceA = f;
				worldNormalA = n;
			}
		}
	}
	// vertices of that face: front (0,1,2), back (2,1,0), edge plane i: (i, previous vertex)
	const int numA = closestFaceA < 2 ? 3 : 2;
	int ia0, ia1, ia2 = 0;
	if (closestFaceA == 0)
		ia0 = 0, ia1 = 1, ia2 = 2;
	else if (closestFaceA == 1)
		ia0 = 2, ia1 = 1, ia2 = 0;
	else
		ia0 = closestFaceA - 2, ia1 = (closestFaceA - 2 + 2) % 3;
	float4 a1[3];
	a1[0] = transformPoint(ia0 == 0 ? vA[0] : (ia0 == 1 ? vA[1] : vA[2]), posA, ornA);
	a1[1] = transformPoint(ia1 == 0 ? vA[0] : (ia1 == 1 ? vA[1] : vA[2]), posA, ornA);
	a1[2] = transformPoint(ia2 == 0 ? vA[0] : (ia2 == 1 ? vA[1] : vA[2]), posA, ornA);

	// ---- clipFacesAndFindContactsKernel
	float4* pIn = bufA;
	float4* pOut = bufB;
	for (int e0 = 0; e0 < numA; e0++)
	{
		const float4 aw = e0 == 0 ? a1[0] : (e0 == 1 ? a1[1] : a1[2]);
		const int e1 = (e0 + 1) % numA;
		const float4 bw = e1 == 0 ? a1[0] : (e1 == 1 ? a1[1] : a1[2]);
		const float4 worldEdge0 = sub3(aw, bw);
		const float4 planeNormalWS = neg3(cross3(worldEdge0, worldNormalA));
		const float planeEqWS = -dot3(aw, planeNormalWS);
		const int numOut = clipFaceGlobalSerial(pIn, numVertsIn, planeNormalWS, planeEqWS, pOut);
		float4* t = pOut;
		pOut = pIn;
		pIn = t;
		numVertsIn = numOut;
	}
	int numContactsOut = 0;
	{
		const float planeEqWS = -dot3(worldNormalA, a1[0]);
		for (int i = 0; i < numVertsIn; i++)
		{
			float4 pt = pIn[i];
			float depth = dot3(worldNormalA, pt) + planeEqWS;
			if (depth <= -1e30f) depth = -1e30f;
			if (depth <= 0.02f)
			{
				pt.w = depth;
				pOut[numContactsOut++] = pt;
			}
		}
	}
	if (numContactsOut <= 0) return;
	const float4* pts = pOut;

	// ---- b3ExtractManifoldSequentialGlobal with nearNormal = -sep (b3NewContactReduction.h:10-91, 127)
	int idx[4] = {0, 1, 2, 3};
	int numPoints = numContactsOut;
	if (numContactsOut > 4)
	{
		const float4 nearNormal = neg3(sep);
		const int nPoints = numContactsOut;
		float4 center = mk4(0, 0, 0);
		for (int i = 0; i < nPoints; i++)
		{
			const float4 p = pts[i];
			center.x += p.x;
			center.y += p.y;
			center.z += p.z;
		}
		{
			const float sc = 1.0f / (float)nPoints;
			center.x *= sc;
			center.y *= sc;
			center.z *= sc;
		}
		const float4 aVector = sub3(pts[0], center);
		float4 u = cross3(nearNormal, aVector);
		float4 v = cross3(nearNormal, u);
		u = normalized3(u);
		v = normalized3(v);
		const float4 nu = neg3(u), nv = neg3(v);
		float minW = FLT_MAX;
		int minIndex = -1;
		float m0 = FLT_MIN, m1 = FLT_MIN, m2 = FLT_MIN, m3 = FLT_MIN;
		for (int ie = 0; ie < nPoints; ie++)
		{
			const float4 p = pts[ie];
			if (p.w < minW)
			{
				minW = p.w;
				minIndex = ie;
			}
			const float4 r = sub3(p, center);
			float f = dot3(u, r);
			if (f < m0)
			{
				m0 = f;
				idx[0] = ie;
			}
			f = dot3(nu, r);
			if (f < m1)
			{
				m1 = f;
				idx[1] = ie;
			}
			f = dot3(v, r);
			if (f < m2)
			{
				m2 = f;
				idx[2] = ie;
			}
			f = dot3(nv, r);
			if (f < m3)
			{
				m3 = f;
				idx[3] = ie;
			}
		}
		if (idx[0] != minIndex && idx[1] != minIndex && idx[2] != minIndex && idx[3] != minIndex) idx[0] = minIndex;
		numPoints = 4;
	}

	// ---- append (b3NewContactReduction.h:129-166)
	const unsigned int slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
	if (slot >= (unsigned int)a.maxContacts) return;  // clamped afterwards, OVF_CONTACTS raised
	b3b200_contact4* c = &a.contacts[slot];
	float4* cw = reinterpret_cast<float4*>(c);
	for (int i = 0; i < 4; i++) cw[i] = i < numPoints ? pts[idx[i]] : mk4(0, 0, 0, 0);
	cw[4] = mk4(sep.x, sep.y, sep.z, (float)numPoints);
	int4 t;
	t.x = (int)(0u | (45874u << 16));  // restitution 0, friction (0.7f * 0xffff)
	t.y = it.y;                         // the reference stores its concave-pair index; the triangle index is the stable equivalent
	t.z = invMassA == 0.f ? -bodyA : bodyA;
	t.w = invMassB == 0.f ? -bodyB : bodyB;
	reinterpret_cast<int4*>(c)[5] = t;
	reinterpret_cast<int4*>(c)[6] = make_int4(-1, -1, 0, 0);  // child indices are not recorded on this path (:143-144)
}

__global__ void __launch_bounds__(128, 5) concaveSmallKernel(CcArgs a, const int4* __restrict__ items)
{
	int numItems = (int)a.ctr[CTR_CONCAVE_SURVIVORS];
	if (numItems > a.maxItems) numItems = a.maxItems;
	for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < numItems; s += gridDim.x * blockDim.x) concaveSmallThread(a, items[s]);
}

__global__ void __launch_bounds__(CC_THREADS) concaveContactKernel(CcArgs a, const int4* __restrict__ items)
{
	__shared__ float4 bufAll[CC_WARPS][2][CC_MAX_POLY];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	float4* bufA = bufAll[warp][0];
	float4* bufB = bufAll[warp][1];
	int numItems = (int)a.ctr[CTR_CONCAVE_SURVIVORS_BACK];  // the items with a larger hull B sit at the back end of the list
	if (numItems > a.maxItems) numItems = a.maxItems;
	// static striding: a dynamic cursor (as in satKernel) measured 2.7x slower here
	const int warpsTotal = gridDim.x * CC_WARPS;
	{
	for (int s = blockIdx.x * CC_WARPS + warp; s < numItems; s += warpsTotal)
	{
		__syncwarp();
		const int4 it = items[a.maxItems - 1 - s];
		const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
		const int cA = a.coll[bodyA], cB = a.coll[bodyB];
		float4 posA = a.pose[2 * bodyA], posB = a.pose[2 * bodyB];
		const float4 ornA = a.pose[2 * bodyA + 1];
		float4 ornB = a.pose[2 * bodyB + 1];
		const float invMassA = posA.w, invMassB = posB.w;
		posA.w = 0.f;
		posB.w = 0.f;

		// ---- the triangle as a 5-face convex: front, back, three edge planes (b3FindConcaveSatAxis.h:588-690)
		const b3b200_convex_polyhedron* cvA = &a.convex[__ldg(&a.collidables[cA].shapeIndex)];
		const b3b200_face* faceA = &a.faces[__ldg(&cvA->faceOffset) + it.y];
		const float4 plane = __ldg(reinterpret_cast<const float4*>(&faceA->plane));
		const int idxOffA = __ldg(&faceA->indexOffset), vOffA = __ldg(&cvA->vertexOffset);
		float4 vA[3];
#pragma unroll
		for (int i = 0; i < 3; i++) vA[i] = __ldg(&a.vertices[vOffA + __ldg(&a.indices[idxOffA + i])]);
		const float4 normal = mk4(plane.x, plane.y, plane.z);
		float4 triN[5];
		triN[0] = normal;
		triN[1] = neg3(normal);
		{
			int prev = 2;
#pragma unroll
			for (int i = 0; i < 3; i++)
			{
				triN[2 + i] = normalized3(cross3(normal, sub3(vA[prev], vA[i])));
				prev = i;
			}
		}
		float4 localCenter = add3(add3(vA[0], vA[1]), vA[2]);
		localCenter = scale3(localCenter, 1.f / 3.f);

		// ---- compound child of B (:694-707)
		int shapeB;
		if (it.z >= 0)
		{
			const b3b200_child_shape* ch = &a.childShapes[it.z];
			const float4 cp = __ldg(reinterpret_cast<const float4*>(&ch->childPosition));
			const float4 co = __ldg(reinterpret_cast<const float4*>(&ch->childOrientation));
			const float4 np = transformPoint(cp, posB, ornB);
			ornB = quatMul(ornB, co);
			posB = np;
			shapeB = __ldg(&a.collidables[__ldg(&ch->shapeIndex)].shapeIndex);
		}
		else
			shapeB = __ldg(&a.collidables[cB].shapeIndex);
		const HullRef hB = loadHull(a.convex, shapeB);

		const float4 c0 = transformPoint(localCenter, posA, ornA);
		const float4 c1 = transformPoint(hB.localCenter, posB, ornB);
		const float4 deltaC2 = sub3(c0, c1);

		// ---- SAT: axes in the reference's order -- 5 triangle faces, faces of B, (triangle edge) x (edge of B)
		const int nFB = hB.numFaces, nEB = hB.numUniqueEdges;
		const int total = 5 + nFB + 3 * nEB;
		float bestD = FLT_MAX;
		int bestK = -1;
		float4 bestAxis = mk4(0, 0, 0);
		bool separated = false;
		// Exact-safe skip for the later rounds of hulls with many axes (same argument as satWarp, narrowphase.cu): along
		// an oriented unit axis n (deltaC2 . n >= 0, the triangle lies "ahead" of B) the overlap is
		//     min(maxT - minH, maxH - minT)  >=  min(maxT - c1.n + rB,  max_k(b_k . n) + c1.n - minT)
		// with the triangle's projection exact (3 vertices), rB the inscribed radius of B and b_k the two vertices of B
		// (relative to its centre c1) that reach furthest towards the triangle.  An axis whose bound exceeds the best
		// depth of the previous rounds by more than supEps can neither separate nor become the strict minimum.
		const bool tight = total > 32;
		const float supEps = 1e-3f + 1e-6f * (fabsf(posA.x) + fabsf(posA.y) + fabsf(posA.z) + fabsf(posB.x) + fabsf(posB.y) + fabsf(posB.z));
		float4 sup0 = mk4(0, 0, 0), sup1 = mk4(0, 0, 0);
		if (tight)
		{
			const float4 dirL = quatRotate(quatInverse(ornB), deltaC2);
			const int nv = hB.numVertices < 64 ? hB.numVertices : 64;
			int taken = -1;
#pragma unroll 1
			for (int kk = 0; kk < 2; kk++)
			{
				float best = -FLT_MAX;
				int bi = -1;
				for (int i = lane; i < nv; i += 32)
				{
					if (i == taken) continue;
					const float sc = dot3(__ldg(&a.vertices[hB.vertexOffset + i]), dirL);
					if (sc > best)
					{
						best = sc;
						bi = i;
					}
				}
				warpArgMax(best, bi);
				if (bi < 0) bi = 0;
				const float4 sv = quatRotate(ornB, sub3(__ldg(&a.vertices[hB.vertexOffset + bi]), hB.localCenter));
				if (kk == 0)
					sup0 = sv;
				else
					sup1 = sv;
				taken = bi;
			}
		}
		float curMin = FLT_MAX;  // warp-wide best depth of the finished rounds
		for (int base = 0; base < total; base += 32)
		{
			const int k = base + lane;
			float d = FLT_MAX;
			bool sepHere = false;
			do
			{
				if (k >= total) break;
				float4 axis;
				if (k < 5)
				{
					if (k == 1) break;  // -normal: the oriented axis and its depth are identical to face 0's, never a strict minimum
					axis = quatRotate(ornA, triN[k]);
				}
				else if (k < 5 + nFB)
				{
					const b3b200_face* f = &a.faces[hB.faceOffset + (k - 5)];
					if (__ldg(&f->pad1) != 0) break;  // bitwise +-duplicate of an earlier face normal (flag set at registration)
					axis = quatRotate(ornB, __ldg(reinterpret_cast<const float4*>(&f->plane)));
				}
				else
				{
					const int e = k - 5 - nFB;
					const int e0 = e / nEB, e1 = e - e0 * nEB;
					const float4 edgeA = e0 == 0 ? sub3(vA[1], vA[0]) : (e0 == 1 ? sub3(vA[2], vA[1]) : sub3(vA[0], vA[2]));
					const float4 edge0World = quatRotate(ornA, edgeA);
					const float4 edge1World = quatRotate(ornB, __ldg(&a.uniqueEdges[hB.uniqueEdgesOffset + e1]));
					const float4 cr = cross3(edge0World, edge1World);
					if (almostZero(cr)) break;
					axis = normalized3(cr);
				}
				if (dot3(deltaC2, axis) < 0) axis = mk4(axis.x * -1.f, axis.y * -1.f, axis.z * -1.f);
				float minT, maxT, minH, maxH;
				projectTri(vA, posA, ornA, axis, minT, maxT);
				if (tight && curMin < FLT_MAX)
				{
					const float c1n = dot3(c1, axis);
					const float l1 = fmaxf(dot3(sup0, axis), dot3(sup1, axis)) + c1n - minT;
					const float l2 = maxT - c1n + hB.radius * (1.0f - 2e-4f);
					if (fminf(l1, l2) > curMin + supEps) break;
				}
				projectHull(hB, posB, ornB, axis, a.vertices, minH, maxH);
				if (maxT < minH || maxH < minT)
				{
					sepHere = true;
					break;
				}
				const float d0 = maxT - minH, d1 = maxH - minT;
				d = d0 < d1 ? d0 : d1;
				if (d < bestD)
				{
					bestD = d;
					bestK = k;
					bestAxis = axis;
				}
			} while (false);
			if (__any_sync(FULL, sepHere))
			{
				separated = true;
				break;
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) d = fminf(d, __shfl_xor_sync(FULL, d, o));
			curMin = fminf(curMin, d);
		}
		if (separated) continue;
		const int myK = bestK;
		warpArgMin(bestD, bestK);
		if (bestK < 0) continue;
		const int src = __ffs(__ballot_sync(FULL, myK == bestK)) - 1;
		float4 sep = mk4(__shfl_sync(FULL, bestAxis.x, src), __shfl_sync(FULL, bestAxis.y, src), __shfl_sync(FULL, bestAxis.z, src));
		if (dot3(neg3(deltaC2), sep) > 0.0f) sep = neg3(sep);

		// ---- b3FindClippingFaces: incident face of B (most aligned), reference face of the triangle prism (least aligned)
		int closestFaceB = -1;
		{
			float dmax = -FLT_MAX;
			for (int f = lane; f < nFB; f += 32)
			{
				const float4 n = __ldg(reinterpret_cast<const float4*>(&a.faces[hB.faceOffset + f].plane));
				const float d = dot3(quatRotate(ornB, mk4(n.x, n.y, n.z)), sep);
				if (d > dmax)
				{
					dmax = d;
					closestFaceB = f;
				}
			}
			warpArgMax(dmax, closestFaceB);
		}
		if (closestFaceB < 0) continue;
		int numVertsIn;
		{
			const b3b200_face* polyB = &a.faces[hB.faceOffset + closestFaceB];
			const int idxOff = __ldg(&polyB->indexOffset);
			numVertsIn = __ldg(&polyB->numIndices);
			if (numVertsIn > CC_MAX_POLY) numVertsIn = CC_MAX_POLY;
			for (int e = lane; e < numVertsIn; e += 32) bufA[e] = transformPoint(__ldg(&a.vertices[hB.vertexOffset + __ldg(&a.indices[idxOff + e])]), posB, ornB);
			__syncwarp();
		}
		int closestFaceA = 0;
		float4 worldNormalA = mk4(0, 0, 0);
		{
			float dmin = FLT_MAX;
#pragma unroll
			for (int f = 0; f < 5; f++)
			{
				const float4 n = quatRotate(ornA, triN[f]);
				const float d = dot3(n, sep);
				if (d < dmin)
				{
					dmin = d;
					closestFaceA = f;
					worldNormalA = n;
				}
			}
		}
		// vertices of that face: front (0,1,2), back (2,1,0), edge plane i: (i, previous vertex)
		const int numA = closestFaceA < 2 ? 3 : 2;
		int ia0, ia1, ia2 = 0;
		if (closestFaceA == 0)
			ia0 = 0, ia1 = 1, ia2 = 2;
		else if (closestFaceA == 1)
			ia0 = 2, ia1 = 1, ia2 = 0;
		else
			ia0 = closestFaceA - 2, ia1 = (closestFaceA - 2 + 2) % 3;
		float4 a1[3];
		a1[0] = transformPoint(ia0 == 0 ? vA[0] : (ia0 == 1 ? vA[1] : vA[2]), posA, ornA);
		a1[1] = transformPoint(ia1 == 0 ? vA[0] : (ia1 == 1 ? vA[1] : vA[2]), posA, ornA);
		a1[2] = transformPoint(ia2 == 0 ? vA[0] : (ia2 == 1 ? vA[1] : vA[2]), posA, ornA);

		// ---- clipFacesAndFindContactsKernel
		float4* pIn = bufA;
		float4* pOut = bufB;
		for (int e0 = 0; e0 < numA; e0++)
		{
			const float4 aw = e0 == 0 ? a1[0] : (e0 == 1 ? a1[1] : a1[2]);
			const int e1 = (e0 + 1) % numA;
			const float4 bw = e1 == 0 ? a1[0] : (e1 == 1 ? a1[1] : a1[2]);
			const float4 worldEdge0 = sub3(aw, bw);
			const float4 planeNormalWS = neg3(cross3(worldEdge0, worldNormalA));
			const float planeEqWS = -dot3(aw, planeNormalWS);
			const int numOut = clipFaceGlobalWarp(pIn, numVertsIn, planeNormalWS, planeEqWS, pOut, lane);
			float4* t = pOut;
			pOut = pIn;
			pIn = t;
			numVertsIn = numOut;
		}
		int numContactsOut = 0;
		{
			const float planeEqWS = -dot3(worldNormalA, a1[0]);
			for (int base = 0; base < numVertsIn; base += 32)
			{
				const int i = base + lane;
				bool keep = false;
				float4 pt = mk4(0, 0, 0);
				if (i < numVertsIn)
				{
					pt = pIn[i];
					float depth = dot3(worldNormalA, pt) + planeEqWS;
					if (depth <= -1e30f) depth = -1e30f;
					if (depth <= 0.02f)
					{
						keep = true;
						pt.w = depth;
					}
				}
				const unsigned int m = __ballot_sync(FULL, keep);
				if (keep) pOut[numContactsOut + __popc(m & ((1u << lane) - 1u))] = pt;
				numContactsOut += __popc(m);
			}
			__syncwarp();
		}
		if (numContactsOut <= 0) continue;
		const float4* pts = pOut;

		// ---- b3ExtractManifoldSequentialGlobal with nearNormal = -sep (b3NewContactReduction.h:10-91, 127)
		int idx0 = 0, idx1 = 1, idx2 = 2, idx3 = 3;
		int numPoints = numContactsOut;
		if (numContactsOut > 4)
		{
			const float4 nearNormal = neg3(sep);
			const int nPoints = numContactsOut > 64 ? 64 : numContactsOut;
			float4 center = mk4(0, 0, 0);
			for (int i = 0; i < nPoints; i++)
			{
				const float4 p = pts[i];
				center.x += p.x;
				center.y += p.y;
				center.z += p.z;
			}
			{
				const float sc = 1.0f / (float)nPoints;
				center.x *= sc;
				center.y *= sc;
				center.z *= sc;
			}
			const float4 aVector = sub3(pts[0], center);
			float4 u = cross3(nearNormal, aVector);
			float4 v = cross3(nearNormal, u);
			u = normalized3(u);
			v = normalized3(v);
			const float4 nu = neg3(u), nv = neg3(v);
			float minW = FLT_MAX;
			int minIndex = -1;
			float m0 = FLT_MIN, m1 = FLT_MIN, m2 = FLT_MIN, m3 = FLT_MIN;
			int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
			for (int ie = lane; ie < nPoints; ie += 32)
			{
				const float4 p = pts[ie];
				if (p.w < minW)
				{
					minW = p.w;
					minIndex = ie;
				}
				const float4 r = sub3(p, center);
				float f = dot3(u, r);
				if (f < m0)
				{
					m0 = f;
					i0 = ie;
				}
				f = dot3(nu, r);
				if (f < m1)
				{
					m1 = f;
					i1 = ie;
				}
				f = dot3(v, r);
				if (f < m2)
				{
					m2 = f;
					i2 = ie;
				}
				f = dot3(nv, r);
				if (f < m3)
				{
					m3 = f;
					i3 = ie;
				}
			}
			warpArgMin(minW, minIndex);
			warpArgMin(m0, i0);
			warpArgMin(m1, i1);
			warpArgMin(m2, i2);
			warpArgMin(m3, i3);
			if (i0 >= 0) idx0 = i0;
			if (i1 >= 0) idx1 = i1;
			if (i2 >= 0) idx2 = i2;
			if (i3 >= 0) idx3 = i3;
			if (idx0 != minIndex && idx1 != minIndex && idx2 != minIndex && idx3 != minIndex) idx0 = minIndex;
			numPoints = 4;
		}

		// ---- append (b3NewContactReduction.h:129-166)
		unsigned int slot = 0;
		if (lane == 0) slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
		slot = __shfl_sync(FULL, slot, 0);
		if (slot >= (unsigned int)a.maxContacts) continue;  // clamped afterwards, OVF_CONTACTS raised
		b3b200_contact4* c = &a.contacts[slot];
		float4* cw = reinterpret_cast<float4*>(c);
		if (lane < 4)
		{
			const int id = lane == 0 ? idx0 : (lane == 1 ? idx1 : (lane == 2 ? idx2 : idx3));
			cw[lane] = lane < numPoints ? pts[id] : mk4(0, 0, 0, 0);
		}
		else if (lane == 4)
			cw[4] = mk4(sep.x, sep.y, sep.z, (float)numPoints);
		else if (lane == 5)
		{
			int4 t;
			t.x = (int)(0u | (45874u << 16));  // restitution 0, friction (0.7f * 0xffff)
			t.y = it.y;                         // the reference stores its concave-pair index; the triangle index is the stable equivalent
			t.z = invMassA == 0.f ? -bodyA : bodyA;
			t.w = invMassB == 0.f ? -bodyB : bodyB;
			reinterpret_cast<int4*>(c)[5] = t;
		}
		else if (lane == 6)
			reinterpret_cast<int4*>(c)[6] = make_int4(-1, -1, 0, 0);  // child indices are not recorded on this path (:143-144)
	}
	}
}

__global__ void clampConcaveKernel(unsigned int* ctr, int maxItems)
{
	if (ctr[CTR_CONCAVE_PAIRS] > (unsigned int)maxItems)
	{
		ctr[CTR_CONCAVE_PAIRS] = (unsigned int)maxItems;
		ctr[CTR_OVERFLOW] |= OVF_CONCAVE;
	}
}
}  // namespace

int launchConcave(World* w, cudaStream_t s)
{
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_CONCAVE_PAIRS], 0, sizeof(unsigned int), s));
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_CONCAVE_SURVIVORS], 0, sizeof(unsigned int), s));
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_CONCAVE_SURVIVORS_BACK], 0, sizeof(unsigned int), s));
	CcArgs a;
	a.pairs = w->bp.pairs.ptr;
	a.ctr = w->dCounters.ptr;
	a.pose = w->dPose.ptr;
	a.coll = w->dCollidableIdx.ptr;
	a.collidables = w->dCollidables.ptr;
	a.convex = w->dConvex.ptr;
	a.vertices = w->dVertices.ptr;
	a.uniqueEdges = w->dUniqueEdges.ptr;
	a.faces = w->dFaces.ptr;
	a.indices = w->dIndices.ptr;
	a.childShapes = w->dChildShapes.ptr;
	a.aabbs = w->bp.aabbs.ptr;
	a.meshNodes = w->dMeshNodes.ptr;
	a.meshTris = w->dMeshTris.ptr;
	a.meshInfos = w->dMeshInfos.ptr;
	a.contacts = w->dContacts.ptr;
	a.maxContacts = w->cfg.maxContactCapacity;
	a.maxItems = (int)w->dConcavePairs.cap;
	concaveCullKernel<<<w->smCount * 8, CULL_WARPS * 32, 0, s>>>(a, reinterpret_cast<const int*>(w->dConcaveSurvivors.ptr), (int)(w->dConcaveSurvivors.cap * 4), w->dConcavePairs.ptr);
	B3_LAUNCH_CHECK();
	clampConcaveKernel<<<1, 1, 0, s>>>(w->dCounters.ptr, a.maxItems);
	B3_LAUNCH_CHECK();
	concaveQuickKernel<<<w->smCount * 4, 256, 0, s>>>(a, w->dConcavePairs.ptr, w->dConcaveSurvivors.ptr);
	B3_LAUNCH_CHECK();
	concaveSmallKernel<<<w->smCount * 8, 128, 0, s>>>(a, w->dConcaveSurvivors.ptr);
	B3_LAUNCH_CHECK();
	concaveContactKernel<<<w->smCount * 8, CC_THREADS, 0, s>>>(a, w->dConcaveSurvivors.ptr);
	B3_LAUNCH_CHECK();
	return 0;
}

}  // namespace b3b200
