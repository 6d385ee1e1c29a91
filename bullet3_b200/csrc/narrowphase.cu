// narrowphase.cu -- contact generation, one warp per overlapping pair.
//
// Replaces b3GpuNarrowPhase::computeContacts -> GpuSatCollision::
// computeConvexConvexContactsGPUSAT (b3GpuNarrowPhase.cpp:751-809,
// b3ConvexHullContact.cpp:2558-4408: ~12 kernels + 7..15 host round trips per
// step) with a short chain of kernels (quick reject -> SAT -> clip, plus the primitive and
// trimesh kernels when such shapes exist) connected by device-side work queues and counters,
// without any host synchronisation.
//
// The arithmetic follows the reference's shared CPU headers operation by
// operation (the oracle the parity tests use):
//   SAT        b3FindSeparatingAxis / b3TestSepAxis / b3ProjectAxis
//              (shared/b3FindSeparatingAxis.h:4-195)
//   clipping   b3ClipHullAgainstHull / b3ClipFaceAgainstHull / b3ClipFace
//              (shared/b3ContactConvexConvexSAT.h:20-266)
//   reduction  b3ReduceContacts (shared/b3ReduceContacts.h:4-87)
// but the work is laid out for a 32-wide warp: lanes own separating-axis
// candidates (face normals of A, of B, edge x edge), the arg-min over axes is a
// shuffle reduction that breaks ties towards the lower axis index (== the
// reference's "first strict minimum" in sequential order), Sutherland-Hodgman
// clipping runs one polygon edge per lane with ballot-compacted output in
// shared memory, and the manifold reduction is four more shuffle arg-mins.
#include "internal.h"

namespace b3b200
{
constexpr int NP_THREADS = 128;
constexpr int NP_WARPS = NP_THREADS / 32;
constexpr int MAX_POLY = 64;  // b3Config::m_maxVerticesPerFace (b3Config.h:27)
constexpr int SAT_EDGES = 96;  // world-space edge directions staged per hull (a 32-vertex triangulated hull has 90)
constexpr int SUP_K = 2;      // support vertices per hull for the tight edge-axis bound
constexpr int WORK_CHUNK = 4;  // items a warp claims per atomic in the warp-per-item kernels
#define FULL 0xffffffffu

struct HullRef
{
	float4 localCenter;
	float radius;  // inscribed radius about localCenter (set at registration)
	int faceOffset, numFaces, numVertices, vertexOffset, uniqueEdgesOffset, numUniqueEdges;
};

B3_D HullRef loadHull(const b3b200_convex_polyhedron* __restrict__ convex, int shapeIndex)
{
	const b3b200_convex_polyhedron* h = &convex[shapeIndex];
	HullRef r;
	r.localCenter = __ldg(reinterpret_cast<const float4*>(&h->localCenter));
	const int4* t = reinterpret_cast<const int4*>(&h->radius);  // radius, faceOffset, numFaces, numVertices
	int4 a = __ldg(t), b = __ldg(t + 1);                         // vertexOffset, uniqueEdgesOffset, numUniqueEdges, unused
	r.radius = __int_as_float(a.x);
	r.faceOffset = a.y;
	r.numFaces = a.z;
	r.numVertices = a.w;
	r.vertexOffset = b.x;
	r.uniqueEdgesOffset = b.y;
	r.numUniqueEdges = b.z;
	return r;
}

// b3ProjectAxis (shared/b3FindSeparatingAxis.h:4-34)
B3_D void projectAxis(const HullRef& hull, const float4& pos, const float4& orn, const float4& dir, const float4* __restrict__ vertices, float& mn, float& mx)
{
	mn = FLT_MAX;
	mx = -FLT_MAX;
	const float4 localDir = quatRotate(quatInverse(orn), dir);
	const float offset = dot3(pos, dir);
	const float4* v = vertices + hull.vertexOffset;
	for (int i = 0; i < hull.numVertices; i++)
	{
		float dp = dot3(__ldg(&v[i]), localDir);
		// "if (dp < mn) mn = dp" as one FMNMX: same value (only the sign of a zero can differ, which no later
		// operation on mn / mx observes)
		mn = fminf(mn, dp);
		mx = fmaxf(mx, dp);
	}
	if (mn > mx)
	{
		float t = mn;
		mn = mx;
		mx = t;
	}
	mn += offset;
	mx += offset;
}

// b3TestSepAxis (shared/b3FindSeparatingAxis.h:36-55)
B3_D bool testSepAxis(const HullRef& hA, const HullRef& hB, const float4& posA, const float4& ornA, const float4& posB, const float4& ornB,
					  const float4& axis, const float4* __restrict__ vertices, float& depth)
{
	float min0, max0, min1, max1;
	projectAxis(hA, posA, ornA, axis, vertices, min0, max0);
	projectAxis(hB, posB, ornB, axis, vertices, min1, max1);
	if (max0 < min1 || max1 < min0) return false;
	float d0 = max0 - min1;
	float d1 = max1 - min0;
	depth = d0 < d1 ? d0 : d1;
	return true;
}

// lexicographic (value, index) arg-min across the warp; index < 0 = "no candidate"
B3_D void warpArgMin(float& d, int& k)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		float od = __shfl_xor_sync(FULL, d, o);
		int ok = __shfl_xor_sync(FULL, k, o);
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
		float od = __shfl_xor_sync(FULL, d, o);
		int ok = __shfl_xor_sync(FULL, k, o);
		if (ok >= 0 && (k < 0 || od > d || (od == d && ok < k)))
		{
			d = od;
			k = ok;
		}
	}
}

B3_D bool almostZero(const float4& v)
{
	// b3IsAlmostZero (shared/b3Float4.h:58-63); x > 1e-6 (double) <=> x > 1e-6f for float x
	return !(fabsf(v.x) > 1e-6f || fabsf(v.y) > 1e-6f || fabsf(v.z) > 1e-6f);
}

B3_D float4 lerp3(const float4& a, const float4& b, float t)
{
	return mk4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, 0.f);
}

// One Sutherland-Hodgman pass (b3ClipFace, shared/b3ContactConvexConvexSAT.h:20-68),
// one polygon edge per lane, outputs compacted in order.  Returns the new vertex count.
B3_D int clipFaceWarp(const float4* in, int numIn, const float4& n, float eq, float4* out, int lane)
{
	if (numIn < 2) return 0;
	int numOut = 0;
	for (int base = 0; base < numIn; base += 32)
	{
		int ve = base + lane;
		int c = 0;
		float4 o0 = mk4(0, 0, 0), o1 = mk4(0, 0, 0);
		if (ve < numIn)
		{
			float4 first = in[ve == 0 ? numIn - 1 : ve - 1];
			float4 end = in[ve];
			float ds = dot3(n, first) + eq;
			float de = dot3(n, end) + eq;
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
		unsigned int m1 = __ballot_sync(FULL, c >= 1);
		unsigned int m2 = __ballot_sync(FULL, c == 2);
		unsigned int lt = (1u << lane) - 1u;
		int pos = numOut + __popc(m1 & lt) + __popc(m2 & lt);
		if (c >= 1 && pos < MAX_POLY) out[pos] = o0;
		if (c == 2 && pos + 1 < MAX_POLY) out[pos + 1] = o1;
		numOut += __popc(m1) + __popc(m2);
	}
	__syncwarp();
	return numOut < MAX_POLY ? numOut : MAX_POLY;
}

struct NpArgs
{
	const b3b200_int4* pairs;
	b3b200_int4* pairsOut;
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
	const float4* childSpheres;  // per child shape: bounding-sphere centre in the compound's frame, radius in w (< 0: not a hull)
	b3b200_contact4* contacts;
	int maxContacts;
	int maxWorkItems;
	float clipMin, clipMax;
};

// One side of a (child) pair: the convex hull and its world transform.  For a child of a compound the
// transform is composed exactly like the reference kernels do (sat.cl:836-862,
// b3ConvexHullContact.cpp:1806-1832): pos' = quatRotate(orn, childPos) + pos, orn' = orn * childOrn.
struct Side
{
	int shape;
	float4 pos, orn;
	float invMass;
};
B3_D bool resolveSide(const NpArgs& a, int body, int child, Side& s)
{
	float4 pos = a.pose[2 * body], orn = a.pose[2 * body + 1];
	s.invMass = pos.w;
	int coll;
	if (child >= 0)
	{
		const b3b200_child_shape* ch = &a.childShapes[child];
		const float4 cp = __ldg(reinterpret_cast<const float4*>(&ch->childPosition));
		const float4 co = __ldg(reinterpret_cast<const float4*>(&ch->childOrientation));
		const float4 r = quatRotate(orn, cp);
		pos = mk4(r.x + pos.x, r.y + pos.y, r.z + pos.z, 0.f);
		orn = quatMul(orn, co);
		coll = __ldg(&ch->shapeIndex);
	}
	else
		coll = a.coll[body];
	if (coll < 0 || __ldg(&a.collidables[coll].shapeType) != B3B200_SHAPE_CONVEX_HULL) return false;
	s.shape = __ldg(&a.collidables[coll].shapeIndex);
	pos.w = 0.f;
	s.pos = pos;
	s.orn = orn;
	return true;
}

// convex hull vs convex hull: b3ContactConvexConvexSAT (shared/b3ContactConvexConvexSAT.h:407-484)
// Part 1: b3FindSeparatingAxis.  Returns false when the hulls are separated; otherwise *sepOut is the
// minimum-penetration axis (the reference's sepNormalWorldSpace).
B3_D bool satWarp(const NpArgs& a, int shapeA, int shapeB, float4 posA, float4 ornA, float4 posB, float4 ornB,
				  float4* bufA, float4* bufB, float4* sup, int* queue, unsigned char* aliveA, unsigned char* aliveB, int lane, float4* sepOut)
{
	posA.w = 0.f;
	posB.w = 0.f;
	const HullRef hA = loadHull(a.convex, shapeA);
	const HullRef hB = loadHull(a.convex, shapeB);

	// ---- b3FindSeparatingAxis
	const float4 c0 = transformPoint(hA.localCenter, posA, ornA);
	const float4 c1 = transformPoint(hB.localCenter, posB, ornB);
	const float4 deltaC2 = sub3(c0, c1);

	const int nFA = hA.numFaces, nFB = hB.numFaces, nEA = hA.numUniqueEdges, nEB = hB.numUniqueEdges;
	const int nF = nFA + nFB;
	const int total = nF + nEA * nEB;
	const float invNEB = 1.0f / (float)(nEB > 0 ? nEB : 1);
	float bestD = FLT_MAX;
	int bestK = -1;
	float4 bestAxis = mk4(0, 0, 0);

	// World-space edge directions are rotated once per hull (the same arithmetic the reference
	// repeats for every pair of edges) and staged in shared memory.
	const bool staged = nEA <= SAT_EDGES && nEB <= SAT_EDGES;
	if (staged)
	{
		for (int e = lane; e < nEA + nEB; e += 32)
		{
			const bool onA = e < nEA;
			const float4 ed = __ldg(&a.uniqueEdges[onA ? hA.uniqueEdgesOffset + e : hB.uniqueEdgesOffset + (e - nEA)]);
			const float4 w = quatRotate(onA ? ornA : ornB, ed);
			if (onA)
				bufA[e] = w;
			else
				bufB[e - nEA] = w;
		}
		__syncwarp();
	}
	// Exact-safe skip for edge x edge axes: with r = inscribed radii about the hull centres, the
	// overlap depth along a unit axis n is >= rA + rB - |deltaC2 . n|.  An axis whose lower bound
	// already exceeds the best depth found so far (2e-4 relative safety margin, >> FP32 rounding)
	// cannot separate and cannot become the strict minimum: the reference's result is unchanged.
	const float rsum = (hA.radius + hB.radius) * (1.0f - 2e-4f);
	float curMin = FLT_MAX;  // warp-uniform upper bound of the final minimum depth

	// Tighter exact-safe skip for hull pairs with many edge pairs.  The support of a hull along n is at least the
	// largest projection of ANY subset of its vertices, so with a_k / b_k = SUP_K vertices of A / B (relative to the
	// hull centres c0 / c1) the overlap along an oriented unit axis n (deltaC2 . n >= 0) obeys
	//     maxB - minA >= max_k(b_k . n) - min_k(a_k . n) - deltaC2 . n      maxA - minB >= deltaC2 . n + rA + rB
	// The subset is the SUP_K vertices of each hull that reach furthest towards the other hull, which are the
	// supports for the axes that matter (those roughly along deltaC2).  An axis whose lower bound exceeds the best
	// depth so far by more than `supEps` (>> the FP32 error of either evaluation) cannot separate and cannot become
	// the strict minimum, so the reference's result is unchanged.
	const bool tight = nEA * nEB >= 128;
	const float supEps = 1e-3f + 1e-6f * (fabsf(posA.x) + fabsf(posA.y) + fabsf(posA.z) + fabsf(posB.x) + fabsf(posB.y) + fabsf(posB.z));
	if (tight)
	{
#pragma unroll 1
		for (int side = 0; side < 2; side++)
		{
			const HullRef& h = side == 0 ? hA : hB;
			const float4 orn = side == 0 ? ornA : ornB;
			const float4 dirL = quatRotate(quatInverse(orn), side == 0 ? neg3(deltaC2) : deltaC2);
			const int nv = h.numVertices < 64 ? h.numVertices : 64;
			unsigned long long taken = 0ull;
			for (int kk = 0; kk < SUP_K; kk++)
			{
				float best = -FLT_MAX;
				int bi = -1;
				for (int i = lane; i < nv; i += 32)
				{
					if ((taken >> i) & 1ull) continue;
					const float sc = dot3(__ldg(&a.vertices[h.vertexOffset + i]), dirL);
					if (sc > best)
					{
						best = sc;
						bi = i;
					}
				}
				warpArgMax(best, bi);
				if (bi < 0) bi = 0;  // fewer than SUP_K vertices: repeat one (still a valid subset)
				taken |= 1ull << bi;
				if (lane == 0)
				{
					const float4 v = __ldg(&a.vertices[h.vertexOffset + bi]);
					sup[side * SUP_K + kk] = quatRotate(orn, sub3(v, h.localCenter));
				}
			}
		}
		__syncwarp();
	}

	// Candidate axes (index k = position in the reference's sequential order: faces of A, faces
	// of B, edge pairs) are filtered 32 at a time and compacted into a per-warp queue, so the
	// expensive projection test always runs with (nearly) all 32 lanes busy.
	int qn = 0;
	const unsigned int ltMask = (1u << lane) - 1u;
	// Phase 0 generates the face candidates, phase 1 the edge x edge candidates; both feed the same queue.
	// Edge pairs are pruned by rows and columns first: the cheap per-pair bound below drops (e0, e1) when
	// (deltaC2 . (e0 x e1))^2 < 0.999 S^2 |e0 x e1|^2.  Since deltaC2 . (e0 x e1) = (deltaC2 x e0) . e1 and (deltaC2 x e0) is
	// perpendicular to e0, the left side is at most |deltaC2 x e0|^2 |e0 x e1|^2 / |e0|^2: an edge e0 with
	// |deltaC2 x e0|^2 < 0.99 S^2 |e0|^2 (the 1 % gap dwarfs the FP32 error of either side) loses EVERY pair it is part of, and
	// likewise an edge e1.  Only the surviving rows x columns are enumerated; exactly the pairs that would have passed the
	// per-pair bound reach the later tests, so the result is unchanged.  S is taken when the face candidates have all been
	// generated (any earlier, larger curMin only prunes less).
	int phase = 0, next = 0, limit = nF;
	int nAA = nEA, nAB = nEB;
	bool useLists = false;
	float invNAB = invNEB;
	for (;;)
	{
		while (qn < 32 && next < limit)
		{
			const int p = next + lane;
			bool cand = false;
			int k = p;
			if (phase == 0)
			{
				if (p < nF)
				{
					const b3b200_face* f = p < nFA ? &a.faces[hA.faceOffset + p] : &a.faces[hB.faceOffset + (p - nFA)];
					// a face whose normal is bitwise +-equal to an earlier face's gives the identical depth
					// and can never win the strict "d < dmin" (flag set at registration, world.cu)
					cand = __ldg(&f->pad1) == 0;
				}
			}
			else if (p < limit)
			{
				const int i0 = __float2int_rz(((float)p + 0.5f) * invNAB), i1 = p - i0 * nAB;  // == p / nAB (exact for these ranges), without the integer division
				const int e0 = useLists ? (int)aliveA[i0] : i0, e1 = useLists ? (int)aliveB[i1] : i1;
				k = nF + e0 * nEB + e1;
				const float4 edge0 = staged ? bufA[e0] : quatRotate(ornA, __ldg(&a.uniqueEdges[hA.uniqueEdgesOffset + e0]));
				const float4 edge1 = staged ? bufB[e1] : quatRotate(ornB, __ldg(&a.uniqueEdges[hB.uniqueEdgesOffset + e1]));
				const float4 cr = cross3(edge0, edge1);
				cand = !almostZero(cr);
				const float S = rsum - curMin;
				if (cand && S > 0.f)
				{
					const float dd = dot3(deltaC2, cr);
					const float len2 = dot3(cr, cr);
					if (dd * dd < S * S * len2 * 0.999f) cand = false;
					if (cand && tight)
					{
						float mnA = FLT_MAX, mxA = -FLT_MAX, mnB = FLT_MAX, mxB = -FLT_MAX;
#pragma unroll
						for (int kk = 0; kk < SUP_K; kk++)
						{
							const float pa = dot3(sup[kk], cr), pb = dot3(sup[SUP_K + kk], cr);
							mnA = fminf(mnA, pa);
							mxA = fmaxf(mxA, pa);
							mnB = fminf(mnB, pb);
							mxB = fmaxf(mxB, pb);
						}
						const float m = sqrtf(len2);
						const float ddAbs = fabsf(dd);
						// oriented axis = +-cr / m with deltaC2 . axis >= 0
						const float l1 = (dd < 0.f ? (mxA - mnB) : (mxB - mnA)) - ddAbs;
						const float l2 = ddAbs + rsum * m;
						if (fminf(l1, l2) > (curMin + supEps) * m) cand = false;
					}
				}
			}
			const unsigned int m = __ballot_sync(FULL, cand);
			if (cand) queue[qn + __popc(m & ltMask)] = k;
			qn += __popc(m);
			next += 32;
			__syncwarp();
		}
		if (phase == 0 && next >= limit && (qn < 32 || rsum - curMin > 0.f))
		{
			// all faces generated: switch to the edge pairs (unless no face has been projected yet -- then one round first, so
			// that there is a bound to prune with)
			if (!(qn > 0 && curMin == FLT_MAX))
			{
				const float S0 = rsum - curMin;
				useLists = staged && S0 > 0.f && nEA * nEB >= 64;
				if (useLists)
				{
					nAA = 0;
					nAB = 0;
					for (int base = 0; base < nEA + nEB; base += 32)
					{
						const int e = base + lane;
						bool alive = false;
						const bool onA = e < nEA;
						if (e < nEA + nEB)
						{
							const float4 ed = onA ? bufA[e] : bufB[e - nEA];
							const float4 u = cross3(deltaC2, ed);
							alive = dot3(u, u) >= 0.99f * S0 * S0 * dot3(ed, ed);
						}
						const unsigned int mA = __ballot_sync(FULL, alive && onA), mB = __ballot_sync(FULL, alive && !onA);
						if (alive && onA) aliveA[nAA + __popc(mA & ltMask)] = (unsigned char)e;
						if (alive && !onA) aliveB[nAB + __popc(mB & ltMask)] = (unsigned char)(e - nEA);
						nAA += __popc(mA);
						nAB += __popc(mB);
					}
					__syncwarp();
				}
				phase = 1;
				next = 0;
				limit = nAA * nAB;
				invNAB = 1.0f / (float)(nAB > 0 ? nAB : 1);
				if (qn < 32 && limit > 0) continue;  // top the queue up with edge candidates before projecting
			}
		}
		if (qn == 0) break;
		const int take = qn < 32 ? qn : 32;
		bool separated = false;
		float d = FLT_MAX;
		if (lane < take)
		{
			const int k = queue[lane];
			float4 axis;
			if (k < nF)
			{
				const bool onA = k < nFA;
				const b3b200_face* f = onA ? &a.faces[hA.faceOffset + k] : &a.faces[hB.faceOffset + (k - nFA)];
				axis = quatRotate(onA ? ornA : ornB, __ldg(reinterpret_cast<const float4*>(&f->plane)));
			}
			else
			{
				const int e = k - nF;
				const int e0 = __float2int_rz(((float)e + 0.5f) * invNEB), e1 = e - e0 * nEB;
				const float4 edge0 = staged ? bufA[e0] : quatRotate(ornA, __ldg(&a.uniqueEdges[hA.uniqueEdgesOffset + e0]));
				const float4 edge1 = staged ? bufB[e1] : quatRotate(ornB, __ldg(&a.uniqueEdges[hB.uniqueEdgesOffset + e1]));
				axis = normalized3(cross3(edge0, edge1));
			}
			if (dot3(deltaC2, axis) < 0) axis = mk4(axis.x * -1.f, axis.y * -1.f, axis.z * -1.f);
			if (!testSepAxis(hA, hB, posA, ornA, posB, ornB, axis, a.vertices, d))
				separated = true;
			else if (d < bestD || (d == bestD && k < bestK))
			{
				bestD = d;
				bestK = k;
				bestAxis = axis;
			}
		}
		if (__any_sync(FULL, separated)) return false;
		{
			float m = separated ? FLT_MAX : d;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
			curMin = fminf(curMin, m);
		}
		// pop the processed entries
		const int rest = qn - take;
		int moved = 0;
		if (lane < rest) moved = queue[take + lane];
		__syncwarp();
		if (lane < rest) queue[lane] = moved;
		qn = rest;
		__syncwarp();
	}
	const int myK = bestK;
	warpArgMin(bestD, bestK);
	if (bestK < 0) return false;
	const int src = __ffs(__ballot_sync(FULL, myK == bestK)) - 1;
	float4 sep = mk4(__shfl_sync(FULL, bestAxis.x, src), __shfl_sync(FULL, bestAxis.y, src), __shfl_sync(FULL, bestAxis.z, src));
	if (dot3(neg3(deltaC2), sep) > 0.0f) sep = neg3(sep);
	*sepOut = sep;
	return true;
}

// Part 2: b3ClipHullHullSingle -- clip the incident face of B against the reference face of A, reduce to
// <= 4 points, append the contact.
B3_D void clipWarp(const NpArgs& a, int pairIndex, int bodyA, int bodyB, int shapeA, int shapeB, int childA, int childB,
				   float4 posA, float4 ornA, float4 posB, float4 ornB, float invMassA, float invMassB, const float4 sep,
				   float4* bufA, float4* bufB, int lane)
{
	posA.w = 0.f;
	posB.w = 0.f;
	const HullRef hA = loadHull(a.convex, shapeA);
	const HullRef hB = loadHull(a.convex, shapeB);
	const int nFA = hA.numFaces, nFB = hB.numFaces;

	// b3ClipHullHullSingle round-trips both orientations through a b3Transform
	// (setRotation -> getRotation, shared/b3ContactConvexConvexSAT.h:323-337); the
	// clipper below sees those re-derived quaternions.
	ornA = quatFromMat(matFromQuat(ornA));
	ornB = quatFromMat(matFromQuat(ornB));

	// ---- b3ClipHullAgainstHull: incident face of B = most aligned with sep
	int closestFaceB = -1;
	{
		float dmax = -FLT_MAX;
		for (int f = lane; f < nFB; f += 32)
		{
			float4 normal = __ldg(reinterpret_cast<const float4*>(&a.faces[hB.faceOffset + f].plane));
			float d = dot3(quatRotate(ornB, normal), sep);
			if (d > dmax)
			{
				dmax = d;
				closestFaceB = f;
			}
		}
		warpArgMax(dmax, closestFaceB);
	}
	if (closestFaceB < 0) return;
	int numVertsIn;
	{
		const b3b200_face* polyB = &a.faces[hB.faceOffset + closestFaceB];
		const int idxOff = __ldg(&polyB->indexOffset);
		numVertsIn = __ldg(&polyB->numIndices);
		if (numVertsIn > MAX_POLY) numVertsIn = MAX_POLY;
		for (int e = lane; e < numVertsIn; e += 32)
		{
			float4 b = __ldg(&a.vertices[hB.vertexOffset + __ldg(&a.indices[idxOff + e])]);
			bufA[e] = transformPoint(b, posB, ornB);
		}
		__syncwarp();
	}

	// ---- b3ClipFaceAgainstHull: reference face of A = least aligned with sep
	int closestFaceA = -1;
	{
		float dmin = FLT_MAX;
		for (int f = lane; f < nFA; f += 32)
		{
			float4 normal = __ldg(reinterpret_cast<const float4*>(&a.faces[hA.faceOffset + f].plane));
			float d = dot3(quatRotate(ornA, mk4(normal.x, normal.y, normal.z)), sep);
			if (d < dmin)
			{
				dmin = d;
				closestFaceA = f;
			}
		}
		warpArgMin(dmin, closestFaceA);
	}
	if (closestFaceA < 0) return;

	const b3b200_face* polyA = &a.faces[hA.faceOffset + closestFaceA];
	const float4 planeA = __ldg(reinterpret_cast<const float4*>(&polyA->plane));
	const int idxOffA = __ldg(&polyA->indexOffset);
	const int numVerticesA = __ldg(&polyA->numIndices);
	const float4 planeNormalA = mk4(planeA.x, planeA.y, planeA.z);
	const float4 worldPlaneAnormal1 = quatRotate(ornA, planeNormalA);

	float4* pIn = bufA;
	float4* pOut = bufB;
	for (int e0 = 0; e0 < numVerticesA; e0++)
	{
		const float4 va = __ldg(&a.vertices[hA.vertexOffset + __ldg(&a.indices[idxOffA + e0])]);
		const float4 vb = __ldg(&a.vertices[hA.vertexOffset + __ldg(&a.indices[idxOffA + ((e0 + 1) % numVerticesA)])]);
		const float4 edge0 = sub3(va, vb);
		const float4 worldEdge0 = quatRotate(ornA, edge0);
		const float4 planeNormalWS = neg3(cross3(worldEdge0, worldPlaneAnormal1));
		const float4 worldA1 = transformPoint(va, posA, ornA);
		const float planeEqWS = -dot3(worldA1, planeNormalWS);
		int numOut = clipFaceWarp(pIn, numVertsIn, planeNormalWS, planeEqWS, pOut, lane);
		float4* t = pOut;
		pOut = pIn;
		pIn = t;
		numVertsIn = numOut;
	}

	// ---- keep points behind the witness face (b3ContactConvexConvexSAT.h:143-173)
	int numContactsOut = 0;
	{
		const float4 planeNormalWS = worldPlaneAnormal1;
		const float planeEqWS = planeA.w - dot3(planeNormalWS, posA);
		for (int base = 0; base < numVertsIn; base += 32)
		{
			int i = base + lane;
			bool keep = false;
			float4 pt = mk4(0, 0, 0);
			if (i < numVertsIn)
			{
				pt = pIn[i];
				float depth = dot3(planeNormalWS, pt) + planeEqWS;
				if (depth <= a.clipMin) depth = a.clipMin;
				if (depth <= a.clipMax)
				{
					keep = true;
					pt.w = depth;
				}
			}
			unsigned int m = __ballot_sync(FULL, keep);
			if (keep) pOut[numContactsOut + __popc(m & ((1u << lane) - 1u))] = pt;
			numContactsOut += __popc(m);
		}
		__syncwarp();
	}
	if (numContactsOut <= 0) return;
	const float4* pts = pOut;

	// ---- b3ReduceContacts (shared/b3ReduceContacts.h:4-87)
	int idx0 = 0, idx1 = 1, idx2 = 2, idx3 = 3;
	int numPoints = numContactsOut;
	if (numContactsOut > 4)
	{
		int nPoints = numContactsOut > 64 ? 64 : numContactsOut;
		float4 center = mk4(0, 0, 0);
		for (int i = 0; i < nPoints; i++)
		{
			float4 p = pts[i];
			center.x += p.x;
			center.y += p.y;
			center.z += p.z;
		}
		{
			float s = 1.0f / (float)nPoints;
			center.x *= s;
			center.y *= s;
			center.z *= s;
		}
		float4 aVector = sub3(pts[0], center);
		float4 u = cross3(sep, aVector);
		float4 v = cross3(sep, u);
		u = normalized3(u);
		v = normalized3(v);
		const float4 nu = neg3(u), nv = neg3(v);
		float minW = FLT_MAX;
		int minIndex = -1;
		float m0 = FLT_MIN, m1 = FLT_MIN, m2 = FLT_MIN, m3 = FLT_MIN;
		int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
		for (int ie = lane; ie < nPoints; ie += 32)
		{
			float4 p = pts[ie];
			if (p.w < minW)
			{
				minW = p.w;
				minIndex = ie;
			}
			float4 r = sub3(p, center);
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

	// ---- append (b3ClipHullHullSingle, shared/b3ContactConvexConvexSAT.h:364-397)
	unsigned int slot = 0;
	if (lane == 0) slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
	slot = __shfl_sync(FULL, slot, 0);
	if (slot >= (unsigned int)a.maxContacts) return;  // clamped afterwards, OVF_CONTACTS raised
	b3b200_contact4* c = &a.contacts[slot];
	float4* cw = reinterpret_cast<float4*>(c);
	if (lane < 4)
	{
		int id = lane == 0 ? idx0 : (lane == 1 ? idx1 : (lane == 2 ? idx2 : idx3));
		float4 p = lane < numPoints ? pts[id] : mk4(0, 0, 0, 0);
		cw[lane] = p;
	}
	else if (lane == 4)
	{
		cw[4] = mk4(sep.x, sep.y, sep.z, (float)numPoints);
	}
	else if (lane == 5)
	{
		int4 t;
		t.x = (int)(0u | (45874u << 16));  // restitutionCmp = 0, frictionCmp = 45874
		t.y = 0;                            // batchIdx
		t.z = invMassA == 0.f ? -bodyA : bodyA;
		t.w = invMassB == 0.f ? -bodyB : bodyB;
		reinterpret_cast<int4*>(c)[5] = t;
	}
	else if (lane == 6)
	{
		int4 t;
		t.x = childA;
		t.y = childB;
		t.z = 0;
		t.w = 0;
		reinterpret_cast<int4*>(c)[6] = t;
	}
	if (lane == 0 && pairIndex >= 0) a.pairsOut[pairIndex].z = (int)slot;
}

// ---------------------------------------------------------------- small pairs: one THREAD per item
// Pairs of small hulls (boxes, tetrahedra, the box children of compounds: <= 8 vertices, <= 6 faces, <= 6 edge
// directions each, i.e. at most 48 axes over 16 vertices) leave most lanes of a warp idle in satKernel / clipKernel.
// They are queued separately and handled like the reference does it on the CPU: one thread runs b3FindSeparatingAxis,
// b3ClipHullAgainstHull and b3ReduceContacts serially, in the reference's statement order (shared/
// b3FindSeparatingAxis.h:57-195, shared/b3ContactConvexConvexSAT.h:20-266, 270-405, shared/b3ReduceContacts.h:4-87),
// so that 32 pairs share a warp.
constexpr int SMALL_VERTS = 8, SMALL_FACES = 6, SMALL_EDGES = 6;
constexpr int SMALL_POLY = 16;  // a face of <= 8 vertices clipped by <= 8 planes

// b3ClipFace (shared/b3ContactConvexConvexSAT.h:20-68)
B3_D int clipFaceSerial(const float4* in, int numIn, const float4& n, float eq, float4* out)
{
	int numOut = 0;
	if (numIn < 2) return 0;
	float4 first = in[numIn - 1];
	float ds = dot3(n, first) + eq;
	for (int ve = 0; ve < numIn; ve++)
	{
		const float4 end = in[ve];
		const float de = dot3(n, end) + eq;
		if (ds < 0)
		{
			if (de < 0)
			{
				if (numOut < SMALL_POLY) out[numOut++] = end;
			}
			else if (numOut < SMALL_POLY)
				out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
		}
		else if (de < 0)
		{
			if (numOut < SMALL_POLY) out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
			if (numOut < SMALL_POLY) out[numOut++] = end;
		}
		first = end;
		ds = de;
	}
	return numOut;
}

// b3ClipHullHullSingle + b3ReduceContacts + append for ONE item in one thread (the second half of the reference's
// b3ContactConvexConvexSAT, in its statement order), given the separating axis.  Works for any hull as long as the two
// chosen faces are small polygons (incident face + one vertex per clipping plane <= SMALL_POLY); returns 1 without touching
// anything when they are not -- the caller then leaves the item to the warp-per-item clipKernel.
B3_D int clipThread(const NpArgs& a, const int4 it, int bodyA, int bodyB, const Side& A, const Side& B, const HullRef& hA, const HullRef& hB, const float4& posA,
					const float4& posB, const float4& ornA, const float4& ornB, const float4& sep)
{
	// ---- b3ClipHullHullSingle: orientations round-trip through b3Transform (:323-337)
	const float4 ornA2 = quatFromMat(matFromQuat(ornA)), ornB2 = quatFromMat(matFromQuat(ornB));
	float4 bufA[SMALL_POLY], bufB[SMALL_POLY];
	// b3ClipHullAgainstHull: incident face of B
	int closestFaceB = -1;
	{
		float dmax = -FLT_MAX;
		for (int f = 0; f < hB.numFaces; f++)
		{
			const float4 normal = __ldg(reinterpret_cast<const float4*>(&a.faces[hB.faceOffset + f].plane));
			const float d = dot3(quatRotate(ornB2, normal), sep);
			if (d > dmax)
			{
				dmax = d;
				closestFaceB = f;
			}
		}
	}
	if (closestFaceB < 0) return 0;
	int numVertsIn;
	{
		const b3b200_face* polyB = &a.faces[hB.faceOffset + closestFaceB];
		const int idxOff = __ldg(&polyB->indexOffset);
		numVertsIn = __ldg(&polyB->numIndices);
		if (numVertsIn > SMALL_POLY) return 1;
		for (int e = 0; e < numVertsIn; e++) bufA[e] = transformPoint(__ldg(&a.vertices[hB.vertexOffset + __ldg(&a.indices[idxOff + e])]), posB, ornB2);
	}
	// b3ClipFaceAgainstHull: reference face of A
	int closestFaceA = -1;
	{
		float dm = FLT_MAX;
		for (int f = 0; f < hA.numFaces; f++)
		{
			const float4 normal = __ldg(reinterpret_cast<const float4*>(&a.faces[hA.faceOffset + f].plane));
			const float d = dot3(quatRotate(ornA2, mk4(normal.x, normal.y, normal.z)), sep);
			if (d < dm)
			{
				dm = d;
				closestFaceA = f;
			}
		}
	}
	if (closestFaceA < 0) return 0;
	const b3b200_face* polyA = &a.faces[hA.faceOffset + closestFaceA];
	const float4 planeA = __ldg(reinterpret_cast<const float4*>(&polyA->plane));
	const int idxOffA = __ldg(&polyA->indexOffset);
	const int numVerticesA = __ldg(&polyA->numIndices);
	if (numVertsIn + numVerticesA > SMALL_POLY) return 1;  // every clipping plane can add one vertex
	const float4 worldPlaneAnormal1 = quatRotate(ornA2, mk4(planeA.x, planeA.y, planeA.z));
	float4* pIn = bufA;
	float4* pOut = bufB;
	for (int e0 = 0; e0 < numVerticesA; e0++)
	{
		const float4 va = __ldg(&a.vertices[hA.vertexOffset + __ldg(&a.indices[idxOffA + e0])]);
		const float4 vb = __ldg(&a.vertices[hA.vertexOffset + __ldg(&a.indices[idxOffA + ((e0 + 1) % numVerticesA)])]);
		const float4 worldEdge0 = quatRotate(ornA2, sub3(va, vb));
		const float4 planeNormalWS = neg3(cross3(worldEdge0, worldPlaneAnormal1));
		const float4 worldA1 = transformPoint(va, posA, ornA2);
		const float planeEqWS = -dot3(worldA1, planeNormalWS);
		const int numOut = clipFaceSerial(pIn, numVertsIn, planeNormalWS, planeEqWS, pOut);
		float4* t = pOut;
		pOut = pIn;
		pIn = t;
		numVertsIn = numOut;
	}
	int numContactsOut = 0;
	{
		const float planeEqWS = planeA.w - dot3(worldPlaneAnormal1, posA);
		for (int i = 0; i < numVertsIn; i++)
		{
			float4 pt = pIn[i];
			float depth = dot3(worldPlaneAnormal1, pt) + planeEqWS;
			if (depth <= a.clipMin) depth = a.clipMin;
			if (depth <= a.clipMax)
			{
				pt.w = depth;
				pOut[numContactsOut++] = pt;
			}
		}
	}
	if (numContactsOut <= 0) return 0;
	const float4* pts = pOut;

	// ---- b3ReduceContacts
	int idx[4] = {0, 1, 2, 3};
	int numPoints = numContactsOut;
	if (numContactsOut > 4)
	{
		const int nP = numContactsOut;
		float4 center = mk4(0, 0, 0);
		for (int i = 0; i < nP; i++) center = add3(center, pts[i]);
		center = scale3(center, 1.0f / (float)nP);
		const float4 aVector = sub3(pts[0], center);
		float4 u = cross3(sep, aVector);
		float4 v = cross3(sep, u);
		u = normalized3(u);
		v = normalized3(v);
		float minW = FLT_MAX;
		int minIndex = -1;
		float m0 = FLT_MIN, m1 = FLT_MIN, m2 = FLT_MIN, m3 = FLT_MIN;
		for (int ie = 0; ie < nP; ie++)
		{
			if (pts[ie].w < minW)
			{
				minW = pts[ie].w;
				minIndex = ie;
			}
			const float4 r = sub3(pts[ie], center);
			float f = dot3(u, r);
			if (f < m0)
			{
				m0 = f;
				idx[0] = ie;
			}
			f = dot3(neg3(u), r);
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
			f = dot3(neg3(v), r);
			if (f < m3)
			{
				m3 = f;
				idx[3] = ie;
			}
		}
		if (idx[0] != minIndex && idx[1] != minIndex && idx[2] != minIndex && idx[3] != minIndex) idx[0] = minIndex;
		numPoints = 4;
	}

	// ---- append
	const unsigned int slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
	if (slot >= (unsigned int)a.maxContacts) return 0;
	b3b200_contact4* c = &a.contacts[slot];
	float4* cw = reinterpret_cast<float4*>(c);
	for (int i = 0; i < 4; i++) cw[i] = i < numPoints ? pts[idx[i]] : mk4(0, 0, 0, 0);
	cw[4] = mk4(sep.x, sep.y, sep.z, (float)numPoints);
	int4 t;
	t.x = (int)(0u | (45874u << 16));
	t.y = 0;
	t.z = A.invMass == 0.f ? -bodyA : bodyA;
	t.w = B.invMass == 0.f ? -bodyB : bodyB;
	reinterpret_cast<int4*>(c)[5] = t;
	reinterpret_cast<int4*>(c)[6] = make_int4(it.y, it.z, 0, 0);
	a.pairsOut[it.x].z = (int)slot;
	return 0;
}

B3_D void smallPairThread(const NpArgs& a, const int4 it)
{
	const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
	Side A, B;
	if (!resolveSide(a, bodyA, it.y, A) || !resolveSide(a, bodyB, it.z, B)) return;
	float4 posA = A.pos, posB = B.pos;
	posA.w = 0.f;
	posB.w = 0.f;
	const float4 ornA = A.orn, ornB = B.orn;
	const HullRef hA = loadHull(a.convex, A.shape), hB = loadHull(a.convex, B.shape);

	// ---- b3FindSeparatingAxis
	const float4 c0 = transformPoint(hA.localCenter, posA, ornA);
	const float4 c1 = transformPoint(hB.localCenter, posB, ornB);
	const float4 deltaC2 = sub3(c0, c1);
	float dmin = FLT_MAX;
	float4 sep = mk4(0, 0, 0);
#pragma unroll 1
	for (int side = 0; side < 2; side++)
	{
		const HullRef& h = side ? hB : hA;
		const float4 orn = side ? ornB : ornA;
		for (int i = 0; i < h.numFaces; i++)
		{
			const b3b200_face* f = &a.faces[h.faceOffset + i];
			if (__ldg(&f->pad1) != 0) continue;  // +-duplicate of an earlier normal: identical depth, never the strict minimum
			float4 n = quatRotate(orn, __ldg(reinterpret_cast<const float4*>(&f->plane)));
			if (dot3(deltaC2, n) < 0) n = mk4(n.x * -1.f, n.y * -1.f, n.z * -1.f);
			float d;
			if (!testSepAxis(hA, hB, posA, ornA, posB, ornB, n, a.vertices, d)) return;
			if (d < dmin)
			{
				dmin = d;
				sep = n;
			}
		}
	}
	{
		float4 edgeB[SMALL_EDGES];
		for (int e1 = 0; e1 < hB.numUniqueEdges; e1++) edgeB[e1] = quatRotate(ornB, __ldg(&a.uniqueEdges[hB.uniqueEdgesOffset + e1]));
		for (int e0 = 0; e0 < hA.numUniqueEdges; e0++)
		{
			const float4 edge0World = quatRotate(ornA, __ldg(&a.uniqueEdges[hA.uniqueEdgesOffset + e0]));
			for (int e1 = 0; e1 < hB.numUniqueEdges; e1++)
			{
				float4 cr = cross3(edge0World, edgeB[e1]);
				if (almostZero(cr)) continue;
				cr = normalized3(cr);
				if (dot3(deltaC2, cr) < 0) cr = mk4(cr.x * -1.f, cr.y * -1.f, cr.z * -1.f);
				float dist;
				if (!testSepAxis(hA, hB, posA, ornA, posB, ornB, cr, a.vertices, dist)) return;
				if (dist < dmin)
				{
					dmin = dist;
					sep = cr;
				}
			}
		}
	}
	if (dot3(neg3(deltaC2), sep) > 0.0f) sep = neg3(sep);

	clipThread(a, it, bodyA, bodyB, A, B, hA, hB, posA, posB, ornA, ornB, sep);
}

// The list is filled from both ends: pairs of box-like hulls (<= 3 edge directions each: 6 + 6 + 9 axes) from the front, the
// other small pairs (up to 44 axes) from the back, so that the lanes of a warp run loops of similar length.
__global__ void __launch_bounds__(128, 6) smallPairKernel(NpArgs a, const int4* __restrict__ smallItems)
{
	int nFront = (int)a.ctr[CTR_SMALL_ITEMS], nBack = (int)a.ctr[CTR_SMALL_ITEMS_BACK];
	if (nFront > a.maxWorkItems) nFront = a.maxWorkItems;
	if (nBack > a.maxWorkItems - nFront) nBack = a.maxWorkItems - nFront;
	const int numItems = nFront + nBack;
	for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < numItems; s += gridDim.x * blockDim.x)
		smallPairThread(a, smallItems[s < nFront ? s : a.maxWorkItems - 1 - (s - nFront)]);
}

// ---------------------------------------------------------------- stage 1: quick reject
// One THREAD per broadphase pair.  Convex-convex pairs and every child pair of compound shapes
// (compound x compound, compound x convex) become work items (pair, childA, childB).  For each
// it picks the face of A and the face of B that look most likely to separate (largest
// local-space dot with the centre offset) and runs the reference's own test (b3TestSepAxis,
// identical arithmetic) on just those two axes.  Both axes are members of b3FindSeparatingAxis'
// candidate list, so a separation found here is a separation the reference finds too: the reject
// is exact, not conservative.  Surviving items are appended for the warp-per-item stage.
// The reference culls child pairs with a tree-vs-tree walk of two quantized BVHs
// (findCompoundPairsKernel, sat.cl:947-1301); any conservative cull gives the same contacts.
constexpr int CULL_THREADS = 256;

B3_D int bestFace(const NpArgs& a, const HullRef& h, const float4& localDir)
{
	int best = 0;
	float bestDot = -FLT_MAX;
	for (int f = 0; f < h.numFaces; f++)
	{
		float4 n = __ldg(reinterpret_cast<const float4*>(&a.faces[h.faceOffset + f].plane));
		float d = dot3(n, localDir);
		if (d > bestDot)
		{
			bestDot = d;
			best = f;
		}
	}
	return best;
}

// true = the pair may touch (keep it)
B3_D bool quickTest(const NpArgs& a, const Side& A, const Side& B)
{
	const HullRef hA = loadHull(a.convex, A.shape);
	const HullRef hB = loadHull(a.convex, B.shape);
	const float4 c0 = transformPoint(hA.localCenter, A.pos, A.orn);
	const float4 c1 = transformPoint(hB.localCenter, B.pos, B.orn);
	const float4 deltaC2 = sub3(c0, c1);
#pragma unroll 1
	for (int which = 0; which < 2; which++)
	{
		const HullRef& h = which == 0 ? hA : hB;
		const float4 orn = which == 0 ? A.orn : B.orn;
		const int f = bestFace(a, h, quatRotate(quatInverse(orn), which == 0 ? neg3(deltaC2) : deltaC2));
		float4 normal = __ldg(reinterpret_cast<const float4*>(&a.faces[h.faceOffset + f].plane));
		float4 axis = quatRotate(orn, normal);
		if (dot3(deltaC2, axis) < 0) axis = mk4(axis.x * -1.f, axis.y * -1.f, axis.z * -1.f);
		float d;
		if (!testSepAxis(hA, hB, A.pos, A.orn, B.pos, B.orn, axis, a.vertices, d)) return false;
	}
	return true;
}

// warp-aggregated append of the kept items: small x small hull pairs go to the thread-per-item list, the rest to the
// warp-per-item list
B3_D void pushClassified(const NpArgs& a, bool keep, int small, const int4& it, int4* __restrict__ items, int4* __restrict__ smallItems, int lane)
{
	const unsigned int mg = __ballot_sync(FULL, keep && small == 0), mf = __ballot_sync(FULL, keep && small == 1), mb = __ballot_sync(FULL, keep && small == 2);
	const unsigned int lt = (1u << lane) - 1u;
	if (mg)
	{
		unsigned int slot = 0;
		if (lane == 0) slot = atomicAdd(&a.ctr[CTR_SURVIVORS], (unsigned int)__popc(mg));
		slot = __shfl_sync(FULL, slot, 0) + __popc(mg & lt);
		if (keep && small == 0 && slot < (unsigned int)a.maxWorkItems) items[slot] = it;
	}
	if (mf | mb)
	{
		// front and back cursors of the small-item list.  The ends cannot cross: every small item comes from a distinct
		// broadphase pair or raw child item, and the list has room for all of those together (world.cu).
		unsigned int sf = 0, sb = 0;
		if (lane == 0)
		{
			if (mf) sf = atomicAdd(&a.ctr[CTR_SMALL_ITEMS], (unsigned int)__popc(mf));
			if (mb) sb = atomicAdd(&a.ctr[CTR_SMALL_ITEMS_BACK], (unsigned int)__popc(mb));
		}
		sf = __shfl_sync(FULL, sf, 0) + __popc(mf & lt);
		sb = __shfl_sync(FULL, sb, 0) + __popc(mb & lt);
		if (keep && small == 1 && sf < (unsigned int)a.maxWorkItems) smallItems[sf] = it;
		if (keep && small == 2 && sb < (unsigned int)a.maxWorkItems) smallItems[(unsigned int)a.maxWorkItems - 1u - sb] = it;
	}
}

// 0 = not a pair of small hulls, 1 = both box-like (<= 3 edge directions), 2 = other small pair
B3_D int smallClass(const NpArgs& a, int shapeA, int shapeB)
{
	const HullRef hA = loadHull(a.convex, shapeA), hB = loadHull(a.convex, shapeB);
	const bool sm = hA.numVertices <= SMALL_VERTS && hA.numFaces <= SMALL_FACES && hA.numUniqueEdges <= SMALL_EDGES && hB.numVertices <= SMALL_VERTS &&
					hB.numFaces <= SMALL_FACES && hB.numUniqueEdges <= SMALL_EDGES;
	if (!sm) return 0;
	return hA.numUniqueEdges <= 3 && hB.numUniqueEdges <= 3 ? 1 : 2;
}

// conservative world-space bounding sphere of one side: centre = hull centre, radius = circumscribed radius about it
// (stored in the convex entry's unused word at registration, world.cu)
B3_D float4 boundSphere(const NpArgs& a, const Side& s, float& radius)
{
	const b3b200_convex_polyhedron* h = &a.convex[s.shape];
	radius = __int_as_float(__ldg(&h->unused));
	return transformPoint(__ldg(reinterpret_cast<const float4*>(&h->localCenter)), s.pos, s.orn);
}

// stage 1b: the exact quick reject for the raw child items of compound pairs, one THREAD per item
__global__ void __launch_bounds__(CULL_THREADS) npChildCullKernel(NpArgs a, const int4* __restrict__ rawItems, int4* __restrict__ items,
																  int4* __restrict__ smallItems)
{
	int numRaw = (int)a.ctr[CTR_COMPOUND_PAIRS];
	if (numRaw > a.maxWorkItems) numRaw = a.maxWorkItems;
	const int lane = threadIdx.x & 31;
	for (int base = blockIdx.x * CULL_THREADS; base < numRaw; base += gridDim.x * CULL_THREADS)
	{
		const int r = base + threadIdx.x;
		bool keep = false;
		int small = 0;
		int4 it = make_int4(0, 0, 0, 0);
		if (r < numRaw)
		{
			it = rawItems[r];
			Side A, B;
			if (resolveSide(a, a.pairs[it.x].x, it.y, A) && resolveSide(a, a.pairs[it.x].y, it.z, B))
			{
				keep = quickTest(a, A, B);
				small = keep ? smallClass(a, A.shape, B.shape) : 0;
			}
		}
		pushClassified(a, keep, small, it, items, smallItems, lane);
	}
}

// Per-warp queues keep the three kinds of work apart, so that each runs with (nearly) full lanes instead of idling the lanes
// that hold another kind of pair: convex pairs of two small hulls (6 faces / 8 vertices: short loops), convex pairs with a larger
// hull (up to 60 faces / 32 vertices), and pairs with a compound, whose child pairs are enumerated one per LANE
// (ncu before: the child loops ran at 4-6 of 32 lanes, the quick reject's face search at 6).
constexpr int CULL_WARPS = CULL_THREADS / 32;
constexpr int CQ_TAKE = 16;  // compound pairs expanded together

struct CullCompound
{
	int pair, firstA, nA, firstB, nB;  // first child shape (-1: a lone hull), number of children
};

__global__ void __launch_bounds__(CULL_THREADS, 4) npCullKernel(NpArgs a, int4* __restrict__ items, int4* __restrict__ rawItems, int* __restrict__ meshPairs,
															 int maxMeshPairs, int4* __restrict__ smallItems)
{
	const int numPairs = (int)a.ctr[CTR_PAIRS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__shared__ int queueAll[CULL_WARPS][3][64];          // 0: small convex pairs, 1: convex pairs with a larger hull, 2: pairs with a compound
	__shared__ CullCompound compAll[CULL_WARPS][CQ_TAKE];
	__shared__ float4 frameAll[CULL_WARPS][CQ_TAKE][2][5];  // per queued compound pair and side: rotation rows, position, lone sphere
	__shared__ int offAll[CULL_WARPS][CQ_TAKE + 1];
	int(*queue)[64] = queueAll[warp];
	CullCompound* comp = compAll[warp];
	int* off = offAll[warp];
	int qn0 = 0, qn1 = 0, qn2 = 0;  // warp-uniform
	const unsigned int lt = (1u << lane) - 1u;
	for (int base = blockIdx.x * CULL_THREADS; base < numPairs || (qn0 | qn1 | qn2) > 0; base += gridDim.x * CULL_THREADS)
	{
		const int p = base + threadIdx.x;
		int kind = -1;
		if (base < numPairs && p < numPairs)
		{
			const int bodyA = a.pairs[p].x, bodyB = a.pairs[p].y;
			const int cA = a.coll[bodyA], cB = a.coll[bodyB];
			if (cA >= 0 && cB >= 0)
			{
				const int typeA = __ldg(&a.collidables[cA].shapeType), typeB = __ldg(&a.collidables[cB].shapeType);
				if (typeA == B3B200_SHAPE_CONVEX_HULL && typeB == B3B200_SHAPE_CONVEX_HULL)
				{
					Side A, B;
					if (resolveSide(a, bodyA, -1, A) && resolveSide(a, bodyB, -1, B))
					{
						// bounding spheres first (conservative, like the child pairs below); the exact quick reject runs below on
						// the warp's queues of such pairs, 32 at a time
						float rA, rB;
						const float4 sA = boundSphere(a, A, rA), sB = boundSphere(a, B, rB);
						const float4 d = sub3(sA, sB);
						const float rr = (rA + rB) * 1.001f + 1e-3f;
						if (dot3(d, d) <= rr * rr)
						{
							const b3b200_convex_polyhedron *hA = &a.convex[A.shape], *hB = &a.convex[B.shape];
							kind = __ldg(&hA->numVertices) <= SMALL_VERTS && __ldg(&hA->numFaces) <= SMALL_FACES && __ldg(&hB->numVertices) <= SMALL_VERTS &&
										   __ldg(&hB->numFaces) <= SMALL_FACES
									   ? 0
									   : 1;
						}
					}
				}
				else if ((typeA == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS || typeA == B3B200_SHAPE_CONVEX_HULL) &&
						 (typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS || typeB == B3B200_SHAPE_CONVEX_HULL))
				{
					// at least one compound: its child pairs are expanded below (two static bodies never collide, sat.cl:975-978)
					if (!(a.pose[2 * bodyA].w == 0.f && a.pose[2 * bodyB].w == 0.f)) kind = 2;
				}
				else if (meshPairs && typeA == B3B200_SHAPE_CONCAVE_TRIMESH &&
						 (typeB == B3B200_SHAPE_CONVEX_HULL || typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS || typeB == B3B200_SHAPE_SPHERE))
				{
					// trimesh (as A, b3BvhTraversal.h:35) x hull / compound: listed for concaveCullKernel, which then
					// need not scan all pairs again
					const unsigned int slot = atomicAdd(&a.ctr[CTR_MESH_PAIRS], 1u);
					if (slot < (unsigned int)maxMeshPairs) meshPairs[slot] = p;
				}
			}
		}
		{
			const unsigned int m0 = __ballot_sync(FULL, kind == 0), m1 = __ballot_sync(FULL, kind == 1), m2 = __ballot_sync(FULL, kind == 2);
			if (kind == 0) queue[0][qn0 + __popc(m0 & lt)] = p;
			if (kind == 1) queue[1][qn1 + __popc(m1 & lt)] = p;
			if (kind == 2) queue[2][qn2 + __popc(m2 & lt)] = p;
			qn0 += __popc(m0);
			qn1 += __popc(m1);
			qn2 += __popc(m2);
			__syncwarp();
		}
		const bool last = base + (int)(gridDim.x * CULL_THREADS) >= numPairs;
		// ---- the exact quick reject on a full warp's worth of queued convex pairs of one kind (or on what is left at the end)
#pragma unroll 1
		for (int k = 0; k < 2; k++)
		{
			int qn = k ? qn1 : qn0;
			while (qn >= 32 || (last && qn > 0))
			{
				const int take = qn < 32 ? qn : 32;
				bool keep = false;
				int small = 0;
				int q = -1;
				if (lane < take)
				{
					q = queue[k][qn - take + lane];
					Side A, B;
					if (resolveSide(a, a.pairs[q].x, -1, A) && resolveSide(a, a.pairs[q].y, -1, B))
					{
						keep = quickTest(a, A, B);
						small = keep ? smallClass(a, A.shape, B.shape) : 0;
					}
				}
				qn -= take;
				__syncwarp();
				pushClassified(a, keep, small, make_int4(q, -1, -1, 0), items, smallItems, lane);
			}
			if (k)
				qn1 = qn;
			else
				qn0 = qn;
		}
		// ---- child pairs of the queued compound pairs, one per lane.  Child pairs whose bounding spheres touch go to the raw
		// child-item queue; npChildCullKernel runs the exact quick reject on them.  The spheres come from a per-child table in the
		// compound's frame (world.cu), so a child costs one matrix-vector product here instead of a composed transform.
		while (qn2 >= CQ_TAKE || (last && qn2 > 0))
		{
			const int take = qn2 < CQ_TAKE ? qn2 : CQ_TAKE;
			int cnt = 0;
			if (lane < take)
			{
				const int q = queue[2][qn2 - take + lane];
				const int bodyA = a.pairs[q].x, bodyB = a.pairs[q].y;
				const int cA = a.coll[bodyA], cB = a.coll[bodyB];
				const bool compA = __ldg(&a.collidables[cA].shapeType) == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS,
						   compB = __ldg(&a.collidables[cB].shapeType) == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS;
				CullCompound c;
				c.pair = q;
				c.firstA = compA ? __ldg(&a.collidables[cA].shapeIndex) : -1;
				c.nA = compA ? __ldg(&a.collidables[cA].numChildShapes) : 1;
				c.firstB = compB ? __ldg(&a.collidables[cB].shapeIndex) : -1;
				c.nB = compB ? __ldg(&a.collidables[cB].numChildShapes) : 1;
				comp[lane] = c;
				cnt = c.nA * c.nB;
#pragma unroll
				for (int side = 0; side < 2; side++)
				{
					const int body = side ? bodyB : bodyA, cc = side ? cB : cA;
					const Mat3 m = matFromQuat(a.pose[2 * body + 1]);
					float4* f = frameAll[warp][lane][side];
					f[0] = m.r0;
					f[1] = m.r1;
					f[2] = m.r2;
					f[3] = a.pose[2 * body];
					float4 lone = mk4(0, 0, 0, -1.f);
					if (!(side ? compB : compA))
					{
						const b3b200_convex_polyhedron* h = &a.convex[__ldg(&a.collidables[cc].shapeIndex)];
						lone = __ldg(reinterpret_cast<const float4*>(&h->localCenter));
						lone.w = __int_as_float(__ldg(&h->unused));
					}
					f[4] = lone;
				}
			}
			// exclusive scan of the child-pair counts
			int incl = cnt;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const int t = __shfl_up_sync(FULL, incl, o);
				if (lane >= o) incl += t;
			}
			if (lane < take) off[lane] = incl - cnt;
			const int total = __shfl_sync(FULL, incl, take - 1);
			if (lane == 0) off[take] = total;
			qn2 -= take;
			__syncwarp();
			for (int cb = 0; cb < total; cb += 32)
			{
				const int c = cb + lane;
				bool hit = false;
				int4 it = make_int4(0, 0, 0, 0);
				if (c < total)
				{
					int sl = 0;
					while (off[sl + 1] <= c) sl++;  // (<= CQ_TAKE entries)
					const CullCompound cc = comp[sl];
					const int r = c - off[sl];
					const int i = r / cc.nB, j = r - i * cc.nB;
					const float4* fA = frameAll[warp][sl][0];
					const float4* fB = frameAll[warp][sl][1];
					const float4 lA = cc.firstA >= 0 ? __ldg(&a.childSpheres[cc.firstA + i]) : fA[4];
					const float4 lB = cc.firstB >= 0 ? __ldg(&a.childSpheres[cc.firstB + j]) : fB[4];
					if (lA.w >= 0.f && lB.w >= 0.f)
					{
						Mat3 mA, mB;
						mA.r0 = fA[0];
						mA.r1 = fA[1];
						mA.r2 = fA[2];
						mB.r0 = fB[0];
						mB.r1 = fB[1];
						mB.r2 = fB[2];
						const float4 sA = add3(matMulVec(mA, lA), fA[3]), sB = add3(matMulVec(mB, lB), fB[3]);
						const float4 d = sub3(sA, sB);
						const float rr = (lA.w + lB.w) * 1.001f + 2e-3f;
						hit = dot3(d, d) <= rr * rr;
						it = make_int4(cc.pair, cc.firstA >= 0 ? cc.firstA + i : -1, cc.firstB >= 0 ? cc.firstB + j : -1, 0);
					}
				}
				const unsigned int m = __ballot_sync(FULL, hit);
				if (m)
				{
					unsigned int slot = 0;
					if (lane == 0) slot = atomicAdd(&a.ctr[CTR_COMPOUND_PAIRS], (unsigned int)__popc(m));
					slot = __shfl_sync(FULL, slot, 0) + __popc(m & lt);
					if (hit && slot < (unsigned int)a.maxWorkItems) rawItems[slot] = it;
				}
			}
			__syncwarp();
		}
	}
}

// ---------------------------------------------------------------- stage 2: SAT, one warp per work item
// Small, hot kernel (fits the instruction cache, ~85 registers): finds the minimum-penetration axis
// and appends (item, axis) to the overlap list.
__global__ void __launch_bounds__(NP_THREADS, 5) satKernel(NpArgs a, const int4* __restrict__ items, int4* __restrict__ overlapItems,
															   float4* __restrict__ overlapSep)
{
	__shared__ float4 bufAll[NP_WARPS][2][SAT_EDGES];
	__shared__ float4 supAll[NP_WARPS][2 * SUP_K];
	__shared__ int queueAll[NP_WARPS][64];
	__shared__ unsigned char aliveAll[NP_WARPS][2][SAT_EDGES];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	float4* bufA = bufAll[warp][0];
	float4* bufB = bufAll[warp][1];
	float4* sup = supAll[warp];
	int* queue = queueAll[warp];
	int numItems = (int)a.ctr[CTR_SURVIVORS];
	if (numItems > a.maxWorkItems) numItems = a.maxWorkItems;
	// Item cost varies by three orders of magnitude (box x box: 15 axes, two 32-vertex hulls: ~3000), so warps claim
	// items dynamically, WORK_CHUNK at a time, instead of striding.
	for (;;)
	{
		int base = 0;
		if (lane == 0) base = (int)atomicAdd(&a.ctr[CTR_CURSOR_SAT], (unsigned int)WORK_CHUNK);
		base = __shfl_sync(FULL, base, 0);
		if (base >= numItems) break;
		const int end = base + WORK_CHUNK < numItems ? base + WORK_CHUNK : numItems;
		for (int s = base; s < end; s++)
		{
			const int4 it = items[s];
			const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
			Side A, B;
			if (resolveSide(a, bodyA, it.y, A) && resolveSide(a, bodyB, it.z, B))
			{
				float4 sep;
				const bool hit = satWarp(a, A.shape, B.shape, A.pos, A.orn, B.pos, B.orn, bufA, bufB, sup, queue, aliveAll[warp][0], aliveAll[warp][1], lane, &sep);
				if (hit && lane == 0)
				{
					const unsigned int slot = atomicAdd(&a.ctr[CTR_OVERLAPS], 1u);
					overlapItems[slot] = it;
					overlapSep[slot] = sep;
				}
			}
			__syncwarp();
		}
	}
}

// ---------------------------------------------------------------- stage 3a: clipping, one THREAD per overlapping item
// The faces that get clipped are small polygons even for the large hulls (triangles, quads), so the warp-per-item clipKernel
// runs its polygon steps on 1-6 lanes.  One thread per item does the same work serially (clipThread); the rare item whose
// faces do not fit its 16-vertex polygon buffers is passed on to clipKernel through a fallback list.
__global__ void __launch_bounds__(128) clipThreadKernel(NpArgs a, const int4* __restrict__ overlapItems, const float4* __restrict__ overlapSep,
														  int4* __restrict__ fallbackItems, float4* __restrict__ fallbackSep)
{
	const int numOverlaps = (int)a.ctr[CTR_OVERLAPS];
	const int lane = threadIdx.x & 31;
	for (int base = blockIdx.x * blockDim.x; base < numOverlaps; base += gridDim.x * blockDim.x)
	{
		const int s = base + (int)threadIdx.x;
		bool fallback = false;
		int4 it = make_int4(0, 0, 0, 0);
		float4 sep = mk4(0, 0, 0);
		if (s < numOverlaps)
		{
			it = overlapItems[s];
			sep = overlapSep[s];
			const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
			Side A, B;
			if (resolveSide(a, bodyA, it.y, A) && resolveSide(a, bodyB, it.z, B))
			{
				float4 posA = A.pos, posB = B.pos;
				posA.w = 0.f;
				posB.w = 0.f;
				const HullRef hA = loadHull(a.convex, A.shape), hB = loadHull(a.convex, B.shape);
				fallback = clipThread(a, it, bodyA, bodyB, A, B, hA, hB, posA, posB, A.orn, B.orn, mk4(sep.x, sep.y, sep.z)) != 0;
			}
		}
		__syncwarp();
		const unsigned int m = __ballot_sync(FULL, fallback);
		if (m)
		{
			unsigned int slot = 0;
			if (lane == 0) slot = atomicAdd(&a.ctr[CTR_CLIP_FALLBACK], (unsigned int)__popc(m));
			slot = __shfl_sync(FULL, slot, 0) + __popc(m & ((1u << lane) - 1u));
			if (fallback)
			{
				fallbackItems[slot] = it;
				fallbackSep[slot] = sep;
			}
		}
	}
}

// ---------------------------------------------------------------- stage 3: clipping + reduction + append
__global__ void __launch_bounds__(NP_THREADS) clipKernel(NpArgs a, const int4* __restrict__ overlapItems, const float4* __restrict__ overlapSep, int countIndex)
{
	__shared__ float4 bufAll[NP_WARPS][2][MAX_POLY];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	float4* bufA = bufAll[warp][0];
	float4* bufB = bufAll[warp][1];
	const int numOverlaps = (int)a.ctr[countIndex];
	for (;;)
	{
		int base = 0;
		if (lane == 0) base = (int)atomicAdd(&a.ctr[CTR_CURSOR_CLIP], (unsigned int)WORK_CHUNK);
		base = __shfl_sync(FULL, base, 0);
		if (base >= numOverlaps) break;
		const int end = base + WORK_CHUNK < numOverlaps ? base + WORK_CHUNK : numOverlaps;
		for (int s = base; s < end; s++)
		{
			const int4 it = overlapItems[s];
			const float4 sep = overlapSep[s];
			const int bodyA = a.pairs[it.x].x, bodyB = a.pairs[it.x].y;
			Side A, B;
			if (resolveSide(a, bodyA, it.y, A) && resolveSide(a, bodyB, it.z, B))
				clipWarp(a, it.x, bodyA, bodyB, A.shape, B.shape, it.y, it.z, A.pos, A.orn, B.pos, B.orn, A.invMass, B.invMass, mk4(sep.x, sep.y, sep.z), bufA, bufB, lane);
			__syncwarp();
		}
	}
}

// ---------------------------------------------------------------- primitives: plane x convex, plane x compound
// One THREAD per broadphase pair that involves a plane.  Restates the reference's host twins
// computeContactPlaneConvex / computeContactPlaneCompound (b3ConvexHullContact.cpp:1272-1395,
// 2182-2320; device version primitiveContacts.cl:584-726) operation by operation, including the
// b3Transform inverse/compose arithmetic (b3Transform.h:90-93,183-197; b3Matrix3x3.h operator*).
B3_D Mat3 matTranspose(const Mat3& m)
{
	Mat3 t;
	t.r0 = mk4(m.r0.x, m.r1.x, m.r2.x);
	t.r1 = mk4(m.r0.y, m.r1.y, m.r2.y);
	t.r2 = mk4(m.r0.z, m.r1.z, m.r2.z);
	return t;
}
// b3Matrix3x3 operator*(m1, m2): row i = (m2.tdotx(m1[i]), m2.tdoty(m1[i]), m2.tdotz(m1[i]))
B3_D Mat3 matMul(const Mat3& m1, const Mat3& m2)
{
	Mat3 r;
	r.r0 = mk4(m2.r0.x * m1.r0.x + m2.r1.x * m1.r0.y + m2.r2.x * m1.r0.z, m2.r0.y * m1.r0.x + m2.r1.y * m1.r0.y + m2.r2.y * m1.r0.z,
			   m2.r0.z * m1.r0.x + m2.r1.z * m1.r0.y + m2.r2.z * m1.r0.z);
	r.r1 = mk4(m2.r0.x * m1.r1.x + m2.r1.x * m1.r1.y + m2.r2.x * m1.r1.z, m2.r0.y * m1.r1.x + m2.r1.y * m1.r1.y + m2.r2.y * m1.r1.z,
			   m2.r0.z * m1.r1.x + m2.r1.z * m1.r1.y + m2.r2.z * m1.r1.z);
	r.r2 = mk4(m2.r0.x * m1.r2.x + m2.r1.x * m1.r2.y + m2.r2.x * m1.r2.z, m2.r0.y * m1.r2.x + m2.r1.y * m1.r2.y + m2.r2.y * m1.r2.z,
			   m2.r0.z * m1.r2.x + m2.r1.z * m1.r2.y + m2.r2.z * m1.r2.z);
	return r;
}

constexpr int MAX_PLANE_CONVEX_POINTS = 64;

B3_D void planeConvexThread(const NpArgs& a, int pairIndex, int planeBody, int convexBody, int child, const Side& B)
{
	const float4 posA = a.pose[2 * planeBody], ornA = a.pose[2 * planeBody + 1];
	const float invMassA = posA.w;
	const int cA = a.coll[planeBody];
	const float4 planeEq = __ldg(reinterpret_cast<const float4*>(&a.faces[__ldg(&a.collidables[cA].shapeIndex)].plane));
	const float4 planeNormal = mk4(planeEq.x, planeEq.y, planeEq.z);
	const float4 planeNormalWorld = quatRotate(ornA, planeNormal);
	const float planeConstant = planeEq.w;
	const HullRef hB = loadHull(a.convex, B.shape);
	const Mat3 Mb = matFromQuat(B.orn), Ma = matFromQuat(ornA);
	const Mat3 MbT = matTranspose(Mb), MaT = matTranspose(Ma);
	// planeInConvex.getBasis() = convexWorldTransform.inverse().basis * planeTransform.basis
	const Mat3 pic = matMul(MbT, Ma);
	const float4 planeNormalInConvex = matMulVec(pic, neg3(planeNormal));
	// planeTransform.inverse(): basis MaT, origin MaT * (-posA)
	const float4 invOriginA = matMulVec(MaT, neg3(mk4(posA.x, posA.y, posA.z)));
	float maxDot = -1e30f;
	float4 pts[MAX_PLANE_CONVEX_POINTS];
	int numPoints = 0;
	for (int i = 0; i < hB.numVertices; i++)
	{
		const float4 vtx = __ldg(&a.vertices[hB.vertexOffset + i]);
		const float curDot = dot3(vtx, planeNormalInConvex);
		if (curDot > maxDot)
		{
			maxDot = curDot;
			if (numPoints == MAX_PLANE_CONVEX_POINTS) numPoints--;  // make sure the deepest point is always included
		}
		if (numPoints < MAX_PLANE_CONVEX_POINTS)
		{
			const float4 r = matMulVec(Mb, vtx);
			float4 vtxWorld = mk4(r.x + B.pos.x, r.y + B.pos.y, r.z + B.pos.z);
			const float4 q = matMulVec(MaT, vtxWorld);
			const float4 vtxInPlane = mk4(q.x + invOriginA.x, q.y + invOriginA.y, q.z + invOriginA.z);
			const float dist = dot3(planeNormal, vtxInPlane) - planeConstant;
			if (dist < 0.f)
			{
				vtxWorld.w = dist;
				pts[numPoints++] = vtxWorld;
			}
		}
	}
	int idx[4] = {0, 1, 2, 3};
	int numReduced = numPoints;
	if (numPoints > 4)
	{
		// extractManifoldSequentialGlobal (b3ConvexHullContact.cpp:418-501) == b3ReduceContacts
		const int nP = numPoints > 64 ? 64 : numPoints;
		float4 center = mk4(0, 0, 0);
		for (int i = 0; i < nP; i++) center = add3(center, pts[i]);
		center = scale3(center, 1.0f / (float)nP);
		const float4 aVector = sub3(pts[0], center);
		float4 u = cross3(planeNormalInConvex, aVector);
		float4 v = cross3(planeNormalInConvex, u);
		u = normalized3(u);
		v = normalized3(v);
		float minW = FLT_MAX;
		int minIndex = -1;
		float m0 = FLT_MIN, m1 = FLT_MIN, m2 = FLT_MIN, m3 = FLT_MIN;
		for (int ie = 0; ie < nP; ie++)
		{
			if (pts[ie].w < minW)
			{
				minW = pts[ie].w;
				minIndex = ie;
			}
			const float4 r = sub3(pts[ie], center);
			float f = dot3(u, r);
			if (f < m0)
			{
				m0 = f;
				idx[0] = ie;
			}
			f = dot3(neg3(u), r);
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
			f = dot3(neg3(v), r);
			if (f < m3)
			{
				m3 = f;
				idx[3] = ie;
			}
		}
		if (idx[0] != minIndex && idx[1] != minIndex && idx[2] != minIndex && idx[3] != minIndex) idx[0] = minIndex;
		numReduced = 4;
	}
	if (numReduced <= 0) return;
	const unsigned int slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
	if (slot >= (unsigned int)a.maxContacts) return;
	b3b200_contact4* c = &a.contacts[slot];
	float4* cw = reinterpret_cast<float4*>(c);
	for (int i = 0; i < 4; i++) cw[i] = i < numReduced ? pts[idx[i]] : mk4(0, 0, 0, 0);
	cw[4] = mk4(-planeNormalWorld.x, -planeNormalWorld.y, -planeNormalWorld.z, (float)numReduced);
	int4 t;
	t.x = (int)(0u | (45874u << 16));  // setFrictionCoeff(0.7) -> (unsigned short)(0.7 * 0xffff)
	t.y = pairIndex;                   // m_batchIdx = pairIndex
	t.z = invMassA == 0.f ? -planeBody : planeBody;
	t.w = B.invMass == 0.f ? -convexBody : convexBody;
	reinterpret_cast<int4*>(c)[5] = t;
	int4 t2;
	t2.x = -1;
	t2.y = child;  // child shape of the compound this contact belongs to (-1 for a plain hull)
	t2.z = 0;
	t2.w = 0;
	reinterpret_cast<int4*>(c)[6] = t2;
	a.pairsOut[pairIndex].z = (int)slot;
}

// one-point contact of the sphere paths (m_childIndexA/B = -1)
B3_D void appendOnePoint(const NpArgs& a, int pairIndex, int bodyA, int bodyB, const float4& normalOnB, const float4& pointWithDepth, int childB = -1)
{
	const unsigned int slot = atomicAdd(&a.ctr[CTR_CONTACTS], 1u);
	if (slot >= (unsigned int)a.maxContacts) return;
	b3b200_contact4* c = &a.contacts[slot];
	float4* cw = reinterpret_cast<float4*>(c);
	cw[0] = pointWithDepth;
	cw[1] = cw[2] = cw[3] = mk4(0, 0, 0, 0);
	cw[4] = mk4(normalOnB.x, normalOnB.y, normalOnB.z, 1.f);
	int4 t;
	t.x = (int)(0u | (45874u << 16));
	t.y = pairIndex;
	t.z = a.pose[2 * bodyA].w == 0.f ? -bodyA : bodyA;
	t.w = a.pose[2 * bodyB].w == 0.f ? -bodyB : bodyB;
	reinterpret_cast<int4*>(c)[5] = t;
	reinterpret_cast<int4*>(c)[6] = make_int4(-1, childB, 0, 0);
	a.pairsOut[pairIndex].z = (int)slot;
}

// computeContactSphereConvex, host twin (b3ConvexHullContact.cpp:2323-2470; signedDistanceFromPointToPlane :342-349,
// IsPointInPolygon :362-416), operation by operation.  A = sphere, B = convex hull.
B3_D void sphereConvexThread(const NpArgs& a, int pairIndex, int sphereBody, int convexBody, int child)
{
	const float radius = __ldg(&a.collidables[a.coll[sphereBody]].radius);
	const float4 spherePos1 = a.pose[2 * sphereBody];
	Side side;
	if (!resolveSide(a, convexBody, child, side)) return;  // child >= 0: that child shape of a compound
	const float4 pos = side.pos, quat = side.orn;
	const Mat3 basis = matFromQuat(quat), inv = matTranspose(basis);
	const float4 invOrigin = matMulVec(inv, neg3(mk4(pos.x, pos.y, pos.z)));
	const float4 spherePos = add3(matMulVec(inv, spherePos1), invOrigin);
	const HullRef h = loadHull(a.convex, side.shape);
	float4 closestPnt = mk4(0, 0, 0), localHitNormal = mk4(0, 0, 0);
	float minDist = -1000000.f;
	bool bCollide = true;
	for (int f = 0; f < h.numFaces; f++)
	{
		const b3b200_face* face = &a.faces[h.faceOffset + f];
		const float4 pl = __ldg(reinterpret_cast<const float4*>(&face->plane));
		const float4 n = mk4(pl.x, pl.y, pl.z);
		float dist = dot3(n, spherePos) + pl.w;
		const float4 pntReturn = sub3(spherePos, scale3(n, dist));
		if (dist > radius)
		{
			bCollide = false;
			break;
		}
		if (dist > 0)
		{
			bool inPoly = true;
			float4 outP = mk4(0, 0, 0);
			const int numIdx = __ldg(&face->numIndices), idxOff = __ldg(&face->indexOffset);
			if (numIdx < 2)
				inPoly = false;
			else
			{
				float4 b = __ldg(&a.vertices[h.vertexOffset + __ldg(&a.indices[idxOff + numIdx - 1])]);
				for (int i = 0; i != numIdx; ++i)
				{
					const float4 av = b;
					b = __ldg(&a.vertices[h.vertexOffset + __ldg(&a.indices[idxOff + i])]);
					const float4 ab = sub3(b, av), ap = sub3(spherePos, av);
					const float4 v = cross3(ab, n);
					if (dot3(ap, v) > 0.f)
					{
						const float ab_m2 = dot3(ab, ab);
						const float rt = ab_m2 != 0.f ? dot3(ab, ap) / ab_m2 : 0.f;
						if (rt <= 0.f)
							outP = av;
						else if (rt >= 1.f)
							outP = b;
						else
						{
							const float s = 1.f - rt;
							outP = mk4(s * av.x + rt * b.x, s * av.y + rt * b.y, s * av.z + rt * b.z);
						}
						inPoly = false;
						break;
					}
				}
			}
			if (inPoly)
			{
				if (dist > minDist)
				{
					minDist = dist;
					closestPnt = pntReturn;
					localHitNormal = n;
				}
			}
			else
			{
				const float4 tmp = sub3(spherePos, outP);
				const float l2 = dot3(tmp, tmp);
				if (l2 < radius * radius)
				{
					dist = sqrtf(l2);
					if (dist > minDist)
					{
						minDist = dist;
						closestPnt = outP;
						localHitNormal = scale3(tmp, 1.0f / dist);
					}
				}
				else
				{
					bCollide = false;
					break;
				}
			}
		}
		else if (dist > minDist)
		{
			minDist = dist;
			closestPnt = pntReturn;
			localHitNormal = n;
		}
	}
	if (bCollide && minDist > -10000)
	{
		const float4 normalOnSurfaceB1 = matMulVec(basis, localHitNormal);
		float4 pOnB1 = add3(matMulVec(basis, closestPnt), mk4(pos.x, pos.y, pos.z));
		const float actualDepth = minDist - radius;
		if (actualDepth < 0)
		{
			pOnB1.w = actualDepth;
			appendOnePoint(a, pairIndex, sphereBody, convexBody, normalOnSurfaceB1, pOnB1, child);
		}
	}
}

// computeContactPlaneSphere (kernels/primitiveContacts.cl:728-790).  The reference has no host twin of it.
B3_D void planeSphereThread(const NpArgs& a, int pairIndex, int planeBody, int sphereBody)
{
	const float4 planeEq = __ldg(reinterpret_cast<const float4*>(&a.faces[__ldg(&a.collidables[a.coll[planeBody]].shapeIndex)].plane));
	const float radius = __ldg(&a.collidables[a.coll[sphereBody]].radius);
	float4 posA = a.pose[2 * planeBody], posB = a.pose[2 * sphereBody];
	const float4 ornA = a.pose[2 * planeBody + 1], ornB = a.pose[2 * sphereBody + 1];
	posA.w = 0.f;
	posB.w = 0.f;
	const float4 planeNormal = mk4(planeEq.x, planeEq.y, planeEq.z);
	const float planeConstant = planeEq.w;
	const float4 invOrnA = quatInverse(ornA), invPosA = quatRotate(invOrnA, neg3(posA));
	const float4 cipOrn = quatMul(invOrnA, ornB), cipPos = add3(quatRotate(invOrnA, posB), invPosA);
	const float4 picOrn = quatMul(quatInverse(ornB), ornA);
	const float4 vtx = scale3(quatRotate(picOrn, neg3(planeNormal)), radius);
	const float4 vtxInPlane = add3(quatRotate(cipOrn, vtx), cipPos);
	const float distance = dot3(planeNormal, vtxInPlane) - planeConstant;
	if (distance < 0.f)
	{
		const float4 projected = sub3(vtxInPlane, scale3(planeNormal, distance));
		const float4 world = add3(quatRotate(ornA, projected), posA);
		const float4 normalOnSurfaceB = quatRotate(ornA, planeNormal);
		float4 pOnB = add3(world, scale3(normalOnSurfaceB, distance));
		pOnB.w = distance;
		appendOnePoint(a, pairIndex, planeBody, sphereBody, neg3(normalOnSurfaceB), pOnB);
	}
}

// sphere x sphere (kernels/primitiveContacts.cl:926-972).  The reference has no host twin of it.
B3_D void sphereSphereThread(const NpArgs& a, int pairIndex, int bodyA, int bodyB)
{
	const float radiusA = __ldg(&a.collidables[a.coll[bodyA]].radius), radiusB = __ldg(&a.collidables[a.coll[bodyB]].radius);
	const float4 posA = a.pose[2 * bodyA], posB = a.pose[2 * bodyB];
	const float4 diff = sub3(posA, posB);
	const float len = sqrtf(dot3(diff, diff));
	if (len <= (radiusA + radiusB))
	{
		const float dist = len - (radiusA + radiusB);
		float4 normalOnSurfaceB = mk4(1.f, 0.f, 0.f);
		if (len > 0.00001f) normalOnSurfaceB = mk4(diff.x / len, diff.y / len, diff.z / len);
		float4 contactPosB = add3(mk4(posB.x, posB.y, posB.z), scale3(normalOnSurfaceB, radiusB));
		contactPosB.w = dist;
		appendOnePoint(a, pairIndex, bodyA, bodyB, normalOnSurfaceB, contactPosB);
	}
}

__global__ void __launch_bounds__(128) npPrimitiveKernel(NpArgs a)
{
	const int numPairs = (int)a.ctr[CTR_PAIRS];
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < numPairs; p += gridDim.x * blockDim.x)
	{
		int bodyA = a.pairs[p].x, bodyB = a.pairs[p].y;
		int cA = a.coll[bodyA], cB = a.coll[bodyB];
		if (cA < 0 || cB < 0) continue;
		int typeA = __ldg(&a.collidables[cA].shapeType), typeB = __ldg(&a.collidables[cB].shapeType);
		if (typeA != B3B200_SHAPE_PLANE && typeB != B3B200_SHAPE_PLANE && (typeA == B3B200_SHAPE_SPHERE || typeB == B3B200_SHAPE_SPHERE))
		{
			if (typeA == B3B200_SHAPE_SPHERE && typeB == B3B200_SHAPE_SPHERE)
				sphereSphereThread(a, p, bodyA, bodyB);
			else if (typeA == B3B200_SHAPE_SPHERE && typeB == B3B200_SHAPE_CONVEX_HULL)
				sphereConvexThread(a, p, bodyA, bodyB, -1);
			else if (typeA == B3B200_SHAPE_CONVEX_HULL && typeB == B3B200_SHAPE_SPHERE)
				sphereConvexThread(a, p, bodyB, bodyA, -1);
			else if ((typeA == B3B200_SHAPE_SPHERE && typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS) ||
					 (typeA == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS && typeB == B3B200_SHAPE_SPHERE))
			{
				// processCompoundPairsPrimitivesKernel (kernels/primitiveContacts.cl:975-1097): every child against the sphere
				const bool compIsB = typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS;
				const int sphereBody = compIsB ? bodyA : bodyB, compBody = compIsB ? bodyB : bodyA, cC = compIsB ? cB : cA;
				const int first = __ldg(&a.collidables[cC].shapeIndex), n = __ldg(&a.collidables[cC].numChildShapes);
				for (int cI = 0; cI < n; cI++) sphereConvexThread(a, p, sphereBody, compBody, first + cI);
			}
			continue;
		}
		if (typeB == B3B200_SHAPE_PLANE && typeA != B3B200_SHAPE_PLANE)
		{
			// the plane is always passed first (b3ConvexHullContact.cpp:2668-2690)
			int t = bodyA;
			bodyA = bodyB;
			bodyB = t;
			t = cA;
			cA = cB;
			cB = t;
			t = typeA;
			typeA = typeB;
			typeB = t;
		}
		if (typeA != B3B200_SHAPE_PLANE) continue;
		if (typeB == B3B200_SHAPE_CONVEX_HULL)
		{
			Side B;
			if (resolveSide(a, bodyB, -1, B)) planeConvexThread(a, p, bodyA, bodyB, -1, B);
		}
		else if (typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS)
		{
			const int first = __ldg(&a.collidables[cB].shapeIndex), n = __ldg(&a.collidables[cB].numChildShapes);
			for (int cI = 0; cI < n; cI++)
			{
				Side B;
				if (resolveSide(a, bodyB, first + cI, B)) planeConvexThread(a, p, bodyA, bodyB, first + cI, B);
			}
		}
		else if (typeB == B3B200_SHAPE_SPHERE)
			planeSphereThread(a, p, bodyA, bodyB);
	}
}

__global__ void clampContactsKernel(unsigned int* ctr, int maxContacts, int maxWorkItems)
{
	if (ctr[CTR_CONTACTS] > (unsigned int)maxContacts)
	{
		ctr[CTR_CONTACTS] = (unsigned int)maxContacts;
		ctr[CTR_OVERFLOW] |= OVF_CONTACTS;
	}
	// work items (child pairs of compounds + surviving pairs) beyond the queue capacity were dropped: say so
	if (ctr[CTR_COMPOUND_PAIRS] > (unsigned int)maxWorkItems || ctr[CTR_SURVIVORS] > (unsigned int)maxWorkItems ||
		ctr[CTR_SMALL_ITEMS] + ctr[CTR_SMALL_ITEMS_BACK] > (unsigned int)maxWorkItems)
		ctr[CTR_OVERFLOW] |= OVF_COMPOUND;
}

int launchNarrowphase(World* w)
{
	cudaStream_t s = w->stream;
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_CONTACTS], 0, sizeof(unsigned int), s));
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_SURVIVORS], 0, 2 * sizeof(unsigned int), s));  // + CTR_OVERLAPS
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_COMPOUND_PAIRS], 0, sizeof(unsigned int), s));
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_SMALL_ITEMS], 0, 3 * sizeof(unsigned int), s));  // + CTR_SMALL_ITEMS_BACK, CTR_CLIP_FALLBACK
	B3_CUDA_CHECK(cudaMemsetAsync(&w->dCounters.ptr[CTR_CURSOR_SAT], 0, 4 * sizeof(unsigned int), s));  // + CLIP, CONCAVE cursors, CTR_MESH_PAIRS
	NpArgs a;
	a.pairs = w->bp.pairs.ptr;
	a.pairsOut = w->bp.pairs.ptr;
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
	a.childSpheres = w->dChildSpheres.ptr;
	a.contacts = w->dContacts.ptr;
	a.maxContacts = w->cfg.maxContactCapacity;
	a.maxWorkItems = (int)w->dSurvivors.cap;
	a.clipMin = w->clipMinDist;
	a.clipMax = w->clipMaxDist;
	// Branches (they share no buffer; contacts are appended through one atomic counter):
	//   P  npPrimitiveKernel                       (pairs only)
	//   C  trimesh chain                           (after npCullKernel listed the trimesh pairs)
	//   S  smallPairKernel                         (after both cull kernels filled the small-item list)
	//   M  satKernel -> clipKernel on `s`          (after npChildCullKernel released the raw queue = overlap list)
	const bool fork = w->npOverlap;
	cudaStream_t sP = fork ? w->npStream[0] : s, sC = fork ? w->npStream[1] : s, sS = fork ? w->npStream[2] : s;
	if (fork)
	{
		B3_CUDA_CHECK(cudaEventRecord(w->evNpFork[0], s));  // counters cleared, pairs final
		B3_CUDA_CHECK(cudaStreamWaitEvent(sP, w->evNpFork[0], 0));
	}
	if (w->hasPlanes)
	{
		npPrimitiveKernel<<<w->smCount * 8, 128, 0, sP>>>(a);
		B3_LAUNCH_CHECK();
	}
	// the overlap list is not live yet: it doubles as the raw child-item queue of compound pairs
	// (the survivor list of the trimesh path is not live yet either: it doubles as the list of trimesh pairs)
	npCullKernel<<<w->smCount * 8, CULL_THREADS, 0, s>>>(a, w->dSurvivors.ptr, w->dOverlapPairs.ptr, w->hasConcave ? reinterpret_cast<int*>(w->dConcaveSurvivors.ptr) : nullptr,
														 w->hasConcave ? (int)(w->dConcaveSurvivors.cap * 4) : 0, w->dSmallItems.ptr);
	B3_LAUNCH_CHECK();
	if (w->hasConcave && fork)
	{
		B3_CUDA_CHECK(cudaEventRecord(w->evNpFork[1], s));
		B3_CUDA_CHECK(cudaStreamWaitEvent(sC, w->evNpFork[1], 0));
		B3_TRY(launchConcave(w, sC));
	}
	if (!w->childShapes.empty())
	{
		npChildCullKernel<<<w->smCount * 8, CULL_THREADS, 0, s>>>(a, w->dOverlapPairs.ptr, w->dSurvivors.ptr, w->dSmallItems.ptr);
		B3_LAUNCH_CHECK();
	}
	if (fork)
	{
		B3_CUDA_CHECK(cudaEventRecord(w->evNpFork[0], s));
		B3_CUDA_CHECK(cudaStreamWaitEvent(sS, w->evNpFork[0], 0));
	}
	smallPairKernel<<<w->smCount * 16, 128, 0, sS>>>(a, w->dSmallItems.ptr);
	B3_LAUNCH_CHECK();
	if (w->timing) B3_CUDA_CHECK(cudaEventRecord(w->evSat[0], s));  // stage_timings()[7] = this kernel (alone only without the forks)
	satKernel<<<w->smCount * 12, NP_THREADS, 0, s>>>(a, w->dSurvivors.ptr, w->dOverlapPairs.ptr, w->dOverlapSep.ptr);
	B3_LAUNCH_CHECK();
	if (w->timing) B3_CUDA_CHECK(cudaEventRecord(w->evSat[1], s));
	if (!fork)
	{
		// thread-per-item clip; what does not fit its buffers comes back in a fallback list.  The lists reuse buffers that are
		// dead by now on this stream: the SAT work items and the small-pair items (not with forked branches: the small-pair
		// kernel may still be reading its list)
		clipThreadKernel<<<w->smCount * 16, 128, 0, s>>>(a, w->dOverlapPairs.ptr, w->dOverlapSep.ptr, w->dSurvivors.ptr, reinterpret_cast<float4*>(w->dSmallItems.ptr));
		B3_LAUNCH_CHECK();
		clipKernel<<<w->smCount * 8, NP_THREADS, 0, s>>>(a, w->dSurvivors.ptr, reinterpret_cast<const float4*>(w->dSmallItems.ptr), (int)CTR_CLIP_FALLBACK);
	}
	else
		clipKernel<<<w->smCount * 8, NP_THREADS, 0, s>>>(a, w->dOverlapPairs.ptr, w->dOverlapSep.ptr, (int)CTR_OVERLAPS);
	B3_LAUNCH_CHECK();
	if (w->hasConcave && !fork) B3_TRY(launchConcave(w, s));
	if (fork)
	{
		cudaStream_t side[3] = {sP, sC, sS};
		for (int i = 0; i < 3; i++)
		{
			B3_CUDA_CHECK(cudaEventRecord(w->evNpJoin[i], side[i]));
			B3_CUDA_CHECK(cudaStreamWaitEvent(s, w->evNpJoin[i], 0));
		}
	}
	clampContactsKernel<<<1, 1, 0, s>>>(w->dCounters.ptr, w->cfg.maxContactCapacity, a.maxWorkItems);
	B3_LAUNCH_CHECK();
	return 0;
}

}  // namespace b3b200
