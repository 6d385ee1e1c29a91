// mpr.cu -- the Minkowski-portal-refinement stage of the reference's GPU narrowphase.
//
// Replaces mprPenetrationKernel (kernels/mpr.cl:14-89), which b3ConvexHullContact.cpp:2817-2850 runs over all convex x convex
// pairs when useMprGpu is set (:19): per pair it runs b3MprPenetration (shared/b3MprPenetration.h:825-888 = portal discovery
// :345-476, refinement :478-518, penetration info :728-823) and, when the hulls penetrate, appends a ONE-point contact
// (position = barycentric point of the portal, normal = -direction, depth = -distance) and records the direction as the
// pair's separating-normal seed.  The pipeline of this build computes its contacts with the SAT / clipping path the
// reference's CPU headers define (narrowphase.cu), so this stage is offered as what it is in the reference: a stand-alone
// kernel over a pair list.  The reference's only golden vectors for the narrowphase are four serialized launches of exactly
// this kernel (data/unittest_data.zip, expected contact totals 0 / 1 / 46 / 98,
// test/OpenCL/AllBullet3Kernels/testExecuteBullet3NarrowphaseKernels.cpp:397-413): tests/golden/mpr_*.npz replays them.
//
// One thread per pair; the arithmetic follows the header's C++ path statement by statement (b3Vector3 operators touch x, y, z
// only; b3QuatRotate = b3Quaternion path, b3TransformPoint = b3Transform path, common.cuh), compiled without FMA contraction,
// so that results equal the header compiled by g++ bit for bit.
#include "internal.h"
#include <algorithm>

namespace b3b200
{
namespace
{
constexpr float MPR_TOLERANCE = 1E-6f;
constexpr int MPR_MAX_ITERATIONS = 1000;

struct Support
{
	float4 v, v1, v2;  // point of the Minkowski difference, of A, of B
};
struct Simplex
{
	Support ps[4];
	int last;
};

struct MprArgs
{
	const b3b200_rigid_body* bodies;
	const b3b200_collidable* collidables;
	const b3b200_convex_polyhedron* convex;
	const float4* vertices;
};

B3_D bool isZero(float v) { return fabsf(v) < FLT_EPSILON; }
// b3MprEq (:96-115)
B3_D bool eq(float a_, float b_)
{
	const float ab = fabsf(a_ - b_);
	if (fabsf(ab) < FLT_EPSILON) return true;
	const float a = fabsf(a_), b = fabsf(b_);
	return b > a ? ab < FLT_EPSILON * b : ab < FLT_EPSILON * a;
}
B3_D bool vecEq(const float4& a, const float4& b) { return eq(a.x, b.x) && eq(a.y, b.y) && eq(a.z, b.z); }
B3_D void scaleIn(float4& d, float k)
{
	d.x *= k;
	d.y *= k;
	d.z *= k;
}
B3_D void addIn(float4& v, const float4& w)
{
	v.x += w.x;
	v.y += w.y;
	v.z += w.z;
}
B3_D void normalizeIn(float4& d)
{
	const float k = 1.f / sqrtf(dot3(d, d));
	scaleIn(d, k);
}

// b3MprConvexSupport (:135-177): world direction -> hull frame, b3MaxDot over the vertices (first strict maximum), back to world
B3_D float4 convexSupport(const MprArgs& a, int body, const float4& dirW)
{
	const float4* bp = reinterpret_cast<const float4*>(&a.bodies[body]);
	const float4 pos = bp[0], orn = bp[1];
	const float4 localDir = quatRotate(quatInverse(orn), mk4(dirW.x, dirW.y, dirW.z, 0.f));
	const int col = a.bodies[body].collidableIdx;
	const b3b200_convex_polyhedron* hull = &a.convex[a.collidables[col].shapeIndex];
	float4 p = mk4(0, 0, 0, 0);
	const int nV = hull->numVertices;
	if (0 < nV)
	{
		const float4* v = a.vertices + hull->vertexOffset;
		float maxDot = -INFINITY;
		int idx = -1;
		for (int i = 0; i < nV; i++)
		{
			const float d = dot3(v[i], localDir);
			if (d > maxDot)
			{
				maxDot = d;
				idx = i;
			}
		}
		if (idx < 0) idx = 0;
		p = v[idx];
	}
	return transformPoint(p, pos, orn);
}
// b3MprSupport (:179-192)
B3_D void support(const MprArgs& a, int bodyA, int bodyB, const float4& dir, Support& s)
{
	s.v1 = convexSupport(a, bodyA, dir);
	s.v2 = convexSupport(a, bodyB, mk4(dir.x * -1.f, dir.y * -1.f, dir.z * -1.f, 0.f));
	s.v = sub3(s.v1, s.v2);
}
// b3PortalDir (:251-262)
B3_D float4 portalDir(const Simplex& p)
{
	float4 dir = cross3(sub3(p.ps[2].v, p.ps[1].v), sub3(p.ps[3].v, p.ps[1].v));
	normalizeIn(dir);
	return dir;
}
// portalEncapsulesOrigin (:264-270) / portalCanEncapsuleOrigin (:298-305)
B3_D bool dotNonNegative(const float4& a, const float4& b)
{
	const float d = dot3(a, b);
	return isZero(d) || d > 0.f;
}
// portalReachTolerance (:272-296)
B3_D bool reachTolerance(const Simplex& p, const Support& v4, const float4& dir)
{
	const float dv1 = dot3(p.ps[1].v, dir), dv2 = dot3(p.ps[2].v, dir), dv3 = dot3(p.ps[3].v, dir), dv4 = dot3(v4.v, dir);
	float dot1 = dv4 - dv1;
	const float dot2 = dv4 - dv2, dot3_ = dv4 - dv3;
	dot1 = dot1 < dot2 ? dot1 : dot2;
	dot1 = dot1 < dot3_ ? dot1 : dot3_;
	return eq(dot1, MPR_TOLERANCE) || dot1 < MPR_TOLERANCE;
}
// b3ExpandPortal (:307-343)
B3_D void expandPortal(Simplex& p, const Support& v4)
{
	const float4 v4v0 = cross3(v4.v, p.ps[0].v);
	float d = dot3(p.ps[1].v, v4v0);
	if (d > 0.f)
	{
		d = dot3(p.ps[2].v, v4v0);
		p.ps[d > 0.f ? 1 : 3] = v4;
	}
	else
	{
		d = dot3(p.ps[3].v, v4v0);
		p.ps[d > 0.f ? 2 : 1] = v4;
	}
}

// b3DiscoverPortal (:345-476)
B3_D int discoverPortal(const MprArgs& a, int bodyA, int bodyB, Simplex& p)
{
	const float4 zero = mk4(0, 0, 0, 0);
	// b3FindOrigin: vertex 0 = centre of the portal
	p.ps[0].v1 = reinterpret_cast<const float4*>(&a.bodies[bodyA])[0];
	p.ps[0].v2 = reinterpret_cast<const float4*>(&a.bodies[bodyB])[0];
	p.ps[0].v = sub3(p.ps[0].v1, p.ps[0].v2);
	p.last = 0;
	if (vecEq(p.ps[0].v, zero)) addIn(p.ps[0].v, mk4(FLT_EPSILON * 10.f, 0.f, 0.f, 0.f));
	// vertex 1 = support in the direction of the origin
	float4 dir = p.ps[0].v;
	scaleIn(dir, -1.f);
	normalizeIn(dir);
	support(a, bodyA, bodyB, dir, p.ps[1]);
	p.last = 1;
	float d = dot3(p.ps[1].v, dir);
	if (isZero(d) || d < 0.f) return -1;
	// vertex 2
	dir = cross3(p.ps[0].v, p.ps[1].v);
	if (isZero(dot3(dir, dir))) return vecEq(p.ps[1].v, zero) ? 1 : 2;  // origin on v1 | on the v0-v1 segment
	normalizeIn(dir);
	support(a, bodyA, bodyB, dir, p.ps[2]);
	d = dot3(p.ps[2].v, dir);
	if (isZero(d) || d < 0.f) return -1;
	p.last = 2;
	// vertex 3 direction
	float4 va = sub3(p.ps[1].v, p.ps[0].v), vb = sub3(p.ps[2].v, p.ps[0].v);
	dir = cross3(va, vb);
	normalizeIn(dir);
	// portal faces oriented "outside" the origin
	d = dot3(dir, p.ps[0].v);
	if (d > 0.f)
	{
		const Support t = p.ps[1];
		p.ps[1] = p.ps[2];
		p.ps[2] = t;
		scaleIn(dir, -1.f);
	}
	while (p.last + 1 < 4)
	{
		support(a, bodyA, bodyB, dir, p.ps[3]);
		d = dot3(p.ps[3].v, dir);
		if (isZero(d) || d < 0.f) return -1;
		bool cont = false;
		// origin outside (v1, v0, v3): v2 := v3
		va = cross3(p.ps[1].v, p.ps[3].v);
		d = dot3(va, p.ps[0].v);
		if (d < 0.f && !isZero(d))
		{
			p.ps[2] = p.ps[3];
			cont = true;
		}
		if (!cont)
		{
			// origin outside (v3, v0, v2): v1 := v3
			va = cross3(p.ps[3].v, p.ps[2].v);
			d = dot3(va, p.ps[0].v);
			if (d < 0.f && !isZero(d))
			{
				p.ps[1] = p.ps[3];
				cont = true;
			}
		}
		if (cont)
		{
			va = sub3(p.ps[1].v, p.ps[0].v);
			vb = sub3(p.ps[2].v, p.ps[0].v);
			dir = cross3(va, vb);
			normalizeIn(dir);
		}
		else
			p.last = 3;
	}
	return 0;
}

// b3RefinePortal (:478-518)
B3_D int refinePortal(const MprArgs& a, int bodyA, int bodyB, Simplex& p)
{
	for (int i = 0; i < MPR_MAX_ITERATIONS; i++)
	{
		const float4 dir = portalDir(p);
		if (dotNonNegative(dir, p.ps[1].v)) return 0;  // the origin is inside the portal
		Support v4;
		support(a, bodyA, bodyB, dir, v4);
		if (!dotNonNegative(v4.v, dir) || reachTolerance(p, v4, dir)) return -1;
		expandPortal(p, v4);
	}
	return -1;
}

// b3FindPos (:520-588): barycentric coordinates of the origin in the portal's tetrahedron
B3_D float4 findPos(const Simplex& p)
{
	const float4 dir = portalDir(p);
	float b[4];
	b[0] = dot3(cross3(p.ps[1].v, p.ps[2].v), p.ps[3].v);
	b[1] = dot3(cross3(p.ps[3].v, p.ps[2].v), p.ps[0].v);
	b[2] = dot3(cross3(p.ps[0].v, p.ps[1].v), p.ps[3].v);
	b[3] = dot3(cross3(p.ps[2].v, p.ps[1].v), p.ps[0].v);
	float sum = b[0] + b[1] + b[2] + b[3];
	if (isZero(sum) || sum < 0.f)
	{
		b[0] = 0.f;
		b[1] = dot3(cross3(p.ps[2].v, p.ps[3].v), dir);
		b[2] = dot3(cross3(p.ps[3].v, p.ps[1].v), dir);
		b[3] = dot3(cross3(p.ps[1].v, p.ps[2].v), dir);
		sum = b[1] + b[2] + b[3];
	}
	const float inv = 1.f / sum;
	float4 p1 = mk4(0, 0, 0, 0), p2 = mk4(0, 0, 0, 0);
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		float4 vec = p.ps[i].v1;
		scaleIn(vec, b[i]);
		addIn(p1, vec);
		vec = p.ps[i].v2;
		scaleIn(vec, b[i]);
		addIn(p2, vec);
	}
	scaleIn(p1, inv);
	scaleIn(p2, inv);
	float4 pos = p1;
	addIn(pos, p2);
	scaleIn(pos, 0.5f);
	return pos;
}

B3_D float dist2(const float4& a, const float4& b)
{
	const float4 ab = sub3(a, b);
	return dot3(ab, ab);
}
// _b3MprVec3PointSegmentDist2 (:597-660), always with a witness
B3_D float pointSegmentDist2(const float4& P, const float4& x0, const float4& b, float4& witness)
{
	float4 d = sub3(b, x0);
	const float4 a = sub3(x0, P);
	float t = -1.f * dot3(a, d);
	t /= dot3(d, d);
	if (t < 0.f || isZero(t))
	{
		witness = x0;
		return dist2(x0, P);
	}
	if (t > 1.f || eq(t, 1.f))
	{
		witness = b;
		return dist2(b, P);
	}
	witness = d;
	scaleIn(witness, t);
	addIn(witness, x0);
	return dist2(witness, P);
}
// b3MprVec3PointTriDist2 (:662-741) with a witness
B3_D float pointTriDist2(const float4& P, const float4& x0, const float4& B, const float4& C, float4& witness)
{
	float4 d1 = sub3(B, x0), d2 = sub3(C, x0);
	const float4 a = sub3(x0, P);
	const float v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
	const float s = (q * r - w * p) / (w * v - r * r);
	const float t = (-s * r - q) / w;
	if ((isZero(s) || s > 0.f) && (eq(s, 1.f) || s < 1.f) && (isZero(t) || t > 0.f) && (eq(t, 1.f) || t < 1.f) && (eq(t + s, 1.f) || t + s < 1.f))
	{
		scaleIn(d1, s);
		scaleIn(d2, t);
		witness = x0;
		addIn(witness, d1);
		addIn(witness, d2);
		return dist2(witness, P);
	}
	float dist = pointSegmentDist2(P, x0, B, witness);
	float4 witness2;
	float d2_ = pointSegmentDist2(P, x0, C, witness2);
	if (d2_ < dist)
	{
		dist = d2_;
		witness = witness2;
	}
	d2_ = pointSegmentDist2(P, B, C, witness2);
	if (d2_ < dist)
	{
		dist = d2_;
		witness = witness2;
	}
	return dist;
}

// b3FindPenetr (:743-790)
B3_D void findPenetr(const MprArgs& a, int bodyA, int bodyB, Simplex& p, float& depth, float4& pdir, float4& pos)
{
	const float4 zero = mk4(0, 0, 0, 0);
	unsigned long iterations = 1UL;
	for (int i = 0; i < MPR_MAX_ITERATIONS; i++)
	{
		const float4 dir = portalDir(p);
		Support v4;
		support(a, bodyA, bodyB, dir, v4);
		if (reachTolerance(p, v4, dir) || iterations == (unsigned long)MPR_MAX_ITERATIONS)
		{
			depth = pointTriDist2(zero, p.ps[1].v, p.ps[2].v, p.ps[3].v, pdir);
			depth = sqrtf(depth);
			if (isZero(pdir.x) && isZero(pdir.y) && isZero(pdir.z)) pdir = dir;
			normalizeIn(pdir);
			pos = findPos(p);
			return;
		}
		expandPortal(p, v4);
		iterations++;
	}
}

// b3MprPenetration (:825-888).  Returns 0 (penetrating: depth, dir, pos valid) or -1.
B3_D int mprPenetration(const MprArgs& a, int bodyA, int bodyB, float4& sepAxis, int& hasSepAxis, float& depth, float4& dir, float4& pos)
{
	Simplex portal;
	hasSepAxis = 0;
	depth = 0.f;
	dir = pos = mk4(0, 0, 0, 0);
	int res = discoverPortal(a, bodyA, bodyB, portal);
	switch (res)
	{
		case 0:
			res = refinePortal(a, bodyA, bodyB, portal);
			if (res < 0) return -1;
			findPenetr(a, bodyA, bodyB, portal, depth, dir, pos);
			hasSepAxis = 1;
			sepAxis = mk4(-dir.x, -dir.y, -dir.z, 0.f);
			break;
		case 1:
			// b3FindPenetrTouch (:792-806): touching contact on the portal's v1
			depth = 0.f;
			dir = mk4(0, 0, 0, 0);
			pos = portal.ps[1].v1;
			addIn(pos, portal.ps[1].v2);
			scaleIn(pos, 0.5f);
			break;
		case 2:
			// b3FindPenetrSegment (:808-823): the origin lies on the v0-v1 segment
			pos = portal.ps[1].v1;
			addIn(pos, portal.ps[1].v2);
			scaleIn(pos, 0.5f);
			dir = portal.ps[1].v;
			depth = sqrtf(dot3(dir, dir));
			normalizeIn(dir);
			break;
		default:
			hasSepAxis = 0;
			return -1;
	}
	return 0;
}

// mprPenetrationKernel (kernels/mpr.cl:14-89)
__global__ void __launch_bounds__(128) mprPenetrationKernel(b3b200_int4* __restrict__ pairs, MprArgs a, float4* __restrict__ sepNormals, int* __restrict__ hasSepAxis,
															b3b200_contact4* __restrict__ contactsOut, unsigned int* __restrict__ nContactsOut, int contactCapacity,
															int numPairs, b3b200_mpr_result* __restrict__ results)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numPairs) return;
	const int bodyA = pairs[i].x, bodyB = pairs[i].y;
	if (results)
	{
		results[i].result = -2;  // pair skipped by the kernel's type / mass tests
		results[i].depth = 0.f;
	}
	if (a.bodies[bodyA].invMass == 0 && a.bodies[bodyB].invMass == 0) return;
	const int cA = a.bodies[bodyA].collidableIdx, cB = a.bodies[bodyB].collidableIdx;
	if (a.collidables[cA].shapeType != B3B200_SHAPE_CONVEX_HULL || a.collidables[cB].shapeType != B3B200_SHAPE_CONVEX_HULL) return;
	float depth;
	float4 dir, pos, sep = mk4(0, 0, 0, 0);
	int has;
	const int res = mprPenetration(a, bodyA, bodyB, sep, has, depth, dir, pos);
	hasSepAxis[i] = has;
	if (has) sepNormals[i] = sep;
	if (results)
	{
		results[i].result = res;
		results[i].depth = depth;
		results[i].dir[0] = dir.x;
		results[i].dir[1] = dir.y;
		results[i].dir[2] = dir.z;
		results[i].pos[0] = pos.x;
		results[i].pos[1] = pos.y;
		results[i].pos[2] = pos.z;
	}
	if (res == 0)
	{
		const unsigned int dst = atomicAdd(nContactsOut, 1u);
		if (dst < (unsigned int)contactCapacity)
		{
			pairs[i].z = (int)dst;
			b3b200_contact4* c = &contactsOut[dst];
			float4* cw = reinterpret_cast<float4*>(c);
			cw[0] = mk4(pos.x, pos.y, pos.z, -depth);
			cw[1] = cw[2] = cw[3] = mk4(0, 0, 0, 0);
			cw[4] = mk4(-dir.x, -dir.y, -dir.z, 1.f);  // GET_NPOINTS = 1
			int4 t;
			t.x = (int)(0u | (45874u << 16));  // restitution 0, friction (unsigned short)(0.7 * 0xffff)
			t.y = i;                            // m_batchIdx = pairIndex
			t.z = a.bodies[bodyA].invMass == 0 ? -bodyA : bodyA;
			t.w = a.bodies[bodyB].invMass == 0 ? -bodyB : bodyB;
			reinterpret_cast<int4*>(c)[5] = t;
			reinterpret_cast<int4*>(c)[6] = make_int4(-1, -1, 0, 0);
		}
	}
}
}  // namespace

}  // namespace b3b200

using namespace b3b200;

// The kernel on caller-supplied HOST arrays (the layout of the reference kernel's buffers): copies them to the device, runs one
// thread per pair, copies the outputs back.  *numContacts is the running contact counter (in: contacts already in
// contactsOut, out: after this stage), as in the reference.  results (optional): per-pair return value / depth / direction /
// position of b3MprPenetration (-2 = pair skipped by the kernel's shape-type / mass tests).
extern "C" int b3b200_mpr_penetration(int device, b3b200_int4* pairs, int numPairs, const b3b200_rigid_body* bodies, int numBodies, const b3b200_collidable* collidables,
									  int numCollidables, const b3b200_convex_polyhedron* convex, int numConvex, const b3b200_float4* vertices, int numVertices,
									  b3b200_float4* sepNormals, int* hasSepAxis, b3b200_contact4* contactsOut, int contactCapacity, int* numContacts,
									  b3b200_mpr_result* results)
{
	if (numPairs < 0 || numBodies < 0 || contactCapacity < 0 || !numContacts || (numPairs > 0 && (!pairs || !bodies || !collidables || !convex || !vertices || !sepNormals || !hasSepAxis)))
		return B3B200_ERR_INVALID;
	if (contactCapacity > 0 && !contactsOut) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(device));
	if (numPairs == 0) return 0;
	DevBuf<b3b200_int4> dPairs;
	DevBuf<b3b200_rigid_body> dBodies;
	DevBuf<b3b200_collidable> dColl;
	DevBuf<b3b200_convex_polyhedron> dConvex;
	DevBuf<float4> dVerts, dSep;
	DevBuf<int> dHas;
	DevBuf<b3b200_contact4> dContacts;
	DevBuf<unsigned int> dCount;
	DevBuf<b3b200_mpr_result> dRes;
	B3_TRY(dPairs.reserve(numPairs));
	B3_TRY(dBodies.reserve(std::max(numBodies, 1)));
	B3_TRY(dColl.reserve(std::max(numCollidables, 1)));
	B3_TRY(dConvex.reserve(std::max(numConvex, 1)));
	B3_TRY(dVerts.reserve(std::max(numVertices, 1)));
	B3_TRY(dSep.reserve(numPairs));
	B3_TRY(dHas.reserve(numPairs));
	B3_TRY(dContacts.reserve(std::max(contactCapacity, 1)));
	B3_TRY(dCount.reserve(1));
	if (results) B3_TRY(dRes.reserve(numPairs));
	B3_CUDA_CHECK(cudaMemcpy(dPairs.ptr, pairs, sizeof(b3b200_int4) * numPairs, cudaMemcpyHostToDevice));
	B3_CUDA_CHECK(cudaMemcpy(dBodies.ptr, bodies, sizeof(b3b200_rigid_body) * numBodies, cudaMemcpyHostToDevice));
	B3_CUDA_CHECK(cudaMemcpy(dColl.ptr, collidables, sizeof(b3b200_collidable) * numCollidables, cudaMemcpyHostToDevice));
	B3_CUDA_CHECK(cudaMemcpy(dConvex.ptr, convex, sizeof(b3b200_convex_polyhedron) * numConvex, cudaMemcpyHostToDevice));
	B3_CUDA_CHECK(cudaMemcpy(dVerts.ptr, vertices, sizeof(float4) * numVertices, cudaMemcpyHostToDevice));
	B3_CUDA_CHECK(cudaMemcpy(dSep.ptr, sepNormals, sizeof(float4) * numPairs, cudaMemcpyHostToDevice));
	B3_CUDA_CHECK(cudaMemcpy(dHas.ptr, hasSepAxis, sizeof(int) * numPairs, cudaMemcpyHostToDevice));
	const unsigned int n0 = (unsigned int)*numContacts;
	B3_CUDA_CHECK(cudaMemcpy(dCount.ptr, &n0, sizeof(n0), cudaMemcpyHostToDevice));
	MprArgs a;
	a.bodies = dBodies.ptr;
	a.collidables = dColl.ptr;
	a.convex = dConvex.ptr;
	a.vertices = dVerts.ptr;
	mprPenetrationKernel<<<divUp(numPairs, 128), 128>>>(dPairs.ptr, a, dSep.ptr, dHas.ptr, dContacts.ptr, dCount.ptr, contactCapacity, numPairs, results ? dRes.ptr : nullptr);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaDeviceSynchronize());
	unsigned int n1 = 0;
	B3_CUDA_CHECK(cudaMemcpy(&n1, dCount.ptr, sizeof(n1), cudaMemcpyDeviceToHost));
	B3_CUDA_CHECK(cudaMemcpy(pairs, dPairs.ptr, sizeof(b3b200_int4) * numPairs, cudaMemcpyDeviceToHost));
	B3_CUDA_CHECK(cudaMemcpy(sepNormals, dSep.ptr, sizeof(float4) * numPairs, cudaMemcpyDeviceToHost));
	B3_CUDA_CHECK(cudaMemcpy(hasSepAxis, dHas.ptr, sizeof(int) * numPairs, cudaMemcpyDeviceToHost));
	const unsigned int nNew = n1 < (unsigned int)contactCapacity ? n1 : (unsigned int)contactCapacity;
	if (nNew > n0) B3_CUDA_CHECK(cudaMemcpy(contactsOut + n0, dContacts.ptr + n0, sizeof(b3b200_contact4) * (nNew - n0), cudaMemcpyDeviceToHost));
	if (results) B3_CUDA_CHECK(cudaMemcpy(results, dRes.ptr, sizeof(b3b200_mpr_result) * numPairs, cudaMemcpyDeviceToHost));
	*numContacts = (int)n1;
	return 0;
}
