// oracle.cpp -- CPU restatement of the reference's algorithms for the GPU
// rigid-body step.  TEST INFRASTRUCTURE ONLY: nothing under bullet3_b200/ may
// link, import or call this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
//
// Parity pinning: oracle/ref_shim.cpp compiles the UNMODIFIED reference sources
// (b3CpuNarrowPhase's shared headers, b3IntegrateTransforms.h, b3UpdateAabbs.h,
// b3ConvertConstraint4.h, the brute-force AABB test) into oracle/_ref/libb3ref.so
// and tests/test_oracle_vs_ref.py checks every function here bit-for-bit against
// it on seeded inputs.  The reference's own golden vectors for this path are the
// MPR record/replay counts (not on the default CPU path) -- see DESIGN.md.
//
// Plain scalar C++, FP32, compiled with -ffp-contract=off.  Every function
// cites the reference file:line it restates (paths relative to bullet3/src).
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../include/b3b200_types.h"

namespace
{
struct V3
{
	float x, y, z, w;
};
inline V3 mk(float x, float y, float z, float w = 0.f)
{
	V3 v = {x, y, z, w};
	return v;
}
inline V3 ld(const b3b200_float4& f) { return mk(f.x, f.y, f.z, f.w); }
inline b3b200_float4 st(const V3& v)
{
	b3b200_float4 f = {v.x, v.y, v.z, v.w};
	return f;
}
// Bullet3Common/b3Vector3.h (scalar path)
inline float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 add(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(const V3& a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 neg(const V3& a) { return mk(-a.x, -a.y, -a.z); }
inline V3 normalized(const V3& v) { return mul(v, 1.0f / sqrtf(dot(v, v))); }  // b3Vector3.h:802,904

// Bullet3Common/b3Quaternion.h:724-729, 305-310, 868-879
inline V3 quatMulVec(const V3& q, const V3& w)
{
	return mk(q.w * w.x + q.y * w.z - q.z * w.y, q.w * w.y + q.z * w.x - q.x * w.z, q.w * w.z + q.x * w.y - q.y * w.x,
			  -q.x * w.x - q.y * w.y - q.z * w.z);
}
inline V3 quatMul(const V3& a, const V3& b)
{
	return mk(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
			  a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
inline V3 quatInv(const V3& q) { return mk(-q.x, -q.y, -q.z, q.w); }
inline V3 quatRotate(const V3& q, const V3& v)
{
	V3 t = quatMulVec(q, v);
	V3 r = quatMul(t, quatInv(q));
	return mk(r.x, r.y, r.z);
}
struct M3
{
	V3 r[3];
};
// b3Matrix3x3::setRotation (b3Matrix3x3.h:201-262)
inline M3 matFromQuat(const V3& q)
{
	float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
	float s = 2.0f / d;
	float xs = q.x * s, ys = q.y * s, zs = q.z * s;
	float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
	float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
	float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
	M3 m;
	m.r[0] = mk(1.0f - (yy + zz), xy - wz, xz + wy);
	m.r[1] = mk(xy + wz, 1.0f - (xx + zz), yz - wx);
	m.r[2] = mk(xz - wy, yz + wx, 1.0f - (xx + yy));
	return m;
}
// b3Matrix3x3::getRotation, scalar path (b3Matrix3x3.h:462-490)
inline V3 quatFromMat(const M3& m)
{
	const float el[3][3] = {{m.r[0].x, m.r[0].y, m.r[0].z}, {m.r[1].x, m.r[1].y, m.r[1].z}, {m.r[2].x, m.r[2].y, m.r[2].z}};
	float trace = el[0][0] + el[1][1] + el[2][2];
	float temp[4];
	if (trace > 0.0f)
	{
		float s = sqrtf(trace + 1.0f);
		temp[3] = (s * 0.5f);
		s = 0.5f / s;
		temp[0] = ((el[2][1] - el[1][2]) * s);
		temp[1] = ((el[0][2] - el[2][0]) * s);
		temp[2] = ((el[1][0] - el[0][1]) * s);
	}
	else
	{
		int i = el[0][0] < el[1][1] ? (el[1][1] < el[2][2] ? 2 : 1) : (el[0][0] < el[2][2] ? 2 : 0);
		int j = (i + 1) % 3;
		int k = (i + 2) % 3;
		float s = sqrtf(el[i][i] - el[j][j] - el[k][k] + 1.0f);
		temp[i] = s * 0.5f;
		s = 0.5f / s;
		temp[3] = (el[k][j] - el[j][k]) * s;
		temp[j] = (el[j][i] + el[i][j]) * s;
		temp[k] = (el[k][i] + el[i][k]) * s;
	}
	return mk(temp[0], temp[1], temp[2], temp[3]);
}
inline V3 matMul(const M3& m, const V3& v) { return mk(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)); }
// b3TransformPoint, C++ path (Bullet3Common/shared/b3Quat.h:18-24 -> b3Transform.h:90-93)
inline V3 transformPoint(const V3& p, const V3& t, const V3& q) { return add(matMul(matFromQuat(q), p), t); }

struct Hull
{
	const b3b200_convex_polyhedron* h;
	const b3b200_float4* vertices;
	const b3b200_float4* uniqueEdges;
	const b3b200_face* faces;
	const int* indices;
};

// b3ProjectAxis (Bullet3Collision/NarrowPhaseCollision/shared/b3FindSeparatingAxis.h:4-34)
void projectAxis(const Hull& hull, const V3& pos, const V3& orn, const V3& dir, float& mn, float& mx)
{
	mn = FLT_MAX;
	mx = -FLT_MAX;
	V3 localDir = quatRotate(quatInv(orn), dir);
	float offset = dot(pos, dir);
	for (int i = 0; i < hull.h->numVertices; i++)
	{
		float dp = dot(ld(hull.vertices[hull.h->vertexOffset + i]), localDir);
		if (dp < mn) mn = dp;
		if (dp > mx) mx = dp;
	}
	if (mn > mx) std::swap(mn, mx);
	mn += offset;
	mx += offset;
}
// b3TestSepAxis (:36-55)
bool testSepAxis(const Hull& A, const Hull& B, const V3& posA, const V3& ornA, const V3& posB, const V3& ornB, const V3& axis, float& depth)
{
	float min0, max0, min1, max1;
	projectAxis(A, posA, ornA, axis, min0, max0);
	projectAxis(B, posB, ornB, axis, min1, max1);
	if (max0 < min1 || max1 < min0) return false;
	float d0 = max0 - min1, d1 = max1 - min0;
	depth = d0 < d1 ? d0 : d1;
	return true;
}
inline bool almostZero(const V3& v)
{
	// b3IsAlmostZero (Bullet3Common/shared/b3Float4.h:58-63) -- double literal on purpose
	if (fabsf(v.x) > 1e-6 || fabsf(v.y) > 1e-6 || fabsf(v.z) > 1e-6) return false;
	return true;
}
// b3FindSeparatingAxis (:57-195)
bool findSeparatingAxis(const Hull& A, const Hull& B, V3 posA, const V3& ornA, V3 posB, const V3& ornB, V3& sep)
{
	posA.w = 0.f;
	posB.w = 0.f;
	V3 c0 = transformPoint(ld(A.h->localCenter), posA, ornA);
	V3 c1 = transformPoint(ld(B.h->localCenter), posB, ornB);
	V3 deltaC2 = sub(c0, c1);
	float dmin = FLT_MAX;
	sep = mk(0, 0, 0);
	for (int i = 0; i < A.h->numFaces; i++)
	{
		V3 n = quatRotate(ornA, ld(A.faces[A.h->faceOffset + i].plane));
		if (dot(deltaC2, n) < 0) n = mul(n, -1.f);
		float d;
		if (!testSepAxis(A, B, posA, ornA, posB, ornB, n, d)) return false;
		if (d < dmin)
		{
			dmin = d;
			sep = n;
		}
	}
	for (int i = 0; i < B.h->numFaces; i++)
	{
		V3 n = quatRotate(ornB, ld(B.faces[B.h->faceOffset + i].plane));
		if (dot(deltaC2, n) < 0) n = mul(n, -1.f);
		float d;
		if (!testSepAxis(A, B, posA, ornA, posB, ornB, n, d)) return false;
		if (d < dmin)
		{
			dmin = d;
			sep = n;
		}
	}
	for (int e0 = 0; e0 < A.h->numUniqueEdges; e0++)
	{
		V3 edge0World = quatRotate(ornA, ld(A.uniqueEdges[A.h->uniqueEdgesOffset + e0]));
		for (int e1 = 0; e1 < B.h->numUniqueEdges; e1++)
		{
			V3 edge1World = quatRotate(ornB, ld(B.uniqueEdges[B.h->uniqueEdgesOffset + e1]));
			V3 cr = cross(edge0World, edge1World);
			if (!almostZero(cr))
			{
				cr = normalized(cr);
				if (dot(deltaC2, cr) < 0) cr = mul(cr, -1.f);
				float dist;
				if (!testSepAxis(A, B, posA, ornA, posB, ornB, cr, dist)) return false;
				if (dist < dmin)
				{
					dmin = dist;
					sep = cr;
				}
			}
		}
	}
	if (dot(neg(deltaC2), sep) > 0.0f) sep = neg(sep);
	return true;
}

const int MAX_VERTS = 1024;  // B3_MAX_VERTS (shared/b3ContactConvexConvexSAT.h:8)

inline V3 lerp3(const V3& a, const V3& b, float t) { return mk(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t); }
// b3ClipFace (shared/b3ContactConvexConvexSAT.h:20-68)
int clipFace(const V3* in, int numIn, const V3& n, float eq, V3* out)
{
	int numOut = 0;
	if (numIn < 2) return 0;
	V3 first = in[numIn - 1];
	float ds = dot(n, first) + eq;
	for (int ve = 0; ve < numIn; ve++)
	{
		V3 end = in[ve];
		float de = dot(n, end) + eq;
		if (ds < 0)
		{
			if (de < 0)
				out[numOut++] = end;
			else
				out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
		}
		else if (de < 0)
		{
			out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
			out[numOut++] = end;
		}
		first = end;
		ds = de;
	}
	return numOut;
}
// b3ClipFaceAgainstHull (:70-176)
int clipFaceAgainstHull(const V3& sep, const Hull& A, const V3& posA, const V3& ornA, V3* vertsB1, int numB1, V3* vertsB2, float minDist, float maxDist,
						V3* contactsOut, int contactCapacity)
{
	int numContactsOut = 0;
	V3* pIn = vertsB1;
	V3* pOut = vertsB2;
	int numIn = numB1;
	int closestFaceA = -1;
	{
		float dmin = FLT_MAX;
		for (int f = 0; f < A.h->numFaces; f++)
		{
			const b3b200_float4& pl = A.faces[A.h->faceOffset + f].plane;
			V3 n = quatRotate(ornA, mk(pl.x, pl.y, pl.z));
			float d = dot(n, sep);
			if (d < dmin)
			{
				dmin = d;
				closestFaceA = f;
			}
		}
	}
	if (closestFaceA < 0) return 0;
	const b3b200_face& polyA = A.faces[A.h->faceOffset + closestFaceA];
	int numVerticesA = polyA.numIndices;
	for (int e0 = 0; e0 < numVerticesA; e0++)
	{
		V3 a = ld(A.vertices[A.h->vertexOffset + A.indices[polyA.indexOffset + e0]]);
		V3 b = ld(A.vertices[A.h->vertexOffset + A.indices[polyA.indexOffset + ((e0 + 1) % numVerticesA)]]);
		V3 edge0 = sub(a, b);
		V3 worldEdge0 = quatRotate(ornA, edge0);
		V3 worldPlaneAnormal1 = quatRotate(ornA, mk(polyA.plane.x, polyA.plane.y, polyA.plane.z));
		V3 planeNormalWS = neg(cross(worldEdge0, worldPlaneAnormal1));
		V3 worldA1 = transformPoint(a, posA, ornA);
		float planeEqWS = -dot(worldA1, planeNormalWS);
		int numOut = clipFace(pIn, numIn, planeNormalWS, planeEqWS, pOut);
		std::swap(pIn, pOut);
		numIn = numOut;
	}
	{
		V3 planeNormalWS = quatRotate(ornA, mk(polyA.plane.x, polyA.plane.y, polyA.plane.z));
		float planeEqWS = polyA.plane.w - dot(planeNormalWS, posA);
		for (int i = 0; i < numIn; i++)
		{
			float depth = dot(planeNormalWS, pIn[i]) + planeEqWS;
			if (depth <= minDist) depth = minDist;
			if (numContactsOut < contactCapacity)
			{
				if (depth <= maxDist) contactsOut[numContactsOut++] = mk(pIn[i].x, pIn[i].y, pIn[i].z, depth);
			}
		}
	}
	return numContactsOut;
}
// b3ClipHullAgainstHull (:178-266)
int clipHullAgainstHull(const V3& sep, const Hull& A, const Hull& B, const V3& posA, const V3& ornA, const V3& posB, const V3& ornB, V3* vertsB1,
						V3* vertsB2, float minDist, float maxDist, V3* contactsOut, int contactCapacity)
{
	int closestFaceB = -1;
	float dmax = -FLT_MAX;
	for (int f = 0; f < B.h->numFaces; f++)
	{
		const b3b200_float4& pl = B.faces[B.h->faceOffset + f].plane;
		V3 n = quatRotate(ornB, mk(pl.x, pl.y, pl.z));
		float d = dot(n, sep);
		if (d > dmax)
		{
			dmax = d;
			closestFaceB = f;
		}
	}
	if (closestFaceB < 0) return 0;
	int numB1 = 0;
	const b3b200_face& polyB = B.faces[B.h->faceOffset + closestFaceB];
	for (int e0 = 0; e0 < polyB.numIndices && numB1 < MAX_VERTS; e0++)
	{
		V3 b = ld(B.vertices[B.h->vertexOffset + B.indices[polyB.indexOffset + e0]]);
		vertsB1[numB1++] = transformPoint(b, posB, ornB);
	}
	return clipFaceAgainstHull(sep, A, posA, ornA, vertsB1, numB1, vertsB2, minDist, maxDist, contactsOut, contactCapacity);
}
// b3ReduceContacts (shared/b3ReduceContacts.h:4-87)
int reduceContacts(const V3* p, int nPoints, const V3& nearNormal, int idx[4])
{
	if (nPoints == 0) return 0;
	if (nPoints <= 4) return nPoints;
	if (nPoints > 64) nPoints = 64;
	V3 center = mk(0, 0, 0);
	for (int i = 0; i < nPoints; i++) center = add(center, p[i]);
	center = mul(center, 1.0f / (float)nPoints);
	V3 aVector = sub(p[0], center);
	V3 u = cross(nearNormal, aVector);
	V3 v = cross(nearNormal, u);
	u = normalized(u);
	v = normalized(v);
	float minW = FLT_MAX;
	int minIndex = -1;
	float maxDots[4] = {FLT_MIN, FLT_MIN, FLT_MIN, FLT_MIN};
	for (int ie = 0; ie < nPoints; ie++)
	{
		if (p[ie].w < minW)
		{
			minW = p[ie].w;
			minIndex = ie;
		}
		V3 r = sub(p[ie], center);
		float f = dot(u, r);
		if (f < maxDots[0])
		{
			maxDots[0] = f;
			idx[0] = ie;
		}
		f = dot(neg(u), r);
		if (f < maxDots[1])
		{
			maxDots[1] = f;
			idx[1] = ie;
		}
		f = dot(v, r);
		if (f < maxDots[2])
		{
			maxDots[2] = f;
			idx[2] = ie;
		}
		f = dot(neg(v), r);
		if (f < maxDots[3])
		{
			maxDots[3] = f;
			idx[3] = ie;
		}
	}
	if (idx[0] != minIndex && idx[1] != minIndex && idx[2] != minIndex && idx[3] != minIndex) idx[0] = minIndex;
	return 4;
}

Hull hullOf(int shapeIndex, const b3b200_convex_polyhedron* convex, const b3b200_float4* vertices, const b3b200_float4* uniqueEdges,
			const b3b200_face* faces, const int* indices)
{
	Hull h = {&convex[shapeIndex], vertices, uniqueEdges, faces, indices};
	return h;
}

unsigned int hashContact(int a, int b, int ca, int cb)
{
	unsigned int h = (unsigned int)a * 0x9E3779B1u;
	h ^= (unsigned int)b * 0x85EBCA77u + 0x165667B1u + (h << 6) + (h >> 2);
	h ^= (unsigned int)ca * 0xC2B2AE3Du + (h << 6) + (h >> 2);
	h ^= (unsigned int)cb * 0x27D4EB2Fu + (h << 6) + (h >> 2);
	h ^= h >> 16;
	h *= 0x85EBCA6Bu;
	h ^= h >> 13;
	h *= 0xC2B2AE35u;
	h ^= h >> 16;
	return h;
}

// calcJacCoeff / calcRelVel / b3PlaneSpace1 (Bullet3Dynamics/shared/b3ConvertConstraint4.h:5-60)
float calcJacCoeff(const V3& angular0, const V3& angular1, float invMass0, const M3& I0, float invMass1, const M3& I1)
{
	float jmj0 = invMass0;
	float jmj1 = dot(matMul(I0, angular0), angular0);
	float jmj2 = invMass1;
	float jmj3 = dot(matMul(I1, angular1), angular1);
	return -1.f / (jmj0 + jmj1 + jmj2 + jmj3);
}
float calcRelVel(const V3& l0, const V3& l1, const V3& a0, const V3& a1, const V3& linVel0, const V3& angVel0, const V3& linVel1, const V3& angVel1)
{
	return dot(l0, linVel0) + dot(a0, angVel0) + dot(l1, linVel1) + dot(a1, angVel1);
}
void planeSpace1(const V3& n, V3& p, V3& q)
{
	if (fabsf(n.z) > 0.70710678f)
	{
		float a = n.y * n.y + n.z * n.z;
		float k = 1.f / sqrtf(a);
		p = mk(0, -n.z * k, n.y * k);
		q = mk(a * k, -n.x * p.z, n.x * p.y);
	}
	else
	{
		float a = n.x * n.x + n.y * n.y;
		float k = 1.f / sqrtf(a);
		p = mk(-n.y * k, n.x * k, 0);
		q = mk(-n.z * p.y, n.z * p.x, a * k);
	}
}
M3 ldM(const b3b200_mat3x3& m)
{
	M3 r;
	r.r[0] = ld(m.row[0]);
	r.r[1] = ld(m.row[1]);
	r.r[2] = ld(m.row[2]);
	return r;
}
}  // namespace

extern "C" {

// b3ComputeWorldAabb / b3TransformAabb2
// (Bullet3Collision/NarrowPhaseCollision/shared/b3UpdateAabbs.h:8-33,
//  Bullet3Collision/BroadPhaseCollision/shared/b3Aabb.h:24-43); max.w carries the
// intended "is dynamic" flag (the reference reads it out of bounds, SURVEY B#12).
void orc_update_aabbs(const b3b200_rigid_body* bodies, int n, const b3b200_collidable* collidables, const b3b200_aabb* localAabbs, b3b200_aabb* out)
{
	for (int i = 0; i < n; i++)
	{
		const b3b200_rigid_body& b = bodies[i];
		int c = b.collidableIdx;
		if (c < 0 || collidables[c].shapeIndex < 0) continue;
		const b3b200_aabb& l = localAabbs[c];
		V3 lmn = mk(l.min[0], l.min[1], l.min[2]), lmx = mk(l.max[0], l.max[1], l.max[2]);
		V3 half = mul(sub(lmx, lmn), 0.5f);
		half = add(half, mk(0.f, 0.f, 0.f));
		V3 lc = mul(add(lmx, lmn), 0.5f);
		M3 m = matFromQuat(ld(b.quat));
		M3 a;
		for (int r = 0; r < 3; r++) a.r[r] = mk(fabsf(m.r[r].x), fabsf(m.r[r].y), fabsf(m.r[r].z));
		V3 center = transformPoint(lc, ld(b.pos), ld(b.quat));
		V3 extent = mk(dot(half, a.r[0]), dot(half, a.r[1]), dot(half, a.r[2]));
		V3 mn = sub(center, extent), mx = add(center, extent);
		out[i].min[0] = mn.x;
		out[i].min[1] = mn.y;
		out[i].min[2] = mn.z;
		out[i].minIndices[3] = i;
		out[i].max[0] = mx.x;
		out[i].max[1] = mx.y;
		out[i].max[2] = mx.z;
		out[i].signedMaxIndices[3] = b.invMass == 0.f ? 0 : 1;
	}
}

// b3TestAabbAgainstAabb (shared/b3Aabb.h:45-53)
static bool aabbOverlap(const b3b200_aabb& a, const b3b200_aabb& b)
{
	bool overlap = true;
	overlap = (a.min[0] > b.max[0] || a.max[0] < b.min[0]) ? false : overlap;
	overlap = (a.min[2] > b.max[2] || a.max[2] < b.min[2]) ? false : overlap;
	overlap = (a.min[1] > b.max[1] || a.max[1] < b.min[1]) ? false : overlap;
	return overlap;
}

// b3GpuSapBroadphase::calculateOverlappingPairsHost
// (Bullet3OpenCL/BroadphaseCollision/b3GpuSapBroadphase.cpp:862-981): brute force
// small x small, then small x large; pairs ordered (min handle, max handle).
// Returns the number of overlapping pairs (may exceed maxPairs; only maxPairs are stored).
int orc_brute_force_pairs(const b3b200_aabb* aabbs, const int* smallIdx, int nSmall, const int* largeIdx, int nLarge, b3b200_int4* pairsOut, int maxPairs)
{
	int count = 0;
	for (int i = 0; i < nSmall; i++)
	{
		const b3b200_aabb& ai = aabbs[smallIdx[i]];
		for (int j = i + 1; j < nSmall; j++)
		{
			const b3b200_aabb& aj = aabbs[smallIdx[j]];
			if (aabbOverlap(ai, aj))
			{
				int a = ai.minIndices[3], b = aj.minIndices[3];
				if (count < maxPairs)
				{
					pairsOut[count].x = a <= b ? a : b;
					pairsOut[count].y = a <= b ? b : a;
					pairsOut[count].z = -1;
					pairsOut[count].w = -1;
				}
				count++;
			}
		}
	}
	for (int i = 0; i < nSmall; i++)
	{
		const b3b200_aabb& ai = aabbs[smallIdx[i]];
		for (int j = 0; j < nLarge; j++)
		{
			const b3b200_aabb& aj = aabbs[largeIdx[j]];
			if (aabbOverlap(ai, aj))
			{
				int a = aj.minIndices[3], b = ai.minIndices[3];
				if (count < maxPairs)
				{
					pairsOut[count].x = a <= b ? a : b;
					pairsOut[count].y = a <= b ? b : a;
					pairsOut[count].z = -1;
					pairsOut[count].w = -1;
				}
				count++;
			}
		}
	}
	return count;
}

// Same pair set as orc_brute_force_pairs, found with a sort-and-sweep on x so that
// full-size scenes (256k bodies) finish in seconds.  Checked against the brute
// force version in tests/test_oracle.py.  Output is sorted lexicographically.
int orc_sweep_pairs(const b3b200_aabb* aabbs, const int* smallIdx, int nSmall, const int* largeIdx, int nLarge, b3b200_int4* pairsOut, int maxPairs)
{
	std::vector<int> order(smallIdx, smallIdx + nSmall);
	std::sort(order.begin(), order.end(), [&](int a, int b) { return aabbs[a].min[0] < aabbs[b].min[0]; });
	std::vector<std::pair<int, int> > found;
	for (int i = 0; i < nSmall; i++)
	{
		const b3b200_aabb& ai = aabbs[order[i]];
		for (int j = i + 1; j < nSmall; j++)
		{
			const b3b200_aabb& aj = aabbs[order[j]];
			if (ai.max[0] < aj.min[0]) break;
			if (aabbOverlap(ai, aj))
			{
				int a = ai.minIndices[3], b = aj.minIndices[3];
				found.push_back(std::make_pair(a <= b ? a : b, a <= b ? b : a));
			}
		}
		for (int j = 0; j < nLarge; j++)
		{
			const b3b200_aabb& aj = aabbs[largeIdx[j]];
			if (aabbOverlap(ai, aj))
			{
				int a = aj.minIndices[3], b = ai.minIndices[3];
				found.push_back(std::make_pair(a <= b ? a : b, a <= b ? b : a));
			}
		}
	}
	std::sort(found.begin(), found.end());
	int count = (int)found.size();
	for (int i = 0; i < count && i < maxPairs; i++)
	{
		pairsOut[i].x = found[i].first;
		pairsOut[i].y = found[i].second;
		pairsOut[i].z = -1;
		pairsOut[i].w = -1;
	}
	return count;
}

// integrateSingleTransform (Bullet3Dynamics/shared/b3IntegrateTransforms.h:5-55):
// the GPU-path order (position first, then gravity).
void orc_integrate(b3b200_rigid_body* bodies, int n, float timeStep, float angularDamping, const float* gravity)
{
	for (int i = 0; i < n; i++)
	{
		b3b200_rigid_body& b = bodies[i];
		if (b.invMass != 0.f)
		{
			float THRESH = (0.25f * 3.14159254f);
			b.angVel.x *= angularDamping;
			b.angVel.y *= angularDamping;
			b.angVel.z *= angularDamping;
			V3 angvel = ld(b.angVel);
			float fAngle = sqrtf(dot(angvel, angvel));
			if (fAngle * timeStep > THRESH) fAngle = THRESH / timeStep;
			V3 axis;
			if (fAngle < 0.001f)
				axis = mul(angvel, 0.5f * timeStep - (timeStep * timeStep * timeStep) * 0.020833333333f * fAngle * fAngle);
			else
				axis = mul(angvel, sinf(0.5f * fAngle * timeStep) / fAngle);
			V3 dorn = mk(axis.x, axis.y, axis.z, cosf(fAngle * timeStep * 0.5f));
			V3 orn0 = ld(b.quat);
			V3 pq = quatMul(dorn, orn0);
			float len2 = pq.x * pq.x + pq.y * pq.y + pq.z * pq.z + pq.w * pq.w;
			float s = 1.0f / sqrtf(len2);
			b.quat.x = pq.x * s;
			b.quat.y = pq.y * s;
			b.quat.z = pq.z * s;
			b.quat.w = pq.w * s;
			b.pos.x += b.linVel.x * timeStep;
			b.pos.y += b.linVel.y * timeStep;
			b.pos.z += b.linVel.z * timeStep;
			b.linVel.x += gravity[0] * timeStep;
			b.linVel.y += gravity[1] * timeStep;
			b.linVel.z += gravity[2] * timeStep;
		}
	}
}

// b3ContactConvexConvexSAT + b3ClipHullHullSingle for every convex-convex pair
// (shared/b3ContactConvexConvexSAT.h:270-484; loop of b3CpuNarrowPhase::computeContacts,
// Bullet3Collision/NarrowPhaseCollision/b3CpuNarrowPhase.cpp:53-141).
// minDist/maxDist are the clip window (CPU header: -1, 0; GPU kernels: -1e30, 0.02).
// pairContactIndex[i] = contact index of pair i or -1.  Returns the contact count.
int orc_convex_contacts(const b3b200_int4* pairs, int nPairs, const b3b200_rigid_body* bodies, const b3b200_collidable* collidables,
						const b3b200_convex_polyhedron* convex, const b3b200_float4* vertices, const b3b200_float4* uniqueEdges,
						const b3b200_face* faces, const int* indices, float minDist, float maxDist, b3b200_contact4* out, int maxContacts,
						int* pairContactIndex)
{
	int nContacts = 0;
	std::vector<V3> b1(MAX_VERTS), b2(MAX_VERTS), cont(MAX_VERTS);
	for (int p = 0; p < nPairs; p++)
	{
		if (pairContactIndex) pairContactIndex[p] = -1;
		int bodyA = pairs[p].x, bodyB = pairs[p].y;
		int cA = bodies[bodyA].collidableIdx, cB = bodies[bodyB].collidableIdx;
		if (collidables[cA].shapeType != B3B200_SHAPE_CONVEX_HULL || collidables[cB].shapeType != B3B200_SHAPE_CONVEX_HULL) continue;
		Hull A = hullOf(collidables[cA].shapeIndex, convex, vertices, uniqueEdges, faces, indices);
		Hull B = hullOf(collidables[cB].shapeIndex, convex, vertices, uniqueEdges, faces, indices);
		V3 posA = ld(bodies[bodyA].pos), ornA = ld(bodies[bodyA].quat);
		V3 posB = ld(bodies[bodyB].pos), ornB = ld(bodies[bodyB].quat);
		posA.w = 0.f;
		posB.w = 0.f;
		V3 sep;
		if (!findSeparatingAxis(A, B, posA, ornA, posB, ornB, sep)) continue;
		// b3ClipHullHullSingle round-trips both orientations through b3Transform:
		// trA.setRotation(ornA); trAorn = trA.getRotation()  (shared/b3ContactConvexConvexSAT.h:323-337)
		V3 ornA2 = quatFromMat(matFromQuat(ornA)), ornB2 = quatFromMat(matFromQuat(ornB));
		int numContactsOut = clipHullAgainstHull(sep, A, B, posA, ornA2, posB, ornB2, b1.data(), b2.data(), minDist, maxDist, cont.data(), MAX_VERTS);
		if (numContactsOut <= 0) continue;
		int idx[4] = {0, 1, 2, 3};
		int numPoints = reduceContacts(cont.data(), numContactsOut, sep, idx);
		if (nContacts < maxContacts)
		{
			b3b200_contact4& c = out[nContacts];
			memset(&c, 0, sizeof(c));
			c.batchIdx = 0;
			c.bodyAPtrAndSignBit = (bodies[bodyA].invMass == 0) ? -bodyA : bodyA;
			c.bodyBPtrAndSignBit = (bodies[bodyB].invMass == 0) ? -bodyB : bodyB;
			c.frictionCmp = 45874;
			c.restitutionCmp = 0;
			c.childIndexA = -1;
			c.childIndexB = -1;
			for (int k = 0; k < numPoints; k++) c.worldPosB[k] = st(cont[idx[k]]);
			c.worldNormalOnB = st(mk(sep.x, sep.y, sep.z, (float)numPoints));
			if (pairContactIndex) pairContactIndex[p] = nContacts;
			nContacts++;
		}
	}
	return nContacts;
}

// ---------------------------------------------------------------------------------------------
// General contact loop: convex x convex, compound children (x compound / x convex), plane x convex,
// plane x compound -- the dispatch of the reference's host loop
// (Bullet3OpenCL/NarrowphaseCollision/b3ConvexHullContact.cpp:2655-2724).
namespace
{
struct SideO
{
	int shape;
	V3 pos, orn;
	float invMass;
};
// child transform composition: sat.cl:836-862 / b3ConvexHullContact.cpp:1806-1832
bool resolveSide(const b3b200_rigid_body* bodies, const b3b200_collidable* collidables, const b3b200_child_shape* children, int body, int child, SideO& s)
{
	V3 pos = ld(bodies[body].pos), orn = ld(bodies[body].quat);
	s.invMass = bodies[body].invMass;
	int coll;
	if (child >= 0)
	{
		V3 cp = ld(children[child].childPosition), co = ld(children[child].childOrientation);
		V3 r = quatRotate(orn, cp);
		pos = mk(r.x + pos.x, r.y + pos.y, r.z + pos.z);
		orn = quatMul(orn, co);
		coll = children[child].shapeIndex;
	}
	else
		coll = bodies[body].collidableIdx;
	if (coll < 0 || collidables[coll].shapeType != B3B200_SHAPE_CONVEX_HULL) return false;
	s.shape = collidables[coll].shapeIndex;
	pos.w = 0.f;
	s.pos = pos;
	s.orn = orn;
	return true;
}
M3 transposeM(const M3& m)
{
	M3 t;
	t.r[0] = mk(m.r[0].x, m.r[1].x, m.r[2].x);
	t.r[1] = mk(m.r[0].y, m.r[1].y, m.r[2].y);
	t.r[2] = mk(m.r[0].z, m.r[1].z, m.r[2].z);
	return t;
}
// b3Matrix3x3 operator* (Bullet3Common/b3Matrix3x3.h: tdotx/tdoty/tdotz of m2 with the rows of m1)
M3 mulM(const M3& m1, const M3& m2)
{
	M3 r;
	for (int i = 0; i < 3; i++)
		r.r[i] = mk(m2.r[0].x * m1.r[i].x + m2.r[1].x * m1.r[i].y + m2.r[2].x * m1.r[i].z, m2.r[0].y * m1.r[i].x + m2.r[1].y * m1.r[i].y + m2.r[2].y * m1.r[i].z,
					m2.r[0].z * m1.r[i].x + m2.r[1].z * m1.r[i].y + m2.r[2].z * m1.r[i].z);
	return r;
}
}  // namespace

extern "C" int orc_contacts(const b3b200_int4* pairs, int nPairs, const b3b200_rigid_body* bodies, const b3b200_collidable* collidables,
							const b3b200_convex_polyhedron* convex, const b3b200_float4* vertices, const b3b200_float4* uniqueEdges, const b3b200_face* faces,
							const int* indices, const b3b200_child_shape* children, float minDist, float maxDist, b3b200_contact4* out, int maxContacts)
{
	int nContacts = 0;
	std::vector<V3> b1(MAX_VERTS), b2(MAX_VERTS), cont(MAX_VERTS);
	auto convexPair = [&](int bodyA, int bodyB, int childA, int childB, const SideO& A, const SideO& B) {
		Hull hA = hullOf(A.shape, convex, vertices, uniqueEdges, faces, indices);
		Hull hB = hullOf(B.shape, convex, vertices, uniqueEdges, faces, indices);
		V3 sep;
		if (!findSeparatingAxis(hA, hB, A.pos, A.orn, B.pos, B.orn, sep)) return;
		V3 ornA2 = quatFromMat(matFromQuat(A.orn)), ornB2 = quatFromMat(matFromQuat(B.orn));
		int n = clipHullAgainstHull(sep, hA, hB, A.pos, ornA2, B.pos, ornB2, b1.data(), b2.data(), minDist, maxDist, cont.data(), MAX_VERTS);
		if (n <= 0) return;
		int idx[4] = {0, 1, 2, 3};
		int numPoints = reduceContacts(cont.data(), n, sep, idx);
		if (nContacts >= maxContacts) return;
		b3b200_contact4& c = out[nContacts++];
		memset(&c, 0, sizeof(c));
		c.bodyAPtrAndSignBit = (A.invMass == 0) ? -bodyA : bodyA;
		c.bodyBPtrAndSignBit = (B.invMass == 0) ? -bodyB : bodyB;
		c.frictionCmp = 45874;
		c.childIndexA = childA;
		c.childIndexB = childB;
		for (int k = 0; k < numPoints; k++) c.worldPosB[k] = st(cont[idx[k]]);
		c.worldNormalOnB = st(mk(sep.x, sep.y, sep.z, (float)numPoints));
	};
	// computeContactPlaneConvex (b3ConvexHullContact.cpp:1272-1395) for one hull
	auto planeConvex = [&](int pairIndex, int planeBody, int convexBody, int child, const SideO& B) {
		V3 posA = ld(bodies[planeBody].pos), ornA = ld(bodies[planeBody].quat);
		int cA = bodies[planeBody].collidableIdx;
		V3 planeEq = ld(faces[collidables[cA].shapeIndex].plane);
		V3 planeNormal = mk(planeEq.x, planeEq.y, planeEq.z);
		V3 planeNormalWorld = quatRotate(ornA, planeNormal);
		float planeConstant = planeEq.w;
		const b3b200_convex_polyhedron& hB = convex[B.shape];
		M3 Mb = matFromQuat(B.orn), Ma = matFromQuat(ornA);
		M3 MbT = transposeM(Mb), MaT = transposeM(Ma);
		M3 pic = mulM(MbT, Ma);
		V3 planeNormalInConvex = matMul(pic, neg(planeNormal));
		V3 invOriginA = matMul(MaT, neg(mk(posA.x, posA.y, posA.z)));
		float maxDot = -1e30f;
		V3 pts[64];
		int numPoints = 0;
		for (int i = 0; i < hB.numVertices; i++)
		{
			V3 vtx = ld(vertices[hB.vertexOffset + i]);
			float curDot = dot(vtx, planeNormalInConvex);
			if (curDot > maxDot)
			{
				maxDot = curDot;
				if (numPoints == 64) numPoints--;
			}
			if (numPoints < 64)
			{
				V3 vtxWorld = add(matMul(Mb, vtx), mk(B.pos.x, B.pos.y, B.pos.z));
				V3 vtxInPlane = add(matMul(MaT, vtxWorld), invOriginA);
				float dist = dot(planeNormal, vtxInPlane) - planeConstant;
				if (dist < 0.f)
				{
					vtxWorld.w = dist;
					pts[numPoints++] = vtxWorld;
				}
			}
		}
		int idx[4] = {0, 1, 2, 3};
		int numReduced = numPoints;
		if (numPoints > 4) numReduced = reduceContacts(pts, numPoints, planeNormalInConvex, idx);  // extractManifoldSequentialGlobal :418-501
		if (numReduced <= 0 || nContacts >= maxContacts) return;
		b3b200_contact4& c = out[nContacts++];
		memset(&c, 0, sizeof(c));
		c.worldNormalOnB = st(mk(-planeNormalWorld.x, -planeNormalWorld.y, -planeNormalWorld.z, (float)numReduced));
		c.frictionCmp = 45874;
		c.batchIdx = pairIndex;
		c.bodyAPtrAndSignBit = bodies[planeBody].invMass == 0 ? -planeBody : planeBody;
		c.bodyBPtrAndSignBit = bodies[convexBody].invMass == 0 ? -convexBody : convexBody;
		c.childIndexA = -1;
		c.childIndexB = child;
		for (int i = 0; i < numReduced; i++) c.worldPosB[i] = st(pts[idx[i]]);
	};
	auto onePoint = [&](int pairIndex, int bodyA, int bodyB, const V3& normalOnB, const V3& pointWithDepth) {
		if (nContacts >= maxContacts) return;
		b3b200_contact4& c = out[nContacts++];
		memset(&c, 0, sizeof(c));
		c.worldNormalOnB = st(mk(normalOnB.x, normalOnB.y, normalOnB.z, 1.f));
		c.frictionCmp = 45874;
		c.batchIdx = pairIndex;
		c.bodyAPtrAndSignBit = bodies[bodyA].invMass == 0 ? -bodyA : bodyA;
		c.bodyBPtrAndSignBit = bodies[bodyB].invMass == 0 ? -bodyB : bodyB;
		c.childIndexA = -1;
		c.childIndexB = -1;
		c.worldPosB[0] = st(pointWithDepth);
	};
	// computeContactSphereConvex, host twin (b3ConvexHullContact.cpp:2323-2470; signedDistanceFromPointToPlane :342-349,
	// IsPointInPolygon :362-416).  A = sphere, B = convex hull.
	// child >= 0: the convex is that child shape of a compound (processCompoundPairsPrimitivesKernel,
	// kernels/primitiveContacts.cl:975-1097, composes the child transform and calls the same routine)
	auto sphereConvex = [&](int pairIndex, int sphereBody, int convexBody, int child) {
		const float radius = collidables[bodies[sphereBody].collidableIdx].radius;
		V3 spherePos1 = ld(bodies[sphereBody].pos);
		SideO side;
		if (!resolveSide(bodies, collidables, children, convexBody, child, side)) return;
		V3 pos = side.pos, quat = side.orn;
		M3 basis = matFromQuat(quat), inv = transposeM(basis);
		V3 invOrigin = matMul(inv, neg(mk(pos.x, pos.y, pos.z)));
		V3 spherePos = add(matMul(inv, spherePos1), invOrigin);
		const b3b200_convex_polyhedron& h = convex[side.shape];
		V3 closestPnt = mk(0, 0, 0), localHitNormal = mk(0, 0, 0);
		float minDist = -1000000.f;
		bool bCollide = true;
		for (int f = 0; f < h.numFaces; f++)
		{
			const b3b200_face& face = faces[h.faceOffset + f];
			V3 n = mk(face.plane.x, face.plane.y, face.plane.z);
			float dist = dot(n, spherePos) + face.plane.w;
			V3 pntReturn = sub(spherePos, mul(n, dist));
			if (dist > radius)
			{
				bCollide = false;
				break;
			}
			if (dist > 0)
			{
				// IsPointInPolygon
				bool inPoly = true;
				V3 outP = mk(0, 0, 0);
				if (face.numIndices < 2)
					inPoly = false;
				else
				{
					V3 b = ld(vertices[h.vertexOffset + indices[face.indexOffset + face.numIndices - 1]]);
					for (int i = 0; i != face.numIndices; ++i)
					{
						V3 a = b;
						b = ld(vertices[h.vertexOffset + indices[face.indexOffset + i]]);
						V3 ab = sub(b, a), ap = sub(spherePos, a);
						V3 v = cross(ab, n);
						if (dot(ap, v) > 0.f)
						{
							float ab_m2 = dot(ab, ab);
							float rt = ab_m2 != 0.f ? dot(ab, ap) / ab_m2 : 0.f;
							if (rt <= 0.f)
								outP = a;
							else if (rt >= 1.f)
								outP = b;
							else
							{
								float s = 1.f - rt;
								outP = mk(s * a.x + rt * b.x, s * a.y + rt * b.y, s * a.z + rt * b.z);
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
					V3 tmp = sub(spherePos, outP);
					float l2 = dot(tmp, tmp);
					if (l2 < radius * radius)
					{
						dist = sqrtf(l2);
						if (dist > minDist)
						{
							minDist = dist;
							closestPnt = outP;
							localHitNormal = mul(tmp, 1.0f / dist);
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
			V3 normalOnSurfaceB1 = matMul(basis, localHitNormal);
			V3 pOnB1 = add(matMul(basis, closestPnt), mk(pos.x, pos.y, pos.z));
			float actualDepth = minDist - radius;
			if (actualDepth < 0)
			{
				pOnB1.w = actualDepth;
				onePoint(pairIndex, sphereBody, convexBody, normalOnSurfaceB1, pOnB1);
				if (child >= 0) out[nContacts - 1].childIndexB = child;
			}
		}
	};
	// computeContactPlaneSphere (kernels/primitiveContacts.cl:728-790; device only -- no host twin: parity unpinned)
	auto planeSphere = [&](int pairIndex, int planeBody, int sphereBody) {
		V3 planeEq = ld(faces[collidables[bodies[planeBody].collidableIdx].shapeIndex].plane);
		const float radius = collidables[bodies[sphereBody].collidableIdx].radius;
		V3 posA = ld(bodies[planeBody].pos), ornA = ld(bodies[planeBody].quat);
		V3 posB = ld(bodies[sphereBody].pos), ornB = ld(bodies[sphereBody].quat);
		posA.w = 0.f;
		posB.w = 0.f;
		V3 planeNormal = mk(planeEq.x, planeEq.y, planeEq.z);
		const float planeConstant = planeEq.w;
		// trInverse / trMul
		V3 invOrnA = quatInv(ornA), invPosA = quatRotate(invOrnA, neg(posA));
		V3 cipOrn = quatMul(invOrnA, ornB), cipPos = add(quatRotate(invOrnA, posB), invPosA);
		V3 invOrnB = quatInv(ornB);
		V3 picOrn = quatMul(invOrnB, ornA);
		V3 vtx = mul(quatRotate(picOrn, neg(planeNormal)), radius);
		V3 vtxInPlane = add(quatRotate(cipOrn, vtx), cipPos);
		float distance = dot(planeNormal, vtxInPlane) - planeConstant;
		if (distance < 0.f)
		{
			V3 projected = sub(vtxInPlane, mul(planeNormal, distance));
			V3 world = add(quatRotate(ornA, projected), posA);
			V3 normalOnSurfaceB = quatRotate(ornA, planeNormal);
			V3 pOnB = add(world, mul(normalOnSurfaceB, distance));
			pOnB.w = distance;
			onePoint(pairIndex, planeBody, sphereBody, neg(normalOnSurfaceB), pOnB);
		}
	};
	// sphere x sphere (kernels/primitiveContacts.cl:926-972; device only: parity unpinned)
	auto sphereSphere = [&](int pairIndex, int bodyA, int bodyB) {
		const float radiusA = collidables[bodies[bodyA].collidableIdx].radius, radiusB = collidables[bodies[bodyB].collidableIdx].radius;
		V3 posA = ld(bodies[bodyA].pos), posB = ld(bodies[bodyB].pos);
		V3 diff = sub(posA, posB);
		float len = sqrtf(dot(diff, diff));
		if (len <= (radiusA + radiusB))
		{
			float dist = len - (radiusA + radiusB);
			V3 normalOnSurfaceB = mk(1.f, 0.f, 0.f);
			if (len > 0.00001f) normalOnSurfaceB = mk(diff.x / len, diff.y / len, diff.z / len);
			V3 contactPosB = add(posB, mul(normalOnSurfaceB, radiusB));
			contactPosB.w = dist;
			onePoint(pairIndex, bodyA, bodyB, normalOnSurfaceB, contactPosB);
		}
	};
	for (int p = 0; p < nPairs; p++)
	{
		int bodyA = pairs[p].x, bodyB = pairs[p].y;
		int cA = bodies[bodyA].collidableIdx, cB = bodies[bodyB].collidableIdx;
		int typeA = collidables[cA].shapeType, typeB = collidables[cB].shapeType;
		const bool hullA = typeA == B3B200_SHAPE_CONVEX_HULL, hullB = typeB == B3B200_SHAPE_CONVEX_HULL;
		const bool compA = typeA == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS, compB = typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS;
		if (hullA && hullB)
		{
			SideO A, B;
			if (resolveSide(bodies, collidables, children, bodyA, -1, A) && resolveSide(bodies, collidables, children, bodyB, -1, B)) convexPair(bodyA, bodyB, -1, -1, A, B);
		}
		else if ((hullA || compA) && (hullB || compB))
		{
			if (bodies[bodyA].invMass == 0 && bodies[bodyB].invMass == 0) continue;
			int firstA = compA ? collidables[cA].shapeIndex : -1, nA = compA ? collidables[cA].numChildShapes : 1;
			int firstB = compB ? collidables[cB].shapeIndex : -1, nB = compB ? collidables[cB].numChildShapes : 1;
			for (int i = 0; i < nA; i++)
				for (int j = 0; j < nB; j++)
				{
					SideO A, B;
					int ca = compA ? firstA + i : -1, cb = compB ? firstB + j : -1;
					if (resolveSide(bodies, collidables, children, bodyA, ca, A) && resolveSide(bodies, collidables, children, bodyB, cb, B)) convexPair(bodyA, bodyB, ca, cb, A, B);
				}
		}
		else if (typeA == B3B200_SHAPE_PLANE || typeB == B3B200_SHAPE_PLANE)
		{
			int planeBody = typeA == B3B200_SHAPE_PLANE ? bodyA : bodyB, other = typeA == B3B200_SHAPE_PLANE ? bodyB : bodyA;
			int cO = bodies[other].collidableIdx, typeO = collidables[cO].shapeType;
			if (typeO == B3B200_SHAPE_CONVEX_HULL)
			{
				SideO B;
				if (resolveSide(bodies, collidables, children, other, -1, B)) planeConvex(p, planeBody, other, -1, B);
			}
			else if (typeO == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS)
			{
				// computeContactPlaneCompound (b3ConvexHullContact.cpp:2182-2320)
				for (int c = 0; c < collidables[cO].numChildShapes; c++)
				{
					SideO B;
					int child = collidables[cO].shapeIndex + c;
					if (resolveSide(bodies, collidables, children, other, child, B)) planeConvex(p, planeBody, other, child, B);
				}
			}
			else if (typeO == B3B200_SHAPE_SPHERE)
				planeSphere(p, planeBody, other);
		}
		else if (typeA == B3B200_SHAPE_SPHERE || typeB == B3B200_SHAPE_SPHERE)
		{
			if (typeA == B3B200_SHAPE_SPHERE && typeB == B3B200_SHAPE_SPHERE)
				sphereSphere(p, bodyA, bodyB);
			else if (typeA == B3B200_SHAPE_SPHERE && hullB)
				sphereConvex(p, bodyA, bodyB, -1);
			else if (hullA && typeB == B3B200_SHAPE_SPHERE)
				sphereConvex(p, bodyB, bodyA, -1);
			else if ((typeA == B3B200_SHAPE_SPHERE && compB) || (compA && typeB == B3B200_SHAPE_SPHERE))
			{
				const int sphereBody = compB ? bodyA : bodyB, compBody = compB ? bodyB : bodyA;
				const int cC = bodies[compBody].collidableIdx;
				for (int c = 0; c < collidables[cC].numChildShapes; c++) sphereConvex(p, sphereBody, compBody, collidables[cC].shapeIndex + c);
			}
			// sphere x trimesh: orc_concave_contacts
		}
	}
	return nContacts;
}

// ---------------------------------------------------------------------------------------------
// Concave trimesh x convex / compound child: the host twins the reference runs when
// bvhTraversalKernelGPU / findConcaveSeparatingAxisKernelGPU / reduceConcaveContactsOnGPU are off and
// clipConcaveFacesAndFindContactsCPU is on (b3ConvexHullContact.cpp:3513-3569, 3700-3770, 3850-3885, 3951-4003):
//   candidates  b3BvhTraversal (shared/b3BvhTraversal.h:11-122) followed by the exact triangle-AABB test at the top of
//               b3FindConcaveSeparatingAxisKernel (shared/b3FindConcaveSatAxis.h:606-610): the quantized tree only
//               over-approximates, so the active set is "triangle AABB (mesh-local!) overlaps the body's world AABB"
//   SAT         b3FindConcaveSeparatingAxisKernel (:551-795): 5-face triangle prism, faces of A, faces of B, all edges
//   clip        b3FindClippingFaces (:417-505) + clipFacesAndFindContactsKernel (shared/b3ClipFaces.h:66-169)
//   reduce      b3NewContactReductionKernel (shared/b3NewContactReduction.h:93-173)
// Contacts come out in (pair, triangle, child) order; the reference's order is its tree's.
namespace
{
bool satOneSided(const Hull& A, const Hull& B, V3 posA, const V3& ornA, V3 posB, const V3& ornB, const V3& deltaC2, V3& sep, float& dmin)
{
	// b3FindSeparatingAxis of b3FindConcaveSatAxis.h:59-125: only the face normals of A
	posA.w = 0.f;
	posB.w = 0.f;
	for (int i = 0; i < A.h->numFaces; i++)
	{
		V3 n = quatRotate(ornA, ld(A.faces[A.h->faceOffset + i].plane));
		if (dot(deltaC2, n) < 0) n = mul(n, -1.f);
		float d;
		if (!testSepAxis(A, B, posA, ornA, posB, ornB, n, d)) return false;
		if (d < dmin)
		{
			dmin = d;
			sep = n;
		}
	}
	if (dot(neg(deltaC2), sep) > 0.0f) sep = neg(sep);
	return true;
}
bool satEdgeEdge(const Hull& A, const Hull& B, V3 posA, const V3& ornA, V3 posB, const V3& ornB, const V3& deltaC2, V3& sep, float& dmin)
{
	// b3FindSeparatingAxisEdgeEdge (:292-415) with searchAllEdgeEdge = true
	posA.w = 0.f;
	posB.w = 0.f;
	for (int e0 = 0; e0 < A.h->numUniqueEdges; e0++)
	{
		V3 edge0World = quatRotate(ornA, ld(A.uniqueEdges[A.h->uniqueEdgesOffset + e0]));
		for (int e1 = 0; e1 < B.h->numUniqueEdges; e1++)
		{
			V3 edge1World = quatRotate(ornB, ld(B.uniqueEdges[B.h->uniqueEdgesOffset + e1]));
			V3 cr = cross(edge0World, edge1World);
			if (!almostZero(cr))
			{
				cr = normalized(cr);
				if (dot(deltaC2, cr) < 0) cr = mul(cr, -1.f);
				float dist;
				if (!testSepAxis(A, B, posA, ornA, posB, ornB, cr, dist)) return false;
				if (dist < dmin)
				{
					dmin = dist;
					sep = cr;
				}
			}
		}
	}
	if (dot(neg(deltaC2), sep) > 0.0f) sep = neg(sep);
	return true;
}
// clipFaceGlobal (shared/b3ClipFaces.h:21-64): like b3ClipFace but without the numVertsIn < 2 early-out
int clipFaceGlobal(const V3* in, int numIn, const V3& n, float eq, V3* out)
{
	int numOut = 0;
	if (numIn <= 0) return 0;  // the reference would read pVtxIn[-1]; it never gets here with 0 vertices and a non-empty loop
	V3 first = in[numIn - 1];
	float ds = dot(n, first) + eq;
	for (int ve = 0; ve < numIn; ve++)
	{
		V3 end = in[ve];
		float de = dot(n, end) + eq;
		if (ds < 0)
		{
			if (de < 0)
				out[numOut++] = end;
			else
				out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
		}
		else if (de < 0)
		{
			out[numOut++] = lerp3(first, end, (ds * 1.f / (ds - de)));
			out[numOut++] = end;
		}
		first = end;
		ds = de;
	}
	return numOut;
}
}  // namespace

extern "C" int orc_concave_contacts(const b3b200_int4* pairs, int nPairs, const b3b200_rigid_body* bodies, const b3b200_collidable* collidables,
									const b3b200_convex_polyhedron* convex, const b3b200_float4* vertices, const b3b200_float4* uniqueEdges,
									const b3b200_face* faces, const int* indices, const b3b200_child_shape* children, const b3b200_aabb* aabbs,
									b3b200_contact4* out, int maxContacts, int* numCandidatesOut)
{
	const int CAP = 64;  // vertexFaceCapacity (b3ConvexHullContact.cpp:3481)
	int nContacts = 0, nCand = 0;
	for (int p = 0; p < nPairs; p++)
	{
		int bodyA = pairs[p].x, bodyB = pairs[p].y;
		int cA = bodies[bodyA].collidableIdx, cB = bodies[bodyB].collidableIdx;
		if (bodies[bodyA].invMass == 0 && bodies[bodyB].invMass == 0) continue;
		if (collidables[cA].shapeType != B3B200_SHAPE_CONCAVE_TRIMESH) continue;  // only with the mesh as A (b3BvhTraversal.h:35)
		int typeB = collidables[cB].shapeType;
		if (typeB != B3B200_SHAPE_CONVEX_HULL && typeB != B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS && typeB != B3B200_SHAPE_SPHERE) continue;
		const b3b200_convex_polyhedron& mesh = convex[collidables[cA].shapeIndex];
		int nChildren = typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS ? collidables[cB].numChildShapes : 1;
		for (int f = 0; f < mesh.numFaces; f++)
		{
			const b3b200_face face = faces[mesh.faceOffset + f];
			b3b200_float4 vA[3];
			V3 localCenter = mk(0, 0, 0);
			float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
			for (int i = 0; i < 3; i++)
			{
				vA[i] = vertices[mesh.vertexOffset + indices[face.indexOffset + i]];
				localCenter = add(localCenter, ld(vA[i]));
				const float c[3] = {vA[i].x, vA[i].y, vA[i].z};
				for (int k = 0; k < 3; k++)
				{
					mn[k] = c[k] < mn[k] ? c[k] : mn[k];
					mx[k] = c[k] > mx[k] ? c[k] : mx[k];
				}
			}
			const b3b200_aabb& bb = aabbs[bodyB];
			bool overlap = true;
			for (int k = 0; k < 3; k++)
				if (mn[k] > bb.max[k] || mx[k] < bb.min[k]) overlap = false;
			if (!overlap) continue;
			if (typeB == B3B200_SHAPE_SPHERE)
			{
				// computeContactSphereTriangle (kernels/primitiveContacts.cl:1162-1300, called by
				// findConcaveSphereContactsKernel :1305-1373; device only: parity unpinned).  "A" of the contact is the sphere.
				nCand++;
				const float radius = collidables[cB].radius;
				V3 pos = ld(bodies[bodyA].pos), quat = ld(bodies[bodyA].quat);
				pos.w = 0.f;
				V3 spherePos2 = ld(bodies[bodyB].pos);
				spherePos2.w = 0.f;
				V3 invOrn = quatInv(quat), invPos = quatRotate(invOrn, neg(pos));
				V3 sphereCenter = add(quatRotate(invOrn, spherePos2), invPos);
				V3 v0 = ld(vA[0]), v1 = ld(vA[1]), v2 = ld(vA[2]);
				V3 normal = normalized(cross(sub(v1, v0), sub(v2, v0)));
				float distanceFromPlane = dot(sub(sphereCenter, v0), normal);
				if (distanceFromPlane < 0.f)
				{
					distanceFromPlane *= -1.f;
					normal = mul(normal, -1.f);
				}
				bool hasContact = false;
				V3 contactPoint = mk(0, 0, 0);
				if (distanceFromPlane < radius)
				{
					// pointInTriangle (:1100-1133)
					V3 e1 = sub(v1, v0), e2 = sub(v2, v1), e3 = sub(v0, v2);
					float r1 = dot(cross(e1, normal), sub(sphereCenter, v0));
					float r2 = dot(cross(e2, normal), sub(sphereCenter, v1));
					float r3 = dot(cross(e3, normal), sub(sphereCenter, v2));
					if ((r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0))
					{
						hasContact = true;
						contactPoint = sub(sphereCenter, mul(normal, distanceFromPlane));
					}
					else
					{
						const V3 tv[3] = {v0, v1, v2};
						for (int i = 0; i < 3; i++)
						{
							// segmentSqrDistance (:1136-1159)
							V3 from = tv[i], to = tv[(i + 1) % 3];
							V3 diff = sub(sphereCenter, from), v = sub(to, from);
							float t = dot(v, diff);
							if (t > 0)
							{
								float dotVV = dot(v, v);
								if (t < dotVV)
								{
									t /= dotVV;
									diff = sub(diff, mul(v, t));
								}
								else
								{
									t = 1;
									diff = sub(diff, v);
								}
							}
							else
								t = 0;
							V3 nearest = add(from, mul(v, t));
							if (dot(diff, diff) < radius * radius)
							{
								hasContact = true;
								contactPoint = nearest;
							}
						}
					}
				}
				if (!hasContact) continue;
				V3 contactToCenter = sub(sphereCenter, contactPoint);
				float minDist = sqrtf(dot(contactToCenter, contactToCenter));
				if (!(minDist > FLT_EPSILON)) continue;
				V3 hitNormal = normalized(contactToCenter);
				V3 normalOnSurfaceB1 = quatRotate(quat, neg(hitNormal));
				V3 pOnB1 = add(quatRotate(quat, contactPoint), pos);
				float actualDepth = minDist - radius;
				if (actualDepth <= 0.f && dot(normalOnSurfaceB1, normalOnSurfaceB1) > FLT_EPSILON && nContacts < maxContacts)
				{
					b3b200_contact4& c = out[nContacts++];
					memset(&c, 0, sizeof(c));
					c.frictionCmp = 45874;
					c.batchIdx = f;
					c.bodyAPtrAndSignBit = bodies[bodyB].invMass == 0 ? -bodyB : bodyB;  // the sphere
					c.bodyBPtrAndSignBit = bodies[bodyA].invMass == 0 ? -bodyA : bodyA;  // the mesh
					c.childIndexA = -1;
					c.childIndexB = f;
					pOnB1.w = actualDepth;
					c.worldPosB[0] = st(pOnB1);
					c.worldNormalOnB = st(mk(-normalOnSurfaceB1.x, -normalOnSurfaceB1.y, -normalOnSurfaceB1.z, 1.f));
				}
				continue;
			}
			for (int ch = 0; ch < nChildren; ch++)
			{
				nCand++;
				// the triangle as a 5-face convex (front, back, three edge planes)
				b3b200_convex_polyhedron tri;
				memset(&tri, 0, sizeof(tri));
				tri.numVertices = 3;
				tri.numUniqueEdges = 3;
				tri.numFaces = 5;
				b3b200_float4 eA[3] = {st(sub(ld(vA[1]), ld(vA[0]))), st(sub(ld(vA[2]), ld(vA[1]))), st(sub(ld(vA[0]), ld(vA[2])))};
				V3 normal = mk(face.plane.x, face.plane.y, face.plane.z);
				b3b200_face fA[5];
				int iA[12] = {0, 1, 2, 2, 1, 0, 0, 0, 0, 0, 0, 0};
				memset(fA, 0, sizeof(fA));
				fA[0].plane = st(mk(normal.x, normal.y, normal.z, face.plane.w));
				fA[0].indexOffset = 0;
				fA[0].numIndices = 3;
				fA[1].plane = st(mk(-normal.x, -normal.y, -normal.z, dot(normal, ld(vA[0]))));
				fA[1].indexOffset = 3;
				fA[1].numIndices = 3;
				int cur = 6, prev = 2;
				for (int i = 0; i < 3; i++)
				{
					V3 v0 = ld(vA[i]), v1 = ld(vA[prev]);
					V3 en = normalized(cross(normal, sub(v1, v0)));
					fA[2 + i].plane = st(mk(en.x, en.y, en.z, -dot(en, v0)));
					fA[2 + i].numIndices = 2;
					fA[2 + i].indexOffset = cur;
					iA[cur++] = i;
					iA[cur++] = prev;
					prev = i;
				}
				tri.localCenter = st(mul(localCenter, 1.f / 3.f));
				Hull A = {&tri, vA, eA, fA, iA};
				V3 posA = ld(bodies[bodyA].pos), ornA = ld(bodies[bodyA].quat);
				V3 posB = ld(bodies[bodyB].pos), ornB = ld(bodies[bodyB].quat);
				posA.w = 0.f;
				posB.w = 0.f;
				int shapeIndexB = collidables[cB].shapeIndex;
				if (typeB == B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS)
				{
					int child = collidables[cB].shapeIndex + ch;
					V3 newPos = transformPoint(ld(children[child].childPosition), posB, ornB);
					ornB = quatMul(ornB, ld(children[child].childOrientation));
					posB = newPos;
					shapeIndexB = collidables[children[child].shapeIndex].shapeIndex;
				}
				Hull B = hullOf(shapeIndexB, convex, vertices, uniqueEdges, faces, indices);
				V3 c0 = transformPoint(ld(tri.localCenter), posA, ornA);
				V3 c1 = transformPoint(ld(B.h->localCenter), posB, ornB);
				V3 deltaC2 = sub(c0, c1);
				float dmin = FLT_MAX;
				V3 sep = mk(1, 2, 3, 4);
				if (!satOneSided(A, B, posA, ornA, posB, ornB, deltaC2, sep, dmin)) continue;
				if (!satOneSided(B, A, posB, ornB, posA, ornA, deltaC2, sep, dmin)) continue;
				if (!satEdgeEdge(A, B, posA, ornA, posB, ornB, deltaC2, sep, dmin)) continue;
				// b3FindClippingFaces
				V3 b1[CAP], b2[CAP], a1[CAP];
				int closestFaceB = -1;
				float dmax = -FLT_MAX;
				for (int fb = 0; fb < B.h->numFaces; fb++)
				{
					const b3b200_float4& pl = B.faces[B.h->faceOffset + fb].plane;
					float d = dot(quatRotate(ornB, mk(pl.x, pl.y, pl.z)), sep);
					if (d > dmax)
					{
						dmax = d;
						closestFaceB = fb;
					}
				}
				int numB = 0;
				{
					const b3b200_face& polyB = B.faces[B.h->faceOffset + closestFaceB];
					for (int e0 = 0; e0 < polyB.numIndices && numB < CAP; e0++)
						b1[numB++] = transformPoint(ld(B.vertices[B.h->vertexOffset + B.indices[polyB.indexOffset + e0]]), posB, ornB);
				}
				int closestFaceA = -1;
				V3 worldNormalA = mk(0, 0, 0);
				{
					float dm = FLT_MAX;
					for (int fa = 0; fa < 5; fa++)
					{
						V3 n = quatRotate(ornA, mk(fA[fa].plane.x, fA[fa].plane.y, fA[fa].plane.z));
						float d = dot(n, sep);
						if (d < dm)
						{
							dm = d;
							closestFaceA = fa;
							worldNormalA = n;
						}
					}
				}
				int numA = fA[closestFaceA].numIndices;
				for (int e0 = 0; e0 < numA; e0++) a1[e0] = transformPoint(ld(vA[iA[fA[closestFaceA].indexOffset + e0]]), posA, ornA);
				// clipFacesAndFindContactsKernel
				const float minDist = -1e30f, maxDist = 0.02f;
				V3* pIn = b1;
				V3* pOut = b2;
				int numIn = numB;
				for (int e0 = 0; e0 < numA; e0++)
				{
					V3 aw = a1[e0], bw = a1[(e0 + 1) % numA];
					V3 worldEdge0 = sub(aw, bw);
					V3 planeNormalWS = neg(cross(worldEdge0, worldNormalA));
					float planeEqWS = -dot(aw, planeNormalWS);
					int numOut = clipFaceGlobal(pIn, numIn, planeNormalWS, planeEqWS, pOut);
					std::swap(pIn, pOut);
					numIn = numOut;
				}
				int numLocal = 0;
				{
					float planeEqWS = -dot(worldNormalA, a1[0]);
					for (int i = 0; i < numIn; i++)
					{
						float depth = dot(worldNormalA, pIn[i]) + planeEqWS;
						if (depth <= minDist) depth = minDist;
						if (depth <= maxDist) pOut[numLocal++] = mk(pIn[i].x, pIn[i].y, pIn[i].z, depth);
					}
				}
				if (numLocal <= 0) continue;
				// b3NewContactReductionKernel
				int idx[4] = {0, 1, 2, 3};
				int nReduced = reduceContacts(pOut, numLocal, neg(sep), idx);  // b3ExtractManifoldSequentialGlobal == b3ReduceContacts arithmetic
				if (nContacts >= maxContacts) continue;
				b3b200_contact4& c = out[nContacts++];
				memset(&c, 0, sizeof(c));
				c.frictionCmp = 45874;
				c.batchIdx = f;  // the reference stores the concave-pair index (tree order); the triangle index is the stable equivalent
				c.bodyAPtrAndSignBit = bodies[bodyA].invMass == 0 ? -bodyA : bodyA;
				c.bodyBPtrAndSignBit = bodies[bodyB].invMass == 0 ? -bodyB : bodyB;
				c.childIndexA = -1;
				c.childIndexB = -1;
				for (int k = 0; k < nReduced; k++) c.worldPosB[k] = st(pOut[idx[k]]);
				c.worldNormalOnB = st(mk(sep.x, sep.y, sep.z, (float)nReduced));
			}
		}
	}
	if (numCandidatesOut) *numCandidatesOut = nCand;
	return nContacts;
}

// Graph colouring = sequential first-fit in descending priority order, priority =
// (hashContact(bodyA, bodyB, childA, childB) << 32) | (index + 1).  This is the
// batching rule of the new solver (bullet3_b200/csrc/solver.cu); it plays the role
// of b3GpuPgsContactSolver::sortConstraintByBatch3
// (Bullet3OpenCL/RigidBody/b3GpuPgsContactSolver.cpp:1385-1529), whose invariant
// -- no two constraints of a batch share a dynamic body -- it keeps.
// Returns the number of batches; colours[i] in [0, 128).
int orc_colour_contacts(const b3b200_contact4* contacts, int n, int numBodies, int staticIdx, int* colours)
{
	std::vector<unsigned long long> prio(n);
	std::vector<int> order(n);
	for (int i = 0; i < n; i++)
	{
		int a = abs(contacts[i].bodyAPtrAndSignBit), b = abs(contacts[i].bodyBPtrAndSignBit);
		prio[i] = ((unsigned long long)hashContact(a, b, contacts[i].childIndexA, contacts[i].childIndexB) << 32) | (unsigned long long)(i + 1);
		order[i] = i;
	}
	std::sort(order.begin(), order.end(), [&](int x, int y) { return prio[x] > prio[y]; });
	std::vector<unsigned long long> mask(2 * (size_t)numBodies, 0ull);
	int numBatches = 0;
	for (int k = 0; k < n; k++)
	{
		int i = order[k];
		int as = contacts[i].bodyAPtrAndSignBit, bs = contacts[i].bodyBPtrAndSignBit;
		int a = abs(as), b = abs(bs);
		bool aStatic = as < 0 || as == staticIdx, bStatic = bs < 0 || bs == staticIdx;
		unsigned long long m0 = 0, m1 = 0;
		if (!aStatic)
		{
			m0 |= mask[2 * a];
			m1 |= mask[2 * a + 1];
		}
		if (!bStatic)
		{
			m0 |= mask[2 * b];
			m1 |= mask[2 * b + 1];
		}
		int colour = 127;
		if (~m0)
			colour = __builtin_ctzll(~m0);
		else if (~m1)
			colour = 64 + __builtin_ctzll(~m1);
		unsigned long long bit = 1ull << (colour & 63);
		if (!aStatic) mask[2 * a + (colour >> 6)] |= bit;
		if (!bStatic) mask[2 * b + (colour >> 6)] |= bit;
		colours[i] = colour;
		if (colour + 1 > numBatches) numBatches = colour + 1;
	}
	return numBatches;
}

// setConstraint4 (Bullet3Dynamics/shared/b3ConvertConstraint4.h:62-148) as driven by
// b3Solver::convertToConstraints (Bullet3OpenCL/RigidBody/b3Solver.cpp:889-933):
// rows are built with m_initInvInertia.
void orc_build_constraints(const b3b200_contact4* contacts, int n, const b3b200_rigid_body* bodies, const b3b200_inertia* inertias, float dt,
						   float positionDrift, float positionConstraintCoeff, b3b200_constraint4* out)
{
	for (int g = 0; g < n; g++)
	{
		const b3b200_contact4& src = contacts[g];
		b3b200_constraint4& dst = out[g];
		memset(&dst, 0, sizeof(dst));
		int aIdx = abs(src.bodyAPtrAndSignBit), bIdx = abs(src.bodyBPtrAndSignBit);
		V3 posA = ld(bodies[aIdx].pos), linVelA = ld(bodies[aIdx].linVel), angVelA = ld(bodies[aIdx].angVel);
		V3 posB = ld(bodies[bIdx].pos), linVelB = ld(bodies[bIdx].linVel), angVelB = ld(bodies[bIdx].angVel);
		float invMassA = bodies[aIdx].invMass, invMassB = bodies[bIdx].invMass;
		M3 IA = ldM(inertias[aIdx].initInvInertia), IB = ldM(inertias[bIdx].initInvInertia);
		dst.bodyA = aIdx;
		dst.bodyB = bIdx;
		float dtInv = 1.f / dt;
		V3 n3 = mk(src.worldNormalOnB.x, src.worldNormalOnB.y, src.worldNormalOnB.z);
		float npoints = src.worldNormalOnB.w;
		dst.linear = st(mk(n3.x, n3.y, n3.z, 0.7f));
		for (int ic = 0; ic < 4; ic++)
		{
			V3 r0 = sub(ld(src.worldPosB[ic]), posA);
			V3 r1 = sub(ld(src.worldPosB[ic]), posB);
			if (ic >= npoints)
			{
				dst.jacCoeffInv[ic] = 0.f;
				continue;
			}
			V3 angular0 = cross(r0, n3);
			V3 angular1 = neg(cross(r1, n3));
			dst.jacCoeffInv[ic] = calcJacCoeff(angular0, angular1, invMassA, IA, invMassB, IB);
			float relVelN = calcRelVel(n3, neg(n3), angular0, angular1, linVelA, angVelA, linVelB, angVelB);
			float e = 0.f;
			if (relVelN * relVelN < 0.004f) e = 0.f;
			dst.b[ic] = e * relVelN;
			dst.b[ic] += (src.worldPosB[ic].w + positionDrift) * positionConstraintCoeff * dtInv;
			dst.appliedRambdaDt[ic] = 0.f;
		}
		if (npoints > 0)
		{
			V3 center = mk(0, 0, 0);
			for (int i = 0; i < npoints; i++) center = add(center, ld(src.worldPosB[i]));
			center = mul(center, 1.0f / (float)npoints);
			V3 t[2];
			planeSpace1(n3, t[0], t[1]);
			V3 r0 = sub(center, posA), r1 = sub(center, posB);
			for (int i = 0; i < 2; i++)
			{
				V3 a0 = cross(r0, t[i]), a1 = neg(cross(r1, t[i]));
				dst.fJacCoeffInv[i] = calcJacCoeff(a0, a1, invMassA, IA, invMassB, IB);
				dst.fAppliedRambdaDt[i] = 0.f;
			}
			dst.center = st(center);
		}
		for (int i = 0; i < 4; i++)
			if (i < npoints)
				dst.worldPos[i] = src.worldPosB[i];
			else
				dst.worldPos[i] = st(mk(0, 0, 0, 0));
		dst.batchIdx = src.batchIdx;
	}
}

// solveContact<false> / solveFriction / SolveTask::run, in the order of the
// reference's global-batch mode: all iterations of the normal rows batch by batch,
// then all iterations of the friction rows (Bullet3OpenCL/RigidBody/b3Solver.cpp:187-329,
// 347-409; b3GpuPgsContactSolver.cpp:262-311).  `constraints` must be sorted by batch,
// batchOffsets has numBatches+1 entries.
void orc_solve(b3b200_constraint4* cs, const int* batchOffsets, int numBatches, b3b200_rigid_body* bodies, const b3b200_inertia* inertias, int iterations)
{
	for (int phase = 0; phase < 2; phase++)
		for (int iter = 0; iter < iterations; iter++)
			for (int bt = 0; bt < numBatches; bt++)
				for (int i = batchOffsets[bt]; i < batchOffsets[bt + 1]; i++)
				{
					b3b200_constraint4& c = cs[i];
					int aIdx = (int)c.bodyA, bIdx = (int)c.bodyB;
					b3b200_rigid_body& A = bodies[aIdx];
					b3b200_rigid_body& B = bodies[bIdx];
					V3 posA = ld(A.pos), posB = ld(B.pos);
					V3 linVelA = ld(A.linVel), angVelA = ld(A.angVel), linVelB = ld(B.linVel), angVelB = ld(B.angVel);
					float invMassA = A.invMass, invMassB = B.invMass;
					M3 IA = ldM(inertias[aIdx].invInertiaWorld), IB = ldM(inertias[bIdx].invInertiaWorld);
					V3 lin = ld(c.linear);
					if (phase == 0)
					{
						for (int ic = 0; ic < 4; ic++)
						{
							if (c.jacCoeffInv[ic] == 0.f) continue;
							V3 r0 = sub(ld(c.worldPos[ic]), posA), r1 = sub(ld(c.worldPos[ic]), posB);
							V3 angular0 = cross(r0, lin), angular1 = neg(cross(r1, lin));
							float rambdaDt = calcRelVel(lin, neg(lin), angular0, angular1, linVelA, angVelA, linVelB, angVelB) + c.b[ic];
							rambdaDt *= c.jacCoeffInv[ic];
							float prevSum = c.appliedRambdaDt[ic];
							float updated = prevSum;
							updated += rambdaDt;
							updated = std::max(updated, 0.f);
							updated = std::min(updated, FLT_MAX);
							rambdaDt = updated - prevSum;
							c.appliedRambdaDt[ic] = updated;
							V3 linImp0 = mul(mul(lin, invMassA), rambdaDt);
							V3 linImp1 = mul(mul(neg(lin), invMassB), rambdaDt);
							V3 angImp0 = mul(matMul(IA, angular0), rambdaDt);
							V3 angImp1 = mul(matMul(IB, angular1), rambdaDt);
							linVelA = add(linVelA, linImp0);
							angVelA = add(angVelA, angImp0);
							linVelB = add(linVelB, linImp1);
							angVelB = add(angVelB, angImp1);
						}
					}
					else
					{
						if (c.fJacCoeffInv[0] == 0 && c.fJacCoeffInv[0] == 0) continue;
						float sum = 0;
						for (int j = 0; j < 4; j++) sum += c.appliedRambdaDt[j];
						float maxR = 0.7f * sum, minR = -maxR;
						V3 center = ld(c.center);
						V3 n = neg(lin);
						V3 t[2];
						planeSpace1(n, t[0], t[1]);
						V3 r0 = sub(center, posA), r1 = sub(center, posB);
						for (int k = 0; k < 2; k++)
						{
							V3 angular0 = cross(r0, t[k]), angular1 = neg(cross(r1, t[k]));
							float rambdaDt = calcRelVel(t[k], neg(t[k]), angular0, angular1, linVelA, angVelA, linVelB, angVelB);
							rambdaDt *= c.fJacCoeffInv[k];
							float prevSum = c.fAppliedRambdaDt[k];
							float updated = prevSum;
							updated += rambdaDt;
							updated = std::max(updated, minR);
							updated = std::min(updated, maxR);
							rambdaDt = updated - prevSum;
							c.fAppliedRambdaDt[k] = updated;
							V3 linImp0 = mul(mul(t[k], invMassA), rambdaDt);
							V3 linImp1 = mul(mul(neg(t[k]), invMassB), rambdaDt);
							V3 angImp0 = mul(matMul(IA, angular0), rambdaDt);
							V3 angImp1 = mul(matMul(IB, angular1), rambdaDt);
							linVelA = add(linVelA, linImp0);
							angVelA = add(angVelA, angImp0);
							linVelB = add(linVelB, linImp1);
							angVelB = add(angVelB, angImp1);
						}
						V3 ab = normalized(sub(posB, posA));
						V3 ac = normalized(sub(center, posA));
						if (dot(ab, ac) > 0.95f || (invMassA == 0.f || invMassB == 0.f))
						{
							float angNA = dot(n, angVelA), angNB = dot(n, angVelB);
							angVelA = sub(angVelA, mul(n, angNA * 0.1f));
							angVelB = sub(angVelB, mul(n, angNB * 0.1f));
						}
					}
					if (invMassA != 0.f)
					{
						A.linVel = st(linVelA);
						A.angVel = st(angVelA);
					}
					if (invMassB != 0.f)
					{
						B.linVel = st(linVelB);
						B.angVel = st(angVelB);
					}
				}
}

// Mass-splitting Jacobi contact solver, GPU order: per iteration contacts, average, friction, average;
// then the averaged delta is added to the body velocities.
// (Bullet3OpenCL/RigidBody/b3GpuJacobiContactSolver.cpp:699-869; kernels/solverUtils.cl:
//  CountBodiesKernel :392-414, setConstraint4/ContactToConstraintSplitKernel :836-967,
//  SolveContactJacobiKernel :527-651, AverageVelocitiesKernel :428-456,
//  SolveFrictionJacobiKernel :654-811, UpdateBodyVelocitiesKernel :815-833.)
// Split slots are handed out in contact-index order (the reference's atomic order is arbitrary).
// hostOrder = 0: the order of the GPU path (solveContacts, :788-855): per iteration contacts, average, friction, average.
// hostOrder = 1: the order of the reference's host twin solveGroupHost (:462-697): all contact iterations, then all friction
// iterations -- the variant that is pinned bit for bit against the compiled reference (tests/test_oracle_vs_refcl.py); the two
// differ in nothing but this loop nest.
void orc_jacobi_solve_ordered(const b3b200_contact4* contacts, int n, b3b200_rigid_body* bodies, int numBodies, const b3b200_inertia* inertias, int staticIdx,
							  int iterations, float dt, float positionDrift, float positionConstraintCoeff, int hostOrder);
void orc_jacobi_solve(const b3b200_contact4* contacts, int n, b3b200_rigid_body* bodies, int numBodies, const b3b200_inertia* inertias, int staticIdx,
					  int iterations, float dt, float positionDrift, float positionConstraintCoeff)
{
	orc_jacobi_solve_ordered(contacts, n, bodies, numBodies, inertias, staticIdx, iterations, dt, positionDrift, positionConstraintCoeff, 0);
}
void orc_jacobi_solve_ordered(const b3b200_contact4* contacts, int n, b3b200_rigid_body* bodies, int numBodies, const b3b200_inertia* inertias, int staticIdx,
							  int iterations, float dt, float positionDrift, float positionConstraintCoeff, int hostOrder)
{
	std::vector<unsigned int> bodyCount(numBodies, 0), bodyOffset(numBodies, 0);
	std::vector<int> slotA(n, 0), slotB(n, 0);
	for (int i = 0; i < n; i++)
	{
		int pa = contacts[i].bodyAPtrAndSignBit, pb = contacts[i].bodyBPtrAndSignBit;
		if (!((pa < 0) || (pa == staticIdx))) slotA[i] = (int)bodyCount[abs(pa)]++;
		if (!((pb < 0) || (pb == staticIdx))) slotB[i] = (int)bodyCount[abs(pb)]++;
	}
	unsigned int total = 0;
	for (int i = 0; i < numBodies; i++)
	{
		bodyOffset[i] = total;
		total += bodyCount[i];
	}
	std::vector<b3b200_constraint4> cs(n);
	for (int g = 0; g < n; g++)
	{
		const b3b200_contact4& src = contacts[g];
		b3b200_constraint4& dst = cs[g];
		memset(&dst, 0, sizeof(dst));
		int aIdx = abs(src.bodyAPtrAndSignBit), bIdx = abs(src.bodyBPtrAndSignBit);
		V3 posA = ld(bodies[aIdx].pos), linVelA = ld(bodies[aIdx].linVel), angVelA = ld(bodies[aIdx].angVel);
		V3 posB = ld(bodies[bIdx].pos), linVelB = ld(bodies[bIdx].linVel), angVelB = ld(bodies[bIdx].angVel);
		float invMassA = bodies[aIdx].invMass, invMassB = bodies[bIdx].invMass;
		M3 IA = ldM(inertias[aIdx].invInertiaWorld), IB = ldM(inertias[bIdx].invInertiaWorld);
		float countA = invMassA != 0.f ? (float)bodyCount[aIdx] : 1, countB = invMassB != 0.f ? (float)bodyCount[bIdx] : 1;
		auto jac = [&](const V3& a0, const V3& a1) {
			float jmj0 = invMassA, jmj1 = dot(matMul(IA, a0), a0), jmj2 = invMassB, jmj3 = dot(matMul(IB, a1), a1);
			return -1.f / ((jmj0 + jmj1) * countA + (jmj2 + jmj3) * countB);
		};
		dst.bodyA = aIdx;
		dst.bodyB = bIdx;
		float dtInv = 1.f / dt;
		V3 n3 = mk(src.worldNormalOnB.x, src.worldNormalOnB.y, src.worldNormalOnB.z);
		float npoints = src.worldNormalOnB.w;
		dst.linear = st(mk(n3.x, n3.y, n3.z, 0.7f));
		for (int ic = 0; ic < 4; ic++)
		{
			V3 r0 = sub(ld(src.worldPosB[ic]), posA), r1 = sub(ld(src.worldPosB[ic]), posB);
			if (ic >= npoints) continue;
			V3 a0 = cross(r0, n3), a1 = neg(cross(r1, n3));
			dst.jacCoeffInv[ic] = jac(a0, a1);
			float relVelN = calcRelVel(n3, neg(n3), a0, a1, linVelA, angVelA, linVelB, angVelB);
			float e = 0.f;
			dst.b[ic] = e * relVelN;
			dst.b[ic] += (src.worldPosB[ic].w + positionDrift) * positionConstraintCoeff * dtInv;
		}
		if (npoints > 0)
		{
			V3 center = mk(0, 0, 0);
			for (int i = 0; i < npoints; i++) center = add(center, ld(src.worldPosB[i]));
			center = mul(center, 1.0f / (float)npoints);
			V3 t[2];
			planeSpace1(neg(n3), t[0], t[1]);
			V3 r0 = sub(center, posA), r1 = sub(center, posB);
			for (int i = 0; i < 2; i++) dst.fJacCoeffInv[i] = jac(cross(r0, t[i]), neg(cross(r1, t[i])));
			dst.center = st(center);
		}
		for (int i = 0; i < 4; i++)
			if (i < npoints) dst.worldPos[i] = src.worldPosB[i];
	}
	std::vector<V3> dLin(total + 1, mk(0, 0, 0)), dAng(total + 1, mk(0, 0, 0));
	auto average = [&]() {
		for (int i = 0; i < numBodies; i++)
		{
			if (!bodies[i].invMass) continue;
			int off = (int)bodyOffset[i], count = (int)bodyCount[i];
			float factor = 1.f / ((float)count);
			V3 avL = mk(0, 0, 0), avA = mk(0, 0, 0);
			for (int j = 0; j < count; j++)
			{
				avL = add(avL, mul(dLin[off + j], factor));
				avA = add(avA, mul(dAng[off + j], factor));
			}
			for (int j = 0; j < count; j++)
			{
				dLin[off + j] = avL;
				dAng[off + j] = avA;
			}
		}
	};
	// sweep = (iteration, phase) in the order of the chosen variant
	for (int sweep = 0; sweep < 2 * iterations; sweep++)
	{
		{
			const int phase = hostOrder ? (sweep >= iterations ? 1 : 0) : (sweep & 1);
			for (int i = 0; i < n; i++)
			{
				b3b200_constraint4& c = cs[i];
				int aIdx = (int)c.bodyA, bIdx = (int)c.bodyB;
				V3 posA = ld(bodies[aIdx].pos), linVelA = ld(bodies[aIdx].linVel), angVelA = ld(bodies[aIdx].angVel);
				V3 posB = ld(bodies[bIdx].pos), linVelB = ld(bodies[bIdx].linVel), angVelB = ld(bodies[bIdx].angVel);
				float invMassA = bodies[aIdx].invMass, invMassB = bodies[bIdx].invMass;
				M3 IA = ldM(inertias[aIdx].invInertiaWorld), IB = ldM(inertias[bIdx].invInertiaWorld);
				int splitA = (int)bodyOffset[aIdx] + slotA[i], splitB = (int)bodyOffset[bIdx] + slotB[i];
				V3 dLA = mk(0, 0, 0), dAA = mk(0, 0, 0), dLB = mk(0, 0, 0), dAB = mk(0, 0, 0);
				if (invMassA)
				{
					dLA = dLin[splitA];
					dAA = dAng[splitA];
				}
				if (invMassB)
				{
					dLB = dLin[splitB];
					dAB = dAng[splitB];
				}
				V3 lin = ld(c.linear);
				if (phase == 0)
				{
					for (int ic = 0; ic < 4; ic++)
					{
						if (c.jacCoeffInv[ic] == 0.f) continue;
						V3 r0 = sub(ld(c.worldPos[ic]), posA), r1 = sub(ld(c.worldPos[ic]), posB);
						V3 a0 = cross(r0, lin), a1 = neg(cross(r1, lin));
						float rambdaDt = calcRelVel(lin, neg(lin), a0, a1, add(linVelA, dLA), add(angVelA, dAA), add(linVelB, dLB), add(angVelB, dAB)) + c.b[ic];
						rambdaDt *= c.jacCoeffInv[ic];
						float prevSum = c.appliedRambdaDt[ic];
						float updated = prevSum + rambdaDt;
						updated = std::max(updated, 0.f);
						updated = std::min(updated, FLT_MAX);
						rambdaDt = updated - prevSum;
						c.appliedRambdaDt[ic] = updated;
						if (invMassA)
						{
							dLA = add(dLA, mul(mul(lin, invMassA), rambdaDt));
							dAA = add(dAA, mul(matMul(IA, a0), rambdaDt));
						}
						if (invMassB)
						{
							dLB = add(dLB, mul(mul(neg(lin), invMassB), rambdaDt));
							dAB = add(dAB, mul(matMul(IB, a1), rambdaDt));
						}
					}
				}
				else
				{
					if (c.fJacCoeffInv[0] == 0 && c.fJacCoeffInv[0] == 0) continue;
					float sum = 0;
					for (int j = 0; j < 4; j++) sum += c.appliedRambdaDt[j];
					float maxR = 0.7f * sum, minR = -maxR;
					V3 center = ld(c.center);
					V3 nn = neg(lin);
					V3 t[2];
					planeSpace1(nn, t[0], t[1]);
					V3 r0 = sub(center, posA), r1 = sub(center, posB);
					// The reference's host twin (solveFriction, b3GpuJacobiContactSolver.cpp:239-312) sums velocity + delta ONCE before the two
					// tangent directions and damps with that sum, and only touches the deltas of dynamic bodies; the kernel
					// (solveFrictionConstraint, solverUtils.cl:654-790) re-evaluates the sum per direction and damps with the body's own
					// angular velocity.  hostOrder selects the twin's form -- the only other difference between the two variants.
					const V3 sumLA = add(linVelA, dLA), sumAA = add(angVelA, dAA), sumLB = add(linVelB, dLB), sumAB = add(angVelB, dAB);
					for (int k = 0; k < 2; k++)
					{
						V3 a0 = cross(r0, t[k]), a1 = neg(cross(r1, t[k]));
						float rambdaDt = hostOrder ? calcRelVel(t[k], neg(t[k]), a0, a1, sumLA, sumAA, sumLB, sumAB)
												   : calcRelVel(t[k], neg(t[k]), a0, a1, add(linVelA, dLA), add(angVelA, dAA), add(linVelB, dLB), add(angVelB, dAB));
						rambdaDt *= c.fJacCoeffInv[k];
						float prevSum = c.fAppliedRambdaDt[k];
						float updated = prevSum + rambdaDt;
						updated = std::max(updated, minR);
						updated = std::min(updated, maxR);
						rambdaDt = updated - prevSum;
						c.fAppliedRambdaDt[k] = updated;
						if (!hostOrder || invMassA)
						{
							dLA = add(dLA, mul(mul(t[k], invMassA), rambdaDt));
							dAA = add(dAA, mul(matMul(IA, a0), rambdaDt));
						}
						if (!hostOrder || invMassB)
						{
							dLB = add(dLB, mul(mul(neg(t[k]), invMassB), rambdaDt));
							dAB = add(dAB, mul(matMul(IB, a1), rambdaDt));
						}
					}
					V3 ab = normalized(sub(posB, posA)), ac = normalized(sub(center, posA));
					if (dot(ab, ac) > 0.95f || (invMassA == 0.f || invMassB == 0.f))
					{
						float angNA = dot(nn, hostOrder ? sumAA : angVelA), angNB = dot(nn, hostOrder ? sumAB : angVelB);
						if (!hostOrder || invMassA) dAA = sub(dAA, mul(nn, angNA * 0.1f));
						if (!hostOrder || invMassB) dAB = sub(dAB, mul(nn, angNB * 0.1f));
					}
				}
				if (invMassA)
				{
					dLin[splitA] = dLA;
					dAng[splitA] = dAA;
				}
				if (invMassB)
				{
					dLin[splitB] = dLB;
					dAng[splitB] = dAB;
				}
			}
			average();
		}
	}
	if (iterations > 0)
		for (int i = 0; i < numBodies; i++)
			if (bodies[i].invMass && bodyCount[i])
			{
				int off = (int)bodyOffset[i];
				bodies[i].linVel = st(add(ld(bodies[i].linVel), dLin[off]));
				bodies[i].angVel = st(add(ld(bodies[i].angVel), dAng[off]));
			}
}

// b3RadixSort32CL::executeHost (Bullet3OpenCL/ParallelPrimitives/b3RadixSort32CL.cpp:587-646): stable by key
void orc_radix_sort_kv(b3b200_sort_data* data, int n)
{
	std::stable_sort(data, data + n, [](const b3b200_sort_data& a, const b3b200_sort_data& b) { return a.key < b.key; });
}
// b3PrefixScanCL::executeHost (b3PrefixScanCL.cpp:105-119): exclusive
void orc_prefix_scan(const unsigned int* src, unsigned int* dst, int n, unsigned int* sum)
{
	unsigned int s = 0;
	for (int i = 0; i < n; i++)
	{
		unsigned int t = src[i];
		dst[i] = s;
		s += t;
	}
	if (sum) *sum = s;
}
// b3BoundSearchCL::executeHost COUNT (b3BoundSearchCL.cpp:159-203)
void orc_bound_search_count(const b3b200_sort_data* sorted, int n, unsigned int* counts, int numBuckets)
{
	for (int i = 0; i < numBuckets; i++) counts[i] = 0;
	for (int i = 0; i < n; i++)
		if (sorted[i].key < (unsigned int)numBuckets) counts[sorted[i].key]++;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Joints: b3GpuPgsConstraintSolver::solveJoints, host twins (b3GpuPgsConstraintSolver.cpp: setup :346-541,
// sortConstraintByBatch3 :778-905, iterations :664-735 with resolveSingleConstraintRowGeneric2 :575-604, finish
// :959-1050) + getInfo2Point2Point (b3GpuGenericConstraint.cpp:39-116) and, for the fixed joint, the device-only
// getInfo2FixedOrientation (kernels/jointSolver.cl:546-628).  4 iterations, dt = 1/60, ERP 0.2, CFM 0, damping 1
// (:927-942, b3ContactSolverInfo.h:66-74).  The reference's host twin only works for ONE constraint (its
// batchConstraints[] are never initialised on that path), which is what it is pinned against.
extern "C" void orc_solve_joints(b3b200_rigid_body* bodies, int numBodies, const b3b200_inertia* inertias, b3b200_generic_constraint* cs, int n)
{
	struct Row
	{
		V3 normal, rel1, rel2, angA, angB;
		float jacDiagABInv, rhs, cfm, lower, upper, applied;
		int bodyA, bodyB;
	};
	const float fps = 60.f, erp = 0.2f, cfm = 0.f, damping = 1.0f;
	const int iterations = 4;
	std::vector<V3> dLin(numBodies, mk(0, 0, 0)), dAng(numBodies, mk(0, 0, 0));
	std::vector<std::vector<Row> > rows(n);
	for (int i = 0; i < n; i++)
	{
		const b3b200_generic_constraint& c = cs[i];
		int nr = 0;
		if (c.flags & B3B200_CONSTRAINT_FLAG_ENABLED) nr = c.constraintType == B3B200_CONSTRAINT_P2P ? 3 : (c.constraintType == B3B200_CONSTRAINT_FIXED ? 6 : 0);
		if (!nr) continue;
		const b3b200_rigid_body& rbA = bodies[c.rbA];
		const b3b200_rigid_body& rbB = bodies[c.rbB];
		M3 mA = matFromQuat(ld(rbA.quat)), mB = matFromQuat(ld(rbB.quat));
		V3 a1 = matMul(mA, mk(c.pivotInA.x, c.pivotInA.y, c.pivotInA.z)), a2 = matMul(mB, mk(c.pivotInB.x, c.pivotInB.y, c.pivotInB.z));
		V3 a1n = neg(a1);
		const float k = fps * erp;
		V3 normals[6], rel1[6], rel2[6];
		float rhs[6];
		const V3 J1[3] = {mk(0.f, -a1n.z, a1n.y), mk(a1n.z, 0.f, -a1n.x), mk(-a1n.y, a1n.x, 0.f)};
		const V3 J2[3] = {mk(0.f, -a2.z, a2.y), mk(a2.z, 0.f, -a2.x), mk(-a2.y, a2.x, 0.f)};
		const float pa[3] = {rbA.pos.x, rbA.pos.y, rbA.pos.z}, pb[3] = {rbB.pos.x, rbB.pos.y, rbB.pos.z};
		const float a1v[3] = {a1.x, a1.y, a1.z}, a2v[3] = {a2.x, a2.y, a2.z};
		for (int j = 0; j < 3; j++)
		{
			normals[j] = mk(j == 0 ? 1.f : 0.f, j == 1 ? 1.f : 0.f, j == 2 ? 1.f : 0.f);
			rel1[j] = J1[j];
			rel2[j] = J2[j];
			rhs[j] = k * (a2v[j] + pb[j] - a1v[j] - pa[j]);
		}
		if (nr == 6)
		{
			V3 ornA = ld(rbA.quat), ornB = ld(rbB.quat);
			V3 qrelCur = quatMul(ornA, quatInv(ornB));
			V3 q0 = ld(c.relTargetAB);
			V3 dq = mk(q0.x - qrelCur.x, q0.y - qrelCur.y, q0.z - qrelCur.z, q0.w - qrelCur.w);
			V3 sq = mk(q0.x + qrelCur.x, q0.y + qrelCur.y, q0.z + qrelCur.z, q0.w + qrelCur.w);
			float dd = dq.x * dq.x + dq.y * dq.y + dq.z * dq.z + dq.w * dq.w, ss = sq.x * sq.x + sq.y * sq.y + sq.z * sq.z + sq.w * sq.w;
			V3 orn1 = dd < ss ? qrelCur : mk(-qrelCur.x, -qrelCur.y, -qrelCur.z, -qrelCur.w);
			V3 dorn = quatMul(orn1, quatInv(q0));
			if (dorn.w >= 1.f) dorn.w = 1.f;
			float x = dorn.w;
			if (x < -1.f) x = -1.f;
			if (x > 1.f) x = 1.f;
			float angle = 2.f * acosf(x);
			V3 axis = mk(dorn.x, dorn.y, dorn.z);
			float len = dot(axis, axis);
			if (len < FLT_EPSILON * FLT_EPSILON)
				axis = mk(1.f, 0.f, 0.f);
			else
			{
				float sl = sqrtf(len);
				axis = mk(axis.x / sl, axis.y / sl, axis.z / sl);
			}
			V3 diff = mul(axis, -angle);
			const float dv[3] = {diff.x, diff.y, diff.z};
			for (int j = 0; j < 3; j++)
			{
				normals[3 + j] = mk(0, 0, 0);
				rel1[3 + j] = mk(j == 0 ? 1.f : 0.f, j == 1 ? 1.f : 0.f, j == 2 ? 1.f : 0.f);
				rel2[3 + j] = mk(j == 0 ? -1.f : 0.f, j == 1 ? -1.f : 0.f, j == 2 ? -1.f : 0.f);
				rhs[3 + j] = k * dv[j];
			}
		}
		M3 IA = ldM(inertias[c.rbA].invInertiaWorld), IB = ldM(inertias[c.rbB].invInertiaWorld);
		for (int j = 0; j < nr; j++)
		{
			Row r;
			r.lower = -FLT_MAX;
			r.upper = FLT_MAX;
			if (r.upper >= c.breakingImpulseThreshold) r.upper = c.breakingImpulseThreshold;
			if (r.lower <= -c.breakingImpulseThreshold) r.lower = -c.breakingImpulseThreshold;
			r.normal = normals[j];
			r.rel1 = rel1[j];
			r.rel2 = rel2[j];
			r.angA = matMul(IA, rel1[j]);
			r.angB = matMul(IB, rel2[j]);
			V3 iMJlA = mul(normals[j], rbA.invMass), iMJlB = mul(normals[j], rbB.invMass);
			float sum = dot(iMJlA, normals[j]);
			sum += dot(r.angA, rel1[j]);
			sum += dot(iMJlB, normals[j]);
			sum += dot(r.angB, rel2[j]);
			r.jacDiagABInv = fabsf(sum) > FLT_EPSILON ? 1.f / sum : 0.f;
			float vel1Dotn = dot(normals[j], ld(rbA.linVel)) + dot(rel1[j], ld(rbA.angVel));
			float vel2Dotn = -dot(normals[j], ld(rbB.linVel)) + dot(rel2[j], ld(rbB.angVel));
			float relVel = vel1Dotn + vel2Dotn;
			float velocityError = 0.f - relVel * damping;
			r.rhs = rhs[j] * r.jacDiagABInv + velocityError * r.jacDiagABInv;
			r.cfm = cfm;
			r.applied = 0.f;
			r.bodyA = c.rbA;
			r.bodyB = c.rbB;
			rows[i].push_back(r);
		}
	}
	// sortConstraintByBatch3, simdWidth = n + 1
	std::vector<int> order(n), batchOffset(1, 0);
	for (int i = 0; i < n; i++) order[i] = i;
	{
		std::vector<char> used(numBodies, 0);
		int numValid = 0;
		while (numValid < n)
		{
			std::fill(used.begin(), used.end(), 0);
			for (int i = numValid; i < n; i++)
			{
				const b3b200_generic_constraint& c = cs[order[i]];
				bool aStatic = bodies[c.rbA].invMass == 0.f, bStatic = bodies[c.rbB].invMass == 0.f;
				bool unavailable = !aStatic && used[c.rbA];
				if (!unavailable) unavailable = !bStatic && used[c.rbB];
				if (unavailable) continue;
				if (!aStatic) used[c.rbA] = 1;
				if (!bStatic) used[c.rbB] = 1;
				std::swap(order[i], order[numValid]);
				numValid++;
			}
			batchOffset.push_back(numValid);
		}
	}
	for (int it = 0; it < iterations; it++)
		for (size_t b = 0; b + 1 < batchOffset.size(); b++)
			for (int s = batchOffset[b]; s < batchOffset[b + 1]; s++)
			{
				int ci = order[s];
				if (!(cs[ci].flags & B3B200_CONSTRAINT_FLAG_ENABLED)) continue;
				for (size_t j = 0; j < rows[ci].size(); j++)
				{
					Row& r = rows[ci][j];
					float invMassA = bodies[r.bodyA].invMass, invMassB = bodies[r.bodyB].invMass;
					float deltaImpulse = r.rhs - r.applied * r.cfm;
					float deltaVel1Dotn = dot(r.normal, dLin[r.bodyA]) + dot(r.rel1, dAng[r.bodyA]);
					float deltaVel2Dotn = -dot(r.normal, dLin[r.bodyB]) + dot(r.rel2, dAng[r.bodyB]);
					deltaImpulse -= deltaVel1Dotn * r.jacDiagABInv;
					deltaImpulse -= deltaVel2Dotn * r.jacDiagABInv;
					float sum = r.applied + deltaImpulse;
					if (sum < r.lower)
					{
						deltaImpulse = r.lower - r.applied;
						r.applied = r.lower;
					}
					else if (sum > r.upper)
					{
						deltaImpulse = r.upper - r.applied;
						r.applied = r.upper;
					}
					else
						r.applied = sum;
					dLin[r.bodyA] = add(dLin[r.bodyA], mul(mul(r.normal, invMassA), deltaImpulse));
					dAng[r.bodyA] = add(dAng[r.bodyA], mul(r.angA, deltaImpulse));
					dLin[r.bodyB] = add(dLin[r.bodyB], mul(mul(neg(r.normal), invMassB), deltaImpulse));
					dAng[r.bodyB] = add(dAng[r.bodyB], mul(r.angB, deltaImpulse));
				}
			}
	for (int i = 0; i < n; i++)
		for (size_t j = 0; j < rows[i].size(); j++)
			if (fabsf(rows[i][j].applied) >= cs[i].breakingImpulseThreshold) cs[i].flags = 0;
	for (int i = 0; i < numBodies; i++)
	{
		if (bodies[i].invMass == 0.f) continue;
		bodies[i].linVel.x += dLin[i].x;
		bodies[i].linVel.y += dLin[i].y;
		bodies[i].linVel.z += dLin[i].z;
		bodies[i].angVel.x += dAng[i].x;
		bodies[i].angVel.y += dAng[i].y;
		bodies[i].angVel.z += dAng[i].z;
	}
}

// ---------------------------------------------------------------------------------------------
// b3GpuRaycast::castRaysHost (src/Bullet3OpenCL/Raycast/b3GpuRaycast.cpp:160-246) with rayConvex (:123-158) and
// sphere_intersect (:99-121).  One deliberate difference: the reference's SHAPE_SPHERE case has no `break` and falls
// through into the convex test with the sphere collidable's shape index; that accident is not restated.
extern "C" void orc_cast_rays(const b3b200_ray_info* rays, int numRays, b3b200_ray_hit* hits, const b3b200_rigid_body* bodies, int numBodies,
							  const b3b200_collidable* collidables, const b3b200_convex_polyhedron* convex, const b3b200_face* faces)
{
	for (int r = 0; r < numRays; r++)
	{
		V3 rayFrom = ld(rays[r].from), rayTo = ld(rays[r].to);
		float hitFraction = hits[r].hitFraction;
		int hitBodyIndex = -1;
		V3 hitNormal = mk(0, 0, 0);
		for (int b = 0; b < numBodies; b++)
		{
			const b3b200_collidable& col = collidables[bodies[b].collidableIdx];
			V3 pos = ld(bodies[b].pos);
			if (col.shapeType == B3B200_SHAPE_SPHERE)
			{
				float radius = col.radius;
				V3 rs = sub(rayFrom, pos), rayDir = sub(rayTo, rayFrom);
				float A = dot(rayDir, rayDir), B = dot(rs, rayDir), C = dot(rs, rs) - (radius * radius);
				float D = B * B - A * C;
				if (D > 0.0f)
				{
					float t = (-B - sqrtf(D)) / A;
					if ((t >= 0.0f) && (t < hitFraction))
					{
						hitFraction = t;
						hitBodyIndex = b;
						float s = 1.0f - t;
						V3 hp = mk(s * rayFrom.x + t * rayTo.x, s * rayFrom.y + t * rayTo.y, s * rayFrom.z + t * rayTo.z);
						hitNormal = normalized(sub(hp, pos));
					}
				}
			}
			else if (col.shapeType == B3B200_SHAPE_CONVEX_HULL)
			{
				M3 basis = matFromQuat(ld(bodies[b].quat)), inv = transposeM(basis);
				V3 invOrigin = matMul(inv, neg(mk(pos.x, pos.y, pos.z)));
				V3 fromL = add(matMul(inv, rayFrom), invOrigin), toL = add(matMul(inv, rayTo), invOrigin);
				const b3b200_convex_polyhedron& poly = convex[col.shapeIndex];
				float exitFraction = hitFraction, enterFraction = -0.1f;
				V3 curHitNormal = mk(0, 0, 0);
				bool hit = true;
				for (int i = 0; i < poly.numFaces && hit; i++)
				{
					const b3b200_float4& pl = faces[poly.faceOffset + i].plane;
					V3 n = mk(pl.x, pl.y, pl.z);
					float fromPlaneDist = dot(fromL, n) + pl.w, toPlaneDist = dot(toL, n) + pl.w;
					if (fromPlaneDist < 0.f)
					{
						if (toPlaneDist >= 0.f)
						{
							float fraction = fromPlaneDist / (fromPlaneDist - toPlaneDist);
							if (exitFraction > fraction) exitFraction = fraction;
						}
					}
					else
					{
						if (toPlaneDist < 0.f)
						{
							float fraction = fromPlaneDist / (fromPlaneDist - toPlaneDist);
							if (enterFraction <= fraction)
							{
								enterFraction = fraction;
								curHitNormal = n;
							}
						}
						else
							hit = false;
					}
					if (exitFraction <= enterFraction) hit = false;
				}
				if (hit && !(enterFraction < 0.f))
				{
					hitFraction = enterFraction;
					hitNormal = curHitNormal;
					hitBodyIndex = b;
				}
			}
		}
		if (hitBodyIndex >= 0)
		{
			float s = 1.0f - hitFraction;
			hits[r].hitFraction = hitFraction;
			hits[r].hitPoint = st(mk(s * rayFrom.x + hitFraction * rayTo.x, s * rayFrom.y + hitFraction * rayTo.y, s * rayFrom.z + hitFraction * rayTo.z));
			hits[r].hitNormal = st(hitNormal);
			hits[r].hitBody = hitBodyIndex;
		}
	}
}
