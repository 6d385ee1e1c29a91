// hull.cpp -- host-side convex polyhedron builder.
//
// Stands in for b3ConvexUtility::initializePolyhedralFeatures
// (src/Bullet3Collision/NarrowPhaseCollision/b3ConvexUtility.cpp:26-281, which
// leans on b3ConvexHullComputer) for the one-time shape registration path
// b3GpuNarrowPhase::registerConvexHullShape(const float*, stride, n, scaling)
// (b3GpuNarrowPhase.cpp:298-319).  Output conventions match the reference:
//  * faces are merged coplanar polygons, wound counter-clockwise seen from
//    outside, plane = (outward unit normal, -dot(normal, vertex))   (:84-98)
//  * uniqueEdges are normalised edge directions, de-duplicated up to sign with
//    the reference's 1e-6 per-component test                          (:341-367)
//  * vertices are the hull's extreme points only.
// The algorithm is a plain supporting-plane enumeration (O(n^4)); shapes are
// instanced, so this runs a handful of times at set-up.
#include <math.h>
#include <algorithm>
#include <vector>
#include "hull.h"

namespace b3b200
{
namespace
{
struct D3
{
	double x, y, z;
};
inline D3 sub(const D3& a, const D3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 cross(const D3& a, const D3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double len(const D3& a) { return sqrt(dot(a, a)); }

struct Plane
{
	D3 n;
	double d;  // dot(n, p) == d on the plane
	std::vector<int> verts;
};
}  // namespace

bool buildConvexHull(const std::vector<b3b200_float4>& pointsIn, HullOut& out)
{
	// de-duplicate
	std::vector<D3> pts;
	for (size_t i = 0; i < pointsIn.size(); i++)
	{
		D3 p = {pointsIn[i].x, pointsIn[i].y, pointsIn[i].z};
		bool dup = false;
		for (size_t j = 0; j < pts.size(); j++)
		{
			D3 q = sub(p, pts[j]);
			if (fabs(q.x) < 1e-7 && fabs(q.y) < 1e-7 && fabs(q.z) < 1e-7)
			{
				dup = true;
				break;
			}
		}
		if (!dup) pts.push_back(p);
	}
	const int n = (int)pts.size();
	if (n < 4) return false;
	D3 lo = pts[0], hi = pts[0];
	for (int i = 1; i < n; i++)
	{
		lo.x = std::min(lo.x, pts[i].x);
		lo.y = std::min(lo.y, pts[i].y);
		lo.z = std::min(lo.z, pts[i].z);
		hi.x = std::max(hi.x, pts[i].x);
		hi.y = std::max(hi.y, pts[i].y);
		hi.z = std::max(hi.z, pts[i].z);
	}
	const double scale = len(sub(hi, lo));
	if (!(scale > 0)) return false;
	const double tol = 1e-6 * scale;

	std::vector<Plane> planes;
	for (int i = 0; i < n; i++)
		for (int j = i + 1; j < n; j++)
			for (int k = j + 1; k < n; k++)
			{
				D3 nr = cross(sub(pts[j], pts[i]), sub(pts[k], pts[i]));
				double l = len(nr);
				if (l < 1e-12 * scale * scale) continue;
				nr = {nr.x / l, nr.y / l, nr.z / l};
				double d = dot(nr, pts[i]);
				double mx = -1e300, mn = 1e300;
				for (int p = 0; p < n; p++)
				{
					double t = dot(nr, pts[p]) - d;
					mx = std::max(mx, t);
					mn = std::min(mn, t);
				}
				if (mx <= tol)
				{
				}
				else if (mn >= -tol)
				{
					nr = {-nr.x, -nr.y, -nr.z};
					d = -d;
				}
				else
					continue;
				bool have = false;
				for (size_t q = 0; q < planes.size(); q++)
					if (dot(planes[q].n, nr) > 1.0 - 1e-9 && fabs(planes[q].d - d) < 10 * tol)
					{
						have = true;
						break;
					}
				if (have) continue;
				Plane pl;
				pl.n = nr;
				pl.d = d;
				planes.push_back(pl);
			}
	if (planes.size() < 4) return false;

	std::vector<char> used(n, 0);
	for (size_t q = 0; q < planes.size(); q++)
	{
		Plane& pl = planes[q];
		std::vector<int> on;
		for (int p = 0; p < n; p++)
			if (fabs(dot(pl.n, pts[p]) - pl.d) <= 4 * tol) on.push_back(p);
		if (on.size() < 3) continue;
		D3 c = {0, 0, 0};
		for (size_t t = 0; t < on.size(); t++)
		{
			c.x += pts[on[t]].x;
			c.y += pts[on[t]].y;
			c.z += pts[on[t]].z;
		}
		c = {c.x / on.size(), c.y / on.size(), c.z / on.size()};
		D3 u = sub(pts[on[0]], c);
		double ul = len(u);
		u = {u.x / ul, u.y / ul, u.z / ul};
		D3 v = cross(pl.n, u);
		std::vector<std::pair<double, int> > ang;
		for (size_t t = 0; t < on.size(); t++)
		{
			D3 r = sub(pts[on[t]], c);
			ang.push_back(std::make_pair(atan2(dot(r, v), dot(r, u)), on[t]));
		}
		std::sort(ang.begin(), ang.end());
		// drop collinear (non-extreme) points
		std::vector<int> poly;
		const int m = (int)ang.size();
		for (int t = 0; t < m; t++)
		{
			const D3& a = pts[ang[(t + m - 1) % m].second];
			const D3& b = pts[ang[t].second];
			const D3& cc = pts[ang[(t + 1) % m].second];
			D3 cr = cross(sub(b, a), sub(cc, b));
			if (dot(cr, pl.n) > 1e-9 * scale * scale) poly.push_back(ang[t].second);
		}
		if (poly.size() < 3) continue;
		pl.verts = poly;
		for (size_t t = 0; t < poly.size(); t++) used[poly[t]] = 1;
	}

	// re-index the extreme points (input order is kept)
	std::vector<int> remap(n, -1);
	out.vertices.clear();
	for (int p = 0; p < n; p++)
		if (used[p])
		{
			remap[p] = (int)out.vertices.size();
			b3b200_float4 f = {(float)pts[p].x, (float)pts[p].y, (float)pts[p].z, 0.f};
			out.vertices.push_back(f);
		}
	if (out.vertices.size() < 4) return false;

	out.faces.clear();
	out.indices.clear();
	out.uniqueEdges.clear();
	for (size_t q = 0; q < planes.size(); q++)
	{
		const Plane& pl = planes[q];
		if (pl.verts.size() < 3) continue;
		b3b200_face f;
		f.plane.x = (float)pl.n.x;
		f.plane.y = (float)pl.n.y;
		f.plane.z = (float)pl.n.z;
		// plane[3] = -min_v dot(v, n) over the face's vertices, in float (b3ConvexUtility.cpp:90-98)
		float planeEq = 1e30f;
		for (size_t t = 0; t < pl.verts.size(); t++)
		{
			const b3b200_float4& v = out.vertices[remap[pl.verts[t]]];
			float eq = v.x * f.plane.x + v.y * f.plane.y + v.z * f.plane.z;
			if (planeEq > eq) planeEq = eq;
		}
		f.plane.w = -planeEq;
		f.indexOffset = (int)out.indices.size();
		f.numIndices = (int)pl.verts.size();
		f.pad1 = f.pad2 = 0;
		for (size_t t = 0; t < pl.verts.size(); t++) out.indices.push_back(remap[pl.verts[t]]);
		out.faces.push_back(f);
	}
	// unique edge directions (b3ConvexUtility.cpp:331-367)
	for (size_t fi = 0; fi < out.faces.size(); fi++)
	{
		const b3b200_face& f = out.faces[fi];
		for (int j = 0; j < f.numIndices; j++)
		{
			int k = (j + 1) % f.numIndices;
			int v0 = out.indices[f.indexOffset + j], v1 = out.indices[f.indexOffset + k];
			if (v1 > v0) std::swap(v0, v1);  // b3InternalVertexPair orders (v0 >= v1)
			const b3b200_float4& a = out.vertices[v0];
			const b3b200_float4& b = out.vertices[v1];
			float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
			float inv = 1.0f / sqrtf(ex * ex + ey * ey + ez * ez);
			ex *= inv;
			ey *= inv;
			ez *= inv;
			bool found = false;
			for (size_t p = 0; p < out.uniqueEdges.size(); p++)
			{
				const b3b200_float4& u = out.uniqueEdges[p];
				bool z1 = !(fabsf(u.x - ex) > 1e-6 || fabsf(u.y - ey) > 1e-6 || fabsf(u.z - ez) > 1e-6);
				bool z2 = !(fabsf(u.x + ex) > 1e-6 || fabsf(u.y + ey) > 1e-6 || fabsf(u.z + ez) > 1e-6);
				if (z1 || z2)
				{
					found = true;
					break;
				}
			}
			if (!found)
			{
				b3b200_float4 e = {ex, ey, ez, 0.f};
				out.uniqueEdges.push_back(e);
			}
		}
	}
	// localCenter = vertex average (b3GpuNarrowPhase.cpp:332-337)
	b3b200_convex_polyhedron& poly = out.poly;
	poly = b3b200_convex_polyhedron();
	float cx = 0.f, cy = 0.f, cz = 0.f;
	for (size_t i = 0; i < out.vertices.size(); i++)
	{
		cx += out.vertices[i].x;
		cy += out.vertices[i].y;
		cz += out.vertices[i].z;
	}
	float s = 1.f / out.vertices.size();
	poly.localCenter.x = cx * s;
	poly.localCenter.y = cy * s;
	poly.localCenter.z = cz * s;
	poly.localCenter.w = 0.f;
	poly.numVertices = (int)out.vertices.size();
	poly.numFaces = (int)out.faces.size();
	poly.numUniqueEdges = (int)out.uniqueEdges.size();
	return true;
}

}  // namespace b3b200
