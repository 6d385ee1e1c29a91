// shapes.cu -- registration of the non-convex-hull collision shapes:
// b3GpuNarrowPhase::registerPlaneShape / registerSphereShape / registerCompoundShape / registerConcaveMesh
// (src/Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.cpp:159-231, 370-498).  Host side, one-time.
#include <string.h>
#include <math.h>
#include <algorithm>
#include "internal.h"

using namespace b3b200;

extern "C" int b3b200_register_plane(b3b200_world* w, const float* normal3, float planeConstant)
{
	if (!w || !normal3)
	{
		setLastError("registerPlaneShape: invalid argument");
		return -1;
	}
	int ci = allocateCollidable(w);
	if (ci < 0) return -1;
	// registerFace (:194-200): the plane equation lives in the face table
	b3b200_face f;
	memset(&f, 0, sizeof(f));
	f.plane.x = normal3[0];
	f.plane.y = normal3[1];
	f.plane.z = normal3[2];
	f.plane.w = planeConstant;
	w->faces.push_back(f);
	b3b200_collidable& col = w->collidables[ci];
	col.shapeType = B3B200_SHAPE_PLANE;
	col.shapeIndex = (int)w->faces.size() - 1;
	col.radius = planeConstant;
	b3b200_aabb& a = w->localAabbs[ci];
	for (int k = 0; k < 3; k++)
	{
		a.min[k] = -1e30f;
		a.max[k] = 1e30f;
	}
	a.minIndices[3] = 0;
	a.signedMaxIndices[3] = 0;
	return ci;
}

extern "C" int b3b200_register_sphere(b3b200_world* w, float radius)
{
	if (!w)
	{
		setLastError("registerSphereShape: invalid argument");
		return -1;
	}
	int ci = allocateCollidable(w);
	if (ci < 0) return -1;
	b3b200_collidable& col = w->collidables[ci];
	col.shapeType = B3B200_SHAPE_SPHERE;
	col.shapeIndex = 0;
	col.radius = radius;
	b3b200_aabb& a = w->localAabbs[ci];
	for (int k = 0; k < 3; k++)
	{
		a.min[k] = -radius;
		a.max[k] = radius;
	}
	a.minIndices[3] = 0;
	a.signedMaxIndices[3] = 0;
	return ci;
}

extern "C" int b3b200_register_compound(b3b200_world* w, const b3b200_child_shape* children, int numChildren)
{
	if (!w || !children || numChildren <= 0)
	{
		setLastError("registerCompoundShape: invalid argument");
		return -1;
	}
	if ((int)w->childShapes.size() + numChildren >= w->cfg.maxCompoundChildShapes)
	{
		setLastError("registerCompoundShape: exceeding m_maxCompoundChildShapes (%d)", w->cfg.maxCompoundChildShapes);
		return -1;
	}
	for (int i = 0; i < numChildren; i++)
	{
		int cc = children[i].shapeIndex;
		if (cc < 0 || cc >= (int)w->collidables.size() || w->collidables[cc].shapeType != B3B200_SHAPE_CONVEX_HULL)
		{
			setLastError("registerCompoundShape: child %d does not reference a convex hull collidable", i);
			return -1;
		}
	}
	int ci = allocateCollidable(w);
	if (ci < 0) return -1;
	b3b200_collidable& col = w->collidables[ci];
	col.shapeType = B3B200_SHAPE_COMPOUND_OF_CONVEX_HULLS;
	col.shapeIndex = (int)w->childShapes.size();
	// The reference builds a quantized BVH over the children here (:432-474).  This build culls child
	// pairs by their world AABBs directly (see narrowphase.cu), which yields the same contacts.
	col.compoundBvhIndex = -1;
	float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
	for (int i = 0; i < numChildren; i++)
	{
		b3b200_child_shape ch = children[i];
		ch.shapeType = B3B200_SHAPE_CONVEX_HULL;
		w->childShapes.push_back(ch);
		const b3b200_aabb& la = w->localAabbs[ch.shapeIndex];
		const float pos[3] = {ch.childPosition.x, ch.childPosition.y, ch.childPosition.z};
		const float orn[4] = {ch.childOrientation.x, ch.childOrientation.y, ch.childOrientation.z, ch.childOrientation.w};
		float amn[3], amx[3];
		transformAabbHost(la.min, la.max, 0.f, pos, orn, amn, amx);  // b3TransformAabb, margin 0 (:404-411)
		for (int k = 0; k < 3; k++)
		{
			if (amn[k] < mn[k]) mn[k] = amn[k];
			if (amx[k] > mx[k]) mx[k] = amx[k];
		}
	}
	// NB: `col` may dangle after push_backs on other vectors only; collidables was not resized since
	w->collidables[ci].numChildShapes = numChildren;
	b3b200_aabb& a = w->localAabbs[ci];
	for (int k = 0; k < 3; k++)
	{
		a.min[k] = mn[k];
		a.max[k] = mx[k];
	}
	a.minIndices[3] = 0;
	a.signedMaxIndices[3] = 0;
	return ci;
}

// ---------------------------------------------------------------------------------------------------------
// registerConcaveMesh (b3GpuNarrowPhase.cpp:521-605) + registerConcaveMeshShape (:607-668): the triangles become the
// faces of one entry of the convex table (plane = normalised (v1-v0)x(v2-v0), c = -n.v0, scaled vertices), exactly
// like the reference.  The reference then builds a quantized b3OptimizedBvh; its traversal is followed by an exact
// triangle-AABB test (shared/b3FindConcaveSatAxis.h:606-610), so any conservative tree gives the same active set.
// This build uses its own binary AABB tree with exact float boxes over the (scaled) triangles (median split of the
// centroids along the widest axis, <= 4 triangles per leaf), laid out for the warp traversal of concave.cu.
namespace
{
struct TriBox
{
	float mn[3], mx[3], c[3];
	int tri;
};
struct MeshBuilder
{
	std::vector<TriBox>& t;
	std::vector<b3b200_float4>& nodes;
	std::vector<int>& order;
	size_t nodeBase, triBase;
	// builds the subtree of t[lo, hi) into node slot `slot` (already allocated)
	void build(size_t slot, int lo, int hi)
	{
		float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f}, cmn[3] = {1e30f, 1e30f, 1e30f}, cmx[3] = {-1e30f, -1e30f, -1e30f};
		for (int i = lo; i < hi; i++)
			for (int k = 0; k < 3; k++)
			{
				mn[k] = std::min(mn[k], t[i].mn[k]);
				mx[k] = std::max(mx[k], t[i].mx[k]);
				cmn[k] = std::min(cmn[k], t[i].c[k]);
				cmx[k] = std::max(cmx[k], t[i].c[k]);
			}
		b3b200_float4 n0, n1;
		n0.x = mn[0];
		n0.y = mn[1];
		n0.z = mn[2];
		n1.x = mx[0];
		n1.y = mx[1];
		n1.z = mx[2];
		const int count = hi - lo;
		if (count <= 4)
		{
			int first = (int)(order.size() - triBase);
			for (int i = lo; i < hi; i++) order.push_back(t[i].tri);
			memcpy(&n0.w, &first, 4);
			memcpy(&n1.w, &count, 4);
			nodes[nodeBase + 2 * slot] = n0;
			nodes[nodeBase + 2 * slot + 1] = n1;
			return;
		}
		int axis = 0;
		for (int k = 1; k < 3; k++)
			if (cmx[k] - cmn[k] > cmx[axis] - cmn[axis]) axis = k;
		const int mid = (lo + hi) / 2;
		std::nth_element(t.begin() + lo, t.begin() + mid, t.begin() + hi, [axis](const TriBox& a, const TriBox& b) {
			return a.c[axis] < b.c[axis] || (a.c[axis] == b.c[axis] && a.tri < b.tri);
		});
		const size_t left = (nodes.size() - nodeBase) / 2;
		nodes.resize(nodes.size() + 4);
		int li = (int)left, zero = 0;
		memcpy(&n0.w, &li, 4);
		memcpy(&n1.w, &zero, 4);
		nodes[nodeBase + 2 * slot] = n0;
		nodes[nodeBase + 2 * slot + 1] = n1;
		build(left, lo, mid);
		build(left + 1, mid, hi);
	}
};
}  // namespace

extern "C" int b3b200_register_concave(b3b200_world* w, const float* vertices, int numVertices, const int* triIndices, int numIndices, const float* scaling3)
{
	if (!w || !vertices || !triIndices || !scaling3 || numVertices <= 0 || numIndices < 3 || numIndices % 3)
	{
		setLastError("registerConcaveMesh: invalid argument");
		return -1;
	}
	for (int i = 0; i < numIndices; i++)
		if (triIndices[i] < 0 || triIndices[i] >= numVertices)
		{
			setLastError("registerConcaveMesh: index %d out of range", i);
			return -1;
		}
	int ci = allocateCollidable(w);
	if (ci < 0) return -1;
	const float sx = scaling3[0], sy = scaling3[1], sz = scaling3[2];
	const int numTris = numIndices / 3;
	b3b200_convex_polyhedron cv;
	memset(&cv, 0, sizeof(cv));
	cv.uniqueEdgesOffset = (int)w->uniqueEdges.size();
	cv.faceOffset = (int)w->faces.size();
	cv.numFaces = numTris;
	cv.numVertices = numVertices;
	cv.vertexOffset = (int)w->vertices.size();
	auto vtx = [&](int i, float* o) {
		o[0] = vertices[3 * i] * sx;
		o[1] = vertices[3 * i + 1] * sy;
		o[2] = vertices[3 * i + 2] * sz;
	};
	std::vector<TriBox> boxes((size_t)numTris);
	for (int f = 0; f < numTris; f++)
	{
		float v0[3], v1[3], v2[3];
		vtx(triIndices[3 * f], v0);
		vtx(triIndices[3 * f + 1], v1);
		vtx(triIndices[3 * f + 2], v2);
		// ((vert1 - vert0).cross(vert2 - vert0)).normalize(): b3Vector3::normalize is *= 1/length (b3Vector3.h:303-306, 251-260)
		const float a[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, b[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
		float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
		const float rlen = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
		n[0] *= rlen;
		n[1] *= rlen;
		n[2] *= rlen;
		b3b200_face face;
		memset(&face, 0, sizeof(face));
		face.plane.x = n[0];
		face.plane.y = n[1];
		face.plane.z = n[2];
		face.plane.w = -(n[0] * v0[0] + n[1] * v0[1] + n[2] * v0[2]);
		face.indexOffset = (int)w->indices.size();
		face.numIndices = 3;
		w->faces.push_back(face);
		for (int k = 0; k < 3; k++) w->indices.push_back(triIndices[3 * f + k]);
		TriBox& tb = boxes[f];
		tb.tri = f;
		for (int k = 0; k < 3; k++)
		{
			tb.mn[k] = std::min(v0[k], std::min(v1[k], v2[k]));
			tb.mx[k] = std::max(v0[k], std::max(v1[k], v2[k]));
			tb.c[k] = (v0[k] + v1[k] + v2[k]) * (1.f / 3.f);
		}
	}
	float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
	for (int i = 0; i < numVertices; i++)
	{
		float v[3];
		vtx(i, v);
		b3b200_float4 o;
		o.x = v[0];
		o.y = v[1];
		o.z = v[2];
		o.w = 0.f;
		w->vertices.push_back(o);
		for (int k = 0; k < 3; k++)
		{
			mn[k] = std::min(mn[k], v[k]);
			mx[k] = std::max(mx[k], v[k]);
		}
	}
	w->convex.push_back(cv);
	b3b200_collidable& col = w->collidables[ci];
	col.shapeType = B3B200_SHAPE_CONCAVE_TRIMESH;
	col.shapeIndex = (int)w->convex.size() - 1;
	col.bvhIndex = (int)w->bvhInfos.size();
	b3b200_aabb& a = w->localAabbs[ci];
	for (int k = 0; k < 3; k++)
	{
		a.min[k] = mn[k];
		a.max[k] = mx[k];
	}
	a.minIndices[3] = 0;
	a.signedMaxIndices[3] = 0;
	// b3BvhInfo header as the reference computes it (b3QuantizedBvh::setQuantizationValues, margin 1); the quantized
	// node / subtree tables themselves are not produced (numNodes = numSubTrees = 0)
	b3b200_bvh_info info;
	memset(&info, 0, sizeof(info));
	const float lo[3] = {mn[0] - 1.f, mn[1] - 1.f, mn[2] - 1.f}, hi[3] = {mx[0] + 1.f, mx[1] + 1.f, mx[2] + 1.f};
	info.aabbMin.x = lo[0];
	info.aabbMin.y = lo[1];
	info.aabbMin.z = lo[2];
	info.aabbMax.x = hi[0];
	info.aabbMax.y = hi[1];
	info.aabbMax.z = hi[2];
	info.quantization.x = 65533.f / (hi[0] - lo[0]);
	info.quantization.y = 65533.f / (hi[1] - lo[1]);
	info.quantization.z = 65533.f / (hi[2] - lo[2]);
	w->bvhInfos.push_back(info);
	b3b200_int4 mi;
	mi.x = (int)(w->meshNodes.size() / 2);
	mi.z = (int)w->meshTris.size();
	mi.w = numTris;
	MeshBuilder mb = {boxes, w->meshNodes, w->meshTris, w->meshNodes.size(), w->meshTris.size()};
	w->meshNodes.resize(w->meshNodes.size() + 2);
	mb.build(0, 0, numTris);
	mi.y = (int)(w->meshNodes.size() / 2) - mi.x;
	w->meshInfos.push_back(mi);
	w->uploaded = false;
	return ci;
}
