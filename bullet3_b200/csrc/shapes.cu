// shapes.cu -- registration of the non-convex-hull collision shapes:
// b3GpuNarrowPhase::registerPlaneShape / registerSphereShape / registerCompoundShape / registerConcaveMesh
// (src/Bullet3OpenCL/RigidBody/b3GpuNarrowPhase.cpp:159-231, 370-498).  Host side, one-time.
#include <string.h>
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

extern "C" int b3b200_register_concave(b3b200_world* w, const float* vertices, int numVertices, const int* triIndices, int numIndices, const float* scaling3)
{
	(void)w;
	(void)vertices;
	(void)numVertices;
	(void)triIndices;
	(void)numIndices;
	(void)scaling3;
	setLastError("registerConcaveMesh: not built yet");
	return -1;
}
