// hull.h -- host-side convex polyhedron builder (see hull.cpp)
#pragma once
#include <vector>
#include "../../include/b3b200_types.h"

namespace b3b200
{
struct HullOut
{
	std::vector<b3b200_float4> vertices, uniqueEdges;
	std::vector<b3b200_face> faces;
	std::vector<int> indices;
	b3b200_convex_polyhedron poly;
};
// replaces b3ConvexUtility::initializePolyhedralFeatures
// (src/Bullet3Collision/NarrowPhaseCollision/b3ConvexUtility.cpp:26)
bool buildConvexHull(const std::vector<b3b200_float4>& points, HullOut& out);
}  // namespace b3b200
