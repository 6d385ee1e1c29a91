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

// ---------------------------------------------------------------------------------------------------------
// The reference's quantized BVH tables (b3QuantizedBvhNodeData 16 B, b3BvhSubtreeInfoData 32 B) for a set of leaf boxes, as
// b3QuantizedBvh::buildInternal produces them (b3QuantizedBvh.cpp:34-62, 79-88, 99-176, 178-208, 210-261, 263-290; helpers
// b3QuantizedBvh.h:189-268, 314-341, 389-398): 16-bit boxes inside the padded shape box (margin 1, 65533 steps; minima rounded
// down to even, maxima rounded up to odd), leaves reordered in place by the mean split along the axis of largest variance of
// the UNQUANTIZED-again leaf centres (with the reference's fallback to the middle when a side gets < 1/3), nodes in depth-first
// order with escape indices (-subtree size on internal nodes), 2 * leaves slots (the last one stays zero), and a subtree
// header for every child subtree of <= 128 nodes whose parent is larger (or one header for a tree that small).
// This build's own narrowphase walks different structures (float trees, bounding-sphere child culls); the tables exist so that
// getStatic / serialisation consumers of b3GpuNarrowPhase see what the reference would hand them.
namespace
{
struct QuantizedBvhBuilder
{
	float bvhMin[3], bvhMax[3], quant[3];
	std::vector<b3b200_bvh_node> leaves, nodes;
	std::vector<b3b200_bvh_subtree> headers;
	int cur = 0;

	void setQuantizationValues(const float* mn, const float* mx)
	{
		for (int k = 0; k < 3; k++)
		{
			bvhMin[k] = mn[k] - 1.0f;
			bvhMax[k] = mx[k] + 1.0f;
			quant[k] = 65533.0f / (bvhMax[k] - bvhMin[k]);
		}
	}
	void quantize(unsigned short* out, const float* p, int isMax) const
	{
		for (int k = 0; k < 3; k++)
		{
			const float v = (p[k] - bvhMin[k]) * quant[k];
			out[k] = isMax ? (unsigned short)(((unsigned short)(int)(v + 1.0f)) | 1) : (unsigned short)(((unsigned short)(int)v) & 0xfffe);
		}
	}
	void unQuantize(const unsigned short* in, float* out) const
	{
		for (int k = 0; k < 3; k++) out[k] = (float)in[k] / quant[k] + bvhMin[k];
	}
	void addLeaf(const float* mn, const float* mx, int index)
	{
		b3b200_bvh_node n;
		quantize(n.quantizedAabbMin, mn, 0);
		quantize(n.quantizedAabbMax, mx, 1);
		n.escapeIndexOrTriangleIndex = index;  // part 0
		leaves.push_back(n);
	}
	void centre(int i, float* c) const
	{
		float mn[3], mx[3];
		unQuantize(leaves[i].quantizedAabbMin, mn);
		unQuantize(leaves[i].quantizedAabbMax, mx);
		for (int k = 0; k < 3; k++) c[k] = 0.5f * (mx[k] + mn[k]);
	}
	int calcSplittingAxis(int start, int end) const
	{
		float means[3] = {0.f, 0.f, 0.f}, variance[3] = {0.f, 0.f, 0.f};
		const int num = end - start;
		for (int i = start; i < end; i++)
		{
			float c[3];
			centre(i, c);
			for (int k = 0; k < 3; k++) means[k] += c[k];
		}
		const float inv = 1.0f / (float)num;
		for (int k = 0; k < 3; k++) means[k] *= inv;
		for (int i = start; i < end; i++)
		{
			float c[3];
			centre(i, c);
			for (int k = 0; k < 3; k++)
			{
				const float d = c[k] - means[k];
				variance[k] += d * d;
			}
		}
		const float invv = 1.0f / ((float)num - 1);
		for (int k = 0; k < 3; k++) variance[k] *= invv;
		// b3Vector3::maxAxis
		return variance[0] < variance[1] ? (variance[1] < variance[2] ? 2 : 1) : (variance[0] < variance[2] ? 2 : 0);
	}
	int sortAndCalcSplittingIndex(int start, int end, int axis)
	{
		int split = start;
		const int num = end - start;
		float means[3] = {0.f, 0.f, 0.f};
		for (int i = start; i < end; i++)
		{
			float c[3];
			centre(i, c);
			for (int k = 0; k < 3; k++) means[k] += c[k];
		}
		const float inv = 1.0f / (float)num;
		for (int k = 0; k < 3; k++) means[k] *= inv;
		const float splitValue = means[axis];
		for (int i = start; i < end; i++)
		{
			float c[3];
			centre(i, c);
			if (c[axis] > splitValue)
			{
				std::swap(leaves[i], leaves[split]);
				split++;
			}
		}
		const int range = num / 3;
		if (split <= start + range || split >= end - 1 - range) split = start + (num >> 1);
		return split;
	}
	void header(int nodeIndex)
	{
		const b3b200_bvh_node& n = nodes[nodeIndex];
		b3b200_bvh_subtree h;
		memset(&h, 0, sizeof(h));
		for (int k = 0; k < 3; k++)
		{
			h.quantizedAabbMin[k] = n.quantizedAabbMin[k];
			h.quantizedAabbMax[k] = n.quantizedAabbMax[k];
		}
		h.rootNodeIndex = nodeIndex;
		h.subtreeSize = n.escapeIndexOrTriangleIndex >= 0 ? 1 : -n.escapeIndexOrTriangleIndex;
		headers.push_back(h);
	}
	void buildTree(int start, int end)
	{
		const int num = end - start, curIndex = cur;
		if (num == 1)
		{
			nodes[cur++] = leaves[start];
			return;
		}
		const int axis = calcSplittingAxis(start, end);
		const int split = sortAndCalcSplittingIndex(start, end, axis);
		const int internal = cur;
		quantize(nodes[internal].quantizedAabbMin, bvhMax, 0);
		quantize(nodes[internal].quantizedAabbMax, bvhMin, 1);
		for (int i = start; i < end; i++)
		{
			// mergeInternalNodeAabb takes the leaf's UNQUANTIZED box and quantizes it again
			float mn[3], mx[3];
			unQuantize(leaves[i].quantizedAabbMin, mn);
			unQuantize(leaves[i].quantizedAabbMax, mx);
			unsigned short qmn[3], qmx[3];
			quantize(qmn, mn, 0);
			quantize(qmx, mx, 1);
			for (int k = 0; k < 3; k++)
			{
				if (nodes[internal].quantizedAabbMin[k] > qmn[k]) nodes[internal].quantizedAabbMin[k] = qmn[k];
				if (nodes[internal].quantizedAabbMax[k] < qmx[k]) nodes[internal].quantizedAabbMax[k] = qmx[k];
			}
		}
		cur++;
		const int left = cur;
		buildTree(start, split);
		const int right = cur;
		buildTree(split, end);
		const int escape = cur - curIndex;
		if (escape * (int)sizeof(b3b200_bvh_node) > 2048)  // MAX_SUBTREE_SIZE_IN_BYTES
		{
			const int ls = nodes[left].escapeIndexOrTriangleIndex >= 0 ? 1 : -nodes[left].escapeIndexOrTriangleIndex;
			const int rs = nodes[right].escapeIndexOrTriangleIndex >= 0 ? 1 : -nodes[right].escapeIndexOrTriangleIndex;
			if (ls * (int)sizeof(b3b200_bvh_node) <= 2048) header(left);
			if (rs * (int)sizeof(b3b200_bvh_node) <= 2048) header(right);
		}
		nodes[internal].escapeIndexOrTriangleIndex = -escape;
	}
	void build()
	{
		b3b200_bvh_node zero;
		memset(&zero, 0, sizeof(zero));
		nodes.assign(2 * leaves.size(), zero);
		cur = 0;
		if (!leaves.empty()) buildTree(0, (int)leaves.size());
		if (headers.empty() && !nodes.empty()) header(0);
	}
	// appends the tables to the world's and returns the b3BvhInfo (b3GpuNarrowPhase.cpp:575-603)
	b3b200_bvh_info append(World* w)
	{
		b3b200_bvh_info info;
		memset(&info, 0, sizeof(info));
		info.aabbMin.x = bvhMin[0];
		info.aabbMin.y = bvhMin[1];
		info.aabbMin.z = bvhMin[2];
		info.aabbMax.x = bvhMax[0];
		info.aabbMax.y = bvhMax[1];
		info.aabbMax.z = bvhMax[2];
		info.quantization.x = quant[0];
		info.quantization.y = quant[1];
		info.quantization.z = quant[2];
		info.numNodes = (int)nodes.size();
		info.numSubTrees = (int)headers.size();
		info.nodeOffset = (int)w->bvhNodes.size();
		info.subTreeOffset = (int)w->bvhSubtrees.size();
		w->bvhNodes.insert(w->bvhNodes.end(), nodes.begin(), nodes.end());
		w->bvhSubtrees.insert(w->bvhSubtrees.end(), headers.begin(), headers.end());
		return info;
	}
};
}  // namespace

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
	// The reference builds a quantized BVH over the children (:439-490): the tables are produced below for its consumers; this
	// build's narrowphase culls child pairs by bounding spheres instead (narrowphase.cu), which yields the same contacts.
	col.compoundBvhIndex = (int)w->bvhInfos.size();
	float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
	std::vector<float> childBoxes((size_t)numChildren * 6);
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
			childBoxes[(size_t)i * 6 + k] = amn[k];
			childBoxes[(size_t)i * 6 + 3 + k] = amx[k];
		}
	}
	{
		QuantizedBvhBuilder qb;
		qb.setQuantizationValues(mn, mx);
		for (int i = 0; i < numChildren; i++) qb.addLeaf(&childBoxes[(size_t)i * 6], &childBoxes[(size_t)i * 6 + 3], i);
		qb.build();
		b3b200_bvh_info info = qb.append(w);
		info.numNodes = numChildren;  // the reference stores the number of CHILDREN here (:442, :470) while it appends all 2 n node slots
		w->bvhInfos.push_back(info);
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
	// the reference's b3OptimizedBvh tables (b3OptimizedBvh.cpp:27-190): quantization box = the SCALED mesh box padded by 1, leaves =
	// the boxes of the UNSCALED triangles (b3GpuNarrowPhase.cpp:562-573 hands the tree the caller's vertices), each widened to at
	// least 0.002 per axis, leaf index = triangle index
	{
		QuantizedBvhBuilder qb;
		qb.setQuantizationValues(mn, mx);
		for (int f = 0; f < numTris; f++)
		{
			float tmn[3] = {1e18f, 1e18f, 1e18f}, tmx[3] = {-1e18f, -1e18f, -1e18f};  // B3_LARGE_FLOAT
			for (int c = 0; c < 3; c++)
			{
				const float* v = &vertices[3 * triIndices[3 * f + c]];
				for (int k = 0; k < 3; k++)
				{
					tmn[k] = std::min(tmn[k], v[k]);
					tmx[k] = std::max(tmx[k], v[k]);
				}
			}
			for (int k = 0; k < 3; k++)
				if (tmx[k] - tmn[k] < 0.002f)
				{
					tmx[k] = tmx[k] + 0.001f;
					tmn[k] = tmn[k] - 0.001f;
				}
			qb.addLeaf(tmn, tmx, f);
		}
		qb.build();
		w->bvhInfos.push_back(qb.append(w));
	}
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
