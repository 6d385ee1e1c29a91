// raycast.cu -- b3GpuRigidBodyPipeline::castRays / b3GpuRaycast::castRays
// (src/Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.cpp:671-680, src/Bullet3OpenCL/Raycast/b3GpuRaycast.cpp:249-373).
// Result semantics are those of the reference's host twin castRaysHost (:160-246): closest hit, ties to the lower body
// index, hull normal in the hull's local frame.  The reference walks the bodies sequentially per ray (host twin) or
// builds a PLBVH and sorts ray-body pairs (GPU path); here
//   rayCandidatesKernel  one warp per (ray, chunk of 2048 bodies): lanes cull by a segment-vs-world-AABB slab test,
//                        run the exact body test on the survivors and merge the best (fraction, body) of the chunk
//                        into a 64-bit per-ray key with atomicMin (non-negative float bits order like the floats)
//   rayResolveKernel     one thread per ray: recomputes the winning body's hit and writes the b3RayHit
// i.e. O(rays x bodies) AABB tests (32 B each, L2-resident): the path for a few picking rays or small worlds.
// Many rays against many bodies go through a linear BVH rebuilt per call (the role of b3GpuParallelLinearBvh,
// b3GpuParallelLinearBvh.cpp:108-557, with a flatter layout):
//   rayBoundsKernel / rayMortonKernel   30-bit Morton code of every small body's AABB centre inside the bounds of all centres
//   radixSortKV32                       bodies into Morton order
//   rayLeafKernel / raySuperKernel      implicit 32-ary tree over that order: leaf = body AABB, chunk = union of 32 leaves,
//                                       super = union of 32 chunks (N = 262 144 -> 8 192 chunks, 256 supers)
//   rayTreeKernel                       one warp per ray: lanes test 32 supers / 32 chunks / 32 leaves at a time against the
//                                       segment clipped to the best hit so far, exact body test on the surviving lanes;
//                                       the broadphase's "large" bodies (static ground / meshes) are tested directly
// Both paths give the same answer: culling is conservative, the exact test always runs against the caller's cap, and
// the winner is the minimum of (fraction bits, body index).
#include "internal.h"

namespace b3b200
{
namespace
{
constexpr int RC_CHUNK = 2048;
#define FULL 0xffffffffu

struct RayArgs
{
	const b3b200_ray_info* rays;
	b3b200_ray_hit* hits;
	unsigned long long* best;
	int numRays, numBodies, numChunks;
	const float4* pose;
	const int* coll;
	const b3b200_collidable* collidables;
	const b3b200_convex_polyhedron* convex;
	const b3b200_face* faces;
	const b3b200_aabb* aabbs;
};

// exact test of one body; returns true and the entering fraction / normal when the ray hits it before `limit`
// (rayConvex / sphere_intersect, b3GpuRaycast.cpp:99-158, with the transform arithmetic of castRaysHost :192-200)
B3_D bool rayBody(const RayArgs& a, int b, const float4& rayFrom, const float4& rayTo, float limit, float& tOut, float4& normalOut)
{
	const int ci = a.coll[b];
	if (ci < 0) return false;
	const int type = __ldg(&a.collidables[ci].shapeType);
	const float4 pos = a.pose[2 * b];
	if (type == B3B200_SHAPE_SPHERE)
	{
		const float radius = __ldg(&a.collidables[ci].radius);
		const float4 rs = sub3(rayFrom, pos);
		const float4 rayDir = sub3(rayTo, rayFrom);
		const float A = dot3(rayDir, rayDir);
		const float B = dot3(rs, rayDir);
		const float C = dot3(rs, rs) - (radius * radius);
		const float D = B * B - A * C;
		if (D > 0.0f)
		{
			const float t = (-B - sqrtf(D)) / A;
			if ((t >= 0.0f) && (t < limit))
			{
				// hitPoint.setInterpolate3(from, to, t); hitNormal = (hitPoint - pos).normalize()
				const float s = 1.0f - t;
				const float4 hp = mk4(s * rayFrom.x + t * rayTo.x, s * rayFrom.y + t * rayTo.y, s * rayFrom.z + t * rayTo.z);
				tOut = t;
				normalOut = normalized3(sub3(hp, pos));
				return true;
			}
		}
		return false;
	}
	if (type != B3B200_SHAPE_CONVEX_HULL) return false;  // "Raytest: unsupported shape type" (:236-243)
	const float4 quat = a.pose[2 * b + 1];
	const Mat3 basis = matFromQuat(quat);
	Mat3 inv;
	inv.r0 = mk4(basis.r0.x, basis.r1.x, basis.r2.x);
	inv.r1 = mk4(basis.r0.y, basis.r1.y, basis.r2.y);
	inv.r2 = mk4(basis.r0.z, basis.r1.z, basis.r2.z);
	const float4 invOrigin = matMulVec(inv, neg3(mk4(pos.x, pos.y, pos.z)));
	const float4 fromL = add3(matMulVec(inv, rayFrom), invOrigin), toL = add3(matMulVec(inv, rayTo), invOrigin);
	const b3b200_convex_polyhedron* poly = &a.convex[__ldg(&a.collidables[ci].shapeIndex)];
	const int numFaces = __ldg(&poly->numFaces), faceOffset = __ldg(&poly->faceOffset);
	float exitFraction = limit;
	float enterFraction = -0.1f;
	float4 curHitNormal = mk4(0, 0, 0);
	for (int i = 0; i < numFaces; i++)
	{
		const float4 plane = __ldg(reinterpret_cast<const float4*>(&a.faces[faceOffset + i].plane));
		const float fromPlaneDist = dot3(fromL, plane) + plane.w;
		const float toPlaneDist = dot3(toL, plane) + plane.w;
		if (fromPlaneDist < 0.f)
		{
			if (toPlaneDist >= 0.f)
			{
				const float fraction = fromPlaneDist / (fromPlaneDist - toPlaneDist);
				if (exitFraction > fraction) exitFraction = fraction;
			}
		}
		else
		{
			if (toPlaneDist < 0.f)
			{
				const float fraction = fromPlaneDist / (fromPlaneDist - toPlaneDist);
				if (enterFraction <= fraction)
				{
					enterFraction = fraction;
					curHitNormal = mk4(plane.x, plane.y, plane.z);
				}
			}
			else
				return false;
		}
		if (exitFraction <= enterFraction) return false;
	}
	if (enterFraction < 0.f) return false;
	tOut = enterFraction;
	normalOut = curHitNormal;
	return true;
}

// conservative: does the segment from + t (to - from), t in [0, tMax], touch the box?
B3_D bool segmentHitsAabb(const float4& from, const float4& to, const float4& mn, const float4& mx, float tMax = 1.f)
{
	float t0 = 0.f, t1 = tMax;
	const float f[3] = {from.x, from.y, from.z}, d[3] = {to.x - from.x, to.y - from.y, to.z - from.z};
	const float lo[3] = {mn.x, mn.y, mn.z}, hi[3] = {mx.x, mx.y, mx.z};
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		if (fabsf(d[k]) < 1e-12f)
		{
			if (f[k] < lo[k] - 1e-4f || f[k] > hi[k] + 1e-4f) return false;
		}
		else
		{
			const float inv = 1.0f / d[k];
			float ta = (lo[k] - 1e-4f - f[k]) * inv, tb = (hi[k] + 1e-4f - f[k]) * inv;
			if (ta > tb)
			{
				const float t = ta;
				ta = tb;
				tb = t;
			}
			t0 = fmaxf(t0, ta - 1e-5f);
			t1 = fminf(t1, tb + 1e-5f);
			if (t0 > t1) return false;
		}
	}
	return true;
}

__global__ void rayInitKernel(RayArgs a)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.numRays) return;
	const float limit = a.hits[r].hitFraction;
	a.best[r] = ((unsigned long long)__float_as_uint(limit > 0.f ? limit : 0.f) << 32) | 0xffffffffull;
}

__global__ void __launch_bounds__(128) rayCandidatesKernel(RayArgs a)
{
	const int lane = threadIdx.x & 31;
	const long long warpId = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long total = (long long)a.numRays * a.numChunks;
	const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long wi = warpId; wi < total; wi += warps)
	{
		const int r = (int)(wi / a.numChunks), chunk = (int)(wi - (long long)r * a.numChunks);
		const float4 from = *reinterpret_cast<const float4*>(&a.rays[r].from), to = *reinterpret_cast<const float4*>(&a.rays[r].to);
		const float limit = a.hits[r].hitFraction;
		unsigned long long key = ~0ull;
		const int begin = chunk * RC_CHUNK, end = min(begin + RC_CHUNK, a.numBodies);
		for (int b = begin + lane; b < end; b += 32)
		{
			const float4 mn = *reinterpret_cast<const float4*>(a.aabbs[b].min), mx = *reinterpret_cast<const float4*>(a.aabbs[b].max);
			if (!segmentHitsAabb(from, to, mn, mx, fmaxf(limit, 0.f))) continue;
			float t;
			float4 n;
			if (rayBody(a, b, from, to, limit, t, n))
			{
				const unsigned long long k = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned int)b;
				key = k < key ? k : key;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			const unsigned long long other = __shfl_xor_sync(FULL, key, o);
			key = other < key ? other : key;
		}
		if (lane == 0 && key != ~0ull) atomicMin(&a.best[r], key);
	}
}

__global__ void rayResolveKernel(RayArgs a)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.numRays) return;
	const unsigned long long key = a.best[r];
	const unsigned int body = (unsigned int)(key & 0xffffffffull);
	if (body == 0xffffffffu) return;  // nothing closer than the caller's limit
	const float4 from = *reinterpret_cast<const float4*>(&a.rays[r].from), to = *reinterpret_cast<const float4*>(&a.rays[r].to);
	float t;
	float4 n;
	if (!rayBody(a, (int)body, from, to, a.hits[r].hitFraction, t, n)) return;
	const float s = 1.0f - t;
	b3b200_ray_hit h = a.hits[r];
	h.hitFraction = t;
	h.hitBody = (int)body;
	h.hitPoint.x = s * from.x + t * to.x;  // b3Vector3::setInterpolate3 (b3Vector3.h:513-530)
	h.hitPoint.y = s * from.y + t * to.y;
	h.hitPoint.z = s * from.z + t * to.z;
	h.hitPoint.w = 0.f;
	h.hitNormal.x = n.x;
	h.hitNormal.y = n.y;
	h.hitNormal.z = n.z;
	h.hitNormal.w = 0.f;
	a.hits[r] = h;
}

// ---------------------------------------------------------------- linear BVH path
struct RayTree
{
	const float4* leaf;   // 2 per sorted body: {min.xyz, bits(body)}, {max.xyz, 0}
	const float4* chunk;  // 2 per 32 leaves
	const float4* super;  // 2 per 32 chunks
	int numLeaves, numChunks, numSupers;
	const int* largeMap;
	int numLarge;
};

B3_D unsigned int orderedBits(float f)
{
	const unsigned int u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
B3_D float fromOrderedBits(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
B3_D unsigned int expandBits10(unsigned int v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

// bounds[0..2] = min, bounds[3..5] = max of the AABB centres (ordered-int encoding), cleared to (~0, 0) by the host
__global__ void __launch_bounds__(256) rayBoundsKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int n, unsigned int* __restrict__ bounds)
{
	float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const int b = smallMap[i];
		const float4 lo = *reinterpret_cast<const float4*>(aabbs[b].min), hi = *reinterpret_cast<const float4*>(aabbs[b].max);
		const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
#pragma unroll
		for (int k = 0; k < 3; k++)
		{
			mn[k] = fminf(mn[k], c[k]);
			mx[k] = fmaxf(mx[k], c[k]);
		}
	}
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			mn[k] = fminf(mn[k], __shfl_xor_sync(FULL, mn[k], o));
			mx[k] = fmaxf(mx[k], __shfl_xor_sync(FULL, mx[k], o));
		}
		if ((threadIdx.x & 31) == 0 && mn[k] <= mx[k])
		{
			atomicMin(&bounds[k], orderedBits(mn[k]));
			atomicMax(&bounds[3 + k], orderedBits(mx[k]));
		}
	}
}

__global__ void __launch_bounds__(256) rayMortonKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int n,
													   const unsigned int* __restrict__ bounds, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int b = smallMap[i];
	const float4 lo = *reinterpret_cast<const float4*>(aabbs[b].min), hi = *reinterpret_cast<const float4*>(aabbs[b].max);
	const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
	unsigned int q[3];
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		const float mn = fromOrderedBits(bounds[k]), mx = fromOrderedBits(bounds[3 + k]);
		const float ext = mx - mn;
		float f = ext > 0.f ? (c[k] - mn) / ext * 1023.0f : 0.f;
		f = fminf(fmaxf(f, 0.f), 1023.0f);  // (NaN / inf positions land in cell 0 / 1023: only the tree's tightness suffers)
		q[k] = (unsigned int)f;
	}
	keys[i] = expandBits10(q[0]) | (expandBits10(q[1]) << 1) | (expandBits10(q[2]) << 2);
	vals[i] = (unsigned int)b;
}

// one warp per 32 consecutive entries: copy (leaf level) and union
__global__ void __launch_bounds__(256) rayLeafKernel(const b3b200_aabb* __restrict__ aabbs, const unsigned int* __restrict__ order, int n, float4* __restrict__ leaf,
													 float4* __restrict__ chunk)
{
	const int lane = threadIdx.x & 31;
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c * 32 >= n) return;
	const int i = c * 32 + lane;
	float4 mn = mk4(FLT_MAX, FLT_MAX, FLT_MAX), mx = mk4(-FLT_MAX, -FLT_MAX, -FLT_MAX);
	if (i < n)
	{
		const int b = (int)order[i];
		mn = *reinterpret_cast<const float4*>(aabbs[b].min);
		mx = *reinterpret_cast<const float4*>(aabbs[b].max);
		mn.w = __int_as_float(b);
		mx.w = 0.f;
		leaf[2 * i] = mn;
		leaf[2 * i + 1] = mx;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		mn.x = fminf(mn.x, __shfl_xor_sync(FULL, mn.x, o));
		mn.y = fminf(mn.y, __shfl_xor_sync(FULL, mn.y, o));
		mn.z = fminf(mn.z, __shfl_xor_sync(FULL, mn.z, o));
		mx.x = fmaxf(mx.x, __shfl_xor_sync(FULL, mx.x, o));
		mx.y = fmaxf(mx.y, __shfl_xor_sync(FULL, mx.y, o));
		mx.z = fmaxf(mx.z, __shfl_xor_sync(FULL, mx.z, o));
	}
	if (lane == 0)
	{
		chunk[2 * c] = mk4(mn.x, mn.y, mn.z);
		chunk[2 * c + 1] = mk4(mx.x, mx.y, mx.z);
	}
}

__global__ void __launch_bounds__(256) raySuperKernel(const float4* __restrict__ chunk, int numChunks, float4* __restrict__ super)
{
	const int lane = threadIdx.x & 31;
	const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (s * 32 >= numChunks) return;
	const int c = s * 32 + lane;
	float4 mn = mk4(FLT_MAX, FLT_MAX, FLT_MAX), mx = mk4(-FLT_MAX, -FLT_MAX, -FLT_MAX);
	if (c < numChunks)
	{
		mn = chunk[2 * c];
		mx = chunk[2 * c + 1];
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		mn.x = fminf(mn.x, __shfl_xor_sync(FULL, mn.x, o));
		mn.y = fminf(mn.y, __shfl_xor_sync(FULL, mn.y, o));
		mn.z = fminf(mn.z, __shfl_xor_sync(FULL, mn.z, o));
		mx.x = fmaxf(mx.x, __shfl_xor_sync(FULL, mx.x, o));
		mx.y = fmaxf(mx.y, __shfl_xor_sync(FULL, mx.y, o));
		mx.z = fmaxf(mx.z, __shfl_xor_sync(FULL, mx.z, o));
	}
	if (lane == 0)
	{
		super[2 * s] = mn;
		super[2 * s + 1] = mx;
	}
}

B3_D unsigned long long warpMinKey(unsigned long long key)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const unsigned long long other = __shfl_xor_sync(FULL, key, o);
		key = other < key ? other : key;
	}
	return key;
}

__global__ void __launch_bounds__(128) rayTreeKernel(RayArgs a, RayTree t)
{
	const int lane = threadIdx.x & 31;
	const int warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = warpId; r < a.numRays; r += warps)
	{
		const float4 from = *reinterpret_cast<const float4*>(&a.rays[r].from), to = *reinterpret_cast<const float4*>(&a.rays[r].to);
		const float cap = a.hits[r].hitFraction;
		unsigned long long key = ~0ull;
		for (int l = lane; l < t.numLarge; l += 32)
		{
			const int b = t.largeMap[l];
			float tt;
			float4 n;
			if (rayBody(a, b, from, to, cap, tt, n))
			{
				const unsigned long long k = ((unsigned long long)__float_as_uint(tt) << 32) | (unsigned int)b;
				key = k < key ? k : key;
			}
		}
		// the segment is clipped to the best hit so far (inclusive, so that ties between bodies are still all seen)
		float tCull = cap > 0.f ? cap : 0.f;
		{
			const unsigned long long k = warpMinKey(key);
			if (k != ~0ull) tCull = fminf(tCull, __uint_as_float((unsigned int)(k >> 32)));
		}
		for (int s0 = 0; s0 < t.numSupers; s0 += 32)
		{
			const int s = s0 + lane;
			const bool hitS = s < t.numSupers && segmentHitsAabb(from, to, t.super[2 * s], t.super[2 * s + 1], tCull);
			unsigned int maskS = __ballot_sync(FULL, hitS);
			while (maskS)
			{
				const int sIdx = s0 + __ffs(maskS) - 1;
				maskS &= maskS - 1;
				const int c = sIdx * 32 + lane;
				const bool hitC = c < t.numChunks && segmentHitsAabb(from, to, t.chunk[2 * c], t.chunk[2 * c + 1], tCull);
				unsigned int maskC = __ballot_sync(FULL, hitC);
				while (maskC)
				{
					const int cIdx = sIdx * 32 + __ffs(maskC) - 1;
					maskC &= maskC - 1;
					const int i = cIdx * 32 + lane;
					bool found = false;
					float tt = 0.f;
					if (i < t.numLeaves)
					{
						const float4 mn = t.leaf[2 * i], mx = t.leaf[2 * i + 1];
						if (segmentHitsAabb(from, to, mn, mx, tCull))
						{
							const int b = __float_as_int(mn.w);
							float4 n;
							if (rayBody(a, b, from, to, cap, tt, n))
							{
								const unsigned long long k = ((unsigned long long)__float_as_uint(tt) << 32) | (unsigned int)b;
								key = k < key ? k : key;
								found = true;
							}
						}
					}
					if (__any_sync(FULL, found))
					{
						float m = found ? tt : FLT_MAX;
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
						tCull = fminf(tCull, m);
					}
				}
			}
		}
		key = warpMinKey(key);
		if (lane == 0)
		{
			const float lim = cap > 0.f ? cap : 0.f;
			a.best[r] = key != ~0ull ? key : (((unsigned long long)__float_as_uint(lim) << 32) | 0xffffffffull);
		}
	}
}
}  // namespace
}  // namespace b3b200

using namespace b3b200;

extern "C" int b3b200_cast_rays(b3b200_world* w, const b3b200_ray_info* rays, int numRays, b3b200_ray_hit* hits)
{
	if (!w || w->device < 0 || !w->uploaded) return B3B200_ERR_STATE;
	if (numRays < 0 || (numRays > 0 && (!rays || !hits))) return B3B200_ERR_INVALID;
	if (numRays == 0 || w->numBodies == 0) return 0;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	cudaStream_t s = w->stream;
	if (!w->aabbsValid) B3_TRY(launchUpdateAabbs(w));
	B3_TRY(w->dRays.reserve((size_t)numRays * 2));
	B3_TRY(w->dRayHits.reserve((size_t)numRays * 3));
	B3_TRY(w->dRayBest.reserve((size_t)numRays));
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dRays.ptr, rays, sizeof(b3b200_ray_info) * (size_t)numRays, cudaMemcpyHostToDevice, s));
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dRayHits.ptr, hits, sizeof(b3b200_ray_hit) * (size_t)numRays, cudaMemcpyHostToDevice, s));
	RayArgs a;
	a.rays = reinterpret_cast<const b3b200_ray_info*>(w->dRays.ptr);
	a.hits = reinterpret_cast<b3b200_ray_hit*>(w->dRayHits.ptr);
	a.best = w->dRayBest.ptr;
	a.numRays = numRays;
	a.numBodies = w->numBodies;
	a.numChunks = divUp(w->numBodies, RC_CHUNK);
	a.pose = w->dPose.ptr;
	a.coll = w->dCollidableIdx.ptr;
	a.collidables = w->dCollidables.ptr;
	a.convex = w->dConvex.ptr;
	a.faces = w->dFaces.ptr;
	a.aabbs = w->bp.aabbs.ptr;
	const int numSmall = w->bp.numSmall;
	const bool useTree = w->rayAccel == 1 || (w->rayAccel < 0 && numSmall >= 4096 && numRays >= 64);
	if (useTree && numSmall > 0)
	{
		const int numChunks = divUp(numSmall, 32), numSupers = divUp(numChunks, 32);
		B3_TRY(w->dRayKeys.reserve((size_t)numSmall));
		B3_TRY(w->dRayOrder.reserve((size_t)numSmall));
		B3_TRY(w->dRayTree.reserve(2 * ((size_t)numSmall + numChunks + numSupers)));
		B3_TRY(w->dRayBounds.reserve(8));
		const unsigned int init[6] = {~0u, ~0u, ~0u, 0u, 0u, 0u};
		B3_CUDA_CHECK(cudaMemcpyAsync(w->dRayBounds.ptr, init, sizeof(init), cudaMemcpyHostToDevice, s));
		rayBoundsKernel<<<std::min(divUp(numSmall, 256), w->smCount * 4), 256, 0, s>>>(w->bp.aabbs.ptr, w->bp.smallMap.ptr, numSmall, w->dRayBounds.ptr);
		B3_LAUNCH_CHECK();
		rayMortonKernel<<<divUp(numSmall, 256), 256, 0, s>>>(w->bp.aabbs.ptr, w->bp.smallMap.ptr, numSmall, w->dRayBounds.ptr, w->dRayKeys.ptr, w->dRayOrder.ptr);
		B3_LAUNCH_CHECK();
		B3_TRY(radixSortKV32(s, w->raySortTmp, w->dRayKeys.ptr, w->dRayOrder.ptr, numSmall, 30));
		RayTree t;
		float4* leaf = w->dRayTree.ptr;
		float4* chunk = leaf + 2 * (size_t)numSmall;
		float4* super = chunk + 2 * (size_t)numChunks;
		rayLeafKernel<<<divUp(numChunks * 32, 256), 256, 0, s>>>(w->bp.aabbs.ptr, w->dRayOrder.ptr, numSmall, leaf, chunk);
		B3_LAUNCH_CHECK();
		raySuperKernel<<<divUp(numSupers * 32, 256), 256, 0, s>>>(chunk, numChunks, super);
		B3_LAUNCH_CHECK();
		t.leaf = leaf;
		t.chunk = chunk;
		t.super = super;
		t.numLeaves = numSmall;
		t.numChunks = numChunks;
		t.numSupers = numSupers;
		t.largeMap = w->bp.largeMap.ptr;
		t.numLarge = w->bp.numLarge;
		const int blocks = std::min(divUp(numRays, 4), w->smCount * 16);
		rayTreeKernel<<<std::max(blocks, 1), 128, 0, s>>>(a, t);
		B3_LAUNCH_CHECK();
	}
	else
	{
		rayInitKernel<<<divUp(numRays, 256), 256, 0, s>>>(a);
		B3_LAUNCH_CHECK();
		const long long warpsNeeded = (long long)numRays * a.numChunks;
		const int blocks = (int)std::min<long long>((warpsNeeded + 3) / 4, (long long)w->smCount * 16);
		rayCandidatesKernel<<<std::max(blocks, 1), 128, 0, s>>>(a);
		B3_LAUNCH_CHECK();
	}
	rayResolveKernel<<<divUp(numRays, 256), 256, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemcpyAsync(hits, w->dRayHits.ptr, sizeof(b3b200_ray_hit) * (size_t)numRays, cudaMemcpyDeviceToHost, s));
	B3_CUDA_CHECK(cudaStreamSynchronize(s));
	return 0;
}

// -1 = choose per call (tree for >= 64 rays against >= 4096 small bodies), 0 = always the brute-force kernel, 1 = always the tree
extern "C" int b3b200_set_ray_accel(b3b200_world* w, int mode)
{
	if (!w || mode < -1 || mode > 1) return B3B200_ERR_INVALID;
	w->rayAccel = mode;
	return 0;
}
