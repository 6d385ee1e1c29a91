// raycast.cu -- b3GpuRigidBodyPipeline::castRays / b3GpuRaycast::castRays
// (src/Bullet3OpenCL/RigidBody/b3GpuRigidBodyPipeline.cpp:671-680, src/Bullet3OpenCL/Raycast/b3GpuRaycast.cpp:249-373).
// Result semantics are those of the reference's host twin castRaysHost (:160-246): closest hit, ties to the lower body
// index, hull normal in the hull's local frame.  The reference walks the bodies sequentially per ray (host twin) or
// builds a PLBVH and sorts ray-body pairs (GPU path); here
//   rayCandidatesKernel  one warp per (ray, chunk of 2048 bodies): lanes cull by a segment-vs-world-AABB slab test,
//                        run the exact body test on the survivors and merge the best (fraction, body) of the chunk
//                        into a 64-bit per-ray key with atomicMin (non-negative float bits order like the floats)
//   rayResolveKernel     one thread per ray: recomputes the winning body's hit and writes the b3RayHit
// i.e. O(rays x bodies) AABB tests (32 B each, L2-resident): meant for picking / sensor rays, not for ray tracing.
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

// conservative: does the segment from + t (to - from), t in [0, 1], touch the box?
B3_D bool segmentHitsAabb(const float4& from, const float4& to, const float4& mn, const float4& mx)
{
	float t0 = 0.f, t1 = 1.f;
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
			if (!segmentHitsAabb(from, to, mn, mx)) continue;
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
	rayInitKernel<<<divUp(numRays, 256), 256, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	const long long warpsNeeded = (long long)numRays * a.numChunks;
	const int blocks = (int)std::min<long long>((warpsNeeded + 3) / 4, (long long)w->smCount * 16);
	rayCandidatesKernel<<<std::max(blocks, 1), 128, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	rayResolveKernel<<<divUp(numRays, 256), 256, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemcpyAsync(hits, w->dRayHits.ptr, sizeof(b3b200_ray_hit) * (size_t)numRays, cudaMemcpyDeviceToHost, s));
	B3_CUDA_CHECK(cudaStreamSynchronize(s));
	return 0;
}
