// dynamics.cu -- AoS<->SoA packing, world AABB update and transform
// integration.  Memory-bound streaming kernels: every body record is moved with
// 128-bit loads/stores and each thread owns one body.
//
//   updateAabbs : setupGpuAabbsFull -> initializeGpuAabbsFull
//                 (b3GpuRigidBodyPipeline.cpp:500-560, shared/b3UpdateAabbs.h:8-33,
//                  b3TransformAabb2 shared/b3Aabb.h:24-43)        112 B / body
//   integrate   : integrateTransformsKernel / integrateSingleTransform
//                 (b3GpuRigidBodyPipeline.cpp:465-498, b3IntegrateTransforms.h:5-55)
//                 fused with the next step's AABB update           128+32 B / body
#include "internal.h"

namespace b3b200
{
constexpr int DYN_THREADS = 256;

__global__ void __launch_bounds__(DYN_THREADS) packSoAKernel(const b3b200_rigid_body* __restrict__ aos, int n, float4* __restrict__ pose,
															 float4* __restrict__ vel, int* __restrict__ coll)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4* p = reinterpret_cast<const float4*>(&aos[i]);
	float4 pos = p[0], quat = p[1], lin = p[2], ang = p[3], tail = p[4];
	pos.w = tail.y;  // invMass rides in pos.w
	pose[2 * i] = pos;
	pose[2 * i + 1] = quat;
	vel[2 * i] = lin;
	vel[2 * i + 1] = ang;
	coll[i] = __float_as_int(tail.x);
}

__global__ void __launch_bounds__(DYN_THREADS) unpackSoAKernel(b3b200_rigid_body* __restrict__ aos, int n, const float4* __restrict__ pose,
															   const float4* __restrict__ vel, const int* __restrict__ coll)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4* p = reinterpret_cast<float4*>(&aos[i]);
	float4 pos = pose[2 * i];
	float oldW = p[0].w;
	pos.w = oldW;  // the reference keeps whatever the caller stored in pos.w
	p[0] = pos;
	p[1] = pose[2 * i + 1];
	p[2] = vel[2 * i];
	p[3] = vel[2 * i + 1];
	// collidable index and inverse mass only change through the halo path (ghost slots); keep the AoS view in step
	aos[i].collidableIdx = coll[i];
	aos[i].invMass = pose[2 * i].w;
}

B3_D void computeWorldAabb(const float4& pos, const float4& quat, const float4& lmn, const float4& lmx, float4& outMin, float4& outMax)
{
	// b3TransformAabb2 with margin 0 (shared/b3Aabb.h:24-43)
	float4 half = mk4(0.5f * (lmx.x - lmn.x), 0.5f * (lmx.y - lmn.y), 0.5f * (lmx.z - lmn.z));
	half = mk4(half.x + 0.f, half.y + 0.f, half.z + 0.f);
	float4 lc = mk4(0.5f * (lmx.x + lmn.x), 0.5f * (lmx.y + lmn.y), 0.5f * (lmx.z + lmn.z));
	Mat3 m = matFromQuat(quat);
	float4 a0 = mk4(fabsf(m.r0.x), fabsf(m.r0.y), fabsf(m.r0.z));
	float4 a1 = mk4(fabsf(m.r1.x), fabsf(m.r1.y), fabsf(m.r1.z));
	float4 a2 = mk4(fabsf(m.r2.x), fabsf(m.r2.y), fabsf(m.r2.z));
	float4 r = matMulVec(m, lc);
	float4 c = mk4(r.x + pos.x, r.y + pos.y, r.z + pos.z);
	float4 e = mk4(dot3(half, a0), dot3(half, a1), dot3(half, a2));
	outMin = mk4(c.x - e.x, c.y - e.y, c.z - e.z);
	outMax = mk4(c.x + e.x, c.y + e.y, c.z + e.z);
}

B3_D void writeAabb(b3b200_aabb* __restrict__ aabbs, int i, float4 mn, float4 mx, float invMass)
{
	mn.w = __int_as_float(i);
	mx.w = __int_as_float(invMass == 0.f ? 0 : 1);
	float4* q = reinterpret_cast<float4*>(&aabbs[i]);
	q[0] = mn;
	q[1] = mx;
}

__global__ void __launch_bounds__(DYN_THREADS) updateAabbsKernel(const float4* __restrict__ pose, const int* __restrict__ coll, int n,
																 const b3b200_collidable* __restrict__ collidables, const b3b200_aabb* __restrict__ localAabbs,
																 b3b200_aabb* __restrict__ aabbs)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 pos = pose[2 * i], quat = pose[2 * i + 1];
	int c = coll[i];
	if (c < 0) return;
	if (__ldg(&collidables[c].shapeIndex) < 0) return;
	const float4* l = reinterpret_cast<const float4*>(&localAabbs[c]);
	float4 mn, mx;
	computeWorldAabb(pos, quat, __ldg(l), __ldg(l + 1), mn, mx);
	writeAabb(aabbs, i, mn, mx, pos.w);
}

template <bool WITH_AABB>
__global__ void __launch_bounds__(DYN_THREADS) integrateKernel(float4* __restrict__ pose, float4* __restrict__ vel, const int* __restrict__ coll, int n,
															   float dt, float angularDamping, float gx, float gy, float gz,
															   const b3b200_collidable* __restrict__ collidables, const b3b200_aabb* __restrict__ localAabbs,
															   b3b200_aabb* __restrict__ aabbs)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 pos = pose[2 * i], quat = pose[2 * i + 1];
	const float invMass = pos.w;
	if (invMass != 0.f)
	{
		float4 lin = vel[2 * i], ang = vel[2 * i + 1];
		const float THRESH = (0.25f * 3.14159254f);
		ang.x *= angularDamping;
		ang.y *= angularDamping;
		ang.z *= angularDamping;
		float fAngle = sqrtf(dot3(ang, ang));
		if (fAngle * dt > THRESH) fAngle = THRESH / dt;
		float k;
		if (fAngle < 0.001f)
			k = 0.5f * dt - (dt * dt * dt) * 0.020833333333f * fAngle * fAngle;
		else
			k = sinf(0.5f * fAngle * dt) / fAngle;
		float4 dorn = mk4(ang.x * k, ang.y * k, ang.z * k, cosf(fAngle * dt * 0.5f));
		float4 pq = quatMul(dorn, quat);
		float len2 = pq.x * pq.x + pq.y * pq.y + pq.z * pq.z + pq.w * pq.w;
		float s = 1.0f / sqrtf(len2);
		quat = mk4(pq.x * s, pq.y * s, pq.z * s, pq.w * s);
		pos.x += lin.x * dt;
		pos.y += lin.y * dt;
		pos.z += lin.z * dt;
		lin.x += gx * dt;
		lin.y += gy * dt;
		lin.z += gz * dt;
		pose[2 * i] = pos;
		pose[2 * i + 1] = quat;
		vel[2 * i] = lin;
		vel[2 * i + 1] = ang;
	}
	if (WITH_AABB)
	{
		int c = coll[i];
		if (c >= 0 && __ldg(&collidables[c].shapeIndex) >= 0)
		{
			const float4* l = reinterpret_cast<const float4*>(&localAabbs[c]);
			float4 mn, mx;
			computeWorldAabb(pos, quat, __ldg(l), __ldg(l + 1), mn, mx);
			writeAabb(aabbs, i, mn, mx, invMass);
		}
	}
}

int launchPackSoA(World* w)
{
	int n = w->numBodies;
	if (n == 0) return 0;
	packSoAKernel<<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(w->dBodiesAoS.ptr, n, w->dPose.ptr, w->dVel.ptr, w->dCollidableIdx.ptr);
	B3_LAUNCH_CHECK();
	w->soaDirty = false;
	w->partValid = false;  // poses (and inverse masses) were replaced: the solver re-partitions the bodies
	return 0;
}

// pipelined host stepping (world.cu, b3b200_step_host_async): body records come from / go to a staging buffer instead of the
// world's own AoS array
int launchPackSoAFrom(World* w, const b3b200_rigid_body* src)
{
	int n = w->numBodies;
	if (n == 0) return 0;
	packSoAKernel<<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(src, n, w->dPose.ptr, w->dVel.ptr, w->dCollidableIdx.ptr);
	B3_LAUNCH_CHECK();
	w->soaDirty = true;    // the world's own AoS array does not hold this state
	w->partValid = false;  // poses (and inverse masses) were replaced: the solver re-partitions the bodies
	w->aabbsValid = false;
	return 0;
}
int launchUnpackSoATo(World* w, b3b200_rigid_body* dst)
{
	int n = w->numBodies;
	if (n == 0) return 0;
	unpackSoAKernel<<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(dst, n, w->dPose.ptr, w->dVel.ptr, w->dCollidableIdx.ptr);
	B3_LAUNCH_CHECK();
	return 0;
}

int launchUnpackSoA(World* w)
{
	int n = w->numBodies;
	if (n == 0) return 0;
	unpackSoAKernel<<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(w->dBodiesAoS.ptr, n, w->dPose.ptr, w->dVel.ptr, w->dCollidableIdx.ptr);
	B3_LAUNCH_CHECK();
	w->soaDirty = false;
	return 0;
}

int launchUpdateAabbs(World* w)
{
	int n = w->numBodies;
	if (n == 0) return 0;
	updateAabbsKernel<<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(w->dPose.ptr, w->dCollidableIdx.ptr, n, w->dCollidables.ptr, w->dLocalAabbs.ptr, w->bp.aabbs.ptr);
	B3_LAUNCH_CHECK();
	w->aabbsValid = true;
	return 0;
}

int launchIntegrate(World* w, float dt, bool alsoAabbs)
{
	int n = w->numBodies;
	if (n == 0) return 0;
	if (alsoAabbs)
		integrateKernel<true><<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(w->dPose.ptr, w->dVel.ptr, w->dCollidableIdx.ptr, n, dt, w->angularDamping,
																					w->gravity[0], w->gravity[1], w->gravity[2], w->dCollidables.ptr, w->dLocalAabbs.ptr, w->bp.aabbs.ptr);
	else
		integrateKernel<false><<<divUp(n, DYN_THREADS), DYN_THREADS, 0, w->stream>>>(w->dPose.ptr, w->dVel.ptr, w->dCollidableIdx.ptr, n, dt, w->angularDamping,
																					 w->gravity[0], w->gravity[1], w->gravity[2], w->dCollidables.ptr, w->dLocalAabbs.ptr, w->bp.aabbs.ptr);
	B3_LAUNCH_CHECK();
	w->soaDirty = true;
	w->aabbsValid = alsoAabbs;
	return 0;
}

}  // namespace b3b200
