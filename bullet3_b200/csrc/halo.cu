// halo.cu -- ghost-body halo records for the spatial slab decomposition of one large scene over several
// GPUs (no counterpart in the reference, which is single-device).  A rank owns the bodies
// registered first ([0, numOwned)); the trailing bodies of its world are GHOST SLOTS that mirror the
// neighbours' boundary bodies.  Every step, after integration:
//   b3b200_halo_pack    selects the owned dynamic bodies whose world AABB reaches into [lo, hi] along `axis`
//                       and writes one 176-byte record per body into a caller-supplied DEVICE buffer
//                       (pose, velocity, both inverse inertias, collidable index, global id)
//   (the caller moves the buffer to the neighbour: NCCL send/recv over NVLink, bullet3_b200/slab.py)
//   b3b200_halo_unpack  scatters received records into the ghost slots and parks the unused slots
// Both run on the world's stream; pack returns the record count to the host (one 4-byte read-back).
// Migration (a body whose centre has left the slab changes owner) uses the same records:
//   b3b200_halo_emigrate  selects the owned dynamic bodies whose centre along `axis` lies in [lo, hi] (a range OUTSIDE the
//                         slab), writes their records, frees their slots (parked, static, id -1) and returns the slot list
//   b3b200_halo_adopt     writes received records into the free owned slots the caller names
// Every slot carries its global body id (b3b200_halo_set_ids at start-up; pack / emigrate send it, unpack / adopt store it).
#include "internal.h"

namespace b3b200
{
struct HaloRecord
{
	float4 pos;  // w = invMass
	float4 quat;
	float4 linVel;
	float4 angVel;
	float4 invInertiaWorld[3];
	float4 initInvInertia[3];
	int4 meta;  // x = collidable index, y = global body id, z = owner rank
};
static_assert(sizeof(HaloRecord) == 176, "halo record layout");

__global__ void haloPackKernel(const float4* __restrict__ pose, const float4* __restrict__ vel, const b3b200_inertia* __restrict__ inertias,
							   const int* __restrict__ coll, const b3b200_aabb* __restrict__ aabbs, int numOwned, int axis, float lo, float hi,
							   int globalIdBase, const int* __restrict__ ids, int rank, HaloRecord* __restrict__ dst, int capacity, unsigned int* __restrict__ count)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	bool take = false;
	if (i < numOwned)
	{
		const float invMass = pose[2 * i].w;
		const float mn = aabbs[i].min[axis], mx = aabbs[i].max[axis];
		take = invMass != 0.f && mx >= lo && mn <= hi;
	}
	const unsigned int m = __ballot_sync(0xffffffffu, take);
	if (!m) return;
	const int lane = threadIdx.x & 31;
	unsigned int base = 0;
	if (lane == 0) base = atomicAdd(count, (unsigned int)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (!take) return;
	const unsigned int slot = base + __popc(m & ((1u << lane) - 1u));
	if (slot >= (unsigned int)capacity) return;
	HaloRecord r;
	r.pos = pose[2 * i];
	r.quat = pose[2 * i + 1];
	r.linVel = vel[2 * i];
	r.angVel = vel[2 * i + 1];
	const float4* I = reinterpret_cast<const float4*>(&inertias[i]);
	for (int k = 0; k < 3; k++)
	{
		r.invInertiaWorld[k] = I[k];
		r.initInvInertia[k] = I[3 + k];
	}
	r.meta = make_int4(coll[i], ids ? ids[i] : globalIdBase + i, rank, 0);
	dst[slot] = r;
}

__global__ void haloUnpackKernel(float4* __restrict__ pose, float4* __restrict__ vel, b3b200_inertia* __restrict__ inertias, int* __restrict__ coll,
								 int* __restrict__ ghostGlobalId, const HaloRecord* __restrict__ src, int count, int firstSlot, int numSlots)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= numSlots) return;
	const int i = firstSlot + k;
	if (k < count)
	{
		const HaloRecord r = src[k];
		pose[2 * i] = r.pos;
		pose[2 * i + 1] = r.quat;
		vel[2 * i] = r.linVel;
		vel[2 * i + 1] = r.angVel;
		float4* I = reinterpret_cast<float4*>(&inertias[i]);
		for (int j = 0; j < 3; j++)
		{
			I[j] = r.invInertiaWorld[j];
			I[3 + j] = r.initInvInertia[j];
		}
		coll[i] = r.meta.x;
		ghostGlobalId[i] = r.meta.y;
	}
	else
	{
		// parked: static, far away, one slot per 1024 units so that parked ghosts never overlap anything
		// (indexed by the absolute body slot: the parked slots of the left and of the right neighbour must not coincide)
		pose[2 * i] = mk4(1.0e6f + 1024.0f * (float)i, -1.0e6f, 1.0e6f, 0.f);
		pose[2 * i + 1] = mk4(0.f, 0.f, 0.f, 1.f);
		vel[2 * i] = mk4(0, 0, 0, 0);
		vel[2 * i + 1] = mk4(0, 0, 0, 0);
		ghostGlobalId[i] = -1;
	}
}

B3_D void parkSlot(float4* __restrict__ pose, float4* __restrict__ vel, int* __restrict__ ids, int i)
{
	// parked: static, far away, one slot per 1024 units so that parked slots never overlap anything
	// (indexed by the absolute body slot: the parked slots of the left and of the right neighbour must not coincide)
	pose[2 * i] = mk4(1.0e6f + 1024.0f * (float)i, -1.0e6f, 1.0e6f, 0.f);
	pose[2 * i + 1] = mk4(0.f, 0.f, 0.f, 1.f);
	vel[2 * i] = mk4(0, 0, 0, 0);
	vel[2 * i + 1] = mk4(0, 0, 0, 0);
	ids[i] = -1;
}

__global__ void haloEmigrateKernel(float4* __restrict__ pose, float4* __restrict__ vel, const b3b200_inertia* __restrict__ inertias, const int* __restrict__ coll,
								   int* __restrict__ ids, int numOwned, int axis, float lo, float hi, int rank, HaloRecord* __restrict__ dst,
								   int* __restrict__ slots, int capacity, unsigned int* __restrict__ count)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	bool take = false;
	if (i < numOwned)
	{
		const float4 p = pose[2 * i];
		const float c = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
		take = p.w != 0.f && ids[i] >= 0 && c >= lo && c <= hi;
	}
	const unsigned int m = __ballot_sync(0xffffffffu, take);
	if (!m) return;
	const int lane = threadIdx.x & 31;
	unsigned int base = 0;
	if (lane == 0) base = atomicAdd(count, (unsigned int)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (!take) return;
	const unsigned int slot = base + __popc(m & ((1u << lane) - 1u));
	if (slot >= (unsigned int)capacity) return;  // stays here this time; the host reports the overflow
	HaloRecord r;
	r.pos = pose[2 * i];
	r.quat = pose[2 * i + 1];
	r.linVel = vel[2 * i];
	r.angVel = vel[2 * i + 1];
	const float4* I = reinterpret_cast<const float4*>(&inertias[i]);
	for (int k = 0; k < 3; k++)
	{
		r.invInertiaWorld[k] = I[k];
		r.initInvInertia[k] = I[3 + k];
	}
	r.meta = make_int4(coll[i], ids[i], rank, 0);
	dst[slot] = r;
	slots[slot] = i;
	parkSlot(pose, vel, ids, i);
}

__global__ void haloAdoptKernel(float4* __restrict__ pose, float4* __restrict__ vel, b3b200_inertia* __restrict__ inertias, int* __restrict__ coll,
								int* __restrict__ ids, const HaloRecord* __restrict__ src, const int* __restrict__ slots, int count)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= count) return;
	const int i = slots[k];
	const HaloRecord r = src[k];
	pose[2 * i] = r.pos;
	pose[2 * i + 1] = r.quat;
	vel[2 * i] = r.linVel;
	vel[2 * i + 1] = r.angVel;
	float4* I = reinterpret_cast<float4*>(&inertias[i]);
	for (int j = 0; j < 3; j++)
	{
		I[j] = r.invInertiaWorld[j];
		I[3 + j] = r.initInvInertia[j];
	}
	coll[i] = r.meta.x;
	ids[i] = r.meta.y;
}

}  // namespace b3b200

using namespace b3b200;

extern "C" int b3b200_halo_record_size(void) { return (int)sizeof(HaloRecord); }

extern "C" int b3b200_halo_pack(b3b200_world* w, int axis, float lo, float hi, int numOwned, int globalIdBase, int rank, void* dstDevice, int capacity,
								int* countOut)
{
	if (!w || w->device < 0 || !w->uploaded || axis < 0 || axis > 2 || numOwned < 0 || numOwned > w->numBodies || !dstDevice || capacity < 0 || !countOut)
		return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	if (!w->aabbsValid) B3_TRY(launchUpdateAabbs(w));
	unsigned int* ctr = &w->dCounters.ptr[CTR_HALO];
	B3_CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), w->stream));
	if (numOwned > 0)
	{
		haloPackKernel<<<divUp(numOwned, 256), 256, 0, w->stream>>>(w->dPose.ptr, w->dVel.ptr, w->dInertias.ptr, w->dCollidableIdx.ptr, w->bp.aabbs.ptr, numOwned, axis, lo,
																	hi, globalIdBase, w->haloIdsSet ? w->dGhostGlobalId.ptr : nullptr, rank, (HaloRecord*)dstDevice, capacity, ctr);
		B3_LAUNCH_CHECK();
	}
	unsigned int n = 0;
	B3_CUDA_CHECK(cudaMemcpyAsync(&n, ctr, sizeof(n), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	if ((int)n > capacity)
	{
		setLastError("halo_pack: %u boundary bodies exceed the halo capacity %d", n, capacity);
		*countOut = capacity;
		return B3B200_ERR_CAPACITY;
	}
	*countOut = (int)n;
	return 0;
}

extern "C" int b3b200_halo_unpack(b3b200_world* w, const void* srcDevice, int count, int firstGhostSlot, int numGhostSlots)
{
	if (!w || w->device < 0 || !w->uploaded || count < 0 || firstGhostSlot < 0 || numGhostSlots < 0 || firstGhostSlot + numGhostSlots > w->numBodies ||
		count > numGhostSlots || (count > 0 && !srcDevice))
		return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	B3_TRY(w->dGhostGlobalId.reserve(std::max(w->numBodies, 1)));
	if (numGhostSlots > 0)
	{
		haloUnpackKernel<<<divUp(numGhostSlots, 256), 256, 0, w->stream>>>(w->dPose.ptr, w->dVel.ptr, w->dInertias.ptr, w->dCollidableIdx.ptr, w->dGhostGlobalId.ptr,
																		   (const HaloRecord*)srcDevice, count, firstGhostSlot, numGhostSlots);
		B3_LAUNCH_CHECK();
	}
	w->aabbsValid = false;
	w->soaDirty = true;
	w->partValid = false;
	return 0;
}

extern "C" int b3b200_halo_ghost_ids(b3b200_world* w, int* dst, int n)
{
	if (!w || w->device < 0 || !dst || n < 0 || n > w->numBodies) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	B3_TRY(w->dGhostGlobalId.reserve(std::max(w->numBodies, 1)));
	if (n) B3_CUDA_CHECK(cudaMemcpyAsync(dst, w->dGhostGlobalId.ptr, sizeof(int) * n, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}

// global id of every body slot (-1 = none); from then on pack sends these instead of globalIdBase + slot
extern "C" int b3b200_halo_set_ids(b3b200_world* w, const int* ids, int n)
{
	if (!w || w->device < 0 || !w->uploaded || !ids || n != w->numBodies) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	B3_TRY(w->dGhostGlobalId.reserve(std::max(w->numBodies, 1)));
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dGhostGlobalId.ptr, ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	w->haloIdsSet = true;
	return 0;
}

extern "C" int b3b200_halo_emigrate(b3b200_world* w, int axis, float lo, float hi, int numOwned, int rank, void* dstDevice, int capacity, int* slotsOut,
									int* countOut)
{
	if (!w || w->device < 0 || !w->uploaded || !w->haloIdsSet || axis < 0 || axis > 2 || numOwned < 0 || numOwned > w->numBodies || !dstDevice || capacity < 0 ||
		!slotsOut || !countOut)
		return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	B3_TRY(w->dHaloSlots.reserve((size_t)std::max(capacity, 1)));
	unsigned int* ctr = &w->dCounters.ptr[CTR_HALO];
	B3_CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), w->stream));
	if (numOwned > 0)
	{
		haloEmigrateKernel<<<divUp(numOwned, 256), 256, 0, w->stream>>>(w->dPose.ptr, w->dVel.ptr, w->dInertias.ptr, w->dCollidableIdx.ptr, w->dGhostGlobalId.ptr,
																		numOwned, axis, lo, hi, rank, (HaloRecord*)dstDevice, w->dHaloSlots.ptr, capacity, ctr);
		B3_LAUNCH_CHECK();
	}
	unsigned int n = 0;
	B3_CUDA_CHECK(cudaMemcpyAsync(&n, ctr, sizeof(n), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	const int moved = (int)n < capacity ? (int)n : capacity;
	if (moved) B3_CUDA_CHECK(cudaMemcpyAsync(slotsOut, w->dHaloSlots.ptr, sizeof(int) * (size_t)moved, cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	*countOut = moved;
	if (moved)
	{
		w->aabbsValid = false;
		w->soaDirty = true;
		w->partValid = false;
	}
	if ((int)n > capacity)
	{
		setLastError("halo_emigrate: %u bodies left the slab, only %d fit the buffer (the rest move next time)", n, capacity);
		return B3B200_ERR_CAPACITY;
	}
	return 0;
}

extern "C" int b3b200_halo_adopt(b3b200_world* w, const void* srcDevice, int count, const int* slots)
{
	if (!w || w->device < 0 || !w->uploaded || !w->haloIdsSet || count < 0 || (count > 0 && (!srcDevice || !slots))) return B3B200_ERR_INVALID;
	if (count == 0) return 0;
	for (int k = 0; k < count; k++)
		if (slots[k] < 0 || slots[k] >= w->numBodies) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	B3_TRY(w->dHaloSlots.reserve((size_t)count));
	B3_CUDA_CHECK(cudaMemcpyAsync(w->dHaloSlots.ptr, slots, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, w->stream));
	haloAdoptKernel<<<divUp(count, 256), 256, 0, w->stream>>>(w->dPose.ptr, w->dVel.ptr, w->dInertias.ptr, w->dCollidableIdx.ptr, w->dGhostGlobalId.ptr,
															  (const HaloRecord*)srcDevice, w->dHaloSlots.ptr, count);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));  // `slots` is the caller's host memory
	w->aabbsValid = false;
	w->soaDirty = true;
	w->partValid = false;
	return 0;
}
