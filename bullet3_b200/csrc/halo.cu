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
#include <dlfcn.h>
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

// ---------------------------------------------------------------- slab step behind the C ABI: NCCL called from C++
// The message of one side is a fixed-capacity buffer [header record | capacity records]; the header's meta.x is the record
// count, written on the device.  So neither side ever needs the count on the host: no read-back, no stream synchronisation;
// pack, ncclSend / ncclRecv and unpack are all queued on the world's stream behind the step's kernels.
__global__ void haloHeaderKernel(HaloRecord* __restrict__ dst, const unsigned int* __restrict__ count, int capacity, unsigned int* __restrict__ overflow)
{
	const unsigned int n = *count;
	HaloRecord h;
	memset(&h, 0, sizeof(h));
	h.meta = make_int4((int)(n < (unsigned int)capacity ? n : (unsigned int)capacity), (int)n, 0, 0);
	dst[0] = h;
	if (n > (unsigned int)capacity) atomicOr(overflow, (unsigned int)OVF_HALO);
}
__global__ void haloUnpackMsgKernel(float4* __restrict__ pose, float4* __restrict__ vel, b3b200_inertia* __restrict__ inertias, int* __restrict__ coll,
									int* __restrict__ ghostGlobalId, const HaloRecord* __restrict__ msg, int firstSlot, int numSlots)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= numSlots) return;
	const int i = firstSlot + k;
	int count = msg[0].meta.x;
	if (count > numSlots) count = numSlots;
	if (k < count)
	{
		const HaloRecord r = msg[1 + k];
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
		parkSlot(pose, vel, ghostGlobalId, i);
}

// NCCL through dlopen: the library has no link-time dependency on it (a single-GPU user never needs it), and inside a
// process that already holds an NCCL (torch's) the same copy is used.  Only the handful of entry points of the exchange.
struct NcclApi
{
	void* lib = nullptr;
	int (*GetUniqueId)(void*) = nullptr;
	int (*CommInitRank)(void**, int, b3b200_nccl_id, int) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int loadNccl()
{
	if (g_nccl.lib) return 0;
	void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h)
	{
		setLastError("slab mode needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
		return B3B200_ERR_STATE;
	}
	NcclApi a;
	a.lib = h;
	a.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
	a.CommInitRank = (int (*)(void**, int, b3b200_nccl_id, int))dlsym(h, "ncclCommInitRank");
	a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
	a.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
	a.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
	a.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
	a.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
	a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
	if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.GroupStart || !a.GroupEnd || !a.Send || !a.Recv)
	{
		setLastError("slab mode: the NCCL library lacks a required entry point");
		return B3B200_ERR_STATE;
	}
	g_nccl = a;
	return 0;
}
#define B3_NCCL_CHECK(expr)                                                                                              \
	do                                                                                                                   \
	{                                                                                                                    \
		int _r = (expr);                                                                                                 \
		if (_r != 0)                                                                                                     \
		{                                                                                                                \
			setLastError("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");     \
			return B3B200_ERR_CUDA;                                                                                      \
		}                                                                                                                \
	} while (0)

void slabDestroy(World* w)
{
	if (w->slab.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(w->slab.comm);
	w->slab.comm = nullptr;
	w->slab.active = false;
}

// pack -> exchange -> unpack, all queued on the world's stream
int slabExchange(World* w)
{
	SlabState& sl = w->slab;
	cudaStream_t st = w->stream;
	if (!w->aabbsValid) B3_TRY(launchUpdateAabbs(w));
	const size_t msgBytes = sizeof(HaloRecord) * (size_t)(sl.maxGhosts + 1);
	const float big = 3.0e38f;
	for (int side = 0; side < 2; side++)
	{
		if (!(side == 0 ? sl.hasLeft : sl.hasRight)) continue;
		unsigned int* ctr = &w->dCounters.ptr[side == 0 ? CTR_HALO : CTR_HALO_RIGHT];
		B3_CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
		HaloRecord* msg = reinterpret_cast<HaloRecord*>(sl.sendBuf[side].ptr);
		const float a = side == 0 ? -big : sl.hi - sl.margin, b = side == 0 ? sl.lo + sl.margin : big;
		if (sl.numOwned > 0)
		{
			haloPackKernel<<<divUp(sl.numOwned, 256), 256, 0, st>>>(w->dPose.ptr, w->dVel.ptr, w->dInertias.ptr, w->dCollidableIdx.ptr, w->bp.aabbs.ptr, sl.numOwned,
																   sl.axis, a, b, sl.globalIdBase, w->haloIdsSet ? w->dGhostGlobalId.ptr : nullptr, sl.rank, msg + 1,
																   sl.maxGhosts, ctr);
			B3_LAUNCH_CHECK();
		}
		haloHeaderKernel<<<1, 1, 0, st>>>(msg, ctr, sl.maxGhosts, &w->dCounters.ptr[CTR_OVERFLOW]);
		B3_LAUNCH_CHECK();
	}
	B3_NCCL_CHECK(g_nccl.GroupStart());
	for (int side = 0; side < 2; side++)
	{
		if (!(side == 0 ? sl.hasLeft : sl.hasRight)) continue;
		const int peer = side == 0 ? sl.rank - 1 : sl.rank + 1;
		B3_NCCL_CHECK(g_nccl.Send(sl.sendBuf[side].ptr, msgBytes, 0 /* ncclInt8 */, peer, sl.comm, st));
		B3_NCCL_CHECK(g_nccl.Recv(sl.recvBuf[side].ptr, msgBytes, 0, peer, sl.comm, st));
	}
	B3_NCCL_CHECK(g_nccl.GroupEnd());
	B3_TRY(w->dGhostGlobalId.reserve(std::max(w->numBodies, 1)));
	int slot = sl.firstGhostSlot;
	for (int side = 0; side < 2; side++)
	{
		if (!(side == 0 ? sl.hasLeft : sl.hasRight)) continue;
		haloUnpackMsgKernel<<<divUp(sl.maxGhosts, 256), 256, 0, st>>>(w->dPose.ptr, w->dVel.ptr, w->dInertias.ptr, w->dCollidableIdx.ptr, w->dGhostGlobalId.ptr,
																	   reinterpret_cast<const HaloRecord*>(sl.recvBuf[side].ptr), slot, sl.maxGhosts);
		B3_LAUNCH_CHECK();
		slot += sl.maxGhosts;
	}
	w->aabbsValid = false;
	w->soaDirty = true;
	w->partValid = false;
	return 0;
}

}  // namespace b3b200

using namespace b3b200;

// ---- slab decomposition driven from C / C++: the caller creates one world per rank (owned bodies first, then
// ghost slots), hands every rank the same NCCL unique id, and calls b3b200_slab_step instead of b3b200_step.
extern "C" int b3b200_slab_unique_id(b3b200_nccl_id* out)
{
	if (!out) return B3B200_ERR_INVALID;
	B3_TRY(loadNccl());
	B3_NCCL_CHECK(g_nccl.GetUniqueId(out));
	return 0;
}
extern "C" int b3b200_slab_init(b3b200_world* w, const b3b200_slab_config* cfg, const b3b200_nccl_id* id)
{
	if (w) w->dropStepGraphs();  // may change what a step launches
	if (!w || w->device < 0 || !w->uploaded || !cfg || !id) return B3B200_ERR_INVALID;
	if (cfg->axis < 0 || cfg->axis > 2 || cfg->rank < 0 || cfg->rank >= cfg->numRanks || cfg->maxGhosts < 0 || cfg->numOwned < 0 || cfg->numOwned > w->numBodies ||
		cfg->firstGhostSlot < 0)
		return B3B200_ERR_INVALID;
	if (w->numWorlds > 1)
	{
		setLastError("slab_init: a world that batches independent worlds (b3b200_set_current_world) cannot be slab-decomposed; shard the worlds across ranks instead");
		return B3B200_ERR_STATE;
	}
	const bool hasLeft = cfg->rank > 0, hasRight = cfg->rank < cfg->numRanks - 1;
	if (cfg->firstGhostSlot + cfg->maxGhosts * ((int)hasLeft + (int)hasRight) > w->numBodies)
	{
		setLastError("slab_init: the ghost slots [%d, +%d x %d) do not fit the world's %d bodies", cfg->firstGhostSlot, (int)hasLeft + (int)hasRight, cfg->maxGhosts, w->numBodies);
		return B3B200_ERR_INVALID;
	}
	B3_TRY(loadNccl());
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	slabDestroy(w);
	SlabState& sl = w->slab;
	sl.axis = cfg->axis;
	sl.lo = cfg->lo;
	sl.hi = cfg->hi;
	sl.margin = cfg->margin;
	sl.numOwned = cfg->numOwned;
	sl.firstGhostSlot = cfg->firstGhostSlot;
	sl.maxGhosts = cfg->maxGhosts;
	sl.globalIdBase = cfg->globalIdBase;
	sl.rank = cfg->rank;
	sl.numRanks = cfg->numRanks;
	sl.hasLeft = hasLeft;
	sl.hasRight = hasRight;
	const size_t msgBytes = sizeof(HaloRecord) * (size_t)(sl.maxGhosts + 1);
	for (int side = 0; side < 2; side++)
	{
		B3_TRY(sl.sendBuf[side].reserve(msgBytes));
		B3_TRY(sl.recvBuf[side].reserve(msgBytes));
		B3_CUDA_CHECK(cudaMemsetAsync(sl.sendBuf[side].ptr, 0, msgBytes, w->stream));
		B3_CUDA_CHECK(cudaMemsetAsync(sl.recvBuf[side].ptr, 0, msgBytes, w->stream));
	}
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	B3_NCCL_CHECK(g_nccl.CommInitRank(&sl.comm, sl.numRanks, *id, sl.rank));
	sl.active = true;
	return 0;
}
// ghost slots <- the neighbours' current boundary bands (call once before the first step)
extern "C" int b3b200_slab_exchange(b3b200_world* w)
{
	if (!w || w->device < 0 || !w->uploaded || !w->slab.active) return B3B200_ERR_STATE;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	return slabExchange(w);
}
// halo records sent by this rank in the last exchange (left, right); synchronises the stream
extern "C" int b3b200_slab_last_counts(b3b200_world* w, int* left, int* right)
{
	if (!w || w->device < 0 || !w->slab.active || !left || !right) return B3B200_ERR_STATE;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	unsigned int c[2] = {0, 0};
	B3_CUDA_CHECK(cudaMemcpyAsync(&c[0], &w->dCounters.ptr[CTR_HALO], sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaMemcpyAsync(&c[1], &w->dCounters.ptr[CTR_HALO_RIGHT], sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	*left = w->slab.hasLeft ? (int)c[0] : 0;
	*right = w->slab.hasRight ? (int)c[1] : 0;
	return 0;
}
extern "C" int b3b200_slab_shutdown(b3b200_world* w)
{
	if (w) w->dropStepGraphs();  // may change what a step launches
	if (!w) return B3B200_ERR_INVALID;
	if (w->device >= 0) cudaSetDevice(w->device);
	slabDestroy(w);
	return 0;
}

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
	if (w) w->dropStepGraphs();  // may change what a step launches
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
	if (w) w->dropStepGraphs();  // may change what a step launches
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
	if (w) w->dropStepGraphs();  // may change what a step launches
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
	if (w) w->dropStepGraphs();  // may change what a step launches
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
