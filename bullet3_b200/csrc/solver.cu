// solver.cu -- two-level batched PGS contact solver.
//
// Replaces b3GpuPgsContactSolver::solveContacts (b3GpuPgsContactSolver.cpp:568-1103): 6 radix sorts, contacts binned
// into 8 x 4 x 8 wrapped spatial cells, a serial batching thread per cell (batchingKernelsNew.cl:144-231), contact ->
// constraint, then 2 * I * 8 launches in which one work-group walks the batches of one cell between local barriers
// (solveContact.cl:397-482) with the body velocities in global memory.
//
// Same two-level idea, rebuilt around what a B200 SM can hold:
//
//  partition  (every few steps) the dynamic bodies are sorted along a Morton curve and cut into BLOCKS of S <= 2176
//             bodies -- one block per CTA of the persistent iteration kernel, whose 96 bytes of solver state per body
//             (velocities, position + inverse mass, world inverse inertia) live in that CTA's shared memory for the
//             whole solve.
//  setup      every contact is INTERIOR to a block (both dynamic bodies in it, or one dynamic + one static body) or
//             CROSS (two blocks).  Interior contacts are coloured per block inside shared memory, cross contacts
//             globally with one 128-bit colour mask per body; rows (setConstraint4, b3ConvertConstraint4.h:62-148)
//             are written as 32-row structure-of-arrays tiles in solve order, so a warp streams a tile with 512-byte
//             coalesced loads: 128 B per row and pass in the normal phase, 80 B in the friction phase, against 2 x 192 B
//             for the 176-byte b3ContactConstraint4 rows.
//  iterate    ONE persistent cooperative kernel.  Per iteration: the cross colours (global velocities through L2, a
//             grid barrier per colour -- 5-10 of them), then every CTA runs all the colours of its block between
//             __syncthreads() with the velocities in shared memory.  The batch index a contact reports (b3Contact4::
//             m_batchIdx) is its position in that order: cross colour k -> k, interior colour m -> Kc + m (interior
//             colours of different blocks share no dynamic body, so together they form one valid batch); solving
//             the batches one after the other on the CPU gives the same Gauss-Seidel order ("same batching").
//             Order of the phases = the reference's: all iterations of the normal rows, then all iterations of the
//             friction rows (solveContact<false> / solveFriction, b3Solver.cpp:187-329).
#include "internal.h"

namespace b3b200
{
// threads of the iteration kernels, per phase (normal rows: 44 registers of row data per tile in flight; friction: 20):
// as many warps as keep the row solve free of spills -- more warps = fewer tiles per warp and colour
constexpr int ITER_THREADS_NORMAL = 256, ITER_THREADS_FRICTION = 256;
constexpr int SETUP_THREADS = 512;
constexpr int S_MAX = 2688;       // dynamic bodies per block: (S_MAX + NSTATIC) * (80 + 2) B = 220 KB of shared memory
constexpr int S_MIN = 1024;       // below this a block is not worth a grid barrier
constexpr int NSTATIC = 64;       // static bodies a block can keep in its own slots (more -> those contacts go the global way)
constexpr int MAX_BLOCKS = 8192;  // block-start table of the scatter kernel lives in shared memory
constexpr int JP_CONTACT_CAP = 16384;  // reproducible colouring keeps 5 B of state per interior contact in shared memory
// normal tile, 11 float4 per row: {n, slot pair} | 4 x ({r0 x n, jacCoeffInv}, {-(r1 x n), b}) | lambda[4] | {bodyA, bodyB, batch, contact}
// (the angular Jacobians are constants of the step -- positions do not move during the solve -- so they are built once
// here instead of in each of the 2 * I passes like solveContact does: the row solve is a chain of dependent FP32
// operations, and this halves its length)
constexpr int NT_FIELDS = 11, NT_LAMBDA = 9, NT_TAIL = 10;
constexpr int NT_STRIDE = NT_FIELDS * 32;
// friction tile, 4 float4 per row: {centre, damping flag} | {t0, fJacCoeffInv[0]} | {t1, fJacCoeffInv[1]} | {fLambda[2], -, -}
constexpr int FT_FIELDS = 4, FT_LAMBDA = 3;
constexpr int FT_STRIDE = FT_FIELDS * 32;
constexpr unsigned int INVALID_IDS = 0xffffffffu;
constexpr int ITER_SMEM_PER_SLOT = (int)(sizeof(float4) * 5 + sizeof(unsigned short));
constexpr int ITER_SMEM_MAX = ITER_SMEM_PER_SLOT * (S_MAX + NSTATIC);

// scratch layout (unsigned ints, zeroed before the setup kernels): [blockCount | blockCursor | crossHist | crossCursor | misc]
enum
{
	MISC_CROSS_COUNT = 0,  // length of the cross list
	MISC_TILE_CURSOR = 1,  // interior tiles handed out so far
	MISC_MAX_KI = 2,       // largest number of interior colours of any block
	MISC_KC = 3,           // number of cross colours
	MISC_NUM = 16
};

// ---------------------------------------------------------------- grid barrier
// One monotonically increasing arrival counter (zeroed by the host before the launch).  Thread 0 of every CTA arrives
// with a gpu-scope RELEASE add and spins with relaxed loads until all CTAs of this generation have arrived, then
// fences; bar.sync on either side extends the ordering to the rest of the CTA.  All CTAs are co-resident (cooperative
// launch).  MEASURED CAVEAT: the fence by thread 0 does not stop the OTHER threads' plain loads from hitting stale lines
// in the SM's L1, so every value another CTA may have written during the kernel is read with ld.global.cg (__ldcg),
// volatile or an atomic -- never a plain load.
struct GridBarrier
{
	unsigned int* counter;
	unsigned int numBlocks;
	unsigned int target;
	unsigned int seen;
	B3_D void init(unsigned int* c, unsigned int nb)
	{
		counter = c;
		numBlocks = nb;
		target = 0;
		seen = 0;
	}
	// split form: arrive() publishes this CTA's writes and signals, wait() blocks until every CTA has arrived;
	// independent loads may be issued in between
	B3_D void arrive()
	{
		target += numBlocks;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
			seen = target - 1u;
		}
	}
	B3_D void wait()
	{
		if (threadIdx.x == 0)
		{
			// Poll with relaxed loads.  No acquire fence afterwards: it costs 0.15-0.3 us per barrier (MEMBAR + CCTL.IVALL, measured
			// with tools/cu/barrier_bench.cu) and buys nothing here, because every datum another CTA writes during the kernel is
			// read with L2-only loads (ld.cg / volatile / atomics) that are issued after this loop has seen the last arrival, and the
			// arriving side has made its stores visible at L2 before its arrival (release).
			unsigned int v = seen;
			while ((int)(v - target) < 0)
			{
				asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
			}
		}
		__syncthreads();
	}
	B3_D void sync()
	{
		arrive();
		wait();
	}
};

B3_HD unsigned int hashU32(unsigned int h)
{
	h ^= h >> 16;
	h *= 0x85EBCA6Bu;
	h ^= h >> 13;
	h *= 0xC2B2AE35u;
	h ^= h >> 16;
	return h;
}

// counters[key] += 1 for every lane with key >= 0, one atomic per distinct key in the warp.  Returns the lane's own slot.
// Whole warp calls.
B3_D unsigned int warpCountByKey(unsigned int* counters, int key, int lane)
{
	const unsigned int peers = __match_any_sync(0xffffffffu, key);
	unsigned int base = 0;
	const int leader = __ffs(peers) - 1;
	if (key >= 0 && lane == leader) base = atomicAdd(&counters[key], (unsigned int)__popc(peers));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + (unsigned int)__popc(peers & ((1u << lane) - 1u));
}

// ================================================================ partition
// float -> unsigned int whose order is the float order
B3_D unsigned int orderedBits(float f)
{
	const unsigned int u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
B3_D float fromOrderedBits(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void __launch_bounds__(256) partBoundsKernel(const float4* __restrict__ pose, int n, unsigned int* __restrict__ bounds)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float4 p = mk4(0, 0, 0, 0);
	if (i < n) p = pose[2 * i];
	const bool dyn = i < n && p.w != 0.f && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
	unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
	if (dyn)
	{
		lo[0] = hi[0] = orderedBits(p.x);
		lo[1] = hi[1] = orderedBits(p.y);
		lo[2] = hi[2] = orderedBits(p.z);
	}
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
		hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
	}
	const unsigned int m = __ballot_sync(0xffffffffu, dyn);
	if ((threadIdx.x & 31) == 0 && m)
	{
#pragma unroll
		for (int k = 0; k < 3; k++)
		{
			atomicMin(&bounds[k], lo[k]);
			atomicMax(&bounds[3 + k], hi[k]);
		}
		atomicAdd(&bounds[6], (unsigned int)__popc(m));
	}
}

B3_D unsigned int spread10(unsigned int v)
{
	v &= 1023u;
	v = (v | (v << 16)) & 0x030000FFu;
	v = (v | (v << 8)) & 0x0300F00Fu;
	v = (v | (v << 4)) & 0x030C30C3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

__global__ void __launch_bounds__(256) partKeysKernel(const float4* __restrict__ pose, int n, const unsigned int* __restrict__ bounds,
													  unsigned int* __restrict__ keys, unsigned int* __restrict__ vals, const int* __restrict__ worldOf, int worldShift)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 p = pose[2 * i];
	unsigned int key = 0xffffffffu;  // static bodies sort last
	if (p.w != 0.f)
	{
		key = 0x3fffffffu;  // a dynamic body at a non-finite position still gets a block
		if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z))
		{
			const float mnx = fromOrderedBits(bounds[0]), mny = fromOrderedBits(bounds[1]), mnz = fromOrderedBits(bounds[2]);
			const float ex = fromOrderedBits(bounds[3]) - mnx, ey = fromOrderedBits(bounds[4]) - mny, ez = fromOrderedBits(bounds[5]) - mnz;
			// cubic cells: the same scale on the three axes, so that a key range is a compact region whatever the aspect of the scene
			const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-6f));
			const float sc = 1023.f / ext;
			const unsigned int qx = (unsigned int)fminf(fmaxf((p.x - mnx) * sc, 0.f), 1023.f);
			const unsigned int qy = (unsigned int)fminf(fmaxf((p.y - mny) * sc, 0.f), 1023.f);
			const unsigned int qz = (unsigned int)fminf(fmaxf((p.z - mnz) * sc, 0.f), 1023.f);
			key = spread10(qx) | (spread10(qz) << 1) | (spread10(qy) << 2);
		}
		// batched independent worlds: world-major order (the top bits of the Morton key inside a world), so that a block of
		// consecutive ranks holds whole worlds and no contact crosses blocks when the block size is a multiple of a world
		if (worldOf) key = ((unsigned int)worldOf[i] << worldShift) | (key >> (30 - worldShift));
	}
	keys[i] = key;
	vals[i] = (unsigned int)i;
}

// rank r in Morton order -> block r / S, slot r % S; statics (key 0xffffffff) -> -1
__global__ void __launch_bounds__(256) partAssignKernel(const unsigned int* __restrict__ keys, const unsigned int* __restrict__ vals, int n, int S,
														int* __restrict__ bodyLoc)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const unsigned int body = vals[r];
	bodyLoc[body] = keys[r] == 0xffffffffu ? -1 : (((r / S) << 12) | (r % S));
}

static int blockSizeFor(const World* w)
{
	const int n = std::max(w->numBodies, 1);
	if (n <= S_MAX) return S_MAX;  // one block: the whole solve runs in one CTA without any grid barrier
	// k blocks per CTA (k = 1: the block state stays in shared memory for the whole solve), all CTAs equally loaded
	const int k = divUp(n, w->smCount * S_MAX);
	int S = divUp(n, w->smCount * k);
	S = std::max(S, S_MIN);
	S = std::min(S, S_MAX);
	// batched worlds of equal size: a block holds whole worlds (the ranks are world-major), so every contact is interior to a
	// block and the iteration kernels never meet a grid barrier
	const int D = w->worldDynBodies;
	if (w->numWorlds > 1 && D > 0 && D <= S_MAX)
	{
		int m = divUp(S, D);
		if (m * D > S_MAX) m = S_MAX / D;
		S = std::max(m, 1) * D;
	}
	return S;
}

static int ensurePartition(World* w)
{
	const int n = w->numBodies;
	if (w->partValid && w->partAge < w->partInterval && w->partBodies == n)
	{
		w->partAge++;
		return 0;
	}
	cudaStream_t s = w->stream;
	w->partS = blockSizeFor(w);
	w->partBlocksMax = std::max(divUp(std::max(n, 1), w->partS), 1);
	if (w->partBlocksMax > MAX_BLOCKS)
	{
		setLastError("solver: %d bodies need more than %d blocks", n, MAX_BLOCKS);
		return B3B200_ERR_INVALID;
	}
	const size_t nb = (size_t)std::max(n, 1);
	B3_TRY(w->dBodyLoc.reserve(nb));
	B3_TRY(w->dPartKeys.reserve(nb));
	B3_TRY(w->dPartVals.reserve(nb));
	B3_TRY(w->dPartBounds.reserve(8));
	// {min.xyz = 0xffffffff, max.xyz = 0, 0, 0} (memsets: capturable into the step graph, unlike a copy from pageable memory)
	B3_CUDA_CHECK(cudaMemsetAsync(w->dPartBounds.ptr, 0xff, 3 * sizeof(unsigned int), s));
	B3_CUDA_CHECK(cudaMemsetAsync(w->dPartBounds.ptr + 3, 0, 5 * sizeof(unsigned int), s));
	if (n > 0)
	{
		partBoundsKernel<<<divUp(n, 256), 256, 0, s>>>(w->dPose.ptr, n, w->dPartBounds.ptr);
		B3_LAUNCH_CHECK();
		int worldShift = 30;
		if (w->numWorlds > 1)
		{
			int wbits = 1;
			while ((1 << wbits) < w->numWorlds) wbits++;
			worldShift = std::max(31 - wbits, 1);  // keys stay below 0x80000000 (statics: 0xffffffff)
			if (worldShift > 30) worldShift = 30;
		}
		partKeysKernel<<<divUp(n, 256), 256, 0, s>>>(w->dPose.ptr, n, w->dPartBounds.ptr, w->dPartKeys.ptr, w->dPartVals.ptr,
													 w->numWorlds > 1 ? w->dBodyWorld.ptr : nullptr, worldShift);
		B3_LAUNCH_CHECK();
		B3_TRY(radixSortKV32(s, w->partSortTmp, w->dPartKeys.ptr, w->dPartVals.ptr, n, 32));
		partAssignKernel<<<divUp(n, 256), 256, 0, s>>>(w->dPartKeys.ptr, w->dPartVals.ptr, n, w->partS, w->dBodyLoc.ptr);
		B3_LAUNCH_CHECK();
	}
	w->partValid = true;
	w->partAge = 1;
	w->partBodies = n;
	return 0;
}

// ================================================================ setup
struct SetupArgs
{
	b3b200_contact4* contacts;
	unsigned int* ctr;
	const float4* pose;
	const b3b200_inertia* inertias;
	const int* bodyLoc;
	unsigned long long* bodyMask;  // 2 per body: colours of the body's CROSS contacts
	unsigned long long* bodyPrio;  // 1 per body (reproducible cross colouring)
	int* contactBlock;             // owner block, -1 = cross
	unsigned int* contactSlots;    // slotA | slotB << 16 of an interior contact
	int* contactColour;            // colour inside its class (cross colour / interior colour of its block), -2 = left out
	int2* contactPair;             // dynamic bodies of the contact (-1 = that side is static): what the colouring needs, 8 coalesced bytes
	int* tileSrc;                  // per row slot of every tile: the contact it is built from, -1 = padding
	// body state in BLOCK ORDER (rank r = block * S + slot): what the iteration kernels read and write; filled by solverGatherKernel
	float4 *gLin, *gAng, *gInerA, *gInerB, *gPos;
	unsigned char* gBoundary;      // rank has cross contacts
	const unsigned int* partVals;  // rank -> body
	const unsigned int* partBounds;
	const float4* vel;
	unsigned int* blockCount;      // scratch, see the enum above
	unsigned int* blockCursor;
	unsigned int* crossHist;
	unsigned int* ctaBase;  // [setup grid][numBlocksMax + 1]: where each CTA's contacts go inside every block's list (+ in the cross list)
	unsigned int* crossCursor;
	unsigned int* misc;
	int* blockStatics;             // NSTATIC per block, -1 = free
	unsigned int* blockStart;      // numBlocksMax + 1
	unsigned int* blockList;       // interior contacts grouped by block
	unsigned int* crossList;
	unsigned int* blockTileBase;   // per block
	unsigned int* blockTileOff;    // (MAX_BATCHES + 1) per block: first tile of every interior colour, relative to the base
	unsigned int* crossTileOff;    // MAX_BATCHES + 1: first tile of every cross colour, relative to crossTileBase
	float4* tilesN;
	float4* tilesF;
	unsigned int tileCap;          // tiles in the buffers; interior tiles grow from 0, cross tiles sit at the top
	int numBodies;
	int numBlocksMax;
	int S;
	int staticIdx;
	int colouring;  // 0 = reproducible (priority rounds), 1 = single pass first fit with atomics
	float dt, positionDrift, positionConstraintCoeff;
};

B3_D float4 matRowMul(const float4& r0, const float4& r1, const float4& r2, const float4& v) { return mk4(dot3(r0, v), dot3(r1, v), dot3(r2, v)); }

// calcJacCoeff (b3ConvertConstraint4.h:50-60)
B3_D float calcJacCoeff(const float4& angular0, const float4& angular1, float invMass0, const float4* I0, float invMass1, const float4* I1)
{
	float jmj0 = invMass0;
	float jmj1 = dot3(matRowMul(I0[0], I0[1], I0[2], angular0), angular0);
	float jmj2 = invMass1;
	float jmj3 = dot3(matRowMul(I1[0], I1[1], I1[2], angular1), angular1);
	return -1.f / (jmj0 + jmj1 + jmj2 + jmj3);
}
// b3PlaneSpace1 (b3ConvertConstraint4.h:5-34)
B3_D void planeSpace1(const float4& n, float4& p, float4& q)
{
	if (fabsf(n.z) > 0.70710678f)
	{
		float a = n.y * n.y + n.z * n.z;
		float k = 1.f / sqrtf(a);
		p = mk4(0.f, -n.z * k, n.y * k);
		q = mk4(a * k, -n.x * p.z, n.x * p.y);
	}
	else
	{
		float a = n.x * n.x + n.y * n.y;
		float k = 1.f / sqrtf(a);
		p = mk4(-n.y * k, n.x * k, 0.f);
		q = mk4(-n.z * p.y, n.z * p.x, a * k);
	}
}

// setConstraint4 (b3ConvertConstraint4.h:62-148) -> one row of a normal tile + one row of a friction tile.
// `idsWord` goes into the w of the first field (slot pair of an interior row; unused by cross rows), `batch` is the global
// batch index.  b = e * relVelN + ... with e = 0 in the reference: the velocities are not needed here.
B3_D void buildRow(const SetupArgs& s, int c, unsigned int idsWord, int batch, float4* __restrict__ tn, float4* __restrict__ tf)
{
	const b3b200_contact4* src = &s.contacts[c];
	const float4* cw = reinterpret_cast<const float4*>(src);
	float4 wp[4] = {cw[0], cw[1], cw[2], cw[3]};
	const float4 nrm = cw[4];
	const int4 ids = reinterpret_cast<const int4*>(src)[5];
	const int aIdx = abs(ids.z), bIdx = abs(ids.w);
	const float4 posA = s.pose[2 * aIdx], posB = s.pose[2 * bIdx];
	const float invMassA = posA.w, invMassB = posB.w;
	// quirk kept from the reference: rows are built with the LOCAL initial inverse inertia
	// (solverSetup.cl:254,260 / b3Solver.cpp:911,917), solved with the world one.
	const float4* IA = reinterpret_cast<const float4*>(&s.inertias[aIdx].initInvInertia);
	const float4* IB = reinterpret_cast<const float4*>(&s.inertias[bIdx].initInvInertia);
	float4 ia[3] = {__ldg(IA), __ldg(IA + 1), __ldg(IA + 2)};
	float4 ib[3] = {__ldg(IB), __ldg(IB + 1), __ldg(IB + 2)};

	const float dtInv = 1.f / s.dt;
	const float npoints = nrm.w;
	float jac[4], bb[4];
	float4 ang0[4], ang1[4];
	const float4 n = mk4(nrm.x, nrm.y, nrm.z);
#pragma unroll
	for (int ic = 0; ic < 4; ic++)
	{
		float4 r0 = sub3(wp[ic], posA);
		float4 r1 = sub3(wp[ic], posB);
		ang0[ic] = ang1[ic] = mk4(0, 0, 0);
		if ((float)ic >= npoints)
		{
			jac[ic] = 0.f;
			bb[ic] = 0.f;
			continue;
		}
		float4 angular0 = cross3(r0, n);
		float4 angular1 = neg3(cross3(r1, n));
		ang0[ic] = angular0;
		ang1[ic] = angular1;
		jac[ic] = calcJacCoeff(angular0, angular1, invMassA, ia, invMassB, ib);
		float b = 0.f;  // e * relVelN, e = 0
		b += (wp[ic].w + s.positionDrift) * s.positionConstraintCoeff * dtInv;
		bb[ic] = b;
	}
	float fjac[2] = {0.f, 0.f};
	float4 center = mk4(0, 0, 0);
	float4 t0 = mk4(0, 0, 0), t1 = mk4(0, 0, 0);
	if (npoints > 0)
	{
		for (int i = 0; (float)i < npoints && i < 4; i++)
		{
			center.x += wp[i].x;
			center.y += wp[i].y;
			center.z += wp[i].z;
		}
		float inv = 1.0f / (float)npoints;
		center.x *= inv;
		center.y *= inv;
		center.z *= inv;
		planeSpace1(neg3(n), t0, t1);  // the tangents solveFriction uses (of -linear, b3Solver.cpp:278-282)
		float4 r0 = sub3(center, posA), r1 = sub3(center, posB);
		{
			float4 a0 = cross3(r0, t0), a1 = neg3(cross3(r1, t0));
			fjac[0] = calcJacCoeff(a0, a1, invMassA, ia, invMassB, ib);
		}
		{
			float4 a0 = cross3(r0, t1), a1 = neg3(cross3(r1, t1));
			fjac[1] = calcJacCoeff(a0, a1, invMassA, ia, invMassB, ib);
		}
	}
	tn[0] = mk4(nrm.x, nrm.y, nrm.z, __uint_as_float(idsWord));
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		const bool on = (float)i < npoints;
		tn[(1 + 2 * i) * 32] = on ? mk4(ang0[i].x, ang0[i].y, ang0[i].z, jac[i]) : mk4(0, 0, 0, 0);
		tn[(2 + 2 * i) * 32] = on ? mk4(ang1[i].x, ang1[i].y, ang1[i].z, bb[i]) : mk4(0, 0, 0, 0);
	}
	tn[NT_LAMBDA * 32] = mk4(0, 0, 0, 0);  // appliedRambdaDt
	// the two bodies as the iteration kernels address them: rank in block order, or -(body + 2) for a body without a rank (static)
	const int la = s.bodyLoc[aIdx], lb = s.bodyLoc[bIdx];
	int4 tail;
	tail.x = la >= 0 ? (la >> 12) * s.S + (la & 4095) : -(aIdx + 2);
	tail.y = lb >= 0 ? (lb >> 12) * s.S + (lb & 4095) : -(bIdx + 2);
	tail.z = batch;
	tail.w = c;
	reinterpret_cast<int4*>(tn)[NT_TAIL * 32] = tail;
	// angular damping of solveFriction (b3Solver.cpp:317-328): its condition only depends on positions
	float damp = 0.f;
	{
		const float4 ab = normalized3(sub3(posB, posA));
		const float4 ac = normalized3(sub3(center, posA));
		if (dot3(ab, ac) > 0.95f || (invMassA == 0.f || invMassB == 0.f)) damp = 1.f;
	}
	tf[0] = mk4(center.x, center.y, center.z, damp);
	tf[32] = mk4(t0.x, t0.y, t0.z, fjac[0]);
	tf[64] = mk4(t1.x, t1.y, t1.z, fjac[1]);
	tf[FT_LAMBDA * 32] = mk4(0, 0, 0, 0);  // fAppliedRambdaDt
}

B3_D void buildPadding(float4* __restrict__ tn, float4* __restrict__ tf)
{
	tn[0] = mk4(0, 0, 0, __uint_as_float(INVALID_IDS));
#pragma unroll
	for (int i = 1; i < NT_TAIL; i++) tn[i * 32] = mk4(0, 0, 0, 0);
	int4 tail;
	tail.x = 0;
	tail.y = 0;
	tail.z = -1;
	tail.w = -1;  // no contact: padding
	reinterpret_cast<int4*>(tn)[NT_TAIL * 32] = tail;
#pragma unroll
	for (int i = 0; i < FT_FIELDS; i++) tf[i * 32] = mk4(0, 0, 0, 0);
}

B3_D void contactBodies(const SetupArgs& s, int c, int& a, int& b, bool& aStatic, bool& bStatic, int& la, int& lb)
{
	const int4 ids = reinterpret_cast<const int4*>(&s.contacts[c])[5];
	a = abs(ids.z);
	b = abs(ids.w);
	la = s.bodyLoc[a];
	lb = s.bodyLoc[b];
	aStatic = ids.z < 0 || ids.z == s.staticIdx || la < 0;
	bStatic = ids.w < 0 || ids.w == s.staticIdx || lb < 0;
}

// ---- K1: classify every contact (interior to which block / cross), count per block, clear the cross masks.  This is the
// only setup kernel that reads the body ids out of the 112-byte contact records (one 32-byte sector each) and looks the
// bodies up in the partition; the kernels after it work on the packed per-contact words it leaves.
__global__ void __launch_bounds__(SETUP_THREADS) solverClassifyKernel(SetupArgs s)
{
	// the per-block contact counts are gathered per CTA in shared memory and added to the global table once at the end (a few
	// hundred hot addresses hit by every warp were what bounded this kernel)
	extern __shared__ unsigned int sCount[];  // numBlocksMax + 1
	for (int i = threadIdx.x; i <= s.numBlocksMax; i += blockDim.x) sCount[i] = 0u;
	__syncthreads();
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	const int nContacts = (int)s.ctr[CTR_CONTACTS];
	for (int i = tid; i < s.numBodies; i += stride)
	{
		s.bodyMask[2 * i] = 0ull;
		s.bodyMask[2 * i + 1] = 0ull;
		s.bodyPrio[i] = 0ull;
	}
	for (int base = tid - lane; base < nContacts; base += stride)
	{
		const int c = base + lane;
		int owner = -1;
		if (c < nContacts)
		{
			int a, b, la, lb;
			bool aStatic, bStatic;
			contactBodies(s, c, a, b, aStatic, bStatic, la, lb);
			if (aStatic != bStatic)
				owner = (aStatic ? lb : la) >> 12;
			else if (!aStatic && (la >> 12) == (lb >> 12))
				owner = la >> 12;
			unsigned int sa = (unsigned int)(la & 4095), sb = (unsigned int)(lb & 4095);
			if (owner >= 0 && aStatic != bStatic)
			{
				// the static body gets one of the block's own static slots (find or insert)
				const int g = aStatic ? a : b;
				int* table = s.blockStatics + (size_t)owner * NSTATIC;
				int slot = -1;
				for (int k = 0; k < NSTATIC; k++)
				{
					int v = *((volatile int*)&table[k]);
					if (v == -1) v = atomicCAS(&table[k], -1, g), v = (v == -1) ? g : v;
					if (v == g)
					{
						slot = k;
						break;
					}
				}
				if (slot < 0)
					owner = -1;  // more than NSTATIC static bodies under one block: this contact goes the global way
				else if (aStatic)
					sa = (unsigned int)(s.S + slot);
				else
					sb = (unsigned int)(s.S + slot);
			}
			s.contactBlock[c] = owner;
			s.contactSlots[c] = sa | (sb << 16);
			s.contactPair[c] = make_int2(aStatic ? -1 : a, bStatic ? -1 : b);
		}
		warpCountByKey(sCount, owner, lane);
		// (cross contacts are counted under the extra key numBlocksMax)
		warpCountByKey(sCount, c < nContacts && owner < 0 ? s.numBlocksMax : -1, lane);
	}
	__syncthreads();
	// one atomic per (CTA, block): the returned base is where THIS CTA's contacts of that block go in the block's list.  The
	// scatter kernel (same grid, same contact -> CTA mapping) then needs shared-memory cursors only.
	unsigned int* mine = s.ctaBase + (size_t)blockIdx.x * (s.numBlocksMax + 1);
	for (int i = threadIdx.x; i <= s.numBlocksMax; i += blockDim.x)
	{
		const unsigned int cnt = sCount[i];
		unsigned int* dst = i < s.numBlocksMax ? &s.blockCount[i] : &s.misc[MISC_CROSS_COUNT];
		mine[i] = cnt ? atomicAdd(dst, cnt) : 0u;
	}
}

// lowest colour not in (m0, m1), -2 when all 128 are taken
B3_D int lowestFree(unsigned long long m0, unsigned long long m1)
{
	if (~m0) return __ffsll((long long)~m0) - 1;
	if (~m1) return 64 + __ffsll((long long)~m1) - 1;
	return -2;
}

// Single-pass colouring: the contact takes the lowest colour free on its dynamic bodies by setting the colour's bit in
// the bodies' masks with atomicOr -- whoever flips a bit from 0 to 1 owns that (body, colour) -- and retries with fresh
// masks when it loses a race.  The bodies are taken in index order, so two contacts can never hold one bit each and wait
// for the other's; a contact that loses on its second body gives the first bit back.  The colours depend on how the races
// resolve (not reproducible from run to run).  MASKS: global (cross contacts) or shared (interior contacts of a block).
template <typename LoadMask>
B3_D int colourFirstFit(unsigned long long* masks, int i0, int i1, int nb, LoadMask ld)
{
	if (nb == 2 && i0 > i1)
	{
		const int t = i0;
		i0 = i1;
		i1 = t;
	}
	for (;;)
	{
		unsigned long long m0 = 0ull, m1 = 0ull;
		if (nb > 0)
		{
			m0 |= ld(&masks[2 * i0]);
			m1 |= ld(&masks[2 * i0 + 1]);
		}
		if (nb > 1)
		{
			m0 |= ld(&masks[2 * i1]);
			m1 |= ld(&masks[2 * i1 + 1]);
		}
		const int colour = lowestFree(m0, m1);
		if (colour < 0) return -2;
		const unsigned long long bit = 1ull << (colour & 63);
		const int word = colour >> 6;
		if (nb > 0 && (atomicOr(&masks[2 * i0 + word], bit) & bit)) continue;
		if (nb > 1 && (atomicOr(&masks[2 * i1 + word], bit) & bit))
		{
			atomicAnd(&masks[2 * i0 + word], ~bit);
			continue;
		}
		return colour;
	}
}

// ---- K2: interior contacts -> their block's list; cross contacts -> cross list (+ first-fit colour)
__global__ void __launch_bounds__(SETUP_THREADS) solverScatterKernel(SetupArgs s)
{
	extern __shared__ unsigned int sStart[];  // numBlocksMax + 1
	__shared__ unsigned int sWarp[SETUP_THREADS / 32];
	__shared__ unsigned int sCarry;
	__shared__ unsigned int sCrossHist[MAX_BATCHES];  // cross colours of this CTA's contacts, added to the global histogram at the end
	for (int i = threadIdx.x; i < MAX_BATCHES; i += blockDim.x) sCrossHist[i] = 0u;
	// this CTA's cursors into the blocks' lists and into the cross list, from the ranges the classify kernel reserved
	unsigned int* sCursor = sStart + s.numBlocksMax + 1;  // numBlocksMax + 1
	for (int i = threadIdx.x; i <= s.numBlocksMax; i += blockDim.x) sCursor[i] = s.ctaBase[(size_t)blockIdx.x * (s.numBlocksMax + 1) + i];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// exclusive scan of the block counts (every CTA computes its own copy; CTA 0 also publishes it)
	if (threadIdx.x == 0) sCarry = 0;
	__syncthreads();
	for (int base = 0; base < s.numBlocksMax; base += SETUP_THREADS)
	{
		const int i = base + threadIdx.x;
		const unsigned int v = i < s.numBlocksMax ? s.blockCount[i] : 0u;
		unsigned int incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) sWarp[warp] = incl;
		__syncthreads();
		unsigned int before = sCarry;
		for (int k = 0; k < warp; k++) before += sWarp[k];
		if (i < s.numBlocksMax) sStart[i] = before + incl - v;
		__syncthreads();
		if (threadIdx.x == SETUP_THREADS - 1) sCarry = before + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) sStart[s.numBlocksMax] = sCarry;
	__syncthreads();
	if (blockIdx.x == 0)
		for (int i = threadIdx.x; i <= s.numBlocksMax; i += SETUP_THREADS) s.blockStart[i] = sStart[i];

	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int nContacts = (int)s.ctr[CTR_CONTACTS];
	const unsigned int contactCap = s.colouring == 0 ? (unsigned int)JP_CONTACT_CAP : 0xffffffffu;
	for (int base = tid - lane; base < nContacts; base += stride)
	{
		const int c = base + lane;
		int owner = c < nContacts ? s.contactBlock[c] : -1;
		// interior: a place in the block's list (one atomic per distinct block in the warp)
		const bool wasCross = c < nContacts && owner < 0;
		const unsigned int pos = warpCountByKey(sCursor, owner, lane);
		if (owner >= 0)
		{
			if (pos >= contactCap)
			{
				owner = -1;
				s.contactBlock[c] = -1;
			}
			else
				s.blockList[sStart[owner] + pos] = (unsigned int)c;
		}
		const bool cross = c < nContacts && owner < 0;
		int colour = -1;
		if (cross && s.colouring == 1)
		{
			const int2 pr = s.contactPair[c];
			int i0 = 0, i1 = 0, nb = 0;
			if (pr.x >= 0) i0 = pr.x, nb = 1;
			if (pr.y >= 0)
			{
				if (nb)
					i1 = pr.y;
				else
					i0 = pr.y;
				nb++;
			}
			colour = colourFirstFit(s.bodyMask, i0, i1, nb, [](const unsigned long long* p) { return __ldcg(p); });
			s.contactColour[c] = colour;
			if (colour < 0)
			{
				// more than B3_MAX_NUM_BATCHES colours at one body: the reference errors out
				// (b3GpuPgsContactSolver.cpp:1497-1502); here the contact is left out of this step's solve
				s.contacts[c].batchIdx = -2;
				atomicOr(&s.ctr[CTR_OVERFLOW], (unsigned int)OVF_BATCHES);
			}
		}
		__syncwarp();
		warpCountByKey(sCrossHist, colour, lane);
		// cross list: the contacts the classify kernel counted go to this CTA's reserved range; the few that overflowed their
		// block's cap just now take a slot behind all reserved ranges
		const unsigned int slotR = warpCountByKey(sCursor, wasCross ? s.numBlocksMax : -1, lane);
		if (wasCross)
			s.crossList[slotR] = (unsigned int)c;
		else if (cross)
			s.crossList[atomicAdd(&s.misc[MISC_CROSS_COUNT], 1u)] = (unsigned int)c;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < MAX_BATCHES; i += blockDim.x)
		if (sCrossHist[i]) atomicAdd(&s.crossHist[i], sCrossHist[i]);
}

// ---- K2b (reproducible mode only): the cross contacts are coloured by priority rounds (Jones-Plassmann) in ONE CTA:
// per round every uncoloured contact posts its priority on its dynamic bodies, the contact that is top on all of them takes
// the lowest colour free on them.  Equals the sequential first fit in descending priority order: a pure function of the
// contact array.
B3_D unsigned long long contactPrio(int c) { return ((unsigned long long)hashU32((unsigned int)c * 0x9E3779B1u + 0x7F4A7C15u) << 32) | (unsigned long long)(unsigned int)(c + 1); }

__global__ void __launch_bounds__(1024) solverCrossColourKernel(SetupArgs s)
{
	__shared__ int sLeft;
	const int n = (int)s.misc[MISC_CROSS_COUNT];
	const int lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < n; i += blockDim.x) s.contactColour[s.crossList[i]] = -1;
	__syncthreads();
	for (int round = 0; round < 100000; round++)
	{
		if (threadIdx.x == 0) sLeft = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x)
		{
			const int c = (int)s.crossList[i];
			if (s.contactColour[c] != -1) continue;  // own earlier write
			int a, b, la, lb;
			bool aStatic, bStatic;
			contactBodies(s, c, a, b, aStatic, bStatic, la, lb);
			const unsigned long long prio = contactPrio(c);
			if (!aStatic) atomicMax(&s.bodyPrio[a], prio);
			if (!bStatic) atomicMax(&s.bodyPrio[b], prio);
		}
		__syncthreads();
		for (int base = (int)threadIdx.x - lane; base < n; base += blockDim.x)
		{
			const int i = base + lane;
			int got = -1;
			if (i < n)
			{
				const int c = (int)s.crossList[i];
				if (s.contactColour[c] == -1)
				{
					int a, b, la, lb;
					bool aStatic, bStatic;
					contactBodies(s, c, a, b, aStatic, bStatic, la, lb);
					const unsigned long long prio = contactPrio(c);
					volatile unsigned long long* vp = s.bodyPrio;
					if ((aStatic || vp[a] == prio) && (bStatic || vp[b] == prio))
					{
						unsigned long long m0 = 0ull, m1 = 0ull;
						if (!aStatic) m0 |= __ldcg(&s.bodyMask[2 * a]), m1 |= __ldcg(&s.bodyMask[2 * a + 1]);
						if (!bStatic) m0 |= __ldcg(&s.bodyMask[2 * b]), m1 |= __ldcg(&s.bodyMask[2 * b + 1]);
						const int colour = lowestFree(m0, m1);
						if (colour >= 0)
						{
							const unsigned long long bit = 1ull << (colour & 63);
							const int word = colour >> 6;
							if (!aStatic) __stcg(&s.bodyMask[2 * a + word], __ldcg(&s.bodyMask[2 * a + word]) | bit);
							if (!bStatic) __stcg(&s.bodyMask[2 * b + word], __ldcg(&s.bodyMask[2 * b + word]) | bit);
							got = colour;
						}
						else
						{
							s.contacts[c].batchIdx = -2;
							atomicOr(&s.ctr[CTR_OVERFLOW], (unsigned int)OVF_BATCHES);
						}
						s.contactColour[c] = colour;
						if (!aStatic) vp[a] = 0ull;
						if (!bStatic) vp[b] = 0ull;
					}
					else
						sLeft = 1;
				}
			}
			__syncwarp();
			warpCountByKey(s.crossHist, got, lane);
		}
		__syncthreads();
		const int left = sLeft;
		__syncthreads();
		if (!left) break;
	}
}

// first tile of every colour (tiles of 32 rows) from the colour histogram; returns the number of colours in use
B3_D int tileOffsetsFromHist(const unsigned int* hist, unsigned int* off, int lane)
{
	// one warp
	unsigned int run = 0;
	int numColours = 0;
	for (int base = 0; base < MAX_BATCHES; base += 32)
	{
		const unsigned int raw = hist[base + lane];
		const unsigned int v = (raw + 31u) >> 5;
		unsigned int incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		off[base + lane] = run + incl - v;
		const unsigned int nz = __ballot_sync(0xffffffffu, raw != 0);
		if (nz) numColours = base + 32 - __clz(nz);
		run += __shfl_sync(0xffffffffu, incl, 31);
	}
	if (lane == 0) off[MAX_BATCHES] = run;
	return numColours;
}

// ---- K3: one CTA per block: colour the block's interior contacts in shared memory, lay out its tiles, build its rows
__global__ void __launch_bounds__(SETUP_THREADS) solverBlockSetupKernel(SetupArgs s)
{
	extern __shared__ unsigned long long sm64[];
	__shared__ unsigned int sHist[MAX_BATCHES], sCursor[MAX_BATCHES], sOff[MAX_BATCHES + 1], sCrossHist[MAX_BATCHES];
	__shared__ unsigned int sTileBase;
	__shared__ int sKc, sKi, sLeft;
	const int slots = s.S + NSTATIC;
	unsigned long long* sMask = sm64;                  // 2 per slot
	unsigned long long* sPrio = sm64 + 2 * slots;      // 1 per slot (reproducible mode)
	unsigned int* sCSlots = reinterpret_cast<unsigned int*>(sPrio + slots);  // per contact (reproducible mode)
	signed char* sCCol = reinterpret_cast<signed char*>(sCSlots + JP_CONTACT_CAP);
	const int lane = threadIdx.x & 31;

	// number of cross colours: interior colour m of any block is global batch Kc + m
	if (threadIdx.x < MAX_BATCHES) sCrossHist[threadIdx.x] = s.crossHist[threadIdx.x];
	__syncthreads();
	if (threadIdx.x < 32)
	{
		int kc = 0;
		for (int base = 0; base < MAX_BATCHES; base += 32)
		{
			const unsigned int nz = __ballot_sync(0xffffffffu, sCrossHist[base + lane] != 0);
			if (nz) kc = base + 32 - __clz(nz);
		}
		if (lane == 0)
		{
			sKc = kc;
			if (blockIdx.x == 0) s.misc[MISC_KC] = (unsigned int)kc;
		}
	}
	__syncthreads();
	const int Kc = sKc;
	{
		// the row slots of the cross tiles (top of the buffers) start out as padding; solverCrossBuildKernel fills them
		unsigned int crossTiles = 0;
		for (int k = 0; k < MAX_BATCHES; k++) crossTiles += (sCrossHist[k] + 31u) >> 5;
		const size_t first = (size_t)(s.tileCap - crossTiles) * 32u;
		for (size_t k = (size_t)blockIdx.x * SETUP_THREADS + threadIdx.x; k < (size_t)crossTiles * 32u; k += (size_t)gridDim.x * SETUP_THREADS) s.tileSrc[first + k] = -1;
	}

	for (int blk = blockIdx.x; blk < s.numBlocksMax; blk += gridDim.x)
	{
		const unsigned int first = s.blockStart[blk];
		const unsigned int counted = s.blockStart[blk + 1] - first;
		unsigned int n = counted;  // (every position of the segment was handed to exactly one contact)
		if (s.colouring == 0 && n > (unsigned int)JP_CONTACT_CAP) n = JP_CONTACT_CAP;  // the scatter kernel sent the rest the global way
		const unsigned int* list = s.blockList + first;
		for (int i = threadIdx.x; i < 2 * slots; i += SETUP_THREADS) sMask[i] = 0ull;
		if (threadIdx.x < MAX_BATCHES)
		{
			sHist[threadIdx.x] = 0;
			sCursor[threadIdx.x] = 0;
		}
		__syncthreads();
		// ---- colouring
		if (s.colouring == 1)
		{
			for (unsigned int i = threadIdx.x; i < n; i += SETUP_THREADS)
			{
				const int c = (int)list[i];
				const unsigned int sl = s.contactSlots[c];
				const int sa = (int)(sl & 0xffffu), sb = (int)(sl >> 16);
				int i0 = 0, i1 = 0, nb = 0;
				if (sa < s.S) i0 = sa, nb = 1;
				if (sb < s.S)
				{
					if (nb)
						i1 = sb;
					else
						i0 = sb;
					nb++;
				}
				int colour = colourFirstFit(sMask, i0, i1, nb, [](const unsigned long long* p) { return *((volatile const unsigned long long*)p); });
				if (colour >= 0 && Kc + colour >= MAX_BATCHES) colour = -2;  // (its mask bits stay set: harmless, the block only loses a colour)
				s.contactColour[c] = colour;
				if (colour >= 0) atomicAdd(&sHist[colour], 1u);
			}
		}
		else
		{
			for (int i = threadIdx.x; i < slots; i += SETUP_THREADS) sPrio[i] = 0ull;
			for (unsigned int i = threadIdx.x; i < n; i += SETUP_THREADS)
			{
				sCSlots[i] = s.contactSlots[list[i]];
				sCCol[i] = -1;
			}
			__syncthreads();
			for (int round = 0; round < 100000; round++)
			{
				if (threadIdx.x == 0) sLeft = 0;
				// (the priorities are cleared between the rounds, not by the winners while the others still compare: no thread then
				// reads a word another one writes in the same phase)
				if (round > 0)
				{
					for (int i = threadIdx.x; i < slots; i += SETUP_THREADS) sPrio[i] = 0ull;
					__syncthreads();
				}
				for (unsigned int i = threadIdx.x; i < n; i += SETUP_THREADS)
				{
					if (sCCol[i] != -1) continue;
					const unsigned int sl = sCSlots[i];
					const int sa = (int)(sl & 0xffffu), sb = (int)(sl >> 16);
					const unsigned long long prio = contactPrio((int)list[i]);
					if (sa < s.S) atomicMax(&sPrio[sa], prio);
					if (sb < s.S) atomicMax(&sPrio[sb], prio);
				}
				__syncthreads();
				for (unsigned int i = threadIdx.x; i < n; i += SETUP_THREADS)
				{
					if (sCCol[i] != -1) continue;
					const unsigned int sl = sCSlots[i];
					const int sa = (int)(sl & 0xffffu), sb = (int)(sl >> 16);
					const unsigned long long prio = contactPrio((int)list[i]);
					volatile unsigned long long* vp = sPrio;
					const bool da = sa < s.S, db = sb < s.S;
					if ((!da || vp[sa] == prio) && (!db || vp[sb] == prio))
					{
						unsigned long long m0 = 0ull, m1 = 0ull;
						if (da) m0 |= sMask[2 * sa], m1 |= sMask[2 * sa + 1];
						if (db) m0 |= sMask[2 * sb], m1 |= sMask[2 * sb + 1];
						int colour = lowestFree(m0, m1);
						if (colour >= 0)
						{
							const unsigned long long bit = 1ull << (colour & 63);
							const int word = colour >> 6;
							if (da) sMask[2 * sa + word] |= bit;
							if (db) sMask[2 * sb + word] |= bit;
							if (Kc + colour >= MAX_BATCHES) colour = -2;
						}
						sCCol[i] = (signed char)colour;
						if (colour >= 0) atomicAdd(&sHist[colour], 1u);
					}
					else
						sLeft = 1;
				}
				__syncthreads();
				const int left = sLeft;
				__syncthreads();
				if (!left) break;
			}
			for (unsigned int i = threadIdx.x; i < n; i += SETUP_THREADS) s.contactColour[list[i]] = (int)sCCol[i];
		}
		__syncthreads();
		// ---- tile layout of the block
		if (threadIdx.x < 32)
		{
			const int ki = tileOffsetsFromHist(sHist, sOff, lane);
			__syncwarp();
			if (lane == 0)
			{
				sKi = ki;
				const unsigned int nTiles = sOff[MAX_BATCHES];
				unsigned int base = atomicAdd(&s.misc[MISC_TILE_CURSOR], nTiles);
				sTileBase = base;
				s.blockTileBase[blk] = base;
				atomicMax(&s.misc[MISC_MAX_KI], (unsigned int)ki);
			}
		}
		__syncthreads();
		for (int i = threadIdx.x; i <= MAX_BATCHES; i += SETUP_THREADS) s.blockTileOff[(size_t)blk * (MAX_BATCHES + 1) + i] = sOff[i];
		const unsigned int tileBase = sTileBase;
		// (capacity: interior + cross contacts <= contact capacity, and the buffers hold capacity / 32 + padding tiles)
		// ---- which contact every row slot of the block's tiles is built from (the rows themselves are built by
		// solverBuildRowsKernel, a warp per tile at full occupancy: this CTA's 512 threads cannot hide the latency of its
		// block's ~4 000 contact / pose / inertia gathers)
		{
			const unsigned int nSlots = sOff[MAX_BATCHES] * 32u;
			for (unsigned int k = threadIdx.x; k < nSlots; k += SETUP_THREADS) s.tileSrc[(size_t)tileBase * 32u + k] = -1;
		}
		__syncthreads();
		for (unsigned int i = threadIdx.x; i < n; i += SETUP_THREADS)
		{
			const int c = (int)list[i];
			const int colour = s.contactColour[c];  // this thread's own write
			if (colour < 0)
			{
				s.contacts[c].batchIdx = -2;
				atomicOr(&s.ctr[CTR_OVERFLOW], (unsigned int)OVF_BATCHES);
				continue;
			}
			const unsigned int rank = atomicAdd(&sCursor[colour], 1u);
			const unsigned int tile = tileBase + sOff[colour] + (rank >> 5);
			s.contacts[c].batchIdx = Kc + colour;
			s.tileSrc[(size_t)tile * 32u + (rank & 31u)] = c;
		}
		__syncthreads();
	}
}

// ---- K4: where the cross rows go: the top of the tile buffers, in colour order
__global__ void __launch_bounds__(SETUP_THREADS) solverCrossBuildKernel(SetupArgs s)
{
	__shared__ unsigned int sHist[MAX_BATCHES], sOff[MAX_BATCHES + 1];
	const int lane = threadIdx.x & 31;
	if (threadIdx.x < MAX_BATCHES) sHist[threadIdx.x] = s.crossHist[threadIdx.x];
	__syncthreads();
	if (threadIdx.x < 32)
	{
		const int kc = tileOffsetsFromHist(sHist, sOff, lane);
		if (blockIdx.x == 0 && lane == 0)
		{
			s.ctr[CTR_BATCHES] = (unsigned int)kc + s.misc[MISC_MAX_KI];
			s.ctr[CTR_COLOUR_ROUNDS] = (unsigned int)kc;  // diagnostics: number of cross colours = grid barriers per pass
		}
	}
	__syncthreads();
	const unsigned int crossTiles = sOff[MAX_BATCHES];
	const unsigned int crossBase = s.tileCap - crossTiles;
	if (blockIdx.x == 0)
		for (int i = threadIdx.x; i <= MAX_BATCHES; i += SETUP_THREADS) s.crossTileOff[i] = sOff[i];
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int n = (int)s.misc[MISC_CROSS_COUNT];
	for (int base = tid - lane; base < n; base += stride)
	{
		const int i = base + lane;
		const int c = i < n ? (int)s.crossList[i] : 0;
		const int colour = i < n ? s.contactColour[c] : -1;
		const unsigned int rank = warpCountByKey(s.crossCursor, colour, lane);
		if (colour < 0) continue;
		const unsigned int tile = crossBase + sOff[colour] + (rank >> 5);
		s.contacts[c].batchIdx = colour;
		s.tileSrc[(size_t)tile * 32u + (rank & 31u)] = c;
	}
}

// ---- K5: the rows, a warp per tile: lane r builds row r from the contact the layout kernels put there (scattered reads of
// whole 112-byte contact records, coalesced 512-byte stores of the row fields)
__global__ void __launch_bounds__(256, 2) solverBuildRowsKernel(SetupArgs s)
{
	const unsigned int interiorTiles = s.misc[MISC_TILE_CURSOR];
	const unsigned int crossTiles = s.crossTileOff[MAX_BATCHES];
	const unsigned int total = interiorTiles + crossTiles;
	const int lane = threadIdx.x & 31;
	for (unsigned int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < total; t += gridDim.x * (blockDim.x >> 5))
	{
		const bool interior = t < interiorTiles;
		const unsigned int tile = interior ? t : s.tileCap - crossTiles + (t - interiorTiles);
		const int c = s.tileSrc[(size_t)tile * 32u + lane];
		float4* tn = s.tilesN + (size_t)tile * NT_STRIDE + lane;
		float4* tf = s.tilesF + (size_t)tile * FT_STRIDE + lane;
		if (c < 0)
			buildPadding(tn, tf);
		else
			buildRow(s, c, interior ? s.contactSlots[c] : 0u, s.contacts[c].batchIdx, tn, tf);
	}
}

// ---- K6: body state into block order.  The iteration kernels then load / store a block with coalesced copies, and the cross rows
// address their bodies by rank (no indirection through the partition on the critical path of a pass).
__global__ void __launch_bounds__(256) solverGatherKernel(SetupArgs s, int numRanks)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRanks) return;
	const int numDyn = (int)s.partBounds[6];
	float4 lin = mk4(0, 0, 0), ang = lin, ia = lin, ib = lin, pos = lin;
	unsigned char boundary = 0;
	if (r < numDyn)
	{
		const int g = (int)s.partVals[r];
		const float4 ps = s.pose[2 * g];
		const float4* I = reinterpret_cast<const float4*>(&s.inertias[g].invInertiaWorld);
		const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
		lin = s.vel[2 * g];
		ang = s.vel[2 * g + 1];
		pos = ps;
		ia = mk4(r0.x, r0.y, r0.z, r1.y);
		ib = mk4(r1.z, r2.z, ps.w, 0.f);
		boundary = (s.bodyMask[2 * g] | s.bodyMask[2 * g + 1]) != 0ull;
	}
	s.gLin[r] = lin;
	s.gAng[r] = ang;
	s.gInerA[r] = ia;
	s.gInerB[r] = ib;
	s.gPos[r] = pos;
	s.gBoundary[r] = boundary;
}
__global__ void __launch_bounds__(256) solverScatterBackKernel(const float4* __restrict__ gLin, const float4* __restrict__ gAng, const float4* __restrict__ gInerB,
															  const unsigned int* __restrict__ partVals, const unsigned int* __restrict__ partBounds, float4* __restrict__ vel)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= (int)partBounds[6]) return;
	if (gInerB[r].z == 0.f) return;  // (no longer dynamic)
	const int g = (int)partVals[r];
	vel[2 * g] = gLin[r];
	vel[2 * g + 1] = gAng[r];
}

// ================================================================ iterations
struct IterArgs
{
	float4* tilesN;
	float4* tilesF;
	unsigned int tileCap;
	const unsigned int* ctr;
	const float4* pose;
	float4* vel;
	const b3b200_inertia* inertias;
	float4 *gLin, *gAng;               // velocities in block order (rank = block * S + slot): the cross rows' global memory
	const float4 *gInerA, *gInerB, *gPos;
	const unsigned char* gBoundary;
	const unsigned int* partBounds;    // [6] = number of dynamic bodies
	const int* blockStatics;
	const unsigned int* blockTileBase;
	const unsigned int* blockTileOff;
	const unsigned int* crossTileOff;
	const unsigned int* misc;
	const unsigned long long* bodyMask;  // != 0 -> the body has cross contacts
	unsigned int* bar;
	int iterations;
	int S;
	int numBlocksMax;
	unsigned long long* probe;  // development aid: globaltimer stamps of CTA 0 (nullptr = off)
};

// Iteration arithmetic: explicit FMAs.  This file is built with --fmad=false, so the only fused operations are the ones
// written here; the shared-memory path and the global path produce the same bits for the same Gauss-Seidel order.
// (The oracle evaluates the reference's unfused expressions; the bar for velocities is 1e-4 relative.)
B3_D float fdot3(const float4& a, const float4& b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
B3_D float4 fcross3(const float4& a, const float4& b)
{
	return mk4(__fmaf_rn(a.y, b.z, -(a.z * b.y)), __fmaf_rn(a.z, b.x, -(a.x * b.z)), __fmaf_rn(a.x, b.y, -(a.y * b.x)));
}
B3_D float4 fmatRowMul(const float4& r0, const float4& r1, const float4& r2, const float4& v) { return mk4(fdot3(r0, v), fdot3(r1, v), fdot3(r2, v)); }
// v + (dir * k) * s
B3_D float4 faddScaled(const float4& v, const float4& dir, float k, float s)
{
	return mk4(__fmaf_rn(dir.x * k, s, v.x), __fmaf_rn(dir.y * k, s, v.y), __fmaf_rn(dir.z * k, s, v.z));
}
B3_D float4 faddScaled1(const float4& v, const float4& dir, float s) { return mk4(__fmaf_rn(dir.x, s, v.x), __fmaf_rn(dir.y, s, v.y), __fmaf_rn(dir.z, s, v.z)); }
B3_D float fcalcRelVel(const float4& l0, const float4& l1, const float4& a0, const float4& a1, const float4& linVel0, const float4& angVel0,
					   const float4& linVel1, const float4& angVel1)
{
	return fdot3(l0, linVel0) + fdot3(a0, angVel0) + fdot3(l1, linVel1) + fdot3(a1, angVel1);
}

struct BodyConst
{
	float4 pos;  // w = inverse mass
	float4 i0, i1, i2;
};

// solveContact<false> (b3Solver.cpp:187-266): the four points of one manifold, with the angular Jacobians of the row.
// j0[i] = {r0 x n, jacCoeffInv}, j1[i] = {-(r1 x n), b}.
//
// The reference walks the points one after the other, each reading the velocities the previous one wrote: a chain of
// ~50 dependent FP32 operations per point, ~1 us per row on one warp -- and every colour of every pass waits for one such
// row.  The same sequence is evaluated here through its linear structure: with J_i the 12-component Jacobian of point i
// and D_i the velocity change per unit impulse, relVel_i = J_i . v0 + sum_{j<i} (J_i . D_j) dLambda_j.  J_i . v0 (4 dot
// products), the couplings J_i . D_j (6 scalars) and D_i depend on nothing the chain produces, so they issue back to back;
// what stays sequential is one FMA + the clamp per point, and the velocities are updated once at the end.  Same
// Gauss-Seidel order, same clamping; the results differ from the point-by-point evaluation by reassociation only
// (~1e-7 relative; the parity bar of the velocities is 1e-4).
B3_D void solveNormalCore(const float4& nId, const float4* j0, const float4* j1, float4& applied, const BodyConst& A, const BodyConst& B, float4& linVelA,
						  float4& angVelA, float4& linVelB, float4& angVelB)
{
	const float4 n = mk4(nId.x, nId.y, nId.z);
	const float invMassA = A.pos.w, invMassB = B.pos.w;
	float4 dAngA[4], dAngB[4];
	float rv[4];
	// (n . n)(invMassA + invMassB): the linear part of every coupling J_i . D_j
	const float linCoupling = fdot3(n, n) * (invMassA + invMassB);
	const float relLin = fdot3(n, linVelA) - fdot3(n, linVelB);
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		dAngA[i] = fmatRowMul(A.i0, A.i1, A.i2, j0[i]);
		dAngB[i] = fmatRowMul(B.i0, B.i1, B.i2, j1[i]);
		rv[i] = (relLin + fdot3(j0[i], angVelA)) + (fdot3(j1[i], angVelB) + j1[i].w);
	}
	const float lam[4] = {applied.x, applied.y, applied.z, applied.w};
	float d[4], out[4];
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		float r = rv[i];
#pragma unroll
		for (int j = 0; j < i; j++)
		{
			const float c = (linCoupling + fdot3(j0[i], dAngA[j])) + fdot3(j1[i], dAngB[j]);  // J_i . D_j (independent of the chain)
			r = __fmaf_rn(c, d[j], r);
		}
		const float jac = j0[i].w;
		float updated = __fmaf_rn(r, jac, lam[i]);
		updated = fmaxf(updated, 0.f);
		updated = fminf(updated, FLT_MAX);
		d[i] = jac == 0.f ? 0.f : updated - lam[i];  // (a point the manifold does not have: solveContact skips it)
		out[i] = jac == 0.f ? lam[i] : updated;
	}
	applied = mk4(out[0], out[1], out[2], out[3]);
	const float dsum = (d[0] + d[1]) + (d[2] + d[3]);
	linVelA = faddScaled1(linVelA, n, invMassA * dsum);
	linVelB = faddScaled1(linVelB, n, -(invMassB * dsum));
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		angVelA = faddScaled1(angVelA, dAngA[i], d[i]);
		angVelB = faddScaled1(angVelB, dAngB[i], d[i]);
	}
}

// solveFriction (b3Solver.cpp:268-329), the two tangent directions through the same linear structure.  cf = {centre, damping
// flag}, t0 / t1 = {tangent, fJacCoeffInv}, fl = fAppliedRambdaDt; returns false when the row has no friction
B3_D bool solveFrictionCore(const float4& nId, const float4& cf, const float4& t0, const float4& t1, float4& fl, const float4& applied, const BodyConst& A,
							const BodyConst& B, float4& linVelA, float4& angVelA, float4& linVelB, float4& angVelB)
{
	if (t0.w == 0.f && t0.w == 0.f) return false;  // (sic, b3Solver.cpp:274)
	const float invMassA = A.pos.w, invMassB = B.pos.w;
	float sum = 0.f;
	sum += applied.x;
	sum += applied.y;
	sum += applied.z;
	sum += applied.w;
	const float frictionCoeff = 0.7f;
	const float maxR = frictionCoeff * sum;
	const float minR = -maxR;
	const float4 n = neg3(mk4(nId.x, nId.y, nId.z));
	const float4 r0 = sub3(cf, A.pos), r1 = sub3(cf, B.pos);
	const float4 ta = mk4(t0.x, t0.y, t0.z), tb = mk4(t1.x, t1.y, t1.z);
	const float4 a0 = fcross3(r0, ta), a1 = neg3(fcross3(r1, ta));
	const float4 b0 = fcross3(r0, tb), b1 = neg3(fcross3(r1, tb));
	const float4 dA0 = fmatRowMul(A.i0, A.i1, A.i2, a0), dB0 = fmatRowMul(B.i0, B.i1, B.i2, a1);
	const float4 dA1 = fmatRowMul(A.i0, A.i1, A.i2, b0), dB1 = fmatRowMul(B.i0, B.i1, B.i2, b1);
	const float4 relV = sub3(linVelA, linVelB);
	const float rv0 = (fdot3(ta, relV) + fdot3(a0, angVelA)) + fdot3(a1, angVelB);
	const float rv1 = (fdot3(tb, relV) + fdot3(b0, angVelA)) + fdot3(b1, angVelB);
	const float c10 = (fdot3(tb, ta) * (invMassA + invMassB) + fdot3(b0, dA0)) + fdot3(b1, dB0);
	float u0 = __fmaf_rn(rv0, t0.w, fl.x);
	u0 = fminf(fmaxf(u0, minR), maxR);
	const float d0 = u0 - fl.x;
	float u1 = __fmaf_rn(__fmaf_rn(c10, d0, rv1), t1.w, fl.y);
	u1 = fminf(fmaxf(u1, minR), maxR);
	const float d1 = u1 - fl.y;
	linVelA = faddScaled1(faddScaled1(linVelA, ta, invMassA * d0), tb, invMassA * d1);
	linVelB = faddScaled1(faddScaled1(linVelB, ta, -(invMassB * d0)), tb, -(invMassB * d1));
	angVelA = faddScaled1(faddScaled1(angVelA, dA0, d0), dA1, d1);
	angVelB = faddScaled1(faddScaled1(angVelB, dB0, d0), dB1, d1);
	if (cf.w != 0.f)
	{
		// angular damping for point constraint (b3Solver.cpp:317-328); the condition was evaluated by the setup
		const float angNA = fdot3(n, angVelA);
		const float angNB = fdot3(n, angVelB);
		angVelA = faddScaled1(angVelA, n, -(angNA * 0.1f));
		angVelB = faddScaled1(angVelB, n, -(angNB * 0.1f));
	}
	fl = mk4(u0, u1, 0.f, 0.f);
	return true;
}

// ---- rows as they sit in registers between the (early) load and the solve
template <int PHASE>
struct RowRegs
{
	float4 nId;
	float4 j0[PHASE == 0 ? 4 : 1];  // normal: {r0 x n, jacCoeffInv} per point | friction: {centre, damping flag}
	float4 j1[PHASE == 0 ? 4 : 2];  // normal: {-(r1 x n), b} per point      | friction: {t0, fJacCoeffInv[0]}, {t1, fJacCoeffInv[1]}
	float4 applied;                 // lambda[4] of the normal rows
	float4 fl;                      // friction: fLambda[2]
};
template <int PHASE>
B3_D void loadRowRegs(const float4* __restrict__ tilesN, const float4* __restrict__ tilesF, unsigned int tile, int lane, RowRegs<PHASE>& r)
{
	const float4* tn = tilesN + (size_t)tile * NT_STRIDE + lane;
	r.nId = tn[0];
	r.applied = tn[NT_LAMBDA * 32];
	if (PHASE == 0)
	{
#pragma unroll
		for (int i = 0; i < 4; i++)
		{
			r.j0[i] = tn[(1 + 2 * i) * 32];
			r.j1[i] = tn[(2 + 2 * i) * 32];
		}
	}
	else
	{
		const float4* tf = tilesF + (size_t)tile * FT_STRIDE + lane;
		r.j0[0] = tf[0];
		r.j1[0] = tf[32];
		r.j1[1] = tf[64];
		r.fl = tf[FT_LAMBDA * 32];
	}
}
template <int PHASE>
B3_D bool solveRowCore(RowRegs<PHASE>& r, const BodyConst& A, const BodyConst& B, float4& linVelA, float4& angVelA, float4& linVelB, float4& angVelB)
{
	if (PHASE == 0)
	{
		solveNormalCore(r.nId, r.j0, r.j1, r.applied, A, B, linVelA, angVelA, linVelB, angVelB);
		return true;
	}
	return solveFrictionCore(r.nId, r.j0[0], r.j1[0], r.j1[PHASE == 0 ? 0 : 1], r.fl, r.applied, A, B, linVelA, angVelA, linVelB, angVelB);
}
template <int PHASE>
B3_D void storeRowLambda(float4* tilesN, float4* tilesF, unsigned int tile, int lane, const RowRegs<PHASE>& r)
{
	if (PHASE == 0)
		tilesN[(size_t)tile * NT_STRIDE + NT_LAMBDA * 32 + lane] = r.applied;
	else
		tilesF[(size_t)tile * FT_STRIDE + FT_LAMBDA * 32 + lane] = r.fl;
}

// The CTA's shared-memory state: one 16-byte array per quantity (a warp gathers 32 random slots per load: with 16-byte
// strides the eight lanes of a quarter-warp can hit all 32 banks, with the 32-byte {lin, ang} records of the global
// array only half of them -- shared-memory wavefronts are what bounds the interior loop).  The world inverse inertia is
// kept as its upper triangle (R diag R^T is symmetric up to rounding): 2 loads per body instead of 4 with the position.
struct SharedBodies
{
	float4* lin;        // linear velocity
	float4* ang;        // angular velocity
	float4* inerA;      // {Ixx, Ixy, Ixz, Iyy}
	float4* inerB;      // {Iyz, Izz, inverse mass, -}
	float4* pos;        // position (friction rows: r = centre - pos)
};
// interior row: bodies are slots of the CTA's shared-memory state
template <int PHASE>
B3_D void solveRowShared(float4* tilesN, float4* tilesF, unsigned int tile, int lane, RowRegs<PHASE>& r, const SharedBodies& sb_)
{
	const unsigned int ids = __float_as_uint(r.nId.w);
	if (ids == INVALID_IDS) return;
	const int sa = (int)(ids & 0xffffu), sb = (int)(ids >> 16);
	BodyConst A, B;
	{
		const float4 a0 = sb_.inerA[sa], a1 = sb_.inerB[sa], b0 = sb_.inerA[sb], b1 = sb_.inerB[sb];
		A.i0 = mk4(a0.x, a0.y, a0.z), A.i1 = mk4(a0.y, a0.w, a1.x), A.i2 = mk4(a0.z, a1.x, a1.y);
		B.i0 = mk4(b0.x, b0.y, b0.z), B.i1 = mk4(b0.y, b0.w, b1.x), B.i2 = mk4(b0.z, b1.x, b1.y);
		A.pos = mk4(0, 0, 0, a1.z);
		B.pos = mk4(0, 0, 0, b1.z);
		if (PHASE == 1)
		{
			const float4 pa = sb_.pos[sa], pb = sb_.pos[sb];
			A.pos = mk4(pa.x, pa.y, pa.z, a1.z);
			B.pos = mk4(pb.x, pb.y, pb.z, b1.z);
		}
	}
	float4 linVelA = sb_.lin[sa], angVelA = sb_.ang[sa];
	float4 linVelB = sb_.lin[sb], angVelB = sb_.ang[sb];
	if (!solveRowCore<PHASE>(r, A, B, linVelA, angVelA, linVelB, angVelB)) return;
	storeRowLambda<PHASE>(tilesN, tilesF, tile, lane, r);
	if (A.pos.w != 0.f)
	{
		sb_.lin[sa] = linVelA;
		sb_.ang[sa] = angVelA;
	}
	if (B.pos.w != 0.f)
	{
		sb_.lin[sb] = linVelB;
		sb_.ang[sb] = angVelB;
	}
}

// cross row: bodies through global memory (L2).  Everything but the row itself is fetched AFTER the grid barrier: the
// velocities have to be (one L2 round trip), and the constants of the two bodies ride along in the same round trip.
struct CrossBodies
{
	float4 *gLin, *gAng;
	const float4 *gInerA, *gInerB;
	const float4* pose;  // bodies without a rank (static): original arrays
	const float4* vel;
	const b3b200_inertia* inertias;
};
B3_D void loadCrossBody(const CrossBodies& cb, int id, BodyConst& C, float4& lin, float4& ang)
{
	if (id >= 0)
	{
		lin = __ldcg(&cb.gLin[id]);
		ang = __ldcg(&cb.gAng[id]);
		const float4 a0 = __ldg(&cb.gInerA[id]), a1 = __ldg(&cb.gInerB[id]);
		C.i0 = mk4(a0.x, a0.y, a0.z), C.i1 = mk4(a0.y, a0.w, a1.x), C.i2 = mk4(a0.z, a1.x, a1.y);
		C.pos = mk4(0, 0, 0, a1.z);
	}
	else
	{
		const int g = -id - 2;
		lin = cb.vel[2 * g];
		ang = cb.vel[2 * g + 1];
		C.pos = cb.pose[2 * g];
		const float4* I = reinterpret_cast<const float4*>(&cb.inertias[g].invInertiaWorld);
		C.i0 = __ldg(I), C.i1 = __ldg(I + 1), C.i2 = __ldg(I + 2);
		C.pos.w = 0.f;  // never written back
	}
}
template <int PHASE>
B3_D void solveRowGlobal(float4* tilesN, float4* tilesF, const CrossBodies& cb, const float4* gPos, unsigned int tile, int lane, RowRegs<PHASE>& r, const int4& tail)
{
	const int a = tail.x, b = tail.y;
	if (tail.w < 0) return;  // padding
	BodyConst A, B;
	float4 linVelA, angVelA, linVelB, angVelB;
	loadCrossBody(cb, a, A, linVelA, angVelA);
	loadCrossBody(cb, b, B, linVelB, angVelB);
	if (PHASE == 1)
	{
		// friction rows need the positions (r = centre - pos)
		const float4 pa = a >= 0 ? __ldg(&gPos[a]) : cb.pose[2 * (-a - 2)], pb = b >= 0 ? __ldg(&gPos[b]) : cb.pose[2 * (-b - 2)];
		A.pos = mk4(pa.x, pa.y, pa.z, A.pos.w);
		B.pos = mk4(pb.x, pb.y, pb.z, B.pos.w);
	}
	if (!solveRowCore<PHASE>(r, A, B, linVelA, angVelA, linVelB, angVelB)) return;
	storeRowLambda<PHASE>(tilesN, tilesF, tile, lane, r);
	if (A.pos.w != 0.f)
	{
		__stcg(&cb.gLin[a], linVelA);
		__stcg(&cb.gAng[a], angVelA);
	}
	if (B.pos.w != 0.f)
	{
		__stcg(&cb.gLin[b], linVelB);
		__stcg(&cb.gAng[b], angVelB);
	}
}
B3_D int4 loadTail(const float4* tilesN, unsigned int tile, int lane) { return reinterpret_cast<const int4*>(tilesN)[(size_t)tile * NT_STRIDE + NT_TAIL * 32 + lane]; }

// The rows of a step (~150 MB on the bench scene) do not stay in L2 between passes, and a warp holds only one tile ahead in
// registers: with the row solve down to a few hundred cycles, a tile's HBM latency would be exposed at every tile.  One
// lane asks for the whole tile (one contiguous range) to be pulled into L2 a few tiles ahead; the register loads then hit L2.
B3_D void prefetchL2(const void* p, unsigned int bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
template <int PHASE>
B3_D void prefetchTile(const float4* tilesN, const float4* tilesF, unsigned int tile)
{
	const float4* tn = tilesN + (size_t)tile * NT_STRIDE;
	if (PHASE == 0)
		prefetchL2(tn, NT_TAIL * 512);
	else
	{
		prefetchL2(tn, 512);
		prefetchL2(tn + NT_LAMBDA * 32, 512);
		prefetchL2(tilesF + (size_t)tile * FT_STRIDE, FT_FIELDS * 512);
	}
}
constexpr int PF_AHEAD = 4;  // tiles (of the same warp) between the L2 prefetch and the solve

struct BlockView
{
	int blk;
	int count;        // dynamic bodies in the block
	unsigned int tileBase;
	unsigned int numTiles;
	int numColours;
};

// all colours of one block, velocities in shared memory.  Warp w owns the tiles w, w + ITER_WARPS, ... of the block's tile
// sequence (the same warp every pass, so a tile's lambdas are read back by the thread that wrote them) and fetches its
// next tile into a second register set while it solves the current one; one __syncthreads() per colour.
template <int PHASE, int ITER_WARPS>
B3_D void solveBlockInterior(float4* tilesN, float4* tilesF, unsigned int tileBase, unsigned int numTiles, int numColours,
												const unsigned int* sTileOff, SharedBodies sb_, unsigned long long* probe)
{
	int probeN = 256;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned int T = (unsigned int)warp;
	int m = 0;
	RowRegs<PHASE> ra, rb;
	if (lane == 0)
		for (int j = 2; j < PF_AHEAD; j++)
			if (T + j * ITER_WARPS < numTiles) prefetchTile<PHASE>(tilesN, tilesF, tileBase + T + j * ITER_WARPS);
	if (T < numTiles) loadRowRegs<PHASE>(tilesN, tilesF, tileBase + T, lane, ra);
	while (T < numTiles)
	{
		unsigned int T2 = T + ITER_WARPS;
		if (T2 < numTiles) loadRowRegs<PHASE>(tilesN, tilesF, tileBase + T2, lane, rb);
		if (lane == 0)
		{
			if (T + PF_AHEAD * ITER_WARPS < numTiles) prefetchTile<PHASE>(tilesN, tilesF, tileBase + T + PF_AHEAD * ITER_WARPS);
			if (T + (PF_AHEAD + 1) * ITER_WARPS < numTiles) prefetchTile<PHASE>(tilesN, tilesF, tileBase + T + (PF_AHEAD + 1) * ITER_WARPS);
		}

		while (T >= sTileOff[m + 1])
		{
			__syncthreads();
			m++;
			if (probe && threadIdx.x == 0 && probeN < 500)
			{
				unsigned long long t_;
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
				probe[probeN++] = (t_ << 8) | (unsigned long long)m;
			}
		}
		if (probe && threadIdx.x == 0 && probeN < 500)
		{
			unsigned long long t_;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
			probe[probeN++] = (t_ << 8) | 200ull;
		}
		solveRowShared<PHASE>(tilesN, tilesF, tileBase + T, lane, ra, sb_);
		if (probe && threadIdx.x == 0 && probeN < 500)
		{
			unsigned long long t_;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
			probe[probeN++] = (t_ << 8) | 201ull;
		}
		T = T2;
		if (T >= numTiles) break;
		T2 = T + ITER_WARPS;
		if (T2 < numTiles) loadRowRegs<PHASE>(tilesN, tilesF, tileBase + T2, lane, ra);
		while (T >= sTileOff[m + 1])
		{
			__syncthreads();
			m++;
		}
		solveRowShared<PHASE>(tilesN, tilesF, tileBase + T, lane, rb, sb_);
		T = T2;
	}
	for (; m < numColours; m++) __syncthreads();
}

// cross colours of one pass: global velocities, a grid barrier per colour; the next colour's rows are fetched while the
// other CTAs arrive.  Trailing colours of at most one tile per warp of a CTA are solved by CTA 0 alone between __syncthreads().
template <int PHASE, int ITER_WARPS>
B3_D void solveCrossColours(float4* tilesN, float4* tilesF, CrossBodies cb, const float4* gPos, GridBarrier& bar, int Kc, int tailStart,
											   const unsigned int* sCrossOff, unsigned int crossBase)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned int gw = (unsigned int)warp * gridDim.x + blockIdx.x;  // tiles are dealt round-robin to the CTAs
	const unsigned int gstride = ITER_WARPS * gridDim.x;
	RowRegs<PHASE> pre;
	int4 tail;
	// this warp's first tile of colour 0, fetched while the other CTAs arrive
	unsigned int t = sCrossOff[0] + gw;
	if (tailStart > 0 && t < sCrossOff[1])
	{
		loadRowRegs<PHASE>(tilesN, tilesF, crossBase + t, lane, pre);
		tail = loadTail(tilesN, crossBase + t, lane);
	}
	bar.wait();  // (the caller arrived after publishing this CTA's velocities)
	for (int k = 0; k < tailStart; k++)
	{
		if (t < sCrossOff[k + 1])
		{
			solveRowGlobal<PHASE>(tilesN, tilesF, cb, gPos, crossBase + t, lane, pre, tail);
			// (a colour with more tiles than the grid has warps: the rest without the early fetch)
			for (t += gstride; t < sCrossOff[k + 1]; t += gstride)
			{
				loadRowRegs<PHASE>(tilesN, tilesF, crossBase + t, lane, pre);
				tail = loadTail(tilesN, crossBase + t, lane);
				solveRowGlobal<PHASE>(tilesN, tilesF, cb, gPos, crossBase + t, lane, pre, tail);
			}
		}
		bar.arrive();
		t = sCrossOff[k + 1] + gw;
		if (k + 1 < tailStart && t < sCrossOff[k + 2])
		{
			loadRowRegs<PHASE>(tilesN, tilesF, crossBase + t, lane, pre);
			tail = loadTail(tilesN, crossBase + t, lane);
		}
		bar.wait();
	}
	if (tailStart < Kc)
	{
		if (blockIdx.x == 0)
		{
			for (int k = tailStart; k < Kc; k++)
			{
				t = sCrossOff[k] + (unsigned int)warp;
				if (t < sCrossOff[k + 1])
				{
					loadRowRegs<PHASE>(tilesN, tilesF, crossBase + t, lane, pre);
					tail = loadTail(tilesN, crossBase + t, lane);
					solveRowGlobal<PHASE>(tilesN, tilesF, cb, gPos, crossBase + t, lane, pre, tail);
				}
				__syncthreads();
			}
		}
		bar.sync();
	}
}

#define B3_PROBE(i)                                                              \
	do                                                                           \
	{                                                                            \
		if (s.probe && blockIdx.x == 0 && threadIdx.x == 0 && probeN < 250)      \
		{                                                                        \
			unsigned long long t_;                                               \
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));               \
			s.probe[probeN++] = (t_ << 4) | (unsigned long long)(i);             \
		}                                                                        \
	} while (0)
// One launch per phase (all iterations of the normal rows, then all iterations of the friction rows): a kernel that holds only
// one phase's code keeps its row registers without spilling -- with ~6 KB of L1 left beside the block state, every spill
// is an L2 round trip on the critical path.
template <int PHASE, int ITER_THREADS>
__global__ void __launch_bounds__(ITER_THREADS, 1) solverIterateKernel(IterArgs s)
{
	constexpr int ITER_WARPS = ITER_THREADS / 32;
	constexpr int TAIL_TILES = ITER_WARPS;  // cross colours this small are solved by CTA 0 alone (cheaper than a grid barrier)
	int probeN = PHASE * 100;
	extern __shared__ float4 smem4[];
	__shared__ unsigned int sTileOff[MAX_BATCHES + 1], sCrossOff[MAX_BATCHES + 1];
	__shared__ int sNumBoundary;
	const int slots = s.S + NSTATIC;
	SharedBodies sb_;
	sb_.lin = smem4;
	sb_.ang = smem4 + slots;
	sb_.inerA = smem4 + 2 * slots;
	sb_.inerB = smem4 + 3 * slots;
	sb_.pos = smem4 + 4 * slots;
	unsigned short* sBoundary = reinterpret_cast<unsigned short*>(smem4 + 5 * slots);  // slots of the bodies that have cross contacts
	GridBarrier bar;
	bar.init(s.bar, gridDim.x);

	const int numDyn = (int)s.partBounds[6];
	const int numBlocks = (numDyn + s.S - 1) / s.S;
	const int Kc = (int)s.misc[MISC_KC];
	const bool resident = numBlocks <= (int)gridDim.x;
	for (int i = threadIdx.x; i <= MAX_BATCHES; i += ITER_THREADS) sCrossOff[i] = s.crossTileOff[i];
	if (threadIdx.x == 0) sNumBoundary = 0;
	__syncthreads();
	const unsigned int crossBase = s.tileCap - sCrossOff[MAX_BATCHES];
	int tailStart = Kc;
	while (tailStart > 0 && sCrossOff[tailStart] - sCrossOff[tailStart - 1] <= (unsigned int)TAIL_TILES) tailStart--;
	CrossBodies cb;
	cb.gLin = s.gLin;
	cb.gAng = s.gAng;
	cb.gInerA = s.gInerA;
	cb.gInerB = s.gInerB;
	cb.pose = s.pose;
	cb.vel = s.vel;
	cb.inertias = s.inertias;

	auto view = [&](int blk) {
		BlockView v;
		v.blk = blk;
		v.count = min(s.S, numDyn - blk * s.S);
		v.tileBase = s.blockTileBase[blk];
		for (int i = threadIdx.x; i <= MAX_BATCHES; i += ITER_THREADS) sTileOff[i] = s.blockTileOff[(size_t)blk * (MAX_BATCHES + 1) + i];
		__syncthreads();
		v.numTiles = sTileOff[MAX_BATCHES];
		int kc = 0;
		// colours in use = last colour whose tile range is not empty
		for (int m = MAX_BATCHES; m > 0; m--)
			if (sTileOff[m] != sTileOff[m - 1])
			{
				kc = m;
				break;
			}
		v.numColours = kc;
		return v;
	};
	// the block's state: coalesced copies out of the block-ordered arrays (+ the few static bodies the block refers to, from the
	// body arrays); lists the boundary slots
	auto loadBlock = [&](const BlockView& v, bool listBoundary) {
		const size_t first = (size_t)v.blk * s.S;
		for (int k = threadIdx.x; k < v.count; k += ITER_THREADS)
		{
			sb_.lin[k] = __ldcg(&s.gLin[first + k]);
			sb_.ang[k] = __ldcg(&s.gAng[first + k]);
			sb_.inerA[k] = __ldg(&s.gInerA[first + k]);
			sb_.inerB[k] = __ldg(&s.gInerB[first + k]);
			sb_.pos[k] = __ldg(&s.gPos[first + k]);
			if (listBoundary && s.gBoundary[first + k]) sBoundary[atomicAdd(&sNumBoundary, 1)] = (unsigned short)k;
		}
		for (int k = threadIdx.x; k < NSTATIC; k += ITER_THREADS)
		{
			const int g = s.blockStatics[(size_t)v.blk * NSTATIC + k];
			if (g < 0) continue;
			const float4 ps = s.pose[2 * g];
			const float4* I = reinterpret_cast<const float4*>(&s.inertias[g].invInertiaWorld);
			const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
			sb_.pos[s.S + k] = ps;
			sb_.inerA[s.S + k] = mk4(r0.x, r0.y, r0.z, r1.y);
			sb_.inerB[s.S + k] = mk4(r1.z, r2.z, 0.f, 0.f);  // a static slot is never written back
			sb_.lin[s.S + k] = s.vel[2 * g];
			sb_.ang[s.S + k] = s.vel[2 * g + 1];
		}
	};
	auto storeBlock = [&](const BlockView& v) {
		const size_t first = (size_t)v.blk * s.S;
		for (int k = threadIdx.x; k < v.count; k += ITER_THREADS)
		{
			__stcg(&s.gLin[first + k], sb_.lin[k]);
			__stcg(&s.gAng[first + k], sb_.ang[k]);
		}
	};

	B3_PROBE(0);
	BlockView mine;
	mine.blk = -1;
	mine.count = 0;
	mine.tileBase = 0;
	mine.numTiles = 0;
	mine.numColours = 0;
	if (resident && (int)blockIdx.x < numBlocks)
	{
		mine = view((int)blockIdx.x);
		loadBlock(mine, true);
		__syncthreads();
	}

	{
#pragma unroll 1
		for (int iter = 0; iter < s.iterations; iter++)
		{
			if (Kc > 0)
			{
				// the velocities the cross rows need are in global memory after this: resident blocks publish their boundary
				// bodies (the others never left global memory)
				B3_PROBE(1);
				if (mine.blk >= 0)
				{
					const int nB = sNumBoundary;
					for (int i = threadIdx.x; i < nB; i += ITER_THREADS)
					{
						const int k = sBoundary[i];
						__stcg(&s.gLin[(size_t)mine.blk * s.S + k], sb_.lin[k]);
						__stcg(&s.gAng[(size_t)mine.blk * s.S + k], sb_.ang[k]);
					}
				}
				B3_PROBE(2);
				bar.arrive();
				solveCrossColours<PHASE, ITER_WARPS>(s.tilesN, s.tilesF, cb, s.gPos, bar, Kc, tailStart, sCrossOff, crossBase);
				B3_PROBE(3);
				if (mine.blk >= 0)
				{
					const int nB = sNumBoundary;
					for (int i = threadIdx.x; i < nB; i += ITER_THREADS)
					{
						const int k = sBoundary[i];
						sb_.lin[k] = __ldcg(&s.gLin[(size_t)mine.blk * s.S + k]);
						sb_.ang[k] = __ldcg(&s.gAng[(size_t)mine.blk * s.S + k]);
					}
					__syncthreads();
				}
				B3_PROBE(4);
			}
			if (resident)
			{
				if (mine.blk >= 0)
				{
					solveBlockInterior<PHASE, ITER_WARPS>(s.tilesN, s.tilesF, mine.tileBase, mine.numTiles, mine.numColours, sTileOff, sb_, (PHASE == 0 && blockIdx.x == 0 && iter == 1) ? s.probe : nullptr);
				}
			}
			else
			{
				for (int blk = blockIdx.x; blk < numBlocks; blk += gridDim.x)
				{
					__syncthreads();
					const BlockView v = view(blk);
					loadBlock(v, false);
					__syncthreads();
					solveBlockInterior<PHASE, ITER_WARPS>(s.tilesN, s.tilesF, v.tileBase, v.numTiles, v.numColours, sTileOff, sb_, nullptr);
					storeBlock(v);
				}
			}
		}
	}
	B3_PROBE(5);
	if (mine.blk >= 0)
	{
		__syncthreads();
		storeBlock(mine);
	}
	B3_PROBE(6);
}

// ================================================================ export (b3b200_get_constraints)
// tiles -> b3ContactConstraint4 records (reference layout), compacted with one atomic cursor; the host sorts them by batch
__global__ void __launch_bounds__(256) solverExportKernel(const float4* __restrict__ tilesN, const float4* __restrict__ tilesF, unsigned int tileCap,
														  const unsigned int* __restrict__ misc, const unsigned int* __restrict__ crossTileOff,
														  const b3b200_contact4* __restrict__ contacts, b3b200_constraint4* __restrict__ out,
														  unsigned int* __restrict__ cursor, unsigned int capacity)
{
	const unsigned int interiorTiles = misc[MISC_TILE_CURSOR];
	const unsigned int crossTiles = crossTileOff[MAX_BATCHES];
	const unsigned int total = interiorTiles + crossTiles;
	const int lane = threadIdx.x & 31;
	for (unsigned int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < total; t += gridDim.x * (blockDim.x >> 5))
	{
		const unsigned int tile = t < interiorTiles ? t : tileCap - crossTiles + (t - interiorTiles);
		const float4* tn = tilesN + (size_t)tile * NT_STRIDE + lane;
		const float4* tf = tilesF + (size_t)tile * FT_STRIDE + lane;
		const int4 tail = reinterpret_cast<const int4*>(tn)[NT_TAIL * 32];
		const bool valid = tail.w >= 0;
		const unsigned int m = __ballot_sync(0xffffffffu, valid);
		if (!m) continue;
		unsigned int slot = 0;
		if (lane == 0) slot = atomicAdd(cursor, (unsigned int)__popc(m));
		slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
		if (!valid || slot >= capacity) continue;
		float4* dw = reinterpret_cast<float4*>(&out[slot]);
		const float4 nId = tn[0];
		dw[0] = mk4(nId.x, nId.y, nId.z, 0.7f);
		// m_worldPos: the points of the contact the row was built from (the tiles keep the angular Jacobians instead)
		const float4* cw = reinterpret_cast<const float4*>(&contacts[tail.w]);
		const float npoints = cw[4].w;
		float jac[4], bb[4];
#pragma unroll
		for (int i = 0; i < 4; i++)
		{
			jac[i] = tn[(1 + 2 * i) * 32].w;
			bb[i] = tn[(2 + 2 * i) * 32].w;
			const float4 p = cw[i];
			dw[1 + i] = ((float)i < npoints) ? mk4(p.x, p.y, p.z, 0.f) : mk4(0, 0, 0, 0);
		}
		const float4 cf = tf[0], t0 = tf[32], t1 = tf[64], fl = tf[FT_LAMBDA * 32];
		dw[5] = mk4(cf.x, cf.y, cf.z, 0.f);
		dw[6] = mk4(jac[0], jac[1], jac[2], jac[3]);
		dw[7] = mk4(bb[0], bb[1], bb[2], bb[3]);
		dw[8] = tn[NT_LAMBDA * 32];
		dw[9] = mk4(t0.w, t1.w, fl.x, fl.y);
		const int4 cids = reinterpret_cast<const int4*>(&contacts[tail.w])[5];
		int4 o;
		o.x = abs(cids.z);  // (the tiles address the bodies by their rank in block order)
		o.y = abs(cids.w);
		o.z = tail.z;
		o.w = 0;
		reinterpret_cast<int4*>(dw)[10] = o;
	}
}

// ================================================================ launchers
static size_t tileCapacity(const World* w, int blocksMax)
{
	const size_t nc = (size_t)std::max(w->cfg.maxContactCapacity, 1);
	return nc / 32 + std::min(nc, (size_t)blocksMax * MAX_BATCHES) + MAX_BATCHES + 2;
}

static int fillSetupArgs(World* w, SetupArgs& s)
{
	const int B = w->partBlocksMax;
	const size_t nc = (size_t)std::max(w->cfg.maxContactCapacity, 1);
	const size_t nb = (size_t)std::max(w->numBodies, 1);
	const size_t tiles = tileCapacity(w, B);
	B3_TRY(w->dTilesN.reserve(tiles * NT_STRIDE));
	B3_TRY(w->dTilesF.reserve(tiles * FT_STRIDE));
	B3_TRY(w->dContactBlock.reserve(nc));
	B3_TRY(w->dContactSlots.reserve(nc));
	B3_TRY(w->dContactColour.reserve(nc));
	B3_TRY(w->dContactPair.reserve(nc));
	B3_TRY(w->dTileSrc.reserve(tiles * 32));
	B3_TRY(w->dBlockList.reserve(nc));
	B3_TRY(w->dCrossList.reserve(nc));
	B3_TRY(w->dBodyMask.reserve(2 * nb));
	B3_TRY(w->dBodyPrio.reserve(2 * nb));
	B3_TRY(w->dSolverScratch.reserve((size_t)2 * B + 2 * MAX_BATCHES + MISC_NUM));
	B3_TRY(w->dCtaBase.reserve((size_t)w->smCount * 4 * ((size_t)B + 1)));  // (smCount * 4 = the grid of the classify / scatter kernels)
	B3_TRY(w->dBlockStatics.reserve((size_t)B * NSTATIC));
	B3_TRY(w->dBlockStart.reserve((size_t)B + 1));
	B3_TRY(w->dBlockTileBase.reserve((size_t)B));
	B3_TRY(w->dBlockTileOff.reserve((size_t)B * (MAX_BATCHES + 1)));
	B3_TRY(w->dCrossTileOff.reserve(MAX_BATCHES + 1));
	s.contacts = w->dContacts.ptr;
	s.ctr = w->dCounters.ptr;
	s.pose = w->dPose.ptr;
	s.inertias = w->dInertias.ptr;
	s.bodyLoc = w->dBodyLoc.ptr;
	s.bodyMask = w->dBodyMask.ptr;
	s.bodyPrio = reinterpret_cast<unsigned long long*>(w->dBodyPrio.ptr);
	s.contactBlock = w->dContactBlock.ptr;
	s.contactSlots = w->dContactSlots.ptr;
	s.contactColour = w->dContactColour.ptr;
	s.contactPair = w->dContactPair.ptr;
	{
		const size_t ranks = (size_t)B * (size_t)w->partS;
		B3_TRY(w->dSortedState.reserve(5 * ranks));
		B3_TRY(w->dSortedBoundary.reserve(ranks));
		s.gLin = w->dSortedState.ptr;
		s.gAng = s.gLin + ranks;
		s.gInerA = s.gAng + ranks;
		s.gInerB = s.gInerA + ranks;
		s.gPos = s.gInerB + ranks;
		s.gBoundary = w->dSortedBoundary.ptr;
	}
	s.partVals = w->dPartVals.ptr;
	s.partBounds = w->dPartBounds.ptr;
	s.vel = w->dVel.ptr;
	s.tileSrc = w->dTileSrc.ptr;
	unsigned int* scr = w->dSolverScratch.ptr;
	s.blockCount = scr;
	s.blockCursor = scr + B;
	s.crossHist = scr + 2 * B;
	s.ctaBase = w->dCtaBase.ptr;
	s.crossCursor = scr + 2 * B + MAX_BATCHES;
	s.misc = scr + 2 * B + 2 * MAX_BATCHES;
	s.blockStatics = w->dBlockStatics.ptr;
	s.blockStart = w->dBlockStart.ptr;
	s.blockList = w->dBlockList.ptr;
	s.crossList = w->dCrossList.ptr;
	s.blockTileBase = w->dBlockTileBase.ptr;
	s.blockTileOff = w->dBlockTileOff.ptr;
	s.crossTileOff = w->dCrossTileOff.ptr;
	s.tilesN = w->dTilesN.ptr;
	s.tilesF = w->dTilesF.ptr;
	s.tileCap = (unsigned int)tiles;
	s.numBodies = w->numBodies;
	s.numBlocksMax = B;
	s.S = w->partS;
	s.staticIdx = w->static0Index;
	s.colouring = w->solverColouring;
	s.dt = 1.f / 60.f;  // the reference solver ignores deltaTime (b3GpuPgsContactSolver.cpp:672)
	s.positionDrift = 0.005f;
	s.positionConstraintCoeff = 0.2f;
	return 0;
}

int launchSolverSetup(World* w)
{
	B3_TRY(ensurePartition(w));
	SetupArgs s;
	B3_TRY(fillSetupArgs(w, s));
	cudaStream_t st = w->stream;
	const int B = w->partBlocksMax;
	B3_CUDA_CHECK(cudaMemsetAsync(w->dSolverScratch.ptr, 0, sizeof(unsigned int) * ((size_t)2 * B + 2 * MAX_BATCHES + MISC_NUM), st));
	B3_CUDA_CHECK(cudaMemsetAsync(w->dBlockStatics.ptr, 0xff, sizeof(int) * (size_t)B * NSTATIC, st));
	const int grid = w->smCount * 4;
	solverClassifyKernel<<<grid, SETUP_THREADS, sizeof(unsigned int) * ((size_t)B + 1), st>>>(s);
	B3_LAUNCH_CHECK();
	solverScatterKernel<<<grid, SETUP_THREADS, sizeof(unsigned int) * 2 * ((size_t)B + 1), st>>>(s);
	B3_LAUNCH_CHECK();
	if (w->solverColouring == 0)
	{
		solverCrossColourKernel<<<1, 1024, 0, st>>>(s);
		B3_LAUNCH_CHECK();
	}
	{
		const int slots = w->partS + NSTATIC;
		size_t smem = sizeof(unsigned long long) * 2 * (size_t)slots;
		if (w->solverColouring == 0) smem += sizeof(unsigned long long) * (size_t)slots + (size_t)JP_CONTACT_CAP * 5 + 16;
		if (!w->solverAttrSet)
		{
			B3_CUDA_CHECK(cudaFuncSetAttribute(solverBlockSetupKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
			B3_CUDA_CHECK(cudaFuncSetAttribute(solverIterateKernel<0, ITER_THREADS_NORMAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ITER_SMEM_MAX)));
			B3_CUDA_CHECK(cudaFuncSetAttribute(solverIterateKernel<1, ITER_THREADS_FRICTION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ITER_SMEM_MAX)));
			w->solverAttrSet = true;
		}
		solverBlockSetupKernel<<<std::min(B, w->smCount * 2), SETUP_THREADS, smem, st>>>(s);
		B3_LAUNCH_CHECK();
	}
	solverCrossBuildKernel<<<grid, SETUP_THREADS, 0, st>>>(s);
	B3_LAUNCH_CHECK();
	solverBuildRowsKernel<<<w->smCount * 6, 256, 0, st>>>(s);
	B3_LAUNCH_CHECK();
	{
		const int numRanks = B * w->partS;
		solverGatherKernel<<<divUp(numRanks, 256), 256, 0, st>>>(s, numRanks);
		B3_LAUNCH_CHECK();
	}
	w->solverMisc = s.misc;
	return 0;
}

int launchSolverIterate(World* w)
{
	if (!w->partValid || !w->solverMisc)
	{
		setLastError("solver_iterate without solver_setup");
		return B3B200_ERR_STATE;
	}
	IterArgs s;
	s.tilesN = w->dTilesN.ptr;
	s.tilesF = w->dTilesF.ptr;
	s.tileCap = (unsigned int)tileCapacity(w, w->partBlocksMax);
	s.ctr = w->dCounters.ptr;
	s.pose = w->dPose.ptr;
	s.vel = w->dVel.ptr;
	s.inertias = w->dInertias.ptr;
	{
		const size_t ranks = (size_t)w->partBlocksMax * (size_t)w->partS;
		s.gLin = w->dSortedState.ptr;
		s.gAng = s.gLin + ranks;
		s.gInerA = s.gAng + ranks;
		s.gInerB = s.gInerA + ranks;
		s.gPos = s.gInerB + ranks;
		s.gBoundary = w->dSortedBoundary.ptr;
	}
	s.partBounds = w->dPartBounds.ptr;
	s.blockStatics = w->dBlockStatics.ptr;
	s.blockTileBase = w->dBlockTileBase.ptr;
	s.blockTileOff = w->dBlockTileOff.ptr;
	s.crossTileOff = w->dCrossTileOff.ptr;
	s.misc = w->solverMisc;
	s.bodyMask = w->dBodyMask.ptr;
	s.bar = w->dGridBarrier.ptr;
	s.iterations = w->solverIterations;
	s.S = w->partS;
	s.numBlocksMax = w->partBlocksMax;
	s.probe = w->dSolverProbe.ptr;  // nullptr unless b3b200_debug_solver_probe armed it
	w->soaDirty = true;
	const size_t smem = (size_t)ITER_SMEM_PER_SLOT * (size_t)(w->partS + NSTATIC);
	// one CTA per SM at most (cooperative: all CTAs co-resident); small worlds use as many CTAs as they have blocks
	const int grid = std::max(1, std::min(w->smCount, w->partBlocksMax));
	dim3 g(grid), b0(ITER_THREADS_NORMAL), b1(ITER_THREADS_FRICTION);
	void* args[] = {&s};
	B3_CUDA_CHECK(cudaMemsetAsync(w->dGridBarrier.ptr, 0, sizeof(unsigned int) * 4, w->stream));
	B3_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)solverIterateKernel<0, ITER_THREADS_NORMAL>, g, b0, args, smem, w->stream));
	g_launchCount++;
	B3_CUDA_CHECK(cudaMemsetAsync(w->dGridBarrier.ptr, 0, sizeof(unsigned int) * 4, w->stream));
	B3_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)solverIterateKernel<1, ITER_THREADS_FRICTION>, g, b1, args, smem, w->stream));
	g_launchCount++;
	{
		const int n = std::max(w->numBodies, 1);
		solverScatterBackKernel<<<divUp(n, 256), 256, 0, w->stream>>>(s.gLin, s.gAng, s.gInerB, w->dPartVals.ptr, w->dPartBounds.ptr, w->dVel.ptr);
		B3_LAUNCH_CHECK();
	}
	return 0;
}

// b3b200_get_constraints: rows in the reference layout, sorted by batch on the host
int exportConstraints(World* w, std::vector<b3b200_constraint4>& out, std::vector<int>& batchOffsets)
{
	out.clear();
	batchOffsets.assign(1, 0);
	if (!w->solverMisc) return 0;
	unsigned int nContacts = 0;
	B3_CUDA_CHECK(cudaMemcpyAsync(&nContacts, &w->dCounters.ptr[CTR_CONTACTS], sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	if (nContacts == 0) return 0;
	DevBuf<b3b200_constraint4> tmp;
	DevBuf<unsigned int> cursor;
	B3_TRY(tmp.reserve(nContacts));
	B3_TRY(cursor.reserve(1));
	B3_CUDA_CHECK(cudaMemsetAsync(cursor.ptr, 0, sizeof(unsigned int), w->stream));
	solverExportKernel<<<w->smCount * 4, 256, 0, w->stream>>>(w->dTilesN.ptr, w->dTilesF.ptr, (unsigned int)tileCapacity(w, w->partBlocksMax), w->solverMisc,
															 w->dCrossTileOff.ptr, w->dContacts.ptr, tmp.ptr, cursor.ptr, nContacts);
	B3_LAUNCH_CHECK();
	unsigned int n = 0;
	B3_CUDA_CHECK(cudaMemcpyAsync(&n, cursor.ptr, sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	n = std::min(n, nContacts);
	std::vector<b3b200_constraint4> raw(n);
	if (n) B3_CUDA_CHECK(cudaMemcpy(raw.data(), tmp.ptr, sizeof(b3b200_constraint4) * n, cudaMemcpyDeviceToHost));
	int nb = 0;
	for (unsigned int i = 0; i < n; i++) nb = std::max(nb, raw[i].batchIdx + 1);
	std::vector<int> count(nb + 1, 0);
	for (unsigned int i = 0; i < n; i++) count[raw[i].batchIdx + 1]++;
	for (int b = 0; b < nb; b++) count[b + 1] += count[b];
	batchOffsets.assign(count.begin(), count.end());
	out.resize(n);
	std::vector<int> cur(count.begin(), count.end() - 1);
	for (unsigned int i = 0; i < n; i++) out[cur[raw[i].batchIdx]++] = raw[i];
	return 0;
}

}  // namespace b3b200
