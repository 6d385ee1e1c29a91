// broadphase.cu -- pair finding on sm_100a.
//
// Two algorithms behind the b3GpuBroadphaseInterface contract
// (b3GpuBroadphaseInterface.h:12-40):
//   GRID : uniform-grid cell hash (128^3 wrapped, like gridBroadphase.cl:13-21)
//          -> stable KV radix sort -> cell starts -> 27-cell scan.  Unlike the
//          reference grid (which silently drops pairs whose AABB is wider than
//          the fixed 3.0 cell) the cell edge is re-derived
//          every call from the widest small AABB, so the pair SET equals the
//          brute-force host twin calculateOverlappingPairsHost
//          (b3GpuSapBroadphase.cpp:862-981) bit for bit.
//   SAP  : 1-axis sweep on the axis of largest centre variance
//          (b3GpuSapBroadphase.cpp:997-1231): FloatFlip keys -> sort -> gather ->
//          forward sweep.
// Static ("large") proxies are tested brute force against all small ones
// (computePairsKernelTwoArrays, sap.cl:66).
//
// Pairs are staged per warp in shared memory and flushed with ONE global atomic
// per >=64 pairs and coalesced 16-byte stores (the reference does one global
// atomic_add per pair: gridBroadphase.cl:151, sap.cl:296).  No host round trip:
// axis choice, cell size and pair count stay on the device.
#include "internal.h"

namespace b3b200
{
constexpr int BP_THREADS = 128;
constexpr int GRID_DIM = 128;
constexpr int GRID_CELLS = GRID_DIM * GRID_DIM * GRID_DIM;
constexpr int STAGE_CAP = 96;  // per-warp staging: flush at >= 64 (larger stages were measured: no gain, the one pair counter is not the bottleneck)

// scalars layout (32-bit words)
enum
{
	SC_MAXEXT_BITS = 0,
	SC_CELL = 1,
	SC_INVCELL = 2,
	SC_AXIS = 3,
	SC_SUM = 4,   // 3 floats
	SC_SUM2 = 7,  // 3 floats
	SC_EXTSUM = 10,       // float: sum of the small AABBs' widest extents
	SC_WIDE_THRESH = 11,  // float: a small AABB wider than this is "wide" (kept out of the grid); FLT_MAX when there is none
	SC_WIDE_COUNT = 12,   // number of wide AABBs this step
	SC_COUNT = 16
};

B3_D bool aabbOverlap(const float4& mn1, const float4& mx1, const float4& mn2, const float4& mx2)
{
	// b3TestAabbAgainstAabb (shared/b3Aabb.h:45-53): inclusive, NaN => overlapping
	bool overlap = true;
	overlap = (mn1.x > mx2.x || mx1.x < mn2.x) ? false : overlap;
	overlap = (mn1.z > mx2.z || mx1.z < mn2.z) ? false : overlap;
	overlap = (mn1.y > mx2.y || mx1.y < mn2.y) ? false : overlap;
	return overlap;
}

// All 32 lanes must call this together.
B3_D void stageFlush(int2* stage, int& count, int lane, b3b200_int4* __restrict__ pairs, unsigned int* __restrict__ ctr, int maxPairs)
{
	unsigned int base = 0;
	if (lane == 0) base = atomicAdd(&ctr[CTR_PAIRS], (unsigned int)count);
	base = __shfl_sync(0xffffffffu, base, 0);
	for (int k = lane; k < count; k += 32)
	{
		unsigned int dst = base + k;
		if (dst < (unsigned int)maxPairs)
		{
			int2 p = stage[k];
			b3b200_int4 o;
			o.x = p.x;
			o.y = p.y;
			o.z = -1;  // clearOverlappingPairsKernel (updateAabbsKernel.cl:15) folded in
			o.w = -1;
			pairs[dst] = o;
		}
	}
	__syncwarp();
	count = 0;
}

B3_D void stagePush(bool hit, int a, int b, int2* stage, int& count, int lane, b3b200_int4* __restrict__ pairs, unsigned int* __restrict__ ctr, int maxPairs)
{
	unsigned int m = __ballot_sync(0xffffffffu, hit);
	if (m)
	{
		if (hit)
		{
			int pos = count + __popc(m & ((1u << lane) - 1u));
			stage[pos] = make_int2(a < b ? a : b, a < b ? b : a);
		}
		count += __popc(m);
		__syncwarp();
		if (count >= STAGE_CAP - 32) stageFlush(stage, count, lane, pairs, ctr, maxPairs);
	}
}

// ---------------------------------------------------------------- parameters
__global__ void __launch_bounds__(256) bpPrepKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int n, unsigned int* __restrict__ scal)
{
	float ext = 0.f, es = 0.f;
	float sx = 0.f, sy = 0.f, sz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const float4* p = reinterpret_cast<const float4*>(&aabbs[smallMap[i]]);
		float4 mn = __ldg(p), mx = __ldg(p + 1);
		const float e = fmaxf(mx.x - mn.x, fmaxf(mx.y - mn.y, mx.z - mn.z));
		ext = fmaxf(ext, e);
		if (e < 1e30f) es += e;
		float cx = (mx.x + mn.x) * 0.5f, cy = (mx.y + mn.y) * 0.5f, cz = (mx.z + mn.z) * 0.5f;
		sx += cx;
		sy += cy;
		sz += cz;
		qx += cx * cx;
		qy += cy * cy;
		qz += cz * cz;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, o));
		es += __shfl_xor_sync(0xffffffffu, es, o);
		sx += __shfl_xor_sync(0xffffffffu, sx, o);
		sy += __shfl_xor_sync(0xffffffffu, sy, o);
		sz += __shfl_xor_sync(0xffffffffu, sz, o);
		qx += __shfl_xor_sync(0xffffffffu, qx, o);
		qy += __shfl_xor_sync(0xffffffffu, qy, o);
		qz += __shfl_xor_sync(0xffffffffu, qz, o);
	}
	// one set of atomics per CTA (8 warps -> shared memory -> warp 0): the 7 scalars are global hot spots
	__shared__ float part[8][8];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0)
	{
		part[warp][0] = ext;
		part[warp][1] = sx;
		part[warp][2] = sy;
		part[warp][3] = sz;
		part[warp][4] = qx;
		part[warp][5] = qy;
		part[warp][6] = qz;
		part[warp][7] = es;
	}
	__syncthreads();
	if (threadIdx.x < 8)
	{
		float v = part[0][threadIdx.x];
		for (int k = 1; k < 8; k++) v = threadIdx.x == 0 ? fmaxf(v, part[k][0]) : v + part[k][threadIdx.x];
		float* f = reinterpret_cast<float*>(scal);
		if (threadIdx.x == 0)
			atomicMax(&scal[SC_MAXEXT_BITS], __float_as_uint(v));
		else if (threadIdx.x <= 3)
			atomicAdd(&f[SC_SUM + (threadIdx.x - 1)], v);
		else if (threadIdx.x <= 6)
			atomicAdd(&f[SC_SUM2 + (threadIdx.x - 4)], v);
		else
			atomicAdd(&f[SC_EXTSUM], v);
	}
}

__global__ void bpParamsKernel(unsigned int* scal, int n)
{
	float* f = reinterpret_cast<float*>(scal);
	float ext = __uint_as_float(scal[SC_MAXEXT_BITS]);
	// the cell edge must be >= the widest AABB IN THE GRID for the 27-cell scan to be exact.  Normally that is the widest small
	// AABB; but one long body among many small ones would blow every cell up (and the pair kernel towards O(n^2)), so AABBs
	// wider than 4 x the mean extent are kept out of the grid when they exist -- gridCountKernel lists them, wideSmallKernel
	// tests them against everything, like the static "large" proxies -- and the cell is sized for the rest.
	const float robust = 4.0f * f[SC_EXTSUM] / fmaxf((float)n, 1.0f);
	float thresh = FLT_MAX;
	if (n >= 64 && ext > robust && robust > 0.f)
	{
		thresh = robust;
		ext = robust;
	}
	f[SC_WIDE_THRESH] = thresh;
	float cell = fmaxf(ext * 1.01f + 1e-6f, 1e-3f);
	f[SC_CELL] = cell;
	f[SC_INVCELL] = 1.0f / cell;
	// axis of largest variance (b3GpuSapBroadphase.cpp:1041-1051)
	float nn = (float)n;
	float vx = f[SC_SUM2 + 0] - f[SC_SUM + 0] * f[SC_SUM + 0] / nn;
	float vy = f[SC_SUM2 + 1] - f[SC_SUM + 1] * f[SC_SUM + 1] / nn;
	float vz = f[SC_SUM2 + 2] - f[SC_SUM + 2] * f[SC_SUM + 2] / nn;
	int axis = 0;
	float best = vx;
	if (vy > best)
	{
		axis = 1;
		best = vy;
	}
	if (vz > best) axis = 2;
	scal[SC_AXIS] = (unsigned int)axis;
}

// ---------------------------------------------------------------- grid
B3_D int3 cellOf(const float4& mn, const float4& mx, float invCell)
{
	float cx = (mx.x + mn.x) * 0.5f, cy = (mx.y + mn.y) * 0.5f, cz = (mx.z + mn.z) * 0.5f;
	// (clamped before the conversion: a far-away or non-finite centre must not overflow the int, and c.x + 1 must not either)
	const float lim = 1.0e9f;
	return make_int3((int)fminf(fmaxf(floorf(cx * invCell), -lim), lim), (int)fminf(fmaxf(floorf(cy * invCell), -lim), lim), (int)fminf(fmaxf(floorf(cz * invCell), -lim), lim));
}
// Batched independent worlds ("key = (worldId << k) + cell"): every world gets its own 8 x 8 x 8 block of the 128^3
// cell table (4096 blocks, then they repeat), i.e. its cell coordinates are shifted before the wrap.  Worlds that stand at the
// same coordinates therefore do not pile into the same cells, neighbouring cells stay neighbours, and exactness never depends
// on the shift: the pair test compares the world ids.
B3_D int3 worldCellShift(int world)
{
	return make_int3((world & 15) << 3, ((world >> 4) & 15) << 3, ((world >> 8) & 15) << 3);
}
B3_D unsigned int cellKey(int x, int y, int z)
{
	return ((unsigned int)(z & (GRID_DIM - 1)) << 14) | ((unsigned int)(y & (GRID_DIM - 1)) << 7) | (unsigned int)(x & (GRID_DIM - 1));
}

// Counting sort by cell (the keys ARE small integers: no radix passes): count with the rank inside the cell as a by-product,
// dense exclusive scan over the 128^3 cells, scatter.  The dense table gives the sorted range of ANY run of cells -- also of
// empty ones -- with two loads, which is what the pair kernel needs.
__global__ void __launch_bounds__(256) gridCountKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int n,
													   const unsigned int* scal, unsigned int* __restrict__ keys, unsigned int* __restrict__ rankInCell,
													   unsigned int* __restrict__ cellCount, const int* __restrict__ worldOf, unsigned int* wideCount,
													   int* __restrict__ wideList)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float invCell = __uint_as_float(scal[SC_INVCELL]);
	int idx = smallMap[i];
	const float4* p = reinterpret_cast<const float4*>(&aabbs[idx]);
	{
		const float4 mn = __ldg(p), mx = __ldg(p + 1);
		if (fmaxf(mx.x - mn.x, fmaxf(mx.y - mn.y, mx.z - mn.z)) > __uint_as_float(scal[SC_WIDE_THRESH]))
		{
			// a wide AABB: not in the grid (its key marks it for the scatter kernel), listed for wideSmallKernel
			wideList[atomicAdd(wideCount, 1u)] = idx;
			keys[i] = 0xffffffffu;
			return;
		}
	}
	int3 c = cellOf(__ldg(p), __ldg(p + 1), invCell);
	if (worldOf)
	{
		const int3 sh = worldCellShift(__ldg(&worldOf[idx]));
		c.x += sh.x;
		c.y += sh.y;
		c.z += sh.z;
	}
	const unsigned int key = cellKey(c.x, c.y, c.z);
	keys[i] = key;
	rankInCell[i] = atomicAdd(&cellCount[key], 1u);
}

// AABBs into cell order (the order inside a cell is the order of the atomics: the pair SET does not depend on it)
__global__ void __launch_bounds__(256) gridScatterKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int n,
														 const unsigned int* __restrict__ keys, const unsigned int* __restrict__ rankInCell,
														 const unsigned int* __restrict__ cellStart, b3b200_aabb* __restrict__ sorted, const int* __restrict__ worldOf)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (keys[i] == 0xffffffffu) return;  // a wide AABB (gridCountKernel)
	const float4* p = reinterpret_cast<const float4*>(&aabbs[smallMap[i]]);
	const float4 mn = __ldg(p);
	float4 mx = __ldg(p + 1);
	if (worldOf) mx.w = __int_as_float(__ldg(&worldOf[smallMap[i]]));  // the sorted copy carries the world id for the pair kernel
	float4* q = reinterpret_cast<float4*>(&sorted[cellStart[keys[i]] + rankInCell[i]]);
	q[0] = mn;
	q[1] = mx;
}

// gather AABBs into sorted order (coalesced 2x128-bit per proxy) -- SAP path
__global__ void __launch_bounds__(256) gatherKernel(const b3b200_aabb* __restrict__ aabbs, const unsigned int* __restrict__ keys, const unsigned int* __restrict__ vals,
													int n, b3b200_aabb* __restrict__ sorted, const int* __restrict__ worldOf)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4* p = reinterpret_cast<const float4*>(&aabbs[vals[i]]);
	float4 mn = __ldg(p), mx = __ldg(p + 1);
	if (worldOf) mx.w = __int_as_float(__ldg(&worldOf[vals[i]]));
	float4* q = reinterpret_cast<float4*>(&sorted[i]);
	q[0] = mn;
	q[1] = mx;
}

// One thread per AABB (in cell order).  Its candidates are the 3 x 3 x 3 cells around its own: per (y, z) row the cells x-1..x+1
// are ONE contiguous range of the sorted array, [cellStart[key(x-1)], cellStart[key(x+1) + 1]) -- two loads per row, all rows'
// loads independent of each other (the first version walked 27 cells with a dependent cellStart -> key -> AABB chain each).
// A pair is emitted by the body with the lower sorted index; rows whose keys are all smaller than the body's own cannot
// hold a higher index, so away from the wrap-around of the 128^3 hash only 5 of the 9 rows are visited.  Consecutive threads
// are neighbours in space: their candidate ranges overlap and are served by L1.
constexpr int GRID_ROWS = 27;
constexpr int GRID_UNROLL = 4;
__global__ void __launch_bounds__(BP_THREADS) gridFindPairsKernel(const b3b200_aabb* __restrict__ sorted, const unsigned int* __restrict__ cellStart, int n,
																  const unsigned int* __restrict__ scal, b3b200_int4* __restrict__ pairs, unsigned int* __restrict__ ctr, int maxPairs,
																  int hasWorlds)
{
	__shared__ int2 stageAll[BP_THREADS / 32][STAGE_CAP];
	__shared__ int2 sRange[GRID_ROWS][BP_THREADS];  // [first, end) per row of cells (9 rows, or 27 single cells next to the wrap-around)
	const int lane = threadIdx.x & 31, t = threadIdx.x;
	int2* stage = stageAll[threadIdx.x >> 5];
	int count = 0;
	const float invCell = __uint_as_float(scal[SC_INVCELL]);
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	n = (int)__ldg(&cellStart[GRID_CELLS]);  // bodies in the grid (all small AABBs but the wide ones)
	bool valid = i < n;
	float4 mnA = mk4(0, 0, 0), mxA = mk4(0, 0, 0);
	int idA = 0;
	int nr = 0;
	if (valid)
	{
		const float4* p = reinterpret_cast<const float4*>(&sorted[i]);
		mnA = p[0];
		mxA = p[1];
		idA = __float_as_int(mnA.w);
		int3 c = cellOf(mnA, mxA, invCell);
		if (hasWorlds)
		{
			const int3 sh = worldCellShift(__float_as_int(mxA.w));
			c.x += sh.x;
			c.y += sh.y;
			c.z += sh.z;
		}
		const int xw = c.x & (GRID_DIM - 1), yw = c.y & (GRID_DIM - 1), zw = c.z & (GRID_DIM - 1);
		const bool inner = yw >= 1 && yw <= GRID_DIM - 2 && zw >= 1 && zw <= GRID_DIM - 2;  // no (y, z) wrap: key order == (z, y, x) order
		const bool xInner = xw >= 1 && xw <= GRID_DIM - 2;
		auto add = [&](int lo, int hi) {
			if (inner && hi <= i + 1) return;   // nothing behind this body in the range
			if (inner && lo <= i) lo = i + 1;   // (own row: start right behind the body itself)
			if (hi > lo) sRange[nr++][t] = make_int2(lo, hi);
		};
#pragma unroll
		for (int r = 0; r < 9; r++)
		{
			const int dz = r / 3 - 1, dy = r % 3 - 1;
			if (inner && (dz < 0 || (dz == 0 && dy < 0))) continue;  // all of that row sorts before this body
			if (xInner)
			{
				const unsigned int k0 = cellKey(c.x - 1, c.y + dy, c.z + dz);
				add((int)__ldg(&cellStart[k0]), (int)__ldg(&cellStart[k0 + 3]));
			}
			else
			{
				for (int dx = -1; dx <= 1; dx++)
				{
					const unsigned int k = cellKey(c.x + dx, c.y + dy, c.z + dz);
					add((int)__ldg(&cellStart[k]), (int)__ldg(&cellStart[k + 1]));
				}
			}
		}
	}
	int r = -1, j = 0, jend = 0;
	if (nr == 0) valid = false;
	for (;;)
	{
		// GRID_UNROLL candidates per lane and round: their loads are issued together (one dependent load per candidate was the
		// kernel's critical path)
		int jj[GRID_UNROLL];
#pragma unroll
		for (int k = 0; k < GRID_UNROLL; k++) jj[k] = -1;
		if (valid)
		{
#pragma unroll
			for (int k = 0; k < GRID_UNROLL; k++)
			{
				while (valid && j >= jend)
				{
					if (++r < nr)
					{
						const int2 rg = sRange[r][t];
						j = rg.x;
						jend = rg.y;
					}
					else
						valid = false;
				}
				if (valid) jj[k] = j++;
			}
		}
		float4 mnB[GRID_UNROLL], mxB[GRID_UNROLL];
#pragma unroll
		for (int k = 0; k < GRID_UNROLL; k++)
		{
			mnB[k] = mk4(0, 0, 0);
			mxB[k] = mnB[k];
			if (jj[k] > i)
			{
				const float4* p = reinterpret_cast<const float4*>(&sorted[jj[k]]);
				mnB[k] = __ldg(p);
				mxB[k] = __ldg(p + 1);
			}
		}
#pragma unroll
		for (int k = 0; k < GRID_UNROLL; k++)
		{
			const bool hit = jj[k] > i && aabbOverlap(mnA, mxA, mnB[k], mxB[k]) && (!hasWorlds || __float_as_int(mxB[k].w) == __float_as_int(mxA.w));
			stagePush(hit, idA, __float_as_int(mnB[k].w), stage, count, lane, pairs, ctr, maxPairs);
		}
		if (!__any_sync(0xffffffffu, valid)) break;
	}
	if (count) stageFlush(stage, count, lane, pairs, ctr, maxPairs);
}

// ---------------------------------------------------------------- SAP
B3_D unsigned int floatFlip(float f)
{
	// flipFloatKernel (sap.cl:353): IEEE-754 -> order-preserving u32
	unsigned int u = __float_as_uint(f);
	unsigned int mask = (unsigned int)(-(int)(u >> 31)) | 0x80000000u;
	return u ^ mask;
}
B3_D float axisOf(const float4& v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

__global__ void __launch_bounds__(256) sapKeyKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int n,
													const unsigned int* __restrict__ scal, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int axis = (int)scal[SC_AXIS];
	int idx = smallMap[i];
	float4 mn = __ldg(reinterpret_cast<const float4*>(&aabbs[idx]));
	keys[i] = floatFlip(axisOf(mn, axis));
	vals[i] = (unsigned int)idx;
}

// Warp-cooperative sweep with a shared-memory window (the role of the LDS window of computePairsKernelLocalSharedMemory,
// sap.cl:231-300): a warp owns 32 consecutive bodies of the sorted order; the candidates behind them are loaded ONCE per warp,
// 32 at a time and coalesced, into shared memory, and every lane tests the 32 staged candidates against its own body.  A lane
// stops for good at the first candidate whose minimum lies beyond its maximum on the sweep axis (the reference's break test);
// the warp stops when all its lanes have.  Hits are kept as one bit per (lane, candidate); after each window the warp
// reserves room for all of them with one atomic and the lanes write their pairs.  Same pair set as one thread per body.
__global__ void __launch_bounds__(BP_THREADS) sapSweepKernel(const b3b200_aabb* __restrict__ sorted, int n, const unsigned int* __restrict__ scal,
															 b3b200_int4* __restrict__ pairs, unsigned int* __restrict__ ctr, int maxPairs, int hasWorlds)
{
	__shared__ float4 winAll[BP_THREADS / 32][2][32];
	const int lane = threadIdx.x & 31;
	float4* wmn = winAll[threadIdx.x >> 5][0];
	float4* wmx = winAll[threadIdx.x >> 5][1];
	const int axis = (int)scal[SC_AXIS];
	const int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
	if (base >= n) return;
	const int i = base + lane;
	bool active = i < n;
	float4 mnA = mk4(0, 0, 0), mxA = mk4(0, 0, 0);
	float limit = 0.f;
	if (active)
	{
		const float4* p = reinterpret_cast<const float4*>(&sorted[i]);
		mnA = p[0];
		mxA = p[1];
		limit = axisOf(mxA, axis);
	}
	const int idA = __float_as_int(mnA.w);
	for (int start = base + 1; start < n; start += 32)
	{
		const int j = start + lane;
		if (j < n)
		{
			const float4* p = reinterpret_cast<const float4*>(&sorted[j]);
			wmn[lane] = __ldg(p);
			wmx[lane] = __ldg(p + 1);
		}
		__syncwarp();
		const int cnt = n - start < 32 ? n - start : 32;
		unsigned int hits = 0u;
		if (active)
		{
			// candidates [start, start + cnt); this lane's own successors start at i + 1
			int t = i + 1 - start;
			if (t < 0) t = 0;
			for (; t < cnt; t++)
			{
				const float4 mnB = wmn[t];
				// computePairsKernelLocalSharedMemory break test (sap.cl:231-300)
				if (limit < axisOf(mnB, axis))
				{
					active = false;
					break;
				}
				const float4 mxB = wmx[t];
				if (aabbOverlap(mnA, mxA, mnB, mxB) && (!hasWorlds || __float_as_int(mxB.w) == __float_as_int(mxA.w))) hits |= 1u << t;
			}
		}
		// emit: one reservation per warp and window
		const int mine = __popc(hits);
		int incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int v = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += v;
		}
		const int total = __shfl_sync(0xffffffffu, incl, 31);
		if (total)
		{
			unsigned int slot = 0;
			if (lane == 0) slot = atomicAdd(&ctr[CTR_PAIRS], (unsigned int)total);
			slot = __shfl_sync(0xffffffffu, slot, 0) + (unsigned int)(incl - mine);
			while (hits)
			{
				const int t = __ffs(hits) - 1;
				hits &= hits - 1;
				if (slot < (unsigned int)maxPairs)
				{
					const int idB = __float_as_int(wmn[t].w);
					b3b200_int4 o;
					o.x = idA < idB ? idA : idB;
					o.y = idA < idB ? idB : idA;
					o.z = -1;  // clearOverlappingPairsKernel (updateAabbsKernel.cl:15) folded in
					o.w = -1;
					pairs[slot] = o;
				}
				slot++;
			}
		}
		if (!__any_sync(0xffffffffu, active)) break;
		__syncwarp();
	}
}

// ---------------------------------------------------------------- large x small
__global__ void __launch_bounds__(BP_THREADS) largeSmallKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int nSmall,
															   const int* __restrict__ largeMap, int nLarge,
															   b3b200_int4* __restrict__ pairs, unsigned int* __restrict__ ctr, int maxPairs,
															   const int* __restrict__ worldOf, const int* __restrict__ largeStart)
{
	__shared__ int2 stageAll[BP_THREADS / 32][STAGE_CAP];
	const int lane = threadIdx.x & 31;
	int2* stage = stageAll[threadIdx.x >> 5];
	int count = 0;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = i < nSmall;
	float4 mnA = mk4(0, 0, 0), mxA = mk4(0, 0, 0);
	int idA = 0;
	if (valid)
	{
		const float4* p = reinterpret_cast<const float4*>(&aabbs[smallMap[i]]);
		mnA = __ldg(p);
		mxA = __ldg(p + 1);
		idA = __float_as_int(mnA.w);
	}
	// batched worlds: largeMap is grouped by world and a body only meets the large proxies of its own world
	int first = 0, num = nLarge;
	if (worldOf)
	{
		first = num = 0;
		if (valid)
		{
			const int wd = __ldg(&worldOf[smallMap[i]]);
			first = __ldg(&largeStart[wd]);
			num = __ldg(&largeStart[wd + 1]) - first;
		}
	}
	const int numMax = worldOf ? __reduce_max_sync(0xffffffffu, num) : nLarge;
	for (int l = 0; l < numMax; l++)
	{
		bool hit = false;
		int idB = 0;
		if (l < num)
		{
			const float4* p = reinterpret_cast<const float4*>(&aabbs[largeMap[first + l]]);
			float4 mnB = __ldg(p), mxB = __ldg(p + 1);
			hit = valid && aabbOverlap(mnA, mxA, mnB, mxB);
			idB = __float_as_int(mnB.w);
		}
		stagePush(hit, idA, idB, stage, count, lane, pairs, ctr, maxPairs);
	}
	if (count) stageFlush(stage, count, lane, pairs, ctr, maxPairs);
}

// wide x everything: the AABBs gridCountKernel kept out of the grid against every small AABB (a pair of two wide ones is
// emitted by the one with the lower index).  Usually there is none and every thread leaves after one load.
__global__ void __launch_bounds__(BP_THREADS) wideSmallKernel(const b3b200_aabb* __restrict__ aabbs, const int* __restrict__ smallMap, int nSmall,
															  const int* __restrict__ wideList, const unsigned int* __restrict__ scal, b3b200_int4* __restrict__ pairs,
															  unsigned int* __restrict__ ctr, int maxPairs, const int* __restrict__ worldOf)
{
	const int nWide = (int)scal[SC_WIDE_COUNT];
	if (nWide == 0) return;
	__shared__ int2 stageAll[BP_THREADS / 32][STAGE_CAP];
	const int lane = threadIdx.x & 31;
	int2* stage = stageAll[threadIdx.x >> 5];
	int count = 0;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = i < nSmall;
	float4 mnA = mk4(0, 0, 0), mxA = mk4(0, 0, 0);
	int idxA = -1, wA = 0;
	bool wideA = false;
	if (valid)
	{
		idxA = smallMap[i];
		const float4* p = reinterpret_cast<const float4*>(&aabbs[idxA]);
		mnA = __ldg(p);
		mxA = __ldg(p + 1);
		wideA = fmaxf(mxA.x - mnA.x, fmaxf(mxA.y - mnA.y, mxA.z - mnA.z)) > __uint_as_float(scal[SC_WIDE_THRESH]);
		if (worldOf) wA = __ldg(&worldOf[idxA]);
	}
	const int idA = __float_as_int(mnA.w);
	for (int l = 0; l < nWide; l++)
	{
		const int idxB = __ldg(&wideList[l]);
		const float4* p = reinterpret_cast<const float4*>(&aabbs[idxB]);
		const float4 mnB = __ldg(p), mxB = __ldg(p + 1);
		bool hit = valid && idxB != idxA && !(wideA && idxB < idxA) && aabbOverlap(mnA, mxA, mnB, mxB);
		if (hit && worldOf) hit = __ldg(&worldOf[idxB]) == wA;
		stagePush(hit, idA, __float_as_int(mnB.w), stage, count, lane, pairs, ctr, maxPairs);
	}
	if (count) stageFlush(stage, count, lane, pairs, ctr, maxPairs);
}

__global__ void clampPairsKernel(unsigned int* ctr, int maxPairs)
{
	if (ctr[CTR_PAIRS] > (unsigned int)maxPairs)
	{
		ctr[CTR_PAIRS] = (unsigned int)maxPairs;
		ctr[CTR_OVERFLOW] |= OVF_PAIRS;
	}
}

// ---------------------------------------------------------------- host side
int Broadphase::init(int kind_, int device_, cudaStream_t stream_, int maxProxies_, int maxPairs_)
{
	kind = kind_;
	device = device_;
	maxProxies = maxProxies_;
	maxPairs = maxPairs_;
	B3_CUDA_CHECK(cudaSetDevice(device));
	if (stream_)
	{
		stream = stream_;
		ownStream = false;
	}
	else
	{
		B3_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		ownStream = true;
	}
	B3_TRY(counters.reserve(CTR_COUNT));
	B3_CUDA_CHECK(cudaMemsetAsync(counters.ptr, 0, sizeof(unsigned int) * CTR_COUNT, stream));
	ctr = counters.ptr;
	B3_TRY(scalars.reserve(SC_COUNT));
	B3_TRY(pairs.reserve(maxPairs > 0 ? maxPairs : 1));
	B3_CUDA_CHECK(cudaEventCreate(&ev0));
	B3_CUDA_CHECK(cudaEventCreate(&ev1));
	return 0;
}

void Broadphase::destroy()
{
	cudaSetDevice(device);
	if (stream) cudaStreamSynchronize(stream);
	if (ev0) cudaEventDestroy(ev0);
	if (ev1) cudaEventDestroy(ev1);
	ev0 = ev1 = nullptr;
	if (ownStream && stream) cudaStreamDestroy(stream);
	stream = 0;
	aabbs.release();
	smallMap.release();
	largeMap.release();
	pairs.release();
	counters.release();
	keys.release();
	vals.release();
	sortedAabbs.release();
	cellStart.release();
	cellCnt.release();
	scanTotals.release();
	scalars.release();
}

int Broadphase::reset()
{
	aabbsCPU.clear();
	smallIdx.clear();
	largeIdx.clear();
	numSmall = numLarge = numAabbs = 0;
	return 0;
}

int Broadphase::createProxy(const float* mn, const float* mx, int userPtr, bool large)
{
	if ((int)aabbsCPU.size() >= maxProxies)
	{
		setLastError("broadphase: exceeding the number of proxies (%d)", maxProxies);
		return B3B200_ERR_CAPACITY;
	}
	// b3GpuSapBroadphase::createProxy / createLargeProxy (b3GpuSapBroadphase.cpp:1233-1264)
	b3b200_aabb a;
	a.min[0] = mn[0];
	a.min[1] = mn[1];
	a.min[2] = mn[2];
	a.minIndices[3] = userPtr;
	a.max[0] = mx[0];
	a.max[1] = mx[1];
	a.max[2] = mx[2];
	a.signedMaxIndices[3] = (int)aabbsCPU.size();
	(large ? largeIdx : smallIdx).push_back((int)aabbsCPU.size());
	aabbsCPU.push_back(a);
	return 0;
}

int Broadphase::writeAabbs()
{
	B3_CUDA_CHECK(cudaSetDevice(device));
	numAabbs = (int)aabbsCPU.size();
	numSmall = (int)smallIdx.size();
	numLarge = (int)largeIdx.size();
	B3_TRY(aabbs.reserve(numAabbs > 0 ? numAabbs : 1));
	B3_TRY(smallMap.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(largeMap.reserve(numLarge > 0 ? numLarge : 1));
	B3_TRY(keys.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(vals.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(sortedAabbs.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(wideList.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(cellStart.reserve(GRID_CELLS + 4));
	B3_TRY(cellCnt.reserve(GRID_CELLS + 4));
	B3_TRY(scanTotals.reserve((size_t)largeScanChunks(GRID_CELLS + 4)));
	B3_TRY(sortTmp.keysAlt.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(sortTmp.valsAlt.reserve(numSmall > 0 ? numSmall : 1));
	B3_TRY(sortTmp.blockHist.reserve((size_t)256 * divUp(numSmall > 0 ? numSmall : 1, 2048)));
	if (numAabbs) B3_CUDA_CHECK(cudaMemcpyAsync(aabbs.ptr, aabbsCPU.data(), sizeof(b3b200_aabb) * numAabbs, cudaMemcpyHostToDevice, stream));
	if (numSmall) B3_CUDA_CHECK(cudaMemcpyAsync(smallMap.ptr, smallIdx.data(), sizeof(int) * numSmall, cudaMemcpyHostToDevice, stream));
	if (numLarge) B3_CUDA_CHECK(cudaMemcpyAsync(largeMap.ptr, largeIdx.data(), sizeof(int) * numLarge, cudaMemcpyHostToDevice, stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(stream));
	return 0;
}

int Broadphase::calculatePairs(int maxPairsNow)
{
	if (maxPairsNow > maxPairs) maxPairsNow = maxPairs;
	if (maxPairsNow < 0) maxPairsNow = 0;
	cudaStream_t s = stream;
	B3_CUDA_CHECK(cudaMemsetAsync(&ctr[CTR_PAIRS], 0, sizeof(unsigned int), s));
	if (numSmall > 0)
	{
		B3_CUDA_CHECK(cudaMemsetAsync(scalars.ptr, 0, sizeof(float) * SC_COUNT, s));
		int prepBlocks = divUp(numSmall, 256);
		if (prepBlocks > 592) prepBlocks = 592;
		bpPrepKernel<<<prepBlocks, 256, 0, s>>>(aabbs.ptr, smallMap.ptr, numSmall, reinterpret_cast<unsigned int*>(scalars.ptr));
		B3_LAUNCH_CHECK();
		bpParamsKernel<<<1, 1, 0, s>>>(reinterpret_cast<unsigned int*>(scalars.ptr), numSmall);
		B3_LAUNCH_CHECK();
		const unsigned int* scal = reinterpret_cast<const unsigned int*>(scalars.ptr);
		if (kind == B3B200_BP_GRID)
		{
			unsigned int* cellCount = reinterpret_cast<unsigned int*>(cellCnt.ptr);
			unsigned int* cellBegin = reinterpret_cast<unsigned int*>(cellStart.ptr);
			B3_CUDA_CHECK(cudaMemsetAsync(cellCount, 0, sizeof(unsigned int) * (GRID_CELLS + 4), s));
			gridCountKernel<<<divUp(numSmall, 256), 256, 0, s>>>(aabbs.ptr, smallMap.ptr, numSmall, scal, keys.ptr, vals.ptr, cellCount, worldOf,
																 reinterpret_cast<unsigned int*>(scalars.ptr) + SC_WIDE_COUNT, wideList.ptr);
			B3_LAUNCH_CHECK();
			// GRID_CELLS + 4 entries: cellStart[key + 3] of the last cells reads the total
			B3_TRY(exclusiveScanLargeU32(s, cellCount, cellBegin, GRID_CELLS + 4, scanTotals.ptr, nullptr));
			gridScatterKernel<<<divUp(numSmall, 256), 256, 0, s>>>(aabbs.ptr, smallMap.ptr, numSmall, keys.ptr, vals.ptr, cellBegin, sortedAabbs.ptr, worldOf);
			B3_LAUNCH_CHECK();
			gridFindPairsKernel<<<divUp(numSmall, BP_THREADS), BP_THREADS, 0, s>>>(sortedAabbs.ptr, cellBegin, numSmall, scal, pairs.ptr, ctr, maxPairsNow, worldOf ? 1 : 0);
			B3_LAUNCH_CHECK();
			wideSmallKernel<<<divUp(numSmall, BP_THREADS), BP_THREADS, 0, s>>>(aabbs.ptr, smallMap.ptr, numSmall, wideList.ptr, scal, pairs.ptr, ctr, maxPairsNow, worldOf);
			B3_LAUNCH_CHECK();
		}
		else
		{
			sapKeyKernel<<<divUp(numSmall, 256), 256, 0, s>>>(aabbs.ptr, smallMap.ptr, numSmall, scal, keys.ptr, vals.ptr);
			B3_LAUNCH_CHECK();
			B3_TRY(radixSortKV32(s, sortTmp, keys.ptr, vals.ptr, numSmall, 32));
			gatherKernel<<<divUp(numSmall, 256), 256, 0, s>>>(aabbs.ptr, keys.ptr, vals.ptr, numSmall, sortedAabbs.ptr, worldOf);
			B3_LAUNCH_CHECK();
			sapSweepKernel<<<divUp(numSmall, BP_THREADS), BP_THREADS, 0, s>>>(sortedAabbs.ptr, numSmall, scal, pairs.ptr, ctr, maxPairsNow, worldOf ? 1 : 0);
			B3_LAUNCH_CHECK();
		}
		if (numLarge > 0)
		{
			largeSmallKernel<<<divUp(numSmall, BP_THREADS), BP_THREADS, 0, s>>>(aabbs.ptr, smallMap.ptr, numSmall, largeMap.ptr, numLarge, pairs.ptr, ctr, maxPairsNow, worldOf,
																				worldOf ? largeStart.ptr : nullptr);
			B3_LAUNCH_CHECK();
		}
	}
	clampPairsKernel<<<1, 1, 0, s>>>(ctr, maxPairsNow);
	B3_LAUNCH_CHECK();
	return 0;
}

}  // namespace b3b200

using namespace b3b200;

// ------------------------------------------------------------------ C ABI
extern "C" int b3b200_bp_create(int kind, int device, void* stream, int maxProxies, int maxPairs, b3b200_broadphase** out)
{
	if (!out || maxProxies <= 0 || maxPairs < 0 || (kind != B3B200_BP_SAP && kind != B3B200_BP_GRID)) return B3B200_ERR_INVALID;
	b3b200_broadphase* bp = new b3b200_broadphase();
	int r = bp->init(kind, device, (cudaStream_t)stream, maxProxies, maxPairs);
	if (r < 0)
	{
		bp->destroy();
		delete bp;
		return r;
	}
	*out = bp;
	return 0;
}
extern "C" int b3b200_bp_destroy(b3b200_broadphase* bp)
{
	if (!bp) return B3B200_ERR_INVALID;
	bp->destroy();
	delete bp;
	return 0;
}
extern "C" int b3b200_bp_create_proxy(b3b200_broadphase* bp, const float* mn, const float* mx, int userPtr)
{
	if (!bp || !mn || !mx) return B3B200_ERR_INVALID;
	return bp->createProxy(mn, mx, userPtr, false);
}
extern "C" int b3b200_bp_create_large_proxy(b3b200_broadphase* bp, const float* mn, const float* mx, int userPtr)
{
	if (!bp || !mn || !mx) return B3B200_ERR_INVALID;
	return bp->createProxy(mn, mx, userPtr, true);
}
extern "C" int b3b200_bp_write_aabbs(b3b200_broadphase* bp)
{
	if (!bp) return B3B200_ERR_INVALID;
	return bp->writeAabbs();
}
extern "C" int b3b200_bp_set_aabbs(b3b200_broadphase* bp, const b3b200_aabb* a, int n)
{
	if (!bp || !a || n != (int)bp->aabbsCPU.size()) return B3B200_ERR_INVALID;
	for (int i = 0; i < n; i++) bp->aabbsCPU[i] = a[i];
	if (n != bp->numAabbs) return bp->writeAabbs();
	B3_CUDA_CHECK(cudaSetDevice(bp->device));
	B3_CUDA_CHECK(cudaMemcpyAsync(bp->aabbs.ptr, a, sizeof(b3b200_aabb) * n, cudaMemcpyHostToDevice, bp->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(bp->stream));
	return 0;
}
extern "C" int b3b200_bp_calculate_pairs(b3b200_broadphase* bp, int maxPairs)
{
	if (!bp) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(bp->device));
	if (bp->numAabbs != (int)bp->aabbsCPU.size()) B3_TRY(bp->writeAabbs());
	B3_CUDA_CHECK(cudaEventRecord(bp->ev0, bp->stream));
	B3_TRY(bp->calculatePairs(maxPairs));
	B3_CUDA_CHECK(cudaEventRecord(bp->ev1, bp->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(bp->stream));
	cudaEventElapsedTime(&bp->lastMs, bp->ev0, bp->ev1);
	return 0;
}
extern "C" int b3b200_bp_num_overlap(b3b200_broadphase* bp)
{
	if (!bp) return B3B200_ERR_INVALID;
	unsigned int n = 0;
	B3_CUDA_CHECK(cudaSetDevice(bp->device));
	B3_CUDA_CHECK(cudaMemcpyAsync(&n, &bp->ctr[CTR_PAIRS], sizeof(n), cudaMemcpyDeviceToHost, bp->stream));
	B3_CUDA_CHECK(cudaStreamSynchronize(bp->stream));
	return (int)n;
}
extern "C" int b3b200_bp_get_pairs(b3b200_broadphase* bp, b3b200_int4* dst, int capacity, int* numPairs)
{
	if (!bp || !numPairs || capacity < 0) return B3B200_ERR_INVALID;
	int n = b3b200_bp_num_overlap(bp);
	if (n < 0) return n;
	*numPairs = n;
	int m = n < capacity ? n : capacity;
	if (m > 0 && dst)
	{
		B3_CUDA_CHECK(cudaMemcpyAsync(dst, bp->pairs.ptr, sizeof(b3b200_int4) * m, cudaMemcpyDeviceToHost, bp->stream));
		B3_CUDA_CHECK(cudaStreamSynchronize(bp->stream));
	}
	return 0;
}
extern "C" int b3b200_bp_device_pairs(b3b200_broadphase* bp, void** p)
{
	if (!bp || !p) return B3B200_ERR_INVALID;
	*p = bp->pairs.ptr;
	return 0;
}
extern "C" int b3b200_bp_device_aabbs(b3b200_broadphase* bp, void** p)
{
	if (!bp || !p) return B3B200_ERR_INVALID;
	*p = bp->aabbs.ptr;
	return 0;
}
extern "C" int b3b200_bp_last_ms(b3b200_broadphase* bp, float* ms)
{
	if (!bp || !ms) return B3B200_ERR_INVALID;
	*ms = bp->lastMs;
	return 0;
}
