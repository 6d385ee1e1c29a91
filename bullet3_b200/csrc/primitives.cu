// primitives.cu -- radix sort / scan / bound search / fill for sm_100a.
//
// Radix sort: LSD, 8-bit digits.  Per pass: (1) per-tile digit histogram,
// (2) one exclusive scan over the digit-major [256][numTiles] table,
// (3) stable scatter.  Ranking inside a tile is done with __match_any_sync:
// every warp owns 8 rounds of 32 consecutive keys, peers with the same digit
// rank themselves with a popc of the lower-lane mask -- no shared-memory
// atomics, no bank-conflicted counters, and the order (warp, round, lane) is
// exactly input order, so the sort is stable like b3RadixSort32CL
// (b3RadixSort32CL.cpp:12-646; its executeHost twin :587-646 is the oracle).
#include "primitives.cuh"

namespace b3b200
{
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rsHistKernel(const KeyT* __restrict__ keys, int n, int shift,
														   unsigned int* __restrict__ blockHist, int numBlocks)
{
	__shared__ unsigned int hist[256];
	hist[threadIdx.x] = 0;
	__syncthreads();
	int base = blockIdx.x * RS_TILE;
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++)
	{
		int i = base + r * RS_THREADS + threadIdx.x;
		if (i < n)
		{
			unsigned int d = (unsigned int)(keys[i] >> shift) & 255u;
			atomicAdd(&hist[d], 1u);
		}
	}
	__syncthreads();
	blockHist[threadIdx.x * numBlocks + blockIdx.x] = hist[threadIdx.x];
}

template <typename KeyT, bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS) rsScatterKernel(const KeyT* __restrict__ keysIn, const unsigned int* __restrict__ valsIn,
															  KeyT* __restrict__ keysOut, unsigned int* __restrict__ valsOut,
															  int n, int shift, const unsigned int* __restrict__ blockOffsets, int numBlocks,
															  const unsigned int* __restrict__ digitTotals)
{
	__shared__ unsigned int whist[RS_WARPS][257];
	__shared__ unsigned int digitBase[256];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_THREADS) (&whist[0][0])[i] = 0;
	{
		// exclusive scan of the 256 digit totals (every CTA its own copy: 256 values)
		const unsigned int v = digitTotals[threadIdx.x];
		unsigned int incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		digitBase[threadIdx.x] = incl - v;  // exclusive inside the warp
		__syncthreads();
		unsigned int before = 0;
		for (int w = 0; w < warp; w++) before += digitBase[w * 32 + 31] + digitTotals[w * 32 + 31];
		__syncthreads();
		digitBase[threadIdx.x] = before + incl - v;
	}
	__syncthreads();

	const int warpBase = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
	KeyT key[RS_ITEMS];
	unsigned int val[RS_ITEMS];
	unsigned int rank[RS_ITEMS];
	unsigned int dig[RS_ITEMS];
	const unsigned int ltMask = (1u << lane) - 1u;
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++)
	{
		int i = warpBase + r * 32 + lane;
		bool valid = i < n;
		key[r] = valid ? keysIn[i] : (KeyT)0;
		if (HAS_VALS) val[r] = valid ? valsIn[i] : 0u;
		dig[r] = valid ? ((unsigned int)(key[r] >> shift) & 255u) : 256u;
	}
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++)
	{
		unsigned int d = dig[r];
		unsigned int peers = __match_any_sync(0xffffffffu, d);
		unsigned int before = whist[warp][d];
		__syncwarp();
		if ((peers & ltMask) == 0) whist[warp][d] = before + __popc(peers);
		__syncwarp();
		rank[r] = before + __popc(peers & ltMask);
	}
	__syncthreads();
	{
		// thread d turns the per-warp counts of digit d into global write cursors
		int d = threadIdx.x;
		unsigned int run = digitBase[d] + blockOffsets[d * numBlocks + blockIdx.x];
#pragma unroll
		for (int w = 0; w < RS_WARPS; w++)
		{
			unsigned int t = whist[w][d];
			whist[w][d] = run;
			run += t;
		}
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++)
	{
		if (dig[r] < 256u)
		{
			unsigned int pos = whist[warp][dig[r]] + rank[r];
			keysOut[pos] = key[r];
			if (HAS_VALS) valsOut[pos] = val[r];
		}
	}
}

// single-CTA exclusive scan (tables here are small: 256 x numTiles)
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) scanKernel(const unsigned int* src, unsigned int* dst, int n, unsigned int* total)
{
	__shared__ unsigned int warpSums[32];
	__shared__ unsigned int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int base = 0; base < n; base += SCAN_THREADS * 4)
	{
		int i = base + threadIdx.x * 4;
		unsigned int v0 = i < n ? src[i] : 0, v1 = i + 1 < n ? src[i + 1] : 0, v2 = i + 2 < n ? src[i + 2] : 0, v3 = i + 3 < n ? src[i + 3] : 0;
		unsigned int sum = v0 + v1 + v2 + v3;
		unsigned int incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) warpSums[warp] = incl;
		__syncthreads();
		if (warp == 0)
		{
			unsigned int w = warpSums[lane];
			unsigned int wi = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= o) wi += t;
			}
			warpSums[lane] = wi - w;  // exclusive
		}
		__syncthreads();
		unsigned int excl = carry + warpSums[warp] + (incl - sum);
		if (i < n) dst[i] = excl;
		if (i + 1 < n) dst[i + 1] = excl + v0;
		if (i + 2 < n) dst[i + 2] = excl + v0 + v1;
		if (i + 3 < n) dst[i + 3] = excl + v0 + v1 + v2;
		__syncthreads();
		if (threadIdx.x == SCAN_THREADS - 1) carry = excl + sum;
		__syncthreads();
	}
	if (total && threadIdx.x == 0) *total = carry;
}

// ---- exclusive scan of a LARGE array (the 128^3 cell table of the grid broadphase) over many CTAs: (1) per-chunk totals,
// (2) every CTA scans its chunk again and adds the totals of the chunks before it (at most a few thousand values: one
// block-wide reduction).  3 x n x 4 bytes of traffic, no inter-CTA waiting.
constexpr int LSCAN_THREADS = 512;
constexpr int LSCAN_ITEMS = 8;
constexpr int LSCAN_CHUNK = LSCAN_THREADS * LSCAN_ITEMS;
__global__ void __launch_bounds__(LSCAN_THREADS) largeScanTotalsKernel(const unsigned int* __restrict__ src, int n, unsigned int* __restrict__ chunkTotals)
{
	__shared__ unsigned int warpSums[LSCAN_THREADS / 32];
	const int base = blockIdx.x * LSCAN_CHUNK;
	unsigned int sum = 0;
	const uint4* p = reinterpret_cast<const uint4*>(src + base);
#pragma unroll
	for (int r = 0; r < LSCAN_ITEMS / 4; r++)
	{
		const int i = base + (r * LSCAN_THREADS + (int)threadIdx.x) * 4;
		if (i + 3 < n)
		{
			const uint4 v = p[r * LSCAN_THREADS + threadIdx.x];
			sum += v.x + v.y + v.z + v.w;
		}
		else
			for (int k = 0; k < 4; k++)
				if (i + k < n) sum += src[i + k];
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
	if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = sum;
	__syncthreads();
	if (threadIdx.x < 32)
	{
		unsigned int v = threadIdx.x < LSCAN_THREADS / 32 ? warpSums[threadIdx.x] : 0u;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if (threadIdx.x == 0) chunkTotals[blockIdx.x] = v;
	}
}
__global__ void __launch_bounds__(LSCAN_THREADS) largeScanApplyKernel(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst, int n,
																	  const unsigned int* __restrict__ chunkTotals, unsigned int* __restrict__ total)
{
	__shared__ unsigned int warpSums[LSCAN_THREADS / 32];
	__shared__ unsigned int sBase;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// totals of the chunks before this one
	unsigned int before = 0;
	for (int k = threadIdx.x; k < (int)blockIdx.x; k += LSCAN_THREADS) before += chunkTotals[k];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
	if (lane == 0) warpSums[warp] = before;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		unsigned int v = 0;
		for (int k = 0; k < LSCAN_THREADS / 32; k++) v += warpSums[k];
		sBase = v;
	}
	__syncthreads();
	const unsigned int chunkBase = sBase;
	__syncthreads();
	// this chunk: every thread owns LSCAN_ITEMS consecutive values
	const int first = blockIdx.x * LSCAN_CHUNK + threadIdx.x * LSCAN_ITEMS;
	unsigned int v[LSCAN_ITEMS];
	unsigned int sum = 0;
#pragma unroll
	for (int k = 0; k < LSCAN_ITEMS; k++)
	{
		v[k] = first + k < n ? src[first + k] : 0u;
		sum += v[k];
	}
	unsigned int incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) warpSums[warp] = incl;
	__syncthreads();
	if (warp == 0)
	{
		unsigned int w = lane < LSCAN_THREADS / 32 ? warpSums[lane] : 0u, wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
			if (lane >= o) wi += t;
		}
		if (lane < LSCAN_THREADS / 32) warpSums[lane] = wi - w;
	}
	__syncthreads();
	unsigned int run = chunkBase + warpSums[warp] + (incl - sum);
#pragma unroll
	for (int k = 0; k < LSCAN_ITEMS; k++)
	{
		if (first + k < n) dst[first + k] = run;
		run += v[k];
	}
	if (total && first <= n - 1 && n - 1 < first + LSCAN_ITEMS) *total = run;  // the thread that owns the last element
}
int exclusiveScanLargeU32(cudaStream_t s, const unsigned int* src, unsigned int* dst, int n, unsigned int* chunkTotals, unsigned int* totalDevice)
{
	if (n <= 0) return 0;
	const int chunks = divUp(n, LSCAN_CHUNK);
	largeScanTotalsKernel<<<chunks, LSCAN_THREADS, 0, s>>>(src, n, chunkTotals);
	B3_LAUNCH_CHECK();
	largeScanApplyKernel<<<chunks, LSCAN_THREADS, 0, s>>>(src, dst, n, chunkTotals, totalDevice);
	B3_LAUNCH_CHECK();
	return 0;
}
int largeScanChunks(int n) { return divUp(n > 0 ? n : 1, LSCAN_CHUNK); }

// ---- the digit table of one radix pass, [256][numTiles] digit-major: CTA d scans row d (exclusive, in place) and leaves the
// row total in digitTotals[d]; the scatter kernel adds the exclusive scan of the 256 totals itself.  Replaces one serial CTA
// walking the whole table (19 us per pass at 262 144 keys) by 256 short rows in parallel.
__global__ void __launch_bounds__(256) rsRowScanKernel(unsigned int* __restrict__ blockHist, int numTiles, unsigned int* __restrict__ digitTotals)
{
	__shared__ unsigned int warpSums[8];
	__shared__ unsigned int carry;
	unsigned int* row = blockHist + (size_t)blockIdx.x * numTiles;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < numTiles; base += 256)
	{
		const int i = base + threadIdx.x;
		const unsigned int v = i < numTiles ? row[i] : 0u;
		unsigned int incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) warpSums[warp] = incl;
		__syncthreads();
		unsigned int before = carry;
		for (int k = 0; k < warp; k++) before += warpSums[k];
		if (i < numTiles) row[i] = before + incl - v;
		__syncthreads();
		if (threadIdx.x == 255) carry = before + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) digitTotals[blockIdx.x] = carry;
}

int exclusiveScanU32(cudaStream_t s, const unsigned int* src, unsigned int* dst, int n, unsigned int* totalDevice)
{
	scanKernel<<<1, SCAN_THREADS, 0, s>>>(src, dst, n, totalDevice);
	B3_LAUNCH_CHECK();
	return 0;
}

template <typename KeyT, bool HAS_VALS>
static int radixSortImpl(cudaStream_t s, RadixSortTemp& tmp, KeyT* keys, KeyT* keysAlt, unsigned int* vals, unsigned int* valsAlt, int n, int numBits)
{
	if (n <= 1) return 0;
	int numBlocks = divUp(n, RS_TILE);
	B3_TRY(tmp.blockHist.reserve((size_t)256 * numBlocks + 256));  // + the 256 digit totals
	int passes = (numBits + 7) / 8;
	KeyT* kin = keys;
	KeyT* kout = keysAlt;
	unsigned int* vin = vals;
	unsigned int* vout = valsAlt;
	for (int p = 0; p < passes; p++)
	{
		int shift = p * 8;
		rsHistKernel<KeyT><<<numBlocks, RS_THREADS, 0, s>>>(kin, n, shift, tmp.blockHist.ptr, numBlocks);
		B3_LAUNCH_CHECK();
		rsRowScanKernel<<<256, 256, 0, s>>>(tmp.blockHist.ptr, numBlocks, tmp.blockHist.ptr + (size_t)256 * numBlocks);
		B3_LAUNCH_CHECK();
		rsScatterKernel<KeyT, HAS_VALS><<<numBlocks, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, tmp.blockHist.ptr, numBlocks, tmp.blockHist.ptr + (size_t)256 * numBlocks);
		B3_LAUNCH_CHECK();
		KeyT* t = kin;
		kin = kout;
		kout = t;
		unsigned int* tv = vin;
		vin = vout;
		vout = tv;
	}
	if (kin != keys)
	{
		B3_CUDA_CHECK(cudaMemcpyAsync(keys, kin, sizeof(KeyT) * n, cudaMemcpyDeviceToDevice, s));
		if (HAS_VALS) B3_CUDA_CHECK(cudaMemcpyAsync(vals, vin, sizeof(unsigned int) * n, cudaMemcpyDeviceToDevice, s));
	}
	return 0;
}

int radixSortKV32(cudaStream_t s, RadixSortTemp& tmp, unsigned int* keys, unsigned int* vals, int n, int numBits)
{
	B3_TRY(tmp.keysAlt.reserve(n));
	B3_TRY(tmp.valsAlt.reserve(n));
	return radixSortImpl<unsigned int, true>(s, tmp, keys, tmp.keysAlt.ptr, vals, tmp.valsAlt.ptr, n, numBits);
}
int radixSortKeys32(cudaStream_t s, RadixSortTemp& tmp, unsigned int* keys, int n, int numBits)
{
	B3_TRY(tmp.keysAlt.reserve(n));
	return radixSortImpl<unsigned int, false>(s, tmp, keys, tmp.keysAlt.ptr, nullptr, nullptr, n, numBits);
}
int radixSortKV64(cudaStream_t s, RadixSortTemp& tmp, unsigned long long* keys, unsigned int* vals, int n, int numBits)
{
	B3_TRY(tmp.keys64Alt.reserve(n));
	B3_TRY(tmp.valsAlt.reserve(n));
	return radixSortImpl<unsigned long long, true>(s, tmp, keys, tmp.keys64Alt.ptr, vals, tmp.valsAlt.ptr, n, numBits);
}

// ------------------------------------------------------------------ misc kernels
__global__ void fillKernel(unsigned int* dst, unsigned int value, int n, int offset)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[offset + i] = value;
}

// b3BoundSearchCL COUNT (b3BoundSearchCL.cpp:74-203): counts[b] = #{i : sorted[i].key == b}
__global__ void boundCountKernel(const b3b200_sort_data* __restrict__ sorted, int n, unsigned int* __restrict__ lower, unsigned int* __restrict__ upper, int numBuckets)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned int k = sorted[i].key;
	if (k >= (unsigned int)numBuckets) return;
	if (i == 0 || sorted[i - 1].key != k) lower[k] = i;
	if (i == n - 1 || sorted[i + 1].key != k) upper[k] = i + 1;
}
// b3BoundSearchCL BOUND_LOWER / BOUND_UPPER (b3BoundSearchCL.cpp:74-110, host twin :139-172): dst[k] = first index holding key k /
// one past the last; entries of keys that do not occur are left as the caller passed them
__global__ void boundLowerUpperKernel(const b3b200_sort_data* __restrict__ sorted, int n, unsigned int* __restrict__ dst, int numBuckets, int upper)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned int k = sorted[i].key;
	if (k >= (unsigned int)numBuckets) return;
	if (!upper && (i == 0 || sorted[i - 1].key != k)) dst[k] = i;
	if (upper && (i == n - 1 || sorted[i + 1].key != k)) dst[k] = i + 1;
}

// b3PrefixScanFloat4CL (b3PrefixScanFloat4CL.cpp:12-120): exclusive scan of the xyz of float4 values (w is not summed, like
// b3Vector3::operator+=), one CTA, warp-shuffle scans
constexpr int SCAN4_THREADS = 1024;
__global__ void __launch_bounds__(SCAN4_THREADS) scanFloat4Kernel(const float4* __restrict__ src, float4* __restrict__ dst, int n)
{
	__shared__ float4 warpSums[32];
	__shared__ float4 carry;
	if (threadIdx.x == 0) carry = make_float4(0.f, 0.f, 0.f, 0.f);
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int base = 0; base < n; base += SCAN4_THREADS)
	{
		const int i = base + threadIdx.x;
		float4 v = i < n ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
		float4 incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const float tx = __shfl_up_sync(0xffffffffu, incl.x, o), ty = __shfl_up_sync(0xffffffffu, incl.y, o), tz = __shfl_up_sync(0xffffffffu, incl.z, o);
			if (lane >= o)
			{
				incl.x += tx;
				incl.y += ty;
				incl.z += tz;
			}
		}
		if (lane == 31) warpSums[warp] = incl;
		__syncthreads();
		if (warp == 0)
		{
			float4 w = warpSums[lane], wi = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const float tx = __shfl_up_sync(0xffffffffu, wi.x, o), ty = __shfl_up_sync(0xffffffffu, wi.y, o), tz = __shfl_up_sync(0xffffffffu, wi.z, o);
				if (lane >= o)
				{
					wi.x += tx;
					wi.y += ty;
					wi.z += tz;
				}
			}
			warpSums[lane] = make_float4(wi.x - w.x, wi.y - w.y, wi.z - w.z, 0.f);  // exclusive
		}
		__syncthreads();
		const float4 c = carry, ws = warpSums[warp];
		if (i < n) dst[i] = make_float4(c.x + ws.x + (incl.x - v.x), c.y + ws.y + (incl.y - v.y), c.z + ws.z + (incl.z - v.z), 0.f);
		__syncthreads();
		if (threadIdx.x == SCAN4_THREADS - 1) carry = make_float4(c.x + ws.x + incl.x, c.y + ws.y + incl.y, c.z + ws.z + incl.z, 0.f);
		__syncthreads();
	}
}

__global__ void subKernel(const unsigned int* lower, const unsigned int* upper, unsigned int* counts, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) counts[i] = upper[i] - lower[i];
}

__global__ void splitSortDataKernel(const b3b200_sort_data* in, unsigned int* k, unsigned int* v, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		k[i] = in[i].key;
		v[i] = in[i].value;
	}
}
__global__ void joinSortDataKernel(b3b200_sort_data* out, const unsigned int* k, const unsigned int* v, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		out[i].key = k[i];
		out[i].value = v[i];
	}
}

}  // namespace b3b200

using namespace b3b200;

// ------------------------------------------------------------------ C ABI (host buffers)
extern "C" int b3b200_radix_sort_kv(int device, b3b200_sort_data* data, int n)
{
	if (n < 0 || (n > 0 && !data)) return B3B200_ERR_INVALID;
	if (n == 0) return 0;
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<b3b200_sort_data> d;
	DevBuf<unsigned int> k, v;
	RadixSortTemp tmp;
	B3_TRY(d.reserve(n));
	B3_TRY(k.reserve(n));
	B3_TRY(v.reserve(n));
	B3_CUDA_CHECK(cudaMemcpy(d.ptr, data, sizeof(b3b200_sort_data) * n, cudaMemcpyHostToDevice));
	splitSortDataKernel<<<divUp(n, 256), 256>>>(d.ptr, k.ptr, v.ptr, n);
	B3_LAUNCH_CHECK();
	B3_TRY(radixSortKV32(0, tmp, k.ptr, v.ptr, n, 32));
	joinSortDataKernel<<<divUp(n, 256), 256>>>(d.ptr, k.ptr, v.ptr, n);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemcpy(data, d.ptr, sizeof(b3b200_sort_data) * n, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int b3b200_radix_sort_keys(int device, unsigned int* keys, int n)
{
	if (n < 0 || (n > 0 && !keys)) return B3B200_ERR_INVALID;
	if (n == 0) return 0;
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<unsigned int> k;
	RadixSortTemp tmp;
	B3_TRY(k.reserve(n));
	B3_CUDA_CHECK(cudaMemcpy(k.ptr, keys, sizeof(unsigned int) * n, cudaMemcpyHostToDevice));
	B3_TRY(radixSortKeys32(0, tmp, k.ptr, n, 32));
	B3_CUDA_CHECK(cudaMemcpy(keys, k.ptr, sizeof(unsigned int) * n, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int b3b200_prefix_scan_u32(int device, const unsigned int* src, unsigned int* dst, int n, unsigned int* sum)
{
	if (n < 0 || (n > 0 && (!src || !dst))) return B3B200_ERR_INVALID;
	if (n == 0)
	{
		if (sum) *sum = 0;
		return 0;
	}
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<unsigned int> a, b, t;
	B3_TRY(a.reserve(n));
	B3_TRY(b.reserve(n));
	B3_TRY(t.reserve(1));
	B3_CUDA_CHECK(cudaMemcpy(a.ptr, src, sizeof(unsigned int) * n, cudaMemcpyHostToDevice));
	B3_TRY(exclusiveScanU32(0, a.ptr, b.ptr, n, t.ptr));
	B3_CUDA_CHECK(cudaMemcpy(dst, b.ptr, sizeof(unsigned int) * n, cudaMemcpyDeviceToHost));
	if (sum) B3_CUDA_CHECK(cudaMemcpy(sum, t.ptr, sizeof(unsigned int), cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int b3b200_bound_search_count(int device, const b3b200_sort_data* sorted, int n, unsigned int* counts, int numBuckets)
{
	if (n < 0 || numBuckets <= 0 || !counts || (n > 0 && !sorted)) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<b3b200_sort_data> d;
	DevBuf<unsigned int> lo, hi, c;
	B3_TRY(lo.reserve(numBuckets));
	B3_TRY(hi.reserve(numBuckets));
	B3_TRY(c.reserve(numBuckets));
	B3_CUDA_CHECK(cudaMemset(lo.ptr, 0, sizeof(unsigned int) * numBuckets));
	B3_CUDA_CHECK(cudaMemset(hi.ptr, 0, sizeof(unsigned int) * numBuckets));
	if (n > 0)
	{
		B3_TRY(d.reserve(n));
		B3_CUDA_CHECK(cudaMemcpy(d.ptr, sorted, sizeof(b3b200_sort_data) * n, cudaMemcpyHostToDevice));
		boundCountKernel<<<divUp(n, 256), 256>>>(d.ptr, n, lo.ptr, hi.ptr, numBuckets);
		B3_LAUNCH_CHECK();
	}
	subKernel<<<divUp(numBuckets, 256), 256>>>(lo.ptr, hi.ptr, c.ptr, numBuckets);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemcpy(counts, c.ptr, sizeof(unsigned int) * numBuckets, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int b3b200_fill_u32(int device, unsigned int* dst, unsigned int value, int n, int offset)
{
	if (n < 0 || offset < 0 || (n > 0 && !dst)) return B3B200_ERR_INVALID;
	if (n == 0) return 0;
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<unsigned int> d;
	B3_TRY(d.reserve((size_t)n + offset));
	B3_CUDA_CHECK(cudaMemcpy(d.ptr, dst, sizeof(unsigned int) * ((size_t)n + offset), cudaMemcpyHostToDevice));
	fillKernel<<<divUp(n, 256), 256>>>(d.ptr, value, n, offset);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemcpy(dst, d.ptr, sizeof(unsigned int) * ((size_t)n + offset), cudaMemcpyDeviceToHost));
	return 0;
}

// b3BoundSearchCL::execute with BOUND_LOWER (0) / BOUND_UPPER (1) / COUNT (2) (b3BoundSearchCL.cpp:74-137).  `dst` comes in with
// the caller's initial values (the reference only writes the entries of keys that occur) and goes out with the result.
extern "C" int b3b200_bound_search(int device, const b3b200_sort_data* sorted, int n, unsigned int* dst, int numBuckets, int option)
{
	if (option == 2) return b3b200_bound_search_count(device, sorted, n, dst, numBuckets);
	if (n < 0 || numBuckets <= 0 || !dst || (n > 0 && !sorted) || option < 0 || option > 2) return B3B200_ERR_INVALID;
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<b3b200_sort_data> d;
	DevBuf<unsigned int> out;
	B3_TRY(out.reserve(numBuckets));
	B3_CUDA_CHECK(cudaMemcpy(out.ptr, dst, sizeof(unsigned int) * numBuckets, cudaMemcpyHostToDevice));
	if (n > 0)
	{
		B3_TRY(d.reserve(n));
		B3_CUDA_CHECK(cudaMemcpy(d.ptr, sorted, sizeof(b3b200_sort_data) * n, cudaMemcpyHostToDevice));
		boundLowerUpperKernel<<<divUp(n, 256), 256>>>(d.ptr, n, out.ptr, numBuckets, option);
		B3_LAUNCH_CHECK();
	}
	B3_CUDA_CHECK(cudaMemcpy(dst, out.ptr, sizeof(unsigned int) * numBuckets, cudaMemcpyDeviceToHost));
	return 0;
}

// b3PrefixScanFloat4CL::execute (b3PrefixScanFloat4CL.cpp:44-93): exclusive scan of xyz; `sum` = the LAST OUTPUT element like the
// reference returns it (dst[n-1], not the grand total: b3PrefixScanFloat4CL.cpp:116-119)
extern "C" int b3b200_prefix_scan_float4(int device, const b3b200_float4* src, b3b200_float4* dst, int n, b3b200_float4* sum)
{
	if (n < 0 || (n > 0 && (!src || !dst))) return B3B200_ERR_INVALID;
	if (n == 0) return 0;
	B3_CUDA_CHECK(cudaSetDevice(device));
	DevBuf<float4> a, b;
	B3_TRY(a.reserve(n));
	B3_TRY(b.reserve(n));
	B3_CUDA_CHECK(cudaMemcpy(a.ptr, src, sizeof(float4) * n, cudaMemcpyHostToDevice));
	scanFloat4Kernel<<<1, SCAN4_THREADS>>>(a.ptr, b.ptr, n);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemcpy(dst, b.ptr, sizeof(float4) * n, cudaMemcpyDeviceToHost));
	if (sum) *sum = dst[n - 1];
	return 0;
}
