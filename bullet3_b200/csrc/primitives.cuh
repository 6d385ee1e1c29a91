// primitives.cuh -- device-side parallel primitives used by the step:
// stable LSD radix sort (8-bit digits, warp match-any ranking), exclusive scan.
// Replaces b3RadixSort32CL / b3PrefixScanCL / b3BoundSearchCL / b3FillCL
// (src/Bullet3OpenCL/ParallelPrimitives/*.cpp) -- same results, new algorithms.
#pragma once
#include "common.cuh"

namespace b3b200
{
struct RadixSortTemp
{
	DevBuf<unsigned int> blockHist;  // [256][numBlocks]
	DevBuf<unsigned int> keysAlt;
	DevBuf<unsigned int> valsAlt;
	DevBuf<unsigned long long> keys64Alt;
};

// Sort n (key,value) pairs by the key bits [0, numBits).  Stable.  The result
// ends up back in keys/vals (ping-pong through the temp buffers).
int radixSortKV32(cudaStream_t s, RadixSortTemp& tmp, unsigned int* keys, unsigned int* vals, int n, int numBits);
int radixSortKeys32(cudaStream_t s, RadixSortTemp& tmp, unsigned int* keys, int n, int numBits);
int radixSortKV64(cudaStream_t s, RadixSortTemp& tmp, unsigned long long* keys, unsigned int* vals, int n, int numBits);

// exclusive scan of n u32; if total != nullptr the grand total is written there (device pointer)
// the same for large n over many CTAs; chunkTotals: scratch of largeScanChunks(n) words
int exclusiveScanLargeU32(cudaStream_t s, const unsigned int* src, unsigned int* dst, int n, unsigned int* chunkTotals, unsigned int* totalDevice);
int largeScanChunks(int n);
int exclusiveScanU32(cudaStream_t s, const unsigned int* src, unsigned int* dst, int n, unsigned int* totalDevice);

}  // namespace b3b200
