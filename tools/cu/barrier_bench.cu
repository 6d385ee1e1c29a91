// development aid: latency of software grid barrier variants on 148 co-resident CTAs (run under gpurun)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned int ldRelaxed(const unsigned int* p)
{
	unsigned int v;
	asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// VARIANT 0: red.release + relaxed poll + fence.acq_rel   1: no acquire fence   2: flags per CTA (st.release each, all polled by a warp-set), no acquire fence
// 3: like 1 but the red is relaxed and preceded by __threadfence() of thread 0 only
// PAYLOAD: every thread stores 2 float4 (st.cg) before arriving and loads 2 (ld.cg) after the wait, from another CTA's slots
template <int VARIANT, int PAYLOAD>
__global__ void __launch_bounds__(256, 1) barrierKernel(unsigned int* counter, unsigned int* flags, float4* data, int iters, long long* out)
{
	unsigned int target = 0;
	const unsigned int nb = gridDim.x;
	float4 acc = make_float4(0, 0, 0, 0);
	long long t0 = 0;
	for (int it = 0; it < iters + 10; it++)
	{
		if (it == 10) t0 = clock64();
		if (PAYLOAD)
		{
			float4 v = make_float4(it, threadIdx.x, blockIdx.x, acc.x);
			__stcg(&data[(size_t)blockIdx.x * 512 + threadIdx.x], v);
			__stcg(&data[(size_t)blockIdx.x * 512 + 256 + threadIdx.x], v);
		}
		if (VARIANT == 2)
		{
			__syncthreads();
			if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"((unsigned int)(it + 1)) : "memory");
			if (threadIdx.x < nb)
			{
				while (ldRelaxed(flags + threadIdx.x * 32) < (unsigned int)(it + 1)) {}
			}
			__syncthreads();
		}
		else
		{
			target += nb;
			__syncthreads();
			if (threadIdx.x == 0)
			{
				if (VARIANT == 3)
				{
					__threadfence();
					asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
				}
				else
					asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
				unsigned int v = target - 1u;
				while ((int)(v - target) < 0) v = ldRelaxed(counter);
				if (VARIANT == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
			}
			__syncthreads();
		}
		if (PAYLOAD)
		{
			const unsigned int src = (blockIdx.x + 1 + it) % nb;
			float4 a = __ldcg(&data[(size_t)src * 512 + threadIdx.x]);
			float4 b = __ldcg(&data[(size_t)src * 512 + 256 + threadIdx.x]);
			if (a.x != (float)it || b.x != (float)it) acc.y += 1.f;  // stale data seen
			acc.x += a.y + b.z;
		}
	}
	long long t1 = clock64();
	if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
	if (acc.y != 0.f && threadIdx.x == 0) out[gridDim.x + blockIdx.x] = (long long)acc.y;
	if (acc.x == 12345.f) out[0] = 0;
}

template <int V, int P>
int run(const char* name, int grid)
{
	unsigned int *counter, *flags;
	float4* data;
	long long* out;
	CHECK(cudaMalloc(&counter, 256));
	CHECK(cudaMalloc(&flags, 148 * 32 * 4));
	CHECK(cudaMalloc(&data, sizeof(float4) * 512 * 148));
	CHECK(cudaMalloc(&out, sizeof(long long) * 2 * 148));
	const int iters = 2000;
	for (int rep = 0; rep < 2; rep++)
	{
		CHECK(cudaMemset(counter, 0, 256));
		CHECK(cudaMemset(flags, 0, 148 * 32 * 4));
		CHECK(cudaMemset(out, 0, sizeof(long long) * 2 * 148));
		int it = iters;
		void* args[] = {&counter, &flags, &data, &it, &out};
		CHECK(cudaLaunchCooperativeKernel((const void*)barrierKernel<V, P>, dim3(grid), dim3(256), args, 0, 0));
		CHECK(cudaDeviceSynchronize());
	}
	long long h[2 * 148];
	CHECK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
	long long stale = 0;
	for (int i = 0; i < grid; i++) stale += h[148 + i];
	int clk = 0;
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	printf("%-44s grid %3d: %7.1f cycles = %.2f us per barrier (clock %d kHz)  stale reads: %lld\n", name, grid, (double)h[0] / iters, (double)h[0] / iters / (clk * 1e-3), clk, stale);
	cudaFree(counter), cudaFree(flags), cudaFree(data), cudaFree(out);
	return 0;
}

int main()
{
	for (int grid : {148, 74, 16})
	{
		run<0, 0>("red.release + poll + fence.acq_rel", grid);
		run<1, 0>("red.release + poll", grid);
		run<3, 0>("threadfence + red.relaxed + poll", grid);
		run<2, 0>("flag per CTA (st.release), polled by threads", grid);
		run<0, 1>("red.release + poll + fence.acq_rel, payload", grid);
		run<1, 1>("red.release + poll, payload", grid);
		run<2, 1>("flag per CTA, payload", grid);
	}
	return 0;
}
