#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k1(float *out, int iters, float a, float b)
{
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 16; k++)
		{
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void __launch_bounds__(256) k2(float *out, int iters, float a, float b)
{
	float2 A = make_float2(a, a), Bv = make_float2(b, b);
	float t = threadIdx.x;
	float2 x0 = make_float2(t, t + 1), x1 = make_float2(t + 2, t + 3), x2 = make_float2(t + 4, t + 5), x3 = make_float2(t + 6, t + 7);
	float2 x4 = make_float2(t + 8, t + 9), x5 = make_float2(t + 10, t + 11), x6 = make_float2(t + 12, t + 13), x7 = make_float2(t + 14, t + 15);
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 16; k++)
		{
			x0 = __ffma2_rn(x0, A, Bv); x1 = __ffma2_rn(x1, A, Bv); x2 = __ffma2_rn(x2, A, Bv); x3 = __ffma2_rn(x3, A, Bv);
			x4 = __ffma2_rn(x4, A, Bv); x5 = __ffma2_rn(x5, A, Bv); x6 = __ffma2_rn(x6, A, Bv); x7 = __ffma2_rn(x7, A, Bv);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x1.x + x2.x + x3.x + x4.x + x5.x + x6.x + x7.x + x0.y + x1.y + x2.y + x3.y + x4.y + x5.y + x6.y + x7.y;
}
// mixed: FFMA2 + independent integer ops, to see whether FFMA2 frees issue slots
__global__ void __launch_bounds__(256) k3(float *out, int iters, float a, float b, int q)
{
	float2 A = make_float2(a, a), Bv = make_float2(b, b);
	float t = threadIdx.x;
	float2 x0 = make_float2(t, t + 1), x1 = make_float2(t + 2, t + 3), x2 = make_float2(t + 4, t + 5), x3 = make_float2(t + 6, t + 7);
	int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 16; k++)
		{
			x0 = __ffma2_rn(x0, A, Bv); x1 = __ffma2_rn(x1, A, Bv); x2 = __ffma2_rn(x2, A, Bv); x3 = __ffma2_rn(x3, A, Bv);
			i0 = (i0 ^ q) + i1; i1 = (i1 ^ q) + i2; i2 = (i2 ^ q) + i3; i3 = (i3 ^ q) + i0;
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x1.x + x2.x + x3.x + x0.y + x1.y + x2.y + x3.y + (float) (i0 + i1 + i2 + i3);
}
__global__ void __launch_bounds__(256) k4(float *out, int iters, float a, float b, int q)
{
	float t = threadIdx.x;
	float x0 = t, x1 = t + 1, x2 = t + 2, x3 = t + 3, x4 = t + 4, x5 = t + 5, x6 = t + 6, x7 = t + 7;
	int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 16; k++)
		{
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
			i0 = (i0 ^ q) + i1; i1 = (i1 ^ q) + i2; i2 = (i2 ^ q) + i3; i3 = (i3 ^ q) + i0;
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float) (i0 + i1 + i2 + i3);
}
int main()
{
	int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
	int blocks = sm * 8, threads = 256, iters = 4096;
	float *d; cudaMalloc(&d, sizeof(float) * blocks * threads);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for(int which = 1; which <= 4; which++)
	{
		float best = 1e9;
		for(int rep = 0; rep < 5; rep++)
		{
			cudaEventRecord(e0);
			if(which == 1) k1<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
			if(which == 2) k2<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
			if(which == 3) k3<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f, 12345);
			if(which == 4) k4<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f, 12345);
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1); if(ms < best) best = ms;
		}
		double fmas = (which == 1 ? 8.0 : which == 2 ? 16.0 : 8.0) * 16.0 * iters * (double) blocks * threads;
		printf("kernel %d: %.3f ms  %.1f TFLOP/s fp32 (%s)\n", which, best, 2 * fmas / (best * 1e-3) / 1e12,
			   which == 1 ? "8 FFMA chains" : which == 2 ? "8 FFMA2 chains" : which == 3 ? "4 FFMA2 + 8 int ops" : "8 FFMA + 8 int ops");
	}
	return 0;
}
