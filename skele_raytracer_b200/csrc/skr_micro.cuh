// skr_micro.cuh -- roofline denominators measured on the device the frames run on (MEASURED_PEAKS.json holds HBM and
// tensor figures only): FP32 FMA throughput and the bandwidth of the memory levels the BVH / primitive fetch of the
// triangle path goes through (shared memory, L1, L2).  Used by bench.py through skr_measure_fp32_peak /
// skr_measure_bandwidth; never on a frame's path.
#pragma once
#include <cuda_runtime.h>

// FP32 FMA peak: 8 independent FFMA chains per thread, register resident.
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b)
{
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 16; k++)
		{
			x0 = fmaf(x0, a, b);
			x1 = fmaf(x1, a, b);
			x2 = fmaf(x2, a, b);
			x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b);
			x5 = fmaf(x5, a, b);
			x6 = fmaf(x6, a, b);
			x7 = fmaf(x7, a, b);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// Shared-memory read bandwidth: every thread streams LDS.128 over a 16 KB window of its CTA (conflict-free: consecutive
// lanes read consecutive float4).  8 loads in flight per thread.
__global__ void __launch_bounds__(256) lds_bw_kernel(float *out, int iters)
{
	__shared__ float4 win[1024];
	for(int i = threadIdx.x; i < 1024; i += 256)
	{
		win[i] = make_float4(i, 1, 2, 3);
	}
	__syncthreads();
	float4 acc = make_float4(0, 0, 0, 0);
	unsigned at = threadIdx.x;
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 8; k++)
		{
			float4 v;
			const unsigned sa = (unsigned) __cvta_generic_to_shared(win + ((at + 256u * k) & 1023u));
			asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa));
			acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
		}
		at += 32u;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// L1 (CACHE = 0: ld.global.ca over a 16 KB window per CTA, resident in the SM's L1 after the first pass) and
// L2 (CACHE = 1: ld.global.cg, which bypasses L1, over a buffer that fits the 126 MB L2 but not L1) read bandwidth.
template <int CACHE>
__global__ void __launch_bounds__(256) gmem_bw_kernel(const float4 *__restrict__ buf, unsigned mask, unsigned cta_stride, float *out, int iters)
{
	float4 acc	= make_float4(0, 0, 0, 0);
	unsigned at = blockIdx.x * cta_stride + threadIdx.x;
	for(int i = 0; i < iters; i++)
	{
#pragma unroll
		for(int k = 0; k < 8; k++)
		{
			const float4 *p = buf + ((at + 256u * k) & mask);
			float4 v;
			if(CACHE == 0)
			{
				asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
			}
			else
			{
				asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
			}
			acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
		}
		at += CACHE == 0 ? 32u : 2048u;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
