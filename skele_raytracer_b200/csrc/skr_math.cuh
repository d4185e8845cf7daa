// skr_math.cuh -- float3 helpers for the device code.
//
// Two flavours of every geometric primitive:
//   *_rn  : explicit round-to-nearest intrinsics that nvcc never contracts into FMA and that keep the
//           reference's operand order (glm 0.9.5.4: dot = (x0*y0 + x1*y1) + x2*y2, reference
//           src/glm/detail/func_geometric.inl:66-73; normalize = v * (1.0f / sqrt(x*x+y*y+z*z)), :256-265).
//           Used where a value feeds a discontinuity (ray generation, hit point, normal, triangle leaf test).
//   plain : ordinary expressions, free to become FFMA.  Used in the per-sphere test loops and in colour math.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define SKR_DEV __device__ __forceinline__

SKR_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SKR_DEV float3 f3(const float4 &v) { return make_float3(v.x, v.y, v.z); }
SKR_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SKR_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SKR_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
SKR_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
SKR_DEV float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
SKR_DEV float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
SKR_DEV float3 &operator+=(float3 &a, float3 b)
{
	a.x += b.x;
	a.y += b.y;
	a.z += b.z;
	return a;
}
SKR_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SKR_DEV float3 cross(float3 x, float3 y) { return f3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }

// ---- uncontracted, reference-ordered ----
SKR_DEV float dot_rn(float3 a, float3 b)
{
	return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
SKR_DEV float3 sub_rn(float3 a, float3 b) { return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
SKR_DEV float3 add_rn(float3 a, float3 b) { return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
SKR_DEV float3 muls_rn(float3 a, float s) { return f3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
SKR_DEV float3 adds_rn(float3 a, float s) { return f3(__fadd_rn(a.x, s), __fadd_rn(a.y, s), __fadd_rn(a.z, s)); }
SKR_DEV float3 cross_rn(float3 x, float3 y)
{
	return f3(__fsub_rn(__fmul_rn(x.y, y.z), __fmul_rn(y.y, x.z)), __fsub_rn(__fmul_rn(x.z, y.x), __fmul_rn(y.z, x.x)),
			  __fsub_rn(__fmul_rn(x.x, y.y), __fmul_rn(y.x, x.y)));
}
// glm::normalize(vec3): v * (1.0f / sqrt(dot)) with IEEE sqrt and divide
SKR_DEV float3 normalize_rn(float3 v)
{
	float sqr = dot_rn(v, v);
	float inv = __fdiv_rn(1.0f, __fsqrt_rn(sqr));
	return muls_rn(v, inv);
}
// colour-path normalize: rsqrt approximation (rel. error ~2^-22), fine for shading terms
SKR_DEV float3 normalize_fast(float3 v) { return v * rsqrtf(dot(v, v)); }

// ---- packed FP32x2 (Blackwell FFMA2/FADD2/FMUL2): two lanes per issue slot.  Measured on B200
// (scripts/micro/fma2_peak.cu): same FP32 peak as FFMA (74 vs 72 TFLOP/s) but half the issue slots, so an
// issue-bound mix of FP and integer/control instructions runs ~1.45x faster.
SKR_DEV float2 f2(float a, float b) { return make_float2(a, b); }
SKR_DEV float2 splat2(float a) { return make_float2(a, a); }
SKR_DEV float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
SKR_DEV float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
SKR_DEV float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// sqrt.approx (MUFU-based, ~2^-22 relative error, NaN for negative inputs): used only to RANK candidate hits; the
// winner's distance is always recomputed with IEEE operations.
SKR_DEV float sqrt_approx(float x)
{
	float r;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

SKR_DEV float u2f(uint32_t u) { return __uint_as_float(u); }
SKR_DEV uint32_t f2u(float f) { return __float_as_uint(f); }

// 256-bit read-only global load (sm_100: LDG.E.256): 32 B per lane in ONE instruction.  The BVH walk is a gather -- every
// lane its own node -- and what it is bound by is the number of load INSTRUCTIONS x lanes (one L1 tag lookup per lane per
// instruction, not bytes): a 64 B node costs two of these instead of four LDG.128.  p must be 32-byte aligned.
SKR_DEV void ldg256(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
		: "l"(p));
}
