// skr_bvh_build.cuh -- device-side LBVH construction at scene upload.
//
//   1. tri_bounds_kernel   : bounds of each MIRRORED triangle (v0, 2*v0-v1, v2) (see skr_bvh.cuh) -- or of the triangle
//                            itself for the shaded-triangles hierarchy --, slightly inflated,
//                            and the scene bounds (block reduce in shared memory + float atomics).
//   2. morton_kernel       : 63-bit Morton code (21 bits/axis) of each box centre.  30-bit codes are not enough:
//                            dragon.scn's ground quad stretches the scene box to +-20 while the model is 0.2 wide.
//   3. radix sort          : LSD, 8 bits/pass, 8 passes over the 64-bit keys, values = triangle ids.
//                            Per pass: per-warp digit histograms (shared-memory counters) -> exclusive scan of the
//                            [digit][warp] table -> stable scatter (warp match + shared-memory running counters).
//   4. karras_kernel       : Karras 2012 "Maximizing Parallelism in the Construction of BVHs": every internal node
//                            finds its key range and split from common-prefix lengths (ties broken by index).
//   5. refit_kernel        : bottom-up; the second thread to reach a node (atomic counter) merges the child boxes
//                            and writes the 64-byte node with both child boxes in it.
//   6. gather_kernel       : triangles re-laid in leaf order, 64 B each (v0 | original index, v1, v2, pad): one 256-bit and one
//                            128-bit load in the leaf test.
#pragma once
#include "skr_math.cuh"

namespace bvhb
{
constexpr int SORT_THREADS		  = 256;
constexpr int SORT_WARPS		  = SORT_THREADS / 32;
constexpr int SORT_ITEMS_PER_WARP = 32 * 32; // each warp owns 1024 consecutive keys (256 for small scenes: sort_items_per_warp)
// Small scenes are launch- and latency-bound: 10 002 triangles in chunks of 1024 keep 10 warps of the whole GPU busy, each
// walking its chunk in 32 dependent steps (18 + 29 us per pass, 8 passes).  Chunks of 256 give 40 warps and 8 steps;
// between 11 k and 45 k keys the chunk grows so that the table of 256 counters per warp still fits the one-round scan.
constexpr int SCAN_SMALL = 11264; // entries of a histogram table scanned in one round (44 KB of shared memory): 44 warps x 256 digits
inline int sort_items_per_warp(int n)
{
	const int max_warps = SCAN_SMALL / 256; // so that the table is scanned in one round
	if(n <= 256 * max_warps)
	{
		return 256;
	}
	if(n <= SORT_ITEMS_PER_WARP * max_warps)
	{
		return ((n + max_warps - 1) / max_warps + 31) / 32 * 32;
	}
	return SORT_ITEMS_PER_WARP;
}

struct Box
{
	float3 lo, hi;
};

__device__ __forceinline__ void atomic_min_f(float *addr, float v)
{
	// valid for any sign: ints order like floats for >= 0, reversed for < 0
	if(v >= 0.0f)
	{
		atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
	}
	else
	{
		atomicMax(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
	}
}
__device__ __forceinline__ void atomic_max_f(float *addr, float v)
{
	if(v >= 0.0f)
	{
		atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
	}
	else
	{
		atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
	}
}

// DEAD-TRIANGLE BOUND g (see tri_bounds_kernel): a ray with |dir| * g < 1e-5 cannot pass the reference's fabs(det) >= 1e-5 test
__device__ __forceinline__ float tri_dead_bound(float3 v0, float3 v1, float3 v2)
{
	const double e1x = (double) v1.x - v0.x, e1y = (double) v1.y - v0.y, e1z = (double) v1.z - v0.z;
	const double e2x = (double) v2.x - v0.x, e2y = (double) v2.y - v0.y, e2z = (double) v2.z - v0.z;
	const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
	const double g	= sqrt(nx * nx + ny * ny + nz * nz) + 1.0e-6 * sqrt((e1x * e1x + e1y * e1y + e1z * e1z) * (e2x * e2x + e2y * e2y + e2z * e2z));
	return __double2float_ru(g * 1.000001);
}

__global__ void init_scene_box_kernel(float *scene_box) // (min.xyz, max.xyz) = (+3e38, -3e38)
{
	if(threadIdx.x < 6)
	{
		scene_box[threadIdx.x] = threadIdx.x < 3 ? 3.0e38f : -3.0e38f;
	}
}

// tris: [T][9] floats as uploaded.  box_lo/box_hi: float4 per triangle.  scene_box: 6 floats (lo, hi), pre-set to +-FLT_MAX.
__global__ void tri_bounds_kernel(const float *__restrict__ tris, int T, float4 *__restrict__ box_lo, float4 *__restrict__ box_hi, float *scene_box, int mirror)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float3 lo = f3(3.0e38f, 3.0e38f, 3.0e38f), hi = f3(-3.0e38f, -3.0e38f, -3.0e38f);
	if(i < T)
	{
		const float *t	= tris + 9 * (size_t) i;
		const float3 v0 = f3(t[0], t[1], t[2]), v1 = f3(t[3], t[4], t[5]), v2 = f3(t[6], t[7], t[8]);
		// mirrored vertex (the reference's query, skr_bvh.cuh), or the vertex itself (shaded-triangles mode)
		const float3 m1 = mirror ? f3(2.0f * v0.x - v1.x, 2.0f * v0.y - v1.y, 2.0f * v0.z - v1.z) : v1;
		lo				= f3(fminf(fminf(v0.x, m1.x), v2.x), fminf(fminf(v0.y, m1.y), v2.y), fminf(fminf(v0.z, m1.z), v2.z));
		hi				= f3(fmaxf(fmaxf(v0.x, m1.x), v2.x), fmaxf(fmaxf(v0.y, m1.y), v2.y), fmaxf(fmaxf(v0.z, m1.z), v2.z));
		// inflate: the reference's float u/v window can accept points a few ulps outside the exact triangle
		const float diag = fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
		const float mag	 = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
		const float pad	 = 1.0e-3f * diag + 4.0e-6f * mag + 1.0e-30f;
		lo				 = f3(lo.x - pad, lo.y - pad, lo.z - pad);
		hi				 = f3(hi.x + pad, hi.y + pad, hi.z + pad);
		// DEAD-TRIANGLE BOUND.  The reference rejects a triangle when fabs(det) < 1e-5 (src/utils.h:190), det = e1 . (dir x e2)
		// = -dir . (e1 x e2): whatever the ray, |det| <= |dir| * |e1 x e2|, and the float evaluation (uncontracted cross and
		// dot, tri_test_ref) adds at most ~11 eps |dir| |e1| |e2|.  With g = |e1 x e2| + 1e-6 |e1| |e2| (rounded up), a ray
		// with |dir| * g < 1e-5 cannot pass the test: traversal skips the triangle -- and every subtree whose largest g fails
		// (refit_kernel keeps the maximum per child).  dragon.scn's triangles are ~2 mm across (|e1 x e2| ~ 4e-6): for the
		// shorter camera rays most of the model can never be hit, which is exactly why the reference's render of it is sparse.
		box_lo[i]		= make_float4(lo.x, lo.y, lo.z, tri_dead_bound(v0, v1, v2));
		box_hi[i]		= make_float4(hi.x, hi.y, hi.z, 0.0f);
	}
	// block reduction of the scene box through shared memory, one atomic per block per component
	__shared__ float red[6][32];
	float v[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
#pragma unroll
	for(int k = 0; k < 6; k++)
	{
#pragma unroll
		for(int off = 16; off > 0; off >>= 1)
		{
			const float o = __shfl_xor_sync(0xffffffffu, v[k], off);
			v[k]		  = k < 3 ? fminf(v[k], o) : fmaxf(v[k], o);
		}
		if((threadIdx.x & 31) == 0)
		{
			red[k][threadIdx.x >> 5] = v[k];
		}
	}
	__syncthreads();
	if(threadIdx.x < 6)
	{
		const int k = threadIdx.x;
		float r		= red[k][0];
		for(int w = 1; w < (int) (blockDim.x >> 5); w++)
		{
			r = k < 3 ? fminf(r, red[k][w]) : fmaxf(r, red[k][w]);
		}
		if(k < 3)
		{
			atomic_min_f(scene_box + k, r);
		}
		else
		{
			atomic_max_f(scene_box + k, r);
		}
	}
}

__device__ __forceinline__ unsigned long long expand21(unsigned v) // spread 21 bits to every third bit
{
	unsigned long long x = v & 0x1fffffull;
	x					 = (x | x << 32) & 0x1f00000000ffffull;
	x					 = (x | x << 16) & 0x1f0000ff0000ffull;
	x					 = (x | x << 8) & 0x100f00f00f00f00full;
	x					 = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x					 = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

// Also pulls OUTSIZED triangles out of the hierarchy (dragon.scn: two ground triangles spanning the whole scene around a
// model 1/200 of its size; inside an LBVH they inflate every ancestor box, so that every ray walks ~10 nodes): a
// triangle whose box surface exceeds SKR_BIG_TRI_FRACTION of the scene box's goes to a short list (big_v, at most
// big_cap entries) that traversal tests first -- any hit ends the query -- and its leaf box collapses to its centre, so
// it no longer widens its ancestors.  (The leaf stays in the tree: harmless, its test is exact whatever the box.)
#define SKR_BIG_TRI_FRACTION 0.1f
// measured on config 4 (dragon.scn 1080p): no class bits 0.218 ms / 15.3 M node visits; 1 bit 0.202 / 14.0 M; 2 bits 0.206; 3 bits
// 0.208-0.211 (13.8-14.0 M visits, but three more levels above every class)
#ifndef SKR_GCLASS_BITS
#define SKR_GCLASS_BITS 0
#endif
#ifndef SKR_MORTON_CUBIC
#define SKR_MORTON_CUBIC 1
#endif
#ifndef SKR_GCLASS_PER_OCTAVE
#define SKR_GCLASS_PER_OCTAVE 2
#endif
__global__ void morton_kernel(float4 *box_lo, float4 *box_hi, const float *__restrict__ scene_box, int T, unsigned long long *__restrict__ keys,
							  unsigned *__restrict__ vals, const float *__restrict__ tris, int *big_count, float4 *__restrict__ big_v, int big_cap,
							  int gclass_bits)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= T)
	{
		return;
	}
	const float3 slo = f3(scene_box[0], scene_box[1], scene_box[2]);
	const float3 shi = f3(scene_box[3], scene_box[4], scene_box[5]);
	const float4 lo = box_lo[i], hi = box_hi[i];
	const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
	const float ex = fmaxf(shi.x - slo.x, 1e-30f), ey = fmaxf(shi.y - slo.y, 1e-30f), ez = fmaxf(shi.z - slo.z, 1e-30f);
	if(big_cap > 0)
	{
		const float bx = hi.x - lo.x, by = hi.y - lo.y, bz = hi.z - lo.z;
		if(bx * by + by * bz + bz * bx > SKR_BIG_TRI_FRACTION * (ex * ey + ey * ez + ez * ex))
		{
			const int slot = atomicAdd(big_count, 1);
			if(slot < big_cap)
			{
				const float *t		= tris + 9 * (size_t) i;
				big_v[3 * slot + 0] = make_float4(t[0], t[1], t[2], __int_as_float(i)); // w: original triangle index
				big_v[3 * slot + 1] = make_float4(t[3], t[4], t[5], 0.0f);
				big_v[3 * slot + 2] = make_float4(t[6], t[7], t[8], 0.0f);
				box_lo[i] = box_hi[i] = make_float4(cx, cy, cz, 0.0f); // (w = 0: its leaf in the tree is never worth a test -- the big list has it)
			}
		}
	}
	// quantise in double: 21 bits exceed the float mantissa's headroom near the top of the range
#if SKR_MORTON_CUBIC
	// ONE cell size for the three axes (the scene box's longest side): with each axis scaled by its own extent a flat scene --
	// dragon.scn: a 0.2-high model on a 40 x 40 ground -- has cells 200 times longer than high, and the top levels of the
	// model's hierarchy are horizontal slabs that span its whole footprint
	const double emax = fmax((double) ex, fmax((double) ey, (double) ez));
	const unsigned qx = (unsigned) fmin(fmax((double) (cx - slo.x) / emax * 2097152.0, 0.0), 2097151.0);
	const unsigned qy = (unsigned) fmin(fmax((double) (cy - slo.y) / emax * 2097152.0, 0.0), 2097151.0);
	const unsigned qz = (unsigned) fmin(fmax((double) (cz - slo.z) / emax * 2097152.0, 0.0), 2097151.0);
#else
	const unsigned qx = (unsigned) fmin(fmax((double) (cx - slo.x) / (double) ex * 2097152.0, 0.0), 2097151.0);
	const unsigned qy = (unsigned) fmin(fmax((double) (cy - slo.y) / (double) ey * 2097152.0, 0.0), 2097151.0);
	const unsigned qz = (unsigned) fmin(fmax((double) (cz - slo.z) / (double) ez * 2097152.0, 0.0), 2097151.0);
#endif
	unsigned long long key = expand21(qx) << 2 | expand21(qy) << 1 | expand21(qz);
	if(gclass_bits > 0)
	{
		// CLASSES BY DEAD-TRIANGLE BOUND (line query only).  Whether a triangle can be hit at all depends on the ray's length:
		// g |dir| >= 1e-5 (tri_bounds_kernel).  In a purely spatial order dead and live triangles alternate, so no subtree is
		// ever dead as a whole and a line that crosses the mesh THROUGH its dead triangles (57 % of dragon.scn's for a unit
		// ray) still descends around every live neighbour's box next to many dead ones.  The class -- 0: g >= 1e-5 (live for
		// every ray of length >= 1), then SKR_GCLASS_PER_OCTAVE classes per halving of g -- goes ABOVE the spatial bits of the
		// key: the top levels of the tree split by class, each class is its own spatial hierarchy with boxes around ITS
		// triangles only, and a ray prunes every class that is dead for it at the top (the per-child g of refit_kernel).
		const float g  = box_lo[i].w; // (0 for a triangle moved to the big list: last class)
		const float r  = g > 0.0f ? 1.0e-5f / g : 3.0e38f;
		const int cmax = (1 << gclass_bits) - 1;
		const int c	   = r <= 1.0f ? 0 : min(cmax, 1 + (int) floorf((float) SKR_GCLASS_PER_OCTAVE * log2f(r)));
		key			   = (key >> gclass_bits) | ((unsigned long long) c << (63 - gclass_bits));
	}
	keys[i] = key;
	vals[i] = (unsigned) i;
}

// ---- radix sort -------------------------------------------------------------------------------
// Table layout: hist[digit * nwarps + warp].  Warp w owns keys [w*ipw, (w+1)*ipw), ipw = sort_items_per_warp(n).

__global__ void sort_hist_kernel(const unsigned long long *__restrict__ keys, int n, int shift, unsigned *__restrict__ hist, int nwarps, int ipw)
{
	__shared__ unsigned cnt[SORT_WARPS][256];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int warp = blockIdx.x * SORT_WARPS + wib;
	for(int d = lane; d < 256; d += 32)
	{
		cnt[wib][d] = 0;
	}
	__syncwarp();
	if(warp < nwarps)
	{
		const int base = warp * ipw;
		for(int it = 0; it < ipw / 32; it++)
		{
			const int i = base + it * 32 + lane;
			if(i < n)
			{
				atomicAdd(&cnt[wib][(unsigned) (keys[i] >> shift) & 255u], 1u);
			}
		}
		__syncwarp();
		for(int d = lane; d < 256; d += 32)
		{
			hist[(size_t) d * nwarps + warp] = cnt[wib][d];
		}
	}
}

// single-block exclusive scan over `len` entries (len = 256 * nwarps), in place
__global__ void sort_scan_kernel(unsigned *__restrict__ data, int len)
{
	__shared__ unsigned warp_sums[32];
	__shared__ unsigned carry;
	if(threadIdx.x == 0)
	{
		carry = 0;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
	__shared__ unsigned buf[SCAN_SMALL];
	if(len <= SCAN_SMALL)
	{
		// small tables (small scenes): ONE round through shared memory (coalesced in, coalesced out) -- every thread scans its
		// own run of consecutive entries, the block scans the run totals -- instead of len / blockDim.x rounds of four barriers
		for(int i = threadIdx.x; i < len; i += blockDim.x)
		{
			buf[i] = data[i];
		}
		__syncthreads();
		const int per = (len + (int) blockDim.x - 1) / (int) blockDim.x, b = (int) threadIdx.x * per;
		unsigned sum = 0;
		for(int k = 0; k < per; k++)
		{
			const int i = b + k;
			if(i < len)
			{
				const unsigned v = buf[i];
				buf[i]			 = sum; // exclusive within the run
				sum += v;
			}
		}
		unsigned s = sum;
#pragma unroll
		for(int off = 1; off < 32; off <<= 1)
		{
			const unsigned o = __shfl_up_sync(0xffffffffu, s, off);
			if(lane >= off)
			{
				s += o;
			}
		}
		if(lane == 31)
		{
			warp_sums[wib] = s;
		}
		__syncthreads();
		if(wib == 0)
		{
			unsigned ws = lane < nw ? warp_sums[lane] : 0u;
#pragma unroll
			for(int off = 1; off < 32; off <<= 1)
			{
				const unsigned o = __shfl_up_sync(0xffffffffu, ws, off);
				if(lane >= off)
				{
					ws += o;
				}
			}
			warp_sums[lane] = ws; // inclusive
		}
		__syncthreads();
		const unsigned prefix = (wib > 0 ? warp_sums[wib - 1] : 0u) + s - sum;
		for(int k = 0; k < per; k++)
		{
			const int i = b + k;
			if(i < len)
			{
				buf[i] += prefix;
			}
		}
		__syncthreads();
		for(int i = threadIdx.x; i < len; i += blockDim.x)
		{
			data[i] = buf[i];
		}
		return;
	}
	for(int base = 0; base < len; base += blockDim.x)
	{
		const int i		 = base + threadIdx.x;
		const unsigned v = i < len ? data[i] : 0u;
		unsigned s		 = v;
#pragma unroll
		for(int off = 1; off < 32; off <<= 1)
		{
			const unsigned o = __shfl_up_sync(0xffffffffu, s, off);
			if(lane >= off)
			{
				s += o;
			}
		}
		if(lane == 31)
		{
			warp_sums[wib] = s;
		}
		__syncthreads();
		if(wib == 0)
		{
			unsigned ws = lane < nw ? warp_sums[lane] : 0u;
#pragma unroll
			for(int off = 1; off < 32; off <<= 1)
			{
				const unsigned o = __shfl_up_sync(0xffffffffu, ws, off);
				if(lane >= off)
				{
					ws += o;
				}
			}
			warp_sums[lane] = ws; // inclusive
		}
		__syncthreads();
		const unsigned prefix = carry + (wib > 0 ? warp_sums[wib - 1] : 0u) + s - v;
		if(i < len)
		{
			data[i] = prefix;
		}
		__syncthreads();
		if(threadIdx.x == blockDim.x - 1)
		{
			carry = prefix + v;
		}
		__syncthreads();
	}
}

__global__ void sort_scatter_kernel(const unsigned long long *__restrict__ keys_in, const unsigned *__restrict__ vals_in, int n, int shift,
									const unsigned *__restrict__ offs, int nwarps, unsigned long long *__restrict__ keys_out,
									unsigned *__restrict__ vals_out, int ipw)
{
	__shared__ unsigned run[SORT_WARPS][256]; // running output cursor per digit for this warp
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int warp = blockIdx.x * SORT_WARPS + wib;
	if(warp >= nwarps)
	{
		return;
	}
	for(int d = lane; d < 256; d += 32)
	{
		run[wib][d] = offs[(size_t) d * nwarps + warp];
	}
	__syncwarp();
	const int base			 = warp * ipw;
	const unsigned lt_mask = (1u << lane) - 1u;
	for(int it = 0; it < ipw / 32; it++)
	{
		const int i				   = base + it * 32 + lane;
		const bool valid		   = i < n;
		const unsigned long long k = valid ? keys_in[i] : 0ull;
		const unsigned v		   = valid ? vals_in[i] : 0u;
		const unsigned digit	   = valid ? ((unsigned) (k >> shift) & 255u) : 256u + (unsigned) lane; // invalid lanes never match
		const unsigned peers	   = __match_any_sync(0xffffffffu, digit);
		const unsigned rank		   = __popc(peers & lt_mask);
		unsigned pos			   = 0;
		if(valid)
		{
			pos = run[wib][digit] + rank;
		}
		__syncwarp();
		if(valid && rank == 0)
		{
			run[wib][digit] += __popc(peers);
		}
		__syncwarp();
		if(valid)
		{
			keys_out[pos] = k;
			vals_out[pos] = v;
		}
	}
}

// ---- Karras hierarchy --------------------------------------------------------------------------

__device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j)
{
	if(j < 0 || j >= n)
	{
		return -1;
	}
	const unsigned long long a = keys[i], b = keys[j];
	if(a == b)
	{
		return 64 + __clz(i ^ j);
	}
	return __clzll((long long) (a ^ b));
}

// parent[0 .. n-2] for internal nodes, parent[n-1 .. 2n-2] for leaves; child < 0 means leaf ~idx
__global__ void karras_kernel(const unsigned long long *__restrict__ keys, int n, int2 *__restrict__ children, int *__restrict__ parent)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n - 1)
	{
		return;
	}
	const int d		= (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	const int dmin	= delta(keys, n, i, i - d);
	int lmax		= 2;
	while(delta(keys, n, i, i + lmax * d) > dmin)
	{
		lmax <<= 1;
	}
	int l = 0;
	for(int t = lmax >> 1; t >= 1; t >>= 1)
	{
		if(delta(keys, n, i, i + (l + t) * d) > dmin)
		{
			l += t;
		}
	}
	const int j		= i + l * d;
	const int dnode = delta(keys, n, i, j);
	int s			= 0;
	int t			= l;
	do
	{
		t = (t + 1) >> 1;
		if(delta(keys, n, i, i + (s + t) * d) > dnode)
		{
			s += t;
		}
	} while(t > 1);
	const int gamma = i + s * d + min(d, 0);
	const int lo = min(i, j), hi = max(i, j);
	const int left	= (lo == gamma) ? ~gamma : gamma;
	const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
	children[i]		= make_int2(left, right);
	if(left < 0)
	{
		parent[n - 1 + gamma] = i;
	}
	else
	{
		parent[left] = i;
	}
	if(right < 0)
	{
		parent[n - 1 + gamma + 1] = i;
	}
	else
	{
		parent[right] = i;
	}
	if(i == 0)
	{
		parent[0] = -1;
	}
}

// ---- refit -------------------------------------------------------------------------------------

__global__ void refit_kernel(int n, const unsigned *__restrict__ sorted_ids, const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi,
							 const int2 *__restrict__ children, const int *__restrict__ parent, float4 *node_lo, float4 *node_hi, int *flags,
							 float4 *__restrict__ nodes)
{
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if(leaf >= n)
	{
		return;
	}
	int node = parent[n - 1 + leaf];
	while(node >= 0)
	{
		__threadfence();
		if(atomicAdd(&flags[node], 1) == 0)
		{
			return; // first to arrive: the sibling subtree is not finished yet
		}
		__threadfence();
		const int2 ch = children[node];
		float4 llo, lhi, rlo, rhi;
		if(ch.x < 0)
		{
			const unsigned t = sorted_ids[~ch.x];
			llo = box_lo[t], lhi = box_hi[t];
		}
		else
		{
			llo = __ldcg(node_lo + ch.x), lhi = __ldcg(node_hi + ch.x);
		}
		if(ch.y < 0)
		{
			const unsigned t = sorted_ids[~ch.y];
			rlo = box_lo[t], rhi = box_hi[t];
		}
		else
		{
			rlo = __ldcg(node_lo + ch.y), rhi = __ldcg(node_hi + ch.y);
		}
		// Left/Right interleaved (skr_bvh.cuh)
		nodes[4 * node + 0] = make_float4(llo.x, rlo.x, llo.y, rlo.y);
		nodes[4 * node + 1] = make_float4(llo.z, rlo.z, lhi.x, rhi.x);
		nodes[4 * node + 2] = make_float4(lhi.y, rhi.y, lhi.z, rhi.z);
		// (.z, .w: the largest dead-triangle bound g under the left / right child, see tri_bounds_kernel)
		nodes[4 * node + 3] = make_float4(__int_as_float(ch.x), __int_as_float(ch.y), llo.w, rlo.w);
		__stcg(node_lo + node, make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), fmaxf(llo.w, rlo.w)));
		__stcg(node_hi + node, make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f));
		node = parent[node];
	}
}

__global__ void gather_tris_kernel(const float *__restrict__ tris, const unsigned *__restrict__ sorted_ids, int n, float4 *__restrict__ tri_v)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n)
	{
		return;
	}
	const float *t	 = tris + 9 * (size_t) sorted_ids[i];
	tri_v[4 * i + 0] = make_float4(t[0], t[1], t[2], __int_as_float((int) sorted_ids[i])); // w: original triangle index
	tri_v[4 * i + 1] = make_float4(t[3], t[4], t[5], 0.0f);
	tri_v[4 * i + 2] = make_float4(t[6], t[7], t[8], 0.0f);
	tri_v[4 * i + 3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // (64 B per triangle: 32-byte aligned for 256-bit loads)
}

__global__ void iota_tris_kernel(const float *__restrict__ tris, int n, float4 *__restrict__ tri_v)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n)
	{
		return;
	}
	const float *t	 = tris + 9 * (size_t) i;
	tri_v[4 * i + 0] = make_float4(t[0], t[1], t[2], __int_as_float(i));
	// (.w: the dead-triangle bound -- the brute-force list has no hierarchy to prune it by: tri_any_hit_line skips the test itself)
	tri_v[4 * i + 1] = make_float4(t[3], t[4], t[5], tri_dead_bound(f3(t[0], t[1], t[2]), f3(t[3], t[4], t[5]), f3(t[6], t[7], t[8])));
	tri_v[4 * i + 2] = make_float4(t[6], t[7], t[8], 0.0f);
	tri_v[4 * i + 3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

} // namespace bvhb
