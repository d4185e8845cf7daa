// skr_bvh.cuh -- LBVH traversal for the triangle half of shade() (reference src/raytrace.h:169-186).
//
// What the reference computes: "does ANY triangle pass triangle_intersection_occurs with t < t_sphere_min?"
// (then the pixel sample is black, src/raytrace.h:221-224).  Three quirks shape the query:
//   * u is negated (src/utils.h:198)  -> the accepted region is the MIRRORED triangle (v0, 2*v0 - v1, v2);
//     the BVH bounds those, the leaf test runs the reference arithmetic on the original vertices;
//   * there is no sign test on t (src/utils.h:211-212) and the caller accepts any t < min_distance
//     -> the query is over the whole LINE t in (-inf, tmax), not a ray;
//   * any accepted hit gives the same answer -> any-hit: stop at the first one.
//
// Node layout (built by skr_bvh_build.cuh), 4 x float4 = 64 B per internal node, children boxes in the parent:
//   n0 = (Lmin.x, Lmin.y, Lmin.z, Lmax.x)  n1 = (Lmax.y, Lmax.z, Rmin.x, Rmin.y)
//   n2 = (Rmin.z, Rmax.x, Rmax.y, Rmax.z)  n3 = (bits(left), bits(right), -, -)
// child index >= 0: internal node; < 0: leaf ~idx, whose triangle is tri_v[3*idx .. 3*idx+2] (leaf order).
#pragma once

#define SKR_BVH_STACK 96

SKR_DEV bool line_hits_box(float3 o, float3 inv, float tmax, float bx0, float by0, float bz0, float bx1, float by1, float bz1)
{
	// slab test without the t >= 0 clamp; fminf/fmaxf drop NaNs (0 * inf when the origin sits on a slab plane
	// of a zero direction component), which only ever widens the interval
	const float tx0 = (bx0 - o.x) * inv.x, tx1 = (bx1 - o.x) * inv.x;
	const float ty0 = (by0 - o.y) * inv.y, ty1 = (by1 - o.y) * inv.y;
	const float tz0 = (bz0 - o.z) * inv.z, tz1 = (bz1 - o.z) * inv.z;
	float tn		= fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fminf(tz0, tz1));
	float tf		= fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
	// widen by a few ulps so that rounding in the slab arithmetic can never cull a box the exact test would keep
	tn -= fabsf(tn) * 4.8e-7f;
	tf += fabsf(tf) * 4.8e-7f;
	return tn <= tf && tn < tmax;
}

template <bool STATS>
SKR_DEV bool tri_leaf_hit(const SceneView &sv, int leaf, float3 o, float3 d, float tmax, Counters &cnt)
{
	const float4 a = __ldg(sv.tri_v + 3 * leaf + 0);
	const float4 b = __ldg(sv.tri_v + 3 * leaf + 1);
	const float4 c = __ldg(sv.tri_v + 3 * leaf + 2);
	if(STATS)
	{
		cnt.tt++;
	}
	float t;
	return tri_test_ref(o, d, f3(a), f3(b), f3(c), t) && t < tmax;
}

template <bool STATS>
SKR_DEV bool tri_any_hit_line(const SceneView &sv, float3 o, float3 d, float tmax, Counters &cnt)
{
	if(sv.bvh == nullptr) // brute force (validation mode, SKR_NO_BVH=1)
	{
		for(int i = 0; i < sv.T; i++)
		{
			if(tri_leaf_hit<STATS>(sv, i, o, d, tmax, cnt))
			{
				return true;
			}
		}
		return false;
	}
	if(sv.bvh_root_is_leaf)
	{
		return tri_leaf_hit<STATS>(sv, 0, o, d, tmax, cnt);
	}
	const float3 inv = f3(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));
	int stack[SKR_BVH_STACK];
	int sp	 = 0;
	int node = 0;
	for(;;)
	{
		const float4 n0 = __ldg(sv.bvh + 4 * node + 0);
		const float4 n1 = __ldg(sv.bvh + 4 * node + 1);
		const float4 n2 = __ldg(sv.bvh + 4 * node + 2);
		const float4 n3 = __ldg(sv.bvh + 4 * node + 3);
		if(STATS)
		{
			cnt.nv++;
		}
		const bool hl = line_hits_box(o, inv, tmax, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
		const bool hr = line_hits_box(o, inv, tmax, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
		const int cl = (int) f2u(n3.x), cr = (int) f2u(n3.y);
		int next = -1; // next internal node to visit, if any
		if(hl)
		{
			if(cl < 0)
			{
				if(tri_leaf_hit<STATS>(sv, ~cl, o, d, tmax, cnt))
				{
					return true;
				}
			}
			else
			{
				next = cl;
			}
		}
		if(hr)
		{
			if(cr < 0)
			{
				if(tri_leaf_hit<STATS>(sv, ~cr, o, d, tmax, cnt))
				{
					return true;
				}
			}
			else if(next < 0)
			{
				next = cr;
			}
			else if(sp < SKR_BVH_STACK)
			{
				stack[sp++] = cr;
			}
			else
			{
				atomicOr(sv.err, 2); // deeper than any LBVH over 63-bit codes + index bits can be; reported, never silent
			}
		}
		if(next < 0)
		{
			if(sp == 0)
			{
				return false;
			}
			next = stack[--sp];
		}
		node = next;
	}
}
