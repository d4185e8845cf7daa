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
// Node layout (built by skr_bvh_build.cuh), 4 x float4 = 64 B per internal node, both child boxes in the parent,
// interleaved Left/Right so that the slab arithmetic of the two boxes runs as packed FP32x2 (FADD2/FMUL2):
//   n0 = (Lmin.x, Rmin.x, Lmin.y, Rmin.y)  n1 = (Lmin.z, Rmin.z, Lmax.x, Rmax.x)
//   n2 = (Lmax.y, Rmax.y, Lmax.z, Rmax.z)  n3 = (bits(left), bits(right), gL, gR)
// gL / gR: the largest dead-triangle bound under the child (skr_bvh_build.cuh: tri_bounds_kernel): a ray with
// |dir| * g < 1e-5 cannot pass the reference's fabs(det) >= 1e-5 test on any triangle there, so the child is skipped.
// child index >= 0: internal node; < 0: leaf ~idx, whose triangle is tri_v[4*idx .. 4*idx+2] (leaf order, 64 B each).
#pragma once

#define SKR_BVH_STACK 96

// Slab test of the LINE (no t >= 0 clamp) against the two child boxes of a node at once.  (b - o) * inv per plane, the
// same two roundings as the scalar form, issued as FADD2 + FMUL2 with Left in .x and Right in .y.  fminf/fmaxf drop
// NaNs (0 * inf when the origin sits on a slab plane of a zero direction component), which only widens the interval;
// the interval is then widened by a few ulps so that rounding can never cull a box the exact test would keep.
SKR_DEV void line_hits_boxes(float3 o, float3 inv, float tmax, const float4 &n0, const float4 &n1, const float4 &n2, bool &hl, bool &hr)
{
	const float2 nox = splat2(-o.x), noy = splat2(-o.y), noz = splat2(-o.z);
	const float2 ix = splat2(inv.x), iy = splat2(inv.y), iz = splat2(inv.z);
	const float2 tx0 = mul2(add2(f2(n0.x, n0.y), nox), ix), tx1 = mul2(add2(f2(n1.z, n1.w), nox), ix);
	const float2 ty0 = mul2(add2(f2(n0.z, n0.w), noy), iy), ty1 = mul2(add2(f2(n2.x, n2.y), noy), iy);
	const float2 tz0 = mul2(add2(f2(n1.x, n1.y), noz), iz), tz1 = mul2(add2(f2(n2.z, n2.w), noz), iz);
	float tnl = fmaxf(fmaxf(fminf(tx0.x, tx1.x), fminf(ty0.x, ty1.x)), fminf(tz0.x, tz1.x));
	float tfl = fminf(fminf(fmaxf(tx0.x, tx1.x), fmaxf(ty0.x, ty1.x)), fmaxf(tz0.x, tz1.x));
	float tnr = fmaxf(fmaxf(fminf(tx0.y, tx1.y), fminf(ty0.y, ty1.y)), fminf(tz0.y, tz1.y));
	float tfr = fminf(fminf(fmaxf(tx0.y, tx1.y), fmaxf(ty0.y, ty1.y)), fmaxf(tz0.y, tz1.y));
	tnl -= fabsf(tnl) * 4.8e-7f;
	tfl += fabsf(tfl) * 4.8e-7f;
	tnr -= fabsf(tnr) * 4.8e-7f;
	tfr += fabsf(tfr) * 4.8e-7f;
	hl = tnl <= tfl && tnl < tmax;
	hr = tnr <= tfr && tnr < tmax;
}

template <bool STATS>
SKR_DEV bool tri_leaf_hit(const SceneView &sv, int leaf, float3 o, float3 d, float tmax, Counters &cnt)
{
	float4 a, b;
	ldg256(sv.tri_v + 4 * leaf, a, b); // 64 B per triangle (v0, v1, v2, -): one 256-bit + one 128-bit load
	const float4 c = __ldg(sv.tri_v + 4 * leaf + 2);
	if(STATS)
	{
		cnt.tt++;
	}
	float t;
	return tri_test_ref(o, d, f3(a), f3(b), f3(c), t) && t < tmax;
}

// The outsized triangles kept out of the hierarchy: any hit ends the query.
template <bool STATS>
SKR_DEV bool tri_big_hit(const SceneView &sv, float3 o, float3 d, float tmax, Counters &cnt)
{
	for(int k = 0; k < sv.nbig; k++)
	{
		const float4 a = __ldg(sv.big_v + 3 * k + 0), b = __ldg(sv.big_v + 3 * k + 1), c = __ldg(sv.big_v + 3 * k + 2);
		if(STATS)
		{
			cnt.tt++;
		}
		float t;
		if(tri_test_ref(o, d, f3(a), f3(b), f3(c), t) && t < tmax)
		{
			return true;
		}
	}
	return false;
}

// Traversal state of one line query, so that the loop can be driven one node at a time (tri_deferred_kernel refills idle
// lanes between steps) as well as to completion (tri_any_hit_line).
struct TriWalk
{
	float3 o, d, inv;
	float tmax;
	float dlen; // |d| rounded up: the child's dead-triangle bound g is tested against 1e-5 / dlen
	int node, sp;
};
SKR_DEV void tri_walk_begin(TriWalk &w, float3 o, float3 d, float tmax)
{
	w.o	   = o;
	w.d	   = d;
	w.inv  = f3(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));
	w.tmax = tmax;
	w.dlen = sqrtf(dot(d, d)) * 1.00001f;
	w.node = 0;
	w.sp   = 0;
}
// one node visit.  Returns 1: a triangle was hit (query over), 0: no node left (query over, no hit), -1: keep going.
template <bool STATS>
SKR_DEV int tri_walk_step(const SceneView &sv, TriWalk &w, int *stack, Counters &cnt)
{
	float4 n0, n1, n2, n3;
	ldg256(sv.bvh + 4 * w.node, n0, n1); // the 64 B node in two 256-bit loads
	ldg256(sv.bvh + 4 * w.node + 2, n2, n3);
	if(STATS)
	{
		cnt.nv++;
	}
	bool hl, hr;
	line_hits_boxes(w.o, w.inv, w.tmax, n0, n1, n2, hl, hr);
	hl = hl && n3.z * w.dlen >= 0.99999e-5f; // dead subtrees: no triangle there can reach fabs(det) >= 1e-5 for this ray
	hr = hr && n3.w * w.dlen >= 0.99999e-5f;
	const int cl = (int) f2u(n3.x), cr = (int) f2u(n3.y);
	int next = -1; // next internal node to visit, if any
	if(hl)
	{
		if(cl < 0)
		{
			if(tri_leaf_hit<STATS>(sv, ~cl, w.o, w.d, w.tmax, cnt))
			{
				return 1;
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
			if(tri_leaf_hit<STATS>(sv, ~cr, w.o, w.d, w.tmax, cnt))
			{
				return 1;
			}
		}
		else if(next < 0)
		{
			next = cr;
		}
		else if(w.sp < SKR_BVH_STACK)
		{
			stack[w.sp++] = cr;
		}
		else
		{
			atomicOr(sv.err, 2); // deeper than any LBVH over 63-bit codes + index bits can be; reported, never silent
		}
	}
	if(next < 0)
	{
		if(w.sp == 0)
		{
			return 0;
		}
		next = stack[--w.sp];
	}
	w.node = next;
	return -1;
}

// One node for a TEAM walk (tri_deferred_kernel): tests both children of `node`; leaf children are tested at once (true =
// a triangle was hit), internal children that the line reaches are returned in kid[0 .. nk).
template <bool STATS>
SKR_DEV bool tri_node_visit(const SceneView &sv, const TriWalk &w, int node, int (&kid)[2], int &nk, Counters &cnt)
{
	float4 n0, n1, n2, n3;
	ldg256(sv.bvh + 4 * node, n0, n1);
	ldg256(sv.bvh + 4 * node + 2, n2, n3);
	if(STATS)
	{
		cnt.nv++;
	}
	bool hl, hr;
	line_hits_boxes(w.o, w.inv, w.tmax, n0, n1, n2, hl, hr);
	hl = hl && n3.z * w.dlen >= 0.99999e-5f;
	hr = hr && n3.w * w.dlen >= 0.99999e-5f;
	const int cl = (int) f2u(n3.x), cr = (int) f2u(n3.y);
	if(hl && cl < 0 && tri_leaf_hit<STATS>(sv, ~cl, w.o, w.d, w.tmax, cnt))
	{
		nk = 0;
		return true;
	}
	if(hr && cr < 0 && tri_leaf_hit<STATS>(sv, ~cr, w.o, w.d, w.tmax, cnt))
	{
		nk = 0;
		return true;
	}
	// internal children the line reaches (no indexed stores: kid[] stays in registers)
	const bool li = hl && cl >= 0, ri = hr && cr >= 0;
	kid[0] = li ? cl : cr;
	kid[1] = cr;
	nk	   = (li ? 1 : 0) + (ri ? 1 : 0);
	return false;
}

template <bool STATS>
SKR_DEV bool tri_any_hit_line(const SceneView &sv, float3 o, float3 d, float tmax, Counters &cnt)
{
	if(sv.bvh == nullptr) // brute force (a handful of triangles, or validation mode SKR_NO_BVH=1)
	{
		// triangles that are dead for this ray (|d| g < 1e-5: the reference's fabs(det) test must fail, skr_bvh_build.cuh) are
		// skipped without the test -- spheres1.scn's two triangles are collinear points, dead for every camera ray
		const float dlen = sqrtf(dot(d, d)) * 1.00001f;
		for(int i = 0; i < sv.T; i++)
		{
			if(__ldg(sv.tri_v + 4 * i + 1).w * dlen < 0.99999e-5f)
			{
				continue;
			}
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
	if(tri_big_hit<STATS>(sv, o, d, tmax, cnt)) // outsized triangles first
	{
		return true;
	}
	TriWalk w;
	tri_walk_begin(w, o, d, tmax);
	int stack[SKR_BVH_STACK];
	for(;;)
	{
		const int r = tri_walk_step<STATS>(sv, w, stack, cnt);
		if(r >= 0)
		{
			return r == 1;
		}
	}
}

// DEFERRED query (single-sample camera rays of non --gillum frames, see tri_deferred_kernel): walk at most SKR_DEFER_STEPS
// nodes in place -- enough for the many lines that only graze the model's bounds -- and answer 1: a triangle is hit,
// 0: none is, 2: undecided.  Undecided rays become candidates: tri_deferred_kernel walks them (from the root again) as
// one dense work list.
#ifndef SKR_DEFER_STEPS
#define SKR_DEFER_STEPS 6
#endif
template <bool STATS>
SKR_DEV int tri_any_hit_line_deferred(const SceneView &sv, float3 o, float3 d, float tmax, Counters &cnt)
{
	if(tri_big_hit<STATS>(sv, o, d, tmax, cnt))
	{
		return 1;
	}
	TriWalk w;
	tri_walk_begin(w, o, d, tmax);
	int stack[SKR_DEFER_STEPS + 1];
	Counters c2; // the prefix of an undecided walk is walked again by tri_deferred_kernel: counted there, not here
	zero(c2);
#pragma unroll 1
	for(int k = 0; k < SKR_DEFER_STEPS; k++)
	{
		const int r = tri_walk_step<STATS>(sv, w, stack, c2);
		if(r >= 0)
		{
			if(STATS)
			{
				cnt.nv += c2.nv;
				cnt.tt += c2.tt;
			}
			return r;
		}
	}
	return 2;
}
