// skr_shaded.cuh -- the opt-in SHADED-TRIANGLES mode (skr_options.shade_triangles, `--shade-triangles`).
//
// NOT reference behaviour: the reference shades every triangle hit black (src/raytrace.h:221-224) and never reads the
// Material each Triangle carries (src/shapes.h:27-33, filled at src/scene.cpp:67-81).  This mode lights triangles with
// those materials.  Its specification is oracle/skr_oracle_ext.inc (brute force on the CPU); semantics in short:
//   closest hit   spheres exactly as the reference (1.0 < t cutoff); triangles by textbook Moller-Trumbore on the ACTUAL
//                 triangle, |det| >= 1e-12, forward hits t > 1e-4; the smaller t wins, the sphere on ties
//   triangle hit  geometric normal flipped against the ray, the triangle's material
//   direct light  ambient + per unshadowed light diffuse + Blinn-Phong specular (src/blinn_phong.h:47-134 without fog)
//   shadows       the reference's sphere test (src/utils.h:42-58) OR any triangle between P + 1e-4 N and the light
// Queries run over a SECOND hierarchy, built over the actual (un-mirrored) triangles the first time a frame asks for it
// (sv.tri_v2 / bvh2 / big_v2): same node layout and builder as the reference query's (skr_bvh_build.cuh).
#pragma once
#include "skr_kernels.cuh"

SKR_DEV bool tri_hit_std(float3 o, float3 dir, float3 v0, float3 v1, float3 v2, float &t)
{
	const float3 e1 = v1 - v0, e2 = v2 - v0;
	const float3 p	= cross(dir, e2);
	const float det = dot(e1, p);
	if(fabsf(det) < 1.0e-12f)
	{
		return false;
	}
	const float inv = __fdiv_rn(1.0f, det);
	const float3 tv = o - v0;
	const float u	= dot(tv, p) * inv;
	if(u < 0.0f || u > 1.0f)
	{
		return false;
	}
	const float3 q = cross(tv, e1);
	const float v  = dot(dir, q) * inv;
	if(v < 0.0f || u + v > 1.0f)
	{
		return false;
	}
	t = dot(e2, q) * inv;
	return t > 1.0e-4f;
}

// RAY (t in (1e-4, tmax)) against the two child boxes of a node; widened by a few ulps like line_hits_boxes
SKR_DEV void ray_hits_boxes(float3 o, float3 inv, float tmax, const float4 &n0, const float4 &n1, const float4 &n2, bool &hl, bool &hr)
{
	const float tx0l = (n0.x - o.x) * inv.x, tx1l = (n1.z - o.x) * inv.x, tx0r = (n0.y - o.x) * inv.x, tx1r = (n1.w - o.x) * inv.x;
	const float ty0l = (n0.z - o.y) * inv.y, ty1l = (n2.x - o.y) * inv.y, ty0r = (n0.w - o.y) * inv.y, ty1r = (n2.y - o.y) * inv.y;
	const float tz0l = (n1.x - o.z) * inv.z, tz1l = (n2.z - o.z) * inv.z, tz0r = (n1.y - o.z) * inv.z, tz1r = (n2.w - o.z) * inv.z;
	float tnl = fmaxf(fmaxf(fminf(tx0l, tx1l), fminf(ty0l, ty1l)), fminf(tz0l, tz1l));
	float tfl = fminf(fminf(fmaxf(tx0l, tx1l), fmaxf(ty0l, ty1l)), fmaxf(tz0l, tz1l));
	float tnr = fmaxf(fmaxf(fminf(tx0r, tx1r), fminf(ty0r, ty1r)), fminf(tz0r, tz1r));
	float tfr = fminf(fminf(fmaxf(tx0r, tx1r), fmaxf(ty0r, ty1r)), fmaxf(tz0r, tz1r));
	tnl -= fabsf(tnl) * 4.8e-7f;
	tfl += fabsf(tfl) * 4.8e-7f;
	tnr -= fabsf(tnr) * 4.8e-7f;
	tfr += fabsf(tfr) * 4.8e-7f;
	hl = tnl <= tfl && tnl < tmax && tfl > 0.0f;
	hr = tnr <= tfr && tnr < tmax && tfr > 0.0f;
}

// Closest (ANY = false) or any (ANY = true) hit of the ray with the actual triangles, t in (1e-4, tbest).
// Returns the ORIGINAL index of the triangle hit (tbest updated) or -1.
template <bool ANY, bool STATS>
SKR_DEV int tri_ray_query(const SceneView &sv, float3 o, float3 d, float &tbest, Counters &cnt)
{
	int hit = -1;
	const auto leaf = [&](const float4 *v) -> bool {
		const float4 a = __ldg(v + 0), b = __ldg(v + 1), c = __ldg(v + 2); // (big_v: 48 B records, tri_v2: 64 B records)
		if(STATS)
		{
			cnt.tt++;
		}
		float t;
		if(tri_hit_std(o, d, f3(a), f3(b), f3(c), t) && t < tbest)
		{
			tbest = t;
			hit	  = __float_as_int(a.w);
			return true;
		}
		return false;
	};
	if(sv.bvh2 == nullptr) // a handful of triangles (or SKR_NO_BVH=1): test them all
	{
		for(int i = 0; i < sv.T; i++)
		{
			if(leaf(sv.tri_v2 + 4 * i) && ANY)
			{
				return hit;
			}
		}
		return hit;
	}
	for(int k = 0; k < sv.nbig2; k++) // outsized triangles, kept out of the hierarchy
	{
		if(leaf(sv.big_v2 + 3 * k) && ANY)
		{
			return hit;
		}
	}
	const float3 inv = f3(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));
	int stack[SKR_BVH_STACK];
	int sp	 = 0;
	int node = 0;
	for(;;)
	{
		float4 n0, n1, n2, n3;
		ldg256(sv.bvh2 + 4 * node, n0, n1);
		ldg256(sv.bvh2 + 4 * node + 2, n2, n3);
		if(STATS)
		{
			cnt.nv++;
		}
		bool hl, hr;
		ray_hits_boxes(o, inv, tbest, n0, n1, n2, hl, hr);
		const int cl = (int) f2u(n3.x), cr = (int) f2u(n3.y);
		int next = -1;
		if(hl)
		{
			if(cl < 0)
			{
				if(leaf(sv.tri_v2 + 4 * (~cl)) && ANY)
				{
					return hit;
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
				if(leaf(sv.tri_v2 + 4 * (~cr)) && ANY)
				{
					return hit;
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
				atomicOr(sv.err, 2);
			}
		}
		if(next < 0)
		{
			if(sp == 0)
			{
				return hit;
			}
			next = stack[--sp];
		}
		node = next;
	}
}

// ambient + diffuse + specular at (p, n) with material (am = ambient (.) ka | power, kd, ks); oracle: ext_direct
template <bool STATS>
SKR_DEV float3 shaded_direct(const float4 *__restrict__ B, const SceneView &sv, bool use_shadows, float4 am, float3 kd, float3 ks, float3 p, float3 n,
							 Counters &cnt)
{
	float3 col			= f3(am);
	const bool has_spec = ks.x != 0.0f || ks.y != 0.0f || ks.z != 0.0f;
	const float3 view	= normalize_fast(sv.cam_pos - p);
	const float3 so		= p + n * 1.0e-4f;
	for(int i = 0; i < sv.L + sv.D; i++)
	{
		const bool point = i < sv.L;
		float3 lhat, lcol;
		float dist = CUDART_INF_F, intensity = 1.0f;
		if(point)
		{
			const float3 lv = f3(B[sv.off_plpos + i]) - p;
			const float d2	= dot(lv, lv);
			dist			= sqrtf(d2);
			lhat			= lv * __fdividef(1.0f, dist);
			intensity		= __fdividef(1.0f, d2);
			lcol			= f3(B[sv.off_plcol + i]);
		}
		else
		{
			lhat = f3(B[sv.off_dldir + (i - sv.L)]);
			lcol = f3(B[sv.off_dlcol + (i - sv.L)]);
		}
		if(use_shadows)
		{
			if(occluded<STATS>(B, sv, p, lhat, cnt)) // the reference's sphere test (counts the shadow ray)
			{
				continue;
			}
			float tb = dist;
			if(sv.T > 0 && tri_ray_query<true, STATS>(sv, so, lhat, tb, cnt) >= 0)
			{
				continue;
			}
		}
		if(STATS)
		{
			cnt.le++;
		}
		col += kd * lcol * (intensity * fmaxf(0.0f, dot(n, lhat)));
		if(has_spec)
		{
			const float3 h = normalize_fast(view + lhat);
			col += ks * lcol * (intensity * pow_fast(fmaxf(0.0f, dot(n, h)), am.w));
		}
	}
	return col;
}

// One thread per pixel of this rank's tiles (same local pixel order as primary_kernel), all samples of the pixel.
template <bool STATS, bool SMEM>
__global__ void __launch_bounds__(SKR_BLOCK) shaded_tris_kernel(const SceneView sv, const FrameParams fp, long long npix)
{
	extern __shared__ float4 smem[];
	const float4 *B = stage_scene<SMEM>(sv, smem);
	Counters cnt;
	zero(cnt);
	const long long lp = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	PixelId p		   = decode_pixel(fp, lp);
	p.valid			   = p.valid && lp < npix;
	const uint32_t pixel = (uint32_t) (p.y * fp.width + p.x);
	const int nsamples	 = fp.max_depth > 0 ? fp.spp : 0;
	float3 sum			 = f3(0.0f, 0.0f, 0.0f);
	uint4 jit			 = make_uint4(0u, 0u, 0u, 0u);
	for(int s = 0; s < nsamples && p.valid; s++)
	{
		float u, v;
		if(fp.grid > 0) // src/main.cpp:52-54, as primary_kernel
		{
			if((s & 3) == 0)
			{
				jit = philox4x32_10(make_uint4(pixel, (uint32_t) s >> 2, 0u, 0u), fp.key);
			}
			const uint32_t jw = (s & 3) == 0 ? jit.x : (s & 3) == 1 ? jit.y : (s & 3) == 2 ? jit.z : jit.w;
			const float r	  = rng_unit(jw);
			u = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(2.0f, __fmul_rn(__fadd_rn((float) p.x, r), fp.inv_w)), 1.0f), fp.angle), fp.aspect);
			v = __fmul_rn(__fsub_rn(1.0f, __fmul_rn(2.0f, __fmul_rn(__fadd_rn((float) p.y, r), fp.inv_h))), fp.angle);
		}
		else // src/main.cpp:73-74
		{
			u = (float) __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(2.0, __dmul_rn((double) p.x + 0.5, (double) fp.inv_w)), 1.0), (double) fp.angle),
								  (double) fp.aspect);
			v = (float) __dmul_rn(__dsub_rn(1.0, __dmul_rn(2.0, __dmul_rn((double) p.y + 0.5, (double) fp.inv_h))), (double) fp.angle);
		}
		const float3 d = add_rn(add_rn(sv.cam_dir, muls_rn(sv.cam_right, u)), muls_rn(sv.cam_up, v));
		const float3 o = sv.cam_pos;
		if(STATS)
		{
			cnt.ch++;
		}
		float ts	 = CUDART_INF_F;
		const int si = closest_sphere<true, STATS>(B, sv, o, d, ts, cnt);
		float tt	 = CUDART_INF_F;
		const int ti = sv.T > 0 ? tri_ray_query<false, STATS>(sv, o, d, tt, cnt) : -1;
		if(si < 0 && ti < 0)
		{
			sum += sv.background;
		}
		else if(ti >= 0 && (si < 0 || tt < ts))
		{
			// the triangle's vertices by ORIGINAL index are not kept; re-derive the normal from the hit triangle's record:
			// tri_ray_query returns the original index, the geometry comes from the raw upload (9 floats per triangle)
			const float *raw = sv.tris_raw + 9 * (size_t) ti;
			const float3 v0 = f3(raw[0], raw[1], raw[2]), v1 = f3(raw[3], raw[4], raw[5]), v2 = f3(raw[6], raw[7], raw[8]);
			const float3 hp = o + d * tt;
			float3 n		= normalize_fast(cross(v1 - v0, v2 - v0));
			if(dot(n, d) > 0.0f)
			{
				n = n * -1.0f;
			}
			float4 am = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
			float3 kd = f3(0.0f, 0.0f, 0.0f), ks = kd;
			if(sv.tri_mat)
			{
				am = __ldg(sv.tri_mat + 3 * ti);
				kd = f3(__ldg(sv.tri_mat + 3 * ti + 1));
				ks = f3(__ldg(sv.tri_mat + 3 * ti + 2));
			}
			sum += shaded_direct<STATS>(B, sv, fp.shadows != 0, am, kd, ks, hp, n, cnt);
		}
		else
		{
			if(STATS)
			{
				cnt.hits++;
			}
			const float3 c	= f3(B[sv.off_geom + si]);
			const float t	= sphere_t_ref(o, d, c, B[sv.off_spec + si].w, ts);
			const float3 hp = add_rn(o, muls_rn(d, t));
			const float3 n	= normalize_fast(sub_rn(hp, c));
			sum += shaded_direct<STATS>(B, sv, fp.shadows != 0, B[sv.off_amb + si], f3(B[sv.off_diff + si]), f3(B[sv.off_spec + si]), hp, n, cnt);
		}
	}
	if(p.valid)
	{
		if(fp.grid > 0)
		{
			const float n2 = (float) fp.spp;
			sum			   = f3(__fdiv_rn(sum.x, n2), __fdiv_rn(sum.y, n2), __fdiv_rn(sum.z, n2));
		}
		write_pixel(fp, lp, p, sum);
	}
	flush_counters<STATS>(fp, cnt);
}
