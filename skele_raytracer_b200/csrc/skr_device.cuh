// skr_device.cuh -- device-side scene view, Philox RNG, intersection and shading routines.
//
// Reference semantics reproduced here (paths under /root/reference):
//   sphere test            src/utils.h:87-121,169-179   (smallest_root / collision_distance / intersection_occurs)
//   triangle test          src/utils.h:181-213          (mirrored-u Moller-Trumbore, no sign test on t)
//   shadow()               src/utils.h:42-76
//   Blinn-Phong terms      src/blinn_phong.h:13-134     (ambient / diffuse / specular / spherical fog)
//   basis + hemisphere     src/utils.h:148-165, src/raytrace.h:22-30
#pragma once
#include "skr_math.cuh"

// ------------------------------------------------------------------------------------------------
// Scene view.  All small per-scene arrays live in ONE float4 blob (built by skr_scene_upload) that each
// CTA stages into shared memory; indices below are float4 offsets into it.
//   geom [S]  (cx, cy, cz, r)
//   prim [S]  (2*(cam-c).xyz, |cam-c|^2 - r^2)      the camera-origin constants of the quadratic
//   amb  [S]  (ambient_light (.) ka .xyz, phong power)
//   diff [S]  (kd.xyz, ior)
//   spec [S]  (ks.xyz, any(ks != 0))
//   plpos[L]  (pos.xyz, 0)   plcol[L] (colour.xyz, 0)
//   dldir[D]  (normalize(dir).xyz, 0)   dlcol[D] (colour.xyz, 0)
//   foga [F]  (scattering, absorption, radius, 0)   fogalb[F] (albedo.xyz, 0)
//   fogp      float[S*L*F]: exp(-min(|c_s - Lp_i|, 2 r_j) * (abs_j + scat_j)), precomputed on the host with the
//             reference's own expression (src/blinn_phong.h:22-29)
// ------------------------------------------------------------------------------------------------
struct SceneView
{
	int S, T, L, D, F;
	int off_geom, off_prim, off_amb, off_diff, off_spec, off_plpos, off_plcol, off_dldir, off_dlcol, off_foga, off_fogalb, off_fogp;
	int blob_f4;		 // blob size in float4
	int blob_in_smem;	 // 1: kernels stage the blob in shared memory
	const float4 *blob;	 // device
	const float4 *tri_v; // 3 float4 per triangle (v0, v1, v2), LBVH leaf order
	const float4 *bvh;	 // 4 float4 per internal node (see skr_bvh.cuh)
	int bvh_root_is_leaf; // T == 1
	float3 cam_pos, cam_dir, cam_up, cam_right, background;
};

struct Counters // per-thread, reduced at kernel exit when STATS
{
	unsigned ch, sh, st, stp, tt, nv, hits, le;
};
SKR_DEV void zero(Counters &c) { c.ch = c.sh = c.st = c.stp = c.tt = c.nv = c.hits = c.le = 0; }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Keying, identical to oracle/skr_oracle.c:
//   counter = (pixel, sample, node, slot), key = (seed_lo, seed_hi)
//   slot 0                      : .x jitter r of the primary sample
//   slot 1 + (call*L + i)*F + j : fog draws (xi, dx, dy, dz) for call (0 diffuse / 1 specular), light i, fog j
//   slot 1 + 2*L*F + c          : (.x, .y) = (r1, r2) of GI child c
// ------------------------------------------------------------------------------------------------
SKR_DEV uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
	for(int r = 0; r < 10; r++)
	{
		uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
		uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
		c			 = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
		k.x += 0x9E3779B9u;
		k.y += 0xBB67AE85u;
	}
	return c;
}
// float(rand()) / float(RAND_MAX) with a 31-bit draw: k * 2^-31 (RAND_MAX rounds to 2^31 as a float)
SKR_DEV float rng_unit(uint32_t x) { return __fmul_rn(__int2float_rn((int) (x >> 1)), 4.6566128730773926e-10f); }
// -1.0f + rand() / float(RAND_MAX / 2)    (src/utils.h:219-221)
SKR_DEV float rng_pm1(uint32_t x) { return __fadd_rn(-1.0f, __fmul_rn(__int2float_rn((int) (x >> 1)), 9.3132257461547852e-10f)); }

struct RngCtx
{
	uint32_t pixel, sample, node;
	uint2 key;
};
SKR_DEV uint4 rng_block(const RngCtx &r, uint32_t slot) { return philox4x32_10(make_uint4(r.pixel, r.sample, r.node, slot), r.key); }

// ------------------------------------------------------------------------------------------------
// Spheres
// ------------------------------------------------------------------------------------------------

// Closest sphere along (o, d) with 1.0 < t < inf, strict minimum, first wins ties (src/raytrace.h:149-165).
// PRIMARY: o is the camera position, use the precomputed (2e, c) constants.
template <bool PRIMARY, bool STATS>
SKR_DEV int closest_sphere(const float4 *__restrict__ B, const SceneView &sv, float3 o, float3 d, float &tmin, Counters &cnt)
{
	const float a  = dot(d, d);
	const float a4 = 4.0f * a;
	const float a2 = 2.0f * a;
	int best	   = -1;
	tmin		   = CUDART_INF_F;
	const int S	   = sv.S;
#pragma unroll 4
	for(int s = 0; s < S; s++)
	{
		float b, cc;
		if(PRIMARY)
		{
			const float4 p = B[sv.off_prim + s];
			b			   = dot(d, f3(p)); // = 2*dot(d, e): scaling by 2 is exact
			cc			   = p.w;
		}
		else
		{
			const float4 g = B[sv.off_geom + s];
			const float3 e = o - f3(g);
			b			   = 2.0f * dot(d, e);
			cc			   = dot(e, e) - g.w * g.w;
		}
		const float disc = b * b - a4 * cc;
		if(STATS)
		{
			cnt.st++;
		}
		if(disc >= 0.0f)
		{
			if(STATS)
			{
				cnt.stp++;
			}
			const float t2 = __fdiv_rn(-b - __fsqrt_rn(disc), a2); // the root smallest_root returns for a > 0
			if(t2 > 1.0f && t2 < tmin)
			{
				tmin = t2;
				best = s;
			}
		}
	}
	return best;
}

// shadow(): ANY sphere with 1.0 < t2 < inf along the normalised direction from p + 1e-6 occludes; no light-distance
// bound, hits within 1.0 ignored (src/utils.h:42-58).  t2 > 1  <=>  q = -b - 2a > 0 and q*q > disc.
template <bool STATS>
SKR_DEV bool occluded(const float4 *__restrict__ B, const SceneView &sv, float3 p, float3 dir, Counters &cnt)
{
	const float3 o = adds_rn(p, 0.000001f);
	const float a  = dot(dir, dir);
	const float a4 = 4.0f * a;
	const float a2 = 2.0f * a;
	const int S	   = sv.S;
	if(STATS)
	{
		cnt.sh++;
	}
	for(int s = 0; s < S; s++)
	{
		const float4 g	 = B[sv.off_geom + s];
		const float3 e	 = o - f3(g);
		const float b	 = 2.0f * dot(dir, e);
		const float cc	 = dot(e, e) - g.w * g.w;
		const float disc = b * b - a4 * cc;
		const float q	 = -b - a2;
		if(STATS)
		{
			cnt.st++;
			cnt.stp += disc >= 0.0f;
		}
		if(disc >= 0.0f && q > 0.0f && q * q > disc)
		{
			return true;
		}
	}
	return false;
}

// ------------------------------------------------------------------------------------------------
// Triangles.  The leaf test is the reference's arithmetic verbatim in uncontracted float ops, so that the
// `fabs(det) < 1e-5` rejection and the u/v window decide exactly as on the CPU.
// ------------------------------------------------------------------------------------------------
SKR_DEV bool tri_test_ref(float3 o, float3 dir, float3 v0, float3 v1, float3 v2, float &t)
{
	const float3 v0v1 = sub_rn(v1, v0);
	const float3 v0v2 = sub_rn(v2, v0);
	const float3 p	  = cross_rn(dir, v0v2);
	const float det	  = dot_rn(v0v1, p);
	if(fabsf(det) < 0.00001f)
	{
		return false;
	}
	const float inv = __fdiv_rn(1.0f, det);
	const float3 tv = sub_rn(o, v0);
	const float u	= __fmul_rn(inv, -dot_rn(tv, p)); // dot(-tv, p) == -dot(tv, p) under round-to-nearest
	if(u < 0.0f || u > 1.0f)
	{
		return false;
	}
	const float3 q = cross_rn(tv, v0v1);
	const float v  = __fmul_rn(dot_rn(dir, q), inv);
	if(v < 0.0f || __fadd_rn(u, v) > 1.0f)
	{
		return false;
	}
	t = __fmul_rn(dot_rn(v0v2, q), inv);
	return true;
}

#include "skr_bvh.cuh" // tri_any_hit_line()

// ------------------------------------------------------------------------------------------------
// Shading
// ------------------------------------------------------------------------------------------------

SKR_DEV float pow_fast(float x, float p) // x >= 0
{
	return p == 0.0f ? 1.0f : __powf(x, p);
}

// bp::spherical_fog_shading (src/blinn_phong.h:19-44) + scattering_phase_function (src/utils.h:216-224)
SKR_DEV float3 fog_term(const float4 *__restrict__ B, const SceneView &sv, const RngCtx &rng, int call, int i, int j, int sidx, float3 kd, float3 lcol,
						float3 lhat, float inv_d2, float3 n)
{
	const float *fogp = reinterpret_cast<const float *>(B + sv.off_fogp);
	const float p_no  = fogp[(sidx * sv.L + i) * sv.F + j];
	const uint4 r	  = rng_block(rng, 1u + (uint32_t) ((call * sv.L + i) * sv.F + j));
	if(rng_unit(r.x) > p_no)
	{
		return kd * lcol * (inv_d2 * fmaxf(0.0f, dot(n, lhat)));
	}
	const float4 fa = B[sv.off_foga + j];
	const float3 nd = f3(lhat.x + rng_pm1(r.y) * fa.x, lhat.y + rng_pm1(r.z) * fa.x, lhat.z + rng_pm1(r.w) * fa.x);
	return f3(B[sv.off_fogalb + j]) * lcol * fmaxf(0.0f, dot(n, nd));
}

// direct_illumination as HEAD computes it (src/raytrace.h:36-44): ambient + diffuse + specular.  One shadow ray per
// light serves both terms (the reference casts the same ray twice).  View direction is towards the CAMERA POSITION
// even for bounce hits (src/blinn_phong.h:93).
template <bool STATS>
SKR_DEV float3 direct_light(const float4 *__restrict__ B, const SceneView &sv, bool use_shadows, const RngCtx &rng, int sidx, float3 p, float3 n,
							Counters &cnt)
{
	const float4 am = B[sv.off_amb + sidx];
	const float3 kd = f3(B[sv.off_diff + sidx]);
	const float4 sp = B[sv.off_spec + sidx];
	const float3 ks = f3(sp);
	float3 col		= f3(am);
	const float3 view = normalize_fast(sv.cam_pos - p);
	for(int i = 0; i < sv.L; i++)
	{
		const float3 lv	 = f3(B[sv.off_plpos + i]) - p;
		const float d2	 = dot(lv, lv);
		const float3 lhat = normalize_rn(sub_rn(f3(B[sv.off_plpos + i]), p)); // also the shadow-ray direction
		if(use_shadows && occluded<STATS>(B, sv, p, lhat, cnt))
		{
			continue;
		}
		if(STATS)
		{
			cnt.le++;
		}
		const float3 lcol  = f3(B[sv.off_plcol + i]);
		const float inv_d2 = __fdividef(1.0f, d2);
		if(sv.F > 0)
		{
			for(int j = 0; j < sv.F; j++)
			{
				col += fog_term(B, sv, rng, 0, i, j, sidx, kd, lcol, lhat, inv_d2, n);
				col += fog_term(B, sv, rng, 1, i, j, sidx, kd, lcol, lhat, inv_d2, n);
			}
		}
		else
		{
			col += kd * lcol * (inv_d2 * fmaxf(0.0f, dot(n, lhat)));
			if(sp.w != 0.0f)
			{
				const float3 h = normalize_fast(view + lhat);
				col += ks * lcol * (inv_d2 * pow_fast(fmaxf(0.0f, dot(n, h)), am.w));
			}
		}
	}
	for(int i = 0; i < sv.D; i++)
	{
		const float3 lhat = f3(B[sv.off_dldir + i]);
		if(use_shadows && occluded<STATS>(B, sv, p, lhat, cnt))
		{
			continue;
		}
		const float3 lcol = f3(B[sv.off_dlcol + i]);
		col += kd * lcol * fmaxf(0.0f, dot(n, lhat));
		if(sp.w != 0.0f)
		{
			const float3 h = normalize_fast(view + lhat);
			col += ks * lcol * pow_fast(fmaxf(0.0f, dot(n, h)), am.w);
		}
	}
	return col;
}

// transform_coordinate_space (src/utils.h:148-165)
SKR_DEV void basis_from_normal(float3 n, float3 &nt, float3 &nb)
{
	if(fabsf(n.x) > fabsf(n.y))
	{
		const float l = __fsqrt_rn(__fadd_rn(__fmul_rn(n.x, n.x), __fmul_rn(n.z, n.z)));
		nt			  = f3(__fdiv_rn(n.z, l), __fdiv_rn(0.0f, l), __fdiv_rn(-n.x, l));
	}
	else
	{
		const float l = __fsqrt_rn(__fadd_rn(__fmul_rn(n.y, n.y), __fmul_rn(n.z, n.z)));
		nt			  = f3(__fdiv_rn(0.0f, l), __fdiv_rn(-n.z, l), __fdiv_rn(n.y, l));
	}
	nb = cross_rn(n, nt);
}

// GI child direction: uniform_sample_hemi (src/raytrace.h:22-30) pushed through the reference's local->world
// transform INCLUDING its bug (perp_to_both.y/.z used for the z-column, src/raytrace.h:123-125).
SKR_DEV float3 gi_child_dir(float r1, float r2, float3 n, float3 nt, float3 nb)
{
	const float s_theta = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(r1, r1)));
	float sn, cs;
	sincospif(2.0f * r2, &sn, &cs); // phi = 2*pi*r2
	const float sx = s_theta * cs, sy = r1, sz = s_theta * sn;
	return f3(sx * nb.x + sy * n.x + sz * nt.x, sx * nb.y + sy * n.y + sz * nb.y, sx * nb.z + sy * n.z + sz * nb.z);
}

// ------------------------------------------------------------------------------------------------
// One closest-hit query = the first half of shade() (src/raytrace.h:149-192).
// Returns: -2 background, -1 triangle (black), >= 0 sphere index with t in tmin.
// ------------------------------------------------------------------------------------------------
template <bool PRIMARY, bool STATS>
SKR_DEV int closest_hit(const float4 *__restrict__ B, const SceneView &sv, float3 o, float3 d, float &tmin, Counters &cnt)
{
	if(STATS)
	{
		cnt.ch++;
	}
	const int s = closest_sphere<PRIMARY, STATS>(B, sv, o, d, tmin, cnt);
	if(sv.T > 0 && tri_any_hit_line<STATS>(sv, o, d, tmin, cnt))
	{
		return -1;
	}
	if(s < 0)
	{
		return -2;
	}
	if(STATS)
	{
		cnt.hits++;
	}
	return s;
}
