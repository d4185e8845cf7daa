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
//   geom [S]  (cx, cy, cz, -r*r)
//   pgeom[S4/2][2]  spheres in PAIRS for the packed FP32x2 test loops: (-cx0,-cx1,-cy0,-cy1), (-cz0,-cz1,-r0^2,-r1^2)
//   pprim[S4/2][2]  same for camera rays: (ex0,ex1,ey0,ey1), (ez0,ez1,cc0,cc1) with e = cam - c, cc = e.e - r^2
//                   S4 = S rounded up to a multiple of 4; the padding spheres can never be hit, so the loops run
//                   unguarded, two pairs per iteration
//   amb  [S]  (ambient_light (.) ka .xyz, phong power)
//   diff [S]  (kd.xyz, ior)
//   spec [S]  (ks.xyz, r)
//   plpos[L]  (pos.xyz, 0)   plcol[L] (colour.xyz, 0)
//   dldir[D]  (normalize(dir).xyz, 0)   dlcol[D] (colour.xyz, 0)
//   foga [F]  (scattering, absorption, radius, 0)   fogalb[F] (albedo.xyz, 0)
//   fogp      float[S*L*F]: exp(-min(|c_s - Lp_i|, 2 r_j) * (abs_j + scat_j)), precomputed on the host with the
//             reference's own expression (src/blinn_phong.h:22-29)
//   cull [1+L][S4/2][3]  per apex A (0: camera, 1+i: point light i) and sphere pair, with u = c - A:
//             (ux0,ux1,uy0,uy1), (uz0,uz1,|u0|^2,|u1|^2), (R0,R1,|u0|,|u1|), R = inflated radius (see cull_pairs)
//   rmask     uint32[S*L]: pair mask of the spheres a shadow ray from ANY point of sphere s towards point light i could
//             hit (the cull_pairs test for the bundle "apex light i, points within r_s of c_s", evaluated on the host),
//             then float[S]: the squared radius around c_s inside which a hit point may use row s
// ------------------------------------------------------------------------------------------------
struct SceneView
{
	int S, S4, T, L, D, F;
	int off_geom, off_pgeom, off_pprim, off_amb, off_diff, off_spec, off_plpos, off_plcol, off_dldir, off_dlcol, off_foga, off_fogalb, off_fogp;
	int off_cull;		 // bundle-culling tables (see cull_pairs), or -1 when the scene has more than 64 spheres
	int cull_shadow;	 // 1: L * (S4/2) <= 64, the per-light masks of a pixel fit one 64-bit word
	int off_rmask;		 // static shadow masks per (receiver sphere, point light) + receiver check radii, or -1
	int blob_f4;		 // blob size in float4
	int blob_in_smem;	 // 1: kernels stage the blob in shared memory
	const float4 *blob;	 // device
	const float4 *tri_v; // 4 float4 = 64 B per triangle (v0 | original index, v1, v2, pad), LBVH leaf order
	const float4 *bvh;	 // 4 float4 per internal node (see skr_bvh.cuh)
	const float4 *big_v; // 3 float4 per outsized triangle, tested before the hierarchy (skr_bvh_build.cuh: morton_kernel)
	int nbig;
	int bvh_root_is_leaf; // T == 1
	// shaded-triangles mode (opt-in, beyond the reference): a second hierarchy over the ACTUAL triangles + their materials
	const float4 *tri_v2, *bvh2, *big_v2;
	int nbig2;
	const float4 *tri_mat; // 3 float4 per triangle in ORIGINAL order, or null
	const float *tris_raw; // the uploaded triangles, 9 floats each, ORIGINAL order
	int *err;			  // device error word (bit 1: BVH traversal stack overflow)
	float3 cam_pos, cam_dir, cam_up, cam_right, background;
};

struct Counters // per-thread, reduced at kernel exit when STATS
{
	unsigned ch, sh, st, stp, tt, nv, hits, le, se;
};
SKR_DEV void zero(Counters &c) { c.ch = c.sh = c.st = c.stp = c.tt = c.nv = c.hits = c.le = c.se = 0; }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Keying, identical to oracle/skr_oracle.c.  One call yields 128 bits, so draws
// the same thread needs together share a block:
//   counter = (pixel, sample, node, slot), key = (seed_lo, seed_hi)
//   jitter   : counter (pixel, sample >> 2, 0, 0), word sample & 3 -> r (31 bits): one call per 4 samples
//   fog      : slot 1 + i*F + j, shared by the diffuse and the specular call: words 0/1 -> their xi (31 bits),
//              words 2/3 -> their three scattering offsets, 10 bits each, offset = -1 + (k + 0.5) / 512
//   GI child : slot 1 + L*F + (c >> 1): children 2m, 2m+1 share a block, (r1, r2) = words (0,1) / (2,3)
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
// -1.0f + rand() / float(RAND_MAX / 2) (src/utils.h:219-221) from a 10-bit draw: -1 + (k + 0.5) / 512
SKR_DEV float rng_pm1_10(uint32_t k) { return fmaf(__uint2float_rn(k) + 0.5f, 1.0f / 512.0f, -1.0f); }

struct RngCtx
{
	uint32_t pixel, sample, node;
	uint2 key;
};
SKR_DEV uint4 rng_block(const RngCtx &r, uint32_t slot) { return philox4x32_10(make_uint4(r.pixel, r.sample, r.node, slot), r.key); }

// ------------------------------------------------------------------------------------------------
// Spheres
// ------------------------------------------------------------------------------------------------

// All sphere tests use the half-b form of the reference's quadratic (src/utils.h:87-121): with e = o - c,
// h = d.e (= b/2), cc = e.e - r^2, a = d.d:   disc/4 = h^2 - a*cc,   t2 = (-h - sqrt(h^2 - a*cc)) / a.
// Same roots, one multiply less per test; the winner's t is then recomputed with the reference's own expression.

#define SKR_TAG_BITS 6
#define SKR_TAG_MASK ((1u << SKR_TAG_BITS) - 1u)
// ranking keys of a sphere pair for incoherent rays: v = (-h - a) - sqrt(d4), positive iff 1.0 < t2 (see closest_sphere_table)
SKR_DEV float2 rank_keys(float2 h, float2 d4, float na)
{
	const float2 hm = add2(f2(-h.x, -h.y), splat2(na));
	return add2(hm, f2(-sqrt_approx(d4.x), -sqrt_approx(d4.y)));
}
// winner of the tagged minimum: index from the tag, u = v + a; keys at or above +inf (negative, NaN, no candidate) and
// v = 0 (t2 = 1.0 exactly) are misses
SKR_DEV void untag(uint32_t wmin, float a, int &best, float &umin)
{
	if(wmin - (SKR_TAG_MASK + 1u) < 0x7f800000u - (SKR_TAG_MASK + 1u))
	{
		best = (int) (wmin & SKR_TAG_MASK);
		umin = __uint_as_float(wmin & ~SKR_TAG_MASK) + a;
	}
}

// Closest sphere along (o, d) with 1.0 < t < inf, strict minimum, first wins ties (src/raytrace.h:149-165).
// PRIMARY: o is the camera position, e and cc come precomputed from the blob.
// Two spheres per step in packed FP32x2.  Selection on u = a*t2 = -h - sqrt(d4) (a > 0 is common to a ray's spheres):
//   t2 > 1        <=>  h < -a  and  d4 >= 0  and  cc + 2h > -a          (as in occluded())
//   u < umin      <=>  m < 0  or  d4 > m*m,  m = -h - umin              (no sqrt)
// so a square root is taken only when a sphere actually becomes the new closest (once or twice per ray).
// ORIGIN_TABLE: G is a pair table of (e, cc) for THIS ray's origin (camera rays: the blob's pprim; bounce rays of one
// hit: the per-group table expand_kernel builds in shared memory) -- the test is then 3 + 2 packed operations.
// Otherwise G is the blob's pgeom (-c, -r^2) and e, cc are formed per test from `o`.
// COHERENT picks branches (whole warps skip) over predicates (no divergence) for the rare selection work.
template <bool ORIGIN_TABLE, bool COHERENT, bool STATS>
SKR_DEV int closest_sphere_table(const float4 *__restrict__ G, int NP, int S, float3 o, float3 d, float &tmin, Counters &cnt)
{
	constexpr bool PRIMARY = ORIGIN_TABLE;
	const float a  = dot(d, d);
	const float na = -a;
	int best	   = -1;
	float umin	   = CUDART_INF_F;
	const bool tagged = NP <= (1 << (SKR_TAG_BITS - 1));
	uint32_t wmin	  = 0xffffffffu;
	const float2 dx = splat2(d.x), dy = splat2(d.y), dz = splat2(d.z), na2 = splat2(na), two = splat2(2.0f);
	const float2 ox = splat2(o.x), oy = splat2(o.y), oz = splat2(o.z);
#pragma unroll 2
	for(int p = 0; p < NP; p++)
	{
		const float4 g0 = G[2 * p], g1 = G[2 * p + 1];
		float2 h, cc;
		if(PRIMARY)
		{
			h  = fma2(dz, f2(g1.x, g1.y), fma2(dy, f2(g0.z, g0.w), mul2(dx, f2(g0.x, g0.y))));
			cc = f2(g1.z, g1.w);
		}
		else
		{
			const float2 ex = add2(ox, f2(g0.x, g0.y)), ey = add2(oy, f2(g0.z, g0.w)), ez = add2(oz, f2(g1.x, g1.y));
			h  = fma2(dz, ez, fma2(dy, ey, mul2(dx, ex)));
			cc = fma2(ez, ez, fma2(ey, ey, fma2(ex, ex, f2(g1.z, g1.w))));
		}
		const float2 d4 = fma2(h, h, mul2(na2, cc));
		if(STATS)
		{
			cnt.st += (2 * p < S) + (2 * p + 1 < S);
			cnt.se += (2 * p < S) + (2 * p + 1 < S);
			cnt.stp += (d4.x >= 0.0f) + (d4.y >= 0.0f);
		}
		if(COHERENT)
		{
			// camera rays are coherent: whole warps skip the selection of spheres their rays' lines miss
			if(d4.x >= 0.0f)
			{
				const float w = fmaf(2.0f, h.x, cc.x), m = -h.x - umin;
				if((h.x < na) & (w > na) & ((m < 0.0f) | (d4.x > m * m)))
				{
					umin = -h.x - sqrt_approx(d4.x); // ranking only: the winner's t is recomputed exactly
					best = 2 * p;
				}
			}
			if(d4.y >= 0.0f)
			{
				const float w = fmaf(2.0f, h.y, cc.y), m = -h.y - umin;
				if((h.y < na) & (w > na) & ((m < 0.0f) | (d4.y > m * m)))
				{
					umin = -h.y - sqrt_approx(d4.y);
					best = 2 * p + 1;
				}
			}
		}
		else
		{
			// bounce rays are incoherent: in every warp some lane's line hits this sphere, so branches would only add
			// divergence.  Rank candidates directly on v = u - a = (-h - a) - sqrt(d4) with an approximate MUFU square
			// root: 1.0 < t2 <=> v > 0.  The sphere index rides in the low SKR_TAG_BITS mantissa bits of v; as UNSIGNED
			// integers the positive floats order like floats and sit below every negative float and NaN (d4 < 0), so
			// one 3-input unsigned minimum per pair keeps the closest valid candidate, lower index on ties.  No branch,
			// no predicate chain; the caller recomputes the winner's t exactly (sphere_t_ref).
			if(tagged)
			{
				const float2 v = rank_keys(h, d4, na);
				wmin		   = __vimin3_u32(wmin, (__float_as_uint(v.x) & ~SKR_TAG_MASK) | (uint32_t) (2 * p),
											  (__float_as_uint(v.y) & ~SKR_TAG_MASK) | (uint32_t) (2 * p + 1));
			}
			else
			{
				// more than 2^SKR_TAG_BITS spheres: compare-and-select on u = -h - sqrt(d4), a < u < umin
				const float ux = -h.x - sqrt_approx(d4.x);
				const float uy = -h.y - sqrt_approx(d4.y);
				if((ux > a) & (ux < umin))
				{
					umin = ux;
					best = 2 * p;
				}
				if((uy > a) & (uy < umin))
				{
					umin = uy;
					best = 2 * p + 1;
				}
			}
		}
	}
	if(!COHERENT && tagged)
	{
		untag(wmin, a, best, umin);
	}
	tmin = best >= 0 ? __fdiv_rn(umin, a) : CUDART_INF_F;
	return best;
}

// closest_sphere_table<false, false> for K rays from the same origin (GI children of one hit): e = o - c and
// cc = e.e - r^2 of a sphere pair are formed once and serve all K rays -- 8 of the 25 instructions per (pair, ray).
// Per ray the operations and their order are those of the one-ray form (tagged ranking): same winner, same t.
// Requires NP <= 2^(SKR_TAG_BITS - 1) (closest_hit_xk checks).
template <int K, bool STATS, bool EXACT_T>
SKR_DEV void closest_sphere_xk(const float4 *__restrict__ G, int NP, int S, float3 o, const float3 (&d)[K], float (&t)[K], int (&s)[K], Counters &cnt)
{
	float a[K];
	uint32_t wm[K];
#pragma unroll
	for(int k = 0; k < K; k++)
	{
		a[k]  = dot(d[k], d[k]);
		wm[k] = 0xffffffffu;
	}
	const float2 ox = splat2(o.x), oy = splat2(o.y), oz = splat2(o.z);
#pragma unroll 2
	for(int p = 0; p < NP; p++)
	{
		const float4 g0 = G[2 * p], g1 = G[2 * p + 1];
		const float2 ex = add2(ox, f2(g0.x, g0.y)), ey = add2(oy, f2(g0.z, g0.w)), ez = add2(oz, f2(g1.x, g1.y));
		const float2 cc = fma2(ez, ez, fma2(ey, ey, fma2(ex, ex, f2(g1.z, g1.w))));
		if(STATS)
		{
			cnt.st += K * ((2 * p < S) + (2 * p + 1 < S));
			cnt.se += K * ((2 * p < S) + (2 * p + 1 < S));
		}
#pragma unroll
		for(int k = 0; k < K; k++)
		{
			const float2 h = fma2(splat2(d[k].z), ez, fma2(splat2(d[k].y), ey, mul2(splat2(d[k].x), ex)));
			const float2 q = fma2(h, h, mul2(splat2(-a[k]), cc));
			if(STATS)
			{
				cnt.stp += (q.x >= 0.0f) + (q.y >= 0.0f);
			}
			const float2 v = rank_keys(h, q, -a[k]);
			wm[k]		   = __vimin3_u32(wm[k], (__float_as_uint(v.x) & ~SKR_TAG_MASK) | (uint32_t) (2 * p),
										  (__float_as_uint(v.y) & ~SKR_TAG_MASK) | (uint32_t) (2 * p + 1));
		}
	}
#pragma unroll
	for(int k = 0; k < K; k++)
	{
		int b	 = -1;
		float um = CUDART_INF_F;
		untag(wm[k], a[k], b, um);
		t[k] = b >= 0 ? (EXACT_T ? __fdiv_rn(um, a[k]) : __fdividef(um, a[k])) : CUDART_INF_F; // exact only where it bounds a triangle query
		s[k] = b;
	}
}

template <bool PRIMARY, bool STATS>
SKR_DEV int closest_sphere(const float4 *__restrict__ B, const SceneView &sv, float3 o, float3 d, float &tmin, Counters &cnt)
{
	return closest_sphere_table<PRIMARY, PRIMARY, STATS>(B + (PRIMARY ? sv.off_pprim : sv.off_pgeom), sv.S4 >> 1, sv.S, o, d, tmin, cnt);
}

// The reference's own expression for the winner's t (src/raytrace.h:197-201 -> src/utils.h:87-110), so that the hit
// point is the reference's.  (Its sqrt/divide run in double and round to float; IEEE float ops here differ from
// that by at most one ulp on rare inputs -- below every tolerance, cheaper than FP64 sequences.)
SKR_DEV float sphere_t_ref(float3 o, float3 d, float3 c, float r, float t_fallback)
{
	const float3 e	 = sub_rn(o, c);
	const float a	 = dot_rn(d, d);
	const float b	 = __fmul_rn(2.0f, dot_rn(d, e));
	const float cc	 = __fsub_rn(dot_rn(e, e), __fmul_rn(r, r));
	const float disc = __fsub_rn(__fmul_rn(b, b), __fmul_rn(__fmul_rn(4.0f, a), cc));
	if(!(disc >= 0.0f))
	{
		return t_fallback; // the two forms disagree about a grazing hit: keep the loop's value
	}
	return __fdiv_rn(__fsub_rn(-b, __fsqrt_rn(disc)), __fmul_rn(2.0f, a));
}

// shadow(): ANY sphere with 1.0 < t2 < inf along the normalised direction from p + 1e-6 occludes; no light-distance
// bound, hits within 1.0 ignored (src/utils.h:42-58).
//   t2 > 1  <=>  disc >= 0  and  -h - a > 0  and  (h + a)^2 > h^2 - a*cc  <=>  ... and  cc + 2h + a > 0
// (the last term is |o + d - c|^2 - r^2: the point at t = 1 lies outside the sphere) -- no sqrt, no divide.
template <bool STATS>
SKR_DEV bool occluded(const float4 *__restrict__ B, const SceneView &sv, float3 p, float3 dir, Counters &cnt)
{
	const float3 o = adds_rn(p, 0.000001f);
	const float a  = dot(dir, dir);
	const float na = -a;
	const int NP   = sv.S4 >> 1;
	const float4 *__restrict__ G = B + sv.off_pgeom;
	const float2 dx = splat2(dir.x), dy = splat2(dir.y), dz = splat2(dir.z), na2 = splat2(na), two = splat2(2.0f);
	const float2 ox = splat2(o.x), oy = splat2(o.y), oz = splat2(o.z);
	if(STATS)
	{
		cnt.sh++;
	}
	for(int p2 = 0; p2 < NP; p2 += 2) // two pairs = four spheres per iteration, one exit test
	{
		bool any = false;
#pragma unroll
		for(int k = 0; k < 2; k++)
		{
			const int pp	= p2 + k;
			const float4 g0 = G[2 * pp], g1 = G[2 * pp + 1];
			const float2 ex = add2(ox, f2(g0.x, g0.y)), ey = add2(oy, f2(g0.z, g0.w)), ez = add2(oz, f2(g1.x, g1.y));
			const float2 h	= fma2(dz, ez, fma2(dy, ey, mul2(dx, ex)));
			const float2 cc = fma2(ez, ez, fma2(ey, ey, fma2(ex, ex, f2(g1.z, g1.w))));
			const float2 d4 = fma2(h, h, mul2(na2, cc));
			const float2 w	= fma2(two, h, cc);
			const bool occ0 = (h.x < na) & (d4.x >= 0.0f) & (w.x > na);
			const bool occ1 = (h.y < na) & (d4.y >= 0.0f) & (w.y > na);
			if(STATS)
			{
				cnt.se += (2 * pp < sv.S) + (2 * pp + 1 < sv.S);
			}
			if(STATS && !any) // count like the reference's loop: up to and including the first occluder
			{
				cnt.st += (2 * pp < sv.S) + ((2 * pp + 1 < sv.S) & !occ0);
				cnt.stp += (d4.x >= 0.0f) + ((d4.y >= 0.0f) & !occ0);
			}
			any |= occ0 | occ1;
		}
		if(any)
		{
			return true;
		}
	}
	return false;
}

// ------------------------------------------------------------------------------------------------
// Bundle culling.  The jitter samples of one pixel (--jsample n: n*n rays through one pixel, src/main.cpp:47-66) are a
// thin bundle of lines through one apex: the camera for the primary rays, a point light for the shadow rays of hit
// points that lie close together.  One conservative test per (bundle, sphere) -- "could ANY line of the bundle pass the
// sphere test?" -- replaces n*n exact tests for every sphere the bundle clearly misses; the exact tests then run over
// the surviving pairs only.  The result is bit-identical to testing every sphere: a culled sphere is one whose exact
// test (h*h - a*cc >= 0 in float) is guaranteed to fail, by the margins below.
//
// Lines through apex A with directions within angle beta of w.  Sphere (c, r), u = c - A.  Distance from c to a line of
// the bundle >= |u| (sin(angle(u, w)) - beta) = dist_c - |u| beta  (sin is 1-Lipschitz), dist_c^2 = |u|^2 - (u.w)^2/|w|^2.
// Cull iff dist_c^2 > (R + |u| beta + m)^2 with
//   R = sqrt(r^2 + 1e-4 |u|^2) + 1e-4 (1 + |A|_1 + |c|_1): covers the rounding of the exact test (its disc/4 carries an
//       error of ~1e-6 |o - c|^2: >= 80x margin), of this test, and lines that miss the apex by a few ulps of the coordinates;
//   m = 0.01 |w| for shadow bundles, whose ray origins lie |w| away from the apex (|o - c| <= |u| + |w|).
// Comparisons are written so that NaN keeps the sphere.  Returns one bit per sphere PAIR.
// ------------------------------------------------------------------------------------------------
template <bool STATS>
SKR_DEV uint32_t cull_pairs(const float4 *__restrict__ C, int NP, int S, float3 w, float beta, float m, Counters &cnt)
{
	if(STATS)
	{
		cnt.se += S; // one bundle test per sphere
	}
	const float2 ninv = splat2(-__fdividef(1.0f, dot(w, w)));
	const float2 wx = splat2(w.x), wy = splat2(w.y), wz = splat2(w.z), b2 = splat2(beta), m2 = splat2(m);
	uint32_t mask = 0;
	for(int p = 0; p < NP; p++)
	{
		const float4 c0 = C[3 * p], c1 = C[3 * p + 1], c2 = C[3 * p + 2];
		const float2 hw	 = fma2(wz, f2(c1.x, c1.y), fma2(wy, f2(c0.z, c0.w), mul2(wx, f2(c0.x, c0.y))));
		const float2 lhs = fma2(mul2(hw, hw), ninv, f2(c1.z, c1.w));
		const float2 X	 = fma2(f2(c2.z, c2.w), b2, add2(f2(c2.x, c2.y), m2));
		const float2 X2	 = mul2(X, X);
		const bool keep	 = !(lhs.x > X2.x) | !(lhs.y > X2.y);
		mask |= (uint32_t) keep << p;
	}
	return mask;
}

// closest_sphere_table<ORIGIN_TABLE = true, COHERENT = true> over the pairs of `mask` only (ascending, so the first
// sphere still wins ties).  `mask` is uniform across the warp: no divergence, broadcast shared-memory reads.
template <bool STATS, bool EXACT_T>
SKR_DEV int closest_sphere_masked(const float4 *__restrict__ G, uint32_t mask, int S, float3 d, float &tmin, Counters &cnt)
{
	const float a  = dot(d, d);
	const float na = -a;
	int best	   = -1;
	float umin	   = CUDART_INF_F;
	const float2 dx = splat2(d.x), dy = splat2(d.y), dz = splat2(d.z), na2 = splat2(na);
	if(STATS)
	{
		cnt.st += S; // the reference's count: it tests every sphere (the culled ones fail by construction)
	}
	for(uint32_t m = mask; m; m &= m - 1)
	{
		const int p		= __ffs(m) - 1;
		const float4 g0 = G[2 * p], g1 = G[2 * p + 1];
		const float2 h	= fma2(dz, f2(g1.x, g1.y), fma2(dy, f2(g0.z, g0.w), mul2(dx, f2(g0.x, g0.y))));
		const float2 cc = f2(g1.z, g1.w);
		const float2 d4 = fma2(h, h, mul2(na2, cc));
		if(STATS)
		{
			cnt.stp += (d4.x >= 0.0f) + (d4.y >= 0.0f);
			cnt.se += (2 * p < S) + (2 * p + 1 < S);
		}
		if(d4.x >= 0.0f)
		{
			const float w = fmaf(2.0f, h.x, cc.x), mm = -h.x - umin;
			if((h.x < na) & (w > na) & ((mm < 0.0f) | (d4.x > mm * mm)))
			{
				umin = -h.x - sqrt_approx(d4.x); // ranking only: the winner's t is recomputed exactly
				best = 2 * p;
			}
		}
		if(d4.y >= 0.0f)
		{
			const float w = fmaf(2.0f, h.y, cc.y), mm = -h.y - umin;
			if((h.y < na) & (w > na) & ((mm < 0.0f) | (d4.y > mm * mm)))
			{
				umin = -h.y - sqrt_approx(d4.y);
				best = 2 * p + 1;
			}
		}
	}
	// (IEEE division like closest_sphere_table: tmin is sphere_t_ref's fallback for grazing hits, and the frame must not depend
	// on which of the two loops found the sphere -- one pixel in 130 000 differed by an ulp between culling on and off)
	tmin = best >= 0 ? __fdiv_rn(umin, a) : CUDART_INF_F;
	return best;
}

// occluded() for TWO shadow rays from the same point (two lights): e = o - c and cc = e.e - r^2 of a sphere pair are formed
// once and serve both rays.  Per ray the operations are those of occluded().  Returns bit 0 / bit 1.
template <bool STATS>
SKR_DEV unsigned occluded_x2(const float4 *__restrict__ B, const SceneView &sv, float3 p, float3 dir0, float3 dir1, Counters &cnt)
{
	const float3 o	= adds_rn(p, 0.000001f);
	const float na0 = -dot(dir0, dir0), na1 = -dot(dir1, dir1);
	const int NP	= sv.S4 >> 1;
	const float4 *__restrict__ G = B + sv.off_pgeom;
	const float2 ox = splat2(o.x), oy = splat2(o.y), oz = splat2(o.z), two = splat2(2.0f);
	if(STATS)
	{
		cnt.sh += 2;
	}
	bool any0 = false, any1 = false;
	for(int p2 = 0; p2 < NP; p2 += 2) // two pairs = four spheres per iteration, one exit test
	{
#pragma unroll
		for(int k = 0; k < 2; k++)
		{
			const int pp	= p2 + k;
			const float4 g0 = G[2 * pp], g1 = G[2 * pp + 1];
			const float2 ex = add2(ox, f2(g0.x, g0.y)), ey = add2(oy, f2(g0.z, g0.w)), ez = add2(oz, f2(g1.x, g1.y));
			const float2 cc = fma2(ez, ez, fma2(ey, ey, fma2(ex, ex, f2(g1.z, g1.w))));
			const float2 h0 = fma2(splat2(dir0.z), ez, fma2(splat2(dir0.y), ey, mul2(splat2(dir0.x), ex)));
			const float2 h1 = fma2(splat2(dir1.z), ez, fma2(splat2(dir1.y), ey, mul2(splat2(dir1.x), ex)));
			const float2 q0 = fma2(h0, h0, mul2(splat2(na0), cc));
			const float2 q1 = fma2(h1, h1, mul2(splat2(na1), cc));
			const float2 w0 = fma2(two, h0, cc), w1 = fma2(two, h1, cc);
			const bool a0 = (h0.x < na0) & (q0.x >= 0.0f) & (w0.x > na0), b0 = (h0.y < na0) & (q0.y >= 0.0f) & (w0.y > na0);
			const bool a1 = (h1.x < na1) & (q1.x >= 0.0f) & (w1.x > na1), b1 = (h1.y < na1) & (q1.y >= 0.0f) & (w1.y > na1);
			if(STATS)
			{
				const int real = (2 * pp < sv.S) + (2 * pp + 1 < sv.S);
				cnt.se += 2 * real;
				if(!any0) // count like the reference's loops: each ray up to and including its first occluder
				{
					cnt.st += (2 * pp < sv.S) + ((2 * pp + 1 < sv.S) & !a0);
					cnt.stp += (q0.x >= 0.0f) + ((q0.y >= 0.0f) & !a0);
				}
				if(!any1)
				{
					cnt.st += (2 * pp < sv.S) + ((2 * pp + 1 < sv.S) & !a1);
					cnt.stp += (q1.x >= 0.0f) + ((q1.y >= 0.0f) & !a1);
				}
			}
			any0 |= a0 | b0;
			any1 |= a1 | b1;
		}
		if(any0 & any1)
		{
			break;
		}
	}
	return (any0 ? 1u : 0u) | (any1 ? 2u : 0u);
}

// occluded() over the pairs of `mask` only (per-lane mask).
template <bool STATS>
SKR_DEV bool occluded_masked(const float4 *__restrict__ B, const SceneView &sv, uint32_t mask, float3 p, float3 dir, Counters &cnt)
{
	const float3 o = adds_rn(p, 0.000001f);
	const float a  = dot(dir, dir);
	const float na = -a;
	const float4 *__restrict__ G = B + sv.off_pgeom;
	const float2 dx = splat2(dir.x), dy = splat2(dir.y), dz = splat2(dir.z), na2 = splat2(na), two = splat2(2.0f);
	const float2 ox = splat2(o.x), oy = splat2(o.y), oz = splat2(o.z);
	if(STATS)
	{
		cnt.sh++;
	}
	for(uint32_t m = mask; m; m &= m - 1)
	{
		const int pp	= __ffs(m) - 1;
		const float4 g0 = G[2 * pp], g1 = G[2 * pp + 1];
		const float2 ex = add2(ox, f2(g0.x, g0.y)), ey = add2(oy, f2(g0.z, g0.w)), ez = add2(oz, f2(g1.x, g1.y));
		const float2 h	= fma2(dz, ez, fma2(dy, ey, mul2(dx, ex)));
		const float2 cc = fma2(ez, ez, fma2(ey, ey, fma2(ex, ex, f2(g1.z, g1.w))));
		const float2 d4 = fma2(h, h, mul2(na2, cc));
		const float2 w	= fma2(two, h, cc);
		const bool occ0 = (h.x < na) & (d4.x >= 0.0f) & (w.x > na);
		const bool occ1 = (h.y < na) & (d4.y >= 0.0f) & (w.y > na);
		if(STATS)
		{
			cnt.stp += (d4.x >= 0.0f) + ((d4.y >= 0.0f) & !occ0);
			cnt.se += (2 * pp < sv.S) + (2 * pp + 1 < sv.S);
		}
		if(occ0 | occ1)
		{
			if(STATS)
			{
				cnt.st += 2 * pp + (occ0 ? 1 : 2); // the reference stops at its first occluder
			}
			return true;
		}
	}
	if(STATS)
	{
		cnt.st += sv.S;
	}
	return false;
}

// Shadow-ray masks of the hit points within `rho` of pc, one NP-bit field per point light (sv.cull_shadow).
template <bool STATS>
SKR_DEV uint64_t shadow_masks(const float4 *__restrict__ B, const SceneView &sv, float3 pc, float rho, Counters &cnt)
{
	const int NP		= sv.S4 >> 1;
	const uint32_t full = NP >= 32 ? 0xffffffffu : (1u << NP) - 1u;
	uint64_t all		= 0;
	for(int i = 0; i < sv.L; i++)
	{
		const float3 w	= pc - f3(B[sv.off_plpos + i]);
		const float len = sqrtf(dot(w, w));
		uint32_t mk		= full;
		if(rho <= 0.45f * len) // asin(x) <= 1.05 x there
		{
			mk = cull_pairs<STATS>(B + sv.off_cull + 3 * NP * (1 + i), NP, sv.S, w, __fdividef(1.07f * rho, len), 0.01f * len, cnt);
		}
		all |= (uint64_t) mk << (i * NP);
	}
	return all;
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

// bp::spherical_fog_shading (src/blinn_phong.h:19-44) + scattering_phase_function (src/utils.h:216-224).
// `r` is the Philox block of this (light, fog); call 0 = from diffuse_shading, 1 = from specular_shading.
SKR_DEV float3 fog_term(const float4 *__restrict__ B, const SceneView &sv, const uint4 &r, int call, float p_no, int j, float3 base, float3 alb_l,
						float3 lhat, float3 n)
{
	// base  = kd * lcol * (inv_d2 * max(0, n.lhat)): the no-interaction return, the same for both calls
	// alb_l = albedo_j * lcol
	if(rng_unit(call ? r.y : r.x) > p_no)
	{
		return base;
	}
	const uint32_t w = call ? r.w : r.z;
	const float fa	 = B[sv.off_foga + j].x;
	const float3 nd	 = f3(fmaf(rng_pm1_10(w & 1023u), fa, lhat.x), fmaf(rng_pm1_10((w >> 10) & 1023u), fa, lhat.y),
						  fmaf(rng_pm1_10((w >> 20) & 1023u), fa, lhat.z));
	return alb_l * fmaxf(0.0f, dot(n, nd));
}

// direct_illumination as HEAD computes it (src/raytrace.h:36-44): ambient + diffuse + specular.  One shadow ray per
// light serves both terms (the reference casts the same ray twice).  View direction is towards the CAMERA POSITION
// even for bounce hits (src/blinn_phong.h:93).
// COHERENT: primary hits (image-coherent warps): per-pixel bundle masks / static per-receiver masks.  Otherwise (bounce
// hits, every lane another receiver): no masks, shadow rays of two lights tested together.
template <bool STATS, bool FOG, bool COHERENT>
SKR_DEV float3 direct_light(const float4 *__restrict__ B, const SceneView &sv, bool use_shadows, const RngCtx &rng, int sidx, float3 p, float3 n,
							Counters &cnt, bool masked = false, uint64_t smask = 0)
{
	const int NP = sv.S4 >> 1;
	// Primary hits without a pixel bundle (single-sample pixels): the static per-receiver masks, if p really lies on
	// sphere sidx.  Not for bounce hits: every lane holds another receiver there, and 32 different per-lane pair loops
	// cost more than the uniform loop over all pairs (config 5: 216 ms against 208 ms).
	const uint32_t *rmask = nullptr;
	if(COHERENT && !masked && use_shadows && sv.off_rmask >= 0)
	{
		const uint32_t *tab = reinterpret_cast<const uint32_t *>(B + sv.off_rmask);
		const float3 q		= p - f3(B[sv.off_geom + sidx]);
		if(dot(q, q) <= __uint_as_float(tab[sv.S * sv.L + sidx]))
		{
			rmask = tab + sidx * sv.L;
		}
	}
	const float4 am = B[sv.off_amb + sidx];
	const float3 kd = f3(B[sv.off_diff + sidx]);
	const float3 ks		= f3(B[sv.off_spec + sidx]);
	const bool has_spec = ks.x != 0.0f || ks.y != 0.0f || ks.z != 0.0f;
	float3 col		= f3(am);
	const bool fog	= FOG && sv.F > 0;
	const float3 view = (has_spec && (!fog || sv.D > 0)) ? normalize_fast(sv.cam_pos - p) : f3(0.0f, 0.0f, 0.0f); // specular terms only
	// no masks (bounce hits): the shadow rays of two lights are tested together, sharing the per-sphere origin terms
	const bool paired = !COHERENT && use_shadows && sv.L >= 2 && sv.L <= 32;
	unsigned occ	  = 0;
	if(paired)
	{
		for(int i = 0; i + 1 < sv.L; i += 2)
		{
			const float3 lv0 = f3(B[sv.off_plpos + i]) - p, lv1 = f3(B[sv.off_plpos + i + 1]) - p;
			occ |= occluded_x2<STATS>(B, sv, p, lv0 * rsqrtf(dot(lv0, lv0)), lv1 * rsqrtf(dot(lv1, lv1)), cnt) << i;
		}
		if(sv.L & 1)
		{
			const float3 lv = f3(B[sv.off_plpos + sv.L - 1]) - p;
			occ |= (occluded<STATS>(B, sv, p, lv * rsqrtf(dot(lv, lv)), cnt) ? 1u : 0u) << (sv.L - 1);
		}
	}
	for(int i = 0; i < sv.L; i++)
	{
		const float3 lv	  = f3(B[sv.off_plpos + i]) - p;
		const float d2	  = dot(lv, lv);
		const float3 lhat = lv * rsqrtf(d2); // also the shadow-ray direction
		if(paired)
		{
			if((occ >> i) & 1u)
			{
				continue;
			}
		}
		else if(use_shadows)
		{
			if((COHERENT && masked) ? occluded_masked<STATS>(B, sv, (uint32_t) (smask >> (i * NP)) & (uint32_t) ((1ull << NP) - 1ull), p, lhat, cnt)
			   : (COHERENT && rmask) ? occluded_masked<STATS>(B, sv, rmask[i], p, lhat, cnt)
									 : occluded<STATS>(B, sv, p, lhat, cnt))
			{
				continue;
			}
		}
		if(STATS)
		{
			cnt.le++;
		}
		const float3 lcol  = f3(B[sv.off_plcol + i]);
		const float inv_d2 = __fdividef(1.0f, d2);
		if(fog)
		{
			const float3 base = kd * lcol * (inv_d2 * fmaxf(0.0f, dot(n, lhat)));
			const float *fogp = reinterpret_cast<const float *>(B + sv.off_fogp) + (sidx * sv.L + i) * sv.F;
			for(int j = 0; j < sv.F; j++)
			{
				const uint4 r	   = rng_block(rng, 1u + (uint32_t) (i * sv.F + j));
				const float3 alb_l = f3(B[sv.off_fogalb + j]) * lcol;
				const float p_no   = fogp[j];
				col += fog_term(B, sv, r, 0, p_no, j, base, alb_l, lhat, n);
				col += fog_term(B, sv, r, 1, p_no, j, base, alb_l, lhat, n);
			}
		}
		else
		{
			col += kd * lcol * (inv_d2 * fmaxf(0.0f, dot(n, lhat)));
			if(has_spec)
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
		if(has_spec)
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
	__sincosf(6.28318530717958648f * r2, &sn, &cs); // phi = 2*pi*r2 in [0, 2pi]: MUFU sin/cos, abs. error ~2^-21
	const float sx = s_theta * cs, sy = r1, sz = s_theta * sn;
	return f3(sx * nb.x + sy * n.x + sz * nt.x, sx * nb.y + sy * n.y + sz * nb.y, sx * nb.z + sy * n.z + sz * nb.z);
}

// ------------------------------------------------------------------------------------------------
// One closest-hit query = the first half of shade() (src/raytrace.h:149-192).
// Returns: -2 background, -1 triangle (black), >= 0 sphere index with t in tmin.
// ------------------------------------------------------------------------------------------------
// `cand` (PRIMARY only, optional): when non-null the triangle query is DEFERRED -- *cand is set for rays whose line
// reaches the hierarchy under the root, and the ray is answered as if no triangle were hit; tri_deferred_kernel settles it.
template <bool PRIMARY, bool STATS, bool TRIS>
SKR_DEV int closest_hit(const float4 *__restrict__ B, const SceneView &sv, float3 o, float3 d, float &tmin, Counters &cnt, bool masked = false,
						uint32_t pmask = 0, bool *cand = nullptr)
{
	if(STATS)
	{
		cnt.ch++;
	}
	const int s = (PRIMARY && masked) ? closest_sphere_masked<STATS, TRIS>(B + sv.off_pprim, pmask, sv.S, d, tmin, cnt)
									  : closest_sphere<PRIMARY, STATS>(B, sv, o, d, tmin, cnt);
	if(TRIS && sv.T > 0)
	{
		if(PRIMARY && cand)
		{
			const int r = tri_any_hit_line_deferred<STATS>(sv, o, d, tmin, cnt);
			if(r == 1)
			{
				return -1;
			}
			*cand = r == 2;
		}
		else if(tri_any_hit_line<STATS>(sv, o, d, tmin, cnt))
		{
			return -1;
		}
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

// K closest-hit queries from one origin (bounce rays; see closest_sphere_xk).
template <int K, bool STATS, bool TRIS>
SKR_DEV void closest_hit_xk(const float4 *__restrict__ B, const SceneView &sv, float3 o, const float3 (&d)[K], float (&t)[K], int (&h)[K], Counters &cnt)
{
	if((sv.S4 >> 1) > (1 << (SKR_TAG_BITS - 1))) // too many spheres for index tags: one by one
	{
#pragma unroll
		for(int k = 0; k < K; k++)
		{
			h[k] = closest_hit<false, STATS, TRIS>(B, sv, o, d[k], t[k], cnt);
		}
		return;
	}
	if(STATS)
	{
		cnt.ch += K;
	}
	closest_sphere_xk<K, STATS, TRIS>(B + sv.off_pgeom, sv.S4 >> 1, sv.S, o, d, t, h, cnt);
#pragma unroll
	for(int k = 0; k < K; k++)
	{
		h[k] = h[k] < 0 ? -2 : h[k]; // no sphere: background, unless a triangle claims the ray
		if(TRIS && sv.T > 0 && tri_any_hit_line<STATS>(sv, o, d[k], t[k], cnt))
		{
			h[k] = -1;
		}
		if(STATS)
		{
			cnt.hits += h[k] >= 0;
		}
	}
}
