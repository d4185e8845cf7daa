/* oracle/skr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.  See skr_oracle.h.
 *
 * Plain-C restatement of the reference's per-pixel tracing loop.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference).  Floating-point expressions keep the reference's operand
 * order, its float/double mix (which calls resolve to the double libm
 * overloads was read off the compiled reference: sqrt in smallest_root and
 * fresnel, exp in the fog term, tan in the camera angle) and glm 0.9.5.4's
 * formulas (normalize = v * (1.0f / sqrt(x*x+y*y+z*z)),
 * src/glm/detail/func_geometric.inl:256-265; dot = (x0*y0 + x1*y1) + x2*y2,
 * :66-73; vec/scalar = per-component divide, src/glm/detail/type_vec3.inl:580-590),
 * and this file is built with -ffp-contract=off like the reference oracle, so
 * that in rng_mode SKRO_RNG_LIBC the port is BIT-IDENTICAL to the reference's
 * own compiled code (tests/test_oracle_port.py).
 *
 * What is deliberately NOT restated: the reference passes `Scene` by value
 * into every call (deep-copying all vectors per ray).  That is an
 * implementation cost, not part of the algorithm; the port reads the scene in
 * place.  Likewise the quadratic is solved once per sphere, not up to three
 * times (src/utils.h:171, src/raytrace.h:157, :197-201) -- same values.
 */
#define _GNU_SOURCE
#include "skr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct
{
	float x, y, z;
} vec3;

static inline vec3 v3(float x, float y, float z)
{
	vec3 r = {x, y, z};
	return r;
}
static inline vec3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }
static inline void st3(float *p, vec3 v)
{
	p[0] = v.x;
	p[1] = v.y;
	p[2] = v.z;
}
static inline vec3 add(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 sub(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 mul(vec3 a, vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline vec3 muls(vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline vec3 divs(vec3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline vec3 adds(vec3 a, float s) { return v3(a.x + s, a.y + s, a.z + s); }
static inline vec3 neg(vec3 a) { return v3(-a.x, -a.y, -a.z); }
/* glm::dot(vec3): tmp = x*y; tmp.x + tmp.y + tmp.z  (func_geometric.inl:66-73) */
static inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* glm::cross (func_geometric.inl:219-228) */
static inline vec3 cross(vec3 x, vec3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
/* glm::length(vec3) = sqrt(dot(v,v)) (func_geometric.inl:108-114) */
static inline float length3(vec3 v) { return sqrtf(dot(v, v)); }
/* glm::normalize(vec3) (func_geometric.inl:256-265; inversesqrt = 1.0f / sqrt(x), func_exponential.inl:226-229) */
static inline vec3 normalize(vec3 v)
{
	float sqr = v.x * v.x + v.y * v.y + v.z * v.z;
	return muls(v, 1.0f / sqrtf(sqr));
}

/* ---------------------------------------------------------------- RNG ---- */

/* Philox4x32-10 (Salmon et al., SC'11), the standard Random123 constants. */
void skro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for(int r = 0; r < 10; r++)
	{
		uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
		uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t) p1;
		uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t) p0;
		c0			= n0;
		c1			= n1;
		c2			= n2;
		c3			= n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0;
	out[1] = c1;
	out[2] = c2;
	out[3] = c3;
}

typedef struct
{
	const skro_scene *scene;
	const skro_options *opt;
	skro_stats stats;
	uint32_t key[2];
	uint32_t pixel, sample;
	uint32_t node_base; /* children of node k are k*node_base + c + 1 */
	uint32_t slot_gi;	/* first slot of the GI (r1,r2) draws */
} ctx_t;

/* Slot map of the keyed stream (identical in the CUDA path, csrc/skr_device.cuh).  Philox hands out 128 bits per
 * call, so independent draws are packed into one block wherever the same thread needs them together:
 *   counter = (pixel, sample, node, slot), key = (seed_lo, seed_hi)
 *   jitter   : counter (pixel, sample >> 2, 0, 0), word sample & 3          -> r of this primary sample (31 bits)
 *   fog      : slot 1 + i*F + j (light i, fog j), shared by the diffuse and the specular call:
 *                word 0 -> xi of the diffuse call (31 bits), word 1 -> xi of the specular call (31 bits),
 *                word 2 -> the diffuse call's three scattering offsets, 10 bits each (bits 0-9, 10-19, 20-29),
 *                word 3 -> the specular call's three offsets, same packing;  offset = -1 + (k + 0.5) / 512
 *   GI child : slot 1 + L*F + (c >> 1): children 2m and 2m+1 share a block, (r1, r2) = words (0,1) / (2,3), 31 bits
 * In SKRO_RNG_LIBC mode none of this applies: every draw is one rand() call in the reference's order.
 */
static inline void philox_block(const ctx_t *c, uint32_t sample, uint32_t node, uint32_t slot, uint32_t out[4])
{
	uint32_t ctr[4] = {c->pixel, sample, node, slot};
	skro_philox4x32_10(ctr, c->key, out);
}
/* static_cast<float>(rand()) / static_cast<float>(RAND_MAX)   (RAND_MAX = 2^31-1, as float 2^31) */
static inline float unit_from_31(uint32_t k) { return (float) (int32_t) k / (float) 2147483647; }
/* -1.0f + rand() / float(RAND_MAX / 2)   src/utils.h:219-221 */
static inline float pm1_from_31(uint32_t k) { return -1.0f + (float) (int32_t) k / (float) (2147483647 / (1 + 1)); }
static inline float pm1_from_10(uint32_t k) { return -1.0f + ((float) k + 0.5f) * (1.0f / 512.0f); }

static inline float draw_jitter(const ctx_t *c)
{
	if(c->opt->rng_mode == SKRO_RNG_LIBC)
	{
		return unit_from_31((uint32_t) rand());
	}
	uint32_t w[4];
	philox_block(c, c->sample >> 2, 0, 0, w);
	return unit_from_31(w[c->sample & 3] >> 1);
}
static inline void draw_gi(const ctx_t *c, uint32_t node, int child, float *r1, float *r2)
{
	if(c->opt->rng_mode == SKRO_RNG_LIBC)
	{
		*r1 = unit_from_31((uint32_t) rand());
		*r2 = unit_from_31((uint32_t) rand());
		return;
	}
	uint32_t w[4];
	philox_block(c, c->sample, node, c->slot_gi + ((uint32_t) child >> 1), w);
	*r1 = unit_from_31(w[(child & 1) * 2] >> 1);
	*r2 = unit_from_31(w[(child & 1) * 2 + 1] >> 1);
}

/* ----------------------------------------------------------- geometry ---- */

/* src/utils.h:87-110.  sqrt and the divide run in double (unqualified sqrt on
 * a float resolves to ::sqrt(double) there), then round to float. */
float skro_smallest_root(float a, float b, float c)
{
	float discriminant = b * b - 4 * a * c;
	if(discriminant < 0)
	{
		return INFINITY;
	}
	float t1 = (float) (((double) (-b) + sqrt((double) discriminant)) / (double) (2 * a));
	float t2 = (float) (((double) (-b) - sqrt((double) discriminant)) / (double) (2 * a));
	if(t1 < t2 && t1 >= 0)
	{
		return t1;
	}
	else if(t2 >= 0)
	{
		return t2;
	}
	return INFINITY;
}

/* src/utils.h:113-121 collision_distance */
static inline float collision_distance(vec3 o, vec3 d, vec3 centre, float radius, ctx_t *c)
{
	vec3 e_c = sub(o, centre);
	float a	 = dot(d, d);
	float b	 = 2 * dot(d, e_c);
	float cc = dot(e_c, e_c) - radius * radius;
	if(c)
	{
		c->stats.sphere_tests++;
		if(b * b - 4 * a * cc >= 0)
		{
			c->stats.sphere_tests_pos++;
		}
	}
	return skro_smallest_root(a, b, cc);
}

/* src/utils.h:169-179 intersection_occurs: 1.0 < t < inf */
static inline int intersection_occurs_t(float distance)
{
	if(distance <= 1.0f || distance == INFINITY)
	{
		return 0;
	}
	return 1;
}

float skro_sphere_hit(const float *o, const float *d, const float *c, float r, int *occurs)
{
	float t = collision_distance(ld3(o), ld3(d), ld3(c), r, NULL);
	*occurs = intersection_occurs_t(t);
	return t;
}

/* src/utils.h:181-213 triangle_intersection_occurs (u negated -> mirrored
 * triangle; no sign test on t) */
static inline int triangle_hit(vec3 o, vec3 dir, const float *tri, float *t, float *u, float *v)
{
	vec3 v0 = ld3(tri), v1 = ld3(tri + 3), v2 = ld3(tri + 6);
	vec3 v0v1 = sub(v1, v0);
	vec3 v0v2 = sub(v2, v0);
	vec3 p	  = cross(dir, v0v2);
	float d	  = dot(v0v1, p);
	if(fabs(d) < 0.00001f)
	{
		return 0;
	}
	float inverse = 1.0f / d;
	vec3 t_vector = sub(o, v0);
	*u			  = inverse * dot(neg(t_vector), p);
	if(*u < 0 || *u > 1)
	{
		return 0;
	}
	vec3 q = cross(t_vector, v0v1);
	*v	   = dot(dir, q) * inverse;
	if(*v < 0 || *u + *v > 1)
	{
		return 0;
	}
	*t = dot(v0v2, q) * inverse;
	return 1;
}

int skro_triangle_hit(const float *o, const float *d, const float *tri9, float *tuv)
{
	tuv[0] = tuv[1] = tuv[2] = 0;
	return triangle_hit(ld3(o), ld3(d), tri9, &tuv[0], &tuv[1], &tuv[2]);
}

/* src/utils.h:42-58 shadow(Scene, P, PointLight): origin P + 1e-6 (broadcast),
 * direction normalised; ANY sphere with 1.0 < t < inf occludes. */
static int shadow_dir(ctx_t *c, const skro_scene *s, vec3 p, vec3 direction)
{
	vec3 o = adds(p, 0.000001f);
	for(int i = 0; i < s->nspheres; i++)
	{
		const float *sp = s->spheres + 18 * i;
		if(intersection_occurs_t(collision_distance(o, direction, ld3(sp), sp[3], c)))
		{
			return 1;
		}
	}
	return 0;
}
static int shadow_point(ctx_t *c, const skro_scene *s, vec3 p, vec3 light_pos)
{
	return shadow_dir(c, s, p, normalize(sub(light_pos, p)));
}
int skro_shadow_point(const skro_scene *scene, const float *p, const float *light6)
{
	return shadow_point(NULL, scene, ld3(p), ld3(light6));
}

/* src/utils.h:148-165 transform_coordinate_space */
static void transform_coordinate_space(vec3 n, vec3 *perp_to_normal, vec3 *perp_to_both)
{
	if(fabsf(n.x) > fabsf(n.y))
	{
		*perp_to_normal = divs(v3(n.z, 0, -n.x), sqrtf(n.x * n.x + n.z * n.z));
	}
	else
	{
		*perp_to_normal = divs(v3(0, -n.z, n.y), sqrtf(n.y * n.y + n.z * n.z));
	}
	*perp_to_both = cross(n, *perp_to_normal);
}
void skro_transform_coordinate_space(const float *n, float *nt, float *nb)
{
	vec3 a, b;
	transform_coordinate_space(ld3(n), &a, &b);
	st3(nt, a);
	st3(nb, b);
}

/* src/raytrace.h:22-30 uniform_sample_hemi (phi = 2.0f * M_PI * r2 in double) */
static vec3 uniform_sample_hemi(float r1, float r2)
{
	float s_theta = sqrtf(1 - r1 * r1);
	float phi	  = (float) (2.0f * M_PI * r2);
	float x		  = s_theta * cosf(phi);
	float z		  = s_theta * sinf(phi);
	return v3(x, r1, z);
}
void skro_uniform_sample_hemi(float r1, float r2, float *out)
{
	st3(out, uniform_sample_hemi(r1, r2));
}

/* --------------------------------------------------- dead-at-HEAD trio ---- */

static float clampf(float a, float b, float input) /* src/utils.h:132-145 */
{
	if(input < a)
	{
		return a;
	}
	else if(input > b)
	{
		return b;
	}
	return input;
}

/* src/blinn_phong.h:137-140 */
static vec3 reflect_direction(vec3 l, vec3 n)
{
	return normalize(sub(l, muls(n, 2.0f * dot(l, n))));
}
void skro_reflect_direction(const float *l, const float *n, float *out)
{
	st3(out, reflect_direction(ld3(l), ld3(n)));
}

/* src/blinn_phong.h:143-153 */
static vec3 refraction(vec3 dir, vec3 n, float ior)
{
	float dn = dot(dir, n);
	float k	 = 1.0f - (ior * ior) * (1.0f - dn * dn);
	if(k < 0.0f)
	{
		return v3(0, 0, 0);
	}
	return sub(muls(dir, ior), muls(n, ior * dot(dir, n) + sqrtf(k)));
}
void skro_refraction(const float *dir, const float *n, float ior, float *out)
{
	st3(out, refraction(ld3(dir), ld3(n), ior));
}

/* src/blinn_phong.h:156-184 (sint's sqrt stays double because it is multiplied
 * into a float*double product; cos_theta's is an exact float sqrt) */
static float fresnel(vec3 ray_direction, vec3 normal, float ior_in)
{
	float cos_internal = clampf(-1.0f, 1.0f, dot(ray_direction, normal));
	float et		   = 1.0f;
	float ior		   = ior_in;
	if(cos_internal > 0)
	{
		float tmp = et;
		et		  = ior;
		ior		  = tmp;
	}
	float sint = (float) ((double) (et / ior) * sqrt((double) fmaxf(0.0f, 1.0f - cos_internal * cos_internal)));
	if(sint >= 1.0f)
	{
		return 1.0f;
	}
	float cos_theta = sqrtf(fmaxf(0.0f, 1 - sint * sint));
	cos_internal	= fabsf(cos_internal);
	float Rs		= ((ior * cos_internal) - (et * cos_theta)) / ((ior * cos_internal) + (et * cos_theta));
	float Rp		= ((et * cos_internal) - (ior * cos_theta)) / ((ior * cos_internal) + (et * cos_theta));
	return (Rs * Rs + Rp * Rp) / 2.0f;
}
float skro_fresnel(const float *dir, const float *n, float ior)
{
	return fresnel(ld3(dir), ld3(n), ior);
}

/* ------------------------------------------------------------ shading ---- */

static vec3 shade(ctx_t *c, vec3 o, vec3 d, int depth, uint32_t node);

/* src/blinn_phong.h:19-44 spherical_fog_shading + src/utils.h:216-224 */
static vec3 spherical_fog_shading(ctx_t *c, uint32_t node, uint32_t slot, int call, const float *light, const float *fog, const float *sphere, vec3 light_direction, vec3 p, vec3 norm)
{
	vec3 lpos = ld3(light), lcol = ld3(light + 3);
	float scattering = fog[0], absorption = fog[1], fog_radius = fog[5];
	float distance = length3(sub(ld3(sphere), lpos));
	if(distance > 2 * fog_radius)
	{
		distance = 2 * fog_radius;
	}
	float probability_no_interaction = (float) exp((double) (-1.0f * distance * (absorption + scattering)));
	const int libc = c->opt->rng_mode == SKRO_RNG_LIBC;
	uint32_t w[4]  = {0, 0, 0, 0};
	if(!libc)
	{
		philox_block(c, c->sample, node, slot, w);
	}
	float random_num = unit_from_31(libc ? (uint32_t) rand() : w[call] >> 1);
	if(random_num > probability_no_interaction)
	{
		distance		= length3(sub(lpos, p));
		float intensity = 1.0f / (fabsf(distance) * fabsf(distance));
		return muls(muls(mul(ld3(sphere + 7), lcol), intensity), fmaxf(0.0f, dot(norm, light_direction)));
	}
	float x = libc ? pm1_from_31((uint32_t) rand()) : pm1_from_10(w[2 + call] & 1023u);
	float y = libc ? pm1_from_31((uint32_t) rand()) : pm1_from_10((w[2 + call] >> 10) & 1023u);
	float z = libc ? pm1_from_31((uint32_t) rand()) : pm1_from_10((w[2 + call] >> 20) & 1023u);
	vec3 nd = v3(light_direction.x + x * scattering, light_direction.y + y * scattering, light_direction.z + z * scattering);
	return muls(mul(ld3(fog + 2), lcol), fmaxf(0.0f, dot(norm, nd)));
}

/* src/blinn_phong.h:47-87 diffuse_shading.  lit[i] records the shadow result
 * so that specular_shading's duplicate shadow ray is not counted twice. */
static vec3 diffuse_shading(ctx_t *c, uint32_t node, const float *sphere, vec3 p, vec3 norm)
{
	const skro_scene *s = c->scene;
	vec3 colour			= v3(0, 0, 0);
	for(int i = 0; i < s->nplights; i++)
	{
		const float *light = s->plights + 6 * i;
		int shadowed	   = 0;
		if(c->opt->use_shadows)
		{
			c->stats.shadow_rays++;
			shadowed = shadow_point(c, s, p, ld3(light));
		}
		if(!shadowed)
		{
			c->stats.light_evals++;
			vec3 light_direction = normalize(sub(ld3(light), p));
			if(s->nfogs > 0)
			{
				for(int j = 0; j < s->nfogs; j++)
				{
					uint32_t slot = 1u + (uint32_t) (i * s->nfogs + j);
					colour		  = add(colour, spherical_fog_shading(c, node, slot, 0, light, s->fogs + 9 * j, sphere, light_direction, p, norm));
				}
			}
			else
			{
				float distance	= length3(sub(ld3(light), p));
				float intensity = 1.0f / (fabsf(distance) * fabsf(distance));
				colour			= add(colour, muls(muls(mul(ld3(sphere + 7), ld3(light + 3)), intensity), fmaxf(0.0f, dot(norm, light_direction))));
			}
		}
	}
	for(int i = 0; i < s->ndlights; i++)
	{
		const float *light = s->dlights + 6 * i;
		int shadowed	   = 0;
		if(c->opt->use_shadows)
		{
			c->stats.shadow_rays++;
			shadowed = shadow_dir(c, s, p, normalize(ld3(light)));
		}
		if(!shadowed)
		{
			vec3 light_direction = normalize(ld3(light));
			colour				 = add(colour, muls(mul(ld3(sphere + 7), ld3(light + 3)), fmaxf(0.0f, dot(norm, light_direction))));
		}
	}
	return colour;
}

/* src/blinn_phong.h:90-134 specular_shading.  View direction is towards the
 * CAMERA POSITION even for bounce rays. */
static vec3 specular_shading(ctx_t *c, uint32_t node, const float *sphere, vec3 p, vec3 norm)
{
	const skro_scene *s = c->scene;
	vec3 colour			= v3(0, 0, 0);
	vec3 view_direction = normalize(sub(ld3(s->camera), p));
	for(int i = 0; i < s->nplights; i++)
	{
		const float *light = s->plights + 6 * i;
		if(!c->opt->use_shadows || !shadow_point(NULL, s, p, ld3(light)))
		{
			vec3 light_direction = normalize(sub(ld3(light), p));
			vec3 hsum			 = add(view_direction, light_direction);
			vec3 half_vector	 = divs(hsum, length3(hsum));
			if(s->nfogs > 0)
			{
				for(int j = 0; j < s->nfogs; j++)
				{
					uint32_t slot = 1u + (uint32_t) (i * s->nfogs + j);
					colour		  = add(colour, spherical_fog_shading(c, node, slot, 1, light, s->fogs + 9 * j, sphere, light_direction, p, norm));
				}
			}
			else
			{
				float distance	= length3(sub(ld3(light), p));
				float intensity = 1.0f / (fabsf(distance) * fabsf(distance));
				colour			= add(colour, muls(muls(mul(ld3(sphere + 10), ld3(light + 3)), intensity), powf(fmaxf(0.0f, dot(norm, half_vector)), sphere[16])));
			}
		}
	}
	for(int i = 0; i < s->ndlights; i++)
	{
		const float *light = s->dlights + 6 * i;
		if(!c->opt->use_shadows || !shadow_dir(NULL, s, p, normalize(ld3(light))))
		{
			vec3 light_direction = normalize(ld3(light));
			vec3 hsum			 = add(view_direction, light_direction);
			vec3 half_vector	 = divs(hsum, length3(hsum));
			colour				 = add(colour, muls(mul(ld3(sphere + 10), ld3(light + 3)), powf(fmaxf(0.0f, dot(norm, half_vector)), sphere[16])));
		}
	}
	return colour;
}

/* src/raytrace.h:36-104 direct_illumination.  HEAD returns at :44; the
 * recursion at :46-103 runs only with opt->fresnel. */
static vec3 direct_illumination(ctx_t *c, uint32_t node, vec3 ray_d, const float *sphere, vec3 p, vec3 norm, int depth)
{
	const skro_scene *s = c->scene;
	vec3 total			= v3(0, 0, 0);
	total				= add(total, mul(ld3(s->ambient), ld3(sphere + 4))); /* bp::ambient_shading src/blinn_phong.h:13-17 */
	total				= add(total, diffuse_shading(c, node, sphere, p, norm));
	total				= add(total, specular_shading(c, node, sphere, p, norm));
	if(!c->opt->fresnel)
	{
		return total;
	}

	float ior			   = sphere[17];
	vec3 specular		   = ld3(sphere + 10);
	float fr			   = fresnel(ray_d, norm, ior);
	vec3 refraction_colour = v3(0, 0, 0);
	vec3 reflection_colour = v3(0, 0, 0);
	uint32_t n_gi		   = c->opt->monte_carlo ? (uint32_t) c->opt->num_path_traces : 0u;
	if((specular.x != 0.0f || specular.y != 0.0f || specular.z != 0.0f) && depth > 0)
	{
		int nl = s->nplights + s->ndlights;
		for(int i = 0; i < nl; i++)
		{
			vec3 light_direction = i < s->nplights ? normalize(sub(ld3(s->plights + 6 * i), p)) : normalize(ld3(s->dlights + 6 * (i - s->nplights)));
			if(fr < 1)
			{
				vec3 rd			  = refraction(ray_d, norm, ior);
				uint32_t child	  = node * c->node_base + (n_gi + 2u * (uint32_t) i) + 1u;
				refraction_colour = muls(shade(c, p, rd, depth - 1, child), fr); /* assignment: last light wins */
			}
			vec3 rd			  = reflect_direction(light_direction, norm);
			uint32_t child	  = node * c->node_base + (n_gi + 2u * (uint32_t) i + 1u) + 1u;
			reflection_colour = add(reflection_colour, mul(muls(specular, 1 - fr), shade(c, p, rd, depth - 1, child)));
		}
	}
	return add(add(total, refraction_colour), reflection_colour);
}

/* src/raytrace.h:107-136 montecarlo_global_illumination (the local->world
 * transform bug at :123-125 is reproduced) */
static vec3 montecarlo_global_illumination(ctx_t *c, uint32_t node, vec3 p, vec3 n, int depth, int num_rays)
{
	vec3 total = v3(0, 0, 0);
	vec3 perp_to_normal, perp_to_both;
	transform_coordinate_space(n, &perp_to_normal, &perp_to_both);
	float probability_dist = (float) (1 / (M_PI));
	for(int i = 0; i < num_rays; i++)
	{
		float r1, r2;
		draw_gi(c, node, i, &r1, &r2);
		vec3 sample	  = uniform_sample_hemi(r1, r2);
		vec3 world	  = v3(sample.x * perp_to_both.x + sample.y * n.x + sample.z * perp_to_normal.x,
						   sample.x * perp_to_both.y + sample.y * n.y + sample.z * perp_to_both.y,
						   sample.x * perp_to_both.z + sample.y * n.z + sample.z * perp_to_both.z);
		vec3 origin	  = adds(p, 0.00001f);
		uint32_t child = node * c->node_base + (uint32_t) i + 1u;
		total		   = add(total, divs(muls(shade(c, origin, world, depth - 1, child), r1), probability_dist));
	}
	total = divs(total, (float) num_rays);
	return total;
}

/* src/raytrace.h:139-227 shade */
static vec3 shade(ctx_t *c, vec3 o, vec3 d, int depth, uint32_t node)
{
	const skro_scene *s = c->scene;
	if(depth <= 0)
	{
		return v3(0, 0, 0);
	}
	c->stats.closest_hit_rays++;

	float min_distance = INFINITY;
	int hit_sphere_idx = -1;
	int hit_a_sphere   = 0;
	for(int i = 0; i < s->nspheres; i++)
	{
		const float *sp = s->spheres + 18 * i;
		float distance	= collision_distance(o, d, ld3(sp), sp[3], c);
		if(intersection_occurs_t(distance))
		{
			hit_a_sphere = 1;
			if(distance < min_distance)
			{
				min_distance   = distance;
				hit_sphere_idx = i;
			}
		}
	}

	int hit_a_triangle = 0;
	for(int i = 0; i < s->ntris; i++)
	{
		float t, u, v;
		c->stats.tri_tests++;
		if(triangle_hit(o, d, s->tris + 9 * i, &t, &u, &v))
		{
			if(t < min_distance)
			{
				min_distance   = t;
				hit_a_sphere   = 0;
				hit_a_triangle = 1;
			}
		}
	}

	if(!hit_a_sphere && !hit_a_triangle)
	{
		return ld3(s->background);
	}

	if(hit_a_sphere)
	{
		c->stats.sphere_hits++;
		const float *sp = s->spheres + 18 * hit_sphere_idx;
		vec3 centre		= ld3(sp);
		vec3 e_c		= sub(o, centre);
		float a			= dot(d, d);
		float b			= 2 * dot(d, e_c);
		float cc		= dot(e_c, e_c) - sp[3] * sp[3];
		float t			= skro_smallest_root(a, b, cc);
		vec3 p			= add(o, muls(d, t));
		vec3 n			= normalize(sub(p, centre));
		vec3 direct		= direct_illumination(c, node, d, sp, p, n, depth);
		if(c->opt->monte_carlo)
		{
			vec3 indirect = montecarlo_global_illumination(c, node, p, n, depth, c->opt->num_path_traces);
			return mul(add(divs(direct, (float) M_PI), muls(indirect, 2.0f)), ld3(sp + 7));
		}
		return direct;
	}
	return v3(0, 0, 0); /* any triangle hit shades black, src/raytrace.h:221-224 */
}

#include "skr_oracle_ext.inc" /* NOT reference behaviour: this repository's shaded-triangles extension */

/* ---------------------------------------------------------- frame loop ---- */

static void ctx_init(ctx_t *c, const skro_scene *scene, const skro_options *opt)
{
	memset(c, 0, sizeof *c);
	c->scene   = scene;
	c->opt	   = opt;
	c->key[0]  = (uint32_t) opt->seed;
	c->key[1]  = (uint32_t) (opt->seed >> 32);
	uint32_t n = opt->monte_carlo ? (uint32_t) opt->num_path_traces : 0u;
	c->node_base = n + 1u + (opt->fresnel ? 2u * (uint32_t) (scene->nplights + scene->ndlights) : 0u);
	c->slot_gi	 = 1u + (uint32_t) scene->nplights * (uint32_t) scene->nfogs;
}

static void stats_add(skro_stats *a, const skro_stats *b)
{
	a->closest_hit_rays += b->closest_hit_rays;
	a->shadow_rays += b->shadow_rays;
	a->sphere_tests += b->sphere_tests;
	a->sphere_tests_pos += b->sphere_tests_pos;
	a->tri_tests += b->tri_tests;
	a->sphere_hits += b->sphere_hits;
	a->light_evals += b->light_evals;
}

/* src/main.cpp:33-86 (pixel loop, without the overrides at :21-24) and :88-100 (quantiser) */
double skro_render(const skro_scene *scene, const skro_options *opt, float *rgb32, unsigned char *rgb8, skro_stats *stats)
{
	const int width = opt->width, height = opt->height;
	int y0 = opt->y0 < 0 ? 0 : opt->y0;
	int y1 = opt->y1 > height ? height : opt->y1;
	int threads = opt->threads < 1 ? 1 : opt->threads;
	if(opt->rng_mode == SKRO_RNG_LIBC)
	{
		srand((unsigned) opt->seed);
	}
	vec3 cam_pos = ld3(scene->camera), cam_dir = ld3(scene->camera + 3), cam_up = ld3(scene->camera + 6), cam_right = ld3(scene->camera + 9);
	skro_stats total;
	memset(&total, 0, sizeof total);
	float *image = (float *) calloc((size_t) width * (size_t) height * 3, sizeof(float));

	struct timespec ts0, ts1;
	clock_gettime(CLOCK_MONOTONIC, &ts0);
#pragma omp parallel num_threads(threads) if(threads > 1)
	{
		ctx_t c;
		ctx_init(&c, scene, opt);
#pragma omp for schedule(dynamic, 1)
		for(int y = y0; y < y1; y++)
		{
			for(int x = 0; x < width; x++)
			{
				float inv_width	   = 1 / (float) width;
				float inv_height   = 1 / (float) height;
				float aspect_ratio = width / (float) height;
				float angle		   = (float) tan(M_PI * 0.5 * opt->fov / 180.);
				vec3 px			   = v3(0, 0, 0);
				c.pixel			   = (uint32_t) (y * width + x);
				if(opt->grid_size > 0)
				{
					for(int i = 0; i < opt->grid_size; i++)
					{
						for(int j = 0; j < opt->grid_size; j++)
						{
							c.sample	= (uint32_t) (i * opt->grid_size + j);
							float r		= draw_jitter(&c);
							float u		= (2 * ((x + r) * inv_width) - 1) * angle * aspect_ratio;
							float v		= (1 - 2 * ((y + r) * inv_height)) * angle;
							vec3 ray_dir = add(add(cam_dir, muls(cam_right, u)), muls(cam_up, v)); /* never normalised, SURVEY F8 */
							px			 = add(px, opt->shade_triangles ? shade_ext(&c, cam_pos, ray_dir, opt->max_depth) : shade(&c, cam_pos, ray_dir, opt->max_depth, 0));
						}
					}
					px = divs(px, (float) (opt->grid_size * opt->grid_size));
				}
				else
				{
					c.sample	 = 0;
					float u		 = (float) ((2 * ((x + 0.5) * inv_width) - 1) * angle * aspect_ratio);
					float v		 = (float) ((1 - 2 * ((y + 0.5) * inv_height)) * angle);
					vec3 ray_dir = add(add(cam_dir, muls(cam_right, u)), muls(cam_up, v));
					px			 = opt->shade_triangles ? shade_ext(&c, cam_pos, ray_dir, opt->max_depth) : shade(&c, cam_pos, ray_dir, opt->max_depth, 0);
				}
				st3(image + 3 * ((size_t) y * width + x), px);
			}
		}
#pragma omp critical
		stats_add(&total, &c.stats);
	}
	clock_gettime(CLOCK_MONOTONIC, &ts1);

	for(int y = y0; y < y1; y++)
	{
		for(int x = 0; x < width; x++)
		{
			size_t i = (size_t) y * width + x;
			for(int k = 0; k < 3; k++)
			{
				float cch = image[3 * i + k];
				if(rgb32)
				{
					rgb32[3 * i + k] = cch;
				}
				if(rgb8)
				{
					float m = cch < 1.0f ? cch : 1.0f; /* std::min(float(1), c): c if c < 1 else 1 (NaN -> 1) */
					rgb8[3 * i + k] = (unsigned char) (int) (m * 255);
				}
			}
		}
	}
	free(image);
	if(stats)
	{
		*stats = total;
	}
	return (double) (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double) (ts1.tv_nsec - ts0.tv_nsec);
}

void skro_shade(const skro_scene *scene, const skro_options *opt, const float *o, const float *d, int depth, float *rgb)
{
	ctx_t c;
	ctx_init(&c, scene, opt);
	if(opt->rng_mode == SKRO_RNG_LIBC)
	{
		srand((unsigned) opt->seed);
	}
	st3(rgb, shade(&c, ld3(o), ld3(d), depth, 0));
}

int skro_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
