/* oracle/skr_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C CPU restatement ("port") of the reference's per-pixel tracing loop
 * (reference: /root/reference/src/main.cpp:19-104 -> src/raytrace.h:139 shade()
 * and everything below it).  See skr_oracle.c for the per-function citations.
 *
 * Pinning:  tests/test_oracle_*.py check this port
 *   - against the reference's only golden vector, renders/testcpu.ppm
 *     (dragon.scn 640x480, byte-exact), committed as tests/golden/;
 *   - against the reference's OWN code compiled in place (oracle/_ref/
 *     libskr_ref.so, built by oracle/Makefile) bit-for-bit: every helper on
 *     random inputs, and whole frames in every mode, including the stochastic
 *     ones, by consuming libc rand() in the reference's order (rng_mode 0);
 *   - against golden float images generated from the reference and committed
 *     under tests/golden/ (so the pin survives where /root/reference and
 *     oracle/_ref are absent).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * arm may use this library; the product (libskr.so) never links or loads it.
 */
#ifndef SKR_ORACLE_H
#define SKR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat scene snapshot -- same layout as include/skr.h (skr_scene_desc). */
typedef struct skro_scene
{
	int nspheres;
	const float *spheres; /* 18 floats: c3 r | ambient3 | diffuse3 | specular3 | transmissive3 | power ior */
	int ntris;
	const float *tris; /* 9 floats: v0 v1 v2 */
	int nplights;
	const float *plights; /* 6 floats: position3 colour3 */
	int ndlights;
	const float *dlights; /* 6 floats: direction3 colour3 */
	int nfogs;
	const float *fogs; /* 9 floats: scattering absorption | albedo3 | radius | centre3 */
	float camera[12];  /* position | direction | up | right (none normalised) */
	float ambient[3];
	float background[3];
	const float *tri_materials; /* optional, 14 floats per triangle (ambient3 diffuse3 specular3 transmissive3 power ior): read ONLY by
								   the non-reference shaded-triangles extension (skr_oracle_ext.inc) */
} skro_scene;

enum
{
	SKRO_RNG_LIBC	= 0, /* rand() in the reference's call order; serial only */
	SKRO_RNG_PHILOX = 1	 /* Philox4x32-10 keyed (seed; pixel, sample, node, slot) -- same keying as the CUDA path */
};

typedef struct skro_options
{
	int width, height;
	float fov;			 /* degrees, Options::fov src/utils.h:30 */
	int max_depth;		 /* Options::max_depth */
	int monte_carlo;	 /* Options::monte_carlo (--gillum given) */
	int num_path_traces; /* Options::num_path_traces */
	int grid_size;		 /* Options::grid_size (--jsample) */
	int use_shadows;	 /* Scene::use_shadows (--shadow) */
	int fresnel;		 /* 0 = HEAD behaviour; 1 = recursion of src/raytrace.h:46-103 live */
	int rng_mode;
	uint64_t seed;
	int threads;
	int y0, y1; /* row window [y0,y1) */
	int shade_triangles; /* NOT reference behaviour: this repository's extension, see skr_oracle_ext.inc */
} skro_options;

typedef struct skro_stats
{
	uint64_t closest_hit_rays; /* shade() calls with depth > 0 */
	uint64_t shadow_rays;	   /* distinct shadow() queries (the reference issues each twice; counted once) */
	uint64_t sphere_tests;	   /* ray/sphere quadratic evaluations (one per sphere per ray, not the reference's up-to-3x) */
	uint64_t sphere_tests_pos; /* ... of which discriminant >= 0 */
	uint64_t tri_tests;		   /* ray/triangle tests */
	uint64_t sphere_hits;	   /* closest-hit rays that ended on a sphere (shaded) */
	uint64_t light_evals;	   /* (shaded hit, point light) pairs that were lit (not shadowed) */
} skro_stats;

/* Renders rows [y0,y1); returns wall seconds of the pixel loop.
 * rgb32 (H*W*3 floats, pre-clamp) and rgb8 (H*W*3 bytes) are optional. */
double skro_render(const skro_scene *scene, const skro_options *opt, float *rgb32, unsigned char *rgb8, skro_stats *stats);

/* one shade() call on an arbitrary ray (rng_mode/seed from opt; pixel=sample=node=0) */
void skro_shade(const skro_scene *scene, const skro_options *opt, const float *o, const float *d, int depth, float *rgb);

/* helpers exposed for known-answer tests against the reference's own functions */
float skro_smallest_root(float a, float b, float c);
float skro_sphere_hit(const float *o, const float *d, const float *c, float r, int *occurs);
int skro_triangle_hit(const float *o, const float *d, const float *tri9, float *tuv);
void skro_transform_coordinate_space(const float *n, float *nt, float *nb);
void skro_uniform_sample_hemi(float r1, float r2, float *out);
int skro_shadow_point(const skro_scene *scene, const float *p, const float *light6);
float skro_fresnel(const float *dir, const float *n, float ior);
void skro_refraction(const float *dir, const float *n, float ior, float *out);
void skro_reflect_direction(const float *l, const float *n, float *out);
void skro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int skro_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
