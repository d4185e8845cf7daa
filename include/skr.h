/* include/skr.h -- C ABI of the B200-native renderer core (libskr.so).
 *
 * This is the drop-in boundary for skele-raytracer's per-pixel tracing loop.
 * The reference has no plugin/FFI interface; the seam is its frame function
 *
 *     void generate_rays_parallel(Scene scene, Options option, char *output)
 *                                              reference: src/main.cpp:19-104
 *
 * called from main() at src/main.cpp:409 (sibling generate_rays, :108-227,
 * called at :404).  A maintainer swaps the body of that function for the four
 * calls below: flatten `Scene` into skr_scene_desc,
 * skr_scene_upload(), skr_render(), write the PPM as before
 * (src/main.cpp:88-100).  Everything crosses the boundary as plain pointers
 * and sizes; no C++ types, no exceptions, no torch types.
 *
 * Conventions: every entry point returns 0 on success and a nonzero
 * skr_status on failure; skr_last_error() then holds a message.  The caller
 * owns all host buffers; the library owns all device memory and its stream.
 * A context is bound to one CUDA device and is not re-entrant.  There is no
 * CPU fallback: without a usable CUDA device skr_init() fails.
 */
#ifndef SKR_H
#define SKR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKR_ABI_VERSION 3

typedef struct skr_ctx skr_ctx;

typedef enum skr_status
{
	SKR_OK			  = 0,
	SKR_ERR_CUDA	  = 1, /* a CUDA runtime call failed (message has the call and cudaGetErrorString) */
	SKR_ERR_ARG		  = 2, /* invalid argument */
	SKR_ERR_NO_SCENE  = 3, /* skr_render before skr_scene_upload */
	SKR_ERR_NO_DEVICE = 4  /* no CUDA device / device index out of range */
} skr_status;

/* Flat, tightly packed mirror of the reference's `Scene` (src/scene.h:13-28).
 * All arrays are float32, row-major, caller-owned; the library copies them.
 *
 *   spheres   [nspheres][18]  Sphere (src/shapes.h:12-24) = SphereCollider (src/SphereCollider.h:8-12)
 *                             + Material (src/material.h:9-26):
 *                             cx cy cz r | ambient rgb | diffuse rgb | specular rgb | transmissive rgb | power ior
 *   tris      [ntris][9]      Triangle (src/shapes.h:27-33) v0 v1 v2 (its material is never read by the reference,
 *                             src/raytrace.h:221-224; see tri_materials below)
 *   plights   [nplights][6]   PointLight (src/lights.h:18-22) position xyz | colour rgb
 *   dlights   [ndlights][6]   DirectionalLight (src/lights.h:12-16) direction xyz | colour rgb
 *                             (the reference parser never stores any, src/scene.cpp:139-163)
 *   fogs      [nfogs][9]      SphericalFog (src/Fog.h:10-32) scattering absorption | albedo rgb | radius | centre xyz
 *   camera    [12]            Camera (src/camera.h:8-32) position | direction | up | right, as the reference
 *                             leaves them: NOT normalised, right = cross(-direction, up)
 *   ambient   [3]             AmbientLight::colour (src/lights.h:8-10)
 *   background[3]             Scene::background
 */
typedef struct skr_scene_desc
{
	int32_t nspheres;
	const float *spheres;
	int32_t ntris;
	const float *tris;
	int32_t nplights;
	const float *plights;
	int32_t ndlights;
	const float *dlights;
	int32_t nfogs;
	const float *fogs;
	float camera[12];
	float ambient[3];
	float background[3];
	/* ABI 3, optional (may be NULL): the Material of every triangle, [ntris][14] = ambient rgb | diffuse rgb | specular rgb |
	 * transmissive rgb | power ior (src/shapes.h:27-33, src/scene.cpp:67-81).  The reference never reads it; only the
	 * opt-in shaded-triangles mode (skr_options.shade_triangles) does. */
	const float *tri_materials;
} skr_scene_desc;

/* Mirror of the reference's `Options` (src/utils.h:26-39) plus the per-frame
 * scene fields main() overrides from the command line (src/main.cpp:393-396). */
typedef struct skr_options
{
	int32_t width;			 /* --width   (Scene::width,  default 1920) */
	int32_t height;			 /* --height  (Scene::height, default 1080) */
	float fov;				 /* --fov     degrees (Options::fov, default 60) */
	int32_t max_depth;		 /* --depth   (Options::max_depth, default 3) */
	int32_t monte_carlo;	 /* --gillum given (Options::monte_carlo) */
	int32_t num_path_traces; /* --gillum n (Options::num_path_traces) */
	int32_t grid_size;		 /* --jsample n (Options::grid_size; n*n samples per pixel) */
	int32_t use_shadows;	 /* --shadow  (Scene::use_shadows) */
	int32_t fresnel;		 /* 0 = behaviour of HEAD; 1 = opt-in: the reflect/refract recursion of
								src/raytrace.h:46-103 that HEAD skips by returning at :44 */
	uint64_t seed;			 /* Philox key; replaces srand(time(0)) (src/main.cpp:400) */
	/* frame split (additive; the reference has no multi-GPU): the image is cut into tile x tile pixel
	 * tiles, row-major tile index k belongs to rank k % world.  world <= 1 renders the whole frame. */
	int32_t rank;
	int32_t world;
	int32_t tile;			/* tile edge in pixels; 0 = default (32) */
	int32_t collect_stats;	/* nonzero: count rays/tests on the device (slower; not for timed runs) */
	int32_t queue_capacity; /* entries (52 B each) per wavefront queue level; 0 = default: four fan-outs of the frame's samples, 32 M .. 128 M */
	/* ABI 3.  0 = behaviour of HEAD: any triangle hit shades black (src/raytrace.h:221-224).  1 = opt-in, explicitly
	 * NON-PARITY extension: triangles are shaded with their own Material -- closest hit on the actual
	 * (un-mirrored) triangle in front of the ray, geometric normal facing the ray, the Blinn-Phong terms of
	 * src/blinn_phong.h:47-134, shadow rays stopped by triangles as well as spheres; fog is ignored.  Not combinable
	 * with monte_carlo / fresnel (SKR_ERR_ARG). */
	int32_t shade_triangles;
} skr_options;

typedef struct skr_stats
{
	/* counters (valid when collect_stats != 0) */
	uint64_t closest_hit_rays; /* shade() invocations with depth > 0 (src/raytrace.h:139) */
	uint64_t shadow_rays;	   /* distinct shadow() queries (src/utils.h:42; the reference issues each twice) */
	uint64_t sphere_tests;
	uint64_t sphere_tests_pos; /* ... with discriminant >= 0 */
	uint64_t tri_tests;		   /* triangle leaf tests */
	uint64_t bvh_node_visits;  /* BVH nodes whose child boxes were tested */
	uint64_t sphere_hits;	   /* closest-hit rays that ended on a sphere */
	uint64_t light_evals;	   /* (shaded hit, light) pairs that were lit */
	/* always valid */
	uint64_t queue_entries; /* wavefront queue entries written */
	uint32_t kernel_launches;
	uint32_t queue_chunks;
	float ms_total;	  /* device time of the whole render (CUDA events on the library's stream) */
	float ms_primary; /* ray generation + closest hit (+ inline shading when no --gillum) */
	float ms_bounce;  /* shade + expand kernels of the --gillum / fresnel wavefront */
	float ms_resolve; /* accumulate -> float/RGB8 */
	float ms_h2d;	  /* 0 for *_device entry points */
	float ms_d2h;
	/* counter (valid when collect_stats != 0), ABI 2: sphere tests the kernels actually EXECUTED, bundle-culling tests
	 * included.  sphere_tests above stays the reference algorithm's count (every sphere per query, shadow loops up to
	 * their first occluder); the difference is what conservative culling proved unnecessary. */
	uint64_t sphere_tests_executed;
} skr_stats;

/* device < 0: use the current CUDA device. */
int skr_init(int device, skr_ctx **out);
void skr_destroy(skr_ctx *ctx);
const char *skr_last_error(const skr_ctx *ctx); /* ctx may be NULL: last error of skr_init on this thread */
int skr_abi_version(void);
const char *skr_build_info(void); /* ABI 3: compiler, target architectures and tuning macros this binary was built with */

/* Scene upload: AoS -> SoA device buffers (spheres/materials/lights), triangles flattened to
 * float4 vertex triples and a device-built LBVH (Morton codes, radix sort, Karras hierarchy, refit). */
int skr_scene_upload(skr_ctx *ctx, const skr_scene_desc *scene);

/* Optional (ABI 3): allocates, ahead of time, everything frames with these options need -- the wavefront queue arena and
 * the accumulators of a --gillum / fresnel tree, the device frame of skr_render -- and loads the kernels they launch,
 * so that the first frame costs what every later frame costs.  Without it the first frame does this itself (outside
 * its device-timed span).  Allocations are kept and reused by later frames and uploads. */
int skr_reserve(skr_ctx *ctx, const skr_options *opt);

/* Renders one frame (or this rank's tiles of it) and copies the result to HOST buffers.
 *   rgb8  : H*W*3 bytes, row-major top-down RGB, (unsigned char)(min(1,c)*255) as src/main.cpp:96; may be NULL
 *   rgb32 : H*W*3 floats, the pre-clamp image (for tests); may be NULL
 * With world > 1 only this rank's tiles are rendered; the other ranks' pixels are written as 0.
 * Page-locked destinations (cudaHostAlloc / cudaHostRegister / skr_pin_host) are the fast path: frames that are one long
 * kernel (>= 4 samples per pixel, no --gillum) are stored into an RGB8-only host frame by the kernel itself, or leave in
 * bands copied while the kernel is still running; pageable memory gets a plain copy after the frame.  Same bytes each way. */
int skr_render(skr_ctx *ctx, const skr_options *opt, uint8_t *rgb8, float *rgb32, skr_stats *stats);

/* Same, results left in DEVICE memory (pointers valid on ctx's device; either may be NULL).
 * With stats == NULL, collect_stats == 0 and no --gillum/fresnel tree the call is ASYNCHRONOUS: the frame (one
 * kernel) is enqueued on skr_stream() and the call returns at once; order later work on that stream or skr_sync().
 * The same holds for skr_render_tiles_device and skr_deinterleave_device. */
int skr_render_device(skr_ctx *ctx, const skr_options *opt, void *d_rgb8, void *d_rgb32, skr_stats *stats);

/* Multi-GPU frame split.  skr_render_tiles_device renders this rank's tiles into a COMPACT tile-major
 * RGB8 buffer d_tiles of skr_tiles_bytes() bytes (local tile j = global tile j*world + rank, each tile
 * tile*tile*3 bytes, row-major inside the tile, edge tiles padded).  After the ranks' buffers have been
 * gathered rank-major into one buffer of world*skr_tiles_bytes() bytes (one NCCL all-gather / gather),
 * skr_deinterleave_device turns it into the row-major H*W*3 frame. */
int64_t skr_tiles_bytes(const skr_options *opt);
int skr_render_tiles_device(skr_ctx *ctx, const skr_options *opt, void *d_tiles, skr_stats *stats);
int skr_deinterleave_device(skr_ctx *ctx, const skr_options *opt, const void *d_gathered, void *d_rgb8);

/* Frame split WITHOUT a collective: renders this rank's tiles and stores every finished pixel straight into n_frames
 * row-major RGB8 frames (H*W*3 bytes each) -- typically one per GPU of the box, the peers' buffers mapped into this
 * process (CUDA IPC, cuMem fabric handles, torch symmetric memory ...).  The stores to peer GPUs travel over NVLink
 * while the kernel is still tracing, so no gather and no de-interleave pass remain; the caller runs the cross-rank
 * barrier after the call (all ranks' kernels complete => every frame is whole) and double-buffers the frames if it
 * reads one while the next is being rendered.  1 <= n_frames <= 8.  Asynchronous like skr_render_device. */
int skr_render_peers_device(skr_ctx *ctx, const skr_options *opt, void *const *d_frames, int n_frames, skr_stats *stats);

/* Frame split whose result leaves over EVERY GPU's PCIe link (ABI 3).  Like skr_render_peers_device, but a finished pixel
 * goes to ONE of the frames: d_frames[min(y / rows_per_frame, n_frames - 1)], y its image row (every buffer is addressed
 * as a whole H*W*3 frame; only its own band of rows is ever written).  With one frame per GPU of the box, after the ranks'
 * barrier GPU k holds rows [k * rows_per_frame, (k + 1) * rows_per_frame) complete in its own memory -- whichever GPU
 * rendered them, the stores crossed NVLink while the kernels were still tracing -- and copies that band to the host
 * itself: the D2H of the frame is N parallel copies instead of one.  rows_per_frame: a positive multiple of 4. */
int skr_render_bands_device(skr_ctx *ctx, const skr_options *opt, void *const *d_frames, int n_frames, int rows_per_frame, skr_stats *stats);

/* Frames straight into HOST memory (ABI 3).  A page-locked, device-mapped host buffer is a valid target of
 * skr_render_peers_device: the kernel that finishes a pixel stores it over PCIe while the rest of the frame is still
 * being traced, and with a frame split every GPU writes its own tiles over its OWN PCIe link -- no device frame, no
 * gather, no D2H copy.  skr_pin_host page-locks and maps a caller-owned buffer (e.g. a shared-memory segment that several
 * one-process-per-GPU ranks map) and returns the pointer to pass as d_frames[k]; it returns 0 when it registered the
 * buffer (undo with skr_unpin_host), 1000 when the buffer was page-locked already (cudaHostAlloc, torch pin_memory:
 * nothing to undo), another skr_status on error. */
int skr_pin_host(skr_ctx *ctx, void *host, size_t bytes, void **d_ptr);
int skr_unpin_host(skr_ctx *ctx, void *host);
/* Device -> host copy enqueued on the library's stream, i.e. behind the frames rendered so far (asynchronous when host_dst is
 * page-locked; skr_sync() waits for it): how a rank copies its band of skr_render_bands_device out. */
int skr_copy_to_host(skr_ctx *ctx, void *host_dst, const void *d_src, size_t bytes);

/* The library's stream as a cudaStream_t (so that callers can order their own work / events after it),
 * and a blocking wait for it. */
void *skr_stream(skr_ctx *ctx);
int skr_sync(skr_ctx *ctx);

/* FP32 FMA microbenchmark (register-resident FFMA chains on every SM): returns TFLOP/s, <0 on error.
 * Used by bench.py as the measured FP32 roofline denominator (MEASURED_PEAKS.json has none). */
double skr_measure_fp32_peak(skr_ctx *ctx, int iters);

/* Read-bandwidth microbenchmarks of the on-chip memory levels the BVH / primitive fetch goes through (ABI 3):
 * level 0 = shared memory (LDS.128), 1 = L1 (ld.global.ca over a window resident in every SM's L1), 2 = L2
 * (ld.global.cg over 64 MB).  Returns GB/s over all SMs, <0 on error.  Denominators for bench.py's triangle-path
 * roofline (MEASURED_PEAKS.json has HBM only). */
double skr_measure_bandwidth(skr_ctx *ctx, int level);

#ifdef __cplusplus
}
#endif
#endif /* SKR_H */
