// skr_api.cu -- implementation of the C ABI in include/skr.h (libskr.so).
//
// Replaces the body of the reference's frame function generate_rays_parallel (reference src/main.cpp:19-104).
// Host side of the renderer: scene flattening/upload (+ device LBVH build), wavefront scheduling (queue levels,
// chunking so that a full fan-out always fits the next level), output copies, timing.  No CPU rendering path exists
// in this file or anywhere in the library: if CUDA is unavailable every entry point fails.
#include "../../include/skr.h"

#include <cuda.h> // types of the stream memory operations only; entry points come from cudaGetDriverEntryPoint
#include <algorithm>
#include <cmath>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "skr_bvh_build.cuh"
#include "skr_kernels.cuh"
#include "skr_micro.cuh"
#include "skr_shaded.cuh"

namespace
{
thread_local std::string g_init_error;

enum
{
	CAT_PRIMARY = 0,
	CAT_BOUNCE	= 1,
	CAT_RESOLVE = 2,
	CAT_COUNT	= 3
};

struct Span
{
	cudaEvent_t a, b;
	int cat;
};

constexpr size_t SMEM_BLOB_LIMIT = 64 * 1024;
constexpr unsigned MAX_QUEUE_CAP = 128u << 20; // entries per level (52 B each, 6.7 GB): fewer, fuller chunks -- config 5: 228 ms at 32 M, 213 ms at 128 M
constexpr unsigned MIN_QUEUE_CAP = 32u << 20;
constexpr int DEFAULT_TILE = 32;
constexpr int MAX_BANDS = 8;
constexpr int BRUTE_FORCE_TRIS = 4;
constexpr int BIG_TRI_CAP = 16;	 // outsized triangles kept out of the hierarchy (morton_kernel)
constexpr int BIG_TRI_MIN_T = 64; // scenes smaller than this keep everything in the tree
} // namespace

struct BvhGraphKey // compared with memcmp: no padding
{
	long long T;
	const void *tris_raw, *tri_v, *bvh, *scratch;
};

// One triangle hierarchy: the triangles in LBVH leaf order, the nodes, the outsized-triangle list and the CUDA graph that
// replays its build.  A context holds two: over the MIRRORED triangles for the reference's line any-hit query, and --
// built on first use -- over the actual triangles for the opt-in shaded-triangles mode.
struct BvhSet
{
	float4 *d_tri_v = nullptr;
	float4 *d_bvh	= nullptr;
	float4 *d_big	= nullptr; // BIG_TRI_CAP x 3 float4 + the int counter behind them
	size_t tri_v_bytes = 0, bvh_bytes = 0;
	BvhGraphKey graph_key{};
	cudaGraphExec_t graph = nullptr;
	bool graph_broken	  = false; // capture failed once: direct launches from then on
	const float4 *bvh	  = nullptr; // result: nodes, or null = test every triangle (a handful, or SKR_NO_BVH=1)
	bool valid			  = false;

	void release()
	{
		cudaFree(d_tri_v), cudaFree(d_bvh), cudaFree(d_big);
		if(graph)
		{
			cudaGraphExecDestroy(graph);
		}
		d_tri_v = d_bvh = d_big = nullptr;
		graph					= nullptr;
	}
};

struct skr_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	int sm_count = 0;

	// scene
	bool have_scene = false;
	SceneView sv{};
	float4 *d_blob = nullptr;
	float *d_tris_raw = nullptr;
	BvhSet bvh_main, bvh_shade;
	float4 *d_tri_mat = nullptr; // 3 float4 per triangle (original order): (ambient (.) ka, power), (kd, ior), (ks, 0); shaded-triangles mode
	size_t tri_mat_bytes = 0;
	bool have_tri_mat = false;
	int n_tris = 0;
	size_t blob_bytes = 0, tris_raw_bytes = 0, scratch_bytes = 0;
	char *d_scratch = nullptr; // LBVH build scratch, kept between uploads
	size_t smem_bytes = 0;

	// frame buffers
	uint8_t *d_rgb8 = nullptr;
	size_t rgb8_bytes = 0;
	float *d_rgb32 = nullptr;
	size_t rgb32_bytes = 0;
	long long *d_accum = nullptr;
	size_t accum_bytes = 0;

	// wavefront queues: ONE arena (counters first, then a / b / c / d of every level), carved by ensure_queues
	std::vector<Queue> queues;
	unsigned queue_cap = 0;
	unsigned *d_counts = nullptr; // one per level (inside the arena)
	int n_levels_alloc = 0;
	char *d_arena = nullptr;
	size_t arena_bytes = 0;
	unsigned *h_count = nullptr; // pinned

	unsigned long long *d_counters = nullptr; // 9 (16 allocated)
	// heavy-first launch order of the tiles (single-kernel frames), cached while scene and frame geometry stay the same
	int *d_tile_order = nullptr;
	size_t tile_order_bytes = 0;
	std::vector<int> tile_order_host;
	int order_band_perm[MAX_BANDS] = {};		   // cached with the order: launch position -> band of rows (skr_render's overlapped copy-out)
	unsigned order_band_start[MAX_BANDS + 1] = {}; // ... and the first launch block of every position
	bool order_has_bands = false;
	std::vector<float> host_spheres;
	unsigned long long scene_gen = 0; // hash of the uploaded spheres and camera (an e2e loop re-uploads the same scene every frame)
	struct OrderKey
	{
		unsigned long long gen;
		int width, height, tile, rank, world, rows_per_band;
		float fov;
	} order_key{};
	bool order_valid = false;
	float4 *d_cand_d = nullptr; // deferred triangle query: candidates (tri_deferred_kernel)
	uint2 *d_cand_px = nullptr;
	size_t cand_d_bytes = 0, cand_px_bytes = 0;
	unsigned *d_cursor = nullptr;			  // device counters of the deferred triangle query (candidates, CTAs done, fetch cursor); self-resetting
	int *d_err = nullptr;
	int *h_err = nullptr; // pinned

	// timing
	std::vector<Span> spans;
	size_t spans_used = 0;
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_x0 = nullptr, ev_x1 = nullptr;

	// overlapped copy-out of skr_render (see FrameParams::band_flag)
	cudaStream_t copy_stream = nullptr;
	unsigned *d_band = nullptr; // [0, MAX_BANDS): counts, [MAX_BANDS, 2 MAX_BANDS): flags
	unsigned band_seq = 0;
	unsigned long long band_geom = 0; // geometry the band counters were last reset for
	CUresult (*wait_value32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;
	CUresult (*write_value32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;

	unsigned launches = 0, chunks = 0;
	bool async_pending = false; // a fire-and-forget frame was enqueued since the error word was last read
	bool timing = true; // false: fire-and-forget frame, no per-kernel events
	unsigned long long queue_entries = 0;
};

namespace
{
int fail(skr_ctx *ctx, int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if(ctx)
	{
		ctx->err = buf;
	}
	else
	{
		g_init_error = buf;
	}
	return code;
}

#define CK(call)                                                                                                    \
	do                                                                                                               \
	{                                                                                                                \
		cudaError_t e_ = (call);                                                                                     \
		if(e_ != cudaSuccess)                                                                                        \
		{                                                                                                            \
			return fail(ctx, SKR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));        \
		}                                                                                                            \
	} while(0)

template <typename T>
cudaError_t ensure(T *&ptr, size_t &have, size_t want)
{
	if(have >= want && ptr)
	{
		return cudaSuccess;
	}
	if(ptr)
	{
		cudaFree(ptr);
		ptr = nullptr;
		have = 0;
	}
	cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ptr), want);
	if(e == cudaSuccess)
	{
		have = want;
	}
	return e;
}

struct V3
{
	float x, y, z;
};
inline V3 ld3(const float *p) { return V3{p[0], p[1], p[2]}; }
inline float hdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

void span_begin(skr_ctx *ctx, int cat)
{
	if(!ctx->timing)
	{
		return;
	}
	if(ctx->spans_used == ctx->spans.size())
	{
		Span s;
		cudaEventCreate(&s.a);
		cudaEventCreate(&s.b);
		s.cat = cat;
		ctx->spans.push_back(s);
	}
	Span &s = ctx->spans[ctx->spans_used];
	s.cat	= cat;
	cudaEventRecord(s.a, ctx->stream);
}
void span_end(skr_ctx *ctx)
{
	if(!ctx->timing)
	{
		return;
	}
	cudaEventRecord(ctx->spans[ctx->spans_used].b, ctx->stream);
	ctx->spans_used++;
}

// ---- kernel dispatch over the template flags -----------------------------------------------------
// SMEM variants exist for every (TRIS, FOG); the global-memory fallback (scene blob > 64 KB) only as the general one.
template <bool GI, bool STATS>
void launch_primary(skr_ctx *ctx, const FrameParams &fp, const Queue &q, long long lp0, long long n)
{
	const SceneView &sv = ctx->sv;
	cudaStream_t st		= ctx->stream;
	const size_t sm		= ctx->smem_bytes;
	// one 8 x 4 pixel block per warp, four warps per CTA.  SKR_PRIMARY_BLOCK overrides for A/B runs: measured on B200, CTAs of
	// 1 / 2 / 4 warps: config 1 0.053 / 0.051 / 0.049 ms, config 2 1.034 / 1.025 / 1.026, config 4 0.310 / 0.308 / 0.303.
	int threads = SKR_BLOCK;
	if(const char *e = getenv("SKR_PRIMARY_BLOCK"))
	{
		const int v = atoi(e);
		if(v == 32 || v == 64 || v == 128)
		{
			threads = v;
		}
	}
	const long long wpc		= threads / 32;
	const unsigned blocks = (unsigned) (((n + 31) / 32 * ((!GI && fp.split) ? 2 : 1) + wpc - 1) / wpc);
	const bool halves = !GI && fp.spp >= 8; // two-half sample sums (and the split over two warps): frames with samples to split
	FrameParams fpl = fp;
	if(threads != SKR_BLOCK)
	{
		fpl.strip_cta = 0; // (write_strip counts on four warps per CTA)
	}
	const auto go = [&](auto kernel, size_t bytes) { kernel<<<blocks, threads, bytes, st>>>(sv, fpl, q, lp0, n); };
	if(!sv.blob_in_smem)
	{
		halves ? go(primary_kernel<GI, STATS, false, true, true, !GI>, 0) : go(primary_kernel<GI, STATS, false, true, true, false>, 0);
		return;
	}
	const bool tris = sv.T > 0, fog = sv.F > 0;
	if(tris && fog)
	{
		halves ? go(primary_kernel<GI, STATS, true, true, true, !GI>, sm) : go(primary_kernel<GI, STATS, true, true, true, false>, sm);
	}
	else if(tris)
	{
		halves ? go(primary_kernel<GI, STATS, true, true, false, !GI>, sm) : go(primary_kernel<GI, STATS, true, true, false, false>, sm);
	}
	else if(fog)
	{
		halves ? go(primary_kernel<GI, STATS, true, false, true, !GI>, sm) : go(primary_kernel<GI, STATS, true, false, true, false>, sm);
	}
	else
	{
		halves ? go(primary_kernel<GI, STATS, true, false, false, !GI>, sm) : go(primary_kernel<GI, STATS, true, false, false, false>, sm);
	}
}

template <bool STATS, bool LEAF>
void launch_shade_expand_v(skr_ctx *ctx, unsigned blocks, const FrameParams &fp, const Queue &in, unsigned start, unsigned count, const Queue &out,
						   int expand)
{
	const SceneView &sv = ctx->sv;
	cudaStream_t st		= ctx->stream;
	const size_t stage	= LEAF ? (size_t) SKR_LEAF_CTA_BYTES : 0;
	const size_t sm		= ctx->smem_bytes + stage;
	if(!sv.blob_in_smem)
	{
		shade_expand_kernel<STATS, false, true, true, LEAF><<<blocks, SKR_BLOCK, stage, st>>>(sv, fp, in, start, count, out, expand);
		return;
	}
	const bool tris = sv.T > 0, fog = sv.F > 0;
	if(tris && fog)
	{
		shade_expand_kernel<STATS, true, true, true, LEAF><<<blocks, SKR_BLOCK, sm, st>>>(sv, fp, in, start, count, out, expand);
	}
	else if(tris)
	{
		shade_expand_kernel<STATS, true, true, false, LEAF><<<blocks, SKR_BLOCK, sm, st>>>(sv, fp, in, start, count, out, expand);
	}
	else if(fog)
	{
		shade_expand_kernel<STATS, true, false, true, LEAF><<<blocks, SKR_BLOCK, sm, st>>>(sv, fp, in, start, count, out, expand);
	}
	else
	{
		shade_expand_kernel<STATS, true, false, false, LEAF><<<blocks, SKR_BLOCK, sm, st>>>(sv, fp, in, start, count, out, expand);
	}
}
// expand: 0 = shade only, 1 = shade + push the children to `out`, 2 = shade + shade the (leaf) children in place
template <bool STATS>
void launch_shade_expand(skr_ctx *ctx, unsigned blocks, const FrameParams &fp, const Queue &in, unsigned start, unsigned count, const Queue &out,
						 int expand)
{
	if(expand == 2)
	{
		launch_shade_expand_v<STATS, true>(ctx, blocks, fp, in, start, count, out, 1);
	}
	else
	{
		launch_shade_expand_v<STATS, false>(ctx, blocks, fp, in, start, count, out, expand);
	}
}

template <bool GI, bool STATS, bool TRIS, bool FOG>
cudaError_t smem_attr_one(int bytes)
{
	cudaError_t e = cudaFuncSetAttribute(primary_kernel<GI, STATS, true, TRIS, FOG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	if(e == cudaSuccess && !GI)
	{
		e = cudaFuncSetAttribute(primary_kernel<GI, STATS, true, TRIS, FOG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	}
	if(e == cudaSuccess && GI)
	{
		e = cudaFuncSetAttribute(shade_expand_kernel<STATS, true, TRIS, FOG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	}
	if(e == cudaSuccess && GI)
	{
		e = cudaFuncSetAttribute(shade_expand_kernel<STATS, true, TRIS, FOG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + SKR_LEAF_CTA_BYTES);
	}
	if(e == cudaSuccess && GI && TRIS && FOG) // once per STATS: these have no (TRIS, FOG) variants
	{
		e = cudaFuncSetAttribute(fresnel_expand_kernel<STATS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
		if(e == cudaSuccess)
		{
			e = cudaFuncSetAttribute(shaded_tris_kernel<STATS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
		}
	}
	return e;
}

int set_smem_attr(skr_ctx *ctx)
{
	const int bytes = (int) ctx->smem_bytes;
	if(bytes + SKR_LEAF_CTA_BYTES > 48 * 1024)
	{
		CK((smem_attr_one<false, false, false, false>(bytes)));
		CK((smem_attr_one<false, false, false, true>(bytes)));
		CK((smem_attr_one<false, false, true, false>(bytes)));
		CK((smem_attr_one<false, false, true, true>(bytes)));
		CK((smem_attr_one<false, true, false, false>(bytes)));
		CK((smem_attr_one<false, true, false, true>(bytes)));
		CK((smem_attr_one<false, true, true, false>(bytes)));
		CK((smem_attr_one<false, true, true, true>(bytes)));
		CK((smem_attr_one<true, false, false, false>(bytes)));
		CK((smem_attr_one<true, false, false, true>(bytes)));
		CK((smem_attr_one<true, false, true, false>(bytes)));
		CK((smem_attr_one<true, false, true, true>(bytes)));
		CK((smem_attr_one<true, true, false, false>(bytes)));
		CK((smem_attr_one<true, true, false, true>(bytes)));
		CK((smem_attr_one<true, true, true, false>(bytes)));
		CK((smem_attr_one<true, true, true, true>(bytes)));
	}
	return SKR_OK;
}

// Builds `set` over the ctx->d_tris_raw triangles: mirrored (reference query) or actual (shaded-triangles mode).
// Leaves the count of outsized triangles in ctx->h_count once the stream has drained.
int build_bvh(skr_ctx *ctx, int T, BvhSet &set, bool mirror)
{
	using namespace bvhb;
	cudaStream_t st = ctx->stream;
	const int B		= 256;
	const int gridT = (T + B - 1) / B;
	set.valid		= false;
	set.bvh			= nullptr;
	// persistent buffers grow on demand and are reused by later uploads (an e2e loop re-uploads the same scene)
	CK(ensure(set.d_tri_v, set.tri_v_bytes, sizeof(float4) * 4 * (size_t) T));
	const char *nobvh = getenv("SKR_NO_BVH");
	// a handful of triangles (spheres1.scn has two): testing them all costs less per ray than a node visit, and the
	// ~27 dependent launches of the build would dominate the upload (0.21 ms against 0.03 ms)
	if((nobvh && nobvh[0] == '1') || T <= BRUTE_FORCE_TRIS)
	{
		iota_tris_kernel<<<gridT, B, 0, st>>>(ctx->d_tris_raw, T, set.d_tri_v);
		CK(cudaGetLastError());
		*ctx->h_count = 0;
		set.valid	  = true;
		return SKR_OK;
	}
	if(!set.d_big)
	{
		CK(cudaMalloc(&set.d_big, sizeof(float4) * 3 * BIG_TRI_CAP + 256));
	}
	int *big_count	  = reinterpret_cast<int *>(set.d_big + 3 * BIG_TRI_CAP);
	const int big_cap = T >= BIG_TRI_MIN_T ? BIG_TRI_CAP : 0;
	CK(ensure(set.d_bvh, set.bvh_bytes, sizeof(float4) * 4 * (size_t) (T - 1)));
	const int ipw	  = sort_items_per_warp(T);
	const int nwarps  = (T + ipw - 1) / ipw;
	const int sblocks = (nwarps + SORT_WARPS - 1) / SORT_WARPS;

	// build scratch: one arena, carved with 256-byte alignment
	size_t off = 0;
	auto carve = [&off](size_t bytes) {
		const size_t at = off;
		off += (bytes + 255) & ~(size_t) 255;
		return at;
	};
	const size_t o_box_lo = carve(sizeof(float4) * T), o_box_hi = carve(sizeof(float4) * T);
	const size_t o_node_lo = carve(sizeof(float4) * T), o_node_hi = carve(sizeof(float4) * T);
	const size_t o_scene = carve(sizeof(float) * 6);
	const size_t o_k0 = carve(sizeof(unsigned long long) * T), o_k1 = carve(sizeof(unsigned long long) * T);
	const size_t o_v0 = carve(sizeof(unsigned) * T), o_v1 = carve(sizeof(unsigned) * T);
	const size_t o_hist = carve(sizeof(unsigned) * 256 * (size_t) nwarps);
	const size_t o_children = carve(sizeof(int2) * T), o_parent = carve(sizeof(int) * 2 * (size_t) T), o_flags = carve(sizeof(int) * T);
	CK(ensure(ctx->d_scratch, ctx->scratch_bytes, off));
	char *base = ctx->d_scratch;
	float4 *box_lo = (float4 *) (base + o_box_lo), *box_hi = (float4 *) (base + o_box_hi);
	float4 *node_lo = (float4 *) (base + o_node_lo), *node_hi = (float4 *) (base + o_node_hi);
	float *scene_box = (float *) (base + o_scene);
	unsigned long long *keys[2] = {(unsigned long long *) (base + o_k0), (unsigned long long *) (base + o_k1)};
	unsigned *vals[2] = {(unsigned *) (base + o_v0), (unsigned *) (base + o_v1)};
	unsigned *hist = (unsigned *) (base + o_hist);
	int2 *children = (int2 *) (base + o_children);
	int *parent = (int *) (base + o_parent), *flags = (int *) (base + o_flags);

	// The build is ~30 small dependent launches: launch-bound.  It is captured once into a CUDA graph and replayed while
	// the triangle count and the buffers stay the same (an e2e loop re-uploads the same scene every frame).
	const BvhGraphKey key{T, ctx->d_tris_raw, set.d_tri_v, set.d_bvh, ctx->d_scratch};
	const char *nograph = getenv("SKR_NO_GRAPH");
	const bool use_graph = !(nograph && nograph[0] == '1');
	if(use_graph && set.graph && memcmp(&key, &set.graph_key, sizeof key) == 0)
	{
		CK(cudaGraphLaunch(set.graph, st));
		CK(cudaMemcpyAsync(ctx->h_count, big_count, sizeof(int), cudaMemcpyDeviceToHost, st)); // read after the caller's sync
		set.bvh	  = set.d_bvh;
		set.valid = true;
		return SKR_OK;
	}
	if(set.graph)
	{
		cudaGraphExecDestroy(set.graph);
		set.graph = nullptr;
	}
	const auto enqueue_build = [&]() {
		init_scene_box_kernel<<<1, 32, 0, st>>>(scene_box);
		cudaMemsetAsync(big_count, 0, sizeof(int), st);
		tri_bounds_kernel<<<gridT, B, 0, st>>>(ctx->d_tris_raw, T, box_lo, box_hi, scene_box, mirror ? 1 : 0);
		morton_kernel<<<gridT, B, 0, st>>>(box_lo, box_hi, scene_box, T, keys[0], vals[0], ctx->d_tris_raw, big_count, set.d_big, big_cap, mirror ? SKR_GCLASS_BITS : 0);
		int cur = 0;
		for(int pass = 0; pass < 8; pass++)
		{
			const int shift = 8 * pass;
			sort_hist_kernel<<<sblocks, SORT_THREADS, 0, st>>>(keys[cur], T, shift, hist, nwarps, ipw);
			sort_scan_kernel<<<1, 1024, 0, st>>>(hist, 256 * nwarps);
			sort_scatter_kernel<<<sblocks, SORT_THREADS, 0, st>>>(keys[cur], vals[cur], T, shift, hist, nwarps, keys[cur ^ 1], vals[cur ^ 1], ipw);
			cur ^= 1;
		}
		cudaMemsetAsync(flags, 0, sizeof(int) * T, st);
		karras_kernel<<<gridT, B, 0, st>>>(keys[cur], T, children, parent);
		refit_kernel<<<gridT, B, 0, st>>>(T, vals[cur], box_lo, box_hi, children, parent, node_lo, node_hi, flags, set.d_bvh);
		gather_tris_kernel<<<gridT, B, 0, st>>>(ctx->d_tris_raw, vals[cur], T, set.d_tri_v);
	};
	bool launched = false;
	if(use_graph && !set.graph_broken && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
	{
		enqueue_build();
		cudaGraph_t g = nullptr;
		if(cudaStreamEndCapture(st, &g) == cudaSuccess && g && cudaGraphInstantiate(&set.graph, g, 0) == cudaSuccess &&
		   cudaGraphLaunch(set.graph, st) == cudaSuccess)
		{
			set.graph_key = key;
			launched	  = true;
		}
		else
		{
			// capture / instantiation not possible here: launch the kernels one by one, now and from now on
			if(set.graph)
			{
				cudaGraphExecDestroy(set.graph);
				set.graph = nullptr;
			}
			set.graph_broken = true;
			cudaGetLastError();
		}
		if(g)
		{
			cudaGraphDestroy(g);
		}
	}
	if(!launched)
	{
		enqueue_build();
	}
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(ctx->h_count, big_count, sizeof(int), cudaMemcpyDeviceToHost, st)); // read after the caller's sync
	set.bvh	  = set.d_bvh;
	set.valid = true;
	return SKR_OK;
}

// The hierarchy over the ACTUAL triangles (shaded-triangles mode), built the first time a frame asks for it.
int ensure_shade_bvh(skr_ctx *ctx)
{
	if(ctx->n_tris <= 0 || ctx->bvh_shade.valid)
	{
		return SKR_OK;
	}
	int rc = build_bvh(ctx, ctx->n_tris, ctx->bvh_shade, false);
	if(rc)
	{
		return rc;
	}
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->sv.tri_v2 = ctx->bvh_shade.d_tri_v;
	ctx->sv.bvh2   = ctx->bvh_shade.bvh;
	ctx->sv.big_v2 = ctx->bvh_shade.d_big;
	ctx->sv.nbig2  = ctx->bvh_shade.bvh ? std::min((int) *ctx->h_count, BIG_TRI_CAP) : 0;
	return SKR_OK;
}

// Kernel functions of the variants a scene's frames launch (skr_reserve loads them ahead of the first frame: CUDA loads a
// kernel's code lazily at its first launch, a few ms each).
template <bool STATS>
void frame_kernels(const skr_ctx *ctx, bool tree, bool leaf, bool fresnel, std::vector<const void *> &out)
{
	const SceneView &sv = ctx->sv;
	const bool tris = sv.T > 0, fog = sv.F > 0;
#define SKR_PICK(K, ...)                                                                                                      \
	(!sv.blob_in_smem ? (const void *) K<__VA_ARGS__, false, true, true SKR_TAIL> :                                           \
	 tris && fog	   ? (const void *) K<__VA_ARGS__, true, true, true SKR_TAIL> :                                            \
	 tris			   ? (const void *) K<__VA_ARGS__, true, true, false SKR_TAIL> :                                           \
	 fog			   ? (const void *) K<__VA_ARGS__, true, false, true SKR_TAIL> :                                           \
						 (const void *) K<__VA_ARGS__, true, false, false SKR_TAIL>)
#define SKR_TAIL , false
	out.push_back(tree ? SKR_PICK(primary_kernel, true, STATS) : SKR_PICK(primary_kernel, false, STATS));
#undef SKR_TAIL
	if(!tree)
	{
#define SKR_TAIL , true
		out.push_back(SKR_PICK(primary_kernel, false, STATS));
#undef SKR_TAIL
	}
	if(tree)
	{
#define SKR_TAIL , false
		out.push_back(SKR_PICK(shade_expand_kernel, STATS));
#undef SKR_TAIL
		if(leaf)
		{
#define SKR_TAIL , true
			out.push_back(SKR_PICK(shade_expand_kernel, STATS));
#undef SKR_TAIL
		}
		if(fresnel)
		{
			out.push_back(sv.blob_in_smem ? (const void *) fresnel_expand_kernel<STATS, true> : (const void *) fresnel_expand_kernel<STATS, false>);
		}
		out.push_back((const void *) resolve_kernel);
	}
#undef SKR_PICK
}

struct Plan
{
	FrameParams fp;
	long long npix_local; // local pixels incl. padding (tiles_local * tile^2)
	long long tiles_local;
	int levels;		  // depth levels of the --gillum / fresnel tree (0: none)
	bool shaded;	  // opt-in shaded-triangles mode (skr_shaded.cuh)
	int rows_per_band; // tile rows per band of skr_render's overlapped copy-out (0: the frame does not leave in bands)
	int band_perm[MAX_BANDS]; // launch position -> band of rows (identity unless tile_launch_order reorders the bands)
	bool defer;		  // triangle queries of the camera rays go through tri_deferred_kernel
	int qlevels;	  // queue levels needed: `levels`, or one less when the leaves are shaded in place
	bool leaf_inline; // depth-1 hits are shaded by the warp that found them (shade_expand_kernel<..., LEAF>)
};

// constants of fdiv (skr_kernels.cuh): l = ceil(log2 d), m = floor(2^32 (2^l - d) / d) + 1, shifts min(l, 1) and max(l - 1, 0)
FastDiv make_fastdiv(uint32_t d)
{
	FastDiv f;
	if(d == 0)
	{
		d = 1;
	}
	uint32_t l = 0;
	while(l < 32 && (1ull << l) < d)
	{
		l++;
	}
	f.m	 = (uint32_t) ((((1ull << l) - d) << 32) / d + 1ull);
	f.s1 = l < 1 ? l : 1;
	f.s2 = l > 1 ? l - 1 : 0;
	return f;
}

int make_plan(skr_ctx *ctx, const skr_options *o, Plan &pl)
{
	if(!o || o->width <= 0 || o->height <= 0)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: width/height must be positive");
	}
	if(o->grid_size < 0 || o->grid_size > 255)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: grid_size (--jsample) must be in [0,255]");
	}
	if(o->monte_carlo && o->num_path_traces < 0)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: num_path_traces (--gillum) must be >= 0");
	}
	if((long long) o->width * o->height > 0x7fffffffLL)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: frame too large");
	}
	if(o->shade_triangles && (o->monte_carlo || o->fresnel))
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: shade_triangles (a non-parity extension) is not combined with monte_carlo / fresnel");
	}
	pl.shaded		= o->shade_triangles != 0;
	pl.rows_per_band = 0;
	const int world = o->world > 1 ? o->world : 1;
	const int rank	= o->world > 1 ? o->rank : 0;
	if(rank < 0 || rank >= world)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: rank %d outside world %d", rank, world);
	}
	const int tile = o->tile > 0 ? o->tile : DEFAULT_TILE;
	if(tile % 8 != 0 || tile > 1024)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_options: tile must be a multiple of 8 (got %d)", tile);
	}
	FrameParams &fp = pl.fp;
	memset(&fp, 0, sizeof fp);
	fp.width		= o->width;
	fp.height		= o->height;
	fp.tile			= tile;
	fp.tiles_x		= (o->width + tile - 1) / tile;
	const int tiles_y = (o->height + tile - 1) / tile;
	fp.tiles_total	= fp.tiles_x * tiles_y;
	fp.rank			= rank;
	fp.world		= world;
	fp.wpr			= tile / 8;
	fp.grid			= o->grid_size;
	fp.spp			= o->grid_size > 0 ? o->grid_size * o->grid_size : 1;
	fp.max_depth	= o->max_depth;
	fp.gi			= o->monte_carlo ? 1 : 0;
	fp.n_gi			= o->monte_carlo ? o->num_path_traces : 0;
	fp.shadows		= o->use_shadows ? 1 : 0;
	fp.fresnel		= o->fresnel ? 1 : 0;
	// src/main.cpp:40-43
	fp.inv_w  = 1 / float(o->width);
	fp.inv_h  = 1 / float(o->height);
	fp.aspect = o->width / float(o->height);
	fp.angle  = (float) tan(M_PI * 0.5 * o->fov / 180.);
	fp.key	  = make_uint2((uint32_t) o->seed, (uint32_t) (o->seed >> 32));
	{
		// bundle culling (skr_device.cuh: cull_pairs).  A pixel's rays are d(r) = D + u(r) R + v(r) U with ONE draw r in
		// [0, 1] for both axes (src/main.cpp:52-54): |d(r) - d(0.5)| <= 0.5 |du R - dv U| + rounding of the float chain.
		const SceneView &sv = ctx->sv;
		const double du = 2.0 * fp.angle * fp.aspect * fp.inv_w, dv = 2.0 * fp.angle * fp.inv_h;
		const double ex = du * sv.cam_right.x - dv * sv.cam_up.x, ey = du * sv.cam_right.y - dv * sv.cam_up.y, ez = du * sv.cam_right.z - dv * sv.cam_up.z;
		const auto len3 = [](float3 v) { return sqrt((double) v.x * v.x + (double) v.y * v.y + (double) v.z * v.z); };
		const double span = len3(sv.cam_dir) + fabs((double) fp.angle * fp.aspect) * len3(sv.cam_right) + fabs((double) fp.angle) * len3(sv.cam_up);
		fp.cull_delta = (float) (0.5 * 1.01 * sqrt(ex * ex + ey * ey + ez * ez) + 4e-6 * span);
		const char *nocull = getenv("SKR_NO_CULL");
		fp.cull = (sv.off_cull >= 0 && fp.grid > 0 && fp.spp >= 4 && !(nocull && nocull[0] == '1') && std::isfinite(fp.cull_delta)) ? 1 : 0;
	}
	fp.node_base = (uint32_t) fp.n_gi + 1u + (fp.fresnel ? 2u * (uint32_t) (ctx->sv.L + ctx->sv.D) : 0u);
	fp.slot_gi	 = 1u + (uint32_t) ctx->sv.L * (uint32_t) ctx->sv.F;
	pl.tiles_local = (fp.tiles_total + world - 1) / world;
	pl.npix_local  = pl.tiles_local * tile * tile;
	fp.fd_tpix		= make_fastdiv((uint32_t) (tile * tile));
	fp.fd_wpr		= make_fastdiv((uint32_t) fp.wpr);
	fp.fd_tiles_x	= make_fastdiv((uint32_t) fp.tiles_x);
	fp.fd_tile		= make_fastdiv((uint32_t) tile);
	fp.fd_world		= make_fastdiv((uint32_t) world);
	fp.fd_width		= make_fastdiv((uint32_t) o->width);
	fp.fd_peer_rows = make_fastdiv(1u);
	fp.fast			= pl.npix_local <= 0xffffffffLL && (long long) fp.tiles_total + world <= 0x7fffffffLL;
	pl.levels	   = ((fp.gi || fp.fresnel) && fp.max_depth > 0) ? fp.max_depth : 0;
	{
		// deferred triangle query (tri_deferred_kernel: teams of lanes per heavy ray): single-sample frames without a wavefront
		// tree, over a real hierarchy.  OPT-IN (SKR_DEFER=1) since the hierarchy is built over cubic Morton cells: with the
		// slab-shaped top levels of the first build, config 4 had lines of 640 node visits and the teams won (0.218 ms against
		// 0.279 for the walk in place); now (4.7 M visits per frame instead of 21 M) the walk in place takes 0.124 ms, teams of
		// 8 / 4 lanes 0.157 / 0.128-0.136.  Same frame either way.
		const char *yes = getenv("SKR_DEFER");
		pl.defer		= !pl.shaded && !fp.gi && !fp.fresnel && fp.spp == 1 && fp.max_depth > 0 && ctx->sv.bvh != nullptr && !ctx->sv.bvh_root_is_leaf &&
				   pl.npix_local <= 0xffffffffLL && yes && yes[0] == '1';
	}
	{
		// leaves in place: plain --gillum trees (the fresnel pass pushes its own leaf children through the queue)
		// Worth it where shading a leaf is expensive against staging it, i.e. with shadow rays (config 5, 31 spheres x 2
		// lights: 152 -> 145 ms); without them the queued consumer wins (config 3: 6.1 ms against 7.1).  SKR_LEAF_INLINE=1 /
		// SKR_NO_LEAF_INLINE=1 force either path (the frame is bit-identical both ways).
		const char *no = getenv("SKR_NO_LEAF_INLINE"), *yes = getenv("SKR_LEAF_INLINE");
		const bool pays = fp.shadows != 0 || (yes && yes[0] == '1');
		pl.leaf_inline	= pl.levels >= 2 && fp.gi && fp.n_gi > 0 && fp.n_gi <= SKR_LEAF_MAX_CHILDREN && !fp.fresnel && pays && !(no && no[0] == '1');
		pl.qlevels	   = pl.leaf_inline ? pl.levels - 1 : pl.levels;
	}
	{
		// Philox counters number the nodes of a sample's tree as child = parent * node_base + c + 1 in 32 bits: the ids of
		// level k stay below node_base^k, so the deepest level (max_depth - 1) must fit.  (The reference would trace
		// node_base^(depth-1) rays per sample there: hours per pixel long before this limit.)
		double top = 1.0;
		for(int k = 1; k < pl.levels; k++)
		{
			top *= (double) fp.node_base;
		}
		if(top > 4294967295.0)
		{
			return fail(ctx, SKR_ERR_ARG, "skr_options: a tree of %u children per hit and depth %d needs more than 32 bits of node ids (%.3g nodes per sample)",
						fp.node_base - 1u, fp.max_depth, top);
		}
	}
	return SKR_OK;
}

int ensure_queues(skr_ctx *ctx, int levels, unsigned cap)
{
	if(levels <= ctx->n_levels_alloc && cap == ctx->queue_cap)
	{
		return SKR_OK;
	}
	// one allocation for all levels; an arena that is already big enough is re-carved, not re-allocated
	const auto up = [](size_t b) { return (b + 255) & ~(size_t) 255; };
	const size_t per_level = up(sizeof(float4) * (size_t) cap) * 3 + up(sizeof(uint32_t) * (size_t) cap);
	const size_t head	   = up(sizeof(unsigned) * (size_t) (levels + 1));
	const size_t want	   = head + per_level * (size_t) levels;
	ctx->queues.clear();
	ctx->n_levels_alloc = 0;
	ctx->queue_cap		= 0;
	if(want > ctx->arena_bytes)
	{
		if(ctx->d_arena)
		{
			cudaFree(ctx->d_arena);
			ctx->d_arena	 = nullptr;
			ctx->arena_bytes = 0;
		}
		CK(cudaMalloc(&ctx->d_arena, want));
		ctx->arena_bytes = want;
	}
	ctx->d_counts = reinterpret_cast<unsigned *>(ctx->d_arena);
	char *at	  = ctx->d_arena + head;
	for(int l = 0; l < levels; l++)
	{
		Queue q{};
		q.a = reinterpret_cast<float4 *>(at), at += up(sizeof(float4) * (size_t) cap);
		q.b = reinterpret_cast<float4 *>(at), at += up(sizeof(float4) * (size_t) cap);
		q.d = reinterpret_cast<float4 *>(at), at += up(sizeof(float4) * (size_t) cap);
		q.c = reinterpret_cast<uint32_t *>(at), at += up(sizeof(uint32_t) * (size_t) cap);
		q.count = ctx->d_counts + l;
		q.cap	= cap;
		ctx->queues.push_back(q);
	}
	ctx->n_levels_alloc = levels;
	ctx->queue_cap		= cap;
	return SKR_OK;
}

int read_count(skr_ctx *ctx, int level, unsigned &out)
{
	CK(cudaMemcpyAsync(ctx->h_count, ctx->d_counts + level, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	out = *ctx->h_count;
	return SKR_OK;
}

Queue queue_view(const skr_ctx *ctx, int level)
{
	return ctx->queues[level];
}

template <bool STATS>
void launch_fresnel_expand(skr_ctx *ctx, unsigned blocks, const FrameParams &fp, const Queue &in, unsigned start, unsigned count, const Queue &out)
{
	if(ctx->sv.blob_in_smem)
	{
		fresnel_expand_kernel<STATS, true><<<blocks, SKR_BLOCK, ctx->smem_bytes, ctx->stream>>>(ctx->sv, fp, in, start, count, out);
	}
	else
	{
		fresnel_expand_kernel<STATS, false><<<blocks, SKR_BLOCK, 0, ctx->stream>>>(ctx->sv, fp, in, start, count, out);
	}
}

// Depth-first over chunks of one queue level.  `depth` is the shade() depth of the hits in this level.  A hit spawns
// children only if the reference would trace them (depth - 1 >= 1): n_gi hemisphere children under --gillum and, in
// fresnel mode, up to 1 refraction + one reflection per light.  Chunks are sized so that the worst-case fan-out fits
// the next level's queue.
template <bool STATS>
int process_level(skr_ctx *ctx, const Plan &pl, int level, unsigned count, int depth)
{
	if(count == 0)
	{
		return SKR_OK;
	}
	const FrameParams &fp = pl.fp;
	const bool deeper	  = depth - 1 >= 1;
	const bool expand_gi  = deeper && fp.gi && fp.n_gi > 0;
	const bool expand_fr  = deeper && fp.fresnel && (ctx->sv.L + ctx->sv.D) > 0;
	const unsigned fan	  = (expand_gi ? (unsigned) fp.n_gi : 0u) + (expand_fr ? 1u + (unsigned) (ctx->sv.L + ctx->sv.D) : 0u);
	const Queue in		  = queue_view(ctx, level);
	ctx->queue_entries += count;
	const bool leaves_here = pl.leaf_inline && depth == 2; // the children of these hits are leaves: shaded in place
	if(fan == 0 || leaves_here)
	{
		span_begin(ctx, CAT_BOUNCE);
		launch_shade_expand<STATS>(ctx, (count + SKR_BLOCK - 1) / SKR_BLOCK, fp, in, 0u, count, in, leaves_here ? 2 : 0);
		span_end(ctx);
		ctx->launches++;
		CK(cudaGetLastError());
		return SKR_OK;
	}
	const Queue out		 = queue_view(ctx, level + 1);
	const unsigned chunk = out.cap / fan;
	for(unsigned s = 0; s < count; s += chunk)
	{
		const unsigned m = count - s < chunk ? count - s : chunk;
		CK(cudaMemsetAsync(out.count, 0, sizeof(unsigned), ctx->stream));
		span_begin(ctx, CAT_BOUNCE);
		launch_shade_expand<STATS>(ctx, (m + SKR_BLOCK - 1) / SKR_BLOCK, fp, in, s, m, out, expand_gi ? 1 : 0);
		ctx->launches++;
		if(expand_fr)
		{
			launch_fresnel_expand<STATS>(ctx, (m + SKR_BLOCK - 1) / SKR_BLOCK, fp, in, s, m, out);
			ctx->launches++;
		}
		span_end(ctx);
		ctx->chunks++;
		CK(cudaGetLastError());
		unsigned next = 0;
		int rc		  = read_count(ctx, level + 1, next);
		if(rc)
		{
			return rc;
		}
		rc = process_level<STATS>(ctx, pl, level + 1, next, depth - 1);
		if(rc)
		{
			return rc;
		}
	}
	return SKR_OK;
}

// Everything a frame allocates, sized and allocated BEFORE its timed span (and by skr_reserve ahead of the first frame):
// the queue arena and the accumulators of a --gillum / fresnel tree.  An allocation that is already big enough is kept.
int prepare_frame(skr_ctx *ctx, const skr_options *o, Plan &pl)
{
	FrameParams &fp = pl.fp;
	if(pl.shaded)
	{
		return ensure_shade_bvh(ctx);
	}
	if(pl.defer)
	{
		CK(ensure(ctx->d_cand_d, ctx->cand_d_bytes, sizeof(float4) * (size_t) pl.npix_local));
		CK(ensure(ctx->d_cand_px, ctx->cand_px_bytes, sizeof(uint2) * (size_t) pl.npix_local));
		fp.cand_d	  = ctx->d_cand_d;
		fp.cand_px	  = ctx->d_cand_px;
		fp.cand_count = ctx->d_cursor;
		fp.defer	  = 1;
		return SKR_OK;
	}
	if(!((fp.gi || fp.fresnel) && pl.levels > 0))
	{
		return SKR_OK;
	}
	const unsigned fan = (unsigned) fp.n_gi + (fp.fresnel ? 1u + (unsigned) (ctx->sv.L + ctx->sv.D) : 0u);
	unsigned cap	   = (unsigned) o->queue_capacity;
	if(o->queue_capacity <= 0)
	{
		// default: four times what one fan-out of all primary samples could need, between 32 M and 128 M entries (1.7 - 6.7 GB
		// per level of 180), and never more than a quarter of the free memory over all levels; an allocation that is
		// already big enough is kept.  Generous on purpose: a level that fits its queue is ONE launch and one count read-back
		// (config 3: 19 launches / 6.13 ms at 33 M entries, 8 launches / 5.86 ms at 128 M; one rank's share at world = 8:
		// 12 launches / 0.99 ms at 8 M, 6 launches / 0.85 ms at 32 M).
		const unsigned long long want = 4ull * (unsigned long long) pl.npix_local * (unsigned) fp.spp * (fan ? fan : 1u);
		cap = want > MAX_QUEUE_CAP ? MAX_QUEUE_CAP : (want < MIN_QUEUE_CAP ? MIN_QUEUE_CAP : (unsigned) want);
		if(ctx->queue_cap >= cap && ctx->n_levels_alloc >= pl.qlevels)
		{
			cap = ctx->queue_cap;
		}
		else
		{
			size_t free_b = 0, total_b = 0;
			if(cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
			{
				const size_t budget = (free_b + ctx->arena_bytes) / 4 / (size_t) pl.qlevels / 52;
				if(budget < cap)
				{
					cap = (unsigned) budget;
				}
			}
		}
	}
	const unsigned need = fan > (unsigned) fp.spp ? fan : (unsigned) fp.spp;
	if(cap < need * SKR_BLOCK)
	{
		cap = need * SKR_BLOCK;
	}
	int rc = ensure_queues(ctx, pl.qlevels, cap);
	if(rc)
	{
		return rc;
	}
	CK(ensure(ctx->d_accum, ctx->accum_bytes, sizeof(long long) * SKR_ACC_STRIDE * (size_t) pl.npix_local));
	fp.accum = ctx->d_accum;
	return SKR_OK;
}

template <bool STATS>
int render_frame(skr_ctx *ctx, const skr_options *o, Plan &pl)
{
	FrameParams &fp = pl.fp;
	cudaStream_t st = ctx->stream;
	const bool tree = fp.gi || fp.fresnel;
	if(pl.shaded)
	{
		span_begin(ctx, CAT_PRIMARY);
		const unsigned blocks = (unsigned) ((pl.npix_local + SKR_BLOCK - 1) / SKR_BLOCK);
		if(ctx->sv.blob_in_smem)
		{
			shaded_tris_kernel<STATS, true><<<blocks, SKR_BLOCK, ctx->smem_bytes, st>>>(ctx->sv, fp, pl.npix_local);
		}
		else
		{
			shaded_tris_kernel<STATS, false><<<blocks, SKR_BLOCK, 0, st>>>(ctx->sv, fp, pl.npix_local);
		}
		span_end(ctx);
		ctx->launches++;
		CK(cudaGetLastError());
		return SKR_OK;
	}
	if(!tree || pl.levels == 0)
	{
		span_begin(ctx, CAT_PRIMARY);
		Queue none{};
		launch_primary<false, STATS>(ctx, fp, none, 0, pl.npix_local);
		ctx->launches++;
		if(pl.defer)
		{
			tri_deferred_kernel<STATS><<<(unsigned) ctx->sm_count * (unsigned) SKR_DEFER_CTAS, SKR_BLOCK, 0, st>>>(ctx->sv, fp);
			ctx->launches++;
		}
		span_end(ctx);
		CK(cudaGetLastError());
		return SKR_OK;
	}
	const Queue q0	  = queue_view(ctx, 0);
	long long batch	  = (long long) (q0.cap / (unsigned) fp.spp) / SKR_BLOCK * SKR_BLOCK;
	for(long long lp0 = 0; lp0 < pl.npix_local; lp0 += batch)
	{
		const long long n = pl.npix_local - lp0 < batch ? pl.npix_local - lp0 : batch;
		CK(cudaMemsetAsync(q0.count, 0, sizeof(unsigned), st));
		span_begin(ctx, CAT_PRIMARY);
		launch_primary<true, STATS>(ctx, fp, q0, lp0, n);
		span_end(ctx);
		ctx->launches++;
		ctx->chunks++;
		CK(cudaGetLastError());
		unsigned c0 = 0;
		int rc		= read_count(ctx, 0, c0);
		if(rc)
		{
			return rc;
		}
		rc = process_level<STATS>(ctx, pl, 0, c0, fp.max_depth);
		if(rc)
		{
			return rc;
		}
	}
	span_begin(ctx, CAT_RESOLVE);
	resolve_kernel<<<(unsigned) ((pl.npix_local + SKR_BLOCK - 1) / SKR_BLOCK), SKR_BLOCK, 0, st>>>(fp, 0, pl.npix_local);
	span_end(ctx);
	ctx->launches++;
	CK(cudaGetLastError());
	return SKR_OK;
}

// Heavy-first launch order.  A frame that is ONE kernel ends with the tail of its last CTAs; in scan order those are the
// bottom rows of the image -- the ground, the most expensive pixels of most scenes -- and with the frame split over 8 GPUs
// (two waves of CTAs each) that tail was a third of the step.  The host classifies this rank's tiles: can any line
// through a pixel of the tile pass a sphere's test?  (The bundle test of cull_pairs, skr_device.cuh, evaluated in double
// for the cone around the tile's centre ray.)  Tiles that can go first, sky tiles last.  When the frame leaves in bands of
// tile rows (skr_render's overlapped copy-out) the order is band by band -- a band's copy starts when its last block is
// done -- with the BANDS heaviest first too: the last band launched, whose copy cannot hide behind the kernel and whose
// tiles are the kernel's tail, is then the lightest one (pl.band_perm tells skr_render which rows complete k-th).
// Purely a permutation of the launch: every pixel computes what it computed before.
int tile_launch_order(skr_ctx *ctx, Plan &pl)
{
	FrameParams &fp			= pl.fp;
	const int rows_per_band = pl.rows_per_band;
	fp.tile_order			= nullptr;
	const char *no = getenv("SKR_NO_TILE_ORDER");
	if(ctx->sv.S == 0 || ctx->sv.T > 0 || (no && no[0] == '1') || pl.tiles_local < 8)
	{
		return SKR_OK; // nothing to tell apart (no spheres), or triangles anywhere: keep scan order
	}
	skr_ctx::OrderKey key;
	memset(&key, 0, sizeof key); // (compared with memcmp: padding included)
	key.gen = ctx->scene_gen, key.width = fp.width, key.height = fp.height, key.tile = fp.tile, key.rank = fp.rank, key.world = fp.world;
	key.rows_per_band = rows_per_band, key.fov = fp.angle;
	if(!ctx->order_valid || memcmp(&key, &ctx->order_key, sizeof key) != 0)
	{
		const SceneView &sv = ctx->sv;
		const auto ray		= [&](double x, double y, double *d) {
			 const double u = (2.0 * (x * fp.inv_w) - 1.0) * fp.angle * fp.aspect, v = (1.0 - 2.0 * (y * fp.inv_h)) * fp.angle;
			 d[0] = sv.cam_dir.x + u * sv.cam_right.x + v * sv.cam_up.x;
			 d[1] = sv.cam_dir.y + u * sv.cam_right.y + v * sv.cam_up.y;
			 d[2] = sv.cam_dir.z + u * sv.cam_right.z + v * sv.cam_up.z;
		};
		// per band of tile rows (one band when the frame does not leave in bands): tiles that can see a sphere, tiles that cannot
		const int tile_rows = (fp.height + fp.tile - 1) / fp.tile;
		const int nbands	= rows_per_band > 0 ? (tile_rows + rows_per_band - 1) / rows_per_band : 1;
		std::vector<std::vector<int>> heavy((size_t) nbands), light((size_t) nbands);
		std::vector<int> &order = ctx->tile_order_host;
		order.clear();
		for(long long lt = 0; lt < pl.tiles_local; lt++)
		{
			const long long gt = lt * fp.world + fp.rank;
			if(gt >= fp.tiles_total)
			{
				light[(size_t) nbands - 1].push_back(fp.tiles_total); // padding slot: decode_pixel marks it invalid
				continue;
			}
			const int tx = (int) (gt % fp.tiles_x), ty = (int) (gt / fp.tiles_x);
			const int band = rows_per_band > 0 ? ty / rows_per_band : 0;
			// cone of the tile: centre ray w, half angle beta from the farthest corner (+ one pixel for the jitter)
			const double x0 = tx * fp.tile - 1.0, x1 = std::min(fp.width, (tx + 1) * fp.tile) + 1.0;
			const double y0 = ty * fp.tile - 1.0, y1 = std::min(fp.height, (ty + 1) * fp.tile) + 1.0;
			double w[3], c[3];
			ray(0.5 * (x0 + x1), 0.5 * (y0 + y1), w);
			const double ww = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
			double sin2 = 0.0;
			for(int k = 0; k < 4; k++)
			{
				ray(k & 1 ? x1 : x0, k & 2 ? y1 : y0, c);
				const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2], cw = c[0] * w[0] + c[1] * w[1] + c[2] * w[2];
				sin2 = std::max(sin2, 1.0 - cw * cw / (cc * ww));
			}
			const double beta = 1.1 * asin(std::min(1.0, sqrt(std::max(0.0, sin2)))) + 1e-6;
			bool sees = !(ww > 0.0) || !(beta < 0.7); // degenerate camera / very wide tiles: call it heavy
			for(int s = 0; s < sv.S && !sees; s++)
			{
				const float *p	= ctx->host_spheres.data() + 18 * (size_t) s;
				const double ux = (double) p[0] - sv.cam_pos.x, uy = (double) p[1] - sv.cam_pos.y, uz = (double) p[2] - sv.cam_pos.z;
				const double uu = ux * ux + uy * uy + uz * uz, hw = ux * w[0] + uy * w[1] + uz * w[2];
				const double X	= fabs((double) p[3]) * 1.01 + sqrt(uu) * beta + 1e-4 * (1.0 + sqrt(uu));
				sees			= !(uu - hw * hw / ww > X * X);
			}
			(sees ? heavy : light)[(size_t) band].push_back((int) gt);
		}
		// Bands are launched heaviest first (share of tiles that see a sphere; ties in row order), so that the LAST band --
		// the one whose copy-out cannot hide behind the kernel, and whose tiles make the kernel's tail -- is the lightest
		std::vector<int> perm((size_t) nbands);
		for(int b = 0; b < nbands; b++)
		{
			perm[(size_t) b] = b;
		}
		std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
			const double na = (double) (heavy[(size_t) a].size() + light[(size_t) a].size()), nb = (double) (heavy[(size_t) b].size() + light[(size_t) b].size());
			return (double) heavy[(size_t) a].size() * nb > (double) heavy[(size_t) b].size() * na;
		});
		const unsigned blocks_per_tile = (unsigned) (fp.tile * fp.tile / 32);
		ctx->order_has_bands		   = rows_per_band > 0 && nbands <= MAX_BANDS;
		for(int k = 0; k < nbands; k++)
		{
			const int b = perm[(size_t) k];
			if(ctx->order_has_bands)
			{
				ctx->order_band_perm[k]	 = b;
				ctx->order_band_start[k] = (unsigned) order.size() * blocks_per_tile;
			}
			order.insert(order.end(), heavy[(size_t) b].begin(), heavy[(size_t) b].end());
			order.insert(order.end(), light[(size_t) b].begin(), light[(size_t) b].end());
		}
		if(ctx->order_has_bands)
		{
			ctx->order_band_start[nbands] = (unsigned) order.size() * blocks_per_tile;
		}
		CK(ensure(ctx->d_tile_order, ctx->tile_order_bytes, sizeof(int) * order.size()));
		CK(cudaMemcpyAsync(ctx->d_tile_order, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, ctx->stream));
		ctx->order_key	 = key;
		ctx->order_valid = true;
	}
	fp.tile_order = ctx->d_tile_order;
	if(rows_per_band > 0 && ctx->order_has_bands)
	{
		for(unsigned k = 0; k < fp.n_bands && k < (unsigned) MAX_BANDS; k++)
		{
			fp.band_start[k] = ctx->order_band_start[k];
			pl.band_perm[k]	 = ctx->order_band_perm[k];
		}
		fp.band_start[fp.n_bands] = ctx->order_band_start[fp.n_bands];
	}
	return SKR_OK;
}

// the device error word after the stream has drained (h_err holds a fresh copy): reported once, then cleared
int check_error_word(skr_ctx *ctx, const char *what)
{
	const int flag		= *ctx->h_err;
	const bool pending	= ctx->async_pending;
	ctx->async_pending	= false;
	if(!flag)
	{
		return SKR_OK;
	}
	*ctx->h_err = 0;
	cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream);
	return fail(ctx, SKR_ERR_CUDA, "internal: %s (flag %d) in %s%s", (flag & 2) ? "BVH traversal stack overflow" : "wavefront queue overflow", flag, what,
				pending ? " or in an asynchronous frame enqueued before it" : "");
}

// common driver: outputs already set in pl.fp
// `after_launch` (optional) runs once the frame's kernels are enqueued, before anything waits for them; `before_launch`
// (optional) once the frame is planned (launch order known), before anything is enqueued for it.
int render_common(skr_ctx *ctx, const skr_options *o, Plan &pl, skr_stats *stats, const std::function<int()> &after_launch = nullptr,
				  const std::function<int()> &before_launch = nullptr)
{
	cudaStream_t st = ctx->stream;
	ctx->spans_used = 0;
	ctx->launches = ctx->chunks = 0;
	ctx->queue_entries = 0;
	const bool want_stats = o->collect_stats != 0;
	const bool tree		  = (pl.fp.gi || pl.fp.fresnel) && pl.levels > 0;
	// Fire-and-forget: with no stats requested and no wavefront tree (whose scheduling reads queue counts back), the
	// frame is ONE kernel; enqueue it and return without touching the host again.  The caller orders later work on
	// skr_stream() or calls skr_sync().  This is what lets back-to-back frames (and the all-gather / de-interleave of
	// the multi-GPU path) queue up behind each other instead of paying a host round trip per frame.
	const bool async = !stats && !want_stats && !tree;
	ctx->timing		 = !async;
	pl.fp.counters = ctx->d_counters;
	pl.fp.err	   = ctx->d_err;
	{
		// whole 8 x 4 pixel blocks leave as 32-bit words when their rows start on word boundaries
		const auto al4 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0u; };
		// (only frames bound for another device or the host: byte stores into the local L2 cost nothing, the staging does)
		bool ok		   = pl.fp.n_peers > 0 && pl.fp.width % 4 == 0 && al4(pl.fp.rgb8);
		for(int k = 0; k < pl.fp.n_peers; k++)
		{
			ok = ok && al4(pl.fp.peers[k]);
		}
		const char *no	  = getenv("SKR_NO_STRIP_WORDS");
		pl.fp.strip_words = (ok && !(no && no[0] == '1')) ? 1 : 0;
	}
	{
		// sample split (primary_kernel): frames whose blocks would fill the GPU's warp slots fewer than five times over -- one
		// rank's share of a 1080p frame at world >= 4 -- and that have samples to split.  Measured (config 2, one rank's share,
		// split / no split): world 2 0.537 / 0.541 ms, world 4 0.279 / 0.285, world 8 0.156 / 0.176; whole frame 1.049 / 0.992
		const char *no	   = getenv("SKR_NO_SPLIT"), *yes = getenv("SKR_SPLIT");
		const long long nb = pl.npix_local / 32, slots = (long long) ctx->sm_count * SKR_MIN_BLOCKS * (SKR_BLOCK / 32);
		// (spp >= 8: the frames launch_primary gives the HALVES variant)
		pl.fp.split = (!tree && !pl.shaded && pl.fp.spp >= 8 && ((nb < 5 * slots && !(no && no[0] == '1')) || (yes && yes[0] == '1'))) ? 1 : 0;
	}
	{
		// whole 32 x 4 strips per CTA (write_strip): 32 x 32 tiles, one block per warp of a 4-warp CTA, RGB8 frames only
		const char *no	= getenv("SKR_NO_STRIP_CTA");
		pl.fp.strip_cta = (pl.fp.strip_words && !pl.fp.split && pl.fp.tile == 32 && SKR_BLOCK == 128 && !pl.fp.rgb32 && !pl.fp.tiles8 && !tree && !pl.shaded &&
						   !(no && no[0] == '1'))
							  ? 1
							  : 0;
	}
	if(!tree && !pl.shaded)
	{
		const int rc_order = tile_launch_order(ctx, pl);
		if(rc_order)
		{
			return rc_order;
		}
	}
	// a single-kernel frame needs no per-kernel events (its span is the frame), no counter reset unless counters were
	// asked for, and no reset of the error word (zero unless a frame failed; cleared again below when read non-zero)
	ctx->timing = !async && tree;
	{
		const int rc_prep = prepare_frame(ctx, o, pl);
		if(rc_prep)
		{
			return rc_prep;
		}
	}
	if(before_launch)
	{
		const int rc_before = before_launch();
		if(rc_before)
		{
			return rc_before;
		}
	}
	if(!async)
	{
		if(want_stats)
		{
			CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * 16, st));
		}
		CK(cudaEventRecord(ctx->ev_begin, st));
	}
	int rc = want_stats ? render_frame<true>(ctx, o, pl) : render_frame<false>(ctx, o, pl);
	if(rc)
	{
		return rc;
	}
	if(after_launch && (rc = after_launch()) != 0)
	{
		return rc;
	}
	if(async)
	{
		ctx->async_pending = true; // its error word is read by skr_sync() or by the next synchronous frame
		return SKR_OK;
	}
	CK(cudaEventRecord(ctx->ev_end, st));
	CK(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
	unsigned long long hc[9] = {0};
	if(want_stats)
	{
		CK(cudaMemcpyAsync(hc, ctx->d_counters, sizeof hc, cudaMemcpyDeviceToHost, st));
	}
	CK(cudaStreamSynchronize(st));
	{
		const int rc_err = check_error_word(ctx, "this frame");
		if(rc_err)
		{
			return rc_err;
		}
	}
	if(stats)
	{
		memset(stats, 0, sizeof *stats);
		stats->closest_hit_rays = hc[0];
		stats->shadow_rays		= hc[1];
		stats->sphere_tests		= hc[2];
		stats->sphere_tests_pos = hc[3];
		stats->tri_tests		= hc[4];
		stats->bvh_node_visits	= hc[5];
		stats->sphere_hits		= hc[6];
		stats->light_evals		= hc[7];
		stats->sphere_tests_executed = hc[8];
		stats->queue_entries	= ctx->queue_entries;
		stats->kernel_launches	= ctx->launches;
		stats->queue_chunks		= ctx->chunks;
		CK(cudaEventElapsedTime(&stats->ms_total, ctx->ev_begin, ctx->ev_end));
		float cat[CAT_COUNT] = {0, 0, 0};
		for(size_t i = 0; i < ctx->spans_used; i++)
		{
			float ms = 0;
			CK(cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b));
			cat[ctx->spans[i].cat] += ms;
		}
		stats->ms_primary = tree ? cat[CAT_PRIMARY] : stats->ms_total;
		stats->ms_bounce  = cat[CAT_BOUNCE];
		stats->ms_resolve = cat[CAT_RESOLVE];
	}
	return SKR_OK;
}

#define REQUIRE_CTX()                                                     \
	if(!ctx)                                                               \
	{                                                                      \
		return fail(nullptr, SKR_ERR_ARG, "null context");                 \
	}                                                                      \
	if(cudaSetDevice(ctx->device) != cudaSuccess)                          \
	{                                                                      \
		return fail(ctx, SKR_ERR_CUDA, "cudaSetDevice(%d) failed", ctx->device); \
	}
#define REQUIRE_SCENE()                                                                      \
	if(!ctx->have_scene)                                                                      \
	{                                                                                         \
		return fail(ctx, SKR_ERR_NO_SCENE, "skr_scene_upload must be called before rendering"); \
	}
} // namespace

extern "C" {

int skr_abi_version(void)
{
	return SKR_ABI_VERSION;
}

#define SKR_STR2(x) #x
#define SKR_STR(x) SKR_STR2(x)
const char *skr_build_info(void)
{
	// what this binary was compiled as (bench.py records it beside its numbers)
	return "libskr abi " SKR_STR(SKR_ABI_VERSION) "; nvcc " SKR_STR(__CUDACC_VER_MAJOR__) "." SKR_STR(__CUDACC_VER_MINOR__) "." SKR_STR(__CUDACC_VER_BUILD__)
#ifdef __CUDA_ARCH_LIST__
		   "; arch list " SKR_STR(__CUDA_ARCH_LIST__)
#endif
		   "; block " SKR_STR(SKR_BLOCK) " x " SKR_STR(SKR_MIN_BLOCKS) " CTAs/SM; gi batch " SKR_STR(SKR_GI_BATCH) "; defer steps " SKR_STR(SKR_DEFER_STEPS)
		   "; built " __DATE__ " " __TIME__;
}

const char *skr_last_error(const skr_ctx *ctx)
{
	return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

int skr_init(int device, skr_ctx **out)
{
	if(!out)
	{
		return fail(nullptr, SKR_ERR_ARG, "skr_init: out is null");
	}
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if(e != cudaSuccess || ndev == 0)
	{
		return fail(nullptr, SKR_ERR_NO_DEVICE, "skr_init: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
	}
	if(device < 0)
	{
		if(cudaGetDevice(&device) != cudaSuccess)
		{
			device = 0;
		}
	}
	if(device >= ndev)
	{
		return fail(nullptr, SKR_ERR_NO_DEVICE, "skr_init: device %d out of range (%d devices)", device, ndev);
	}
	e = cudaSetDevice(device);
	if(e != cudaSuccess)
	{
		return fail(nullptr, SKR_ERR_CUDA, "skr_init: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
	}
	skr_ctx *c = new skr_ctx();
	c->device  = device;
	auto bail  = [&](cudaError_t err, const char *what) {
		 fail(nullptr, SKR_ERR_CUDA, "skr_init: %s: %s", what, cudaGetErrorString(err));
		 skr_destroy(c);
		 return (int) SKR_ERR_CUDA;
	};
	if((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
	{
		return bail(e, "cudaStreamCreate");
	}
	cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
	if((e = cudaMalloc(&c->d_counters, sizeof(unsigned long long) * 16)) != cudaSuccess)
	{
		return bail(e, "cudaMalloc");
	}
	if((e = cudaMalloc(&c->d_cursor, 8 * sizeof(unsigned))) != cudaSuccess || (e = cudaMemset(c->d_cursor, 0, 8 * sizeof(unsigned))) != cudaSuccess)
	{
		return bail(e, "cudaMalloc");
	}
	if((e = cudaMalloc(&c->d_err, sizeof(int))) != cudaSuccess || (e = cudaMemset(c->d_err, 0, sizeof(int))) != cudaSuccess)
	{
		return bail(e, "cudaMalloc");
	}
	if((e = cudaHostAlloc(&c->h_count, sizeof(unsigned), cudaHostAllocDefault)) != cudaSuccess)
	{
		return bail(e, "cudaHostAlloc");
	}
	if((e = cudaHostAlloc(&c->h_err, sizeof(int), cudaHostAllocDefault)) != cudaSuccess)
	{
		return bail(e, "cudaHostAlloc");
	}
	*c->h_err = 0;
	cudaEventCreate(&c->ev_begin);
	cudaEventCreate(&c->ev_end);
	cudaEventCreate(&c->ev_x0);
	cudaEventCreate(&c->ev_x1);
	// overlapped copy-out (optional: without stream memory operations skr_render copies after the kernel)
	{
		void *fw = nullptr, *fr = nullptr;
		cudaDriverEntryPointQueryResult qw, qr;
		if(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fw, cudaEnableDefault, &qw) == cudaSuccess && qw == cudaDriverEntryPointSuccess && fw &&
		   cudaGetDriverEntryPoint("cuStreamWriteValue32", &fr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess && fr &&
		   cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
		   cudaMalloc(&c->d_band, sizeof(unsigned) * 2 * MAX_BANDS) == cudaSuccess && cudaMemset(c->d_band, 0, sizeof(unsigned) * 2 * MAX_BANDS) == cudaSuccess)
		{
			c->wait_value32	 = reinterpret_cast<decltype(c->wait_value32)>(fw);
			c->write_value32 = reinterpret_cast<decltype(c->write_value32)>(fr);
		}
		cudaGetLastError();
	}
	*out = c;
	return SKR_OK;
}

void skr_destroy(skr_ctx *ctx)
{
	if(!ctx)
	{
		return;
	}
	cudaSetDevice(ctx->device);
	if(ctx->stream)
	{
		cudaStreamSynchronize(ctx->stream);
	}
	cudaFree(ctx->d_blob), cudaFree(ctx->d_tris_raw), cudaFree(ctx->d_scratch), cudaFree(ctx->d_tri_mat);
	ctx->bvh_main.release();
	ctx->bvh_shade.release();
	cudaFree(ctx->d_rgb8), cudaFree(ctx->d_rgb32), cudaFree(ctx->d_accum);
	cudaFree(ctx->d_arena);
	cudaFree(ctx->d_cursor), cudaFree(ctx->d_cand_d), cudaFree(ctx->d_cand_px), cudaFree(ctx->d_tile_order);
	cudaFree(ctx->d_counters), cudaFree(ctx->d_err), cudaFree(ctx->d_band);
	if(ctx->copy_stream)
	{
		cudaStreamDestroy(ctx->copy_stream);
	}
	if(ctx->h_count)
	{
		cudaFreeHost(ctx->h_count);
	}
	if(ctx->h_err)
	{
		cudaFreeHost(ctx->h_err);
	}
	for(Span &s : ctx->spans)
	{
		cudaEventDestroy(s.a), cudaEventDestroy(s.b);
	}
	if(ctx->ev_begin)
	{
		cudaEventDestroy(ctx->ev_begin), cudaEventDestroy(ctx->ev_end), cudaEventDestroy(ctx->ev_x0), cudaEventDestroy(ctx->ev_x1);
	}
	if(ctx->stream)
	{
		cudaStreamDestroy(ctx->stream);
	}
	delete ctx;
}

void *skr_stream(skr_ctx *ctx)
{
	return ctx ? (void *) ctx->stream : nullptr;
}

int skr_sync(skr_ctx *ctx)
{
	REQUIRE_CTX();
	if(!ctx->async_pending)
	{
		CK(cudaStreamSynchronize(ctx->stream));
		return SKR_OK;
	}
	// asynchronous frames report through the device error word (queue / traversal-stack overflow): read it here
	CK(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return check_error_word(ctx, "an asynchronous frame");
}

int skr_scene_upload(skr_ctx *ctx, const skr_scene_desc *sc)
{
	REQUIRE_CTX();
	if(!sc)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_scene_upload: scene is null");
	}
	if(sc->nspheres < 0 || sc->ntris < 0 || sc->nplights < 0 || sc->ndlights < 0 || sc->nfogs < 0)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_scene_upload: negative count");
	}
	if(sc->nspheres > 65535)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_scene_upload: at most 65535 spheres (got %d)", sc->nspheres);
	}
	if((sc->nspheres && !sc->spheres) || (sc->ntris && !sc->tris) || (sc->nplights && !sc->plights) || (sc->ndlights && !sc->dlights) ||
	   (sc->nfogs && !sc->fogs))
	{
		return fail(ctx, SKR_ERR_ARG, "skr_scene_upload: null array with nonzero count");
	}
	ctx->have_scene = false;
	CK(cudaMemsetAsync(ctx->d_cursor, 0, 8 * sizeof(unsigned), ctx->stream)); // (a frame that died mid-kernel may have left the counters armed)
	const int S = sc->nspheres, T = sc->ntris, L = sc->nplights, D = sc->ndlights, F = sc->nfogs;
	const int S4 = (S + 3) / 4 * 4;
	SceneView &sv = ctx->sv;
	memset(&sv, 0, sizeof sv);
	sv.S = S, sv.S4 = S4, sv.T = T, sv.L = L, sv.D = D, sv.F = F;
	int off		  = 0;
	sv.off_geom	  = off, off += S;
	sv.off_pgeom  = off, off += S4; // S4/2 pairs x 2 float4
	sv.off_pprim  = off, off += S4;
	sv.off_amb	  = off, off += S;
	sv.off_diff	  = off, off += S;
	sv.off_spec	  = off, off += S;
	sv.off_plpos  = off, off += L;
	sv.off_plcol  = off, off += L;
	sv.off_dldir  = off, off += D;
	sv.off_dlcol  = off, off += D;
	sv.off_foga	  = off, off += F;
	sv.off_fogalb = off, off += F;
	sv.off_fogp	  = off, off += (S * L * F + 3) / 4;
	const int NP  = S4 / 2;
	sv.off_cull	  = -1;
	if(S > 0 && NP <= 32)
	{
		sv.off_cull = off, off += 3 * NP * (1 + L);
	}
	sv.cull_shadow = (sv.off_cull >= 0 && L * NP <= 64) ? 1 : 0;
	sv.off_rmask   = -1;
	{
		const char *nocull = getenv("SKR_NO_CULL");
		if(sv.off_cull >= 0 && L > 0 && !(nocull && nocull[0] == '1'))
		{
			sv.off_rmask = off, off += (S * L + S + 3) / 4;
		}
	}
	sv.blob_f4	  = off > 0 ? off : 1;
	std::vector<float4> blob((size_t) sv.blob_f4, make_float4(0, 0, 0, 0));
	const V3 cam = ld3(sc->camera);
	for(int s = 0; s < S; s++)
	{
		const float *p = sc->spheres + 18 * (size_t) s;
		const V3 c	   = ld3(p);
		const float r  = p[3];
		// e = camera - centre, c-term = e.e - r*r in the reference's operand order (src/utils.h:115-118)
		const V3 e			   = V3{cam.x - c.x, cam.y - c.y, cam.z - c.z};
		const float ee		   = hdot(e, e);
		const float cterm	   = ee - r * r;
		blob[sv.off_geom + s]  = make_float4(c.x, c.y, c.z, -(r * r));
		{
			// pair layout: element (s & 1) of pair s >> 1
			float *g = reinterpret_cast<float *>(&blob[sv.off_pgeom + 2 * (s >> 1)]);
			float *q = reinterpret_cast<float *>(&blob[sv.off_pprim + 2 * (s >> 1)]);
			const int k = s & 1;
			g[0 + k] = -c.x, g[2 + k] = -c.y, g[4 + k] = -c.z, g[6 + k] = -(r * r);
			q[0 + k] = e.x, q[2 + k] = e.y, q[4 + k] = e.z, q[6 + k] = cterm;
		}
		blob[sv.off_amb + s]   = make_float4(sc->ambient[0] * p[4], sc->ambient[1] * p[5], sc->ambient[2] * p[6], p[16]);
		blob[sv.off_diff + s]  = make_float4(p[7], p[8], p[9], p[17]);
		blob[sv.off_spec + s]  = make_float4(p[10], p[11], p[12], r);
	}
	for(int s = S; s < S4; s++)
	{
		// padding spheres: e.e overflows to +inf, so h*h - a*cc is -inf (or NaN) and every `>= 0` test fails
		float *g = reinterpret_cast<float *>(&blob[sv.off_pgeom + 2 * (s >> 1)]);
		float *q = reinterpret_cast<float *>(&blob[sv.off_pprim + 2 * (s >> 1)]);
		const int k = s & 1;
		g[0 + k] = -3.0e19f, g[2 + k] = -3.0e19f, g[4 + k] = -3.0e19f, g[6 + k] = 0.0f;
		q[0 + k] = 0.0f, q[2 + k] = 0.0f, q[4 + k] = 0.0f, q[6 + k] = 3.0e38f;
	}
	for(int i = 0; i < L; i++)
	{
		const float *p		   = sc->plights + 6 * (size_t) i;
		blob[sv.off_plpos + i] = make_float4(p[0], p[1], p[2], 0);
		blob[sv.off_plcol + i] = make_float4(p[3], p[4], p[5], 0);
	}
	for(int i = 0; i < D; i++)
	{
		const float *p	= sc->dlights + 6 * (size_t) i;
		const V3 d		= ld3(p);
		const float inv = 1.0f / sqrtf(d.x * d.x + d.y * d.y + d.z * d.z); // glm::normalize
		blob[sv.off_dldir + i] = make_float4(d.x * inv, d.y * inv, d.z * inv, 0);
		blob[sv.off_dlcol + i] = make_float4(p[3], p[4], p[5], 0);
	}
	if(sv.off_cull >= 0)
	{
		// bundle-culling tables (skr_device.cuh: cull_pairs): apex 0 = camera, 1 + i = point light i
		for(int a = 0; a <= L; a++)
		{
			const V3 A = a == 0 ? cam : ld3(sc->plights + 6 * (size_t) (a - 1));
			for(int s = 0; s < S4; s++)
			{
				float *q	= reinterpret_cast<float *>(&blob[sv.off_cull + 3 * ((size_t) a * NP + (s >> 1))]);
				const int k = s & 1;
				if(s >= S) // padding: always culled (dist^2 = 3e38 > any margin^2)
				{
					q[0 + k] = 0, q[2 + k] = 0, q[4 + k] = 0, q[6 + k] = 3.0e38f, q[8 + k] = 0, q[10 + k] = 0;
					continue;
				}
				const float *p	= sc->spheres + 18 * (size_t) s;
				const double ux = (double) p[0] - A.x, uy = (double) p[1] - A.y, uz = (double) p[2] - A.z, r = p[3];
				const double uu = ux * ux + uy * uy + uz * uz;
				const double R	= sqrt(r * r + 1e-4 * uu) + 1e-4 * (1.0 + fabs(A.x) + fabs(A.y) + fabs(A.z) + fabs(p[0]) + fabs(p[1]) + fabs(p[2]));
				q[0 + k] = (float) ux, q[2 + k] = (float) uy, q[4 + k] = (float) uz, q[6 + k] = (float) uu;
				q[8 + k] = (float) (R * (1.0 + 1e-6)), q[10 + k] = (float) (sqrt(uu) * (1.0 + 1e-6));
			}
		}
	}
	if(sv.off_rmask >= 0)
	{
		// static shadow masks: cull_pairs (skr_device.cuh) in double for the bundle "lines through light i and a point
		// within rho of c_s", rho = r_s with slack for the rounding of the hit point; same margins as the device test
		uint32_t *rm = reinterpret_cast<uint32_t *>(blob.data() + sv.off_rmask);
		float *rchk	 = reinterpret_cast<float *>(rm + (size_t) S * L);
		const uint32_t full = NP >= 32 ? 0xffffffffu : (1u << NP) - 1u;
		for(int s = 0; s < S; s++)
		{
			const float *ps	 = sc->spheres + 18 * (size_t) s;
			const double rho = fabs((double) ps[3]) * 1.001 + 1e-4 * (1.0 + fabs(ps[0]) + fabs(ps[1]) + fabs(ps[2]));
			rchk[s]			 = (float) (rho * rho * (1.0 - 1e-5));
			for(int i = 0; i < L; i++)
			{
				const float *pl = sc->plights + 6 * (size_t) i;
				const double wx = (double) ps[0] - pl[0], wy = (double) ps[1] - pl[1], wz = (double) ps[2] - pl[2];
				const double ww = wx * wx + wy * wy + wz * wz, len = sqrt(ww);
				uint32_t mk		= full;
				if(rho <= 0.45 * len)
				{
					mk					= 0;
					const double beta	= 1.07 * rho / len, m = 0.01 * len;
					const float *ctab	= reinterpret_cast<const float *>(&blob[sv.off_cull + 3 * (size_t) (1 + i) * NP]);
					for(int k = 0; k < S; k++)
					{
						const float *q	= ctab + 12 * (size_t) (k >> 1);
						const int e		= k & 1;
						const double hw = wx * q[0 + e] + wy * q[2 + e] + wz * q[4 + e];
						const double X	= (double) q[8 + e] + (double) q[10 + e] * beta + m;
						if(!((double) q[6 + e] - hw * hw / ww > X * X))
						{
							mk |= 1u << (k >> 1);
						}
					}
				}
				rm[(size_t) s * L + i] = mk;
			}
		}
	}
	float *fogp = reinterpret_cast<float *>(blob.data() + sv.off_fogp);
	for(int j = 0; j < F; j++)
	{
		const float *f			= sc->fogs + 9 * (size_t) j;
		blob[sv.off_foga + j]	= make_float4(f[0], f[1], f[5], 0);
		blob[sv.off_fogalb + j] = make_float4(f[2], f[3], f[4], 0);
	}
	for(int s = 0; s < S; s++)
	{
		const V3 c = ld3(sc->spheres + 18 * (size_t) s);
		for(int i = 0; i < L; i++)
		{
			const V3 lp = ld3(sc->plights + 6 * (size_t) i);
			const V3 dv = V3{c.x - lp.x, c.y - lp.y, c.z - lp.z};
			for(int j = 0; j < F; j++)
			{
				const float *f = sc->fogs + 9 * (size_t) j;
				// src/blinn_phong.h:22-29: distance clamp and probability of no interaction (exp in double)
				float distance = sqrtf(hdot(dv, dv));
				if(distance > 2 * f[5])
				{
					distance = 2 * f[5];
				}
				fogp[((size_t) s * L + i) * F + j] = (float) exp((double) (-1.0f * distance * (f[1] + f[0])));
			}
		}
	}
	sv.cam_pos	  = make_float3(sc->camera[0], sc->camera[1], sc->camera[2]);
	sv.cam_dir	  = make_float3(sc->camera[3], sc->camera[4], sc->camera[5]);
	sv.cam_up	  = make_float3(sc->camera[6], sc->camera[7], sc->camera[8]);
	sv.cam_right  = make_float3(sc->camera[9], sc->camera[10], sc->camera[11]);
	sv.background = make_float3(sc->background[0], sc->background[1], sc->background[2]);
	sv.err		  = ctx->d_err;

	// an earlier asynchronous frame may still be reading the old blob on this stream: the copies below are
	// stream-ordered behind it, and buffers are only ever re-allocated through ensure() (cudaFree synchronises)
	CK(ensure(ctx->d_blob, ctx->blob_bytes, sizeof(float4) * blob.size()));
	CK(cudaMemcpyAsync(ctx->d_blob, blob.data(), sizeof(float4) * blob.size(), cudaMemcpyHostToDevice, ctx->stream));
	sv.blob			= ctx->d_blob;
	const size_t bb = sizeof(float4) * blob.size();
	sv.blob_in_smem = bb <= SMEM_BLOB_LIMIT ? 1 : 0;
	ctx->smem_bytes = sv.blob_in_smem ? bb : 0;
	int rc			= set_smem_attr(ctx);
	if(rc)
	{
		return rc;
	}

	sv.tri_v = nullptr;
	sv.bvh	 = nullptr;
	ctx->n_tris			 = T;
	ctx->bvh_shade.valid = false; // rebuilt by the first shaded-triangles frame of this scene
	ctx->have_tri_mat	 = false;
	std::vector<float4> tmat;
	if(T > 0)
	{
		CK(ensure(ctx->d_tris_raw, ctx->tris_raw_bytes, sizeof(float) * 9 * (size_t) T));
		CK(cudaMemcpyAsync(ctx->d_tris_raw, sc->tris, sizeof(float) * 9 * (size_t) T, cudaMemcpyHostToDevice, ctx->stream));
		rc = build_bvh(ctx, T, ctx->bvh_main, true);
		if(rc)
		{
			return rc;
		}
		sv.tri_v = ctx->bvh_main.d_tri_v;
		sv.bvh	 = ctx->bvh_main.bvh;
		sv.big_v = ctx->bvh_main.d_big;
		if(sc->tri_materials)
		{
			// (ambient_light (.) ka, power), (kd, ior), (ks, 0) per triangle, original order (shaded-triangles mode only)
			tmat.resize(3 * (size_t) T);
			for(int t = 0; t < T; t++)
			{
				const float *m		= sc->tri_materials + 14 * (size_t) t;
				tmat[3 * t + 0]		= make_float4(sc->ambient[0] * m[0], sc->ambient[1] * m[1], sc->ambient[2] * m[2], m[12]);
				tmat[3 * t + 1]		= make_float4(m[3], m[4], m[5], m[13]);
				tmat[3 * t + 2]		= make_float4(m[6], m[7], m[8], 0.0f);
			}
			CK(ensure(ctx->d_tri_mat, ctx->tri_mat_bytes, sizeof(float4) * tmat.size()));
			CK(cudaMemcpyAsync(ctx->d_tri_mat, tmat.data(), sizeof(float4) * tmat.size(), cudaMemcpyHostToDevice, ctx->stream));
			ctx->have_tri_mat = true;
		}
	}
	sv.tri_mat	= ctx->have_tri_mat ? ctx->d_tri_mat : nullptr;
	sv.tris_raw = T > 0 ? ctx->d_tris_raw : nullptr;
	if(T > 0)
	{
		CK(cudaStreamSynchronize(ctx->stream)); // the count of outsized triangles comes back from the build
	}
	// (sphere scenes: nothing to wait for -- the blob copy above is stream-ordered before every frame, and cudaMemcpyAsync
	// from the pageable staging vector has consumed it before it returns)
	sv.nbig			= (T > 0 && sv.bvh) ? std::min((int) *ctx->h_count, BIG_TRI_CAP) : 0;
	ctx->have_scene = true;
	ctx->host_spheres.assign(sc->spheres, sc->spheres + 18 * (size_t) S); // (geometry for the tile classification of tile_launch_order)
	{
		// FNV-1a over what tile_launch_order reads: the same scene uploaded again keeps its cached tile order
		unsigned long long h = 1469598103934665603ull;
		const auto mix		 = [&h](const void *p, size_t n) {
			  const unsigned char *b = static_cast<const unsigned char *>(p);
			  for(size_t i = 0; i < n; i++)
			  {
				  h = (h ^ b[i]) * 1099511628211ull;
			  }
		};
		mix(ctx->host_spheres.data(), sizeof(float) * ctx->host_spheres.size());
		mix(sc->camera, sizeof sc->camera);
		mix(&T, sizeof T);
		ctx->scene_gen = h;
	}
	return SKR_OK;
}

int skr_reserve(skr_ctx *ctx, const skr_options *opt)
{
	REQUIRE_CTX();
	REQUIRE_SCENE();
	Plan pl;
	int rc = make_plan(ctx, opt, pl);
	if(rc)
	{
		return rc;
	}
	if((rc = prepare_frame(ctx, opt, pl)) != 0)
	{
		return rc;
	}
	const size_t npx = (size_t) opt->width * opt->height;
	CK(ensure(ctx->d_rgb8, ctx->rgb8_bytes, npx * 3));
	std::vector<const void *> fns;
	const bool tree = (pl.fp.gi || pl.fp.fresnel) && pl.levels > 0;
	if(opt->collect_stats)
	{
		frame_kernels<true>(ctx, tree, pl.leaf_inline, pl.fp.fresnel != 0, fns);
	}
	else
	{
		frame_kernels<false>(ctx, tree, pl.leaf_inline, pl.fp.fresnel != 0, fns);
	}
	for(const void *f : fns)
	{
		cudaFuncAttributes a;
		CK(cudaFuncGetAttributes(&a, f));
	}
	return SKR_OK;
}

int skr_render_device(skr_ctx *ctx, const skr_options *opt, void *d_rgb8, void *d_rgb32, skr_stats *stats)
{
	REQUIRE_CTX();
	REQUIRE_SCENE();
	Plan pl;
	int rc = make_plan(ctx, opt, pl);
	if(rc)
	{
		return rc;
	}
	pl.fp.rgb8	= static_cast<uint8_t *>(d_rgb8);
	pl.fp.rgb32 = static_cast<float *>(d_rgb32);
	return render_common(ctx, opt, pl, stats);
}

int skr_render(skr_ctx *ctx, const skr_options *opt, uint8_t *rgb8, float *rgb32, skr_stats *stats)
{
	REQUIRE_CTX();
	REQUIRE_SCENE();
	Plan pl;
	int rc = make_plan(ctx, opt, pl);
	if(rc)
	{
		return rc;
	}
	const size_t npx = (size_t) opt->width * opt->height;
	if(rgb8)
	{
		CK(ensure(ctx->d_rgb8, ctx->rgb8_bytes, npx * 3));
		pl.fp.rgb8 = ctx->d_rgb8;
		if(pl.fp.world > 1)
		{
			CK(cudaMemsetAsync(ctx->d_rgb8, 0, npx * 3, ctx->stream));
		}
	}
	if(rgb32)
	{
		CK(ensure(ctx->d_rgb32, ctx->rgb32_bytes, npx * 3 * sizeof(float)));
		pl.fp.rgb32 = ctx->d_rgb32;
		if(pl.fp.world > 1)
		{
			CK(cudaMemsetAsync(ctx->d_rgb32, 0, npx * 3 * sizeof(float), ctx->stream));
		}
	}
	// Frames that are ONE long kernel and leave as RGB8 only, into page-locked memory this device can address: the kernel
	// stores the finished 8 x 4 blocks STRAIGHT into the host frame as 32-bit words (write_block, the path of
	// skr_render_peers_device), which cross PCIe while the rest of the frame is still being traced -- no copy engine, no
	// flags, and the tiles keep the frame-wide heavy-first launch order.  Config 2: 1.07 ms per upload + frame against 1.13
	// for the band copies below and 1.17 for a copy after the kernel; the kernel itself is not slowed (0.995 ms), whereas
	// copy-engine traffic running beside it costs it ~0.045 ms.  Short kernels (single-sample frames) would wait for PCIe.
	{
		const char *no = getenv("SKR_NO_OVERLAP"), *nod = getenv("SKR_NO_HOST_STORES");
		cudaPointerAttributes a;
		if(rgb8 && !rgb32 && pl.levels == 0 && !pl.shaded && pl.fp.world == 1 && pl.fp.spp >= 4 && opt->width % 4 == 0 && npx * 3 >= (1u << 20) &&
		   (reinterpret_cast<uintptr_t>(rgb8) & 3u) == 0u && !(no && no[0] == '1') && !(nod && nod[0] == '1'))
		{
			if(cudaPointerGetAttributes(&a, rgb8) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer != nullptr)
			{
				pl.fp.rgb8		= nullptr;
				pl.fp.peers[0]	= static_cast<uint8_t *>(a.devicePointer);
				pl.fp.n_peers	= 1;
				pl.fp.peer_rows = 0;
				skr_stats local;
				rc = render_common(ctx, opt, pl, &local); // (returns once the stream has drained: the stores have landed)
				if(rc)
				{
					return rc;
				}
				local.ms_d2h = 0.0f; // no copy: the frame's bytes crossed PCIe inside ms_total
				if(stats)
				{
					*stats = local;
				}
				return SKR_OK;
			}
			cudaGetLastError();
		}
	}
	// Otherwise (float frame wanted too, or memory the device cannot address) the copy-out is overlapped with the kernel by
	// the copy engine: the frame leaves in bands of whole tile rows; the
	// copy of band b sits on a second stream behind a stream-ordered wait (cuStreamWaitValue32) on the flag that the last
	// CTA of the band sets (primary_kernel), so all but the last band cross PCIe while the kernel is still tracing.
	struct Band
	{
		size_t y0, y1;
	} bands[MAX_BANDS];
	int nb = 0;
	{
		const FrameParams &fp = pl.fp;
		const char *no		  = getenv("SKR_NO_OVERLAP");
		const int tpix		  = fp.tile * fp.tile;
		// page-locked destinations only: a copy to pageable memory blocks the host until it is done, and it would be
		// enqueued before the kernel it waits for
		const auto pinned = [](const void *p) {
			cudaPointerAttributes a;
			if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
			{
				cudaGetLastError();
				return false;
			}
			return a.type == cudaMemoryTypeHost;
		};
		// (worth its ~0.04 ms of extra stream operations only when the kernel is long against the copy: jittered frames)
		if(ctx->wait_value32 && pl.levels == 0 && !pl.shaded && fp.world == 1 && fp.spp >= 4 && tpix % SKR_BLOCK == 0 && npx * 3 >= (1u << 20) && !(no && no[0] == '1') &&
		   (!rgb8 || pinned(rgb8)) && (!rgb32 || pinned(rgb32)) && (rgb8 || rgb32))
		{
			const int tile_rows = (opt->height + fp.tile - 1) / fp.tile;
			const char *nbe		= getenv("SKR_BANDS"); // (tuning aid) bands per frame, 1 .. 8; default 6
			const int nwant		= nbe && atoi(nbe) >= 1 && atoi(nbe) <= MAX_BANDS ? atoi(nbe) : 6;
			const int rpb		= (tile_rows + nwant - 1) / nwant;
			nb					= (tile_rows + rpb - 1) / rpb;
			for(int b = 0; b < nb; b++)
			{
				bands[b].y0 = (size_t) b * rpb * fp.tile;
				bands[b].y1 = std::min((size_t) opt->height, (size_t) (b + 1) * rpb * fp.tile);
			}
			pl.fp.band_count = ctx->d_band;
			pl.fp.band_flag	 = ctx->d_band + MAX_BANDS;
			pl.fp.n_bands	 = (unsigned) nb;
			for(int b = 0; b <= nb; b++) // scan order; tile_launch_order may permute the bands
			{
				const size_t rows	= std::min((size_t) tile_rows, (size_t) b * rpb);
				pl.fp.band_start[b] = (unsigned) (rows * (size_t) fp.tiles_x * (size_t) (tpix / 32));
				if(b < nb)
				{
					pl.band_perm[b] = b;
				}
			}
			pl.fp.band_seq	 = ++ctx->band_seq;
			pl.rows_per_band = rpb;
		}
	}
	const size_t row8 = (size_t) opt->width * 3, row32 = row8 * sizeof(float);
	if(nb > 0)
	{
		// the per-band CTA counters run on from frame to frame; reset when the band geometry changes (and now and then,
		// far from 32-bit wrap-around) or after a frame that did not complete
		// (the launch order of the bands is part of that geometry and is only known once the frame is planned: the reset sits
		// in `before_launch`, which render_common calls after tile_launch_order and before the kernel)
		const auto reset_counters = [&]() -> int {
			unsigned long long geom = 1469598103934665603ull;
			for(unsigned k = 0; k <= pl.fp.n_bands; k++)
			{
				geom = (geom ^ pl.fp.band_start[k]) * 1099511628211ull;
			}
			if(geom != ctx->band_geom || (pl.fp.band_seq & 0xffffu) == 0u)
			{
				CK(cudaMemsetAsync(ctx->d_band, 0, sizeof(unsigned) * MAX_BANDS, ctx->stream));
				ctx->band_geom = geom;
			}
			return SKR_OK;
		};
		// The waits are enqueued AFTER the kernel they wait for: should the two streams share a hardware queue, the copies
		// then merely line up behind the kernel; a wait enqueued first could block the kernel behind it for ever.
		bool enqueued		   = false;
		const auto copy_bands = [&]() -> int {
			enqueued = true;
			for(int b = 0; b < nb; b++)
			{
				if(ctx->wait_value32((CUstream) ctx->copy_stream, (CUdeviceptr) (pl.fp.band_flag + b), pl.fp.band_seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
				{
					return fail(ctx, SKR_ERR_CUDA, "skr_render: cuStreamWaitValue32 failed");
				}
				if(b == 0)
				{
					CK(cudaEventRecord(ctx->ev_x0, ctx->copy_stream));
				}
				const int rb	= pl.band_perm[b]; // the band of rows launched b-th
				const size_t y0 = bands[rb].y0, rows = bands[rb].y1 - bands[rb].y0;
				if(rgb8)
				{
					CK(cudaMemcpyAsync(rgb8 + y0 * row8, ctx->d_rgb8 + y0 * row8, rows * row8, cudaMemcpyDeviceToHost, ctx->copy_stream));
				}
				if(rgb32)
				{
					CK(cudaMemcpyAsync((char *) rgb32 + y0 * row32, (char *) ctx->d_rgb32 + y0 * row32, rows * row32, cudaMemcpyDeviceToHost, ctx->copy_stream));
				}
			}
			CK(cudaEventRecord(ctx->ev_x1, ctx->copy_stream));
			return SKR_OK;
		};
		skr_stats local;
		rc = render_common(ctx, opt, pl, &local, copy_bands, reset_counters);
		if(enqueued)
		{
			if(rc)
			{
				// the frame failed somewhere: nothing may stay parked on the copy stream -- publish every flag
				for(int b = 0; b < nb; b++)
				{
					ctx->write_value32((CUstream) ctx->stream, (CUdeviceptr) (pl.fp.band_flag + b), pl.fp.band_seq, 0);
				}
				cudaStreamSynchronize(ctx->stream);
			}
			const cudaError_t ce = cudaStreamSynchronize(ctx->copy_stream);
			if(!rc)
			{
				CK(ce);
			}
		}
		if(rc)
		{
			ctx->band_geom = 0;
			return rc;
		}
		CK(cudaEventElapsedTime(&local.ms_d2h, ctx->ev_x0, ctx->ev_x1));
		if(stats)
		{
			*stats = local;
		}
		return SKR_OK;
	}
	skr_stats local;
	rc = render_common(ctx, opt, pl, &local);
	if(rc)
	{
		return rc;
	}
	CK(cudaEventRecord(ctx->ev_x0, ctx->stream));
	if(rgb8)
	{
		CK(cudaMemcpyAsync(rgb8, ctx->d_rgb8, npx * 3, cudaMemcpyDeviceToHost, ctx->stream));
	}
	if(rgb32)
	{
		CK(cudaMemcpyAsync(rgb32, ctx->d_rgb32, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(cudaEventRecord(ctx->ev_x1, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaEventElapsedTime(&local.ms_d2h, ctx->ev_x0, ctx->ev_x1));
	if(stats)
	{
		*stats = local;
	}
	return SKR_OK;
}

int64_t skr_tiles_bytes(const skr_options *opt)
{
	if(!opt || opt->width <= 0 || opt->height <= 0)
	{
		return -1;
	}
	const int tile	  = opt->tile > 0 ? opt->tile : DEFAULT_TILE;
	const int world	  = opt->world > 1 ? opt->world : 1;
	const int64_t tx  = (opt->width + tile - 1) / tile, ty = (opt->height + tile - 1) / tile;
	const int64_t per = (tx * ty + world - 1) / world;
	return per * tile * tile * 3;
}

int skr_render_tiles_device(skr_ctx *ctx, const skr_options *opt, void *d_tiles, skr_stats *stats)
{
	REQUIRE_CTX();
	REQUIRE_SCENE();
	if(!d_tiles)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_render_tiles_device: d_tiles is null");
	}
	Plan pl;
	int rc = make_plan(ctx, opt, pl);
	if(rc)
	{
		return rc;
	}
	pl.fp.tiles8 = static_cast<uint8_t *>(d_tiles);
	return render_common(ctx, opt, pl, stats);
}

static int render_to_frames(skr_ctx *ctx, const skr_options *opt, void *const *d_frames, int n_frames, int rows_per_frame, skr_stats *stats, const char *who)
{
	REQUIRE_CTX();
	REQUIRE_SCENE();
	if(!d_frames || n_frames < 1 || n_frames > 8)
	{
		return fail(ctx, SKR_ERR_ARG, "%s: between 1 and 8 frame pointers are required (got %d)", who, n_frames);
	}
	Plan pl;
	int rc = make_plan(ctx, opt, pl);
	if(rc)
	{
		return rc;
	}
	for(int k = 0; k < n_frames; k++)
	{
		if(!d_frames[k])
		{
			return fail(ctx, SKR_ERR_ARG, "%s: frame pointer %d is null", who, k);
		}
		pl.fp.peers[k] = static_cast<uint8_t *>(d_frames[k]);
	}
	pl.fp.n_peers	= n_frames;
	pl.fp.peer_rows = rows_per_frame;
	pl.fp.fd_peer_rows = make_fastdiv((uint32_t) (rows_per_frame > 0 ? rows_per_frame : 1));
	return render_common(ctx, opt, pl, stats);
}

int skr_render_peers_device(skr_ctx *ctx, const skr_options *opt, void *const *d_frames, int n_frames, skr_stats *stats)
{
	return render_to_frames(ctx, opt, d_frames, n_frames, 0, stats, "skr_render_peers_device");
}

int skr_render_bands_device(skr_ctx *ctx, const skr_options *opt, void *const *d_frames, int n_frames, int rows_per_frame, skr_stats *stats)
{
	if(rows_per_frame <= 0 || rows_per_frame % 4 != 0)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_render_bands_device: rows_per_frame must be a positive multiple of 4 (got %d)", rows_per_frame);
	}
	return render_to_frames(ctx, opt, d_frames, n_frames, rows_per_frame, stats, "skr_render_bands_device");
}

int skr_deinterleave_device(skr_ctx *ctx, const skr_options *opt, const void *d_gathered, void *d_rgb8)
{
	REQUIRE_CTX();
	Plan pl;
	int rc = make_plan(ctx, opt, pl);
	if(rc)
	{
		return rc;
	}
	if(!d_gathered || !d_rgb8)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_deinterleave_device: null buffer");
	}
	const long long n = (long long) opt->width * opt->height;
	deinterleave_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_gathered), static_cast<uint8_t *>(d_rgb8),
																			   opt->width, opt->height, pl.fp.tile, pl.fp.tiles_x, pl.fp.world, pl.tiles_local);
	CK(cudaGetLastError());
	return SKR_OK;
}

double skr_measure_fp32_peak(skr_ctx *ctx, int iters)
{
	if(!ctx || cudaSetDevice(ctx->device) != cudaSuccess)
	{
		return -1.0;
	}
	if(iters <= 0)
	{
		iters = 2048;
	}
	const int blocks = ctx->sm_count * 8, threads = 256;
	float *d_out = nullptr;
	if(cudaMalloc(&d_out, sizeof(float) * (size_t) blocks * threads) != cudaSuccess)
	{
		return -1.0;
	}
	double best = 0.0;
	for(int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(ctx->ev_x0, ctx->stream);
		fma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 1.0000001f, 1.0e-7f);
		cudaEventRecord(ctx->ev_x1, ctx->stream);
		if(cudaStreamSynchronize(ctx->stream) != cudaSuccess)
		{
			cudaFree(d_out);
			return -1.0;
		}
		float ms = 0;
		cudaEventElapsedTime(&ms, ctx->ev_x0, ctx->ev_x1);
		const double flops = 2.0 * 8.0 * 16.0 * (double) iters * (double) blocks * threads;
		const double tf	   = flops / (ms * 1e-3) / 1e12;
		if(rep > 0 && tf > best)
		{
			best = tf;
		}
	}
	cudaFree(d_out);
	return best;
}

int skr_pin_host(skr_ctx *ctx, void *host, size_t bytes, void **d_ptr)
{
	REQUIRE_CTX();
	if(!host || !bytes || !d_ptr)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_pin_host: null argument");
	}
	*d_ptr = nullptr;
	cudaPointerAttributes a;
	const bool known = cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if(!known)
	{
		CK(cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
	}
	CK(cudaHostGetDevicePointer(d_ptr, host, 0));
	return known ? 1000 : SKR_OK; // 1000: was page-locked already (nothing to undo)
}

int skr_copy_to_host(skr_ctx *ctx, void *host_dst, const void *d_src, size_t bytes)
{
	REQUIRE_CTX();
	if(!host_dst || !d_src)
	{
		return fail(ctx, SKR_ERR_ARG, "skr_copy_to_host: null pointer");
	}
	if(bytes)
	{
		CK(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	}
	return SKR_OK;
}

int skr_unpin_host(skr_ctx *ctx, void *host)
{
	REQUIRE_CTX();
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaHostUnregister(host));
	return SKR_OK;
}

double skr_measure_bandwidth(skr_ctx *ctx, int level)
{
	if(!ctx || cudaSetDevice(ctx->device) != cudaSuccess || level < 0 || level > 2)
	{
		return -1.0;
	}
	const int blocks = ctx->sm_count * 8, threads = 256;
	const int iters	 = level == 2 ? 256 : 2048;
	float *d_out	 = nullptr;
	float4 *d_buf	 = nullptr;
	// L1: every CTA walks the same 16 KB window, resident in each SM's L1; L2: one 64 MB buffer streamed by all CTAs
	const size_t buf_f4 = level == 1 ? 1024 : level == 2 ? (size_t) (64u << 20) / 16 : 1;
	if(cudaMalloc(&d_out, sizeof(float) * (size_t) blocks * threads) != cudaSuccess)
	{
		return -1.0;
	}
	if(cudaMalloc(&d_buf, sizeof(float4) * buf_f4) != cudaSuccess || cudaMemsetAsync(d_buf, 0, sizeof(float4) * buf_f4, ctx->stream) != cudaSuccess)
	{
		cudaFree(d_out);
		return -1.0;
	}
	double best = 0.0;
	for(int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(ctx->ev_x0, ctx->stream);
		if(level == 0)
		{
			lds_bw_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters);
		}
		else if(level == 1)
		{
			gmem_bw_kernel<0><<<blocks, threads, 0, ctx->stream>>>(d_buf, 0xffffffffu, 0u, d_out, 0); // (loads the kernel)
			gmem_bw_kernel<0><<<blocks, threads, 0, ctx->stream>>>(d_buf, 1023u, 0u, d_out, 8);		  // warm the window
			cudaEventRecord(ctx->ev_x0, ctx->stream);
			gmem_bw_kernel<0><<<blocks, threads, 0, ctx->stream>>>(d_buf, 1023u, 0u, d_out, iters);
		}
		else
		{
			gmem_bw_kernel<1><<<blocks, threads, 0, ctx->stream>>>(d_buf, (unsigned) (buf_f4 - 1), 4099u * 16u, d_out, iters);
		}
		cudaEventRecord(ctx->ev_x1, ctx->stream);
		if(cudaStreamSynchronize(ctx->stream) != cudaSuccess)
		{
			best = -1.0;
			break;
		}
		float ms = 0;
		cudaEventElapsedTime(&ms, ctx->ev_x0, ctx->ev_x1);
		const double bytes = 16.0 * 8.0 * (double) iters * (double) blocks * threads;
		const double gbs   = bytes / (ms * 1e-3) / 1e9;
		if(rep > 0 && gbs > best)
		{
			best = gbs;
		}
	}
	cudaFree(d_out);
	cudaFree(d_buf);
	return best;
}

} // extern "C"
