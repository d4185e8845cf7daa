// skr_kernels.cuh -- the frame kernels.
//
//   primary_kernel       reference src/main.cpp:33-86 (per-pixel loop: camera / jittered ray generation) fused with the
//                        first half of shade() (closest sphere + triangle any-hit, src/raytrace.h:149-192).  Rays never
//                        touch memory.  Without --gillum it also shades the hit inline (primary hits are image-coherent,
//                        so warps stay converged) and writes the finished pixel; with --gillum it pushes each sphere hit
//                        to wavefront queue level 0.
//   shade_expand_kernel  one thread per queued hit: direct illumination (+ shadow rays) of that hit, then -- if the
//                        reference would recurse (depth - 1 >= 1) -- the fan-out of montecarlo_global_illumination
//                        (src/raytrace.h:107-136): n child rays generated from Philox draws, intersected four at a
//                        time, misses folded into the local sum, sphere hits pushed (warp-aggregated, one reservation
//                        per batch) to the next queue level as the ray that found them; the exact hit point is
//                        computed by the consumer.  Every lane of a warp holds a hit, so shading never idles lanes on
//                        misses.
//   resolve_kernel       accumulators -> float image and RGB8 ((unsigned char)(min(1,c)*255), src/main.cpp:96).
//
// Accumulation across kernels is in signed 64-bit fixed point (2^-32): integer atomics commute, so the image is
// bit-identical from run to run and for any GPU count, whatever order the queues were filled in.
#pragma once
#include "skr_device.cuh"

// 128-thread CTAs, 7 per SM: 72 registers per thread, 28 resident warps.  Measured against 256 x 4 (64 registers, spills
// in the fog variant) and 256 x 3 / 128 x 6 (80 registers): ~2 % faster on configs 2, 3 and 5.
#ifndef SKR_BLOCK
#define SKR_BLOCK 128
#endif
#ifndef SKR_MIN_BLOCKS
#define SKR_MIN_BLOCKS 7
#endif
#define SKR_FIX_SCALE 4294967296.0f
#ifndef SKR_GI_BATCH
// GI children traced together (even): they share the per-sphere origin terms AND one queue reservation.  Measured on
// B200 (config 3 / config 5): 2: 6.94 / 180.1 ms, 4: 6.57 / 169.0 ms, 8: 7.51 / 190.1 ms (spills).
#define SKR_GI_BATCH 4
#endif

struct Queue
{
	// An entry is a HIT AWAITING SHADING, stored as the ray that found it: the consumer recomputes the hit point
	// P = o + d * t with the reference's expression (sphere_t_ref) -- every lane of a consumer warp holds an entry, while
	// in the producer only the 15-30 % of lanes that hit something would run that code.
	float4 *a;		 // (ray origin.xyz, bits(global pixel id))
	float4 *b;		 // (throughput.xyz, bits(node id))
	uint32_t *c;	 // sample | sphere << 16
	float4 *d;		 // (ray direction.xyz, t of the intersection loop = sphere_t_ref's fallback)
	unsigned *count; // device counter
	unsigned cap;
};

// n / d for any 32-bit n by a multiply-high and two shifts (Granlund & Montgomery's round-up form; constants from
// make_fastdiv on the host): the frame's divisors (tile size, tiles per row, world size, width) are run-time values, and a
// hardware-less integer division costs ~25 instructions (32-bit) to ~100 (64-bit) per thread.
struct FastDiv
{
	uint32_t m, s1, s2;
};
SKR_DEV uint32_t fdiv(uint32_t n, const FastDiv f)
{
	const uint32_t t = __umulhi(f.m, n);
	return (t + ((n - t) >> f.s1)) >> f.s2;
}

struct FrameParams
{
	int width, height;
	int tile, tiles_x, tiles_total, rank, world, wpr; // wpr = tile / 8 (warps per tile row-block)
	FastDiv fd_tpix, fd_wpr, fd_tiles_x, fd_tile, fd_world, fd_width, fd_peer_rows;
	int fast; // 1: local pixel indices fit 32 bits (any frame but absurdly thin ones), fdiv instead of 64-bit divisions
	int grid, spp;
	int max_depth, gi, n_gi, shadows, fresnel;
	float angle, aspect, inv_w, inv_h;
	int cull;		  // 1: bundle culling (cull_pairs) for this frame: jittered, >= 4 samples per pixel, <= 64 spheres
	float cull_delta; // bound on |d(r) - d(0.5)| over the jitter draw r of a pixel (ray directions are un-normalised)
	int split;		  // 1: the two halves of a pixel's samples are traced by two neighbouring warps (primary_kernel)
	int strip_words;  // 1: whole 8 x 4 blocks leave as 32-bit words (width % 4 == 0, 4-byte aligned frames)
	int strip_cta;	  // 1: ... and the four blocks of a CTA (a 32 x 4 strip of a 32 x 32 tile) leave together, 96 B per row (write_strip)
	// Launch order of this rank's tiles (single-kernel frames): global tile index per local slot, tiles that can see a sphere
	// FIRST, so that the tail of the kernel is made of cheap sky tiles (tiles_local entries; padding slots = tiles_total), or null
	const int *tile_order;
	uint2 key;
	uint32_t node_base, slot_gi;
	uint8_t *rgb8;	 // row-major frame or null
	float *rgb32;	 // row-major frame or null
	uint8_t *tiles8; // compact tile-major buffer or null
	uint8_t *peers[8]; // skr_render_peers_device: row-major RGB8 frames (one per GPU of the box, peer-mapped) or null
	int n_peers;
	int peer_rows; // 0: every pixel goes to ALL peers[]; > 0: to peers[min(y / peer_rows, n_peers - 1)] only (skr_render_bands_device)
	// Copy-out overlapped with the kernel (skr_render, single-kernel frames): the frame leaves in bands of whole tile rows.
	// Band k (in LAUNCH order: the host launches the bands heaviest first, tile_launch_order) owns the blocks
	// [band_start[k], band_start[k + 1]); the block that completes a band publishes band_seq in band_flag[k], which a
	// stream-ordered wait on the copy stream is parked on.  Null when unused.
	unsigned *band_count, *band_flag;
	unsigned band_start[9]; // (MAX_BANDS + 1)
	unsigned n_bands, band_seq;
	// Deferred triangle query (single-sample frames over a real hierarchy, see tri_deferred_kernel): candidates = camera rays
	// whose line reaches the hierarchy under the root; (direction, tmax) + local pixel index, counter pair (count, CTAs done)
	float4 *cand_d;
	uint2 *cand_px;		  // (y * width + x, index into the compact tile buffer)
	unsigned *cand_count; // [0] candidates, [1] CTAs done, [2] fetch cursor of tri_deferred_kernel
	int defer;
	long long *accum; // SKR_ACC_STRIDE per local pixel (gi / fresnel frames only)
	unsigned long long *counters; // 9 device counters (STATS)
	int *err;
};

struct PixelId
{
	int x, y;
	bool valid;
};

// local pixel index -> image coordinates.  Local order: tile by tile (local tile j = global tile j*world + rank),
// inside a tile warp by warp, each warp an 8x4 pixel block (coherent primary rays).
SKR_DEV PixelId decode_pixel(const FrameParams &fp, long long lp)
{
	const int tpix = fp.tile * fp.tile;
	PixelId p;
	if(fp.fast)
	{
		const uint32_t l  = (uint32_t) lp;
		const uint32_t lt = fdiv(l, fp.fd_tpix);
		const uint32_t r  = l - lt * (uint32_t) tpix;
		const uint32_t w = r >> 5, lane = r & 31u;
		const uint32_t wy = fdiv(w, fp.fd_wpr), wx = w - wy * (uint32_t) fp.wpr;
		const int px = (int) (wx * 8u + (lane & 7u)), py = (int) (wy * 4u + (lane >> 3));
		const uint32_t gt = fp.tile_order ? (uint32_t) __ldg(fp.tile_order + lt) : lt * (uint32_t) fp.world + (uint32_t) fp.rank;
		const uint32_t ty = fdiv(gt, fp.fd_tiles_x), tx = gt - ty * (uint32_t) fp.tiles_x;
		p.x				  = (int) tx * fp.tile + px;
		p.y				  = (int) ty * fp.tile + py;
		p.valid			  = gt < (uint32_t) fp.tiles_total && p.x < fp.width && p.y < fp.height;
		return p;
	}
	const int lt   = (int) (lp / tpix);
	const int r	   = (int) (lp - (long long) lt * tpix);
	const int w = r >> 5, lane = r & 31;
	const int wx = w % fp.wpr, wy = w / fp.wpr;
	const int px = wx * 8 + (lane & 7), py = wy * 4 + (lane >> 3);
	// local tile -> global tile: interleaved over the ranks, visited in the order of fp.tile_order when the host supplied one
	const long long gt = fp.tile_order ? (long long) __ldg(fp.tile_order + lt) : (long long) lt * fp.world + fp.rank;
	p.valid = gt < fp.tiles_total;
	const int tx = (int) (gt % fp.tiles_x), ty = (int) (gt / fp.tiles_x);
	p.x = tx * fp.tile + px;
	p.y = ty * fp.tile + py;
	p.valid = p.valid && p.x < fp.width && p.y < fp.height;
	return p;
}
// image coordinates -> local pixel index (inverse of the above; the pixel must belong to this rank)
SKR_DEV long long encode_pixel(const FrameParams &fp, int x, int y)
{
	const int tx = (int) fdiv((uint32_t) x, fp.fd_tile), ty = (int) fdiv((uint32_t) y, fp.fd_tile);
	const int px = x - tx * fp.tile, py = y - ty * fp.tile;
	const long long gt = (long long) ty * fp.tiles_x + tx;
	const long long lt = (long long) fdiv((uint32_t) gt, fp.fd_world); // gt < tiles_total, an int
	const int w		   = (py >> 2) * fp.wpr + (px >> 3);
	const int lane	   = ((py & 3) << 3) | (px & 7);
	return lt * (long long) (fp.tile * fp.tile) + (w << 5) + lane;
}

// image coordinates -> pixel index in the compact tile-major RGB8 buffer (tiles8): slot of the tile on this rank, row-major inside
SKR_DEV size_t tile_slot_pixel(const FrameParams &fp, int x, int y)
{
	const uint32_t tx = fdiv((uint32_t) x, fp.fd_tile), ty = fdiv((uint32_t) y, fp.fd_tile);
	const uint32_t lt = fdiv(ty * (uint32_t) fp.tiles_x + tx, fp.fd_world); // whatever the launch order
	const uint32_t px = (uint32_t) x - tx * (uint32_t) fp.tile, py = (uint32_t) y - ty * (uint32_t) fp.tile;
	return (size_t) lt * (size_t) (fp.tile * fp.tile) + (size_t) (py * (uint32_t) fp.tile + px);
}
SKR_DEV int peer_of_row(const FrameParams &fp, int y)
{
	return min((int) fdiv((uint32_t) y, fp.fd_peer_rows), fp.n_peers - 1);
}

// Accumulators (--gillum / fresnel frames): per local pixel 3 x int64 fixed point (2^-32) + one flags word.  Integer
// atomics commute, so the frame is bit-identical whatever order the queues were filled in.  Contributions that fixed
// point cannot hold -- NaN (`--gillum 0` divides by zero paths, src/raytrace.h:133), +-inf, |v| >= 2^16 -- are recorded
// OUT OF BAND in the flags word (bit ch: NaN, bit 3 + ch: +huge, bit 6 + ch: -huge) and add nothing to the sum, so any
// number of them per pixel resolves like the reference's float sum would: NaN stays NaN (-> byte 255, std::min(1, NaN)
// = 1), +inf stays +inf, +inf - inf = NaN.
#define SKR_ACC_STRIDE 4
SKR_DEV long long to_fixed(float v, unsigned &flags, int ch)
{
	if(!(fabsf(v) < 65536.0f))
	{
		flags |= (v != v ? 1u : (v > 0.0f ? 8u : 64u)) << ch;
		return 0;
	}
	return __float2ll_rn(v * SKR_FIX_SCALE);
}
SKR_DEV float from_fixed(long long a, unsigned flags, int ch)
{
	const unsigned f = flags >> ch;
	if((f & 1u) || ((f & 8u) && (f & 64u)))
	{
		return CUDART_NAN_F;
	}
	if(f & 8u)
	{
		return CUDART_INF_F;
	}
	if(f & 64u)
	{
		return -CUDART_INF_F;
	}
	return (float) ((double) a * (1.0 / 4294967296.0));
}
SKR_DEV void accum_store(long long *accum, long long lp, float3 c) // first writer of the pixel (primary_kernel)
{
	unsigned fl		= 0;
	long long *a	= accum + SKR_ACC_STRIDE * lp;
	const long long x = to_fixed(c.x, fl, 0), y = to_fixed(c.y, fl, 1), z = to_fixed(c.z, fl, 2);
	reinterpret_cast<longlong2 *>(a)[0] = make_longlong2(x, y);
	reinterpret_cast<longlong2 *>(a)[1] = make_longlong2(z, (long long) fl);
}
// (ex, ey, ez, efl): fixed-point terms already summed elsewhere (leaf hits shaded in place), folded into the same atomics
SKR_DEV void accum_add(long long *accum, long long lp, float3 c, long long ex = 0, long long ey = 0, long long ez = 0, unsigned efl = 0)
{
	unsigned fl				= efl;
	unsigned long long *a	= reinterpret_cast<unsigned long long *>(accum + SKR_ACC_STRIDE * lp);
	const long long x = to_fixed(c.x, fl, 0) + ex, y = to_fixed(c.y, fl, 1) + ey, z = to_fixed(c.z, fl, 2) + ez;
	atomicAdd(a + 0, (unsigned long long) x);
	atomicAdd(a + 1, (unsigned long long) y);
	atomicAdd(a + 2, (unsigned long long) z);
	if(fl)
	{
		atomicOr(a + 3, (unsigned long long) fl);
	}
}
SKR_DEV float3 accum_load(const long long *accum, long long lp)
{
	const longlong2 p = reinterpret_cast<const longlong2 *>(accum + SKR_ACC_STRIDE * lp)[0];
	const longlong2 q = reinterpret_cast<const longlong2 *>(accum + SKR_ACC_STRIDE * lp)[1];
	const unsigned fl = (unsigned) q.y;
	return f3(from_fixed(p.x, fl, 0), from_fixed(p.y, fl, 1), from_fixed(q.x, fl, 2));
}

SKR_DEV uint8_t quantise(float c) // (unsigned char)(std::min(float(1), c) * 255), src/main.cpp:96
{
	const float m = c < 1.0f ? c : 1.0f;
	return (uint8_t) (__float2int_rz(__fmul_rn(m, 255.0f)) & 0xff);
}

SKR_DEV void write_pixel(const FrameParams &fp, long long lp, const PixelId &p, float3 c)
{
	if(fp.rgb32)
	{
		float *o = fp.rgb32 + 3 * ((size_t) p.y * fp.width + p.x);
		o[0]	 = c.x;
		o[1]	 = c.y;
		o[2]	 = c.z;
	}
	if(fp.rgb8)
	{
		uint8_t *o = fp.rgb8 + 3 * ((size_t) p.y * fp.width + p.x);
		o[0]	   = quantise(c.x);
		o[1]	   = quantise(c.y);
		o[2]	   = quantise(c.z);
	}
	if(fp.n_peers > 0)
	{
		// frame split without a collective: the finished pixel goes straight into every GPU's frame; the stores to the
		// peers travel over NVLink while the rest of the kernel is still tracing
		const uint8_t r = quantise(c.x), g = quantise(c.y), b = quantise(c.z);
		const size_t at = 3 * ((size_t) p.y * fp.width + p.x);
		const int k0 = fp.peer_rows > 0 ? peer_of_row(fp, p.y) : 0;
		const int k1 = fp.peer_rows > 0 ? k0 + 1 : fp.n_peers;
		for(int k = k0; k < k1; k++)
		{
			uint8_t *o = fp.peers[k] + at;
			o[0]	   = r;
			o[1]	   = g;
			o[2]	   = b;
		}
	}
	if(fp.tiles8)
	{
		uint8_t *o = fp.tiles8 + 3 * tile_slot_pixel(fp, p.x, p.y);
		o[0]	   = quantise(c.x);
		o[1]	   = quantise(c.y);
		o[2]	   = quantise(c.z);
	}
}

// One finished 8 x 4 pixel block (all 32 lanes of the warp call this; lane = pixel of the block, lane 0 its origin).
// Frames bound for another device or for page-locked host memory (fp.strip_words) leave as 32-bit words: the block is
// quantised into 96 B of shared memory (`stage`, this warp's), then 24 lanes store one word each, 24 B per pixel row --
// instead of 3 byte stores per pixel across NVLink / PCIe.  Ragged blocks and local frames go pixel by pixel.
SKR_DEV void write_block(const FrameParams &fp, long long lp, const PixelId &p, float3 c, uint32_t *stage)
{
	const unsigned lane = threadIdx.x & 31u;
	const int x0 = __shfl_sync(0xffffffffu, p.x, 0), y0 = __shfl_sync(0xffffffffu, p.y, 0);
	const bool words = fp.strip_words && __shfl_sync(0xffffffffu, (int) p.valid, 0) && x0 + 8 <= fp.width; // (uniform)
	if(!words)
	{
		if(p.valid)
		{
			write_pixel(fp, lp, p, c);
		}
		return;
	}
	uint8_t *o = reinterpret_cast<uint8_t *>(stage) + (lane >> 3) * 24 + (lane & 7) * 3;
	o[0]	   = quantise(c.x);
	o[1]	   = quantise(c.y);
	o[2]	   = quantise(c.z);
	if(fp.rgb32 && p.valid)
	{
		float *f = fp.rgb32 + 3 * ((size_t) p.y * fp.width + p.x);
		f[0]	 = c.x;
		f[1]	 = c.y;
		f[2]	 = c.z;
	}
	if(fp.tiles8 && p.valid)
	{
		uint8_t *t = fp.tiles8 + 3 * tile_slot_pixel(fp, p.x, p.y);
		t[0] = o[0], t[1] = o[1], t[2] = o[2];
	}
	__syncwarp();
	if(lane < 24)
	{
		const int r = lane / 6, w = lane - r * 6;
		if(y0 + r < fp.height)
		{
			const size_t at	 = (((size_t) (y0 + r) * fp.width + x0) * 3) / 4 + w;
			const uint32_t v = stage[r * 6 + w];
			if(fp.rgb8)
			{
				reinterpret_cast<uint32_t *>(fp.rgb8)[at] = v;
			}
			const int k0 = fp.peer_rows > 0 ? peer_of_row(fp, y0 + r) : 0;
			const int k1 = fp.peer_rows > 0 ? k0 + 1 : fp.n_peers;
			for(int k = k0; k < k1; k++)
			{
				reinterpret_cast<uint32_t *>(fp.peers[k])[at] = v;
			}
		}
	}
	__syncwarp();
}

// The same for the FOUR blocks of a CTA at once (fp.strip_cta: 32 x 32 tiles, one block per warp, so a CTA covers a 32 x 4 strip
// of its tile): every warp quantises its block into its 96 B of `stage4`, and the LAST warp of the CTA to get there stores
// the strip -- 4 rows x 96 B, 32-byte aligned, whole sectors -- where four warps would each store 4 x 24 B that straddle
// sector boundaries.  That matters for frames stored straight into page-locked HOST memory (skr_render): partial-sector
// writes cross PCIe one small packet each.  Strips that are not wholly inside the image go block by block as before.
SKR_DEV void write_strip(const FrameParams &fp, long long lp, const PixelId &p, float3 c, uint32_t (*stage4)[24], unsigned *done)
{
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const int x0 = __shfl_sync(0xffffffffu, p.x, 0), y0 = __shfl_sync(0xffffffffu, p.y, 0);
	const int xs = x0 - 8 * (int) warp; // origin of the strip: the CTA's warps are neighbours along x
	const bool whole = __shfl_sync(0xffffffffu, (int) p.valid, 0) && xs >= 0 && xs + 32 <= fp.width && y0 + 4 <= fp.height; // (uniform in the CTA)
	if(!whole)
	{
		write_block(fp, lp, p, c, stage4[warp]);
		return;
	}
	uint8_t *o = reinterpret_cast<uint8_t *>(stage4[warp]) + (lane >> 3) * 24 + (lane & 7) * 3;
	o[0]	   = quantise(c.x);
	o[1]	   = quantise(c.y);
	o[2]	   = quantise(c.z);
	__syncwarp();
	unsigned arrived = 0;
	if(lane == 0)
	{
		__threadfence_block();
		arrived = atomicAdd(done, 1u);
	}
	arrived = __shfl_sync(0xffffffffu, arrived, 0);
	if(arrived != SKR_BLOCK / 32 - 1)
	{
		return;
	}
	__threadfence_block();
#pragma unroll
	for(int k = (int) lane; k < 96; k += 32) // word k of the strip: row k / 24, word k % 24 of the row = word (k % 24) % 6 of block (k % 24) / 6
	{
		const int r = k / 24, w = k - r * 24;
		const uint32_t v = stage4[w / 6][r * 6 + w % 6];
		const size_t at	 = (((size_t) (y0 + r) * fp.width + xs) * 3) / 4 + w;
		if(fp.rgb8)
		{
			reinterpret_cast<uint32_t *>(fp.rgb8)[at] = v;
		}
		const int k0 = fp.peer_rows > 0 ? peer_of_row(fp, y0 + r) : 0;
		const int k1 = fp.peer_rows > 0 ? k0 + 1 : fp.n_peers;
		for(int q = k0; q < k1; q++)
		{
			reinterpret_cast<uint32_t *>(fp.peers[q])[at] = v;
		}
	}
}

// Stage the scene blob into shared memory (all threads of the CTA).  SMEM is a template parameter so that the test
// loops compile to LDS (not generic loads) in the common case; scenes too big for shared memory read the blob in place.
template <bool SMEM>
SKR_DEV const float4 *stage_scene(const SceneView &sv, float4 *smem)
{
	if constexpr(!SMEM)
	{
		return sv.blob;
	}
	else
	{
		for(int i = threadIdx.x; i < sv.blob_f4; i += blockDim.x)
		{
			smem[i] = __ldg(sv.blob + i);
		}
		__syncthreads();
		return smem;
	}
}

template <bool STATS>
SKR_DEV void flush_counters(const FrameParams &fp, Counters &c)
{
	if(!STATS)
	{
		return;
	}
	unsigned v[9] = {c.ch, c.sh, c.st, c.stp, c.tt, c.nv, c.hits, c.le, c.se};
#pragma unroll
	for(int k = 0; k < 9; k++)
	{
		unsigned s = v[k];
#pragma unroll
		for(int off = 16; off > 0; off >>= 1)
		{
			s += __shfl_xor_sync(0xffffffffu, s, off);
		}
		if((threadIdx.x & 31) == 0 && s)
		{
			atomicAdd(fp.counters + k, (unsigned long long) s);
		}
	}
}

// Warp-aggregated push.  queue_reserve: ONE atomic per warp for `count` entries (all 32 lanes call it, count uniform);
// queue_store: one entry.
SKR_DEV unsigned queue_reserve(const Queue &q, unsigned count)
{
	unsigned base = 0;
	if(count == 0)
	{
		return 0;
	}
	if((threadIdx.x & 31) == 0)
	{
		base = atomicAdd(q.count, count);
	}
	return __shfl_sync(0xffffffffu, base, 0);
}
SKR_DEV void queue_store(const Queue &q, unsigned idx, float3 o, uint32_t pixel, float3 thr, uint32_t node, uint32_t sample, int sphere, float3 dir, float t,
						 int *err)
{
	if(idx < q.cap)
	{
		q.a[idx] = make_float4(o.x, o.y, o.z, u2f(pixel));
		q.b[idx] = make_float4(thr.x, thr.y, thr.z, u2f(node));
		q.c[idx] = (sample & 0xffffu) | ((uint32_t) sphere << 16);
		q.d[idx] = make_float4(dir.x, dir.y, dir.z, t);
	}
	else
	{
		atomicOr(err, 1); // cannot happen: the host sizes chunks so that a full fan-out fits
	}
}
// one entry per lane that wants one.  Must be called by all 32 lanes.
SKR_DEV void queue_push(const Queue &q, bool want, float3 o, uint32_t pixel, float3 thr, uint32_t node, uint32_t sample, int sphere, float3 dir, float t,
						int *err)
{
	const unsigned mask = __ballot_sync(0xffffffffu, want);
	const unsigned base = queue_reserve(q, (unsigned) __popc(mask));
	if(want)
	{
		queue_store(q, base + __popc(mask & ((1u << (threadIdx.x & 31)) - 1u)), o, pixel, thr, node, sample, sphere, dir, t, err);
	}
}
// candidate of the deferred triangle query; must be called by all 32 lanes
SKR_DEV void cand_push(const FrameParams &fp, bool want, long long lp, const PixelId &p, float3 d, float tmax)
{
	const unsigned mask = __ballot_sync(0xffffffffu, want);
	if(mask == 0u)
	{
		return;
	}
	unsigned base = 0;
	if((threadIdx.x & 31) == 0)
	{
		base = atomicAdd(fp.cand_count, (unsigned) __popc(mask));
	}
	base = __shfl_sync(0xffffffffu, base, 0);
	if(want)
	{
		const unsigned idx = base + __popc(mask & ((1u << (threadIdx.x & 31)) - 1u));
		fp.cand_d[idx]	   = make_float4(d.x, d.y, d.z, tmax);
		fp.cand_px[idx]	   = make_uint2((uint32_t) (p.y * fp.width + p.x), (uint32_t) tile_slot_pixel(fp, p.x, p.y)); // (deferral needs npix_local < 2^32)
	}
}
// consumer side: the hit point of an entry, src/raytrace.h:197-204 (exact t, then P = o + d * t)
SKR_DEV float3 queue_hit_point(const float4 *__restrict__ B, const SceneView &sv, const float4 &qa, const float4 &qd, int sidx)
{
	const float3 o = f3(qa), d = f3(qd);
	const float t  = sphere_t_ref(o, d, f3(B[sv.off_geom + sidx]), B[sv.off_spec + sidx].w, qd.w);
	return add_rn(o, muls_rn(d, t));
}

// ------------------------------------------------------------------------------------------------
// primary_kernel: local pixels [lp0, lp0 + npix)
// ------------------------------------------------------------------------------------------------
// Template flags: GI (--gillum: push hits instead of shading), STATS (device counters), SMEM (scene blob staged in shared
// memory), TRIS (scene has triangles: BVH code compiled in), FOG (scene has spherical fog: fog shading compiled in).
// Sphere-only, fog-free scenes thus run a kernel without the traversal stack or the fog branch in its register budget.
//
// Launch shape: one 8 x 4 pixel block per WARP, four warps per CTA (the scene blob is staged per CTA).
// (Persistent CTAs whose warps fetch blocks from a device counter were measured and dropped: config 4 0.42 ms against 0.31,
// config 1 0.062 against 0.057 -- the 64 800 fetches of a 1080p frame serialise on one L2 address.  DESIGN.md.)
// Finished pixels bound for another device or for page-locked host memory (skr_render_peers_device) are quantised into
// shared memory and leave as 32-bit words, 24 B per pixel row of the block, instead of 3 byte stores each.
// HALVES: frames with >= 8 samples per pixel (see below); single-sample frames run the lean single-loop variant.
template <bool GI, bool STATS, bool SMEM, bool TRIS, bool FOG, bool HALVES>
__global__ void __launch_bounds__(SKR_BLOCK, SKR_MIN_BLOCKS) primary_kernel(const SceneView sv, const FrameParams fp, const Queue q0, long long lp0, long long npix)
{
	extern __shared__ float4 smem[];
	__shared__ uint32_t s_px[SKR_BLOCK / 32][24]; // per warp: one block of RGB8, 4 rows x 24 B
	__shared__ unsigned s_done;					  // warps of this CTA whose block is staged (write_strip)
	if(threadIdx.x == 0)
	{
		s_done = 0u;
	}
	const float4 *B = stage_scene<SMEM>(sv, smem);
	if(!SMEM && fp.strip_cta)
	{
		__syncthreads();
	}
	Counters cnt;
	zero(cnt);
	__shared__ float s_part[HALVES ? 3 : 1][HALVES ? SKR_BLOCK : 1]; // first-half sums (see below)
	const unsigned lane	   = threadIdx.x & 31u;
	const unsigned nblocks = (unsigned) ((npix + 31) / 32);
	// The grid covers the frame: warp w of CTA b takes the 8 x 4 pixel block b * (warps per CTA) + w.
	// HALVES: a pixel's samples are summed as TWO halves, [0, h) and [h, n), added at the end -- by one warp, or (fp.split: frames
	// with too few blocks to fill the GPU for long, i.e. one rank's share at world >= 4) by two neighbouring warps of the
	// CTA that take one half each: twice the work items, half the length of the kernel's tail.  Same arithmetic either
	// way, so the frame does not depend on the split (bit-identical for any world size, tested).
	const unsigned nparts = (HALVES && !GI && fp.split) ? 2u : 1u;
	const unsigned wg	  = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const unsigned blk = wg / nparts, part = wg % nparts;
	const bool in_range = blk < nblocks;
	long long lp		= 0;
	PixelId p;
	p.x = p.y = 0;
	p.valid	  = false;
	float3 sum = f3(0.0f, 0.0f, 0.0f);
	if(in_range)
	{
	const long long g = (long long) blk * 32 + lane;
	lp				  = lp0 + g;
	p				  = decode_pixel(fp, lp);
	p.valid			  = p.valid && g < npix;

	RngCtx rng;
	rng.pixel = (uint32_t) (p.y * fp.width + p.x);
	rng.node  = 0;
	rng.key	  = fp.key;

	const int nsamples = fp.max_depth > 0 ? fp.spp : 0; // depth <= 0: shade() returns black (src/raytrace.h:142-145)
	uint4 jit = make_uint4(0u, 0u, 0u, 0u);

	// Bundle culling (skr_device.cuh, cull_pairs): the pixel's samples are lines through the camera within cull_delta of
	// the ray of r = 0.5.  The warp tests the union of its 32 pixels' surviving pairs, so the loop stays uniform.
	const bool cull	 = fp.cull != 0;
	uint32_t pmask	 = 0;
	float3 pc		 = f3(0.0f, 0.0f, 0.0f); // shadow bundle: hit points within sqrt(rho2) of pc use smask
	float rho2		 = -1.0f;
	uint64_t smask	 = 0;
	if(cull)
	{
		const float uc = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(2.0f, __fmul_rn(__fadd_rn((float) p.x, 0.5f), fp.inv_w)), 1.0f), fp.angle), fp.aspect);
		const float vc = __fmul_rn(__fsub_rn(1.0f, __fmul_rn(2.0f, __fmul_rn(__fadd_rn((float) p.y, 0.5f), fp.inv_h))), fp.angle);
		const float3 dc = add_rn(add_rn(sv.cam_dir, muls_rn(sv.cam_right, uc)), muls_rn(sv.cam_up, vc));
		const float len = sqrtf(dot(dc, dc));
		const int NP	= sv.S4 >> 1;
		uint32_t mk		= NP >= 32 ? 0xffffffffu : (1u << NP) - 1u;
		if(fp.cull_delta <= 0.45f * len)
		{
			mk = cull_pairs<STATS>(B + sv.off_cull, NP, sv.S, dc, __fdividef(1.05f * fp.cull_delta, len), 0.0f, cnt);
		}
		pmask = __reduce_or_sync(0xffffffffu, p.valid ? mk : 0u);
	}
	const int half = HALVES ? (nsamples + 1) >> 1 : nsamples;
#pragma unroll 1
	for(int range = 0; range < (HALVES ? 2 : 1); range++)
	{
	if(HALVES && nparts == 2u && (unsigned) range != part)
	{
		continue;
	}
	const int s_begin = range == 0 ? 0 : half, s_end = range == 0 ? half : nsamples;
	if(HALVES && range == 1 && nparts == 1u)
	{
		// (one warp does both halves: park the first half's sum in shared memory instead of three more live registers)
		s_part[0][threadIdx.x] = sum.x, s_part[1][threadIdx.x] = sum.y, s_part[2][threadIdx.x] = sum.z;
		sum					   = f3(0.0f, 0.0f, 0.0f);
	}
	for(int s = s_begin; s < s_end; s++)
	{
		rng.sample = (uint32_t) s;
		float u, v;
		if(fp.grid > 0)
		{
			// src/main.cpp:52-54: ONE draw for both axes, all-float arithmetic.  Four consecutive samples
			// share one Philox block.
			if((s & 3) == 0 || s == s_begin)
			{
				jit = philox4x32_10(make_uint4(rng.pixel, (uint32_t) s >> 2, 0u, 0u), rng.key);
			}
			const uint32_t jw = (s & 3) == 0 ? jit.x : (s & 3) == 1 ? jit.y : (s & 3) == 2 ? jit.z : jit.w;
			const float r	  = rng_unit(jw);
			u = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(2.0f, __fmul_rn(__fadd_rn((float) p.x, r), fp.inv_w)), 1.0f), fp.angle), fp.aspect);
			v = __fmul_rn(__fsub_rn(1.0f, __fmul_rn(2.0f, __fmul_rn(__fadd_rn((float) p.y, r), fp.inv_h))), fp.angle);
		}
		else
		{
			// src/main.cpp:73-74: evaluated in double (x + 0.5 promotes), stored to float
			u = (float) __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(2.0, __dmul_rn((double) p.x + 0.5, (double) fp.inv_w)), 1.0), (double) fp.angle),
								  (double) fp.aspect);
			v = (float) __dmul_rn(__dsub_rn(1.0, __dmul_rn(2.0, __dmul_rn((double) p.y + 0.5, (double) fp.inv_h))), (double) fp.angle);
		}
		// src/main.cpp:76-77: D + u*R + v*U, never normalised
		const float3 d = add_rn(add_rn(sv.cam_dir, muls_rn(sv.cam_right, u)), muls_rn(sv.cam_up, v));
		const float3 o = sv.cam_pos;

		float t = 0.0f;
		int h	= -3;
		bool cand = false;
		if(p.valid)
		{
			h = closest_hit<true, STATS, TRIS>(B, sv, o, d, t, cnt, cull, pmask, (TRIS && !GI && fp.defer) ? &cand : nullptr);
		}
		if(TRIS && !GI && fp.defer)
		{
			cand_push(fp, cand, lp, p, d, t); // settled by tri_deferred_kernel: the pixel is written below as if no triangle were hit
		}
#ifdef SKR_DEBUG_NV
		if(STATS && !GI) // debugging aid (never in a shipped build): the frame shows node visits / leaf tests of each pixel's query
		{
			sum = f3((float) cnt.nv, (float) cnt.tt, 0.0f);
			cnt.nv = cnt.tt = 0;
			h = -3;
		}
#endif
		if(h == -2)
		{
			sum += sv.background;
		}
		else if(!GI && h >= 0)
		{
			const float3 c	= f3(B[sv.off_geom + h]);
			t				= sphere_t_ref(o, d, c, B[sv.off_spec + h].w, t);
			const float3 hp = add_rn(o, muls_rn(d, t)); // src/raytrace.h:204
			{
				const float3 n	  = normalize_fast(sub_rn(hp, c)); // feeds shading terms only (no --gillum here)
				const bool smcull = cull && fp.shadows != 0 && sv.cull_shadow != 0;
				if(smcull)
				{
					const float3 dp = hp - pc;
					if(!(dot(dp, dp) <= rho2))
					{
						// new bundle around this hit.  Radius: the footprint of the pixel's jitter diagonal on the surface,
						// t * 2 delta / cos(incidence), with slack; samples that still fall outside start another bundle.
						const float ci	= fabsf(dot(n, d)) * rsqrtf(dot(d, d));
						const float rho = __fdividef(2.5f * fp.cull_delta * t, fmaxf(ci, 0.05f));
						pc				= hp;
						rho2			= rho * rho;
						smask			= shadow_masks<STATS>(B, sv, hp, rho, cnt);
					}
				}
				sum += direct_light<STATS, FOG, true>(B, sv, fp.shadows != 0, rng, h, hp, n, cnt, smcull, smask);
			}
		}
		if(GI)
		{
			queue_push(q0, h >= 0, o, rng.pixel, f3(1.0f, 1.0f, 1.0f), 0u, (uint32_t) s, h, d, t, fp.err);
		}
	}
	} // the two halves
	} // in_range
	// first half + second half
	if(!HALVES)
	{
	}
	else if(nparts == 2u)
	{
		if(in_range && part == 1u)
		{
			s_part[0][threadIdx.x - 32] = sum.x, s_part[1][threadIdx.x - 32] = sum.y, s_part[2][threadIdx.x - 32] = sum.z; // the partner warp's slots
		}
		__syncthreads();
		if(in_range && part == 0u)
		{
			sum = f3(__fadd_rn(sum.x, s_part[0][threadIdx.x]), __fadd_rn(sum.y, s_part[1][threadIdx.x]), __fadd_rn(sum.z, s_part[2][threadIdx.x]));
		}
	}
	else if(in_range)
	{
		sum = f3(__fadd_rn(s_part[0][threadIdx.x], sum.x), __fadd_rn(s_part[1][threadIdx.x], sum.y), __fadd_rn(s_part[2][threadIdx.x], sum.z));
	}
	if(in_range && part == 0u)
	{
	if(GI)
	{
		if(p.valid)
		{
			accum_store(fp.accum, lp, sum);
		}
	}
	else
	{
		if(fp.grid > 0)
		{
			const float n2 = (float) fp.spp; // image[y][x] /= (grid*grid), src/main.cpp:68
			sum			   = f3(__fdiv_rn(sum.x, n2), __fdiv_rn(sum.y, n2), __fdiv_rn(sum.z, n2));
		}
		if(fp.strip_cta)
		{
			write_strip(fp, lp, p, sum, s_px, &s_done);
		}
		else
		{
			write_block(fp, lp, p, sum, s_px[threadIdx.x >> 5]);
		}
	}
	if(!GI && fp.band_flag)
	{
		// overlapped copy-out (skr_render): blocks are counted per band of whole tile rows; the last one publishes the flag
		__syncwarp();
		if(lane == 0)
		{
			// (device scope is enough for the blocks that only count: the publishing block's system fence below is cumulative)
			__threadfence();
			unsigned b = 0;
			while(b + 1u < fp.n_bands && blk >= fp.band_start[b + 1u])
			{
				b++;
			}
			const unsigned want = fp.band_start[b + 1u] - fp.band_start[b];
			if((atomicAdd(fp.band_count + b, 1u) + 1u) % want == 0u) // counters run on from frame to frame (same geometry)
			{
				__threadfence_system();
				atomicExch(fp.band_flag + b, fp.band_seq);
			}
		}
	}
	}
	flush_counters<STATS>(fp, cnt);
}

// ------------------------------------------------------------------------------------------------
// Leaf hits shaded in place (shade_expand_kernel<..., LEAF = true>).
//
// The children of a depth-2 hit have depth 1: the reference shades them and recurses no further (src/raytrace.h:142-145
// under :212).  They are the bulk of a --gillum tree (config 5: 90 % of all sphere hits).  Instead of a round trip
// through a global-memory queue (52 B written + read per hit, three 64-bit global atomics per hit) the warp that found
// them shades them itself -- but never with the 15-30 % of lanes that happen to hold a hit: the hits of the warp's 32
// parents are compacted into a per-warp staging ring in shared memory (ballot + popc), and as soon as 32 are pending
// ALL lanes shade one each.  A leaf's contribution is added, in fixed point, to its parent's accumulator in shared
// memory (integer adds commute -> the frame stays bit-identical to the queued path, run to run and for any chunking);
// the parent's lane folds that into its ONE global atomic triple.
//
// Per-warp staging: SKR_LEAF_SLOTS x 2 float4   slot = (child direction, loop t), (weight w = T kd 2pi r1 / n, meta)
//                                               meta = sphere | parent lane << 16 | child index << 21
//                   32 x 4 x u64                per parent lane: fixed-point sum r, g, b + flags
#define SKR_LEAF_SLOTS (32 + 32 * SKR_GI_BATCH)
#define SKR_LEAF_WARP_BYTES (SKR_LEAF_SLOTS * 32 + 32 * 4 * 8)
#define SKR_LEAF_CTA_BYTES ((SKR_BLOCK / 32) * SKR_LEAF_WARP_BYTES)
#define SKR_LEAF_MAX_CHILDREN 2047

struct LeafStage
{
	float4 *slot;			 // this warp's ring
	unsigned long long *acc; // this warp's 32 x 4 parent accumulators
	unsigned pending;		 // warp-uniform
};

// Fold one round's fixed-point terms into the parents' accumulators in shared memory: 64-bit shared-memory atomics
// (ATOMS.CAST.SPIN loops).  Measured alternatives, all slower on B200 and kept out of the tree: MATCH.ANY +
// REDUX.SUM per parent group (config 5: 215 ms against 145), native 32-bit atomics on 16/16/32-bit limbs (144.4 against
// 144.7: no gain), parent-major staging + integer prefix scan with one atomic triple per run (150 against 146).
SKR_DEV void leaf_accumulate(const LeafStage &ls, bool act, int src, long long fx, long long fy, long long fz, unsigned fl)
{
	if(act)
	{
		unsigned long long *a = ls.acc + 4 * src;
		atomicAdd(a + 0, (unsigned long long) fx);
		atomicAdd(a + 1, (unsigned long long) fy);
		atomicAdd(a + 2, (unsigned long long) fz);
		if(fl)
		{
			atomicOr(a + 3, (unsigned long long) fl);
		}
	}
}
// the parent lane's total, as three fixed-point terms + flags
SKR_DEV void leaf_total(const LeafStage &ls, unsigned lane, long long &x, long long &y, long long &z, unsigned &fl)
{
	const unsigned long long *a = ls.acc + 4 * lane;
	x = (long long) a[0], y = (long long) a[1], z = (long long) a[2], fl = (unsigned) a[3];
}

template <bool STATS, bool FOG>
SKR_DEV void leaf_shade_round(const float4 *__restrict__ B, const SceneView &sv, const FrameParams &fp, const LeafStage &ls, unsigned first, unsigned count,
							  float3 o, const RngCtx &rng, Counters &cnt)
{
	const unsigned lane = threadIdx.x & 31u;
	const bool act		= lane < count;
	const unsigned at	= first + (act ? lane : 0u);
	const float4 s0 = ls.slot[2 * at], s1 = ls.slot[2 * at + 1];
	const uint32_t meta = f2u(s1.w);
	const int sidx		= (int) (meta & 0xffffu);
	const int src		= (int) ((meta >> 16) & 31u);
	// the parent's ray origin and RNG coordinates live in the parent lane's registers
	const float3 po = f3(__shfl_sync(0xffffffffu, o.x, src), __shfl_sync(0xffffffffu, o.y, src), __shfl_sync(0xffffffffu, o.z, src));
	RngCtx r2;
	r2.pixel  = __shfl_sync(0xffffffffu, rng.pixel, src);
	r2.sample = __shfl_sync(0xffffffffu, rng.sample, src);
	r2.node	  = __shfl_sync(0xffffffffu, rng.node, src) * fp.node_base + (meta >> 21) + 1u;
	r2.key	  = fp.key;
	long long fx = 0, fy = 0, fz = 0;
	unsigned fl = 0;
	if(act)
	{
		// exactly the arithmetic of a queued leaf entry (queue_hit_point + the shade-only branch of shade_expand_kernel)
		const float4 g	= B[sv.off_geom + sidx];
		const float3 d	= f3(s0);
		const float t	= sphere_t_ref(po, d, f3(g), B[sv.off_spec + sidx].w, s0.w);
		const float3 hp = add_rn(po, muls_rn(d, t));
		const float3 n	= normalize_rn(sub_rn(hp, f3(g)));
		const float3 kd = f3(B[sv.off_diff + sidx]);
		const float3 direct	 = direct_light<STATS, FOG, false>(B, sv, fp.shadows != 0, r2, sidx, hp, n, cnt);
		const float3 contrib = f3(s1) * kd * (direct * 0.318309886183790672f);
		fl = 0;
		fx = to_fixed(contrib.x, fl, 0), fy = to_fixed(contrib.y, fl, 1), fz = to_fixed(contrib.z, fl, 2);
	}
	leaf_accumulate(ls, act, src, fx, fy, fz, fl);
}

// shade whole rounds of 32 while that many are pending; keep the rest at the front of the ring
template <bool STATS, bool FOG>
SKR_DEV void leaf_drain(const float4 *__restrict__ B, const SceneView &sv, const FrameParams &fp, LeafStage &ls, float3 o, const RngCtx &rng, Counters &cnt,
						bool flush)
{
	__syncwarp();
	const unsigned whole = ls.pending & ~31u;
	for(unsigned first = 0; first < whole; first += 32)
	{
		leaf_shade_round<STATS, FOG>(B, sv, fp, ls, first, 32u, o, rng, cnt);
	}
	const unsigned left = ls.pending - whole;
	if(flush)
	{
		if(left)
		{
			leaf_shade_round<STATS, FOG>(B, sv, fp, ls, whole, left, o, rng, cnt);
		}
		ls.pending = 0;
		__syncwarp();
		return;
	}
	if(whole && left)
	{
		const unsigned lane = threadIdx.x & 31u;
		float4 a = make_float4(0, 0, 0, 0), b = a;
		if(lane < left)
		{
			a = ls.slot[2 * (whole + lane)];
			b = ls.slot[2 * (whole + lane) + 1];
		}
		__syncwarp();
		if(lane < left)
		{
			ls.slot[2 * lane]	  = a;
			ls.slot[2 * lane + 1] = b;
		}
	}
	ls.pending = left;
	__syncwarp();
}

// ------------------------------------------------------------------------------------------------
// shade_expand_kernel: queue entries [start, start + count) of `in`
// ------------------------------------------------------------------------------------------------
// LEAF: the children are leaves of the tree (depth 1) and are shaded in place, see above; `out` is unused.
template <bool STATS, bool SMEM, bool TRIS, bool FOG, bool LEAF>
__global__ void __launch_bounds__(SKR_BLOCK, SKR_MIN_BLOCKS) shade_expand_kernel(const SceneView sv, const FrameParams fp, const Queue in, unsigned start, unsigned count,
																  const Queue out, int expand)
{
	extern __shared__ float4 smem[];
	const float4 *B = stage_scene<SMEM>(sv, smem);
	Counters cnt;
	zero(cnt);
	LeafStage ls;
	if constexpr(LEAF)
	{
		char *base = reinterpret_cast<char *>(smem) + (SMEM ? (size_t) sv.blob_f4 * sizeof(float4) : 0) + (threadIdx.x >> 5) * SKR_LEAF_WARP_BYTES;
		ls.slot	   = reinterpret_cast<float4 *>(base);
		ls.acc	   = reinterpret_cast<unsigned long long *>(base + SKR_LEAF_SLOTS * 32);
		ls.pending = 0;
		const unsigned lane = threadIdx.x & 31u;
		reinterpret_cast<ulonglong2 *>(ls.acc + 4 * lane)[0] = make_ulonglong2(0ull, 0ull);
		reinterpret_cast<ulonglong2 *>(ls.acc + 4 * lane)[1] = make_ulonglong2(0ull, 0ull);
		__syncwarp();
	}

	const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = g < count;
	const unsigned i = start + (valid ? g : 0u);

	const float4 qa	  = in.a[i];
	const float4 qb	  = in.b[i];
	const uint32_t qc = in.c[i];
	const float4 qd	  = in.d[i];
	const float3 thr  = f3(qb);
	const int sidx	  = (int) (qc >> 16);
	const float3 hp	  = queue_hit_point(B, sv, qa, qd, sidx);
	RngCtx rng;
	rng.pixel  = f2u(qa.w);
	rng.sample = qc & 0xffffu;
	rng.node   = f2u(qb.w);
	rng.key	   = fp.key;

	float3 contrib = f3(0.0f, 0.0f, 0.0f);
	const float3 n = normalize_rn(sub_rn(hp, f3(B[sv.off_geom + sidx]))); // src/raytrace.h:205
	const float3 kd = f3(B[sv.off_diff + sidx]);
	if(valid)
	{
		const float3 direct = direct_light<STATS, FOG, false>(B, sv, fp.shadows != 0, rng, sidx, hp, n, cnt);
		// with --gillum: (direct / pi) * kd (src/raytrace.h:213); fresnel-only trees return direct as is (:103, :218)
		contrib = fp.gi ? thr * kd * (direct * 0.318309886183790672f) : thr * direct;
		if(fp.gi && fp.n_gi == 0)
		{
			// `--gillum 0`: the reference divides its empty sum by zero paths (total_colour /= num_rays, src/raytrace.h:133),
			// so every shaded hit is NaN and the pixel quantises to 255.  Reproduced as such.
			contrib = f3(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
		}
	}
	if(expand)
	{
		// src/raytrace.h:107-136: acc += r1 * shade(child) / (1/pi); acc /= n; combine 2 * acc * kd
		float3 nt, nb;
		basis_from_normal(n, nt, nb);
		const float3 tk = thr * kd * (6.28318530717958648f / (float) fp.n_gi);
		const float3 o	= adds_rn(hp, 0.00001f);
		// one child: weight; a miss adds the background, a hit is pushed as (o, d, t) for the next level to finish
		const auto finish = [&](float r1, int h, float3 &w) {
			w = tk * r1;
			if(h == -2)
			{
				contrib += w * sv.background;
			}
		};
		// children are traced SKR_GI_BATCH at a time: they share the per-sphere origin terms of their intersection tests
		// (children 2m and 2m+1 also share a Philox block) and ONE queue reservation
		int c = 0;
		for(; c + SKR_GI_BATCH <= fp.n_gi; c += SKR_GI_BATCH)
		{
			float3 d[SKR_GI_BATCH];
			float r1[SKR_GI_BATCH], t[SKR_GI_BATCH];
			int h[SKR_GI_BATCH];
#pragma unroll
			for(int k = 0; k < SKR_GI_BATCH; k += 2)
			{
				const uint4 r = rng_block(rng, fp.slot_gi + ((uint32_t) (c + k) >> 1));
				r1[k]		  = rng_unit(r.x);
				d[k]		  = gi_child_dir(r1[k], rng_unit(r.y), n, nt, nb);
				r1[k + 1]	  = rng_unit(r.z);
				d[k + 1]	  = gi_child_dir(r1[k + 1], rng_unit(r.w), n, nt, nb);
				t[k] = t[k + 1] = 0.0f;
				h[k] = h[k + 1] = -3;
			}
			if(valid)
			{
				closest_hit_xk<SKR_GI_BATCH, STATS, TRIS>(B, sv, o, d, t, h, cnt);
			}
			float3 w[SKR_GI_BATCH];
			unsigned m[SKR_GI_BATCH], total = 0;
#pragma unroll
			for(int k = 0; k < SKR_GI_BATCH; k++)
			{
				finish(r1[k], h[k], w[k]);
				m[k] = __ballot_sync(0xffffffffu, h[k] >= 0);
				total += (unsigned) __popc(m[k]);
			}
			if constexpr(LEAF)
			{
				unsigned at = ls.pending;
#pragma unroll
				for(int k = 0; k < SKR_GI_BATCH; k++)
				{
					if(h[k] >= 0)
					{
						const unsigned idx	 = at + __popc(m[k] & ((1u << (threadIdx.x & 31)) - 1u));
						ls.slot[2 * idx]	 = make_float4(d[k].x, d[k].y, d[k].z, t[k]);
						ls.slot[2 * idx + 1] = make_float4(w[k].x, w[k].y, w[k].z, u2f((uint32_t) h[k] | ((threadIdx.x & 31u) << 16) | ((uint32_t) (c + k) << 21)));
					}
					at += (unsigned) __popc(m[k]);
				}
				ls.pending = at;
				if(at >= 32u)
				{
					leaf_drain<STATS, FOG>(B, sv, fp, ls, o, rng, cnt, false);
				}
			}
			else
			{
				unsigned at = queue_reserve(out, total);
#pragma unroll
				for(int k = 0; k < SKR_GI_BATCH; k++)
				{
					if(h[k] >= 0)
					{
						queue_store(out, at + __popc(m[k] & ((1u << (threadIdx.x & 31)) - 1u)), o, rng.pixel, w[k], rng.node * fp.node_base + (uint32_t) (c + k) + 1u,
									rng.sample, h[k], d[k], t[k], fp.err);
					}
					at += (unsigned) __popc(m[k]);
				}
			}
		}
		uint4 r = make_uint4(0u, 0u, 0u, 0u);
		for(; c < fp.n_gi; c++) // remainder, one by one
		{
			if((c & 1) == 0)
			{
				r = rng_block(rng, fp.slot_gi + ((uint32_t) c >> 1));
			}
			const float r1 = rng_unit((c & 1) ? r.z : r.x), r2 = rng_unit((c & 1) ? r.w : r.y);
			const float3 d = gi_child_dir(r1, r2, n, nt, nb);
			float t		   = 0.0f;
			int h		   = -3;
			if(valid)
			{
				h = closest_hit<false, STATS, TRIS>(B, sv, o, d, t, cnt);
			}
			float3 w;
			finish(r1, h, w);
			if constexpr(LEAF)
			{
				const unsigned m = __ballot_sync(0xffffffffu, h >= 0);
				if(h >= 0)
				{
					const unsigned idx	 = ls.pending + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
					ls.slot[2 * idx]	 = make_float4(d.x, d.y, d.z, t);
					ls.slot[2 * idx + 1] = make_float4(w.x, w.y, w.z, u2f((uint32_t) h | ((threadIdx.x & 31u) << 16) | ((uint32_t) c << 21)));
				}
				ls.pending += (unsigned) __popc(m);
				if(ls.pending >= 32u)
				{
					leaf_drain<STATS, FOG>(B, sv, fp, ls, o, rng, cnt, false);
				}
			}
			else
			{
				queue_push(out, h >= 0, o, rng.pixel, w, rng.node * fp.node_base + (uint32_t) c + 1u, rng.sample, h, d, t, fp.err);
			}
		}
		if constexpr(LEAF)
		{
			leaf_drain<STATS, FOG>(B, sv, fp, ls, o, rng, cnt, true);
		}
	}
	if(valid)
	{
		const int y = (int) fdiv(rng.pixel, fp.fd_width), x = (int) (rng.pixel - (uint32_t) y * (uint32_t) fp.width);
		const long long lp = encode_pixel(fp, x, y);
		if constexpr(LEAF)
		{
			long long ex, ey, ez;
			unsigned efl;
			leaf_total(ls, threadIdx.x & 31u, ex, ey, ez, efl);
			accum_add(fp.accum, lp, contrib, ex, ey, ez, efl);
		}
		else
		{
			accum_add(fp.accum, lp, contrib);
		}
	}
	flush_counters<STATS>(fp, cnt);
}

// ------------------------------------------------------------------------------------------------
// fresnel_expand_kernel (opt-in mode, skr_options.fresnel): the recursion HEAD never reaches because
// direct_illumination returns at src/raytrace.h:44 -- lines :46-103.  Runs over the same queue chunk as
// shade_expand_kernel and pushes into the same next-level queue.  Per hit with a specular material:
//   fr = bp::fresnel(ray.direction, N)                                   src/blinn_phong.h:156-184
//   per light (point, then directional):
//     fr < 1: refraction_colour  = fr * shade(P, bp::refraction(d, N))   ASSIGNMENT: only the last light's survives
//             reflection_colour += (1 - fr) * ks * shade(P, bp::reflect_direction(Lhat, N))   (reflects the LIGHT direction)
//   rays start exactly at P (the 1.0 near cutoff keeps them off their own sphere).
// ------------------------------------------------------------------------------------------------
SKR_DEV float fresnel_ref(float3 dir, float3 n, float ior_in)
{
	float cos_i = fminf(fmaxf(dot_rn(dir, n), -1.0f), 1.0f);
	float et = 1.0f, ior = ior_in;
	if(cos_i > 0.0f)
	{
		et	= ior_in;
		ior = 1.0f;
	}
	const float sint = __fmul_rn(__fdiv_rn(et, ior), __fsqrt_rn(fmaxf(0.0f, __fsub_rn(1.0f, __fmul_rn(cos_i, cos_i)))));
	if(sint >= 1.0f)
	{
		return 1.0f; // total internal reflection
	}
	const float cos_t = __fsqrt_rn(fmaxf(0.0f, __fsub_rn(1.0f, __fmul_rn(sint, sint))));
	cos_i			  = fabsf(cos_i);
	const float den	  = __fadd_rn(__fmul_rn(ior, cos_i), __fmul_rn(et, cos_t)); // the reference uses this denominator for Rs AND Rp
	const float rs	  = __fdiv_rn(__fsub_rn(__fmul_rn(ior, cos_i), __fmul_rn(et, cos_t)), den);
	const float rp	  = __fdiv_rn(__fsub_rn(__fmul_rn(et, cos_i), __fmul_rn(ior, cos_t)), den);
	return __fdiv_rn(__fadd_rn(__fmul_rn(rs, rs), __fmul_rn(rp, rp)), 2.0f);
}
SKR_DEV float3 refraction_ref(float3 dir, float3 n, float ior) // src/blinn_phong.h:143-153
{
	const float dn = dot_rn(dir, n);
	const float k  = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(ior, ior), __fsub_rn(1.0f, __fmul_rn(dn, dn))));
	if(k < 0.0f)
	{
		return f3(0.0f, 0.0f, 0.0f);
	}
	return sub_rn(muls_rn(dir, ior), muls_rn(n, __fadd_rn(__fmul_rn(ior, dn), __fsqrt_rn(k))));
}
SKR_DEV float3 reflect_direction_ref(float3 l, float3 n) // src/blinn_phong.h:137-140
{
	return normalize_rn(sub_rn(l, muls_rn(n, __fmul_rn(2.0f, dot_rn(l, n)))));
}

template <bool STATS, bool SMEM>
__global__ void __launch_bounds__(SKR_BLOCK) fresnel_expand_kernel(const SceneView sv, const FrameParams fp, const Queue in, unsigned start, unsigned count,
																	const Queue out)
{
	extern __shared__ float4 smem[];
	const float4 *B = stage_scene<SMEM>(sv, smem);
	Counters cnt;
	zero(cnt);
	const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = g < count;
	const unsigned i = start + (valid ? g : 0u);
	const float4 qa = in.a[i], qb = in.b[i], qd = in.d[i];
	const uint32_t qc = in.c[i];
	const float3 thr = f3(qb), dir = f3(qd);
	const int sidx	 = (int) (qc >> 16);
	const float3 hp	 = queue_hit_point(B, sv, qa, qd, sidx);
	const uint32_t pixel = f2u(qa.w), node = f2u(qb.w), sample = qc & 0xffffu;
	const float3 n	= normalize_rn(sub_rn(hp, f3(B[sv.off_geom + sidx])));
	const float3 kd = f3(B[sv.off_diff + sidx]);
	const float3 ks = f3(B[sv.off_spec + sidx]);
	const float ior = B[sv.off_diff + sidx].w;
	const bool spec = valid && (ks.x != 0.0f || ks.y != 0.0f || ks.z != 0.0f);
	const float fr	= fresnel_ref(dir, n, ior);
	// weight of this node's direct_illumination result in the pixel: T (.) kd / pi under --gillum, T otherwise
	const float3 wn = fp.gi ? thr * kd * 0.318309886183790672f : thr;
	float3 contrib	= f3(0.0f, 0.0f, 0.0f);
	const int nl	= sv.L + sv.D;
	for(int li = 0; li < nl; li++)
	{
		const float3 lhat = li < sv.L ? normalize_rn(sub_rn(f3(B[sv.off_plpos + li]), hp)) : f3(B[sv.off_dldir + (li - sv.L)]);
#pragma unroll
		for(int kind = 0; kind < 2; kind++) // 0: refraction (only the last light's result survives), 1: reflection
		{
			const bool want = spec && (kind == 1 || (fr < 1.0f && li == nl - 1));
			const float3 d	= kind == 0 ? refraction_ref(dir, n, ior) : reflect_direction_ref(lhat, n);
			const float3 w	= kind == 0 ? wn * fr : wn * ks * (1.0f - fr);
			float t			= 0.0f;
			int h			= -3;
			if(want)
			{
				h = closest_hit<false, STATS, true>(B, sv, hp, d, t, cnt);
			}
			if(h == -2)
			{
				contrib += w * sv.background;
			}
			queue_push(out, h >= 0, hp, pixel, w, node * fp.node_base + (uint32_t) (fp.n_gi + 2 * li + kind) + 1u, sample, h, d, t, fp.err);
		}
	}
	if(valid && (contrib.x != 0.0f || contrib.y != 0.0f || contrib.z != 0.0f))
	{
		const int y = (int) fdiv(pixel, fp.fd_width), x = (int) (pixel - (uint32_t) y * (uint32_t) fp.width);
		const long long lp = encode_pixel(fp, x, y);
		accum_add(fp.accum, lp, contrib);
	}
	flush_counters<STATS>(fp, cnt);
}

// ------------------------------------------------------------------------------------------------
// tri_deferred_kernel: the triangle half of shade() (src/raytrace.h:169-186) for the CANDIDATES primary_kernel set aside.
//
// What bounds a frame like dragon.scn's: 2.6 % of the camera rays -- lines that thread the model without an
// accepted hit -- walk 100-640 nodes each, and walked one ray per lane the frame lasts as long as its heaviest warp's
// serial chain (640 iterations x ~100 instructions x ~9 cycles).  primary_kernel therefore walks at most SKR_DEFER_STEPS
// nodes in place (enough for the ~97 % of rays that only graze the model's bounds) and hands the rest over as ONE dense
// list; here TEAMS of SKR_TEAM lanes walk one ray each from a shared stack in shared memory: every iteration the team pops
// up to SKR_TEAM pending nodes, one per lane, tests their children (leaf triangles at once) and pushes the internal
// children back (ballot + popc give each lane its slot).  A heavy ray's walk is SKR_TEAM times shorter; teams fetch
// their next ray from the list as they finish (persistent CTAs, one warp-aggregated atomic per refill), so nothing waits
// for the slowest ray but the end of the kernel.  Any accepted triangle blackens the pixel primary_kernel wrote
// (src/raytrace.h:221-224); the walk order is irrelevant to an any-hit query: same frame as the walk in place, bit for bit.
// ------------------------------------------------------------------------------------------------
#ifndef SKR_TEAM
// measured on B200, config 4 (walk in place 0.279 ms): teams of 4 0.245 ms, of 8 0.230, of 16 0.263; one pool of (ray, node)
// items per warp instead of team stacks (every lane pops any ray's node): 0.28-0.34
#define SKR_TEAM 8
#endif
#define SKR_TEAM_STACK 128
#ifndef SKR_DEFER_CTAS
// resident CTAs per SM (register budget) = persistent CTAs launched per SM.  Config 4 on B200: 6: 0.235 ms, 8: 0.218, 10: 0.230
// (48 registers), 12: 0.257, 16: 0.316 -- the kernel keeps the L1 tag stage 71 % busy (each lane loads its own 64 B node:
// ~20 distinct lines per warp load), so more resident teams add spills and evict each other's nodes rather than hide latency
#define SKR_DEFER_CTAS 8
#endif
SKR_DEV void write_black(const FrameParams &fp, uint2 px) // the pixel of a candidate whose line hit a triangle
{
	const size_t at = 3 * (size_t) px.x;
	if(fp.rgb32)
	{
		fp.rgb32[at] = fp.rgb32[at + 1] = fp.rgb32[at + 2] = 0.0f;
	}
	if(fp.rgb8)
	{
		fp.rgb8[at] = fp.rgb8[at + 1] = fp.rgb8[at + 2] = 0;
	}
	{
		const int y	 = (int) fdiv(px.x, fp.fd_width);
		const int k0 = fp.peer_rows > 0 ? peer_of_row(fp, y) : 0;
		const int k1 = fp.peer_rows > 0 ? k0 + 1 : fp.n_peers;
		for(int k = k0; k < k1; k++)
		{
			uint8_t *o = fp.peers[k] + at;
			o[0] = o[1] = o[2] = 0;
		}
	}
	if(fp.tiles8)
	{
		uint8_t *o = fp.tiles8 + 3 * (size_t) px.y;
		o[0] = o[1] = o[2] = 0;
	}
}
template <bool STATS>
__global__ void __launch_bounds__(SKR_BLOCK, SKR_DEFER_CTAS) tri_deferred_kernel(const SceneView sv, const FrameParams fp)
{
	constexpr int T		= SKR_TEAM;
	constexpr int TEAMS = 32 / T;
	__shared__ int s_stack[SKR_BLOCK / 32][TEAMS][SKR_TEAM_STACK];
	Counters cnt;
	zero(cnt);
	const unsigned count = *reinterpret_cast<volatile unsigned *>(fp.cand_count);
	const unsigned lane	 = threadIdx.x & 31u;
	const unsigned team = lane / T, j = lane % T;
	const unsigned team_mask = (T == 32 ? 0xffffffffu : ((1u << T) - 1u)) << (team * T); // this team's lanes
	const unsigned team_lt	 = team_mask & ((1u << lane) - 1u);							  // ... below this one
	int *stk = s_stack[threadIdx.x >> 5][team];
	bool active = false; // (uniform within a team)
	bool more	= true;	 // (uniform within the warp) the list may still hold rays
	int sp		= 0;	 // (uniform within a team)
	uint2 px	= make_uint2(0u, 0u);
	TriWalk wk;
	wk.o = wk.d = wk.inv = f3(0.0f, 0.0f, 0.0f);
	wk.tmax = wk.dlen = 0.0f;
	wk.node = wk.sp = 0;
	for(;;)
	{
		const unsigned idle = __ballot_sync(0xffffffffu, !active);
		if(more && idle != 0u)
		{
			// REFILL: the idle teams take the next rays of the list (one fetch per warp)
			const unsigned n_idle = (unsigned) __popc(idle) / T;
			unsigned base		  = 0;
			if(lane == 0)
			{
				base = atomicAdd(fp.cand_count + 2, n_idle);
			}
			base = __shfl_sync(0xffffffffu, base, 0);
			more = base + n_idle < count;
			if(!active)
			{
				const unsigned g = base + (unsigned) __popc(idle & ((1u << (team * T)) - 1u)) / T;
				if(g < count)
				{
					const float4 c = __ldg(fp.cand_d + g);
					px			   = __ldg(fp.cand_px + g);
					tri_walk_begin(wk, sv.cam_pos, f3(c), c.w);
					if(j == 0)
					{
						stk[0] = 0; // the root
					}
					sp	   = 1;
					active = true;
				}
			}
			__syncwarp();
		}
		if(__ballot_sync(0xffffffffu, active) == 0u)
		{
			if(!more)
			{
				break;
			}
			continue;
		}
		// every lane of an active team takes one pending node (as far as there are any); when the stack is nearly full only one
		int take = sp < T ? sp : T;
		if(sp + take > SKR_TEAM_STACK - 2)
		{
			take = 1;
		}
		const bool have = active && (int) j < take;
		const int node	= have ? stk[sp - 1 - (int) j] : 0;
		__syncwarp(); // all reads of the stacks before any write
		int kid[2];
		int nk	 = 0;
		bool hit = false;
		if(have)
		{
			hit = tri_node_visit<STATS>(sv, wk, node, kid, nk, cnt);
		}
		const unsigned b1 = __ballot_sync(0xffffffffu, nk >= 1), b2 = __ballot_sync(0xffffffffu, nk == 2), bh = __ballot_sync(0xffffffffu, hit);
		if(active)
		{
			const int below = sp - take;
			const int off	= below + __popc(b1 & team_lt) + __popc(b2 & team_lt);
			const int total = __popc(b1 & team_mask) + __popc(b2 & team_mask);
			if(below + total > SKR_TEAM_STACK)
			{
				atomicOr(sv.err, 2); // cannot happen below a depth of ~60 with the guard above; reported, never silent
			}
			else
			{
				if(nk >= 1)
				{
					stk[off] = kid[0];
				}
				if(nk == 2)
				{
					stk[off + 1] = kid[1];
				}
			}
			sp = below + total;
			if(bh & team_mask)
			{
				if(j == 0)
				{
					write_black(fp, px);
				}
				active = false;
			}
			else if(sp == 0)
			{
				active = false;
			}
		}
		__syncwarp();
	}
	__syncthreads();
	if(threadIdx.x == 0)
	{
		// every CTA has read the count and made its last fetch: the last one to leave re-arms the counters for the next frame
		__threadfence();
		if(atomicAdd(fp.cand_count + 1, 1u) == gridDim.x - 1u)
		{
			fp.cand_count[0] = 0u;
			fp.cand_count[1] = 0u;
			fp.cand_count[2] = 0u;
			__threadfence();
		}
	}
	flush_counters<STATS>(fp, cnt);
}

__global__ void __launch_bounds__(SKR_BLOCK) resolve_kernel(const FrameParams fp, long long lp0, long long npix)
{
	__shared__ uint32_t s_px[SKR_BLOCK / 32][24];
	const long long g  = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	const long long lp = lp0 + g;
	PixelId p		   = decode_pixel(fp, lp);
	p.valid			   = p.valid && g < npix;
	float3 c		   = f3(0.0f, 0.0f, 0.0f);
	if(p.valid)
	{
		c = accum_load(fp.accum, lp);
		if(fp.grid > 0)
		{
			const float n2 = (float) fp.spp;
			c			   = f3(__fdiv_rn(c.x, n2), __fdiv_rn(c.y, n2), __fdiv_rn(c.z, n2));
		}
	}
	write_block(fp, lp, p, c, s_px[threadIdx.x >> 5]); // a warp = one 8 x 4 pixel block, like primary_kernel
}

// skr_deinterleave_device: gathered rank-major compact tiles -> row-major frame
__global__ void deinterleave_kernel(const uint8_t *__restrict__ gathered, uint8_t *__restrict__ rgb8, int width, int height, int tile, int tiles_x,
									int world, long long tiles_per_rank)
{
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= (long long) width * height)
	{
		return;
	}
	const int x = (int) (i % width), y = (int) (i / width);
	const int tx = x / tile, ty = y / tile;
	const long long gt = (long long) ty * tiles_x + tx;
	const long long r = gt % world, lt = gt / world;
	const size_t src = ((size_t) (r * tiles_per_rank + lt) * tile * tile + (size_t) (y - ty * tile) * tile + (x - tx * tile)) * 3;
	rgb8[3 * i + 0]	 = gathered[src + 0];
	rgb8[3 * i + 1]	 = gathered[src + 1];
	rgb8[3 * i + 2]	 = gathered[src + 2];
}
