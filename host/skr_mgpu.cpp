// host/skr_mgpu.cpp -- implementation of include/skr_mgpu.h (libskr_mgpu.so): one process, one host thread and one
// skr_ctx per GPU.  Frame assembly: peer stores over NVLink straight into GPU 0's frame when every GPU can map it
// (no collective, no de-interleave pass), else one NCCL all-gather of the finished RGB8 tiles.  See the header.
#include "../include/skr_mgpu.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace
{
thread_local std::string g_err;
}

// One persistent host thread per GPU (a --gillum frame schedules its wavefront with host read-backs, so every GPU needs
// its own driver thread; spawning them per frame cost more than a 0.2 ms frame).
struct Worker
{
	std::thread th;
	std::mutex m;
	std::condition_variable cv;
	std::function<void()> job;
	bool has_job = false, quit = false;
	int taken = 0;				// value of `posted` when the worker last took a job
	std::atomic<int> posted{0}; // bumped with every job: workers spin on it briefly before they sleep (frames are ~0.2 ms apart)
};

struct skr_mgpu
{
	int world = 0;
	std::vector<Worker *> workers;
	std::mutex done_m;
	std::condition_variable done_cv;
	int pending = 0;
	// the caller's frame, page-locked and mapped so that every GPU's kernel can store its tiles straight into it
	void *pinned_host = nullptr;
	size_t pinned_bytes = 0;
	bool pinned_by_us = false;
	void *pinned_dev = nullptr;
	bool direct = true; // kernels may store into the host frame (SKR_MGPU_NO_DIRECT=1: assembly via GPU 0 / NCCL as in round 1)
	// default frame assembly where every GPU pair has peer access: row bands (render_bands)
	bool bands = false;
	std::vector<uint8_t *> d_band; // per GPU: a frame-sized buffer of which only the GPU's own band of rows is used
	std::vector<size_t> cap_band;
	std::vector<skr_ctx *> ctx;
	std::vector<ncclComm_t> comm;
	std::vector<uint8_t *> d_tiles;	   // per GPU: its compact tiles
	std::vector<uint8_t *> d_gathered; // per GPU: all ranks' tiles (all-gather result)
	std::vector<size_t> cap_tiles;
	uint8_t *d_frame = nullptr; // GPU 0
	size_t cap_frame = 0;
	bool p2p = false; // every GPU has GPU 0's memory mapped (cudaDeviceEnablePeerAccess)
	std::string err;
};

namespace
{
int fail(skr_mgpu *m, int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	(m ? m->err : g_err) = buf;
	return code;
}
void worker_main(skr_mgpu *m, int i)
{
	cudaSetDevice(i);
	Worker *w = m->workers[i];
	for(;;)
	{
		std::function<void()> job;
		{
			// spin for up to ~200 us (a condition-variable wake-up costs 20-50 us, a fifth of a frame), then sleep
			const auto t0 = std::chrono::steady_clock::now();
			while(w->posted.load(std::memory_order_acquire) == w->taken && std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200))
			{
			}
			std::unique_lock<std::mutex> lk(w->m);
			w->cv.wait(lk, [w] { return w->has_job || w->quit; });
			if(w->quit)
			{
				return;
			}
			job		   = std::move(w->job);
			w->has_job = false;
			w->taken   = w->posted.load(std::memory_order_acquire);
		}
		job();
		{
			std::lock_guard<std::mutex> lk(m->done_m);
			m->pending--;
		}
		m->done_cv.notify_one();
	}
}
// runs f(i) on every GPU's thread and waits for all of them
void on_all(skr_mgpu *m, const std::function<void(int)> &f)
{
	{
		std::lock_guard<std::mutex> lk(m->done_m);
		m->pending = m->world;
	}
	for(int i = 0; i < m->world; i++)
	{
		Worker *w = m->workers[i];
		{
			std::lock_guard<std::mutex> lk(w->m);
			w->job	   = [&f, i] { f(i); };
			w->has_job = true;
		}
		w->posted.fetch_add(1, std::memory_order_release);
		w->cv.notify_one();
	}
	// the caller spins briefly too, then sleeps
	{
		const auto t0 = std::chrono::steady_clock::now();
		while(std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(300))
		{
			std::lock_guard<std::mutex> lk(m->done_m);
			if(m->pending == 0)
			{
				return;
			}
		}
	}
	std::unique_lock<std::mutex> lk(m->done_m);
	m->done_cv.wait(lk, [m] { return m->pending == 0; });
}

// Page-lock + map the caller's frame once (cached while the pointer and size stay the same); buffers that are already
// page-locked (cudaHostAlloc, torch pin_memory) are used as they are.
int pin_frame(skr_mgpu *m, uint8_t *rgb8, size_t bytes)
{
	if(m->pinned_host == rgb8 && m->pinned_bytes >= bytes && m->pinned_dev)
	{
		return SKR_OK;
	}
	cudaSetDevice(0);
	if(m->pinned_by_us && m->pinned_host)
	{
		cudaHostUnregister(m->pinned_host);
	}
	m->pinned_host = nullptr;
	m->pinned_dev  = nullptr;
	void *d		   = nullptr;
	const int rc   = skr_pin_host(m->ctx[0], rgb8, bytes, &d);
	if(rc != SKR_OK && rc != 1000)
	{
		return rc;
	}
	m->pinned_host	= rgb8;
	m->pinned_bytes = bytes;
	m->pinned_by_us = rc == SKR_OK;
	m->pinned_dev	= d;
	return SKR_OK;
}

void sum_stats(skr_stats *stats, const std::vector<skr_stats> &st)
{
	memset(stats, 0, sizeof *stats);
	for(const skr_stats &s : st)
	{
		stats->closest_hit_rays += s.closest_hit_rays;
		stats->shadow_rays += s.shadow_rays;
		stats->sphere_tests += s.sphere_tests;
		stats->sphere_tests_pos += s.sphere_tests_pos;
		stats->tri_tests += s.tri_tests;
		stats->bvh_node_visits += s.bvh_node_visits;
		stats->sphere_hits += s.sphere_hits;
		stats->light_evals += s.light_evals;
		stats->sphere_tests_executed += s.sphere_tests_executed;
		stats->queue_entries += s.queue_entries;
		stats->kernel_launches += s.kernel_launches;
		stats->queue_chunks += s.queue_chunks;
		if(s.ms_total > stats->ms_total)
		{
			stats->ms_total	  = s.ms_total;
			stats->ms_primary = s.ms_primary;
			stats->ms_bounce  = s.ms_bounce;
			stats->ms_resolve = s.ms_resolve;
		}
	}
}

// Default frame assembly on a box whose GPUs all see each other (NVLink): ROW BANDS.  Every GPU renders its interleaved
// tiles (load balance) but stores each finished pixel block, over NVLink and while it is still tracing, into the memory of
// the GPU that owns the block's band of rows (skr_render_bands_device).  When all kernels are done every GPU holds one
// contiguous band of the frame and copies it to the host ITSELF: N copies over N PCIe links instead of one frame through
// GPU 0's.
int render_bands(skr_mgpu *m, const skr_options *opt, uint8_t *rgb8, skr_stats *stats, size_t frame_bytes)
{
	const int W = m->world;
	pin_frame(m, rgb8, frame_bytes); // (a pageable destination works too, only slower: the copies stage through the driver)
	for(int i = 0; i < W; i++)
	{
		if(m->cap_band[i] < frame_bytes)
		{
			cudaSetDevice(i);
			cudaFree(m->d_band[i]);
			m->d_band[i] = nullptr;
			if(cudaMalloc(&m->d_band[i], frame_bytes) != cudaSuccess)
			{
				m->cap_band[i] = 0;
				return fail(m, SKR_ERR_CUDA, "skr_mgpu_render: cudaMalloc(band buffer) failed on GPU %d", i);
			}
			m->cap_band[i] = frame_bytes;
		}
	}
	const int rows		   = ((opt->height + W - 1) / W + 3) / 4 * 4;
	const size_t row_bytes = (size_t) opt->width * 3;
	std::vector<int> rc(W, 0);
	std::vector<std::string> msg(W);
	std::vector<skr_stats> st(W);
	const bool want_stats = stats != nullptr;
	std::vector<void *> frames(m->d_band.begin(), m->d_band.end());
	std::atomic<int> arrived{0}, failed{0};
	on_all(m, [&](int i) {
		skr_options oi = *opt;
		oi.world	   = W;
		oi.rank		   = i;
		memset(&st[i], 0, sizeof st[i]);
		rc[i] = skr_render_bands_device(m->ctx[i], &oi, frames.data(), W, rows, want_stats ? &st[i] : nullptr);
		if(!rc[i])
		{
			rc[i] = skr_sync(m->ctx[i]);
		}
		if(rc[i])
		{
			msg[i] = skr_last_error(m->ctx[i]);
			failed.fetch_add(1);
		}
		// every GPU's kernel must be done before any band is copied out: the driver threads meet here (they all run: one per
		// GPU, all posted by on_all above)
		arrived.fetch_add(1, std::memory_order_acq_rel);
		while(arrived.load(std::memory_order_acquire) < W)
		{
			std::this_thread::yield();
		}
		if(failed.load() != 0)
		{
			return;
		}
		const size_t y0 = std::min((size_t) opt->height, (size_t) i * rows), y1 = std::min((size_t) opt->height, (size_t) (i + 1) * rows);
		if(y1 > y0)
		{
			rc[i] = skr_copy_to_host(m->ctx[i], rgb8 + y0 * row_bytes, m->d_band[i] + y0 * row_bytes, (y1 - y0) * row_bytes);
			if(!rc[i])
			{
				rc[i] = skr_sync(m->ctx[i]);
			}
			if(rc[i])
			{
				msg[i] = skr_last_error(m->ctx[i]);
			}
		}
	});
	for(int i = 0; i < W; i++)
	{
		if(rc[i])
		{
			return fail(m, rc[i], "GPU %d: %s", i, msg[i].c_str());
		}
	}
	if(stats)
	{
		sum_stats(stats, st);
		stats->ms_d2h = 0.0f;
	}
	return SKR_OK;
}

// Frame assembly without peer access: NO device frame at all.  Every GPU renders its interleaved tiles with skr_render_peers_device and
// its kernel stores each finished pixel block straight into the caller's page-locked host frame -- every GPU over its OWN
// PCIe link, while the rest of its tiles are still being traced.  When the GPUs' streams have drained the frame is whole.
int render_direct(skr_mgpu *m, const skr_options *opt, uint8_t *rgb8, skr_stats *stats, size_t frame_bytes)
{
	const int W = m->world;
	int rc0		= pin_frame(m, rgb8, frame_bytes);
	if(rc0)
	{
		return fail(m, rc0, "skr_mgpu_render: cannot page-lock the frame: %s", skr_last_error(m->ctx[0]));
	}
	std::vector<int> rc(W, 0);
	std::vector<std::string> msg(W);
	std::vector<skr_stats> st(W);
	const bool want_stats = stats != nullptr;
	on_all(m, [&](int i) {
		skr_options oi = *opt;
		oi.world	   = W;
		oi.rank		   = i;
		memset(&st[i], 0, sizeof st[i]);
		void *frames[1] = {m->pinned_dev};
		rc[i]			= skr_render_peers_device(m->ctx[i], &oi, frames, 1, want_stats ? &st[i] : nullptr);
		if(!rc[i])
		{
			rc[i] = skr_sync(m->ctx[i]);
		}
		if(rc[i])
		{
			msg[i] = skr_last_error(m->ctx[i]);
		}
	});
	for(int i = 0; i < W; i++)
	{
		if(rc[i])
		{
			return fail(m, rc[i], "GPU %d: %s", i, msg[i].c_str());
		}
	}
	if(stats)
	{
		sum_stats(stats, st);
		stats->ms_d2h = 0.0f; // the copy-out is the kernels' own stores
	}
	return SKR_OK;
}

// Collective-free frame: every GPU renders its interleaved tiles with skr_render_peers_device and stores each finished
// pixel into GPU 0's row-major frame (peer-mapped; the stores cross NVLink while the kernel is still tracing).  When all
// host threads have seen their kernels complete the frame is whole: one D2H copy from GPU 0.
int render_p2p(skr_mgpu *m, const skr_options *opt, uint8_t *rgb8, skr_stats *stats, size_t frame_bytes)
{
	const int W = m->world;
	cudaSetDevice(0);
	if(m->cap_frame < frame_bytes)
	{
		cudaFree(m->d_frame);
		m->d_frame = nullptr;
		if(cudaMalloc(&m->d_frame, frame_bytes) != cudaSuccess)
		{
			m->cap_frame = 0;
			return fail(m, SKR_ERR_CUDA, "skr_mgpu_render: cudaMalloc(frame) failed");
		}
		m->cap_frame = frame_bytes;
	}
	std::vector<int> rc(W, 0);
	std::vector<std::string> msg(W);
	std::vector<skr_stats> st(W);
	std::vector<std::thread> th;
	for(int i = 0; i < W; i++)
	{
		th.emplace_back([&, i]() {
			cudaSetDevice(i);
			skr_options oi = *opt;
			oi.world	   = W;
			oi.rank		   = i;
			memset(&st[i], 0, sizeof st[i]);
			void *frames[1] = {m->d_frame};
			rc[i]			= skr_render_peers_device(m->ctx[i], &oi, frames, 1, &st[i]);
			if(rc[i])
			{
				msg[i] = skr_last_error(m->ctx[i]);
			}
			else if(skr_sync(m->ctx[i]) != SKR_OK)
			{
				rc[i]  = SKR_ERR_CUDA;
				msg[i] = skr_last_error(m->ctx[i]);
			}
		});
	}
	for(std::thread &t : th)
	{
		t.join();
	}
	for(int i = 0; i < W; i++)
	{
		if(rc[i])
		{
			return fail(m, rc[i], "GPU %d: %s", i, msg[i].c_str());
		}
	}
	cudaSetDevice(0);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaStream_t s = (cudaStream_t) skr_stream(m->ctx[0]);
	cudaEventRecord(e0, s);
	cudaError_t ce = cudaMemcpyAsync(rgb8, m->d_frame, frame_bytes, cudaMemcpyDeviceToHost, s);
	cudaEventRecord(e1, s);
	if(ce == cudaSuccess)
	{
		ce = cudaStreamSynchronize(s);
	}
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	if(ce != cudaSuccess)
	{
		return fail(m, SKR_ERR_CUDA, "skr_mgpu_render: copying the frame out failed: %s", cudaGetErrorString(ce));
	}
	if(stats)
	{
		sum_stats(stats, st);
		stats->ms_d2h = ms;
	}
	return SKR_OK;
}
} // namespace

extern "C" {

const char *skr_mgpu_last_error(const skr_mgpu *m)
{
	return m ? m->err.c_str() : g_err.c_str();
}

int skr_mgpu_world(const skr_mgpu *m)
{
	return m ? m->world : 0;
}

void skr_mgpu_destroy(skr_mgpu *m)
{
	if(!m)
	{
		return;
	}
	for(Worker *w : m->workers)
	{
		{
			std::lock_guard<std::mutex> lk(w->m);
			w->quit = true;
		}
		w->cv.notify_one();
		if(w->th.joinable())
		{
			w->th.join();
		}
		delete w;
	}
	m->workers.clear();
	if(m->pinned_by_us && m->pinned_host && !m->ctx.empty())
	{
		skr_unpin_host(m->ctx[0], m->pinned_host);
	}
	for(int i = 0; i < (int) m->ctx.size(); i++)
	{
		cudaSetDevice(i);
		if(i < (int) m->d_tiles.size())
		{
			cudaFree(m->d_tiles[i]);
			cudaFree(m->d_gathered[i]);
		}
		if(i == 0)
		{
			cudaFree(m->d_frame);
		}
		if(i < (int) m->d_band.size())
		{
			cudaFree(m->d_band[i]);
		}
		if(i < (int) m->comm.size() && m->comm[i])
		{
			ncclCommDestroy(m->comm[i]);
		}
		skr_destroy(m->ctx[i]);
	}
	delete m;
}

int skr_mgpu_init(int n_gpus, skr_mgpu **out)
{
	if(!out)
	{
		return fail(nullptr, SKR_ERR_ARG, "skr_mgpu_init: out is null");
	}
	*out	 = nullptr;
	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
	{
		return fail(nullptr, SKR_ERR_NO_DEVICE, "skr_mgpu_init: no CUDA device; this library has no CPU path");
	}
	if(n_gpus <= 0)
	{
		n_gpus = ndev;
	}
	if(n_gpus > ndev)
	{
		return fail(nullptr, SKR_ERR_NO_DEVICE, "skr_mgpu_init: %d GPUs requested, %d visible", n_gpus, ndev);
	}
	skr_mgpu *m = new skr_mgpu();
	m->world	= n_gpus;
	for(int i = 0; i < n_gpus; i++)
	{
		skr_ctx *c = nullptr;
		if(skr_init(i, &c) != SKR_OK)
		{
			fail(nullptr, SKR_ERR_CUDA, "skr_mgpu_init: GPU %d: %s", i, skr_last_error(nullptr));
			skr_mgpu_destroy(m);
			return SKR_ERR_CUDA;
		}
		m->ctx.push_back(c);
	}
	m->d_tiles.assign(n_gpus, nullptr);
	m->d_gathered.assign(n_gpus, nullptr);
	m->cap_tiles.assign(n_gpus, 0);
	m->comm.assign(n_gpus, nullptr);
	{
		const char *nd = getenv("SKR_MGPU_NO_DIRECT");
		m->direct	   = !(nd && nd[0] == '1');
	}
	for(int i = 0; i < n_gpus; i++)
	{
		m->workers.push_back(new Worker());
	}
	for(int i = 0; i < n_gpus; i++)
	{
		m->workers[i]->th = std::thread(worker_main, m, i);
	}
	m->d_band.assign(n_gpus, nullptr);
	m->cap_band.assign(n_gpus, 0);
	{
		// row bands need every GPU to map every other GPU's memory
		const char *nb = getenv("SKR_MGPU_NO_BANDS");
		m->bands	   = m->direct && n_gpus > 1 && !(nb && nb[0] == '1');
		for(int i = 0; i < n_gpus && m->bands; i++)
		{
			cudaSetDevice(i);
			for(int j = 0; j < n_gpus && m->bands; j++)
			{
				if(i == j)
				{
					continue;
				}
				int can = 0;
				if(cudaDeviceCanAccessPeer(&can, i, j) != cudaSuccess || !can)
				{
					m->bands = false;
					break;
				}
				const cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
				if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				{
					m->bands = false;
				}
				cudaGetLastError();
			}
		}
		cudaSetDevice(0);
	}
	// peer path: every other GPU maps GPU 0's memory; its kernels then store finished pixels into GPU 0's frame
	{
		const char *no = getenv("SKR_MGPU_NO_P2P");
		m->p2p		   = !m->direct && n_gpus > 1 && !(no && no[0] == '1');
		for(int i = 1; i < n_gpus && m->p2p; i++)
		{
			int can = 0;
			cudaSetDevice(i);
			if(cudaDeviceCanAccessPeer(&can, i, 0) != cudaSuccess || !can)
			{
				m->p2p = false;
				break;
			}
			const cudaError_t e = cudaDeviceEnablePeerAccess(0, 0);
			if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
			{
				m->p2p = false;
			}
			cudaGetLastError();
		}
		cudaSetDevice(0);
	}
	if(n_gpus > 1 && !m->p2p && !m->direct)
	{
		std::vector<int> devs(n_gpus);
		for(int i = 0; i < n_gpus; i++)
		{
			devs[i] = i;
		}
		ncclResult_t r = ncclCommInitAll(m->comm.data(), n_gpus, devs.data());
		if(r != ncclSuccess)
		{
			fail(nullptr, SKR_ERR_CUDA, "skr_mgpu_init: ncclCommInitAll: %s", ncclGetErrorString(r));
			skr_mgpu_destroy(m);
			return SKR_ERR_CUDA;
		}
	}
	*out = m;
	return SKR_OK;
}

int skr_mgpu_scene_upload(skr_mgpu *m, const skr_scene_desc *scene)
{
	if(!m)
	{
		return fail(nullptr, SKR_ERR_ARG, "null handle");
	}
	std::vector<int> rc(m->world, 0);
	on_all(m, [&](int i) { rc[i] = skr_scene_upload(m->ctx[i], scene); });
	for(int i = 0; i < m->world; i++)
	{
		if(rc[i])
		{
			return fail(m, rc[i], "GPU %d: %s", i, skr_last_error(m->ctx[i]));
		}
	}
	return SKR_OK;
}

int skr_mgpu_render(skr_mgpu *m, const skr_options *opt, uint8_t *rgb8, skr_stats *stats)
{
	if(!m || !opt || !rgb8)
	{
		return fail(m, SKR_ERR_ARG, "skr_mgpu_render: null argument");
	}
	const int W = m->world;
	skr_options o0 = *opt;
	o0.world	   = W;
	o0.rank		   = 0;
	const int64_t tb = skr_tiles_bytes(&o0);
	if(tb <= 0)
	{
		return fail(m, SKR_ERR_ARG, "skr_mgpu_render: bad options");
	}
	const size_t frame_bytes = (size_t) opt->width * opt->height * 3;
	if(m->bands)
	{
		return render_bands(m, opt, rgb8, stats, frame_bytes);
	}
	if(m->direct)
	{
		return render_direct(m, opt, rgb8, stats, frame_bytes);
	}
	if(m->p2p)
	{
		return render_p2p(m, opt, rgb8, stats, frame_bytes);
	}
	for(int i = 0; i < W; i++)
	{
		cudaSetDevice(i);
		if(m->cap_tiles[i] < (size_t) tb)
		{
			cudaFree(m->d_tiles[i]);
			cudaFree(m->d_gathered[i]);
			m->d_tiles[i] = m->d_gathered[i] = nullptr;
			if(cudaMalloc(&m->d_tiles[i], (size_t) tb) != cudaSuccess || cudaMalloc(&m->d_gathered[i], (size_t) tb * W) != cudaSuccess)
			{
				m->cap_tiles[i] = 0;
				return fail(m, SKR_ERR_CUDA, "skr_mgpu_render: cudaMalloc failed on GPU %d", i);
			}
			m->cap_tiles[i] = (size_t) tb;
		}
	}
	cudaSetDevice(0);
	if(m->cap_frame < frame_bytes)
	{
		cudaFree(m->d_frame);
		m->d_frame = nullptr;
		if(cudaMalloc(&m->d_frame, frame_bytes) != cudaSuccess)
		{
			m->cap_frame = 0;
			return fail(m, SKR_ERR_CUDA, "skr_mgpu_render: cudaMalloc(frame) failed");
		}
		m->cap_frame = frame_bytes;
	}

	std::vector<int> rc(W, 0);
	std::vector<std::string> msg(W);
	std::vector<skr_stats> st(W);
	std::vector<float> ms_tail(W, 0.0f);
	std::vector<std::thread> th;
	for(int i = 0; i < W; i++)
	{
		th.emplace_back([&, i]() {
			cudaSetDevice(i);
			skr_options oi = *opt;
			oi.world	   = W;
			oi.rank		   = i;
			memset(&st[i], 0, sizeof st[i]);
			rc[i] = skr_render_tiles_device(m->ctx[i], &oi, m->d_tiles[i], &st[i]);
			if(rc[i])
			{
				msg[i] = skr_last_error(m->ctx[i]);
				// still take part in the collective so that the other ranks do not hang
			}
			cudaStream_t s = (cudaStream_t) skr_stream(m->ctx[i]);
			cudaEvent_t e0, e1;
			cudaEventCreate(&e0);
			cudaEventCreate(&e1);
			cudaEventRecord(e0, s);
			const uint8_t *src = m->d_tiles[i];
			if(W > 1)
			{
				ncclResult_t r = ncclAllGather(m->d_tiles[i], m->d_gathered[i], (size_t) tb, ncclUint8, m->comm[i], s);
				if(r != ncclSuccess && !rc[i])
				{
					rc[i]  = SKR_ERR_CUDA;
					msg[i] = std::string("ncclAllGather: ") + ncclGetErrorString(r);
				}
				src = m->d_gathered[i];
			}
			if(i == 0 && !rc[i])
			{
				rc[i] = skr_deinterleave_device(m->ctx[0], &oi, src, m->d_frame);
				if(rc[i])
				{
					msg[i] = skr_last_error(m->ctx[0]);
				}
				else if(cudaMemcpyAsync(rgb8, m->d_frame, frame_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess)
				{
					rc[i]  = SKR_ERR_CUDA;
					msg[i] = "cudaMemcpyAsync(frame) failed";
				}
			}
			cudaEventRecord(e1, s);
			if(cudaStreamSynchronize(s) != cudaSuccess && !rc[i])
			{
				rc[i]  = SKR_ERR_CUDA;
				msg[i] = "cudaStreamSynchronize failed";
			}
			cudaEventElapsedTime(&ms_tail[i], e0, e1);
			cudaEventDestroy(e0);
			cudaEventDestroy(e1);
		});
	}
	for(std::thread &t : th)
	{
		t.join();
	}
	for(int i = 0; i < W; i++)
	{
		if(rc[i])
		{
			return fail(m, rc[i], "GPU %d: %s", i, msg[i].c_str());
		}
	}
	if(stats)
	{
		sum_stats(stats, st);
		stats->ms_d2h = ms_tail[0];
	}
	return SKR_OK;
}

} // extern "C"
