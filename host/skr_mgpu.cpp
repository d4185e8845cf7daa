// host/skr_mgpu.cpp -- implementation of include/skr_mgpu.h (libskr_mgpu.so): one process, one host thread and one
// skr_ctx per GPU, NCCL all-gather of the finished RGB8 tiles.  See the header for the design.
#include "../include/skr_mgpu.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace
{
thread_local std::string g_err;
}

struct skr_mgpu
{
	int world = 0;
	std::vector<skr_ctx *> ctx;
	std::vector<ncclComm_t> comm;
	std::vector<uint8_t *> d_tiles;	   // per GPU: its compact tiles
	std::vector<uint8_t *> d_gathered; // per GPU: all ranks' tiles (all-gather result)
	std::vector<size_t> cap_tiles;
	uint8_t *d_frame = nullptr; // GPU 0
	size_t cap_frame = 0;
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
	if(n_gpus > 1)
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
	std::vector<std::thread> th;
	for(int i = 0; i < m->world; i++)
	{
		th.emplace_back([&, i]() { rc[i] = skr_scene_upload(m->ctx[i], scene); });
	}
	for(std::thread &t : th)
	{
		t.join();
	}
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
		memset(stats, 0, sizeof *stats);
		for(int i = 0; i < W; i++)
		{
			stats->closest_hit_rays += st[i].closest_hit_rays;
			stats->shadow_rays += st[i].shadow_rays;
			stats->sphere_tests += st[i].sphere_tests;
			stats->sphere_tests_pos += st[i].sphere_tests_pos;
			stats->tri_tests += st[i].tri_tests;
			stats->bvh_node_visits += st[i].bvh_node_visits;
			stats->sphere_hits += st[i].sphere_hits;
			stats->light_evals += st[i].light_evals;
			stats->sphere_tests_executed += st[i].sphere_tests_executed;
			stats->queue_entries += st[i].queue_entries;
			stats->kernel_launches += st[i].kernel_launches;
			stats->queue_chunks += st[i].queue_chunks;
			if(st[i].ms_total > stats->ms_total)
			{
				stats->ms_total	  = st[i].ms_total;
				stats->ms_primary = st[i].ms_primary;
				stats->ms_bounce  = st[i].ms_bounce;
				stats->ms_resolve = st[i].ms_resolve;
			}
		}
		stats->ms_d2h = ms_tail[0];
	}
	return SKR_OK;
}

} // extern "C"
