// host/skr_main.cpp -- command-line front end of the B200 renderer ("raytracer").
//
// Keeps the reference program's surface (reference: src/main.cpp:230-413, README.md:24-33):
//   --path <scn> --output <ppm> [--width n] [--height n] [--fov deg] [--gillum n] [--jsample n] [--depth n]
//   [--parallel true|false] [--shadow]
// same defaults (1920x1080, fov 60, depth 3, src/utils.h:28-33 and src/scene.h:15), same precedence (command line
// over the .scn's film_resolution / max_depth, which are parsed and ignored), same PPM bytes.  The frame function
// generate_rays_parallel (src/main.cpp:19-104) is replaced by calls into libskr.so (include/skr.h).
// Differences, all additive or bug-for-intent:
//   * --parallel is accepted and ignored: the GPU path is always "parallel", there is no SDL preview;
//   * use_shadows starts false (the reference leaves it uninitialised, src/main.cpp:244);
//   * the 640x480 / depth 1 / no-jsample overrides of src/main.cpp:21-24 are not applied;
//   * new flags: --seed n (replaces srand(time(0))), --gpu i, --gpus N (frame split over N GPUs, NCCL gather; 0 = all visible),
//     --stats, --fresnel, --shade-triangles (beyond the reference: triangles lit with their own materials), --keep-directional, --verbose, --no-fog;
//   * a failing CUDA/library call prints the message and exits 1 (the reference never exits nonzero).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../include/skr.h"
#ifdef SKR_WITH_NCCL
#include "../include/skr_mgpu.h"
#endif
#include "scene_parser.h"

int main(int argc, char *argv[])
{
	skr_options opt;
	memset(&opt, 0, sizeof opt);
	opt.width			= 1920; // Scene::width/height defaults, src/scene.h:15
	opt.height			= 1080;
	opt.fov				= 60;	// Options defaults, src/utils.h:28-33
	opt.num_path_traces = 1;
	opt.grid_size		= 0;
	opt.max_depth		= 3;
	opt.seed			= 0;

	const char *path   = nullptr;
	const char *output = nullptr;
	int gpu			   = -1;
	int gpus		   = 1;
	bool want_stats	   = false;
	skr_host::ParseOptions popt;

	for(int i = 0; i < argc; i++)
	{
		auto has_arg = [&]() { return i + 1 < argc; };
		if(strcmp(argv[i], "--gillum") == 0)
		{
			if(has_arg())
			{
				opt.monte_carlo		= 1;
				opt.num_path_traces = atoi(argv[i + 1]);
			}
			else
			{
				std::cerr << "gillum takes an int after flag for the number of paths traced" << std::endl; // continues, like the reference
			}
		}
		if(strcmp(argv[i], "--fov") == 0)
		{
			if(!has_arg())
			{
				std::cerr << "fov takes a float (degrees) after flag for the field of view" << std::endl;
				return 0;
			}
			opt.fov = (float) atof(argv[i + 1]);
		}
		if(strcmp(argv[i], "--jsample") == 0)
		{
			if(!has_arg())
			{
				std::cerr << "jsample takes an int after flag for the supersampling grid size" << std::endl;
				return 0;
			}
			opt.grid_size = atoi(argv[i + 1]);
		}
		if(strcmp(argv[i], "--width") == 0)
		{
			if(!has_arg())
			{
				std::cerr << "width takes an int after flag for the width" << std::endl;
				return 0;
			}
			opt.width = atoi(argv[i + 1]);
		}
		if(strcmp(argv[i], "--height") == 0)
		{
			if(!has_arg())
			{
				std::cerr << "height takes an int after flag for the width" << std::endl;
				return 0;
			}
			opt.height = atoi(argv[i + 1]);
		}
		if(strcmp(argv[i], "--depth") == 0)
		{
			if(!(has_arg() && atoi(argv[i + 1]) > 0))
			{
				std::cerr << "depth takes a positive int after flag for the max depth" << std::endl;
				return 0;
			}
			opt.max_depth = atoi(argv[i + 1]);
		}
		if(strcmp(argv[i], "--path") == 0)
		{
			if(!has_arg())
			{
				std::cerr << "path must be passed after --path" << std::endl;
				return 0;
			}
			path = argv[i + 1];
		}
		if(strcmp(argv[i], "--output") == 0)
		{
			if(!has_arg())
			{
				std::cerr << "output path must be passed after --output" << std::endl;
				return 0;
			}
			output = argv[i + 1];
		}
		if(strcmp(argv[i], "--shadow") == 0)
		{
			opt.use_shadows = 1;
		}
		// --parallel true|false: accepted, no effect
		// additive flags
		if(strcmp(argv[i], "--seed") == 0 && has_arg())
		{
			opt.seed = strtoull(argv[i + 1], nullptr, 10);
		}
		if(strcmp(argv[i], "--gpu") == 0 && has_arg())
		{
			gpu = atoi(argv[i + 1]);
		}
		if(strcmp(argv[i], "--gpus") == 0 && has_arg())
		{
			gpus = atoi(argv[i + 1]);
		}
		if(strcmp(argv[i], "--stats") == 0)
		{
			want_stats = true;
		}
		if(strcmp(argv[i], "--fresnel") == 0)
		{
			opt.fresnel = 1;
		}
		if(strcmp(argv[i], "--shade-triangles") == 0) // non-parity extension: triangles lit with their own materials
		{
			opt.shade_triangles = 1;
		}
		if(strcmp(argv[i], "--verbose") == 0)
		{
			popt.verbose = true;
		}
		if(strcmp(argv[i], "--no-fog") == 0)
		{
			popt.fog = false;
		}
		if(strcmp(argv[i], "--keep-directional") == 0) // store the directional_light lines the reference parser drops (src/scene.cpp:139-163)
		{
			popt.keep_directional = true;
		}
	}
	if(!path)
	{
		std::cerr << "no scene file was passed. Pass with --path path_to_scn" << std::endl;
		return 0;
	}
	if(!output)
	{
		std::cerr << "no output destination was passed. Pass with --output destination_path.ppm" << std::endl;
		return 0;
	}

	skr_host::HostScene scene;
	std::string err;
	auto t0 = std::chrono::steady_clock::now();
	if(!skr_host::parse_scn(path, scene, err, popt))
	{
		printf("%s\n", err.c_str());
		return 0; // the reference exit(0)s on an unreadable scene, src/scene.cpp:22-26
	}
	auto t1 = std::chrono::steady_clock::now();
	printf("\n\nMonte carlo: %d\nvisual display: %d\nfov: %f\nnum paths traced: %d\nsupersample grid size: %d\nmax depth: %d\n", opt.monte_carlo, 0, opt.fov,
		   opt.num_path_traces, opt.grid_size, opt.max_depth); // Options::to_string, src/utils.h:35-38

	if(gpus != 1)
	{
#ifdef SKR_WITH_NCCL
		skr_mgpu *mg = nullptr;
		if(skr_mgpu_init(gpus, &mg) != SKR_OK)
		{
			std::cerr << "skr_mgpu_init failed: " << skr_mgpu_last_error(nullptr) << std::endl;
			return 1;
		}
		const skr_scene_desc mdesc = scene.desc();
		std::vector<unsigned char> frame((size_t) opt.width * opt.height * 3);
		opt.collect_stats = want_stats ? 1 : 0;
		skr_stats mst;
		auto m0 = std::chrono::steady_clock::now();
		if(skr_mgpu_scene_upload(mg, &mdesc) != SKR_OK || skr_mgpu_render(mg, &opt, frame.data(), &mst) != SKR_OK)
		{
			std::cerr << "multi-GPU render failed: " << skr_mgpu_last_error(mg) << std::endl;
			skr_mgpu_destroy(mg);
			return 1;
		}
		auto m1 = std::chrono::steady_clock::now();
		if(!skr_host::write_ppm(output, opt.width, opt.height, frame.data(), err))
		{
			std::cerr << err << std::endl;
			skr_mgpu_destroy(mg);
			return 1;
		}
		printf("***\nWROTE TO PPM\n***\n");
		if(want_stats)
		{
			printf("{\"gpus\": %d, \"upload_plus_render_wall_ms\": %.3f, \"device_ms_slowest_gpu\": %.3f, \"gather_deinterleave_d2h_ms\": %.3f, "
				   "\"closest_hit_rays\": %llu, \"shadow_rays\": %llu, \"kernel_launches\": %u}\n",
				   skr_mgpu_world(mg), std::chrono::duration<double, std::milli>(m1 - m0).count(), mst.ms_total, mst.ms_d2h,
				   (unsigned long long) mst.closest_hit_rays, (unsigned long long) mst.shadow_rays, mst.kernel_launches);
		}
		skr_mgpu_destroy(mg);
		return 0;
#else
		std::cerr << "--gpus needs the NCCL build (libskr_mgpu.so); rebuild where nccl.h is installed" << std::endl;
		return 1;
#endif
	}

	skr_ctx *ctx = nullptr;
	if(skr_init(gpu, &ctx) != SKR_OK)
	{
		std::cerr << "skr_init failed: " << skr_last_error(nullptr) << std::endl;
		return 1;
	}
	const skr_scene_desc desc = scene.desc();
	if(skr_scene_upload(ctx, &desc) != SKR_OK)
	{
		std::cerr << "skr_scene_upload failed: " << skr_last_error(ctx) << std::endl;
		skr_destroy(ctx);
		return 1;
	}
	auto t2 = std::chrono::steady_clock::now();
	std::vector<unsigned char> rgb8((size_t) opt.width * opt.height * 3);
	opt.collect_stats = want_stats ? 1 : 0;
	skr_stats st;
	if(skr_render(ctx, &opt, rgb8.data(), nullptr, &st) != SKR_OK)
	{
		std::cerr << "skr_render failed: " << skr_last_error(ctx) << std::endl;
		skr_destroy(ctx);
		return 1;
	}
	auto t3 = std::chrono::steady_clock::now();
	if(!skr_host::write_ppm(output, opt.width, opt.height, rgb8.data(), err))
	{
		std::cerr << err << std::endl;
		skr_destroy(ctx);
		return 1;
	}
	printf("***\nWROTE TO PPM\n***\n"); // src/main.cpp:101
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
		return std::chrono::duration<double, std::milli>(b - a).count();
	};
	if(want_stats)
	{
		const double rays = (double) st.closest_hit_rays + (double) st.shadow_rays;
		printf("{\"scene\": {\"spheres\": %d, \"triangles\": %d, \"point_lights\": %d, \"fogs\": %d}, \"parse_ms\": %.3f, \"upload_ms\": %.3f, "
			   "\"render_wall_ms\": %.3f, \"device_ms\": %.3f, \"primary_ms\": %.3f, \"bounce_ms\": %.3f, \"resolve_ms\": %.3f, \"d2h_ms\": %.3f, "
			   "\"closest_hit_rays\": %llu, \"shadow_rays\": %llu, \"sphere_tests\": %llu, \"sphere_tests_executed\": %llu, \"tri_tests\": %llu, \"bvh_node_visits\": %llu, "
			   "\"kernel_launches\": %u, \"mrays_per_s\": %.1f}\n",
			   scene.nspheres(), scene.ntris(), scene.nplights(), scene.nfogs(), ms(t0, t1), ms(t1, t2), ms(t2, t3), st.ms_total, st.ms_primary, st.ms_bounce,
			   st.ms_resolve, st.ms_d2h, (unsigned long long) st.closest_hit_rays, (unsigned long long) st.shadow_rays,
			   (unsigned long long) st.sphere_tests, (unsigned long long) st.sphere_tests_executed, (unsigned long long) st.tri_tests, (unsigned long long) st.bvh_node_visits, st.kernel_launches,
			   st.ms_total > 0 ? rays / (st.ms_total * 1e-3) / 1e6 : 0.0);
	}
	skr_destroy(ctx);
	return 0;
}
