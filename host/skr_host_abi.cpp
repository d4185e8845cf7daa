// host/skr_host_abi.cpp -- C entry points over the host-side .scn reader, so that the parity tests (Python,
// ctypes) can compare it with the reference's parser field by field.  Built into libskr_host.so.
#include <cstring>
#include <string>

#include "scene_parser.h"

extern "C" {

struct skr_host_scene
{
	skr_host::HostScene scene;
	std::string error;
};

// flags: bit0 keep directional lights, bit1 drop spherical_fog lines
skr_host_scene *skr_host_parse_scn(const char *path, int flags)
{
	skr_host_scene *h = new skr_host_scene();
	skr_host::ParseOptions o;
	o.keep_directional = (flags & 1) != 0;
	o.fog			   = (flags & 2) == 0;
	if(!skr_host::parse_scn(path ? path : "", h->scene, h->error, o))
	{
		if(h->error.empty())
		{
			h->error = "parse failed";
		}
	}
	return h;
}

const char *skr_host_scene_error(const skr_host_scene *h)
{
	return h->error.c_str();
}

// counts[0..4] = nspheres ntris nplights ndlights nfogs ; [5],[6] = film_resolution ; [7] = max_depth ; [8] = unknown commands
void skr_host_scene_counts(const skr_host_scene *h, int *counts)
{
	const skr_host::HostScene &s = h->scene;
	counts[0] = s.nspheres(), counts[1] = s.ntris(), counts[2] = s.nplights(), counts[3] = s.ndlights(), counts[4] = s.nfogs();
	counts[5] = s.film_width, counts[6] = s.film_height, counts[7] = s.max_depth, counts[8] = s.unknown_commands;
}

void skr_host_scene_desc(const skr_host_scene *h, skr_scene_desc *out)
{
	*out = h->scene.desc();
}

void skr_host_scene_free(skr_host_scene *h)
{
	delete h;
}

// the reader's number conversion, exposed so that the tests can compare it with strtof bit for bit
int skr_host_read_floats(const char *text, float *out, int max);

int skr_host_write_ppm(const char *path, int width, int height, const unsigned char *rgb8)
{
	std::string err;
	return skr_host::write_ppm(path, width, height, rgb8, err) ? 0 : 1;
}
}
