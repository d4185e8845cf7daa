// host/scene_parser.h -- .scn scene description reader for the B200 renderer's host side.
//
// Same grammar and the same "current material" state machine as the reference parser
// (reference: src/scene.cpp:12-227), producing the flat arrays that cross the C ABI
// (include/skr.h: skr_scene_desc) instead of the reference's `Scene` of std::vectors.
#ifndef SKR_HOST_SCENE_PARSER_H
#define SKR_HOST_SCENE_PARSER_H

#include <string>
#include <vector>

#include "../include/skr.h"

namespace skr_host
{
struct HostScene
{
	std::vector<float> spheres;	 // 18 per sphere
	std::vector<float> vertices; // 3 per vertex (parse-time pool, like Scene::vertices)
	std::vector<float> tris;	 // 9 per triangle
	std::vector<float> tri_materials; // 14 per triangle: the current material at the `triangle` line (src/scene.cpp:80);
									  // the reference never reads it, the opt-in shaded-triangles mode does
	std::vector<float> plights;	 // 6 per light
	std::vector<float> dlights;	 // 6 per light (stays empty unless keep_directional, see below)
	std::vector<float> fogs;	 // 9 per fog
	float camera[12] = {0};		 // position | direction | up | right
	float ambient[3] = {0, 0, 0};
	float background[3] = {0, 0, 0};
	// parsed and then ignored by rendering, exactly like the reference
	int film_width = 1920, film_height = 1080;
	int max_depth = 1;
	std::string output_image;
	int unknown_commands = 0;

	int nspheres() const { return (int) (spheres.size() / 18); }
	int ntris() const { return (int) (tris.size() / 9); }
	int nplights() const { return (int) (plights.size() / 6); }
	int ndlights() const { return (int) (dlights.size() / 6); }
	int nfogs() const { return (int) (fogs.size() / 9); }
	skr_scene_desc desc() const; // pointers into this object
};

struct ParseOptions
{
	bool verbose = false; // echo each recognised command like the reference does (src/scene.cpp:33-216)
	// The reference parses `directional_light` lines and then drops them (no push_back, src/scene.cpp:139-163).
	// Default false = same behaviour.
	bool keep_directional = false;
	// `spherical_fog` in the reference is read with the format "fog %f ..." against a line that starts with
	// "spherical_fog", converts nothing and pushes a fog made of uninitialised stack floats
	// (src/scene.cpp:207-212).  That cannot be reproduced; this reader parses the nine intended fields
	// x y z radius r g b scattering absorption (missing trailing fields = 0).  fog = false drops the line.
	bool fog = true;
};

// Returns false and sets `error` on failure (unreadable file, triangle index out of range).
bool parse_scn(const std::string &path, HostScene &out, std::string &error, const ParseOptions &opt = ParseOptions());

// "P6\n<w> <h>\n255\n" + row-major RGB bytes (reference: src/main.cpp:88-100)
bool write_ppm(const std::string &path, int width, int height, const unsigned char *rgb8, std::string &error);
} // namespace skr_host

#endif
