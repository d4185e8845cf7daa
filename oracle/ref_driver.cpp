// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Thin C-ABI driver around the UNMODIFIED reference tracer.  Everything from
// shade() downward (src/raytrace.h, src/blinn_phong.h, src/utils.h, the data
// model headers, glm, and the .scn parser src/scene.cpp) is compiled from the
// sources where they lie under /root/reference/src (see oracle/Makefile, which
// passes -I$(REF)/src and compiles $(REF)/src/scene.cpp in place).  Nothing of
// the reference is copied into this repository; only the built shared object
// lands in oracle/_ref/ (git-ignored).
//
// The only reference code that is *restated* here is the per-pixel frame loop
// generate_rays_parallel (src/main.cpp:19-104), because at HEAD that function
// hard-overrides width/height/depth/grid (src/main.cpp:21-24) and
// main() leaves use_shadows uninitialised (src/main.cpp:244) and
// seeds rand() from time(0) (src/main.cpp:400).  The restated loop
// is the same arithmetic, with those three things turned into arguments.
// tests/test_oracle_ref.py pins it against (a) renders/testcpu.ppm and (b) the
// unmodified main.cpp built with an SDL stub (oracle/_ref/raytracer_unmodified).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load this library.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <fcntl.h>
#include <omp.h>
#include <chrono>

#include "raytrace.h" // reference: src/raytrace.h (pulls in blinn_phong.h, utils.h, scene.h ...)

namespace
{
struct QuietStdout
{
	int saved;
	QuietStdout()
	{
		fflush(stdout);
		saved  = dup(1);
		int dn = open("/dev/null", O_WRONLY);
		dup2(dn, 1);
		close(dn);
	}
	~QuietStdout()
	{
		fflush(stdout);
		dup2(saved, 1);
		close(saved);
	}
};

// Flat scene snapshot layout shared with the port and the CUDA library
// (documented in include/skr.h):
//   spheres      : nspheres  x 18 floats  [cx cy cz r | amb3 | diff3 | spec3 | trans3 | power ior]
//   triangles    : ntris     x  9 floats  [v0 | v1 | v2]
//   point_lights : nplights  x  6 floats  [pos3 | colour3]
//   dir_lights   : ndlights  x  6 floats  [dir3 | colour3]
//   fogs         : nfogs     x  9 floats  [scattering absorption | albedo3 | radius | centre3]
//   camera       : 12 floats [position | direction | up | right]
//   ambient      : 3 floats, background : 3 floats
} // namespace

extern "C" {

struct ref_scene; // opaque: a heap-allocated reference `Scene`

// Parse a .scn with the reference parser (src/scene.cpp:12-227).  The parser
// prints every line and drops "simplesphere.txt" into the CWD
// (src/scene.cpp:96-102); both are neutralised here (stdout muted, CWD moved
// to /tmp for the duration).
ref_scene *ref_parse_scene(const char *path)
{
	char abspath[4096];
	if(!realpath(path, abspath))
	{
		return nullptr;
	}
	char cwd[4096];
	if(!getcwd(cwd, sizeof cwd))
	{
		return nullptr;
	}
	Scene *s = new Scene();
	{
		QuietStdout q;
		if(chdir("/tmp") != 0) {}
		*s = parseScene(std::string(abspath));
		if(chdir(cwd) != 0) {}
	}
	return reinterpret_cast<ref_scene *>(s);
}

void ref_scene_free(ref_scene *h)
{
	delete reinterpret_cast<Scene *>(h);
}

// counts[0..4] = nspheres, ntris, nplights, ndlights, nfogs ; counts[5],[6] = film width,height
void ref_scene_counts(const ref_scene *h, int *counts)
{
	const Scene *s = reinterpret_cast<const Scene *>(h);
	counts[0]	   = (int) s->spheres.size();
	counts[1]	   = (int) s->triangles.size();
	counts[2]	   = (int) s->point_lights.size();
	counts[3]	   = (int) s->directional_lights.size();
	counts[4]	   = (int) s->spherical_fog.size();
	counts[5]	   = s->width;
	counts[6]	   = s->height;
}

static void put3(float *dst, const glm::vec3 &v)
{
	dst[0] = v.x;
	dst[1] = v.y;
	dst[2] = v.z;
}
static glm::vec3 get3(const float *p)
{
	return glm::vec3(p[0], p[1], p[2]);
}

void ref_scene_export(const ref_scene *h, float *spheres, float *tris, float *plights, float *dlights, float *fogs, float *camera, float *ambient, float *background)
{
	const Scene *s = reinterpret_cast<const Scene *>(h);
	for(size_t i = 0; i < s->spheres.size(); i++)
	{
		float *d		 = spheres + 18 * i;
		const Sphere &sp = s->spheres[i];
		put3(d, sp.collider.position);
		d[3] = sp.collider.radius;
		put3(d + 4, sp.material.ambient);
		put3(d + 7, sp.material.diffuse);
		put3(d + 10, sp.material.specular);
		put3(d + 13, sp.material.transmissive);
		d[16] = sp.material.power;
		d[17] = sp.material.ior;
	}
	for(size_t i = 0; i < s->triangles.size(); i++)
	{
		float *d = tris + 9 * i;
		put3(d, s->triangles[i].v0);
		put3(d + 3, s->triangles[i].v1);
		put3(d + 6, s->triangles[i].v2);
	}
	for(size_t i = 0; i < s->point_lights.size(); i++)
	{
		put3(plights + 6 * i, s->point_lights[i].position);
		put3(plights + 6 * i + 3, s->point_lights[i].colour);
	}
	for(size_t i = 0; i < s->directional_lights.size(); i++)
	{
		put3(dlights + 6 * i, s->directional_lights[i].direction);
		put3(dlights + 6 * i + 3, s->directional_lights[i].colour);
	}
	for(size_t i = 0; i < s->spherical_fog.size(); i++)
	{
		float *d			  = fogs + 9 * i;
		const SphericalFog &f = s->spherical_fog[i];
		d[0]				  = f.scattering;
		d[1]				  = f.absorption;
		put3(d + 2, f.albedo);
		d[5] = f.collider.radius;
		put3(d + 6, f.collider.position);
	}
	put3(camera, s->camera.position);
	put3(camera + 3, s->camera.direction);
	put3(camera + 6, s->camera.up);
	put3(camera + 9, s->camera.right);
	put3(ambient, s->ambient_light.colour);
	put3(background, s->background);
}

// Build a reference `Scene` from a flat snapshot (so that tests can feed the
// reference code the very same numbers the CUDA library gets, fog fields
// included -- SURVEY F5).
ref_scene *ref_scene_from_arrays(int nspheres, const float *spheres, int ntris, const float *tris, int nplights, const float *plights, int ndlights, const float *dlights, int nfogs, const float *fogs, const float *camera, const float *ambient, const float *background)
{
	Scene *s = new Scene();
	for(int i = 0; i < nspheres; i++)
	{
		const float *d = spheres + 18 * i;
		Sphere sp;
		sp.collider.position	 = get3(d);
		sp.collider.radius		 = d[3];
		sp.material.ambient		 = get3(d + 4);
		sp.material.diffuse		 = get3(d + 7);
		sp.material.specular	 = get3(d + 10);
		sp.material.transmissive = get3(d + 13);
		sp.material.power		 = d[16];
		sp.material.ior			 = d[17];
		s->spheres.push_back(sp);
	}
	for(int i = 0; i < ntris; i++)
	{
		const float *d = tris + 9 * i;
		Triangle t;
		t.v0 = get3(d);
		t.v1 = get3(d + 3);
		t.v2 = get3(d + 6);
		s->triangles.push_back(t);
	}
	for(int i = 0; i < nplights; i++)
	{
		PointLight l;
		l.position = get3(plights + 6 * i);
		l.colour   = get3(plights + 6 * i + 3);
		s->point_lights.push_back(l);
	}
	for(int i = 0; i < ndlights; i++)
	{
		DirectionalLight l;
		l.direction = get3(dlights + 6 * i);
		l.colour	= get3(dlights + 6 * i + 3);
		s->directional_lights.push_back(l);
	}
	for(int i = 0; i < nfogs; i++)
	{
		const float *d = fogs + 9 * i;
		s->spherical_fog.push_back(SphericalFog(d[0], d[1], get3(d + 2), d[5], get3(d + 6)));
	}
	s->camera.position	= get3(camera);
	s->camera.direction = get3(camera + 3);
	s->camera.up		= get3(camera + 6);
	s->camera.right		= get3(camera + 9);
	s->ambient_light.colour = get3(ambient);
	s->background			= get3(background);
	return reinterpret_cast<ref_scene *>(s);
}

// The restated frame loop: src/main.cpp:33-86 (pixel loop) and :88-100
// (quantiser), without the overrides at :21-24.
//   rgb32 : optional H*W*3 float image (pre-clamp, post-average)
//   rgb8  : optional H*W*3 bytes, (unsigned char)(min(1,c)*255)
//   y0,y1 : row window [y0,y1) actually rendered (bounded CPU samples for the
//           bench; rows outside are left untouched).  Pass 0,height for a frame.
//   threads<=1 : serial (reproducible rand() stream after srand(seed)).
// Returns the wall time of the pixel loop in seconds.
double ref_render(const ref_scene *h, int width, int height, float fov, int max_depth, int monte_carlo, int num_path_traces, int grid_size, int use_shadows, unsigned seed, int threads, int y0, int y1, float *rgb32, unsigned char *rgb8)
{
	Scene scene		  = *reinterpret_cast<const Scene *>(h);
	scene.width		  = width;
	scene.height	  = height;
	scene.use_shadows = use_shadows != 0;
	Options option;
	option.monte_carlo	   = monte_carlo != 0;
	option.num_path_traces = (short) num_path_traces;
	option.fov			   = fov;
	option.grid_size	   = (short) grid_size;
	option.max_depth	   = max_depth;

	srand(seed);
	if(threads < 1)
	{
		threads = 1;
	}
	if(y0 < 0)
	{
		y0 = 0;
	}
	if(y1 > height)
	{
		y1 = height;
	}

	glm::vec3 *image = new glm::vec3[(size_t) width * (size_t) height];

	auto t_begin = std::chrono::steady_clock::now();
	// clang-format off
	#pragma omp parallel for num_threads(threads) if(threads > 1)
	// clang-format on
	for(int y = y0; y < y1; y++)
	{
		for(int x = 0; x < scene.width; x++)
		{
			float inv_width	   = 1 / float(scene.width);
			float inv_height   = 1 / float(scene.height);
			float aspect_ratio = scene.width / float(scene.height);
			float angle		   = tan(M_PI * 0.5 * option.fov / 180.);
			glm::vec3 &px	   = image[(size_t) y * width + x];

			if(option.grid_size > 0)
			{
				for(int i = 0; i < option.grid_size; i++)
				{
					for(int j = 0; j < option.grid_size; j++)
					{
						float r = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
						float u = (2 * ((x + r) * inv_width) - 1) * angle * aspect_ratio;
						float v = (1 - 2 * ((y + r) * inv_height)) * angle;
						glm::vec3 ray_dir(scene.camera.direction + u * scene.camera.right + v * scene.camera.up);
						Ray ray;
						ray.position  = scene.camera.position;
						ray.direction = ray_dir;
						px += shade(ray, scene, option.max_depth, option.monte_carlo, option.num_path_traces);
					}
				}
				px /= (option.grid_size * option.grid_size);
			}
			else
			{
				float u = (2 * ((x + 0.5) * inv_width) - 1) * angle * aspect_ratio;
				float v = (1 - 2 * ((y + 0.5) * inv_height)) * angle;
				glm::vec3 ray_dir(scene.camera.direction + u * scene.camera.right + v * scene.camera.up);
				Ray ray;
				ray.position  = scene.camera.position;
				ray.direction = ray_dir;
				px			  = shade(ray, scene, option.max_depth, option.monte_carlo, option.num_path_traces);
			}
		}
	}
	auto t_end = std::chrono::steady_clock::now();

	for(int y = y0; y < y1; y++)
	{
		for(int x = 0; x < width; x++)
		{
			size_t i		   = (size_t) y * width + x;
			const glm::vec3 &c = image[i];
			if(rgb32)
			{
				rgb32[3 * i + 0] = c.x;
				rgb32[3 * i + 1] = c.y;
				rgb32[3 * i + 2] = c.z;
			}
			if(rgb8)
			{
				rgb8[3 * i + 0] = (unsigned char) (std::min(float(1), c.x) * 255);
				rgb8[3 * i + 1] = (unsigned char) (std::min(float(1), c.y) * 255);
				rgb8[3 * i + 2] = (unsigned char) (std::min(float(1), c.z) * 255);
			}
		}
	}
	delete[] image;
	return std::chrono::duration<double>(t_end - t_begin).count();
}

// ---- single-function probes for known-answer tests of the port -------------

float ref_smallest_root(float a, float b, float c)
{
	return smallest_root(a, b, c);
}

// ray = o[3], d[3]; sphere = c[3], r.  Returns collision_distance; *occurs = intersection_occurs.
float ref_sphere_hit(const float *o, const float *d, const float *c, float r, int *occurs)
{
	Ray ray;
	ray.position  = get3(o);
	ray.direction = get3(d);
	SphereCollider col;
	col.position = get3(c);
	col.radius	 = r;
	*occurs		 = intersection_occurs(ray, col) ? 1 : 0;
	return collision_distance(ray, col);
}

int ref_triangle_hit(const float *o, const float *d, const float *tri9, float *tuv)
{
	Ray ray;
	ray.position  = get3(o);
	ray.direction = get3(d);
	Triangle t;
	t.v0	 = get3(tri9);
	t.v1	 = get3(tri9 + 3);
	t.v2	 = get3(tri9 + 6);
	float tt = 0, u = 0, v = 0;
	bool hit = triangle_intersection_occurs(ray, t, tt, u, v);
	tuv[0]	 = tt;
	tuv[1]	 = u;
	tuv[2]	 = v;
	return hit ? 1 : 0;
}

void ref_transform_coordinate_space(const float *n, float *nt, float *nb)
{
	std::tuple<glm::vec3, glm::vec3> t = transform_coordinate_space(get3(n));
	put3(nt, std::get<0>(t));
	put3(nb, std::get<1>(t));
}

void ref_uniform_sample_hemi(float r1, float r2, float *out)
{
	put3(out, uniform_sample_hemi(r1, r2));
}

// One shade() call (src/raytrace.h:139) on an arbitrary ray.
void ref_shade(const ref_scene *h, int use_shadows, const float *o, const float *d, int depth, int monte_carlo, int num_path_traces, unsigned seed, float *rgb)
{
	Scene scene		  = *reinterpret_cast<const Scene *>(h);
	scene.use_shadows = use_shadows != 0;
	Ray ray;
	ray.position  = get3(o);
	ray.direction = get3(d);
	srand(seed);
	put3(rgb, shade(ray, scene, depth, monte_carlo != 0, (short) num_path_traces));
}

int ref_shadow_point(const ref_scene *h, const float *p, const float *light6)
{
	const Scene &scene = *reinterpret_cast<const Scene *>(h);
	PointLight l;
	l.position = get3(light6);
	l.colour   = get3(light6 + 3);
	return shadow(scene, get3(p), l) ? 1 : 0;
}

float ref_fresnel(const float *dir, const float *n, float ior)
{
	Sphere s;
	s.material.ior = ior;
	return bp::fresnel(get3(dir), get3(n), s);
}

void ref_refraction(const float *dir, const float *n, float ior, float *out)
{
	Sphere s;
	s.material.ior = ior;
	put3(out, bp::refraction(get3(dir), get3(n), s));
}

void ref_reflect_direction(const float *l, const float *n, float *out)
{
	put3(out, bp::reflect_direction(get3(l), get3(n)));
}

int ref_max_threads(void)
{
	return omp_get_max_threads();
}

} // extern "C"
