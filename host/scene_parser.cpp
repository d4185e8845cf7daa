// host/scene_parser.cpp -- see scene_parser.h.  Grammar of the reference's parseScene (src/scene.cpp:12-227).
#include "scene_parser.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace skr_host
{
namespace
{
struct MaterialState // reference: src/material.h:9-26 defaults
{
	float ambient[3]	  = {0, 0, 0};
	float diffuse[3]	  = {0, 0, 0};
	float specular[3]	  = {0, 0, 0};
	float transmissive[3] = {0, 0, 0};
	float power			  = 1.0f;
	float ior			  = 1.0f;
};

// strtof-equivalent fast path for the numbers scene files actually contain ("-12.5", ".043", "1e-3"): up to 15
// significant digits and a small decimal exponent are converted exactly in double (mantissa and 10^k are both exact
// doubles, one correctly rounded multiply or divide) and then rounded to float.  The only way that double rounding
// can differ from strtof's single rounding is a double that sits within one ulp of a float rounding midpoint; that
// case -- and anything unusual (hex, inf/nan, long mantissas, big exponents) -- goes to strtof itself.
// tests/test_host_parser.py checks bit equality with strtof on random decimal strings and on the reference scenes.
inline bool fast_strtof(const char *p, const char **end, float *out)
{
	static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
	const char *q = p;
	while(*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r' || *q == '\v' || *q == '\f')
	{
		q++;
	}
	bool negative = false;
	if(*q == '-' || *q == '+')
	{
		negative = *q == '-';
		q++;
	}
	unsigned long long mant = 0;
	int digits = 0, frac = 0;
	bool any = false;
	while(*q >= '0' && *q <= '9')
	{
		any = true;
		if(mant || *q != '0')
		{
			if(++digits > 15)
			{
				return false;
			}
			mant = mant * 10 + (unsigned) (*q - '0');
		}
		q++;
	}
	if(*q == '.')
	{
		q++;
		while(*q >= '0' && *q <= '9')
		{
			any = true;
			if(mant || *q != '0')
			{
				if(++digits > 15)
				{
					return false;
				}
				mant = mant * 10 + (unsigned) (*q - '0');
			}
			frac++;
			q++;
		}
	}
	if(!any)
	{
		return false; // not a plain decimal (or not a number at all): let strtof decide
	}
	int e10 = 0;
	if(*q == 'e' || *q == 'E')
	{
		const char *r = q + 1;
		bool eneg	  = false;
		if(*r == '-' || *r == '+')
		{
			eneg = *r == '-';
			r++;
		}
		if(*r >= '0' && *r <= '9')
		{
			int ev = 0;
			while(*r >= '0' && *r <= '9')
			{
				if(ev > 1000)
				{
					return false;
				}
				ev = ev * 10 + (*r - '0');
				r++;
			}
			e10 = eneg ? -ev : ev;
			q	= r;
		}
	}
	if(*q == 'x' || *q == 'X' || ((*q | 32) >= 'a' && (*q | 32) <= 'z' && *q != 'e' && *q != 'E'))
	{
		return false; // "0x..", "inf", "nan", "1f"... : strtof's business
	}
	const int k = e10 - frac;
	if(k < -22 || k > 22)
	{
		return false;
	}
	double d = (double) mant;
	d		 = k < 0 ? d / P10[-k] : d * P10[k];
	if(d > 3.0e38 || (d != 0.0 && d < 1.0e-37))
	{
		return false; // overflow / subnormal range: strtof sets errno and rounds specially
	}
	unsigned long long bits;
	memcpy(&bits, &d, sizeof bits);
	const unsigned low = (unsigned) (bits & 0x1fffffffu); // the 29 bits a float drops
	if(low >= 0x0fffffffu && low <= 0x10000001u)
	{
		return false; // within one double ulp of a float midpoint
	}
	*out = (float) (negative ? -d : d);
	*end = q;
	return true;
}

// Reads up to `max` floats after the command word; returns how many were converted (sscanf %f semantics:
// stops at the first token that is not a number).
int read_floats_impl(const char *p, float *out, int max)
{
	int n = 0;
	while(n < max)
	{
		const char *fend = nullptr;
		float v;
		if(fast_strtof(p, &fend, &v))
		{
			out[n++] = v;
			p		 = fend;
			continue;
		}
		char *end = nullptr;
		errno	  = 0;
		v		  = strtof(p, &end);
		if(end == p)
		{
			break;
		}
		out[n++] = v;
		p		 = end;
	}
	return n;
}
} // namespace

int read_floats(const char *p, float *out, int max)
{
	return read_floats_impl(p, out, max);
}

skr_scene_desc HostScene::desc() const
{
	skr_scene_desc d;
	memset(&d, 0, sizeof d);
	d.nspheres = nspheres();
	d.spheres  = spheres.data();
	d.ntris	   = ntris();
	d.tris	   = tris.data();
	d.tri_materials = tri_materials.size() == 14 * (size_t) ntris() && ntris() > 0 ? tri_materials.data() : nullptr;
	d.nplights = nplights();
	d.plights  = plights.data();
	d.ndlights = ndlights();
	d.dlights  = dlights.data();
	d.nfogs	   = nfogs();
	d.fogs	   = fogs.data();
	memcpy(d.camera, camera, sizeof camera);
	memcpy(d.ambient, ambient, sizeof ambient);
	memcpy(d.background, background, sizeof background);
	return d;
}

bool parse_scn(const std::string &path, HostScene &scene, std::string &error, const ParseOptions &opt)
{
	FILE *fp = fopen(path.c_str(), "r");
	if(!fp)
	{
		error = "Can't open file '" + path + "'"; // src/scene.cpp:22-26 prints this and exit(0)s
		return false;
	}
	scene = HostScene();
	MaterialState mat;
	char line[1024]; // the reference assumes no line is longer than 1024 characters (src/scene.cpp:19)
	int lineno = 0;
	bool ok	   = true;
	while(fgets(line, sizeof line, fp))
	{
		lineno++;
		if(line[0] == '#')
		{
			continue; // src/scene.cpp:31-35: only a '#' in column 0 starts a comment
		}
		// first word of the line = the command (what the reference reads with sscanf("%s "))
		char command[100];
		const char *c = line;
		while(*c == ' ' || *c == '\t' || *c == '\r' || *c == '\n' || *c == '\v' || *c == '\f')
		{
			c++;
		}
		int clen = 0;
		while(*c && !(*c == ' ' || *c == '\t' || *c == '\r' || *c == '\n' || *c == '\v' || *c == '\f') && clen < 99)
		{
			command[clen++] = *c++;
		}
		command[clen] = 0;
		if(clen == 0)
		{
			continue; // blank line
		}
		const char *args = c;
		float f[16];
		for(float &v : f)
		{
			v = 0.0f;
		}
		if(strcmp(command, "sphere") == 0)
		{
			read_floats(args, f, 4); // x y z r
			const float rec[18] = {f[0], f[1], f[2], f[3], mat.ambient[0], mat.ambient[1], mat.ambient[2], mat.diffuse[0], mat.diffuse[1],
								   mat.diffuse[2], mat.specular[0], mat.specular[1], mat.specular[2], mat.transmissive[0], mat.transmissive[1],
								   mat.transmissive[2], mat.power, mat.ior};
			scene.spheres.insert(scene.spheres.end(), rec, rec + 18);
			if(opt.verbose)
			{
				printf("Sphere as position (%f, %f, %f) with radius %f\n", f[0], f[1], f[2], f[3]);
			}
		}
		else if(strcmp(command, "vertex") == 0)
		{
			read_floats(args, f, 3);
			scene.vertices.insert(scene.vertices.end(), f, f + 3);
		}
		else if(strcmp(command, "triangle") == 0)
		{
			// indices are read as floats and used as vector subscripts (src/scene.cpp:66-75)
			read_floats(args, f, 3);
			const size_t nv = scene.vertices.size() / 3;
			for(int k = 0; k < 3; k++)
			{
				if(!(f[k] >= 0.0f) || (size_t) f[k] >= nv)
				{
					char buf[256];
					snprintf(buf, sizeof buf, "%s:%d: triangle references vertex %g but only %zu vertices are defined", path.c_str(), lineno, f[k], nv);
					error = buf;
					ok	  = false;
					break;
				}
			}
			if(!ok)
			{
				break;
			}
			for(int k = 0; k < 3; k++)
			{
				const float *v = scene.vertices.data() + 3 * (size_t) f[k];
				scene.tris.insert(scene.tris.end(), v, v + 3);
			}
			{
				const float rec[14] = {mat.ambient[0], mat.ambient[1], mat.ambient[2], mat.diffuse[0], mat.diffuse[1], mat.diffuse[2], mat.specular[0],
									   mat.specular[1], mat.specular[2], mat.transmissive[0], mat.transmissive[1], mat.transmissive[2], mat.power, mat.ior};
				scene.tri_materials.insert(scene.tri_materials.end(), rec, rec + 14);
			}
		}
		else if(strcmp(command, "camera") == 0)
		{
			read_floats(args, f, 10); // pos3 dir3 up3 halfHeightAngle (the angle is never used)
			float *c = scene.camera;
			memcpy(c, f, 9 * sizeof(float));
			// Camera ctor: right = cross(direction * -1.0f, up); nothing is normalised (src/camera.h:24-31,
			// the glm::normalize results at src/scene.cpp:91-93 are discarded)
			const float nx = f[3] * -1.0f, ny = f[4] * -1.0f, nz = f[5] * -1.0f;
			c[9]  = ny * f[8] - f[7] * nz;
			c[10] = nz * f[6] - f[8] * nx;
			c[11] = nx * f[7] - f[6] * ny;
			if(opt.verbose)
			{
				printf("Camera with position (%f, %f, %f) with viewing direction (%f, %f, %f), up (%f, %f, %f), and halfHeightAngle %f\n", f[0], f[1],
					   f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9]);
			}
		}
		else if(strcmp(command, "film_resolution") == 0)
		{
			int w = scene.film_width, h = scene.film_height;
			sscanf(args, "%d %d", &w, &h);
			scene.film_width  = w;
			scene.film_height = h;
		}
		else if(strcmp(command, "background") == 0)
		{
			read_floats(args, f, 3);
			memcpy(scene.background, f, 3 * sizeof(float));
		}
		else if(strcmp(command, "material") == 0)
		{
			// ambient3 diffuse3 specular3 phongCos transmissive3 ior (src/scene.cpp:119-137)
			read_floats(args, f, 14);
			memcpy(mat.ambient, f, 3 * sizeof(float));
			memcpy(mat.diffuse, f + 3, 3 * sizeof(float));
			memcpy(mat.specular, f + 6, 3 * sizeof(float));
			mat.power = f[9];
			memcpy(mat.transmissive, f + 10, 3 * sizeof(float));
			mat.ior = f[13];
		}
		else if(strcmp(command, "directional_light") == 0)
		{
			read_floats(args, f, 6); // colour3 direction3; colour clamped to <= 1 (src/scene.cpp:144-155)
			for(int k = 0; k < 3; k++)
			{
				if(f[k] > 1)
				{
					f[k] = 1;
				}
			}
			if(opt.keep_directional)
			{
				const float rec[6] = {f[3], f[4], f[5], f[0], f[1], f[2]};
				scene.dlights.insert(scene.dlights.end(), rec, rec + 6);
			}
		}
		else if(strcmp(command, "point_light") == 0)
		{
			read_floats(args, f, 6); // colour3 position3
			const float rec[6] = {f[3], f[4], f[5], f[0], f[1], f[2]};
			scene.plights.insert(scene.plights.end(), rec, rec + 6);
		}
		else if(strcmp(command, "ambient_light") == 0)
		{
			read_floats(args, f, 3); // accumulates (src/scene.cpp:188-190)
			scene.ambient[0] += f[0];
			scene.ambient[1] += f[1];
			scene.ambient[2] += f[2];
		}
		else if(strcmp(command, "max_depth") == 0)
		{
			read_floats(args, f, 1);
			scene.max_depth = (int) f[0];
		}
		else if(strcmp(command, "output_image") == 0)
		{
			char out_file[1024] = {0};
			sscanf(args, "%1023s", out_file);
			scene.output_image = out_file;
		}
		else if(strcmp(command, "spherical_fog") == 0)
		{
			if(opt.fog)
			{
				read_floats(args, f, 9); // x y z radius r g b scattering absorption
				const float rec[9] = {f[7], f[8], f[4], f[5], f[6], f[3], f[0], f[1], f[2]};
				scene.fogs.insert(scene.fogs.end(), rec, rec + 9);
			}
		}
		else
		{
			scene.unknown_commands++;
			if(opt.verbose)
			{
				printf("WARNING. Do not know command: %s\n", command);
			}
		}
	}
	fclose(fp);
	return ok;
}

} // namespace skr_host

extern "C" int skr_host_read_floats(const char *text, float *out, int max)
{
	return skr_host::read_floats(text, out, max);
}

namespace skr_host
{
bool write_ppm(const std::string &path, int width, int height, const unsigned char *rgb8, std::string &error)
{
	FILE *fp = fopen(path.c_str(), "wb");
	if(!fp)
	{
		error = "cannot open '" + path + "' for writing";
		return false;
	}
	fprintf(fp, "P6\n%d %d\n255\n", width, height);
	const size_t n	= (size_t) width * height * 3;
	const bool good = fwrite(rgb8, 1, n, fp) == n;
	fclose(fp);
	if(!good)
	{
		error = "short write to '" + path + "'";
	}
	return good;
}
} // namespace skr_host
