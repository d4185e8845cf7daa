"""oracle/oracle_lib.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes bindings for
  * oracle/libskr_oracle.so       the plain-C port (oracle/skr_oracle.c), always buildable;
  * oracle/_ref/libskr_ref.so     the reference's own code compiled in place by oracle/Makefile
                                  (present where `make -C oracle ref` has run: this container;
                                  the built file travels to the GPU box, /root/reference does not);
  * oracle/_ref/libskr_ref_fresnel.so   same with src/raytrace.h:44 removed (row A9).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
arm import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libskr_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libskr_ref.so")
REF_FRESNEL_SO = os.path.join(HERE, "_ref", "libskr_ref_fresnel.so")
REF_O0_SO = os.path.join(HERE, "_ref", "libskr_ref_O0.so")  # the reference's own flags (src/Makefile:2: -g, no -O)
REF_UNMODIFIED = os.path.join(HERE, "_ref", "raytracer_unmodified")
REF_SCENES = os.path.join(HERE, "_ref", "scenes")

_f32p = C.POINTER(C.c_float)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f32p)


@dataclass
class Scene:
    """Flat scene snapshot (layout of include/skr.h: skr_scene_desc)."""

    spheres: np.ndarray = field(default_factory=lambda: np.zeros((0, 18), np.float32))
    tris: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    plights: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    dlights: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    fogs: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    camera: np.ndarray = field(default_factory=lambda: np.zeros(12, np.float32))
    ambient: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    background: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    # NOT part of the reference's render path: per-triangle materials (14 floats: ambient3 diffuse3 specular3
    # transmissive3 power ior), read only by this repository's shaded-triangles extension (skr_oracle_ext.inc)
    tri_materials: np.ndarray | None = None

    def normalised(self) -> "Scene":
        def a(x, shape):
            return np.ascontiguousarray(np.asarray(x, np.float32).reshape(shape))

        return Scene(a(self.spheres, (-1, 18)), a(self.tris, (-1, 9)), a(self.plights, (-1, 6)),
                     a(self.dlights, (-1, 6)), a(self.fogs, (-1, 9)), a(self.camera, (12,)),
                     a(self.ambient, (3,)), a(self.background, (3,)),
                     None if self.tri_materials is None else a(self.tri_materials, (-1, 14)))

    # ---- snapshot files (tests/golden/scenes/*.npz) ----
    def save(self, path: str) -> None:
        s = self.normalised()
        np.savez_compressed(path, spheres=s.spheres, tris=s.tris, plights=s.plights, dlights=s.dlights,
                            fogs=s.fogs, camera=s.camera, ambient=s.ambient, background=s.background)

    @staticmethod
    def load(path: str) -> "Scene":
        z = np.load(path)
        return Scene(z["spheres"], z["tris"], z["plights"], z["dlights"], z["fogs"], z["camera"],
                     z["ambient"], z["background"]).normalised()


@dataclass
class Options:
    width: int = 1920
    height: int = 1080
    fov: float = 60.0
    max_depth: int = 3
    monte_carlo: bool = False
    num_path_traces: int = 1
    grid_size: int = 0
    use_shadows: bool = False
    fresnel: bool = False
    shade_triangles: bool = False  # NOT reference behaviour (this repository's extension)


class _SkroScene(C.Structure):
    _fields_ = [("nspheres", C.c_int), ("spheres", _f32p), ("ntris", C.c_int), ("tris", _f32p),
                ("nplights", C.c_int), ("plights", _f32p), ("ndlights", C.c_int), ("dlights", _f32p),
                ("nfogs", C.c_int), ("fogs", _f32p), ("camera", C.c_float * 12), ("ambient", C.c_float * 3),
                ("background", C.c_float * 3), ("tri_materials", _f32p)]


class _SkroOptions(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fov", C.c_float), ("max_depth", C.c_int),
                ("monte_carlo", C.c_int), ("num_path_traces", C.c_int), ("grid_size", C.c_int),
                ("use_shadows", C.c_int), ("fresnel", C.c_int), ("rng_mode", C.c_int), ("seed", C.c_uint64),
                ("threads", C.c_int), ("y0", C.c_int), ("y1", C.c_int), ("shade_triangles", C.c_int)]


class SkroStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos",
                                          "tri_tests", "sphere_hits", "light_evals")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


RNG_LIBC, RNG_PHILOX = 0, 1


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "skr_oracle.c")
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


def build_ref() -> bool:
    """Compile the reference in place (only possible where /root/reference exists)."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(REF_SO)


class Port:
    """The plain-C restatement."""

    def __init__(self):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.skro_render.restype = C.c_double
        L.skro_render.argtypes = [C.POINTER(_SkroScene), C.POINTER(_SkroOptions), _f32p, C.c_void_p, C.POINTER(SkroStats)]
        L.skro_shade.argtypes = [C.POINTER(_SkroScene), C.POINTER(_SkroOptions), _f32p, _f32p, C.c_int, _f32p]
        L.skro_smallest_root.restype = C.c_float
        L.skro_smallest_root.argtypes = [C.c_float] * 3
        L.skro_sphere_hit.restype = C.c_float
        L.skro_sphere_hit.argtypes = [_f32p, _f32p, _f32p, C.c_float, C.POINTER(C.c_int)]
        L.skro_triangle_hit.restype = C.c_int
        L.skro_triangle_hit.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.skro_transform_coordinate_space.argtypes = [_f32p] * 3
        L.skro_uniform_sample_hemi.argtypes = [C.c_float, C.c_float, _f32p]
        L.skro_shadow_point.restype = C.c_int
        L.skro_shadow_point.argtypes = [C.POINTER(_SkroScene), _f32p, _f32p]
        L.skro_fresnel.restype = C.c_float
        L.skro_fresnel.argtypes = [_f32p, _f32p, C.c_float]
        L.skro_refraction.argtypes = [_f32p, _f32p, C.c_float, _f32p]
        L.skro_reflect_direction.argtypes = [_f32p, _f32p, _f32p]
        L.skro_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.skro_max_threads.restype = C.c_int

    @staticmethod
    def _scene(s: Scene):
        s = s.normalised()
        cs = _SkroScene(len(s.spheres), _fp(s.spheres), len(s.tris), _fp(s.tris), len(s.plights), _fp(s.plights),
                        len(s.dlights), _fp(s.dlights), len(s.fogs), _fp(s.fogs))
        cs.camera[:] = s.camera.tolist()
        cs.ambient[:] = s.ambient.tolist()
        cs.background[:] = s.background.tolist()
        cs.tri_materials = _fp(s.tri_materials) if s.tri_materials is not None and len(s.tri_materials) == len(s.tris) else None
        return cs, s  # keep `s` alive

    @staticmethod
    def _opts(o: Options, rng_mode, seed, threads, y0, y1):
        return _SkroOptions(o.width, o.height, o.fov, o.max_depth, int(o.monte_carlo), o.num_path_traces,
                            o.grid_size, int(o.use_shadows), int(o.fresnel), rng_mode, seed, threads,
                            0 if y0 is None else y0, o.height if y1 is None else y1, int(o.shade_triangles))

    def max_threads(self) -> int:
        return int(self.lib.skro_max_threads())

    def render(self, scene: Scene, opt: Options, rng_mode=RNG_PHILOX, seed=0, threads=0, y0=None, y1=None,
               want_rgb8=True):
        """-> (rgb32 HxWx3 float32, rgb8 HxWx3 uint8, stats dict, seconds)"""
        cs, keep = self._scene(scene)
        if threads <= 0:
            threads = 1 if rng_mode == RNG_LIBC else self.max_threads()
        co = self._opts(opt, rng_mode, seed, threads, y0, y1)
        rgb32 = np.zeros((opt.height, opt.width, 3), np.float32)
        rgb8 = np.zeros((opt.height, opt.width, 3), np.uint8) if want_rgb8 else None
        st = SkroStats()
        secs = self.lib.skro_render(C.byref(cs), C.byref(co), _fp(rgb32), None if rgb8 is None else rgb8.ctypes.data,
                                    C.byref(st))
        return rgb32, rgb8, st.as_dict(), secs

    def shade(self, scene: Scene, opt: Options, o, d, depth, rng_mode=RNG_PHILOX, seed=0):
        cs, keep = self._scene(scene)
        co = self._opts(opt, rng_mode, seed, 1, None, None)
        o = np.asarray(o, np.float32)
        d = np.asarray(d, np.float32)
        out = np.zeros(3, np.float32)
        self.lib.skro_shade(C.byref(cs), C.byref(co), _fp(o), _fp(d), depth, _fp(out))
        return out

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.skro_philox4x32_10(c, k, o)
        return list(o)


class Ref:
    """The reference's own compiled code (oracle/_ref)."""

    def __init__(self, fresnel: bool = False, unoptimised: bool = False):
        path = REF_FRESNEL_SO if fresnel else (REF_O0_SO if unoptimised else REF_SO)
        if not os.path.exists(path):
            build_ref()
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_parse_scene.restype = C.c_void_p
        L.ref_parse_scene.argtypes = [C.c_char_p]
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_scene_export.argtypes = [C.c_void_p] + [_f32p] * 8
        L.ref_scene_from_arrays.restype = C.c_void_p
        L.ref_scene_from_arrays.argtypes = [C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int,
                                            _f32p, _f32p, _f32p, _f32p]
        L.ref_render.restype = C.c_double
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_uint, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p]
        L.ref_smallest_root.restype = C.c_float
        L.ref_smallest_root.argtypes = [C.c_float] * 3
        L.ref_sphere_hit.restype = C.c_float
        L.ref_sphere_hit.argtypes = [_f32p, _f32p, _f32p, C.c_float, C.POINTER(C.c_int)]
        L.ref_triangle_hit.restype = C.c_int
        L.ref_triangle_hit.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.ref_transform_coordinate_space.argtypes = [_f32p] * 3
        L.ref_uniform_sample_hemi.argtypes = [C.c_float, C.c_float, _f32p]
        L.ref_shade.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_uint, _f32p]
        L.ref_shadow_point.restype = C.c_int
        L.ref_shadow_point.argtypes = [C.c_void_p, _f32p, _f32p]
        L.ref_fresnel.restype = C.c_float
        L.ref_fresnel.argtypes = [_f32p, _f32p, C.c_float]
        L.ref_refraction.argtypes = [_f32p, _f32p, C.c_float, _f32p]
        L.ref_reflect_direction.argtypes = [_f32p, _f32p, _f32p]
        L.ref_max_threads.restype = C.c_int

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    def parse(self, path: str) -> Scene:
        """Parse a .scn with the reference's parser (src/scene.cpp) and export the snapshot."""
        h = self.lib.ref_parse_scene(path.encode())
        if not h:
            raise FileNotFoundError(path)
        try:
            cnt = (C.c_int * 7)()
            self.lib.ref_scene_counts(h, cnt)
            s = Scene(np.zeros((cnt[0], 18), np.float32), np.zeros((cnt[1], 9), np.float32),
                      np.zeros((cnt[2], 6), np.float32), np.zeros((cnt[3], 6), np.float32),
                      np.zeros((cnt[4], 9), np.float32), np.zeros(12, np.float32), np.zeros(3, np.float32),
                      np.zeros(3, np.float32))
            self.lib.ref_scene_export(h, _fp(s.spheres), _fp(s.tris), _fp(s.plights), _fp(s.dlights), _fp(s.fogs),
                                      _fp(s.camera), _fp(s.ambient), _fp(s.background))
            s.film = (cnt[5], cnt[6])
            return s
        finally:
            self.lib.ref_scene_free(h)

    def _handle(self, scene: Scene):
        s = scene.normalised()
        return self.lib.ref_scene_from_arrays(len(s.spheres), _fp(s.spheres), len(s.tris), _fp(s.tris),
                                              len(s.plights), _fp(s.plights), len(s.dlights), _fp(s.dlights),
                                              len(s.fogs), _fp(s.fogs), _fp(s.camera), _fp(s.ambient),
                                              _fp(s.background))

    def render(self, scene: Scene, opt: Options, seed=0, threads=1, y0=None, y1=None):
        """-> (rgb32, rgb8, seconds).  threads=1 gives the reproducible rand() stream."""
        assert not opt.fresnel or self.lib._name.endswith("fresnel.so")
        h = self._handle(scene)
        try:
            rgb32 = np.zeros((opt.height, opt.width, 3), np.float32)
            rgb8 = np.zeros((opt.height, opt.width, 3), np.uint8)
            secs = self.lib.ref_render(h, opt.width, opt.height, opt.fov, opt.max_depth, int(opt.monte_carlo),
                                       opt.num_path_traces, opt.grid_size, int(opt.use_shadows), seed, threads,
                                       0 if y0 is None else y0, opt.height if y1 is None else y1, _fp(rgb32),
                                       rgb8.ctypes.data)
            return rgb32, rgb8, secs
        finally:
            self.lib.ref_scene_free(h)

    def shade(self, scene: Scene, opt: Options, o, d, depth, seed=0):
        h = self._handle(scene)
        try:
            o = np.asarray(o, np.float32)
            d = np.asarray(d, np.float32)
            out = np.zeros(3, np.float32)
            self.lib.ref_shade(h, int(opt.use_shadows), _fp(o), _fp(d), depth, int(opt.monte_carlo),
                               opt.num_path_traces, seed, _fp(out))
            return out
        finally:
            self.lib.ref_scene_free(h)


def ppm_bytes(rgb8: np.ndarray) -> bytes:
    """src/main.cpp:88-100: "P6\\n<w> <h>\\n255\\n" + row-major RGB."""
    h, w, _ = rgb8.shape
    return b"P6\n%d %d\n255\n" % (w, h) + rgb8.tobytes()


def ref_available() -> bool:
    return os.path.exists(REF_SO)
