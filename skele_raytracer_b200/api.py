"""ctypes mirror of include/skr.h and of the reference's interface to the frame function.

Reference interface mirrored (names and argument meaning):
    struct Scene                      src/scene.h:13-28      -> class Scene (flat arrays, see skr.h)
    struct Options                    src/utils.h:26-39      -> class Options
    Scene parseScene(std::string)     src/scene.cpp:12       -> parseScene(path)          (host/scene_parser.cpp)
    void generate_rays_parallel(Scene, Options, char *output)
                                      src/main.cpp:19        -> generate_rays_parallel(scene, option, output)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = C.POINTER(C.c_float)


class SkrError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "libskr.so")


def build_info() -> str:
    """skr_build_info(): how the loaded libskr.so was compiled."""
    L = _load("libskr.so")
    L.skr_build_info.restype = C.c_char_p
    return L.skr_build_info().decode()


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if name == "libskr.so" and os.environ.get("SKR_LIB"):  # developer override: A/B-test a differently built library
        path = os.environ["SKR_LIB"]
    if not os.path.exists(path):
        raise SkrError(f"{path} is missing: build it with `make -C host` (or __graft_entry__.build()). "
                       "There is no CPU fallback.")
    return C.CDLL(path)


class _SceneDesc(C.Structure):
    _fields_ = [("nspheres", C.c_int32), ("spheres", _f32p), ("ntris", C.c_int32), ("tris", _f32p),
                ("nplights", C.c_int32), ("plights", _f32p), ("ndlights", C.c_int32), ("dlights", _f32p),
                ("nfogs", C.c_int32), ("fogs", _f32p), ("camera", C.c_float * 12), ("ambient", C.c_float * 3),
                ("background", C.c_float * 3), ("tri_materials", _f32p)]


class _Options(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fov", C.c_float), ("max_depth", C.c_int32),
                ("monte_carlo", C.c_int32), ("num_path_traces", C.c_int32), ("grid_size", C.c_int32),
                ("use_shadows", C.c_int32), ("fresnel", C.c_int32), ("seed", C.c_uint64), ("rank", C.c_int32),
                ("world", C.c_int32), ("tile", C.c_int32), ("collect_stats", C.c_int32),
                ("queue_capacity", C.c_int32), ("shade_triangles", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("closest_hit_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("sphere_tests_pos", C.c_uint64), ("tri_tests", C.c_uint64), ("bvh_node_visits", C.c_uint64),
                ("sphere_hits", C.c_uint64), ("light_evals", C.c_uint64), ("queue_entries", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("queue_chunks", C.c_uint32), ("ms_total", C.c_float),
                ("ms_primary", C.c_float), ("ms_bounce", C.c_float), ("ms_resolve", C.c_float), ("ms_h2d", C.c_float),
                ("ms_d2h", C.c_float), ("sphere_tests_executed", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


@dataclass
class Scene:
    """Flat mirror of the reference's `Scene` (src/scene.h:13-28); array layouts in include/skr.h."""

    spheres: np.ndarray = field(default_factory=lambda: np.zeros((0, 18), np.float32))
    tris: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    plights: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    dlights: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    fogs: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    camera: np.ndarray = field(default_factory=lambda: np.zeros(12, np.float32))
    ambient: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    background: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    tri_materials: np.ndarray | None = None  # [ntris][14], only read by the shaded-triangles mode (see skr.h)

    def _norm(self) -> "Scene":
        def a(x, shape):
            return np.ascontiguousarray(np.asarray(x, np.float32).reshape(shape))

        return Scene(a(self.spheres, (-1, 18)), a(self.tris, (-1, 9)), a(self.plights, (-1, 6)),
                     a(self.dlights, (-1, 6)), a(self.fogs, (-1, 9)), a(self.camera, (12,)), a(self.ambient, (3,)),
                     a(self.background, (3,)), None if self.tri_materials is None else a(self.tri_materials, (-1, 14)))

    def _desc(self):
        s = self._norm()
        d = _SceneDesc(len(s.spheres), s.spheres.ctypes.data_as(_f32p), len(s.tris), s.tris.ctypes.data_as(_f32p),
                       len(s.plights), s.plights.ctypes.data_as(_f32p), len(s.dlights),
                       s.dlights.ctypes.data_as(_f32p), len(s.fogs), s.fogs.ctypes.data_as(_f32p))
        d.camera[:] = s.camera.tolist()
        d.ambient[:] = s.ambient.tolist()
        d.background[:] = s.background.tolist()
        if s.tri_materials is not None and len(s.tri_materials) == len(s.tris) and len(s.tris):
            d.tri_materials = s.tri_materials.ctypes.data_as(_f32p)
        return d, s

    def save(self, path: str) -> None:
        s = self._norm()
        np.savez_compressed(path, spheres=s.spheres, tris=s.tris, plights=s.plights, dlights=s.dlights, fogs=s.fogs,
                            camera=s.camera, ambient=s.ambient, background=s.background)

    @staticmethod
    def load(path: str) -> "Scene":
        z = np.load(path)
        return Scene(z["spheres"], z["tris"], z["plights"], z["dlights"], z["fogs"], z["camera"], z["ambient"],
                     z["background"])._norm()


@dataclass
class Options:
    """Mirror of the reference's `Options` (src/utils.h:26-39) + the per-frame Scene fields main() sets
    (width, height, use_shadows; src/main.cpp:393-396).  Defaults are the reference's."""

    width: int = 1920
    height: int = 1080
    fov: float = 60.0
    max_depth: int = 3
    monte_carlo: bool = False
    num_path_traces: int = 1
    grid_size: int = 0
    use_shadows: bool = False
    fresnel: bool = False
    seed: int = 0
    rank: int = 0
    world: int = 1
    tile: int = 0
    collect_stats: bool = False
    queue_capacity: int = 0
    shade_triangles: bool = False  # opt-in NON-PARITY extension: triangles shaded with their own materials (skr.h)

    def _c(self) -> _Options:
        return _Options(self.width, self.height, self.fov, self.max_depth, int(self.monte_carlo), self.num_path_traces,
                        self.grid_size, int(self.use_shadows), int(self.fresnel), self.seed, self.rank, self.world,
                        self.tile, int(self.collect_stats), self.queue_capacity, int(self.shade_triangles))


_host = None


def _host_lib():
    global _host
    if _host is None:
        L = _load("libskr_host.so")
        L.skr_host_parse_scn.restype = C.c_void_p
        L.skr_host_parse_scn.argtypes = [C.c_char_p, C.c_int]
        L.skr_host_scene_error.restype = C.c_char_p
        L.skr_host_scene_error.argtypes = [C.c_void_p]
        L.skr_host_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.skr_host_scene_desc.argtypes = [C.c_void_p, C.POINTER(_SceneDesc)]
        L.skr_host_scene_free.argtypes = [C.c_void_p]
        L.skr_host_write_ppm.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        _host = L
    return _host


def parseScene(fileName: str, keep_directional: bool = False, fog: bool = True) -> Scene:
    """Scene parseScene(std::string fileName) -- the host-side C++ reader (host/scene_parser.cpp)."""
    L = _host_lib()
    h = L.skr_host_parse_scn(os.fsencode(fileName), (1 if keep_directional else 0) | (0 if fog else 2))
    try:
        err = L.skr_host_scene_error(h)
        if err:
            raise SkrError(err.decode())
        d = _SceneDesc()
        L.skr_host_scene_desc(h, C.byref(d))
        cnt = (C.c_int * 9)()
        L.skr_host_scene_counts(h, cnt)

        def arr(ptr, n, w):
            if n == 0:
                return np.zeros((0, w), np.float32)
            return np.ctypeslib.as_array(ptr, shape=(n, w)).copy()

        s = Scene(arr(d.spheres, d.nspheres, 18), arr(d.tris, d.ntris, 9), arr(d.plights, d.nplights, 6),
                  arr(d.dlights, d.ndlights, 6), arr(d.fogs, d.nfogs, 9), np.array(d.camera[:], np.float32),
                  np.array(d.ambient[:], np.float32), np.array(d.background[:], np.float32),
                  arr(d.tri_materials, d.ntris, 14) if d.tri_materials and d.ntris else None)
        s.film_resolution = (cnt[5], cnt[6])
        s.max_depth = cnt[7]
        s.unknown_commands = cnt[8]
        return s
    finally:
        L.skr_host_scene_free(h)


def write_ppm(path: str, rgb8: np.ndarray) -> None:
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    h, w, _ = rgb8.shape
    if _host_lib().skr_host_write_ppm(os.fsencode(path), w, h, rgb8.ctypes.data):
        raise SkrError(f"cannot write {path}")


class Renderer:
    """One skr_ctx (one CUDA device)."""

    def __init__(self, device: int = -1):
        L = _load("libskr.so")
        L.skr_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.skr_destroy.argtypes = [C.c_void_p]
        L.skr_last_error.restype = C.c_char_p
        L.skr_last_error.argtypes = [C.c_void_p]
        L.skr_scene_upload.argtypes = [C.c_void_p, C.POINTER(_SceneDesc)]
        L.skr_render.argtypes = [C.c_void_p, C.POINTER(_Options), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.skr_reserve.argtypes = [C.c_void_p, C.POINTER(_Options)]
        L.skr_render_device.argtypes = [C.c_void_p, C.POINTER(_Options), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.skr_tiles_bytes.restype = C.c_int64
        L.skr_tiles_bytes.argtypes = [C.POINTER(_Options)]
        L.skr_render_tiles_device.argtypes = [C.c_void_p, C.POINTER(_Options), C.c_void_p, C.POINTER(Stats)]
        L.skr_deinterleave_device.argtypes = [C.c_void_p, C.POINTER(_Options), C.c_void_p, C.c_void_p]
        L.skr_render_peers_device.argtypes = [C.c_void_p, C.POINTER(_Options), C.POINTER(C.c_void_p), C.c_int, C.POINTER(Stats)]
        L.skr_render_bands_device.argtypes = [C.c_void_p, C.POINTER(_Options), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(Stats)]
        L.skr_stream.restype = C.c_void_p
        L.skr_stream.argtypes = [C.c_void_p]
        L.skr_sync.argtypes = [C.c_void_p]
        L.skr_measure_fp32_peak.restype = C.c_double
        L.skr_measure_fp32_peak.argtypes = [C.c_void_p, C.c_int]
        L.skr_pin_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.skr_unpin_host.argtypes = [C.c_void_p, C.c_void_p]
        L.skr_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.skr_measure_bandwidth.restype = C.c_double
        L.skr_measure_bandwidth.argtypes = [C.c_void_p, C.c_int]
        L.skr_abi_version.restype = C.c_int
        self.lib = L
        self.ctx = C.c_void_p()
        rc = L.skr_init(device, C.byref(self.ctx))
        if rc:
            msg = L.skr_last_error(None).decode()
            self.ctx = None
            raise SkrError(f"skr_init failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.skr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise SkrError(f"{what} failed ({rc}): {self.lib.skr_last_error(self.ctx).decode()}")

    def upload(self, scene: Scene) -> None:
        d, keep = scene._desc()
        self._check(self.lib.skr_scene_upload(self.ctx, C.byref(d)), "skr_scene_upload")

    def reserve(self, option: Options) -> None:
        """skr_reserve: allocate queues / accumulators / frame and load the kernels ahead of the first frame."""
        o = option._c()
        self._check(self.lib.skr_reserve(self.ctx, C.byref(o)), "skr_reserve")

    def render(self, option: Options, rgb8: np.ndarray | None = None, rgb32: np.ndarray | None = None,
               want_rgb8: bool = True, want_rgb32: bool = True):
        """skr_render with HOST buffers -> (rgb32 | None, rgb8 | None, Stats)."""
        if rgb8 is None and want_rgb8:
            rgb8 = np.zeros((option.height, option.width, 3), np.uint8)
        if rgb32 is None and want_rgb32:
            rgb32 = np.zeros((option.height, option.width, 3), np.float32)
        st = Stats()
        o = option._c()
        self._check(self.lib.skr_render(self.ctx, C.byref(o), None if rgb8 is None else rgb8.ctypes.data,
                                        None if rgb32 is None else rgb32.ctypes.data, C.byref(st)), "skr_render")
        return rgb32, rgb8, st

    def render_device(self, option: Options, d_rgb8: int = 0, d_rgb32: int = 0, want_stats: bool = True):
        """skr_render_device: raw device pointers (ints), e.g. torch.Tensor.data_ptr().
        want_stats=False passes stats=NULL: asynchronous for single-kernel frames (see include/skr.h)."""
        st = Stats() if want_stats else None
        o = option._c()
        self._check(self.lib.skr_render_device(self.ctx, C.byref(o), d_rgb8 or None, d_rgb32 or None,
                                               C.byref(st) if want_stats else None), "skr_render_device")
        return st

    def tiles_bytes(self, option: Options) -> int:
        o = option._c()
        return int(self.lib.skr_tiles_bytes(C.byref(o)))

    def render_tiles_device(self, option: Options, d_tiles: int, want_stats: bool = True):
        st = Stats() if want_stats else None
        o = option._c()
        self._check(self.lib.skr_render_tiles_device(self.ctx, C.byref(o), d_tiles, C.byref(st) if want_stats else None),
                    "skr_render_tiles_device")
        return st

    def render_peers_device(self, option: Options, d_frames, want_stats: bool = True):
        """skr_render_peers_device: d_frames = device pointers (ints) of the row-major RGB8 frames to fill."""
        st = Stats() if want_stats else None
        o = option._c()
        arr = (C.c_void_p * len(d_frames))(*[int(p) for p in d_frames])
        self._check(self.lib.skr_render_peers_device(self.ctx, C.byref(o), arr, len(d_frames), C.byref(st) if want_stats else None),
                    "skr_render_peers_device")
        return st

    def render_bands_device(self, option: Options, d_frames, rows_per_frame: int, want_stats: bool = True):
        """skr_render_bands_device: a finished pixel of image row y goes to d_frames[min(y // rows_per_frame, len - 1)]."""
        st = Stats() if want_stats else None
        o = option._c()
        arr = (C.c_void_p * len(d_frames))(*[int(p) for p in d_frames])
        self._check(self.lib.skr_render_bands_device(self.ctx, C.byref(o), arr, len(d_frames), rows_per_frame, C.byref(st) if want_stats else None),
                    "skr_render_bands_device")
        return st

    def deinterleave_device(self, option: Options, d_gathered: int, d_rgb8: int) -> None:
        o = option._c()
        self._check(self.lib.skr_deinterleave_device(self.ctx, C.byref(o), d_gathered, d_rgb8),
                    "skr_deinterleave_device")

    def pin_host(self, host_ptr: int, nbytes: int):
        """skr_pin_host -> (device pointer usable as a frame of render_peers_device, registered_here: bool)."""
        d = C.c_void_p()
        rc = self.lib.skr_pin_host(self.ctx, host_ptr, nbytes, C.byref(d))
        if rc not in (0, 1000):
            self._check(rc, "skr_pin_host")
        return int(d.value or 0), rc == 0

    def copy_to_host(self, host_ptr: int, d_ptr: int, nbytes: int) -> None:
        """skr_copy_to_host: D2H on the library's stream (behind the frames rendered so far); sync() waits for it."""
        self._check(self.lib.skr_copy_to_host(self.ctx, host_ptr, d_ptr, nbytes), "skr_copy_to_host")

    def unpin_host(self, host_ptr: int) -> None:
        self._check(self.lib.skr_unpin_host(self.ctx, host_ptr), "skr_unpin_host")

    def stream(self) -> int:
        return int(self.lib.skr_stream(self.ctx) or 0)

    def sync(self) -> None:
        self._check(self.lib.skr_sync(self.ctx), "skr_sync")

    def measure_fp32_peak(self, iters: int = 4096) -> float:
        return float(self.lib.skr_measure_fp32_peak(self.ctx, iters))

    def measure_bandwidth(self, level: int) -> float:
        """GB/s of shared memory (0), L1 (1) or L2 (2) reads, measured by a microbenchmark on this device."""
        return float(self.lib.skr_measure_bandwidth(self.ctx, level))


class MgpuRenderer:
    """include/skr_mgpu.h: ONE process driving every GPU of the box (one host thread and one skr_ctx per GPU); the frame
    arrives in host memory, every GPU storing its tiles over its own PCIe link."""

    def __init__(self, n_gpus: int = 0):
        path = os.path.join(_HERE, "libskr_mgpu.so")
        if not os.path.exists(path):
            raise SkrError(f"{path} is missing (built only where nccl.h is installed): `make -C host`")
        C.CDLL(lib_path(), mode=C.RTLD_GLOBAL)
        L = C.CDLL(path)
        L.skr_mgpu_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.skr_mgpu_destroy.argtypes = [C.c_void_p]
        L.skr_mgpu_last_error.restype = C.c_char_p
        L.skr_mgpu_last_error.argtypes = [C.c_void_p]
        L.skr_mgpu_world.argtypes = [C.c_void_p]
        L.skr_mgpu_scene_upload.argtypes = [C.c_void_p, C.POINTER(_SceneDesc)]
        L.skr_mgpu_render.argtypes = [C.c_void_p, C.POINTER(_Options), C.c_void_p, C.POINTER(Stats)]
        self.lib = L
        self.h = C.c_void_p()
        rc = L.skr_mgpu_init(n_gpus, C.byref(self.h))
        if rc:
            self.h = None
            raise SkrError(f"skr_mgpu_init failed ({rc}): {L.skr_mgpu_last_error(None).decode()}")

    def world(self) -> int:
        return int(self.lib.skr_mgpu_world(self.h))

    def _check(self, rc, what):
        if rc:
            raise SkrError(f"{what} failed ({rc}): {self.lib.skr_mgpu_last_error(self.h).decode()}")

    def upload(self, scene: Scene) -> None:
        d, keep = scene._desc()
        self._check(self.lib.skr_mgpu_scene_upload(self.h, C.byref(d)), "skr_mgpu_scene_upload")

    def render(self, option: Options, rgb8: np.ndarray, want_stats: bool = False):
        st = Stats() if want_stats else None
        o = option._c()
        self._check(self.lib.skr_mgpu_render(self.h, C.byref(o), rgb8.ctypes.data, C.byref(st) if want_stats else None), "skr_mgpu_render")
        return st

    def close(self):
        if getattr(self, "h", None):
            self.lib.skr_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_rays_parallel(scene: Scene, option: Options, output: str, renderer: Renderer | None = None) -> np.ndarray:
    """void generate_rays_parallel(Scene scene, Options option, char *output)  (reference src/main.cpp:19-104):
    renders the frame and writes the binary PPM.  Returns the RGB8 image as well."""
    r = renderer or Renderer()
    try:
        r.upload(scene)
        _, rgb8, _ = r.render(option, want_rgb32=False)
        write_ppm(output, rgb8)
        return rgb8
    finally:
        if renderer is None:
            r.close()
