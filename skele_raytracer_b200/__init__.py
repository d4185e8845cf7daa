"""skele_raytracer_b200 -- B200-native renderer core for skele-raytracer's per-pixel tracing loop.

The product is `libskr.so` (hand-written sm_100a CUDA behind the C ABI in include/skr.h) plus the C++ host
front end in host/.  This Python package is a thin ctypes mirror of the reference's own interface for that
path (`Scene`, `Options`, `parseScene`, `generate_rays_parallel`; reference src/scene.h, src/utils.h:26-39,
src/main.cpp:19) used by the parity tests and bench.py.  It has no rendering code of its own and no CPU
fallback: importing `api` without a built libskr.so raises.
"""
from .api import (MgpuRenderer, Options, Renderer, Scene, SkrError, Stats, build_info, generate_rays_parallel, lib_path,  # noqa: F401
                  parseScene, write_ppm)
