"""The C-ABI boundary: libskr.so loads and exports every symbol include/skr.h declares, the structs the Python
mirror passes have the header's layout, and -- with no GPU -- the library fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

import skele_raytracer_b200 as S
from skele_raytracer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "skr.h")).read()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(skr_[a-z0-9_]+)\s*\(", body)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(S.lib_path())
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    lib.skr_abi_version.restype = C.c_int
    assert lib.skr_abi_version() == int(re.search(r"#define SKR_ABI_VERSION (\d+)", HEADER).group(1))


def test_multi_gpu_library_exports_its_header():
    path = os.path.join(ROOT, "skele_raytracer_b200", "libskr_mgpu.so")
    if not os.path.exists(path):
        pytest.skip("built without NCCL")
    hdr = open(os.path.join(ROOT, "include", "skr_mgpu.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(skr_mgpu_[a-z0-9_]+)\s*\(", body)))
    assert len(names) >= 6
    C.CDLL(S.lib_path(), mode=C.RTLD_GLOBAL)
    lib = C.CDLL(path)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts_match_header():
    # field order of the ctypes mirrors == field order in the header
    for cname, ctype in (("skr_scene_desc", api._SceneDesc), ("skr_options", api._Options), ("skr_stats", api.Stats)):
        m = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (cname, cname), HEADER, re.S)
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        fields = re.findall(r"(\w+)\s*(?:\[\d+\])?\s*;", body)
        assert fields == [f[0] for f in ctype._fields_], cname
    # and the sizes the C compiler gives the header's structs
    src = '#include <stdio.h>\n#include "skr.h"\nint main(){printf("%zu %zu %zu", sizeof(skr_scene_desc), sizeof(skr_options), sizeof(skr_stats));return 0;}'
    exe = "/tmp/skr_sizeof_%d" % os.getpid()
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    os.remove(exe)
    assert sizes == [C.sizeof(api._SceneDesc), C.sizeof(api._Options), C.sizeof(api.Stats)]


def test_tiles_bytes_is_host_arithmetic():
    lib = C.CDLL(S.lib_path())
    lib.skr_tiles_bytes.restype = C.c_int64
    from skele_raytracer_b200 import tiles as T
    for w, h, world, tile in [(1920, 1080, 1, 0), (1920, 1080, 8, 32), (3840, 2160, 4, 64), (100, 37, 3, 16)]:
        o = S.Options(width=w, height=h, world=world, tile=tile)._c()
        assert lib.skr_tiles_bytes(C.byref(o)) == T.tiles_bytes(w, h, world, tile or T.DEFAULT_TILE)


def _cuda_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_present(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(S.SkrError, match="no CUDA device"):
        S.Renderer()
    exe = os.path.join(ROOT, "host", "raytracer")
    if os.path.exists(exe):
        d = os.path.join(ROOT, "tests", "golden")
        scn = os.path.join(d, "tiny.scn")
        r = subprocess.run([exe, "--path", scn, "--output", "/tmp/skr_never.ppm"], capture_output=True, text=True)
        assert r.returncode == 1 and "skr_init failed" in r.stderr


def test_cli_flag_errors_mirror_the_reference():
    exe = os.path.join(ROOT, "host", "raytracer")
    if not os.path.exists(exe):
        pytest.skip("host/raytracer not built")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "no scene file was passed" in r.stderr          # src/main.cpp:381-385
    r = subprocess.run([exe, "--path", "x.scn"], capture_output=True, text=True)
    assert r.returncode == 0 and "no output destination was passed" in r.stderr  # src/main.cpp:387-391
    r = subprocess.run([exe, "--path", "x.scn", "--output", "o.ppm", "--depth", "0"], capture_output=True, text=True)
    assert r.returncode == 0 and "depth takes a positive int" in r.stderr        # src/main.cpp:318-330
    r = subprocess.run([exe, "--path", "/nonexistent.scn", "--output", "o.ppm"], capture_output=True, text=True)
    assert r.returncode == 0 and "Can't open file" in r.stdout                   # src/scene.cpp:22-26
