"""The drop-in, dropped in: the reference's UNMODIFIED main.cpp + scene.cpp (its flag scan, its parser, its main()) with the
body of generate_rays_parallel (src/main.cpp:19-104) replaced by host/dropin/generate_rays_parallel_skr.inc and linked
against libskr.so (`make -C oracle dropin`, built where /root/reference exists; the binary travels in oracle/_ref/).

  * on the GPU box: its PPMs are byte-identical to host/raytracer's for the same flags;
  * one full-size deterministic frame of the CUDA path against the reference's own compiled code (libskr_ref.so),
    closing the chain GPU -> port -> reference at 1920x1080;
  * without a GPU: the patched reference program fails loudly (no CPU fallback hides behind the seam).
"""
import os
import subprocess

import numpy as np
import pytest

import skele_raytracer_b200 as S
from conftest import ROOT
from oracle import oracle_lib as O
from parity import assert_image_parity

DROPIN = os.path.join(ROOT, "oracle", "_ref", "raytracer_dropin")
EXE = os.path.join(ROOT, "host", "raytracer")


def _need():
    if not os.path.exists(DROPIN) or not os.path.isdir(O.REF_SCENES):
        pytest.skip("oracle/_ref/raytracer_dropin not built (needs /root/reference: make -C oracle dropin)")


def _cuda_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.gpu
@pytest.mark.parametrize("scene,flags", [("spheres1", ["--width", "640", "--height", "360"]),
                                         ("bear", ["--width", "800", "--height", "450", "--fov", "50"]),
                                         ("dragon", ["--width", "640", "--height", "480", "--depth", "1"]),
                                         ("test", ["--width", "320", "--height", "200", "--jsample", "2", "--gillum", "3", "--depth", "2"])])
def test_reference_main_with_the_binding_writes_the_same_ppm(tmp_path, scene, flags):
    """Same .scn, same flags: the reference's own front end + libskr.so against this repo's front end + libskr.so."""
    _need()
    scn = os.path.join(O.REF_SCENES, scene + ".scn")
    a, b = tmp_path / "dropin.ppm", tmp_path / "host.ppm"
    env = dict(os.environ, SKR_SEED="9")
    r = subprocess.run([DROPIN, "--path", scn, "--output", str(a), "--parallel", "true", "--shadow"] + flags, capture_output=True, text=True, timeout=300,
                       env=env, cwd=str(tmp_path))  # (the reference parser drops simplesphere.txt into the CWD, src/scene.cpp:96-102)
    assert r.returncode == 0 and "WROTE TO PPM" in r.stdout, r.stderr[-2000:]
    subprocess.run([EXE, "--path", scn, "--output", str(b), "--parallel", "true", "--shadow", "--seed", "9"] + flags, check=True, capture_output=True,
                   timeout=300)
    assert a.read_bytes() == b.read_bytes()


@pytest.mark.gpu
def test_full_size_frame_against_the_reference_itself(scenes):
    """BASELINE config 1 with shadows at 1920x1080: CUDA path vs oracle/_ref/libskr_ref.so (the reference's shade() and
    below, compiled from its own sources), not vs the port."""
    if not O.ref_available():
        pytest.skip("oracle/_ref/libskr_ref.so not shipped")
    ref = O.Ref()
    oo = O.Options(width=1920, height=1080, max_depth=3, use_shadows=True)
    r32, r8, _ = ref.render(scenes["spheres1"], oo, threads=ref.max_threads())
    g = S.Renderer()
    try:
        s = scenes["spheres1"]
        g.upload(S.Scene(s.spheres, s.tris, s.plights, s.dlights, s.fogs, s.camera, s.ambient, s.background))
        g32, g8, _ = g.render(S.Options(width=1920, height=1080, max_depth=3, use_shadows=True))
    finally:
        g.close()
    assert_image_parity(g32, r32, g8, r8, what="spheres1 1080p --shadow vs libskr_ref.so")


@pytest.mark.skipif(_cuda_present(), reason="only meaningful on a box without a GPU")
def test_patched_reference_program_has_no_cpu_fallback(tmp_path):
    _need()
    scn = os.path.join(O.REF_SCENES, "spheres1.scn")
    r = subprocess.run([DROPIN, "--path", scn, "--output", str(tmp_path / "x.ppm"), "--parallel", "true", "--shadow"], capture_output=True, text=True,
                       timeout=120, cwd=str(tmp_path))
    assert r.returncode == 1 and "no CUDA device" in r.stderr and not (tmp_path / "x.ppm").exists()
