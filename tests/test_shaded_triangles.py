"""The opt-in SHADED-TRIANGLES mode (skr_options.shade_triangles, `--shade-triangles`) -- explicitly NOT
reference behaviour: the reference shades every triangle hit black (src/raytrace.h:221-224).  The checker is
oracle/skr_oracle_ext.inc, a brute-force CPU statement of this repository's own extension (flagged non-reference there);
the CUDA path answers the same queries through a second LBVH over the actual (un-mirrored) triangles.
Default-mode parity is untouched: every other test file runs with shade_triangles off.
"""
import os
import subprocess

import numpy as np
import pytest

import skele_raytracer_b200 as S
from conftest import GOLDEN, ROOT, random_scene
from oracle import oracle_lib as O
from parity import assert_image_parity


def one_triangle_scene():
    s = O.Scene()
    s.tris = np.array([[-2, -2, 5, 2, -2, 5, 0, 2, 5]], np.float32)
    s.tri_materials = np.array([[0.5, 0.5, 0.5, 0.8, 0.6, 0.4, 0, 0, 0, 0, 0, 0, 1, 1]], np.float32)
    s.plights = np.array([[0, 0, 0, 10, 10, 10]], np.float32)
    d, up = np.array([0, 0, 1], np.float32), np.array([0, 1, 0], np.float32)
    s.camera = np.concatenate([np.zeros(3, np.float32), d, up, np.cross(-d, up)]).astype(np.float32)
    s.ambient = np.array([0.2, 0.2, 0.2], np.float32)
    s.background = np.array([0.1, 0.2, 0.3], np.float32)
    return s.normalised()


def test_extension_oracle_known_answer(port):
    """A triangle facing the camera, the light at the camera: the centre pixel is ambient * ka + kd * Lc / d^2 (cos = 1);
    with the extension off the frame is the reference's black-or-background."""
    s = one_triangle_scene()
    img, _, _, _ = port.render(s, O.Options(width=33, height=33, max_depth=1, shade_triangles=True))
    c = img[16, 16]
    want = 0.2 * 0.5 + np.array([0.8, 0.6, 0.4]) * 10 / 25.0
    assert np.allclose(c, want, rtol=2e-3), (c, want)
    assert np.allclose(img[0, 0], s.background)
    # extension off = the reference: a 2-colour image, black where its (mirrored) triangle test passes
    ref, _, _, _ = port.render(s, O.Options(width=33, height=33, max_depth=1))
    black = (ref == 0).all(axis=2)
    assert black.any() and np.allclose(ref[~black], s.background)


def to_gpu(s):
    return S.Scene(s.spheres, s.tris, s.plights, s.dlights, s.fogs, s.camera, s.ambient, s.background, s.tri_materials)


def from_parsed(p):
    return O.Scene(p.spheres, p.tris, p.plights, p.dlights, p.fogs, p.camera, p.ambient, p.background, p.tri_materials).normalised()


@pytest.fixture(scope="module")
def gpu():
    r = S.Renderer()
    yield r
    r.close()


@pytest.mark.gpu
def test_one_triangle(gpu, port):
    s = one_triangle_scene()
    kw = dict(width=160, height=120, max_depth=1, use_shadows=True, shade_triangles=True)
    p32, p8, _, _ = port.render(s, O.Options(**kw))
    gpu.upload(to_gpu(s))
    g32, g8, _ = gpu.render(S.Options(**kw))
    assert_image_parity(g32, p32, g8, p8, what="one triangle")


@pytest.mark.gpu
@pytest.mark.parametrize("scene,kw", [("dragon", dict(width=320, height=180, use_shadows=True)),
                                      ("test", dict(width=320, height=180, use_shadows=True)),
                                      ("test", dict(width=160, height=90, grid_size=2, use_shadows=True, seed=3)),
                                      ("dragon", dict(width=256, height=144, fov=35.0))])
def test_reference_scenes_lit(gpu, port, scene, kw):
    """dragon.scn (10 002 triangles, its directional light kept) and test.scn (4 spheres + 1800 triangles, a point light):
    BVH closest hit + shadow any-hit on the device == brute force on the CPU, within 1/255 on >= 99.9 % of the pixels."""
    d = O.REF_SCENES
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref/scenes not shipped")
    parsed = S.parseScene(os.path.join(d, scene + ".scn"), keep_directional=True)
    sc = from_parsed(parsed)
    seed = kw.get("seed", 0)
    okw = {k: v for k, v in kw.items() if k != "seed"}
    p32, p8, pst, _ = port.render(sc, O.Options(shade_triangles=True, **okw), rng_mode=O.RNG_PHILOX, seed=seed)
    gpu.upload(to_gpu(sc))
    g32, g8, gst = gpu.render(S.Options(shade_triangles=True, collect_stats=True, **kw))
    assert_image_parity(g32, p32, g8, p8, what=f"{scene} shaded {kw}")
    assert gst.closest_hit_rays == pst["closest_hit_rays"]
    # lit, not the reference's black silhouette: most triangle pixels carry colour
    off, _, _ = gpu.render(S.Options(**kw))
    tri_px = (off == 0).all(axis=2) & (g32 != 0).any(axis=2)
    assert tri_px.mean() > 0.01
    # the hierarchy prunes: far fewer leaf tests than brute force
    assert gst.tri_tests < 0.05 * pst["tri_tests"]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(5))
def test_random_scenes(gpu, port, seed):
    rng = np.random.default_rng(300 + seed)
    nt = int(rng.choice([1, 3, 40, 700]))
    sc = random_scene(rng, nspheres=int(rng.integers(0, 6)), nplights=int(rng.integers(1, 3)), ntris=nt, ndlights=int(rng.integers(0, 2)))
    if nt >= 700:
        t = sc.tris.reshape(-1, 3, 3)
        sc.tris = (t[:, :1] + 0.15 * (t - t[:, :1])).reshape(-1, 9)
    m = rng.uniform(0, 1, (nt, 14)).astype(np.float32)
    m[:, 6:9] *= rng.integers(0, 2, (nt, 1))
    m[:, 12] = rng.choice([1, 4, 16, 64], nt)
    sc.tri_materials = m
    sc = sc.normalised()
    gpu.upload(to_gpu(sc))
    for kw in [dict(width=160, height=100, use_shadows=True), dict(width=96, height=60, grid_size=2, seed=seed)]:
        s = kw.get("seed", 0)
        okw = {k: v for k, v in kw.items() if k != "seed"}
        p32, p8, _, _ = port.render(sc, O.Options(shade_triangles=True, **okw), rng_mode=O.RNG_PHILOX, seed=s)
        g32, g8, _ = gpu.render(S.Options(shade_triangles=True, **kw))
        assert_image_parity(g32, p32, g8, p8, min_ok=0.997, what=f"random shaded scene {seed} {kw}")
    # frame split: the mode renders this rank's tiles like every other
    full, _, _ = gpu.render(S.Options(width=96, height=60, use_shadows=True, shade_triangles=True))
    acc = np.zeros_like(full)
    for r in range(3):
        acc += gpu.render(S.Options(width=96, height=60, use_shadows=True, shade_triangles=True, rank=r, world=3, tile=16))[0]
    assert np.array_equal(acc, full)


@pytest.mark.gpu
def test_not_combined_with_the_wavefront_tree(gpu):
    gpu.upload(to_gpu(one_triangle_scene()))
    with pytest.raises(S.SkrError, match="shade_triangles"):
        gpu.render(S.Options(width=8, height=8, shade_triangles=True, monte_carlo=True, num_path_traces=2))


@pytest.mark.gpu
def test_cli_flag_renders_a_lit_dragon(tmp_path):
    d = O.REF_SCENES
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref/scenes not shipped")
    exe = os.path.join(ROOT, "host", "raytracer")
    a, b = tmp_path / "lit.ppm", tmp_path / "ref.ppm"
    common = ["--path", os.path.join(d, "dragon.scn"), "--width", "320", "--height", "180"]
    # (no --shadow: the reference uses `directional_light ... 1 -1 -1` as the direction TOWARDS the light, src/blinn_phong.h:77-84,
    # i.e. from below the ground quad, which would shadow everything)
    subprocess.run([exe, "--output", str(a), "--shade-triangles", "--keep-directional"] + common, check=True, capture_output=True, timeout=300)
    subprocess.run([exe, "--output", str(b)] + common, check=True, capture_output=True, timeout=300)
    lit = np.frombuffer(a.read_bytes()[len(b"P6\n320 180\n255\n"):], np.uint8).reshape(180, 320, 3)
    ref = np.frombuffer(b.read_bytes()[len(b"P6\n320 180\n255\n"):], np.uint8).reshape(180, 320, 3)
    black = (ref == 0).all(axis=2)                    # the reference's triangle pixels
    assert black.mean() > 0.05 and (lit[black] > 0).any(axis=1).mean() > 0.9
    assert len(np.unique(lit.reshape(-1, 3), axis=0)) > 50   # shading, not a 2-colour silhouette
