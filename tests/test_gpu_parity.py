"""Parity tests proper: the CUDA path, through the C ABI (libskr.so), against the oracle on the same inputs.

Oracle = oracle/skr_oracle.c (the C port), which tests/test_oracle_*.py pin bit-for-bit to the reference's golden
vector and to the reference's own compiled code.  Nothing here reads /root/reference.
"""
import dataclasses
import hashlib
import os
import subprocess

import numpy as np
import pytest

import skele_raytracer_b200 as S
from conftest import GOLDEN, ROOT, random_scene
from make_cases import GOLDEN_CASES
from oracle import oracle_lib as O
from parity import assert_image_parity, assert_mean_parity, frac_pixels_within
from test_oracle_ref import golden_testcpu

pytestmark = pytest.mark.gpu


def to_gpu_scene(s: O.Scene) -> S.Scene:
    return S.Scene(s.spheres, s.tris, s.plights, s.dlights, s.fogs, s.camera, s.ambient, s.background)


def opts(**kw):
    """-> (oracle Options, GPU Options) from one kwargs dict"""
    seed = kw.pop("seed", 0)
    extra = {k: kw.pop(k) for k in ("rank", "world", "tile", "collect_stats", "queue_capacity") if k in kw}
    return O.Options(**kw), S.Options(seed=seed, **extra, **kw)


@pytest.fixture(scope="module")
def gpu():
    r = S.Renderer()
    yield r
    r.close()


@pytest.fixture(scope="module")
def gscenes(scenes):
    return {k: to_gpu_scene(v) for k, v in scenes.items()}


# ---- golden vectors ----------------------------------------------------------------------------

def test_dragon_640x480_is_byte_identical_to_testcpu_ppm(gpu, gscenes):
    """The reference's only golden render (renders/testcpu.ppm == dragon.scn @ 640x480)."""
    img, sha = golden_testcpu()
    gpu.upload(gscenes["dragon"])
    _, rgb8, _ = gpu.render(S.Options(width=640, height=480, fov=60.0, max_depth=1), want_rgb32=False)
    assert hashlib.sha256(O.ppm_bytes(rgb8)).hexdigest() == sha


# (spheres2 is stochastic even without --gillum/--jsample: its fog record makes the shading draw random numbers)
@pytest.mark.parametrize("key", sorted(k for k, (sc, kw, _) in GOLDEN_CASES.items()
                                       if not kw.get("monte_carlo") and not kw.get("grid_size") and sc != "spheres2"))
def test_deterministic_golden_images_from_the_reference(gpu, gscenes, ref_images, key):
    scene, kw, _ = GOLDEN_CASES[key]
    gpu.upload(gscenes[scene])
    g32, _, _ = gpu.render(S.Options(**kw))
    assert_image_parity(g32, ref_images[key], what=key)


# ---- deterministic modes -----------------------------------------------------------------------

@pytest.mark.parametrize("scene", ["spheres1", "spheres2_nofog", "bear", "test", "dragon"])
@pytest.mark.parametrize("shadows", [False, True])
def test_deterministic_parity_small(gpu, port, scenes, gscenes, scene, shadows):
    oo, go = opts(width=480, height=270, max_depth=1 if not shadows else 3, use_shadows=shadows, collect_stats=True)
    p32, p8, pst, _ = port.render(scenes[scene], oo)
    gpu.upload(gscenes[scene])
    g32, g8, gst = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, what=f"{scene} shadows={shadows}")
    assert gst.closest_hit_rays == pst["closest_hit_rays"]
    assert abs(int(gst.shadow_rays) - pst["shadow_rays"]) <= 1e-4 * max(1, pst["shadow_rays"])


@pytest.mark.parametrize("scene,kw", [("spheres1", dict(max_depth=1)),                       # BASELINE config 1
                                      ("dragon", dict(use_shadows=True)),                    # BASELINE config 4
                                      ("bear", dict(use_shadows=True)), ("test", dict(use_shadows=True, fov=45.0))])
def test_deterministic_parity_1920x1080(gpu, port, scenes, gscenes, scene, kw):
    oo, go = opts(width=1920, height=1080, **kw)
    p32, p8, _, _ = port.render(scenes[scene], oo)
    gpu.upload(gscenes[scene])
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, what=f"{scene} 1080p")


# ---- stochastic modes, same keyed Philox stream on both sides ----------------------------------

STOCH = [("spheres2", dict(width=240, height=135, grid_size=5, use_shadows=True, seed=1)),               # config 2 shape
         ("spheres2", dict(width=96, height=54, max_depth=4, monte_carlo=True, num_path_traces=16, seed=2)),  # config 3 shape
         ("bear", dict(width=64, height=36, monte_carlo=True, num_path_traces=64, grid_size=4, use_shadows=True, seed=3)),  # config 5 shape
         ("spheres1", dict(width=128, height=72, max_depth=3, monte_carlo=True, num_path_traces=5, use_shadows=True, seed=4)),
         ("test", dict(width=96, height=54, max_depth=2, monte_carlo=True, num_path_traces=3, grid_size=2, seed=5)),
         ("spheres2_nofog", dict(width=128, height=72, grid_size=1, seed=6))]


@pytest.mark.parametrize("scene,kw", STOCH)
def test_stochastic_same_stream_parity(gpu, port, scenes, gscenes, scene, kw):
    oo, go = opts(collect_stats=True, **dict(kw))
    p32, p8, pst, _ = port.render(scenes[scene], oo, rng_mode=O.RNG_PHILOX, seed=go.seed)
    gpu.upload(gscenes[scene])
    g32, g8, gst = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, what=f"{scene} {kw}")
    rays_p = pst["closest_hit_rays"] + pst["shadow_rays"]
    rays_g = int(gst.closest_hit_rays + gst.shadow_rays)
    assert abs(rays_g - rays_p) <= 2e-4 * rays_p


def test_config2_full_size_same_stream(gpu, port, scenes, gscenes):
    """BASELINE config 2 at its real size: spheres2 1920x1080 --jsample 5 --shadow."""
    oo, go = opts(width=1920, height=1080, grid_size=5, use_shadows=True, seed=21)
    p32, p8, _, _ = port.render(scenes["spheres2"], oo, rng_mode=O.RNG_PHILOX, seed=21)
    gpu.upload(gscenes["spheres2"])
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, what="config 2")


@pytest.mark.parametrize("scene,kw,rows", [
    ("spheres2", dict(width=1920, height=1080, max_depth=4, monte_carlo=True, num_path_traces=16, seed=31), (500, 540)),          # config 3
    ("bear", dict(width=3840, height=2160, monte_carlo=True, num_path_traces=64, grid_size=4, use_shadows=True, seed=32), (1000, 1004))])  # config 5
def test_full_size_gi_row_window(gpu, port, scenes, gscenes, scene, kw, rows):
    """Configs 3 and 5 at full size: the oracle renders a window of rows of the SAME frame (it would need hours for all
    of it); the GPU renders the whole frame."""
    oo, go = opts(**dict(kw))
    y0, y1 = rows
    p32, _, _, _ = port.render(scenes[scene], oo, rng_mode=O.RNG_PHILOX, seed=go.seed, y0=y0, y1=y1, want_rgb8=False)
    gpu.upload(gscenes[scene])
    g32, _, _ = gpu.render(go, want_rgb8=False)
    assert_image_parity(g32[y0:y1], p32[y0:y1], what=f"{scene} rows {rows}")


# ---- stochastic modes vs the reference's own rand() stream: per-pixel mean, 3 sigma, N seeds ---

@pytest.mark.parametrize("scene,kw,N", [
    ("spheres2", dict(width=64, height=36, grid_size=5, use_shadows=True), 8),
    ("spheres2", dict(width=48, height=27, max_depth=4, monte_carlo=True, num_path_traces=16), 16),
    ("bear", dict(width=32, height=18, monte_carlo=True, num_path_traces=16, grid_size=2, use_shadows=True), 16)])
def test_stochastic_mean_parity_vs_reference_rand_stream(gpu, port, scenes, gscenes, scene, kw, N):
    oo, _ = opts(**dict(kw))
    gpu.upload(gscenes[scene])
    A = np.stack([gpu.render(S.Options(seed=i, **kw), want_rgb8=False)[0] for i in range(N)])
    B = np.stack([port.render(scenes[scene], oo, rng_mode=O.RNG_LIBC, seed=1000 + i)[0] for i in range(N)])
    assert_mean_parity(A, B, what=f"{scene} {kw}")


# ---- opt-in fresnel mode (row A9: the recursion of src/raytrace.h:46-103 that HEAD skips) -------------------------

FRESNEL = [("spheres1", dict(width=320, height=180, max_depth=2, fresnel=True)),
           ("spheres1", dict(width=320, height=180, max_depth=3, use_shadows=True, fresnel=True)),
           ("spheres2_nofog", dict(width=320, height=180, max_depth=3, use_shadows=True, fresnel=True)),
           ("bear", dict(width=320, height=180, max_depth=4, fresnel=True)),
           ("test", dict(width=160, height=90, max_depth=3, fresnel=True)),
           ("spheres2", dict(width=160, height=90, max_depth=3, grid_size=2, use_shadows=True, fresnel=True, seed=8)),
           ("spheres2_nofog", dict(width=96, height=54, max_depth=3, monte_carlo=True, num_path_traces=3, fresnel=True, seed=9))]


@pytest.mark.parametrize("scene,kw", FRESNEL)
def test_fresnel_mode_parity(gpu, port, scenes, gscenes, scene, kw):
    """Oracle: the port with fresnel=1, which tests/test_oracle_port.py pins bit-for-bit to the reference compiled with
    src/raytrace.h:44 removed."""
    oo, go = opts(collect_stats=True, **dict(kw))
    p32, p8, pst, _ = port.render(scenes[scene], oo, rng_mode=O.RNG_PHILOX, seed=go.seed)
    gpu.upload(gscenes[scene])
    g32, g8, gst = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, min_ok=0.998, what=f"fresnel {scene} {kw}")
    # the GPU traces only the refraction child whose result the reference keeps (the last light's), so it may issue
    # fewer closest-hit rays than the port, never more
    assert gst.closest_hit_rays <= pst["closest_hit_rays"]
    assert gst.closest_hit_rays >= 0.5 * pst["closest_hit_rays"]


def test_fresnel_off_is_head_behaviour(gpu, gscenes):
    gpu.upload(gscenes["spheres1"])
    a, _, _ = gpu.render(S.Options(width=160, height=90, max_depth=3))
    b, _, _ = gpu.render(S.Options(width=160, height=90, max_depth=1))
    c, _, _ = gpu.render(S.Options(width=160, height=90, max_depth=1, fresnel=True))  # depth 1: children would have depth 0
    assert np.array_equal(a, b) and np.allclose(a, c, atol=1e-6)


# ---- edge cases ---------------------------------------------------------------------------------

def test_empty_scene_is_background(gpu):
    s = S.Scene(background=np.array([0.25, 0.5, 0.75], np.float32), camera=np.array([0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 0], np.float32))
    gpu.upload(s)
    g32, g8, _ = gpu.render(S.Options(width=37, height=23))
    assert np.allclose(g32, [0.25, 0.5, 0.75]) and (g8 == [63, 127, 191]).all()


def test_depth_zero_is_black(gpu, gscenes):
    gpu.upload(gscenes["bear"])
    g32, g8, _ = gpu.render(S.Options(width=64, height=36, max_depth=0))
    assert not g32.any() and not g8.any()
    g32, _, _ = gpu.render(S.Options(width=64, height=36, max_depth=0, monte_carlo=True, num_path_traces=4))
    assert not g32.any()


def test_gillum_zero_paths_is_nan_like_the_reference(gpu, port, scenes, gscenes):
    """`--gillum 0`: the reference's empty Monte-Carlo sum is divided by 0 -> NaN on every sphere hit -> byte 255."""
    oo, go = opts(width=64, height=36, max_depth=2, monte_carlo=True, num_path_traces=0)
    p32, p8, _, _ = port.render(scenes["spheres1"], oo, rng_mode=O.RNG_PHILOX, seed=0)
    gpu.upload(gscenes["spheres1"])
    g32, g8, _ = gpu.render(go)
    assert np.isnan(p32).any() and np.array_equal(np.isnan(g32), np.isnan(p32))
    assert np.array_equal(g8, p8) and (g8[np.isnan(g32)] == 255).all()


@pytest.mark.parametrize("w,h", [(1, 1), (37, 23), (33, 65), (8, 4), (257, 3)])
def test_ragged_sizes(gpu, port, scenes, gscenes, w, h):
    oo, go = opts(width=w, height=h, use_shadows=True)
    p32, p8, _, _ = port.render(scenes["spheres1"], oo)
    gpu.upload(gscenes["spheres1"])
    g32, g8, _ = gpu.render(go)
    assert frac_pixels_within(g32, p32) >= 1.0 - 2.0 / (w * h) - 0.001


@pytest.mark.parametrize("seed", range(6))
def test_random_scenes(gpu, port, seed):
    """Spheres + triangles in front of / behind each other, directional lights, several fogs, odd --gillum counts."""
    rng = np.random.default_rng(seed)
    sc = random_scene(rng, nspheres=int(rng.integers(1, 12)), nplights=int(rng.integers(1, 4)), ntris=int(rng.integers(0, 60)),
                      nfogs=int(rng.integers(0, 3)), ndlights=int(rng.integers(0, 3)))
    gpu.upload(to_gpu_scene(sc))
    for kw in [dict(width=160, height=100, use_shadows=True), dict(width=96, height=60, grid_size=2, use_shadows=bool(seed & 1), seed=seed),
               dict(width=64, height=40, max_depth=3, monte_carlo=True, num_path_traces=int(rng.integers(1, 7)), use_shadows=True, seed=seed)]:
        oo, go = opts(**dict(kw))
        p32, p8, _, _ = port.render(sc, oo, rng_mode=O.RNG_PHILOX, seed=go.seed)
        g32, g8, _ = gpu.render(go)
        assert_image_parity(g32, p32, g8, p8, min_ok=0.998, what=f"random scene {seed} {kw}")


def test_many_spheres_use_the_global_memory_path(gpu, port):
    """More spheres than the shared-memory staging holds (blob > 64 KB)."""
    rng = np.random.default_rng(5)
    sc = random_scene(rng, nspheres=1200, nplights=2)
    sc.spheres[:, 3] *= 0.2
    oo, go = opts(width=96, height=54, use_shadows=True)
    p32, p8, _, _ = port.render(sc, oo)
    gpu.upload(to_gpu_scene(sc))
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, min_ok=0.998, what="1200 spheres")
    # the wavefront kernels on the same path (shade_expand / fresnel_expand without shared-memory staging)
    oo, go = opts(width=48, height=27, max_depth=2, monte_carlo=True, num_path_traces=2, use_shadows=True, fresnel=True, seed=3)
    p32, p8, _, _ = port.render(sc, oo, rng_mode=O.RNG_PHILOX, seed=3)
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, min_ok=0.995, what="1200 spheres, gillum + fresnel")


@pytest.mark.parametrize("seed", range(4))
def test_random_scenes_everything_on(gpu, port, seed):
    """fresnel + --gillum + --jsample + shadows + fog + directional lights + triangles at once."""
    rng = np.random.default_rng(100 + seed)
    sc = random_scene(rng, nspheres=int(rng.integers(2, 9)), nplights=int(rng.integers(1, 3)), ntris=int(rng.integers(1, 30)),
                      nfogs=int(rng.integers(0, 2)), ndlights=int(rng.integers(0, 2)))
    gpu.upload(to_gpu_scene(sc))
    kw = dict(width=72, height=48, max_depth=3, monte_carlo=True, num_path_traces=int(rng.choice([2, 4, 7])), grid_size=int(rng.choice([0, 2, 3])),
              use_shadows=True, fresnel=True, seed=seed)
    oo, go = opts(**dict(kw))
    p32, p8, _, _ = port.render(sc, oo, rng_mode=O.RNG_PHILOX, seed=go.seed)
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, min_ok=0.997, what=f"everything-on scene {seed} {kw}")


def test_gillum_64_single_sample(gpu, port, scenes, gscenes):
    oo, go = opts(width=48, height=27, max_depth=3, monte_carlo=True, num_path_traces=64, use_shadows=True, seed=12)
    p32, p8, _, _ = port.render(scenes["spheres1"], oo, rng_mode=O.RNG_PHILOX, seed=12)
    gpu.upload(gscenes["spheres1"])
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, what="gillum 64")


# ---- BVH ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("ntris", [1, 2, 3, 33, 1000, 11264, 11265, 20000, 50000])   # (sort chunk sizes: 256, 288, 480, 1024 keys per warp)
def test_bvh_equals_brute_force(gpu, ntris):
    rng = np.random.default_rng(ntris)
    sc = random_scene(rng, nspheres=3, nplights=1, ntris=ntris)
    if ntris >= 1000:
        sc.tris = (sc.tris.reshape(-1, 3, 3)[:, :1] + 0.08 * (sc.tris.reshape(-1, 3, 3) - sc.tris.reshape(-1, 3, 3)[:, :1])).reshape(-1, 9)
    if ntris == 33:
        sc.tris[5:20] = sc.tris[5]  # duplicates -> identical Morton codes
    if ntris == 20000:
        sc.tris[1000:9000] = sc.tris[1000]  # 8000 identical triangles: ties broken by index, depth stays logarithmic
    g = to_gpu_scene(sc)
    go = S.Options(width=200, height=120, use_shadows=True)
    gpu.upload(g)
    a32, _, _ = gpu.render(go)
    os.environ["SKR_NO_BVH"] = "1"
    try:
        gpu.upload(g)
        b32, _, _ = gpu.render(go)
    finally:
        del os.environ["SKR_NO_BVH"]
    assert np.array_equal(a32, b32)


@pytest.mark.parametrize("scene,kw", [("dragon", dict(width=1920, height=1080)), ("dragon", dict(width=333, height=187, fov=25.0)),
                                      ("test", dict(width=640, height=360, use_shadows=True)), ("test", dict(width=200, height=120, rank=1, world=3, tile=16))])
def test_deferred_triangle_query_equals_the_query_in_place(gpu, gscenes, scene, kw):
    """With SKR_DEFER=1 single-sample frames hand the camera rays that are not settled within a few nodes to
    tri_deferred_kernel (a dense list walked by teams of 8 lanes per ray from a shared stack); the frame must be the one the
    in-place query (the default) gives, bit for bit."""
    gpu.upload(gscenes[scene])
    o = S.Options(collect_stats=True, **kw)
    b32, b8, sb = gpu.render(o)
    os.environ["SKR_DEFER"] = "1"
    try:
        a32, a8, sa = gpu.render(o)
    finally:
        del os.environ["SKR_DEFER"]
    assert np.array_equal(a32.view(np.uint32), b32.view(np.uint32)) and np.array_equal(a8, b8)
    # (node / leaf-test counts differ: a team visits the hierarchy in another order and finishes the iteration a hit falls in)
    assert sa.kernel_launches == sb.kernel_launches + 1 and sa.closest_hit_rays == sb.closest_hit_rays


def test_bvh_dragon_1080p_equals_brute_force_window(gpu, port, scenes, gscenes):
    oo, go = opts(width=1920, height=1080)
    gpu.upload(gscenes["dragon"])
    g32, _, st = gpu.render(dataclasses.replace(go, collect_stats=True))
    p32, _, _, _ = port.render(scenes["dragon"], oo, y0=500, y1=560)
    assert np.array_equal(g32[500:560], p32[500:560])
    # the hierarchy must actually prune: far fewer leaf tests than the reference's T per ray
    assert st.tri_tests < 0.02 * 1920 * 1080 * 10002


# ---- size-independent properties at full size ----------------------------------------------------

def test_linearity_in_light_colour_1080p(gpu, scenes):
    """Blinn-Phong is linear in the light colours: frame(2L) - frame(0) == 2 * (frame(L) - frame(0)) on the pre-clamp float
    image, at BASELINE size, shadows on (shadowing does not depend on colour)."""
    import copy
    base = scenes["bear"]
    imgs = []
    for k in (0.0, 1.0, 2.0):
        sc = copy.deepcopy(base)
        sc.plights = sc.plights.copy()
        sc.plights[:, 3:6] *= k
        gpu.upload(to_gpu_scene(sc))
        imgs.append(gpu.render(S.Options(width=1920, height=1080, use_shadows=True), want_rgb8=False)[0].astype(np.float64))
    d1, d2 = imgs[1] - imgs[0], imgs[2] - imgs[0]
    assert np.abs(d2 - 2 * d1).max() <= 1e-5 * max(1.0, np.abs(d2).max())
    assert np.abs(d1).max() > 0.1


def test_background_only_reaches_miss_pixels_1080p(gpu, scenes):
    import copy
    a = copy.deepcopy(scenes["spheres1"])
    b = copy.deepcopy(scenes["spheres1"])
    b.background = np.array([0.9, 0.1, 0.4], np.float32)
    gpu.upload(to_gpu_scene(a))
    ia = gpu.render(S.Options(width=1920, height=1080, use_shadows=True), want_rgb8=False)[0]
    gpu.upload(to_gpu_scene(b))
    ib = gpu.render(S.Options(width=1920, height=1080, use_shadows=True), want_rgb8=False)[0]
    changed = (ia != ib).any(axis=2)
    assert np.allclose(ia[changed], a.background) and np.allclose(ib[changed], b.background)
    assert 0.3 < changed.mean() < 0.9   # spheres1 covers ~40 % of the frame


def test_gi_energy_bookkeeping_matches_ray_counts(gpu, gscenes):
    """Every queue entry is a sphere hit that gets shaded exactly once: queue entries == sphere hits, and closest-hit rays
    == primary samples + n * (hits that the reference would expand)."""
    gpu.upload(gscenes["bear"])
    w, h, n, depth = 320, 180, 8, 3
    st = gpu.render(S.Options(width=w, height=h, max_depth=depth, monte_carlo=True, num_path_traces=n, collect_stats=True, seed=4, use_shadows=True),
                    want_rgb8=False, want_rgb32=False)[2]
    # depth-1 hits (the leaves) are shaded in place by the warp that found them (frames with shadow rays), never queued
    assert 0 < st.queue_entries < st.sphere_hits
    os.environ["SKR_NO_LEAF_INLINE"] = "1"
    try:
        st2 = gpu.render(S.Options(width=w, height=h, max_depth=depth, monte_carlo=True, num_path_traces=n, collect_stats=True, seed=4, use_shadows=True),
                         want_rgb8=False, want_rgb32=False)[2]
    finally:
        del os.environ["SKR_NO_LEAF_INLINE"]
    assert st2.queue_entries == st2.sphere_hits == st.sphere_hits
    # hits at depth >= 2 expand into n children each; depth-1 hits do not.  With depth 3: levels 0 and 1 expand.
    assert (st.closest_hit_rays - w * h) % n == 0


# ---- determinism, chunking, frame split ---------------------------------------------------------

GI_KW = dict(width=160, height=96, max_depth=3, monte_carlo=True, num_path_traces=6, grid_size=2, use_shadows=True, seed=77)


def test_run_to_run_bit_identical(gpu, gscenes):
    gpu.upload(gscenes["spheres2"])
    a, _, _ = gpu.render(S.Options(**GI_KW))
    b, _, _ = gpu.render(S.Options(**GI_KW))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


CULL_CASES = [("spheres2", dict(width=480, height=270, grid_size=5, use_shadows=True, seed=11)),                  # config 2 shape
              ("spheres2", dict(width=317, height=201, grid_size=2, use_shadows=True, fov=120.0, seed=12)),      # wide pixels
              ("spheres2_nofog", dict(width=64, height=36, grid_size=3, use_shadows=True, fov=170.0, seed=13)),  # huge pixels
              ("bear", dict(width=480, height=270, grid_size=4, use_shadows=True, seed=14)),
              ("spheres1", dict(width=320, height=180, grid_size=3, use_shadows=True, seed=15)),
              ("test", dict(width=160, height=90, grid_size=2, use_shadows=True, seed=16)),
              ("bear", dict(width=96, height=54, grid_size=2, monte_carlo=True, num_path_traces=4, use_shadows=True, seed=17)),
              ("bear", dict(width=480, height=270, use_shadows=True)),                                            # static masks only
              ("spheres2_nofog", dict(width=480, height=270, use_shadows=True)),
              ("spheres1", dict(width=96, height=54, max_depth=3, monte_carlo=True, num_path_traces=6, use_shadows=True, seed=18)),
              ("spheres2", dict(width=64, height=36, max_depth=3, monte_carlo=True, num_path_traces=5, use_shadows=True, fresnel=True, seed=19))]


@pytest.mark.parametrize("scene,kw", CULL_CASES)
def test_bundle_culling_is_exact(gpu, gscenes, scene, kw):
    """cull_pairs only ever skips spheres whose exact test fails: frame AND device counters (sphere tests with a
    non-negative discriminant included) are identical with culling off (SKR_NO_CULL=1)."""
    gpu.upload(gscenes[scene])
    o = S.Options(collect_stats=True, **kw)
    a32, a8, sa = gpu.render(o)
    os.environ["SKR_NO_CULL"] = "1"          # read at upload (static per-receiver masks) and per frame (pixel bundles)
    try:
        gpu.upload(gscenes[scene])
        b32, b8, sb = gpu.render(o)
    finally:
        del os.environ["SKR_NO_CULL"]
    assert np.array_equal(a32.view(np.uint32), b32.view(np.uint32)) and np.array_equal(a8, b8)
    for f in ("closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos", "sphere_hits", "light_evals"):
        assert getattr(sa, f) == getattr(sb, f), f


@pytest.mark.parametrize("seed", range(6))
def test_bundle_culling_is_exact_random_scenes(gpu, seed):
    rng = np.random.default_rng(900 + seed)
    sc = to_gpu_scene(random_scene(rng, nspheres=int(rng.integers(1, 60)), nplights=int(rng.integers(1, 4))))
    gpu.upload(sc)
    o = S.Options(width=200, height=120, grid_size=int(rng.integers(0, 5)), use_shadows=True, fov=float(rng.uniform(20, 150)),
                  monte_carlo=bool(seed & 1), num_path_traces=3, max_depth=2, seed=seed, collect_stats=True)
    a32, a8, sa = gpu.render(o)
    os.environ["SKR_NO_CULL"] = "1"
    try:
        gpu.upload(sc)
        b32, b8, sb = gpu.render(o)
    finally:
        del os.environ["SKR_NO_CULL"]
    assert np.array_equal(a32.view(np.uint32), b32.view(np.uint32)) and np.array_equal(a8, b8)
    for f in ("closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos", "sphere_hits", "light_evals"):
        assert getattr(sa, f) == getattr(sb, f), f


@pytest.mark.parametrize("scene,kw", [("spheres2", dict(width=1920, height=1080, grid_size=2, use_shadows=True, seed=5)),
                                      ("dragon", dict(width=1000, height=700, grid_size=2, seed=6)),  # ragged tile columns / rows
                                      ("bear", dict(width=1280, height=720, use_shadows=True, grid_size=3, seed=7))])
def test_overlapped_copy_out_into_pinned_host_memory(gpu, gscenes, scene, kw):
    """skr_render with page-locked destinations copies the frame out in bands while the kernel is still running
    (stream-ordered waits on per-band flags); the bytes are those of the plain copy-after-kernel path."""
    import torch
    gpu.upload(gscenes[scene])
    o = S.Options(**kw)
    ref32, ref8, _ = gpu.render(o)                                   # pageable numpy buffers: copy after the kernel
    for trial in range(3):
        h8 = torch.full((o.height, o.width, 3), 77, dtype=torch.uint8).pin_memory()
        h32 = torch.full((o.height, o.width, 3), -1.0, dtype=torch.float32).pin_memory()
        gpu.render(o, rgb8=h8.numpy(), rgb32=h32.numpy())
        assert np.array_equal(h8.numpy(), ref8) and np.array_equal(h32.numpy().view(np.uint32), ref32.view(np.uint32))
        h8b = torch.zeros((o.height, o.width, 3), dtype=torch.uint8).pin_memory()
        gpu.render(o, rgb8=h8b.numpy(), want_rgb32=False)
        assert np.array_equal(h8b.numpy(), ref8)
    # RGB8 alone into page-locked memory is stored by the kernel itself (no copy); SKR_NO_HOST_STORES=1 sends it through the
    # band copies instead, SKR_NO_OVERLAP=1 through a plain copy after the kernel: the same bytes every way
    for env in ("SKR_NO_HOST_STORES", "SKR_NO_OVERLAP"):
        os.environ[env] = "1"
        try:
            h8 = torch.zeros((o.height, o.width, 3), dtype=torch.uint8).pin_memory()
            gpu.render(o, rgb8=h8.numpy(), want_rgb32=False)
        finally:
            del os.environ[env]
        assert np.array_equal(h8.numpy(), ref8), env
    for nb in ("1", "3", "8"):                      # band count (tuning aid): bands are launched heaviest first
        os.environ["SKR_BANDS"] = nb
        try:
            h8 = torch.zeros((o.height, o.width, 3), dtype=torch.uint8).pin_memory()
            h32 = torch.zeros((o.height, o.width, 3), dtype=torch.float32).pin_memory()
            gpu.render(o, rgb8=h8.numpy(), rgb32=h32.numpy())
        finally:
            del os.environ["SKR_BANDS"]
        assert np.array_equal(h8.numpy(), ref8) and np.array_equal(h32.numpy().view(np.uint32), ref32.view(np.uint32)), nb


@pytest.mark.parametrize("w,h", [(1000, 700), (1920, 1080), (644, 600), (100, 60), (36, 8)])
def test_word_and_strip_stores_stay_inside_the_frame(gpu, gscenes, w, h):
    """Frames bound for other devices / page-locked host memory leave as 32-bit words, whole 32 x 4 strips per CTA where the
    strip lies inside the image: canaries right before and after the frame must survive, ragged right and bottom edges included,
    and the frame must be the one the plain path renders."""
    import torch
    gpu.upload(gscenes["spheres2"])
    o = S.Options(width=w, height=h, grid_size=2, use_shadows=True, seed=3)
    _, ref8, _ = gpu.render(o, want_rgb32=False)
    n, pad = w * h * 3, 4096
    # (a) device frame through skr_render_peers_device
    buf = torch.full((n + 2 * pad,), 0xAB, dtype=torch.uint8, device="cuda")
    gpu.render_peers_device(o, [buf.data_ptr() + pad])
    gpu.sync()
    got = buf.cpu().numpy()
    assert (got[:pad] == 0xAB).all() and (got[pad + n:] == 0xAB).all()
    assert np.array_equal(got[pad:pad + n].reshape(h, w, 3), ref8)
    # (CTAs of other sizes -- SKR_PRIMARY_BLOCK, a tuning aid -- cannot pair up into strips and must fall back to blocks)
    for threads in ("64", "32"):
        os.environ["SKR_PRIMARY_BLOCK"] = threads
        try:
            buf.fill_(0xAB)
            gpu.render_peers_device(o, [buf.data_ptr() + pad])
            gpu.sync()
        finally:
            del os.environ["SKR_PRIMARY_BLOCK"]
        got = buf.cpu().numpy()
        assert (got[:pad] == 0xAB).all() and (got[pad + n:] == 0xAB).all()
        assert np.array_equal(got[pad:pad + n].reshape(h, w, 3), ref8), threads
    # (b) page-locked host frame through skr_render (stored by the kernel itself when the frame is large enough)
    hbuf = torch.full((n + 2 * pad,), 0xCD, dtype=torch.uint8).pin_memory()
    gpu.render(o, rgb8=hbuf.numpy()[pad:pad + n].reshape(h, w, 3), want_rgb32=False)
    got = hbuf.numpy()
    assert (got[:pad] == 0xCD).all() and (got[pad + n:] == 0xCD).all()
    assert np.array_equal(got[pad:pad + n].reshape(h, w, 3), ref8)


def test_queue_capacity_does_not_change_the_image(gpu, gscenes):
    gpu.upload(gscenes["spheres2"])
    a, _, sa = gpu.render(S.Options(**GI_KW))
    b, _, sb = gpu.render(S.Options(queue_capacity=6 * 256, **GI_KW))
    assert sb.queue_chunks > sa.queue_chunks
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("world,tile", [(2, 32), (3, 16), (8, 32), (5, 24), (7, 40)])   # (divisors that are not powers of two: fdiv)
def test_frame_split_is_bit_identical_for_any_world(gpu, gscenes, world, tile):
    gpu.upload(gscenes["spheres2"])
    full32, full8, _ = gpu.render(S.Options(**GI_KW))
    acc32 = np.zeros_like(full32)
    acc8 = np.zeros_like(full8)
    from skele_raytracer_b200 import tiles as T
    own = T.owner_map(GI_KW["width"], GI_KW["height"], world, tile)
    for r in range(world):
        p32, p8, _ = gpu.render(S.Options(rank=r, world=world, tile=tile, **GI_KW))
        assert not p32[own != r].any()          # other ranks' pixels are left zero
        acc32 += p32
        acc8 += p8
    assert np.array_equal(acc32.view(np.uint32), full32.view(np.uint32)) and np.array_equal(acc8, full8)


@pytest.mark.parametrize("world,tile", [(1, 32), (2, 32), (4, 16), (3, 24)])
def test_compact_tiles_and_deinterleave_on_device(gpu, gscenes, world, tile):
    import torch
    from skele_raytracer_b200 import tiles as T
    kw = dict(width=150, height=70, use_shadows=True)
    gpu.upload(gscenes["spheres1"])
    _, full8, _ = gpu.render(S.Options(**kw), want_rgb32=False)
    parts = []
    for r in range(world):
        o = S.Options(rank=r, world=world, tile=tile, **kw)
        buf = torch.zeros(gpu.tiles_bytes(o), dtype=torch.uint8, device="cuda")
        assert buf.numel() == T.tiles_bytes(kw["width"], kw["height"], world, tile)
        gpu.render_tiles_device(o, buf.data_ptr())
        parts.append(buf)
        assert np.array_equal(buf.cpu().numpy() * (T.compact_from_frame(np.ones_like(full8), r, world, tile) > 0),
                              T.compact_from_frame(full8, r, world, tile))
    gathered = torch.cat(parts)
    frame = torch.zeros((kw["height"], kw["width"], 3), dtype=torch.uint8, device="cuda")
    o = S.Options(rank=0, world=world, tile=tile, **kw)
    gpu.deinterleave_device(o, gathered.data_ptr(), frame.data_ptr())
    gpu.sync()
    assert np.array_equal(frame.cpu().numpy(), full8)


def test_render_peers_device_fills_every_frame(gpu, gscenes):
    """skr_render_peers_device: each rank stores its pixels into every frame it is given (here: three local buffers
    standing in for the GPUs of a box); after all ranks have rendered, every frame is the whole image."""
    import torch
    kw = dict(width=200, height=90, grid_size=2, use_shadows=True, seed=2)
    gpu.upload(gscenes["spheres2"])
    _, full8, _ = gpu.render(S.Options(**kw), want_rgb32=False)
    frames = [torch.zeros((90, 200, 3), dtype=torch.uint8, device="cuda") for _ in range(3)]
    for r in range(4):
        gpu.render_peers_device(S.Options(rank=r, world=4, tile=16, **kw), [f.data_ptr() for f in frames])
    gpu.sync()
    for f in frames:
        assert np.array_equal(f.cpu().numpy(), full8)
    with pytest.raises(S.SkrError, match="between 1 and 8"):
        gpu.render_peers_device(S.Options(**kw), [frames[0].data_ptr()] * 9)
    # --gillum frames go through resolve_kernel -> same stores
    kw = dict(width=96, height=54, max_depth=2, monte_carlo=True, num_path_traces=3, seed=2)
    _, full8, _ = gpu.render(S.Options(**kw), want_rgb32=False)
    frames = [torch.zeros((54, 96, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for r in range(2):
        gpu.render_peers_device(S.Options(rank=r, world=2, **kw), [f.data_ptr() for f in frames])
    gpu.sync()
    assert all(np.array_equal(f.cpu().numpy(), full8) for f in frames)


def test_render_device_matches_host_render(gpu, gscenes):
    import torch
    gpu.upload(gscenes["bear"])
    o = S.Options(width=320, height=180, use_shadows=True)
    h32, h8, _ = gpu.render(o)
    d8 = torch.zeros((180, 320, 3), dtype=torch.uint8, device="cuda")
    d32 = torch.zeros((180, 320, 3), dtype=torch.float32, device="cuda")
    gpu.render_device(o, d8.data_ptr(), d32.data_ptr())
    assert np.array_equal(d8.cpu().numpy(), h8) and np.array_equal(d32.cpu().numpy(), h32)



# ---- leaves shaded in place == leaves through the queue, bit for bit -------------------------------

LEAF_CASES = [("spheres2", dict(width=160, height=90, max_depth=4, monte_carlo=True, num_path_traces=16, seed=41)),             # config 3 shape (fog)
              ("bear", dict(width=96, height=54, monte_carlo=True, num_path_traces=64, grid_size=2, use_shadows=True, seed=42)),  # config 5 shape
              ("spheres1", dict(width=128, height=72, max_depth=2, monte_carlo=True, num_path_traces=7, use_shadows=True, seed=43)),   # odd n: remainder loop
              ("test", dict(width=96, height=54, max_depth=3, monte_carlo=True, num_path_traces=5, grid_size=2, seed=44)),        # triangles
              ("spheres2_nofog", dict(width=64, height=36, max_depth=5, monte_carlo=True, num_path_traces=3, use_shadows=True, seed=45)),
              ("bear", dict(width=33, height=17, max_depth=2, monte_carlo=True, num_path_traces=1, seed=46))]


@pytest.mark.parametrize("scene,kw", LEAF_CASES)
def test_leaves_in_place_equal_leaves_through_the_queue(gpu, gscenes, scene, kw):
    """shade_expand_kernel<LEAF>: the depth-1 hits are compacted in shared memory and shaded by the warp that found them;
    frame (float bit patterns) and every device counter must equal the all-queued wavefront (SKR_NO_LEAF_INLINE=1)."""
    gpu.upload(gscenes[scene])
    o = S.Options(collect_stats=True, **kw)
    os.environ["SKR_LEAF_INLINE"] = "1"      # (by default only frames with shadow rays take the in-place path)
    try:
        a32, a8, sa = gpu.render(o)
    finally:
        del os.environ["SKR_LEAF_INLINE"]
    os.environ["SKR_NO_LEAF_INLINE"] = "1"
    try:
        b32, b8, sb = gpu.render(o)
    finally:
        del os.environ["SKR_NO_LEAF_INLINE"]
    assert np.array_equal(a32.view(np.uint32), b32.view(np.uint32)) and np.array_equal(a8, b8)
    for f in ("closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos", "sphere_hits", "light_evals", "tri_tests"):
        assert getattr(sa, f) == getattr(sb, f), f
    assert sa.queue_entries < sb.queue_entries and sa.kernel_launches <= sb.kernel_launches


def test_many_spheres_leaves_in_place_on_the_global_memory_path(gpu, port):
    rng = np.random.default_rng(6)
    sc = random_scene(rng, nspheres=1200, nplights=2)
    sc.spheres[:, 3] *= 0.2
    oo, go = opts(width=48, height=27, max_depth=2, monte_carlo=True, num_path_traces=4, use_shadows=True, seed=3)
    p32, p8, _, _ = port.render(sc, oo, rng_mode=O.RNG_PHILOX, seed=3)
    gpu.upload(to_gpu_scene(sc))
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, min_ok=0.995, what="1200 spheres, gillum, leaves in place")


# ---- hardening ------------------------------------------------------------------------------------

def test_gillum_zero_with_jsample_is_nan_on_every_hit_pixel(gpu, port, scenes, gscenes):
    """`--gillum 0 --jsample 3`: nine NaN contributions per hit pixel.  Non-finite contributions are kept out of band
    (a flags word beside the fixed-point sums), so any number of them resolves like the reference's float sum."""
    oo, go = opts(width=64, height=36, max_depth=2, monte_carlo=True, num_path_traces=0, grid_size=3, seed=7)
    p32, p8, _, _ = port.render(scenes["spheres1"], oo, rng_mode=O.RNG_PHILOX, seed=7)
    gpu.upload(gscenes["spheres1"])
    g32, g8, _ = gpu.render(go)
    assert np.isnan(p32).any() and np.array_equal(np.isnan(g32), np.isnan(p32))
    assert np.array_equal(g8, p8) and (g8[np.isnan(g32)] == 255).all()


def test_infinite_radiance_saturates_instead_of_wrapping(gpu, gscenes):
    """Light colours of 1e30: every lit contribution overflows fixed point many times over; the pixel must come out
    +inf / 255, never a wrapped value."""
    sc = dataclasses.replace(gscenes["spheres1"])
    sc.plights = sc.plights.copy()
    sc.plights[:, 3:6] = 1.0e30
    gpu.upload(sc)
    g32, g8, _ = gpu.render(S.Options(width=64, height=36, max_depth=3, monte_carlo=True, num_path_traces=4, grid_size=2, seed=1))
    d32, d8, _ = gpu.render(S.Options(width=64, height=36, max_depth=1))
    lit = d32.max(axis=2) > 1.0e20
    assert lit.mean() > 0.05 and np.isposinf(g32[lit]).any(axis=1).mean() > 0.98 and (g8[lit].max(axis=1) == 255).mean() > 0.98
    assert not (g32 < 0).any()


def test_node_ids_beyond_32_bits_are_rejected(gpu, gscenes):
    gpu.upload(gscenes["spheres1"])
    with pytest.raises(S.SkrError, match="node ids"):
        gpu.render(S.Options(width=8, height=8, max_depth=8, monte_carlo=True, num_path_traces=64))
    gpu.render(S.Options(width=8, height=8, max_depth=6, monte_carlo=True, num_path_traces=2))  # 3^5 nodes: fine


@pytest.mark.parametrize("kw", [dict(fresnel=True, max_depth=3), dict(monte_carlo=True, num_path_traces=3, max_depth=2, seed=2)])
def test_blob_between_48_and_64_kb(gpu, port, kw):
    """About 600 spheres: the scene blob is staged in shared memory but needs the opt-in size attribute on EVERY kernel that
    stages it (fresnel_expand_kernel included)."""
    rng = np.random.default_rng(11)
    sc = random_scene(rng, nspheres=600, nplights=2)
    sc.spheres[:, 3] *= 0.25
    oo, go = opts(width=64, height=36, use_shadows=True, **dict(kw))
    p32, p8, _, _ = port.render(sc, oo, rng_mode=O.RNG_PHILOX, seed=go.seed)
    gpu.upload(to_gpu_scene(sc))
    g32, g8, _ = gpu.render(go)
    assert_image_parity(g32, p32, g8, p8, min_ok=0.995, what=f"600 spheres {kw}")


def test_reserve_makes_the_first_frame_cost_what_every_frame_costs():
    r = S.Renderer()
    try:
        r.upload(S.Scene.load(os.path.join(GOLDEN, "scenes", "bear.npz")))
        o = S.Options(width=64, height=36, max_depth=3, monte_carlo=True, num_path_traces=8, use_shadows=True, seed=5)
        r.reserve(o)
        first = r.render(o)[2].ms_total
        later = min(r.render(o)[2].ms_total for _ in range(3))
        assert first < 5.0 and first < 20 * later + 1.0, (first, later)
    finally:
        r.close()


def test_async_frame_errors_surface_at_sync(gpu, gscenes):
    """Fire-and-forget frames report through the device error word; skr_sync reads it (nothing to report here)."""
    import torch
    gpu.upload(gscenes["dragon"])
    d8 = torch.zeros((90, 160, 3), dtype=torch.uint8, device="cuda")
    gpu.render_device(S.Options(width=160, height=90), d8.data_ptr(), 0, want_stats=False)
    gpu.sync()
    assert d8.any()


# ---- error behaviour ------------------------------------------------------------------------------

def test_errors_are_reported_not_swallowed(gscenes):
    r = S.Renderer()
    try:
        with pytest.raises(S.SkrError, match="skr_scene_upload must be called"):
            r.render(S.Options(width=8, height=8))
        r.upload(gscenes["spheres1"])
        with pytest.raises(S.SkrError, match="width/height"):
            r.render(S.Options(width=0, height=8))
        with pytest.raises(S.SkrError, match="tile"):
            r.render(S.Options(width=8, height=8, tile=12))
        with pytest.raises(S.SkrError, match="rank"):
            r.render(S.Options(width=8, height=8, rank=2, world=2))
    finally:
        r.close()


# ---- the C++ front end ----------------------------------------------------------------------------

def test_cli_writes_the_same_ppm(gpu, tmp_path):
    exe = os.path.join(ROOT, "host", "raytracer")
    scn = os.path.join(GOLDEN, "tiny.scn")
    out = tmp_path / "tiny.ppm"
    r = subprocess.run([exe, "--path", scn, "--output", str(out), "--width", "200", "--height", "120", "--shadow", "--parallel", "true", "--stats"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "WROTE TO PPM" in r.stdout and '"closest_hit_rays": 24000' in r.stdout
    sc = S.parseScene(scn)
    gpu.upload(sc)
    _, g8, _ = gpu.render(S.Options(width=200, height=120, use_shadows=True), want_rgb32=False)
    assert out.read_bytes() == O.ppm_bytes(g8)
    # and the reference-interface mirror writes the same file
    out2 = tmp_path / "tiny2.ppm"
    S.generate_rays_parallel(sc, S.Options(width=200, height=120, use_shadows=True), str(out2), renderer=gpu)
    assert out2.read_bytes() == out.read_bytes()


def test_cli_multi_gpu_front_end_on_all_visible_gpus(gpu, tmp_path):
    """host/raytracer --gpus 0 = frame split over every visible GPU through libskr_mgpu.so (NCCL all-gather when there is
    more than one); the PPM must be byte-identical to the single-GPU one."""
    exe = os.path.join(ROOT, "host", "raytracer")
    if not os.path.exists(os.path.join(ROOT, "skele_raytracer_b200", "libskr_mgpu.so")):
        pytest.skip("built without NCCL")
    scn = os.path.join(GOLDEN, "tiny.scn")
    a, b = tmp_path / "a.ppm", tmp_path / "b.ppm"
    common = ["--path", scn, "--width", "333", "--height", "187", "--shadow", "--gillum", "4", "--jsample", "2", "--depth", "3", "--seed", "5"]
    subprocess.run([exe, "--output", str(a)] + common, check=True, capture_output=True, timeout=300)
    r = subprocess.run([exe, "--output", str(b), "--gpus", "0", "--stats"] + common, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert '"gpus":' in r.stdout
    assert a.read_bytes() == b.read_bytes()


def test_cli_on_reference_scene_files_if_present(port, tmp_path):
    """The whole front end (.scn text -> flags -> libskr.so -> PPM) on the reference's own scene files: dragon.scn at
    640x480 must reproduce the reference's golden render byte for byte; the sphere scenes are compared with the oracle
    fed the scene as host/scene_parser.cpp read it (spheres2's fog record is parsed as intended there)."""
    d = O.REF_SCENES
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref/scenes not shipped")
    exe = os.path.join(ROOT, "host", "raytracer")
    out = tmp_path / "dragon.ppm"
    subprocess.run([exe, "--path", os.path.join(d, "dragon.scn"), "--output", str(out), "--width", "640", "--height", "480", "--depth", "1",
                    "--parallel", "true"], check=True, capture_output=True, timeout=300)
    assert hashlib.sha256(out.read_bytes()).hexdigest() == golden_testcpu()[1]
    for name, flags, kw in [("bear", ["--shadow"], dict(use_shadows=True)), ("spheres1", ["--fov", "45"], dict(fov=45.0)),
                            ("spheres2", ["--jsample", "2", "--shadow", "--seed", "4"], dict(grid_size=2, use_shadows=True)),
                            ("test", ["--gillum", "2", "--depth", "2", "--seed", "4"], dict(monte_carlo=True, num_path_traces=2, max_depth=2))]:
        out = tmp_path / (name + ".ppm")
        subprocess.run([exe, "--path", os.path.join(d, name + ".scn"), "--output", str(out), "--width", "320", "--height", "180"] + flags, check=True,
                       capture_output=True, timeout=300)
        parsed = S.parseScene(os.path.join(d, name + ".scn"))
        sc = O.Scene(parsed.spheres, parsed.tris, parsed.plights, parsed.dlights, parsed.fogs, parsed.camera, parsed.ambient, parsed.background)
        _, p8, _, _ = port.render(sc, O.Options(width=320, height=180, **kw), rng_mode=O.RNG_PHILOX, seed=4)
        got = np.frombuffer(out.read_bytes()[len(b"P6\n320 180\n255\n"):], np.uint8).reshape(180, 320, 3)
        assert (np.abs(got.astype(int) - p8.astype(int)) <= 1).all(axis=2).mean() >= 0.999, name


# ---- launch order / band targets ------------------------------------------------------------------

@pytest.mark.parametrize("scene,kw", [("spheres2", dict(width=1920, height=1080, grid_size=2, use_shadows=True, seed=5)),
                                      ("bear", dict(width=1000, height=700, use_shadows=True)),
                                      ("spheres1", dict(width=640, height=360, grid_size=2, rank=1, world=3, tile=16, seed=2))])
def test_heavy_first_tile_order_is_only_a_permutation(gpu, gscenes, scene, kw):
    """Single-kernel frames launch the tiles that can see a sphere first (the kernel's tail is then cheap sky tiles);
    every pixel -- and the compact tile buffer of the frame split -- is what scan order gives (SKR_NO_TILE_ORDER=1)."""
    import torch
    gpu.upload(gscenes[scene])
    o = S.Options(**kw)
    a32, a8, _ = gpu.render(o)
    ta = torch.zeros(gpu.tiles_bytes(o), dtype=torch.uint8, device="cuda")
    gpu.render_tiles_device(o, ta.data_ptr())
    os.environ["SKR_NO_TILE_ORDER"] = "1"
    try:
        b32, b8, _ = gpu.render(o)
        tb = torch.zeros(gpu.tiles_bytes(o), dtype=torch.uint8, device="cuda")
        gpu.render_tiles_device(o, tb.data_ptr())
    finally:
        del os.environ["SKR_NO_TILE_ORDER"]
    assert np.array_equal(a32.view(np.uint32), b32.view(np.uint32)) and np.array_equal(a8, b8)
    assert torch.equal(ta, tb)


def test_render_bands_device_sends_each_row_band_to_its_owner(gpu, gscenes):
    """skr_render_bands_device: a finished pixel of row y goes to frame min(y // rows, n - 1) only; the bands of the n frames
    put together are the whole image (here: four local buffers standing in for the GPUs of a box, two ranks rendering)."""
    import torch
    kw = dict(width=200, height=90, grid_size=2, use_shadows=True, seed=2)
    gpu.upload(gscenes["spheres2"])
    _, full8, _ = gpu.render(S.Options(**kw), want_rgb32=False)
    frames = [torch.full((90, 200, 3), 7, dtype=torch.uint8, device="cuda") for _ in range(4)]
    rows = 24
    for r in range(2):
        gpu.render_bands_device(S.Options(rank=r, world=2, tile=16, **kw), [f.data_ptr() for f in frames], rows)
    gpu.sync()
    for k, f in enumerate(frames):
        y0, y1 = k * rows, (90 if k == 3 else (k + 1) * rows)
        got = f.cpu().numpy()
        assert np.array_equal(got[y0:y1], full8[y0:y1])
        mask = np.ones(90, bool)
        mask[y0:y1] = False
        assert (got[mask] == 7).all()           # nothing outside its band
    with pytest.raises(S.SkrError, match="multiple of 4"):
        gpu.render_bands_device(S.Options(**kw), [frames[0].data_ptr()], 10)
    # --gillum frames (resolve_kernel) take the same route
    kw = dict(width=96, height=54, max_depth=2, monte_carlo=True, num_path_traces=3, seed=2)
    _, full8, _ = gpu.render(S.Options(**kw), want_rgb32=False)
    frames = [torch.zeros((54, 96, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for r in range(2):
        gpu.render_bands_device(S.Options(rank=r, world=2, **kw), [f.data_ptr() for f in frames], 28)
    gpu.sync()
    assert np.array_equal(frames[0].cpu().numpy()[:28], full8[:28]) and np.array_equal(frames[1].cpu().numpy()[28:], full8[28:])


@pytest.mark.parametrize("scene,kw", [("spheres2", dict(width=640, height=360, grid_size=5, use_shadows=True, seed=5)),      # config 2 shape: 25 samples
                                      ("bear", dict(width=333, height=187, grid_size=3, use_shadows=True, seed=6)),           # 9 samples: halves of 5 and 4
                                      ("test", dict(width=200, height=120, grid_size=3, seed=7, rank=2, world=4, tile=16))])  # triangles, a rank's share
def test_sample_split_over_two_warps_does_not_change_the_frame(gpu, gscenes, scene, kw):
    """primary_kernel sums a pixel's samples as two halves; small launches (one rank's share at world >= 4) give the halves
    to two neighbouring warps.  Forced on and off (SKR_SPLIT / SKR_NO_SPLIT) the frame must not change by a bit."""
    gpu.upload(gscenes[scene])
    o = S.Options(collect_stats=True, **kw)
    os.environ["SKR_SPLIT"] = "1"
    try:
        a32, a8, sa = gpu.render(o)
    finally:
        del os.environ["SKR_SPLIT"]
    os.environ["SKR_NO_SPLIT"] = "1"
    try:
        b32, b8, sb = gpu.render(o)
    finally:
        del os.environ["SKR_NO_SPLIT"]
    assert np.array_equal(a32.view(np.uint32), b32.view(np.uint32)) and np.array_equal(a8, b8)
    assert sa.closest_hit_rays == sb.closest_hit_rays and sa.shadow_rays == sb.shadow_rays and sa.sphere_tests == sb.sphere_tests
