"""The plain-C port (oracle/skr_oracle.c) pinned against
  (1) the reference's golden vector renders/testcpu.ppm (byte-exact),
  (2) float images rendered by the reference's own shade() and committed under tests/golden/ (bit-exact, every mode,
      the stochastic ones through the libc rand() call order),
  (3) the reference's own functions, compiled in place, on random inputs (bit-exact; needs oracle/_ref).
"""
import hashlib

import numpy as np
import pytest

from conftest import random_scene
from make_cases import GOLDEN_CASES
from oracle import oracle_lib as O
from test_oracle_ref import golden_testcpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_port_reproduces_testcpu_ppm(port, scenes):
    img, sha = golden_testcpu()
    _, rgb8, st, _ = port.render(scenes["dragon"], O.Options(width=640, height=480, fov=60.0, max_depth=1))
    assert hashlib.sha256(O.ppm_bytes(rgb8)).hexdigest() == sha
    assert st["closest_hit_rays"] == 640 * 480 and st["tri_tests"] == 640 * 480 * 10002


@pytest.mark.parametrize("key", sorted(GOLDEN_CASES))
def test_port_bit_identical_to_reference_images(port, scenes, ref_images, key):
    scene, kw, seed = GOLDEN_CASES[key]
    rgb32, _, _, _ = port.render(scenes[scene], O.Options(**kw), rng_mode=O.RNG_LIBC, seed=seed)
    assert np.array_equal(bits(rgb32), bits(ref_images[key]))


def test_philox_known_answers(port):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert port.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert port.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420,
                                                                                                        0x24126ea1]


def test_smallest_root_branches(port):
    L = port.lib
    inf = float("inf")
    assert L.skro_smallest_root(1.0, 0.0, 1.0) == inf           # discriminant < 0
    assert L.skro_smallest_root(1.0, -6.0, 8.0) == 2.0          # both roots positive -> the smaller (t2)
    assert L.skro_smallest_root(1.0, 2.0, -8.0) == inf          # origin inside: t2 < 0 -> never hits
    assert L.skro_smallest_root(1.0, 6.0, 8.0) == inf           # sphere behind
    assert L.skro_smallest_root(4.0, -12.0, 8.0) == 1.0         # a != 1 stays in the quadratic


def test_near_cutoff_is_one_unit(port):
    # intersection_occurs rejects t <= 1.0 (src/utils.h:173)
    import ctypes as C
    o = np.zeros(3, np.float32)
    d = np.array([0, 0, 1], np.float32)
    occ = C.c_int()
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    c = np.array([0, 0, 1.5], np.float32)
    t = port.lib.skro_sphere_hit(fp(o), fp(d), fp(c), 0.5, C.byref(occ))
    assert t == 1.0 and occ.value == 0
    c = np.array([0, 0, 1.75], np.float32)
    t = port.lib.skro_sphere_hit(fp(o), fp(d), fp(c), 0.5, C.byref(occ))
    assert t == 1.25 and occ.value == 1


def test_triangle_test_is_the_mirrored_triangle(port):
    """the reference test == textbook Moller-Trumbore on (v0, 2*v0 - v1, v2), t of either sign."""
    import ctypes as C
    rng = np.random.default_rng(3)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    agree = hits = neg = checked = 0
    for _ in range(4000):
        tri = rng.uniform(-1, 1, 9).astype(np.float32)
        o = rng.uniform(-2, 2, 3).astype(np.float32)
        # aim at the neighbourhood of the mirrored triangle, from either side of the origin
        target = tri[:3] + rng.uniform(-1, 1) * (tri[:3] - tri[3:6]) + rng.uniform(-0.2, 1.2) * (tri[6:] - tri[:3])
        d = ((target - o) * rng.choice([-1.0, 1.0]) * rng.uniform(0.5, 2)).astype(np.float32)
        tuv = np.zeros(3, np.float32)
        h = port.lib.skro_triangle_hit(fp(o), fp(d), fp(tri), fp(tuv))
        v0, v1, v2 = tri[:3].astype(np.float64), tri[3:6].astype(np.float64), tri[6:].astype(np.float64)
        m1 = 2 * v0 - v1
        e1, e2 = m1 - v0, v2 - v0
        p = np.cross(d.astype(np.float64), e2)
        det = e1 @ p
        if abs(det) < 1e-4:
            continue
        tv = o.astype(np.float64) - v0
        u = (tv @ p) / det
        q = np.cross(tv, e1)
        v = (d.astype(np.float64) @ q) / det
        t = (e2 @ q) / det
        inside = min(u, v, 1 - u - v)
        if abs(inside) < 1e-4:
            continue  # on an edge: float vs double may disagree
        expect = inside > 0
        checked += 1
        agree += int(bool(h) == expect)
        if h:
            hits += 1
            neg += tuv[0] < 0
            assert abs(tuv[0] - t) < 1e-3 * max(1, abs(t))
    assert hits > 500 and neg > 200         # hits behind the origin are accepted
    assert agree >= checked - 2


# ---- against the reference's own compiled functions (needs oracle/_ref) ------------------------

def test_helpers_bit_identical_to_reference(port, ref):
    import ctypes as C
    rng = np.random.default_rng(1)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    P, R = port.lib, ref.lib
    for _ in range(3000):
        a, b, c = (np.float32(x) for x in rng.normal(0, 3, 3))
        a = abs(a)
        x, y = P.skro_smallest_root(a, b, c), R.ref_smallest_root(a, b, c)
        assert bits(x) == bits(y)
        o = rng.normal(0, 3, 3).astype(np.float32)
        d = rng.normal(0, 1, 3).astype(np.float32)
        cen = rng.normal(0, 3, 3).astype(np.float32)
        rad = np.float32(rng.uniform(0.1, 3))
        o1, o2 = C.c_int(), C.c_int()
        assert bits(P.skro_sphere_hit(fp(o), fp(d), fp(cen), rad, C.byref(o1))) == bits(R.ref_sphere_hit(fp(o), fp(d), fp(cen), rad, C.byref(o2)))
        assert o1.value == o2.value
        tri = rng.normal(0, 2, 9).astype(np.float32)
        t1, t2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
        h1, h2 = P.skro_triangle_hit(fp(o), fp(d), fp(tri), fp(t1)), R.ref_triangle_hit(fp(o), fp(d), fp(tri), fp(t2))
        assert h1 == h2 and (not h1 or np.array_equal(bits(t1), bits(t2)))
        n = d / np.linalg.norm(d)
        a1, b1, a2, b2 = (np.zeros(3, np.float32) for _ in range(4))
        P.skro_transform_coordinate_space(fp(n), fp(a1), fp(b1))
        R.ref_transform_coordinate_space(fp(n), fp(a2), fp(b2))
        assert np.array_equal(bits(a1), bits(a2)) and np.array_equal(bits(b1), bits(b2))
        r1, r2 = np.float32(rng.uniform()), np.float32(rng.uniform())
        P.skro_uniform_sample_hemi(r1, r2, fp(a1))
        R.ref_uniform_sample_hemi(r1, r2, fp(a2))
        assert np.array_equal(bits(a1), bits(a2))
        ior = np.float32(rng.uniform(1, 1.6))
        assert bits(P.skro_fresnel(fp(d), fp(n), ior)) == bits(R.ref_fresnel(fp(d), fp(n), ior))
        P.skro_refraction(fp(d), fp(n), ior, fp(a1))
        R.ref_refraction(fp(d), fp(n), ior, fp(a2))
        assert np.array_equal(bits(a1), bits(a2))
        P.skro_reflect_direction(fp(d), fp(n), fp(a1))
        R.ref_reflect_direction(fp(d), fp(n), fp(a2))
        assert np.array_equal(bits(a1), bits(a2))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_scenes_bit_identical_to_reference(port, ref, seed):
    rng = np.random.default_rng(seed)
    sc = random_scene(rng, nspheres=7, nplights=3, ntris=20 if seed else 0, nfogs=seed % 2, ndlights=0)
    for kw in [dict(width=64, height=40, max_depth=2, use_shadows=True),
               dict(width=40, height=24, grid_size=2, use_shadows=bool(seed & 1)),
               dict(width=32, height=20, max_depth=3, monte_carlo=True, num_path_traces=3, use_shadows=True)]:
        opt = O.Options(**kw)
        a, _, _ = ref.render(sc, opt, seed=seed + 40, threads=1)
        b, _, _, _ = port.render(sc, opt, rng_mode=O.RNG_LIBC, seed=seed + 40)
        assert np.array_equal(bits(a), bits(b)), kw


@pytest.mark.parametrize("scene", ["spheres1", "spheres2_nofog", "bear"])
def test_fresnel_mode_bit_identical_to_reference_with_line44_removed(port, ref_fresnel, scenes, scene):
    for kw in [dict(width=64, height=36, max_depth=2, fresnel=True), dict(width=48, height=27, max_depth=3, use_shadows=True, fresnel=True)]:
        opt = O.Options(**kw)
        a, _, _ = ref_fresnel.render(scenes[scene], opt, seed=3, threads=1)
        b, _, _, _ = port.render(scenes[scene], opt, rng_mode=O.RNG_LIBC, seed=3)
        assert np.array_equal(bits(a), bits(b)), kw


def test_depth_only_matters_as_zero_at_head(port, scenes):
    """without --gillum, --depth 3 == --depth 1; depth <= 0 is black."""
    s = scenes["bear"]
    a, _, _, _ = port.render(s, O.Options(width=64, height=36, max_depth=1, use_shadows=True))
    b, _, _, _ = port.render(s, O.Options(width=64, height=36, max_depth=3, use_shadows=True))
    z, _, _, _ = port.render(s, O.Options(width=64, height=36, max_depth=0))
    assert np.array_equal(a, b) and not z.any()


def test_philox_mode_matches_libc_mode_in_distribution(port, scenes):
    """The keyed Philox stream must reproduce the reference's per-pixel MEAN (north star: 3 sigma over N seeds)."""
    s = scenes["spheres2"]
    opt = O.Options(width=48, height=27, max_depth=2, monte_carlo=True, num_path_traces=4, use_shadows=True)
    N = 24
    A = np.stack([port.render(s, opt, rng_mode=O.RNG_LIBC, seed=100 + i)[0] for i in range(N)]).astype(np.float64)
    B = np.stack([port.render(s, opt, rng_mode=O.RNG_PHILOX, seed=i)[0] for i in range(N)]).astype(np.float64)
    tol = 3 * np.sqrt(A.var(0, ddof=1) / N + B.var(0, ddof=1) / N) + 1 / 255
    ok = np.abs(A.mean(0) - B.mean(0)) <= tol
    assert ok.mean() >= 0.99, ok.mean()
