"""Bundle culling (skele_raytracer_b200/csrc/skr_device.cuh: cull_pairs; host tables in skr_api.cu) -- the margin
argument checked on the CPU, independently of the kernels.

The claim the kernels rely on: if the conservative test culls a sphere for a bundle of lines (apex A, central direction
w, angular radius beta, extra margin m), then the EXACT test of that sphere -- disc/4 = h*h - a*cc >= 0 evaluated in
float32 the way the kernels evaluate it -- fails for EVERY line of the bundle.  Here both tests are restated in numpy
float32 (FMA emulated through float64) and run over random and adversarial (grazing) bundles.  The GPU suite then checks
the consequence on whole frames: images and counters are bit-identical with culling off (test_bundle_culling_is_exact).
"""
import numpy as np
import pytest

F = np.float32


def fma(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64; one rounding to float32 at the end
    (the float64 addition rounds first -- a double rounding that matters only on exact ties)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def exact_disc(o, d, c, r):
    """disc/4 of the kernels' sphere test for lines (o, d) against sphere (c, r): e = o + (-c), h = d.e,
    cc = e.e - r^2, a = d.d, disc/4 = h*h - a*cc; float32 with the kernels' FMA chains (occluded / closest_sphere_table)."""
    o, d, c, r = (np.asarray(x, F) for x in (o, d, c, r))
    e = (o + (-c)).astype(F)
    h = fma(d[..., 2], e[..., 2], fma(d[..., 1], e[..., 1], (d[..., 0] * e[..., 0]).astype(F)))
    nr2 = (-(r * r)).astype(F)
    cc = fma(e[..., 2], e[..., 2], fma(e[..., 1], e[..., 1], fma(e[..., 0], e[..., 0], nr2)))
    a = fma(d[..., 2], d[..., 2], fma(d[..., 1], d[..., 1], (d[..., 0] * d[..., 0]).astype(F)))
    return fma(h, h, ((-a).astype(F) * cc).astype(F))


def cull_table(apex, c, r):
    """host side (skr_scene_upload): u = c - A, |u|^2, R = sqrt(r^2 + 1e-4 |u|^2) + 1e-4 (1 + |A|_1 + |c|_1), |u| -- in
    double, rounded up by 1e-6 relative, stored as float32."""
    apex, c = np.asarray(apex, np.float64), np.asarray(c, np.float64)
    u = c - apex
    uu = (u * u).sum(-1)
    R = np.sqrt(np.float64(r) ** 2 + 1e-4 * uu) + 1e-4 * (1.0 + np.abs(apex).sum(-1) + np.abs(c).sum(-1))
    return u.astype(F), uu.astype(F), (R * (1 + 1e-6)).astype(F), (np.sqrt(uu) * (1 + 1e-6)).astype(F)


def culled(tab, w, beta, m):
    """device side (cull_pairs): dist_c^2 = |u|^2 - (u.w)^2/|w|^2 > (R + |u| beta + m)^2, float32; NaN keeps."""
    u, uu, R, un = tab
    w = np.asarray(w, F)
    ww = fma(w[..., 2], w[..., 2], fma(w[..., 1], w[..., 1], (w[..., 0] * w[..., 0]).astype(F)))
    ninv = (-(F(1.0) / ww)).astype(F)
    hw = fma(w[..., 2], u[..., 2], fma(w[..., 1], u[..., 1], (w[..., 0] * u[..., 0]).astype(F)))
    lhs = fma((hw * hw).astype(F), ninv, uu)
    X = fma(un, np.asarray(beta, F), (R + np.asarray(m, F)).astype(F))
    return lhs > (X * X).astype(F)


def random_unit(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def perpendicular(rng, w):
    p = np.cross(w, random_unit(rng, len(w)))
    return p / np.linalg.norm(p, axis=1, keepdims=True)


@pytest.mark.parametrize("scale", [1.0, 50.0, 1000.0])
def test_camera_bundles_never_cull_a_sphere_a_ray_could_pass(scale):
    """Apex = ray origin (camera): lines d = w + delta, |delta| <= cull_delta, beta = 1.05 cull_delta / |w|."""
    rng = np.random.default_rng(int(scale))
    n, k = 40000, 24
    apex = rng.uniform(-10, 10, (n, 3)) * scale / 10
    c = apex + random_unit(rng, n) * rng.uniform(0.5, 60, (n, 1)) * scale / 10
    r = rng.uniform(0.01, 30, n) * scale / 10
    w = random_unit(rng, n) * rng.uniform(0.5, 3, (n, 1))          # un-normalised directions, like the reference's
    # aim half of the bundles at the sphere's silhouette, where culling decisions are tight
    u = c - apex
    dist = np.linalg.norm(u, axis=1)
    graze = rng.random(n) < 0.5
    tang = np.sqrt(np.maximum(dist ** 2 - r ** 2, 0))
    side = perpendicular(rng, u)
    aim = u / dist[:, None] * tang[:, None] + side * (r * rng.uniform(0.9, 1.2, n))[:, None]
    w = np.where(graze[:, None] & (dist > r)[:, None], aim / np.linalg.norm(aim, axis=1, keepdims=True) * np.linalg.norm(w, axis=1, keepdims=True), w)
    wl = np.linalg.norm(w, axis=1)
    delta = wl * rng.uniform(1e-5, 0.05, n)
    tab = cull_table(apex, c, r)
    is_culled = culled(tab, w, 1.05 * delta / wl, 0.0)
    assert 0.05 < is_culled.mean() < 0.98                           # the test does cull, and does keep
    # lines of the bundle: the extreme deviations (|delta| = bound) in k directions, and the centre
    worst = np.full(n, -np.inf)
    for j in range(k + 1):
        dv = perpendicular(rng, w) * delta[:, None] if j else np.zeros_like(w)
        d = (w + dv).astype(F)
        worst = np.maximum(worst, exact_disc(apex.astype(F), d, c.astype(F), r.astype(F)))
    bad = is_culled & (worst >= 0)
    assert not bad.any(), f"{bad.sum()} culled spheres pass the exact test"


@pytest.mark.parametrize("scale", [1.0, 20.0])
def test_shadow_bundles_never_cull_an_occluder(scale):
    """Apex = point light L; shadow rays start at hit points p within rho of pc and point at L (direction normalised,
    origin p + 1e-6 as in occluded()); beta = 1.07 rho / |pc - L|, m = 0.01 |pc - L|, bundles with rho > 0.45 |pc - L|
    are not culled at all."""
    rng = np.random.default_rng(7 + int(scale))
    n, k = 40000, 16
    L = rng.uniform(-10, 10, (n, 3)) * scale
    pc = L + random_unit(rng, n) * rng.uniform(1, 40, (n, 1)) * scale
    wl = np.linalg.norm(pc - L, axis=1)
    rho = wl * rng.uniform(1e-5, 0.44, n) * rng.choice([1e-3, 1e-2, 1.0], n)
    # occluders anywhere near the segment, many of them grazing it
    s = rng.uniform(-0.5, 1.5, (n, 1))
    on_line = L + (pc - L) * s
    r = rng.uniform(0.01, 5, n) * scale
    off = perpendicular(rng, pc - L) * (r * rng.uniform(0.0, 3.0, n) + rho * rng.uniform(0, 2, n))[:, None]
    c = on_line + off
    tab = cull_table(L, c, r)
    w = pc - L
    is_culled = culled(tab, w, 1.07 * rho / wl, 0.01 * wl)
    assert 0.05 < is_culled.mean() < 0.98
    worst = np.full(n, -np.inf)
    for j in range(k + 1):
        p = pc + (random_unit(rng, n) * rho[:, None] if j else 0.0)
        lv = (L.astype(F) - p.astype(F)).astype(F)
        d = (lv / np.linalg.norm(lv.astype(np.float64), axis=1, keepdims=True)).astype(F)
        o = (p.astype(F) + F(0.000001)).astype(F)
        worst = np.maximum(worst, exact_disc(o, d, c.astype(F), r.astype(F)))
    bad = is_culled & (worst >= 0)
    assert not bad.any(), f"{bad.sum()} culled occluders pass the exact test"


def borderline_case(rng, n, margins=True):
    """Spheres sized so that the bundle is JUST culled, far away and almost in line with the rays (dist_c << |u|, where the
    rounding of the exact test is largest against its true value), and the worst line of the bundle: deviated by the full
    delta straight towards the sphere."""
    apex = rng.uniform(-30, 30, (n, 3))
    w = random_unit(rng, n) * rng.uniform(0.5, 3, (n, 1))
    wl = np.linalg.norm(w, axis=1)
    ulen = 10 ** rng.uniform(0, 3, n)
    dist_c = ulen * 10 ** rng.uniform(-4, -0.5, n)
    perp = perpendicular(rng, w)
    along = np.sqrt(ulen ** 2 - dist_c ** 2) * rng.choice([-1.0, 1.0], n)      # in front of and behind the apex (lines!)
    u = w / wl[:, None] * along[:, None] + perp * dist_c[:, None]
    c = apex + u
    delta = wl * 10 ** rng.uniform(-7, -3, n)
    beta = 1.05 * delta / wl
    uu = (u * u).sum(1)
    absm = 1e-4 * (1.0 + np.abs(apex).sum(1) + np.abs(c).sum(1)) if margins else 0.0
    X = dist_c * (1 - 1e-5) - ulen * beta - absm                                  # R must stay just below this
    r2 = X ** 2 - (1e-4 * uu if margins else 0.0)
    ok = (X > 0) & (r2 > 0)
    r = np.sqrt(np.where(ok, r2, 1.0))
    d = (w + perp * delta[:, None]).astype(F)
    return ok, apex, c, r, w, beta, d


def test_borderline_bundles_far_small_spheres():
    rng = np.random.default_rng(11)
    ok, apex, c, r, w, beta, d = borderline_case(rng, 200000)
    assert ok.mean() > 0.3
    tab = cull_table(apex, c, r)
    is_culled = culled(tab, w, beta, 0.0) & ok
    assert is_culled.sum() > 0.5 * ok.sum()                                       # they are culled (borderline, but culled) ...
    disc = exact_disc(apex.astype(F), d, c.astype(F), r.astype(F))
    assert not (is_culled & (disc >= 0)).any()                                    # ... and the worst line still fails the exact test
    # the same construction WITHOUT the margins would be wrong: the exact float32 test passes for some "culled" spheres --
    # i.e. this test can tell a sound margin from none
    ok0, apex0, c0, r0, w0, beta0, d0 = borderline_case(np.random.default_rng(11), 200000, margins=False)
    disc0 = exact_disc(apex0.astype(F), d0, c0.astype(F), r0.astype(F))
    assert (ok0 & (disc0 >= 0)).sum() > 100


def test_padding_spheres_are_always_culled_and_nan_keeps():
    tab = (np.zeros((1, 3), F), np.array([3.0e38], F), np.zeros(1, F), np.zeros(1, F))   # the host's padding record
    assert culled(tab, np.array([[0.3, -0.2, 1.0]], F), 0.01, 0.5).all()
    u, uu, R, un = cull_table(np.zeros((1, 3)), np.array([[0.0, 0.0, 5.0]]), 1.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        assert not culled((u, uu, R, un), np.array([[0.0, 0.0, 0.0]], F), 0.0, 0.0).any()  # |w| = 0 -> NaN -> kept
