"""The ranking trick of the incoherent closest-hit loops (skr_device.cuh: rank_keys / untag / closest_sphere_xk): candidates
are ranked by ONE unsigned minimum over float bit patterns whose low 6 mantissa bits carry the sphere index.  This checks,
in numpy, the properties the kernels rely on:
  * as unsigned integers, positive floats order like floats and sit below every negative float, NaN and +inf;
  * the tagged minimum picks the smallest valid key up to the 2^-17 relative resolution the tag costs, and the LOWER index
    among keys that tie at that resolution (the reference keeps the first sphere on ties, src/raytrace.h:149-165);
  * "no candidate" (all keys negative / NaN), and v = +0 (t == 1.0 exactly, which the reference rejects) decode as a miss.
"""
import numpy as np

TAG_BITS = 6
MASK = np.uint32((1 << TAG_BITS) - 1)


def tagged_min(v):
    """v: (n, S) float32 ranking keys, S <= 64 -> (best index or -1, untagged key)."""
    bits = v.astype(np.float32).view(np.uint32)
    idx = np.arange(v.shape[1], dtype=np.uint32)[None, :]
    w = ((bits & ~MASK) | idx).min(axis=1)
    hit = (w - (MASK + np.uint32(1))) < (np.uint32(0x7F800000) - (MASK + np.uint32(1)))      # untag(): wraps for w < 64
    return np.where(hit, (w & MASK).astype(np.int64), -1), (w & ~MASK).view(np.float32)


def test_unsigned_order_of_float_bit_patterns():
    rng = np.random.default_rng(0)
    pos = np.sort(np.abs(rng.standard_cauchy(10000)).astype(np.float32) + np.float32(1e-30))
    assert (np.diff(pos.view(np.uint32).astype(np.int64)) >= 0).all()                         # same order as the floats
    top = pos.view(np.uint32).max()
    for special in (np.float32(-0.0), np.float32(-1e-30), np.float32(-3e38), np.float32(np.nan), np.float32(np.inf), np.float32(-np.inf)):
        assert np.array([special], np.float32).view(np.uint32)[0] > top                       # all above every finite positive


def test_tagged_minimum_picks_the_closest_valid_candidate():
    rng = np.random.default_rng(1)
    n, S = 200000, 31
    v = rng.uniform(-5, 20, (n, S)).astype(np.float32)
    v[rng.random((n, S)) < 0.6] = np.nan                                                      # lines that miss (sqrt of a negative)
    v[rng.random((n, S)) < 0.1] *= -1                                                         # behind the near cutoff
    best, key = tagged_min(v)
    valid = np.where(np.isfinite(v) & (v > 0), v, np.inf)
    exact = valid.min(axis=1)
    none = ~np.isfinite(exact)
    assert (best[none] == -1).all() and (best[~none] >= 0).all()
    picked = valid[np.arange(n), np.maximum(best, 0)]
    rel = (picked[~none] - exact[~none]) / exact[~none]
    assert (rel >= 0).all() and rel.max() <= 2.0 ** -17                                       # the tag costs 6 mantissa bits
    assert np.allclose(key[~none], picked[~none], rtol=2.0 ** -17, atol=0)


def test_ties_go_to_the_lower_index_and_zero_is_a_miss():
    v = np.full((3, 8), np.nan, np.float32)
    v[0, 5] = v[0, 2] = 3.25                                                                  # exact tie -> index 2
    v[1, 4] = 0.0                                                                             # t == 1.0 exactly: rejected like the reference's t > 1.0
    v[2, 7] = 1e-38                                                                           # denormal-small but positive: a hit
    best, _ = tagged_min(v)
    assert best.tolist() == [2, -1, 7]
