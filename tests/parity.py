"""Comparison helpers with the tolerances BASELINE.json's north_star states.

Deterministic modes: GPU vs reference within 1/255 per channel on >= 99.9% of pixels (silhouette / tie-break pixels
exempt).  Stochastic modes, same keyed stream: same bar (the draws are identical, only libm-level rounding differs).
Stochastic modes vs the reference's own rand() stream: per pixel-channel |mean_gpu - mean_ref| <=
3*sqrt(var_gpu/N + var_ref/N) + 1/255 on >= 99% of pixel-channels, N stated by the caller.
"""
import numpy as np

TOL = 1.0 / 255.0
MIN_OK_DET = 0.999
MIN_OK_STOCH = 0.99


def frac_pixels_within(a32, b32, tol=TOL):
    d = np.abs(np.asarray(a32, np.float64) - np.asarray(b32, np.float64))
    d = np.nan_to_num(d, nan=np.inf)
    return float((d <= tol + 1e-7).all(axis=-1).mean())


def frac_pixels_within_u8(a8, b8, tol=1):
    d = np.abs(np.asarray(a8, np.int32) - np.asarray(b8, np.int32))
    return float((d <= tol).all(axis=-1).mean())


def assert_image_parity(gpu32, ref32, gpu8=None, ref8=None, min_ok=MIN_OK_DET, what=""):
    f = frac_pixels_within(gpu32, ref32)
    assert f >= min_ok, f"{what}: only {f:.5f} of pixels within 1/255 (float image)"
    if gpu8 is not None and ref8 is not None:
        f8 = frac_pixels_within_u8(gpu8, ref8)
        assert f8 >= min_ok, f"{what}: only {f8:.5f} of pixels within 1 level (RGB8)"
    return f


def assert_mean_parity(gpu_stack, ref_stack, what=""):
    """stacks: (N, H, W, 3) float images from N independent seeds each."""
    A = np.asarray(gpu_stack, np.float64)
    B = np.asarray(ref_stack, np.float64)
    na, nb = len(A), len(B)
    tol = 3.0 * np.sqrt(A.var(0, ddof=1) / na + B.var(0, ddof=1) / nb) + TOL
    ok = np.abs(A.mean(0) - B.mean(0)) <= tol
    assert ok.mean() >= MIN_OK_STOCH, f"{what}: only {ok.mean():.4f} of pixel-channels within 3 sigma + 1/255 (N={na})"
    return float(ok.mean())
