"""The parity gate, in one place (BASELINE.json north_star / SURVEY.md §7 'hard parts').

FP64 path: per pixel |got - want| <= 1e-9 * |want| + 1e-13 on the PRE-float-cast doubles, and
identical hit/miss sets. The absolute floor covers pixels that graze the silhouette: their path
length (hence tau, I) goes to 0 while the rounding noise of a face-plane z (coordinates ~1) stays
~1e-16, so a purely relative bound is meaningless below ~1e-6. The reference's own noise floor
under a vertex-order permutation is 3.6e-14 relative.
"""
import numpy as np

REL_TOL_FP64 = 1e-9
ABS_FLOOR = 1e-13
# FP32 variant (BASELINE.json north_star: <= 1e-4 relative). Per-step geometry is single precision on
# coordinates relative to the pixel / entry depth; depths relative to the entry point reach ~1, so a float depth carries ~1e-7 of absolute noise: the floor (1e-6) covers silhouette-grazing pixels whose tau and I are themselves ~1e-4.
REL_TOL_FP32 = 1e-4
ABS_FLOOR_FP32 = 1e-6


def assert_image_parity(got_tau, got_I, want_tau, want_I, *, rel=REL_TOL_FP64, floor=ABS_FLOOR, what=""):
    for got, want, name in ((got_tau, want_tau, "tau"), (got_I, want_I, "I")):
        got = np.asarray(got)
        want = np.asarray(want)
        assert got.shape == want.shape, f"{what}{name}: shape {got.shape} vs {want.shape}"
        nan_g, nan_w = np.isnan(got), np.isnan(want)
        assert np.array_equal(nan_g, nan_w), f"{what}{name}: NaN (solid) masks differ"
        err = np.abs(np.where(nan_w, 0.0, got - want))
        tol = rel * np.abs(np.where(nan_w, 0.0, want)) + floor
        bad = err > tol
        assert not bad.any(), (f"{what}{name}: {int(bad.sum())} pixels out of tolerance, "
                               f"max abs err {err.max():.3e}")


def assert_same_hits(got_steps, want_steps, what=""):
    got_hit = np.asarray(got_steps) > 0
    want_hit = np.asarray(want_steps) > 0
    diff = int((got_hit != want_hit).sum())
    assert diff == 0, f"{what}hit/miss sets differ on {diff} pixels"


def float_ulp_distance(a, b):
    """ULP distance of two arrays of doubles that hold float-rounded values."""
    fa = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    fb = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(fa - fb)
