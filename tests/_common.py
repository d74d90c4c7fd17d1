"""Shared helpers of the test-suite (case generators, golden loader, host harness)."""

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GRAVITY_FIELDS = ("potential", "g_e", "g_n", "g_z", "g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz")
TENSOR_FIELDS = GRAVITY_FIELDS[4:]
#: north_star tolerance: max abs error <= 1e-9 * max|field|
TOL = 1e-9


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def max_rel(got, want):
    """max|got - want| / max|want| over the non-NaN entries; NaN patterns must agree."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    scale = np.nanmax(np.abs(want)) if np.isfinite(want).any() else 1.0
    if scale == 0:
        scale = 1.0
    if not np.isfinite(want).any():
        return 0.0
    return float(np.nanmax(np.abs(got - want)) / scale)


def random_prisms(rng, n, region, zrange, half):
    c = np.stack(
        [rng.uniform(region[0], region[1], n), rng.uniform(region[2], region[3], n),
         rng.uniform(zrange[0], zrange[1], n)], axis=1)  # fmt: skip
    h = rng.uniform(half[0], half[1], (n, 3))
    return np.stack([c[:, 0] - h[:, 0], c[:, 0] + h[:, 0], c[:, 1] - h[:, 1], c[:, 1] + h[:, 1],
                     c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], axis=1)  # fmt: skip


def config1(n_prisms=10_000, n_obs=10_000, seed=1, scale=1.0):
    """BASELINE config 1: random prisms below random observers."""
    rng = np.random.default_rng(seed)
    L = 50e3 * scale
    prisms = random_prisms(rng, n_prisms, (-L, L, -L, L), (-10e3, -0.5e3 - 1e3), (50, 1e3))
    density = rng.uniform(-500, 500, n_prisms)
    density[density == 0] = 1.0
    coords = (rng.uniform(-L, L, n_obs), rng.uniform(-L, L, n_obs), rng.uniform(0, 2e3, n_obs))
    return coords, prisms, density


def layer_config2(n=500, seed=2, spacing=200.0):
    """BASELINE config 2: n x n topography layer, observers at 1 km."""
    rng = np.random.default_rng(seed)
    east_c = (np.arange(n) - (n - 1) / 2) * spacing
    north_c = (np.arange(n) - (n - 1) / 2) * spacing
    ee, nn = np.meshgrid(east_c, north_c)
    surface = np.zeros_like(ee)
    for _ in range(8):
        kx, ky = rng.uniform(-1, 1, 2) * 2 * np.pi / (n * spacing) * rng.uniform(1, 6)
        surface += rng.uniform(20, 120) * np.sin(kx * ee + ky * nn + rng.uniform(0, 2 * np.pi))
    surface = np.clip(surface + 150.0, -400.0, 800.0)
    density = np.where(surface >= 0, 2670.0, 1040.0 - 2670.0)
    flat = rng.permutation(n * n)
    k = max(1, n * n // 100)
    surface.ravel()[flat[:k]] = np.nan
    density.ravel()[flat[k:2 * k]] = 0.0
    top = np.where(surface >= 0, surface, 0.0)
    bottom = np.where(surface >= 0, 0.0, surface)
    top[np.isnan(surface)] = np.nan
    coords = (ee.ravel().copy(), nn.ravel().copy(), np.full(ee.size, 1000.0))
    return coords, east_c, north_c, bottom, top, density


# ------------------------------------------------------------------ host harness
_H = None
FS_IDS = {f: i for i, f in enumerate(GRAVITY_FIELDS)}
FS_IDS.update({"acc3": 10, "tensor6": 11, "b": 12, "b_e": 13, "b_n": 14, "b_u": 15})
_dp = ctypes.POINTER(ctypes.c_double)


def harness():
    global _H
    if _H is None:
        _H = ctypes.CDLL(os.path.join(ROOT, "tests", "harness", "libmath_harness.so"))
        _H.hbt_prism_loop.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int64, _dp, _dp, _dp, ctypes.c_int64, _dp, _dp,
            ctypes.c_uint, _dp, ctypes.POINTER(ctypes.c_uint)]  # fmt: skip
        _H.hbt_nout.restype = ctypes.c_int
    return _H


def harness_prism(fs_name, variant, coords, prisms, prm, rules=3):
    """Host build of the product's per-pair math, summed over prisms (SI, choclo signs)."""
    H = harness()
    oe, on, ou = (np.ascontiguousarray(c, dtype=np.float64) for c in coords)
    prisms = np.ascontiguousarray(np.atleast_2d(prisms), dtype=np.float64)
    prm = np.ascontiguousarray(prm, dtype=np.float64)
    fs = FS_IDS[fs_name]
    out = np.zeros((H.hbt_nout(fs), oe.size))
    flags = ctypes.c_uint(0)
    p = lambda a: a.ctypes.data_as(_dp)  # noqa: E731
    H.hbt_prism_loop(fs, variant, oe.size, p(oe), p(on), p(ou), prisms.shape[0], p(prisms), p(prm),
                     rules, p(out), ctypes.byref(flags))
    return out, flags.value
