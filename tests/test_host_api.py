"""
Host logic of the drop-in package without a GPU: signatures, validation order
and messages (reference file:line in each test), and the C ABI itself: the
library loads and exports every symbol include/harmonica_b200.h declares.
"""

import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import harmonica_b200 as hb
from harmonica_b200 import _lib
from _common import ROOT


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "harmonica_b200.h")).read()
    declared = set(re.findall(r"\b(hb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # binds every symbol; raises AttributeError if one is missing
    assert lib.hb200_version() >= 100
    assert isinstance(ctypes.cast(lib.hb200_last_error, ctypes.c_void_p).value, int)


def test_no_cpu_fallback_without_gpu():
    """the product path must fail loudly when no sm_100 device is usable"""
    lib = _lib.load()
    if lib.hb200_device_count() > 0:
        pytest.skip("a B200 is visible")
    with pytest.raises(hb.HarmonicaB200Error, match="no CPU fallback"):
        hb.prism_gravity(([0.0], [0.0], [10.0]), [-1, 1, -1, 1, -2, -1], 2670, "g_z")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "harmonica_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "oracle" not in src.replace("the oracle", ""), name


def test_signatures_match_the_reference():
    """gravity.py:51-60, magnetic.py:28-37, point.py:30-38, layer.py:314-323"""
    def names(fn, n):
        return list(inspect.signature(fn).parameters)[:n]

    assert names(hb.prism_gravity, 8) == ["coordinates", "prisms", "density", "field", "parallel",
                                          "dtype", "progressbar", "disable_checks"]
    assert names(hb.prism_magnetic, 8) == ["coordinates", "prisms", "magnetization", "field",
                                           "parallel", "dtype", "progressbar", "disable_checks"]
    assert names(hb.point_gravity, 7) == ["coordinates", "points", "masses", "field",
                                          "coordinate_system", "parallel", "dtype"]
    assert names(hb.PrismLayer.gravity, 7) == ["self", "coordinates", "field", "density_name",
                                               "thickness_threshold", "parallel", "progressbar"]
    sig = inspect.signature(hb.prism_gravity)
    assert sig.parameters["dtype"].default == "float64" and sig.parameters["parallel"].default is True
    # dipole.py:27-35, tesseroid_gravity.py:36-46, tesseroid_layer.py:282-290,
    # cartesian.py:170-178, gradient_boosted.py:98-108, spherical.py:111-117
    assert names(hb.dipole_magnetic, 8) == ["coordinates", "dipoles", "magnetic_moments", "field",
                                            "parallel", "dtype", "progressbar", "disable_checks"]
    assert names(hb.tesseroid_gravity, 9) == ["coordinates", "tesseroids", "density", "field", "parallel",
                                              "radial_adaptive_discretization", "dtype", "progressbar",
                                              "disable_checks"]
    assert names(hb.TesseroidLayer.gravity, 6) == ["self", "coordinates", "field", "progressbar",
                                                   "density_name", "thickness_threshold"]
    assert names(hb.EquivalentSources.__init__, 7) == ["self", "damping", "points", "depth", "block_size",
                                                       "parallel", "dtype"]
    assert names(hb.EquivalentSourcesGB.__init__, 9) == ["self", "damping", "points", "depth", "block_size",
                                                         "window_size", "parallel", "random_state", "dtype"]
    assert names(hb.EquivalentSourcesSph.__init__, 5) == ["self", "damping", "points", "relative_depth",
                                                          "parallel"]
    for cls in (hb.EquivalentSources, hb.EquivalentSourcesGB, hb.EquivalentSourcesSph):
        assert names(cls.fit, 4) == ["self", "coordinates", "data", "weights"]
        assert names(cls.predict, 2) == ["self", "coordinates"]


COORDS = ([0.0, 10.0], [0.0, 5.0], [10.0, 20.0])
PRISM = [-1, 1, -1, 1, -2, -1]


def test_prism_gravity_errors():
    """gravity.py:196-198, :208-213, prisms/utils.py:24-43"""
    with pytest.raises(ValueError, match="Gravitational field g_x not recognized"):
        hb.prism_gravity(COORDS, PRISM, 2670, "g_x")
    with pytest.raises(ValueError, match=r"Number of elements in density \(2\) mismatch the number of prisms \(1\)"):
        hb.prism_gravity(COORDS, PRISM, [1, 2], "g_z")
    with pytest.raises(ValueError, match="The west boundary can't be greater than the east one"):
        hb.prism_gravity(COORDS, [1, -1, -1, 1, -2, -1], 2670, "g_z")
    with pytest.raises(ValueError, match="The south boundary can't be greater than the north one"):
        hb.prism_gravity(COORDS, [-1, 1, 1, -1, -2, -1], 2670, "g_z")
    with pytest.raises(ValueError, match="The bottom boundary can't be greater than the top one"):
        hb.prism_gravity(COORDS, [-1, 1, -1, 1, -1, -2], 2670, "g_z")


def test_prism_magnetic_errors():
    """magnetic.py:103-107, :447-457"""
    M = ([1.0], [0.0], [2.0])
    with pytest.raises(ValueError, match="Invalid field 'bz'. Please choose one of 'b,b_e,b_n,b_u'."):
        hb.prism_magnetic(COORDS, PRISM, M, "bz")
    with pytest.raises(ValueError, match="Invalid magnetization vectors with '2' elements"):
        hb.prism_magnetic(COORDS, PRISM, ([1.0], [2.0]), "b")
    with pytest.raises(ValueError, match=r"Number of magnetization vectors \(2\) mismatch the number of prisms \(1\)"):
        hb.prism_magnetic(COORDS, PRISM, ([1.0, 1.0], [0.0, 1.0], [2.0, 1.0]), "b")
    with pytest.raises(ValueError, match="west boundary"):
        hb.prism_magnetic(COORDS, [1, -1, -1, 1, -2, -1], M, "b_u")


def test_point_gravity_errors():
    """_forward/utils.py:85-87, point.py:243-246, :310-315"""
    pts, m = ([0.0], [0.0], [-10.0]), [1e6]
    with pytest.raises(ValueError, match="Coordinate system geodetic not recognized."):
        hb.point_gravity(COORDS, pts, m, "g_z", coordinate_system="geodetic")
    with pytest.raises(ValueError, match=r"Number of elements in masses \(2\) mismatch the number of points \(1\)"):
        hb.point_gravity(COORDS, pts, [1, 2], "g_z")
    with pytest.raises(ValueError, match="Gravitational field 'g_x' not recognized"):
        hb.point_gravity(COORDS, pts, m, "g_x")
    with pytest.raises(ValueError, match="Gravitational field 'g_ee' not recognized"):
        hb.point_gravity(COORDS, pts, m, "g_ee", coordinate_system="spherical")
    for f in ("g_n", "g_e"):
        with pytest.raises(NotImplementedError):
            hb.point_gravity(COORDS, pts, m, f, coordinate_system="spherical")


def test_prism_layer_host_logic():
    """layer.py:141-154, :264-312, :378, :435-483"""
    east, north = np.linspace(0, 10, 5), np.linspace(2, 8, 4)
    surface = np.arange(20, dtype=float).reshape(4, 5)
    layer = hb.PrismLayer((east, north), surface, 0.0, properties={"density": 2670 * np.ones((4, 5))})
    assert layer.prism_layer is layer and layer.shape == (4, 5) and layer.size == 20
    assert [float(b) for b in layer.boundaries] == [-1.25, 11.25, 1.0, 9.0]
    assert [float(b) for b in layer.get_prism((0, 2))] == [3.75, 6.25, 1.0, 3.0, 0.0, 2.0]
    prisms = layer._to_prisms()
    assert prisms.shape == (20, 6) and list(prisms[2]) == [3.75, 6.25, 1.0, 3.0, 0.0, 2.0]
    layer.update_top_bottom(-surface, 0.0)  # surface below the reference swaps top and bottom
    assert layer.top.max() == 0.0 and layer.bottom.min() == -19.0
    with pytest.raises(ValueError, match="Gravitational field 'g_x' not recognized."):
        layer.gravity(COORDS, "g_x")
    with pytest.raises(ValueError, match="Passed easting coordinates are not evenly spaced."):
        hb.PrismLayer((np.array([0.0, 1.0, 3.0]), north), np.zeros((4, 3)), 0.0)
    with pytest.raises(ValueError, match="Invalid surface array with shape"):
        hb.PrismLayer((east, north), np.zeros((5, 4)), 0.0)


def test_eqs_host_logic():
    with pytest.raises(ValueError, match="Number of coefficients"):
        hb.eqs_predict(COORDS, ([0.0], [0.0], [-1.0]), [1.0, 2.0])
    with pytest.raises(RuntimeError, match="not fitted"):
        hb.EquivalentSources().predict(COORDS)
    # cartesian.py:164-186: constructor signature and depth validation
    assert list(inspect.signature(hb.EquivalentSources.__init__).parameters)[1:] == [
        "damping", "points", "depth", "block_size", "parallel", "dtype"]
    with pytest.raises(ValueError, match="Found invalid 'depth' value equal to 'deep'"):
        hb.EquivalentSources(depth="deep")
    with pytest.raises(ValueError, match="Depth value cannot be zero"):
        hb.EquivalentSources(depth=0)


def test_shard_argument():
    with pytest.raises(ValueError, match="Invalid shard"):
        _lib.shard_mode("rows")
    assert _lib.shard_mode("sources") == _lib.SHARD_SOURCES


def test_dipole_errors():
    """dipole.py:99-104, :278-289"""
    dip, mom = ([0.0], [0.0], [-10.0]), ([1.0], [0.0], [2.0])
    with pytest.raises(ValueError, match="Invalid field 'bz'. Please choose one of 'b, b_e, b_n, b_u'."):
        hb.dipole_magnetic(COORDS, dip, mom, "bz")
    with pytest.raises(ValueError, match="Invalid magnetic moments with '2' elements"):
        hb.dipole_magnetic(COORDS, dip, ([1.0], [2.0]), "b")
    with pytest.raises(ValueError, match=r"Number of elements in magnetic_moments \(2\) mismatch the number of dipoles \(1\)."):
        hb.dipole_magnetic(COORDS, dip, ([1.0, 1.0], [0.0, 1.0], [2.0, 1.0]), "b")
    assert list(inspect.signature(hb.dipole_magnetic).parameters)[:8] == [
        "coordinates", "dipoles", "magnetic_moments", "field", "parallel", "dtype", "progressbar",
        "disable_checks"]


def test_locality_order_of_tesseroid_observers():
    """the Morton ordering handed to the device: a permutation, NaN-safe, and local (consecutive
    points are neighbours on the sphere)"""
    from harmonica_b200 import _tesseroid as T

    rng = np.random.default_rng(5)
    lon = rng.uniform(-180, 360, 20_000)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, 20_000)))
    lon[7], lat[11] = np.nan, np.nan
    order = T._locality_order(lon, lat)
    assert np.array_equal(np.sort(order), np.arange(lon.size))
    lo, la = np.mod(lon[order], 360.0), lat[order]
    dlon = np.abs(np.diff(lo))
    jump = np.abs(np.diff(la)) + np.minimum(dlon, 360.0 - dlon)
    assert np.nanmedian(jump) < 3.0  # random order: ~100 degrees
    assert not T._already_local(lon, lat)
    assert T._already_local(lon[order], lat[order])


def test_cartesian_locality_order():
    """the observer ordering of the prism wrappers: None for small jobs and grids, else a
    NaN-safe permutation that makes consecutive points neighbours"""
    from harmonica_b200._utils import cartesian_locality_order as order_of

    rng = np.random.default_rng(8)
    e, n = rng.uniform(-5e4, 5e4, 50_000), rng.uniform(0, 2e4, 50_000)
    assert order_of(e[:1000], n[:1000], 10**6) is None  # few observers
    assert order_of(e, n, 100) is None  # few pairs
    ge, gn = np.meshgrid(np.linspace(0, 1e4, 250), np.linspace(0, 1e4, 200))
    assert order_of(ge.ravel(), gn.ravel(), 10**6) is None  # a grid is local already
    e[3], n[5] = np.nan, np.nan
    perm = order_of(e, n, 10**6)
    assert np.array_equal(np.sort(perm), np.arange(e.size))
    jump = np.abs(np.diff(e[perm])) + np.abs(np.diff(n[perm]))
    assert np.nanmedian(jump) < 0.02 * 1.2e5
    assert order_of(e[perm], n[perm], 10**6) is None


def test_bench_records_are_consistent():
    """bench.py: both arms print the same `config` object, and the committed ncu table
    (profiles/executed_per_pair.json) belongs to the kernel sources in the tree — otherwise the
    bench line would say "stale" instead of a roofline fraction"""
    import argparse

    import bench

    wl = dict(desc="w", coords=(np.zeros(4),), n_src=7, n_eval=6)
    args = argparse.Namespace(shard="observers", scaling="strong")
    assert bench.config_for(wl, "layer_gz", args, 2) == bench.config_for(wl, "layer_gz", args, 2)
    assert "gather to rank 0" in bench.config_for(wl, "layer_gz", args, 8)["sharding"]
    for workload in ("layer_gz", "c1_gz", "tensor", "mag_b", "eqs", "tess_gz"):
        entry, why = bench.executed_for(workload)
        assert entry is not None, (workload, why)
        assert entry["fp64"] > 0 and entry["other"] > 0
