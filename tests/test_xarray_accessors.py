"""
The xarray Dataset accessors (``ds.prism_layer`` / ``ds.tesseroid_layer``; reference:
_forward/prisms/layer.py:157-519, _forward/tesseroid_layer.py:161-560) on a stand-in xarray
(tests/_xarray_stub.py; xarray itself is not installed in this image): every member the
reference accessor has must exist and agree with the numpy-level layer classes.
"""

import numpy as np
import numpy.testing as npt
import pytest

import _xarray_stub
from harmonica_b200 import _prism_layer, _tesseroid_layer


@pytest.fixture(scope="module")
def xr():
    m = _xarray_stub.module()
    _prism_layer._register_xarray_accessor(m)
    _tesseroid_layer._register_xarray_accessor(m)
    return m


def _prism_dataset(xr):
    easting, northing = np.linspace(-5e3, 5e3, 6), np.linspace(0, 8e3, 5)
    rng = np.random.default_rng(3)
    surface = rng.uniform(-200, 600, (5, 6))
    density = np.where(surface >= 100.0, 2670.0, -1630.0)
    ds = xr.Dataset({"density": (("northing", "easting"), density)},
                    coords={"easting": easting, "northing": northing})
    ds.prism_layer.update_top_bottom(surface, 100.0)
    layer = _prism_layer.PrismLayer((easting, northing), surface, 100.0, {"density": density})
    return ds, layer


def test_prism_layer_accessor_has_every_member_of_the_reference(xr):
    ds, layer = _prism_dataset(xr)
    acc = ds.prism_layer
    assert acc.dims == ("northing", "easting") and acc.shape == (5, 6) and acc.size == 30
    npt.assert_allclose(acc.spacing, layer.spacing)
    npt.assert_allclose(acc.boundaries, layer.boundaries)          # layer.py:207-226
    npt.assert_array_equal(ds.top.values, layer.top)
    npt.assert_array_equal(ds.bottom.values, layer.bottom)
    npt.assert_array_equal(acc._to_prisms(), layer._to_prisms())    # layer.py:435-456
    for idx in ((0, 0), (2, 3), (4, 5)):
        npt.assert_allclose(acc.get_prism(idx), layer.get_prism(idx))  # layer.py:458-483
    npt.assert_allclose(acc._get_prism_horizontal_boundaries(1000.0, 2000.0),
                        layer._get_prism_horizontal_boundaries(1000.0, 2000.0))
    with pytest.raises(ValueError, match="Invalid surface array"):
        acc.update_top_bottom(np.zeros((3, 3)), 0.0)


def test_tesseroid_layer_accessor_members(xr):
    lon, lat = np.linspace(-10, 10, 5), np.linspace(-20, 20, 9)
    surface = np.full((9, 5), 6371e3) + np.arange(45).reshape(9, 5) * 10.0
    ds = xr.Dataset({"density": (("latitude", "longitude"), np.full((9, 5), 2670.0))},
                    coords={"longitude": lon, "latitude": lat})
    ds.tesseroid_layer.update_top_bottom(surface, 6371e3 - 1e3)
    layer = _tesseroid_layer.TesseroidLayer((lon, lat), surface, 6371e3 - 1e3,
                                            {"density": np.full((9, 5), 2670.0)})
    acc = ds.tesseroid_layer
    assert acc.shape == (9, 5) and acc.size == 45 and acc.dims == ("latitude", "longitude")
    npt.assert_allclose(acc.spacing, layer.spacing)
    npt.assert_allclose(acc.boundaries, layer.boundaries)
    npt.assert_array_equal(acc._to_tesseroids(), layer._to_tesseroids())
    npt.assert_allclose(acc.get_tesseroid((3, 2)), layer.get_tesseroid((3, 2)))


@pytest.mark.gpu
def test_prism_layer_accessor_gravity_runs_the_kernel(xr, hb):
    """ds.prism_layer.gravity(...) == prism_gravity of ds.prism_layer._to_prisms() (the
    reference's own accessor test, test/test_prism_layer.py:422-456)."""
    import oracle as O

    ds, layer = _prism_dataset(xr)
    coords = (np.array([-1e3, 0.0, 2.5e3]), np.array([1e3, 4e3, 7e3]), np.full(3, 1200.0))
    for field in ("g_z", "potential", "g_ee"):
        got = ds.prism_layer.gravity(coords, field)
        want = O.prism_gravity(coords, layer._to_prisms(), layer.properties["density"].ravel(), field)
        assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want)), field
