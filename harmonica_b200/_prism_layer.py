"""
Layer of prisms: drop-in for ``harmonica.prism_layer`` and the
``Dataset.prism_layer.gravity()`` accessor.

Host logic restated from ``harmonica/_forward/prisms/layer.py:23-154, 238-312,
314-456``; the loop ``_forward_gravity_prism_layer`` (:522-633) runs in
``libharmonica_b200.so`` (prism boundaries are synthesised on the device from
the 1-D centre coordinates, the skip rules are applied in the reference's
order, prisms are visited easting-outer / northing-inner).

xarray and verde are optional here: ``PrismLayer`` is a numpy-level container
with the accessor's methods; when xarray is importable the same methods are
also registered as the ``prism_layer`` Dataset accessor.
"""

import ctypes
import warnings

import numpy as np

from . import _lib
from ._utils import broadcast_coordinates, observer_chunks, progress


def _check_regular_grid(easting, northing):
    """layer.py:141-154."""
    if not np.allclose(easting[1] - easting[0], easting[1:] - easting[:-1]):
        raise ValueError("Passed easting coordinates are not evenly spaced.")
    if not np.allclose(northing[1] - northing[0], northing[1:] - northing[:-1]):
        raise ValueError("Passed northing coordinates are not evenly spaced.")


def prism_layer_gravity(
    coordinates,
    easting,
    northing,
    bottom,
    top,
    density,
    field,
    thickness_threshold=None,
    progressbar=False,
    *,
    _stacklevel=2,
):
    """
    Gravity of a regular layer of prisms given as raw arrays.

    ``easting`` (n_e,), ``northing`` (n_n,): prism centres; ``bottom``, ``top``,
    ``density``: (n_n, n_e). This is what ``DatasetAccessorPrismLayer.gravity``
    computes (layer.py:376-433); the result is always float64.
    """
    if field not in _lib.FIELD_IDS:
        raise ValueError(f"Gravitational field '{field}' not recognized.")
    easting = _lib.f64(np.asarray(easting, dtype=np.float64).ravel())
    northing = _lib.f64(np.asarray(northing, dtype=np.float64).ravel())
    _check_regular_grid(easting, northing)
    shape, coords = broadcast_coordinates(coordinates)
    thickness_threshold = 0.0 if thickness_threshold is None else float(thickness_threshold)
    density = _lib.f64(density)
    bottom = _lib.f64(bottom)
    top = _lib.f64(top)
    expected = (northing.size, easting.size)
    for name, arr in (("density", density), ("bottom", bottom), ("top", top)):
        if arr.shape != expected:
            raise ValueError(
                f"Invalid {name} array with shape '{arr.shape}': expected {expected} "
                "(northing, easting)."
            )
    if np.isnan(density).any():
        warnings.warn(
            "Found NaN values in 'density' property of the prisms layer. "
            "Their respective prisms will be ignored.",
            stacklevel=_stacklevel,
        )
    lib = _lib.ensure_init()
    n_obs = coords[0].size
    out = np.empty(n_obs, dtype=np.float64)
    with progress(n_obs, progressbar) as proxy:
        for lo, hi in observer_chunks(n_obs, proxy):
            sub = tuple(np.ascontiguousarray(c[lo:hi]) for c in coords)
            res = np.empty(hi - lo, dtype=np.float64)
            flags = ctypes.c_uint32(0)
            _lib.check(
                lib.hb200_prism_layer_gravity(
                    _lib.ptr(sub[0]), _lib.ptr(sub[1]), _lib.ptr(sub[2]), hi - lo,
                    _lib.ptr(easting), easting.size, _lib.ptr(northing), northing.size,
                    _lib.ptr(bottom), _lib.ptr(top), _lib.ptr(density), thickness_threshold,
                    1 << _lib.FIELD_IDS[field], _lib.SHARD_OBSERVERS, _lib.ptr(res),
                    ctypes.byref(flags),
                )  # fmt: skip
            )
            out[lo:hi] = res
            if proxy is not None:
                proxy.update(hi - lo)
    return out.reshape(shape)


class PrismLayer:
    """
    numpy-level layer of prisms with the methods of
    ``harmonica.DatasetAccessorPrismLayer`` (layer.py:157-519).

    ``layer.prism_layer`` returns the object itself so that code written for the
    xarray accessor (``ds.prism_layer.gravity(...)``) reads the same.
    """

    def __init__(self, coordinates, surface, reference, properties=None):
        easting, northing = (np.asarray(c, dtype=np.float64) for c in coordinates[:2])
        if easting.ndim != 1 or northing.ndim != 1:
            raise ValueError("coordinates must be 1-D easting and northing arrays")
        _check_regular_grid(easting, northing)
        self.easting, self.northing = easting, northing
        self.properties = {k: np.asarray(v) for k, v in (properties or {}).items()}
        self.attrs = {"coords_units": "meters", "properties_units": "SI"}
        self.update_top_bottom(surface, reference)

    @property
    def prism_layer(self):
        return self

    @property
    def dims(self):
        return ("northing", "easting")

    @property
    def shape(self):
        return (self.northing.size, self.easting.size)

    @property
    def size(self):
        return self.northing.size * self.easting.size

    @property
    def spacing(self):
        """(s_north, s_east), layer.py:186-205."""
        return (self.northing[1] - self.northing[0], self.easting[1] - self.easting[0])

    @property
    def boundaries(self):
        """(west, east, south, north) of the whole layer, layer.py:207-226."""
        s_north, s_east = self.spacing
        return (
            self.easting.min() - s_east / 2,
            self.easting.max() + s_east / 2,
            self.northing.min() - s_north / 2,
            self.northing.max() + s_north / 2,
        )

    def update_top_bottom(self, surface, reference):
        """layer.py:264-312: top = max(surface, reference), bottom = min(...)."""
        surface = np.asarray(surface, dtype=np.float64)
        reference = np.asarray(reference, dtype=np.float64)
        if surface.shape != self.shape:
            raise ValueError(
                f"Invalid surface array with shape '{surface.shape}'. "
                + "Its shape should be compatible with the coordinates "
                + "of the layer of prisms."
            )
        if reference.ndim != 0:
            if reference.shape != self.shape:
                raise ValueError(
                    f"Invalid reference array with shape '{reference.shape}'. "
                    + "Its shape should be compatible with the coordinates "
                    + "of the layer of prisms."
                )
        else:
            reference = reference * np.ones(self.shape)
        top = surface.copy()
        bottom = reference.copy()
        reverse = surface < reference
        top[reverse] = reference[reverse]
        bottom[reverse] = surface[reverse]
        self.top, self.bottom = top, bottom

    def _get_prism_horizontal_boundaries(self, easting, northing):
        s_north, s_east = self.spacing
        return (easting - s_east / 2, easting + s_east / 2,
                northing - s_north / 2, northing + s_north / 2)  # fmt: skip

    def _to_prisms(self):
        """(n_prisms, 6) boundaries, row-major over (northing, easting), layer.py:435-456."""
        easting, northing = np.meshgrid(self.easting, self.northing)
        west, east, south, north = self._get_prism_horizontal_boundaries(
            easting.ravel(), northing.ravel()
        )
        return np.vstack((west, east, south, north, self.bottom.ravel(), self.top.ravel())).T

    def get_prism(self, indices):
        """Boundaries of the prism at ``indices = (i_north, i_east)``, layer.py:458-483."""
        west, east, south, north = self._get_prism_horizontal_boundaries(
            self.easting[indices[1]], self.northing[indices[0]]
        )
        return west, east, south, north, self.bottom[indices], self.top[indices]

    def gravity(
        self,
        coordinates,
        field,
        *,
        density_name="density",
        thickness_threshold=None,
        parallel=True,
        progressbar=False,
    ):
        """Same signature and result as ``ds.prism_layer.gravity`` (layer.py:314-433)."""
        if field not in _lib.FIELD_IDS:
            raise ValueError(f"Gravitational field '{field}' not recognized.")
        return prism_layer_gravity(
            coordinates, self.easting, self.northing, self.bottom, self.top,
            self.properties[density_name], field, thickness_threshold, progressbar,
            _stacklevel=3,
        )  # fmt: skip


def prism_layer(coordinates, surface, reference, properties=None):
    """
    Create a layer of prisms of equal horizontal size (layer.py:23-138).

    Returns an ``xarray.Dataset`` with the ``prism_layer`` accessor when xarray
    and verde are importable, otherwise a :class:`PrismLayer`.
    """
    try:
        import verde as vd  # noqa: PLC0415
        import xarray  # noqa: F401, PLC0415
    except ImportError:
        return PrismLayer(coordinates, surface, reference, properties)
    data_names = tuple(properties) if properties else None
    data = tuple(np.asarray(p) for p in properties.values()) if properties else None
    prisms = vd.make_xarray_grid(
        coordinates, data=data, data_names=data_names, dims=("northing", "easting")
    )
    _check_regular_grid(prisms.easting.values, prisms.northing.values)
    prisms.attrs = {"coords_units": "meters", "properties_units": "SI"}
    prisms.prism_layer.update_top_bottom(surface, reference)
    return prisms


def _register_xarray_accessor(xr=None):
    """Register the ``prism_layer`` Dataset accessor on ``xr`` (default: the installed xarray)."""
    if xr is None:
        try:
            import xarray as xr  # noqa: PLC0415
        except ImportError:
            return None

    @xr.register_dataset_accessor("prism_layer")
    class DatasetAccessorPrismLayer:
        """xarray flavour of :class:`PrismLayer` (layer.py:157-158)."""

        def __init__(self, xarray_obj):
            self._obj = xarray_obj

        @property
        def dims(self):
            return ("northing", "easting")

        @property
        def shape(self):
            return (self._obj.northing.size, self._obj.easting.size)

        @property
        def size(self):
            return self._obj.northing.size * self._obj.easting.size

        @property
        def spacing(self):
            return (
                self._obj.northing.values[1] - self._obj.northing.values[0],
                self._obj.easting.values[1] - self._obj.easting.values[0],
            )

        @property
        def boundaries(self):
            """layer.py:207-226."""
            return self._as_numpy(geometry_only=True).boundaries

        def update_top_bottom(self, surface, reference):
            tmp = PrismLayer(
                (self._obj.easting.values, self._obj.northing.values), surface, reference
            )
            self._obj.coords["top"] = (self.dims, tmp.top)
            self._obj.coords["bottom"] = (self.dims, tmp.bottom)

        def _get_prism_horizontal_boundaries(self, easting, northing):
            """layer.py:246-262."""
            return self._as_numpy(geometry_only=True)._get_prism_horizontal_boundaries(
                easting, northing
            )

        def _to_prisms(self):
            return self._as_numpy()._to_prisms()

        def get_prism(self, indices):
            """layer.py:458-483."""
            return self._as_numpy().get_prism(indices)

        def _as_numpy(self, geometry_only=False):
            layer = PrismLayer.__new__(PrismLayer)
            layer.easting = self._obj.easting.values
            layer.northing = self._obj.northing.values
            if not geometry_only:
                layer.top = self._obj.top.values
                layer.bottom = self._obj.bottom.values
                layer.properties = {k: self._obj[k].values for k in self._obj.data_vars}
            return layer

        def gravity(self, coordinates, field, *, density_name="density",
                    thickness_threshold=None, parallel=True, progressbar=False):  # fmt: skip
            return self._as_numpy().gravity(
                coordinates, field, density_name=density_name,
                thickness_threshold=thickness_threshold, progressbar=progressbar,
            )  # fmt: skip

    return DatasetAccessorPrismLayer


DatasetAccessorPrismLayer = _register_xarray_accessor() or PrismLayer
