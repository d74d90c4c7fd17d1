"""
Layer of tesseroids: drop-in for ``harmonica.tesseroid_layer`` and the
``Dataset.tesseroid_layer.gravity()`` accessor
(``harmonica/_forward/tesseroid_layer.py:20-560``).

The host logic (regular-grid and overlap checks, top / bottom from surface and reference,
expansion into ``(n, 6)`` boundaries, NaN and thickness masks) is restated here; the forward
model itself is :func:`harmonica_b200.tesseroid_gravity` (``hb200_tesseroid_gravity``).
Like ``_prism_layer``: :class:`TesseroidLayer` is a numpy-level container with the accessor's
methods; when xarray is importable the same methods are registered as the ``tesseroid_layer``
Dataset accessor.
"""

import warnings

import numpy as np

from ._tesseroid import tesseroid_gravity


def _check_regular_grid(longitude, latitude):
    """tesseroid_layer.py:135-145."""
    if not np.allclose(longitude[1] - longitude[0], longitude[1:] - longitude[:-1]):
        raise ValueError("Passed longitude coordinates are not evenly spaced.")
    if not np.allclose(latitude[1] - latitude[0], latitude[1:] - latitude[:-1]):
        raise ValueError("Passed latitude coordinates are not evenly spaced.")


def _check_overlap(longitude):
    """tesseroid_layer.py:148-158."""
    spacing = longitude[1] - longitude[0]
    if longitude.max() - longitude.min() >= 360 - spacing:
        raise ValueError(
            "Found invalid longitude coordinates that would create overlapping "
            "tesseroids around the globe."
        )


def _discard_thin_tesseroids(tesseroids, density, thickness_threshold):
    """tesseroid_layer.py:525-560."""
    thickness = tesseroids[:, -1] - tesseroids[:, -2]
    keep = np.logical_not(thickness < thickness_threshold)
    return tesseroids[keep, :], density[keep]


class TesseroidLayer:
    """
    numpy-level layer of tesseroids with the methods of
    ``harmonica.DatasetAccessorTesseroidLayer`` (tesseroid_layer.py:161-522).
    ``layer.tesseroid_layer`` returns the object itself so that code written for the xarray
    accessor reads the same.
    """

    def __init__(self, coordinates, surface, reference, properties=None):
        longitude, latitude = (np.asarray(c, dtype=np.float64) for c in coordinates[:2])
        if longitude.ndim != 1 or latitude.ndim != 1:
            raise ValueError("coordinates must be 1-D longitude and latitude arrays")
        _check_regular_grid(longitude, latitude)
        _check_overlap(longitude)
        self.longitude, self.latitude = longitude, latitude
        self.properties = {k: np.asarray(v) for k, v in (properties or {}).items()}
        self.attrs = {"longitude_units": "degrees", "latitude_units": "degrees",
                      "radius_units": "meters", "properties_units": "SI"}  # fmt: skip
        self.update_top_bottom(surface, reference)

    @property
    def tesseroid_layer(self):
        return self

    @property
    def dims(self):
        return ("latitude", "longitude")

    @property
    def spacing(self):
        """(s_latitude, s_longitude), tesseroid_layer.py:180-197."""
        _check_regular_grid(self.longitude, self.latitude)
        return (self.latitude[1] - self.latitude[0], self.longitude[1] - self.longitude[0])

    @property
    def size(self):
        return self.latitude.size * self.longitude.size

    @property
    def shape(self):
        return (self.latitude.size, self.longitude.size)

    @property
    def boundaries(self):
        """(west, east, south, north) of the whole layer, tesseroid_layer.py:211-230."""
        s_latitude, s_longitude = self.spacing
        return (
            self.longitude.min() - s_longitude / 2,
            self.longitude.max() + s_longitude / 2,
            self.latitude.min() - s_latitude / 2,
            self.latitude.max() + s_latitude / 2,
        )

    def update_top_bottom(self, surface, reference):
        """tesseroid_layer.py:232-280: top = max(surface, reference), bottom = min(...)."""
        surface = np.asarray(surface, dtype=np.float64)
        reference = np.asarray(reference, dtype=np.float64)
        if surface.shape != self.shape:
            raise ValueError(
                f"Invalid surface array with shape '{surface.shape}'. "
                + "Its shape should be compatible with the coordinates "
                + "of the layer of tesseroids."
            )
        if reference.ndim != 0:
            if reference.shape != self.shape:
                raise ValueError(
                    f"Invalid reference array with shape '{reference.shape}'. "
                    + "Its shape should be compatible with the coordinates "
                    + "of the layer of tesseroids."
                )
        else:
            reference = reference * np.ones(self.shape)
        top = surface.copy()
        bottom = reference.copy()
        reverse = surface < reference
        top[reverse] = reference[reverse]
        bottom[reverse] = surface[reverse]
        self.top, self.bottom = top, bottom

    def _get_nonans_mask(self, property_name=None):
        """tesseroid_layer.py:385-423."""
        mask = np.logical_and(np.logical_not(np.isnan(self.top)), np.logical_not(np.isnan(self.bottom)))
        if property_name is not None:
            mask_property = np.logical_not(np.isnan(self.properties[property_name]))
            if not mask_property[mask].all():
                warnings.warn(
                    f"Found missing values in '{property_name}' property "
                    + "of the tesseroid layer. The tesseroids with nan as "
                    + f"'{property_name}' will be ignored.",
                    stacklevel=1,
                )
            mask = np.logical_and(mask, mask_property)
        return mask

    def _get_tesseroid_horizontal_boundaries(self, longitude, latitude):
        s_latitude, s_longitude = self.spacing
        return (longitude - s_longitude / 2, longitude + s_longitude / 2,
                latitude - s_latitude / 2, latitude + s_latitude / 2)  # fmt: skip

    def _to_tesseroids(self):
        """(n, 6) boundaries, row-major over (latitude, longitude), tesseroid_layer.py:425-454."""
        longitude, latitude = np.meshgrid(self.longitude, self.latitude)
        west, east, south, north = self._get_tesseroid_horizontal_boundaries(
            longitude.ravel(), latitude.ravel()
        )
        return np.vstack((west, east, south, north, self.bottom.ravel(), self.top.ravel())).T

    def get_tesseroid(self, indices):
        """Boundaries of the tesseroid at ``indices = (i_latitude, i_longitude)`` (:486-522)."""
        west, east, south, north = self._get_tesseroid_horizontal_boundaries(
            self.longitude[indices[1]], self.latitude[indices[0]]
        )
        return west, east, south, north, self.bottom[indices], self.top[indices]

    def gravity(self, coordinates, field, progressbar=False, density_name="density",
                thickness_threshold=None, **kwargs):  # fmt: skip
        """Same signature and result as ``ds.tesseroid_layer.gravity`` (:282-383)."""
        boundaries = self._to_tesseroids()
        density = np.asarray(self.properties[density_name], dtype=np.float64)
        mask = self._get_nonans_mask(property_name=density_name)
        boundaries = boundaries[mask.ravel()]
        density = density[mask]
        if thickness_threshold is not None:
            boundaries, density = _discard_thin_tesseroids(boundaries, density, thickness_threshold)
        return tesseroid_gravity(coordinates, tesseroids=boundaries, density=density, field=field,
                                 progressbar=progressbar, **kwargs)  # fmt: skip


def tesseroid_layer(coordinates, surface, reference, properties=None):
    """
    Create a layer of tesseroids of equal angular size (tesseroid_layer.py:20-132).

    Returns an ``xarray.Dataset`` with the ``tesseroid_layer`` accessor when xarray and verde are
    importable, otherwise a :class:`TesseroidLayer`.
    """
    try:
        import verde as vd  # noqa: PLC0415
        import xarray  # noqa: F401, PLC0415
    except ImportError:
        return TesseroidLayer(coordinates, surface, reference, properties)
    data_names = tuple(properties) if properties else None
    data = tuple(np.asarray(p) for p in properties.values()) if properties else None
    tesseroids = vd.make_xarray_grid(
        coordinates, data=data, data_names=data_names, dims=("latitude", "longitude")
    )
    _check_regular_grid(tesseroids.longitude.values, tesseroids.latitude.values)
    _check_overlap(tesseroids.longitude.values)
    tesseroids.attrs = {"longitude_units": "degrees", "latitude_units": "degrees",
                        "radius_units": "meters", "properties_units": "SI"}  # fmt: skip
    tesseroids.tesseroid_layer.update_top_bottom(surface, reference)
    return tesseroids


def _register_xarray_accessor(xr=None):
    """Register the ``tesseroid_layer`` Dataset accessor on ``xr`` (default: the installed xarray)."""
    if xr is None:
        try:
            import xarray as xr  # noqa: PLC0415
        except ImportError:
            return None

    @xr.register_dataset_accessor("tesseroid_layer")
    class DatasetAccessorTesseroidLayer:
        """xarray flavour of :class:`TesseroidLayer` (tesseroid_layer.py:161-162)."""

        def __init__(self, xarray_obj):
            self._obj = xarray_obj

        def _as_numpy(self):
            layer = TesseroidLayer.__new__(TesseroidLayer)
            layer.longitude = self._obj.longitude.values
            layer.latitude = self._obj.latitude.values
            layer.top = self._obj.top.values
            layer.bottom = self._obj.bottom.values
            layer.properties = {k: self._obj[k].values for k in self._obj.data_vars}
            return layer

        dims = property(lambda self: ("latitude", "longitude"))
        spacing = property(lambda self: self._as_numpy().spacing)
        size = property(lambda self: self._obj.latitude.size * self._obj.longitude.size)
        shape = property(lambda self: (self._obj.latitude.size, self._obj.longitude.size))
        boundaries = property(lambda self: self._as_numpy().boundaries)

        def update_top_bottom(self, surface, reference):
            tmp = TesseroidLayer.__new__(TesseroidLayer)
            tmp.longitude, tmp.latitude = self._obj.longitude.values, self._obj.latitude.values
            tmp.update_top_bottom(surface, reference)
            self._obj.coords["top"] = (self.dims, tmp.top)
            self._obj.coords["bottom"] = (self.dims, tmp.bottom)

        def _to_tesseroids(self):
            return self._as_numpy()._to_tesseroids()

        def get_tesseroid(self, indices):
            return self._as_numpy().get_tesseroid(indices)

        def gravity(self, coordinates, field, progressbar=False, density_name="density",
                    thickness_threshold=None, **kwargs):  # fmt: skip
            return self._as_numpy().gravity(coordinates, field, progressbar=progressbar,
                                            density_name=density_name,
                                            thickness_threshold=thickness_threshold, **kwargs)  # fmt: skip

    return DatasetAccessorTesseroidLayer


DatasetAccessorTesseroidLayer = _register_xarray_accessor() or TesseroidLayer
