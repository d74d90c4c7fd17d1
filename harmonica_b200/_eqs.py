"""
Equivalent-sources prediction: the pair loop behind
``harmonica.EquivalentSources.predict`` and ``harmonica.EquivalentSourcesSph.predict``.

``predict`` (``harmonica/_equivalent_sources/utils.py:77-101``) with the Cartesian Green's
function ``1/distance`` (``cartesian.py:634-644``) or the spherical one
(``spherical.py:412-424``, geocentric longitude/latitude in degrees + radius) runs in
``libharmonica_b200.so``. The module-level functions keep the reference's calling convention
(``predict_numba_parallel(coordinates, points, coeffs, result, greens_function)`` adds into
``result``) so that ``EquivalentSources.predict`` / ``EquivalentSourcesGB._gradient_boosting``
(``gradient_boosted.py:279-286``) can bind them unchanged.
"""

import ctypes

import numpy as np

from . import _lib
from ._utils import broadcast_coordinates


def greens_func_cartesian(east, north, upward, point_east, point_north, point_upward):
    """Marker (and numpy definition) of the Cartesian Green's function, ``1/distance``."""
    return 1 / np.sqrt(
        (east - point_east) ** 2 + (north - point_north) ** 2 + (upward - point_upward) ** 2
    )


def greens_func_spherical(longitude, latitude, radius, point_longitude, point_latitude, point_radius):
    """Marker (and numpy definition) of the spherical Green's function (spherical.py:412-424)."""
    lam, phi = np.radians(longitude), np.radians(latitude)
    lam_p, phi_p = np.radians(point_longitude), np.radians(point_latitude)
    cospsi = np.sin(phi_p) * np.sin(phi) + np.cos(phi_p) * np.cos(phi) * np.cos(lam_p - lam)
    return 1 / np.sqrt((radius - point_radius) ** 2 + 2 * radius * point_radius * (1 - cospsi))


def eqs_predict(coordinates, points, coefs, dtype="float64", *, coordinate_system="cartesian",
                shard="auto"):  # fmt: skip
    """
    ``sum_j coefs[j] / distance(x_i, x'_j)`` for every observation point (no G, no units).

    ``coordinate_system="spherical"``: coordinates and points are (longitude, latitude, radius)
    with angles in degrees, like ``EquivalentSourcesSph``.
    """
    if coordinate_system not in ("cartesian", "spherical"):
        raise ValueError(f"Coordinate system {coordinate_system} not recognized.")
    shape, coords = broadcast_coordinates(coordinates)
    points = tuple(_lib.f64(np.atleast_1d(p).ravel()) for p in points[:3])
    coefs = _lib.f64(np.atleast_1d(coefs).ravel())
    if coefs.size != points[0].size:
        raise ValueError(
            f"Number of coefficients ({coefs.size}) mismatch the number of points "
            f"({points[0].size})"
        )
    lib = _lib.ensure_init()
    entry = lib.hb200_eqs_predict if coordinate_system == "cartesian" else lib.hb200_eqs_predict_spherical
    out = np.empty(coords[0].size, dtype=np.float64)
    flags = ctypes.c_uint32(0)
    _lib.check(
        entry(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), coords[0].size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), _lib.ptr(coefs),
            coefs.size, _lib.shard_mode(shard), _lib.ptr(out), ctypes.byref(flags),
        )  # fmt: skip
    )
    if flags.value & _lib.FLAG_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    return out.astype(dtype, copy=False).reshape(shape)


def _system_of(greens_function):
    name = getattr(greens_function, "__name__", "") if greens_function is not None else ""
    if greens_function is None or name in ("greens_func_cartesian", "greens"):
        return "cartesian"
    if name == "greens_func_spherical":
        return "spherical"
    raise NotImplementedError(
        "only the Cartesian and spherical 1/distance Green's functions run on the GPU"
    )


def predict_numba_parallel(coordinates, points, coeffs, result, greens_function=None):
    """
    Drop-in for ``harmonica._equivalent_sources.utils.predict_numba_parallel``:
    adds the prediction into ``result`` in place.
    """
    pred = eqs_predict(coordinates, points, coeffs, coordinate_system=_system_of(greens_function))
    result += pred.astype(result.dtype).reshape(result.shape)


predict_numba_serial = predict_numba_parallel


def eqs_jacobian(coordinates, points, dtype="float64"):
    """Dense ``(n_obs, n_src)`` matrix of ``1/distance`` (utils.py:54-74), Cartesian."""
    _, coords = broadcast_coordinates(coordinates)
    points = tuple(_lib.f64(np.atleast_1d(p).ravel()) for p in points[:3])
    lib = _lib.ensure_init()
    jac = np.empty((coords[0].size, points[0].size), dtype=np.float64)
    _lib.check(
        lib.hb200_eqs_jacobian(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), coords[0].size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), points[0].size,
            _lib.ptr(jac),
        )  # fmt: skip
    )
    return jac.astype(dtype, copy=False)


class EquivalentSources:
    """
    Prediction half of ``harmonica.EquivalentSources`` (``cartesian.py:33, 353-383``).

    Holds fitted ``points_`` and ``coefs_`` and evaluates ``predict`` on the
    GPU. Fitting (dense Jacobian + least squares) is outside this
    package's scope; pass sources fitted elsewhere.
    """

    coordinate_system = "cartesian"

    def __init__(self, points=None, coefs=None, dtype="float64"):
        self.dtype = dtype
        if points is not None:
            self.points_ = tuple(np.asarray(p).astype(dtype).ravel() for p in points[:3])
        if coefs is not None:
            self.coefs_ = np.asarray(coefs).ravel()

    def predict(self, coordinates):
        if not hasattr(self, "coefs_"):
            raise RuntimeError(f"This {type(self).__name__} instance is not fitted yet.")
        shape = np.broadcast(*coordinates[:3]).shape
        # cartesian.py:377-380: coordinates are cast to self.dtype first
        coordinates = tuple(np.atleast_1d(c).astype(self.dtype).ravel() for c in coordinates[:3])
        data = eqs_predict(coordinates, self.points_, self.coefs_, dtype=self.dtype,
                           coordinate_system=self.coordinate_system)  # fmt: skip
        return data.reshape(shape)

    def jacobian(self, coordinates, points, dtype="float64"):
        """``cartesian.py:385-415``."""
        if self.coordinate_system != "cartesian":
            raise NotImplementedError("the GPU Jacobian is implemented for Cartesian sources")
        return eqs_jacobian(coordinates, points, dtype=dtype)


class EquivalentSourcesSph(EquivalentSources):
    """
    Prediction half of ``harmonica.EquivalentSourcesSph`` (``spherical.py:219-248``):
    coordinates are (longitude, latitude, radius); the result has the dtype of the
    coordinates (``spherical.py:241-244``).
    """

    coordinate_system = "spherical"

    def predict(self, coordinates):
        if not hasattr(self, "coefs_"):
            raise RuntimeError(f"This {type(self).__name__} instance is not fitted yet.")
        shape = np.broadcast(*coordinates[:3]).shape
        dtype = np.asarray(coordinates[0]).dtype
        if dtype.kind != "f":
            dtype = np.dtype("float64")
        coordinates = tuple(np.atleast_1d(c).ravel() for c in coordinates[:3])
        data = eqs_predict(coordinates, self.points_, self.coefs_, dtype=dtype,
                           coordinate_system="spherical")  # fmt: skip
        return data.reshape(shape)
