"""
Equivalent-sources prediction: the pair loop behind
``harmonica.EquivalentSources.predict``.

``predict`` (``harmonica/_equivalent_sources/utils.py:77-101``) with the
Cartesian Green's function ``1/distance`` (``cartesian.py:634-644``) runs in
``libharmonica_b200.so``. The module-level functions keep the reference's
calling convention (``predict_numba_parallel(coordinates, points, coeffs,
result, greens_function)`` adds into ``result``) so that
``EquivalentSources.predict`` / ``EquivalentSourcesGB._gradient_boosting``
(``gradient_boosted.py:279-286``) can bind them unchanged.
"""

import ctypes

import numpy as np

from . import _lib
from ._utils import broadcast_coordinates


def greens_func_cartesian(east, north, upward, point_east, point_north, point_upward):
    """Marker for the only Green's function the GPU path implements (1/distance)."""
    return 1 / np.sqrt(
        (east - point_east) ** 2 + (north - point_north) ** 2 + (upward - point_upward) ** 2
    )


def eqs_predict(coordinates, points, coefs, dtype="float64", *, shard="auto"):
    """``sum_j coefs[j] / |x_i - x'_j|`` for every observation point (no G, no units)."""
    shape, coords = broadcast_coordinates(coordinates)
    points = tuple(_lib.f64(np.atleast_1d(p).ravel()) for p in points[:3])
    coefs = _lib.f64(np.atleast_1d(coefs).ravel())
    if coefs.size != points[0].size:
        raise ValueError(
            f"Number of coefficients ({coefs.size}) mismatch the number of points "
            f"({points[0].size})"
        )
    lib = _lib.ensure_init()
    out = np.empty(coords[0].size, dtype=np.float64)
    flags = ctypes.c_uint32(0)
    _lib.check(
        lib.hb200_eqs_predict(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), coords[0].size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), _lib.ptr(coefs),
            coefs.size, _lib.shard_mode(shard), _lib.ptr(out), ctypes.byref(flags),
        )  # fmt: skip
    )
    if flags.value & _lib.FLAG_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    return out.astype(dtype, copy=False).reshape(shape)


def predict_numba_parallel(coordinates, points, coeffs, result, greens_function=None):
    """
    Drop-in for ``harmonica._equivalent_sources.utils.predict_numba_parallel``:
    adds the prediction into ``result`` in place.
    """
    if greens_function is not None and getattr(greens_function, "__name__", "") not in (
        "greens_func_cartesian",
        "greens",
    ):
        raise NotImplementedError("only the Cartesian 1/distance Green's function runs on GPU")
    result += eqs_predict(coordinates, points, coeffs).astype(result.dtype).reshape(result.shape)


predict_numba_serial = predict_numba_parallel


def eqs_jacobian(coordinates, points, dtype="float64"):
    """Dense ``(n_obs, n_src)`` matrix of ``1/distance`` (utils.py:54-74)."""
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

    def __init__(self, points=None, coefs=None, dtype="float64"):
        self.dtype = dtype
        if points is not None:
            self.points_ = tuple(np.asarray(p).astype(dtype).ravel() for p in points[:3])
        if coefs is not None:
            self.coefs_ = np.asarray(coefs).ravel()

    def predict(self, coordinates):
        if not hasattr(self, "coefs_"):
            raise RuntimeError("This EquivalentSources instance is not fitted yet.")
        shape = np.broadcast(*coordinates[:3]).shape
        # cartesian.py:377-380: coordinates are cast to self.dtype first
        coordinates = tuple(np.atleast_1d(c).astype(self.dtype).ravel() for c in coordinates[:3])
        data = eqs_predict(coordinates, self.points_, self.coefs_, dtype=self.dtype)
        return data.reshape(shape)
