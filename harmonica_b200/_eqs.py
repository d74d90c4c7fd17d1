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


def _least_squares(jacobian, data, weights, damping):
    """
    ``verde.base.least_squares`` restated with the same scikit-learn calls (column scaling with
    ``StandardScaler(with_mean=False)``, then ``LinearRegression`` or ``Ridge(alpha=damping)``
    without intercept); the reference calls it at ``cartesian.py:279-280``. Host side: the dense
    solve is not part of the pairwise hot path.
    """
    from sklearn.linear_model import LinearRegression, Ridge  # noqa: PLC0415
    from sklearn.preprocessing import StandardScaler  # noqa: PLC0415

    if jacobian.shape[0] < jacobian.shape[1]:
        import warnings  # noqa: PLC0415

        warnings.warn(
            "Under-determined problem detected (ndata, nparams)={}.".format(jacobian.shape),
            stacklevel=2,
        )
    scaler = StandardScaler(copy=False, with_mean=False, with_std=True)
    jacobian = scaler.fit_transform(jacobian)
    regr = LinearRegression(fit_intercept=False) if damping is None else Ridge(
        alpha=damping, fit_intercept=False)  # fmt: skip
    regr.fit(jacobian, data.ravel(), sample_weight=weights)
    return regr.coef_ / scaler.scale_


class EquivalentSources:
    """
    ``harmonica.EquivalentSources`` (``cartesian.py:33-644``) with the pair loops on the GPU:
    ``predict`` (:353-383) and the Jacobian of ``fit`` (:385-415) run in
    ``libharmonica_b200.so``; the least-squares solve of ``fit`` (:279-280) stays on the host and
    uses the same scikit-learn calls as verde. Same constructor signature as the reference.

    Not provided here: ``block_size`` (verde's ``BlockReduce``), ``grid``/``profile``/``scatter``
    (verde's ``BaseGridder``); use :meth:`from_fitted` to evaluate sources fitted elsewhere.
    """

    coordinate_system = "cartesian"

    def __init__(self, damping=None, points=None, depth="default", block_size=None, parallel=True,
                 dtype="float64"):  # fmt: skip
        if isinstance(depth, str) and depth != "default":
            raise ValueError(
                f"Found invalid 'depth' value equal to '{depth}'. "
                "It should be 'default' or a numeric value."
            )
        if not isinstance(depth, str) and depth == 0:
            raise ValueError("Depth value cannot be zero. It should be a non-zero numeric value.")
        self.damping = damping
        self.points = points
        self.depth = depth
        self.block_size = block_size
        self.parallel = parallel
        self.dtype = dtype
        self.greens_function = greens_func_cartesian

    @classmethod
    def from_fitted(cls, points, coefs, dtype="float64"):
        """Equivalent sources whose locations and coefficients are already known."""
        self = cls(points=points, dtype=dtype)
        self.points_ = tuple(np.asarray(p).astype(dtype).ravel() for p in points[:3])
        self.coefs_ = np.asarray(coefs).ravel()
        return self

    def _build_points(self, coordinates):
        """Relative-depth sources below the data points (``cartesian.py:283-324``)."""
        if self.block_size is not None:
            raise NotImplementedError("block-averaged sources need verde.BlockReduce")
        if isinstance(self.depth, str):
            # 4.5 x the mean distance to the first neighbour
            # (bordado.neighbor_distance_statistics(coordinates[:2], "median", k=1))
            from scipy.spatial import cKDTree  # noqa: PLC0415

            xy = np.transpose([coordinates[0], coordinates[1]])
            nearest = cKDTree(xy).query(xy, k=2)[0][:, 1]
            self.depth_ = 4.5 * np.mean(nearest)
        else:
            self.depth_ = self.depth
        return coordinates[0], coordinates[1], coordinates[2] - self.depth_

    def fit(self, coordinates, data, weights=None):
        """Fit the coefficients of the equivalent sources (``cartesian.py:236-281``)."""
        coordinates = tuple(np.asarray(c) for c in coordinates[:3])
        data = np.asarray(data)
        if any(c.shape != data.shape for c in coordinates):
            raise ValueError(
                "Coordinate and data arrays must have the same shape. "
                f"Coordinates: {[c.shape for c in coordinates]}, data: {data.shape}."
            )
        if weights is not None:
            weights = np.asarray(weights)
            if weights.shape != data.shape:
                raise ValueError("Weights must have the same shape as the data array.")
            weights = weights.ravel().astype(self.dtype)
        # utils.py:16-39 (cast_fit_input), then 1-D views
        coordinates = tuple(c.astype(self.dtype).ravel() for c in coordinates)
        data = data.astype(self.dtype).ravel()
        self.region_ = (coordinates[0].min(), coordinates[0].max(),
                        coordinates[1].min(), coordinates[1].max())  # fmt: skip
        if self.points is None:
            self.points_ = tuple(p.astype(self.dtype) for p in self._build_points(coordinates))
        else:
            self.depth_ = None
            self.points_ = tuple(np.asarray(p).astype(self.dtype).ravel() for p in self.points[:3])
        jacobian = self.jacobian(coordinates, self.points_, dtype=self.dtype)
        self.coefs_ = _least_squares(jacobian, data, weights, self.damping)
        return self

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
    coordinates (``spherical.py:241-244``). Build it with :meth:`from_fitted`.
    """

    coordinate_system = "spherical"

    def __init__(self, damping=None, points=None, relative_depth=500, parallel=True):
        super().__init__(damping=damping, points=points, depth=relative_depth, parallel=parallel)
        self.relative_depth = relative_depth
        self.greens_function = greens_func_spherical

    @classmethod
    def from_fitted(cls, points, coefs, dtype="float64"):
        self = cls(points=points)
        self.points_ = tuple(np.asarray(p).astype(dtype).ravel() for p in points[:3])
        self.coefs_ = np.asarray(coefs).ravel()
        return self

    def fit(self, coordinates, data, weights=None):
        raise NotImplementedError("fitting spherical equivalent sources is not part of this package")

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
