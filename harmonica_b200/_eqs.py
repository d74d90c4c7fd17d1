"""
Equivalent sources: the pair loop behind ``harmonica.EquivalentSources.predict`` /
``EquivalentSourcesSph.predict`` and the fits that precede it (``EquivalentSources.fit``,
``EquivalentSourcesGB.fit``, ``EquivalentSourcesSph.fit``), device-resident.

``predict`` (``harmonica/_equivalent_sources/utils.py:77-101``) with the Cartesian Green's
function ``1/distance`` (``cartesian.py:634-644``) or the spherical one
(``spherical.py:412-424``, geocentric longitude/latitude in degrees + radius) runs in
``libharmonica_b200.so``. The module-level functions keep the reference's calling convention
(``predict_numba_parallel(coordinates, points, coeffs, result, greens_function)`` adds into
``result``) so that ``EquivalentSources.predict`` / ``EquivalentSourcesGB._gradient_boosting``
(``gradient_boosted.py:279-286``) can bind them unchanged.
"""

import ctypes
import warnings

import numpy as np

from . import _gridding, _lib
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


def eqs_jacobian_spherical(coordinates, points, dtype="float64"):
    """The same matrix with ``greens_func_spherical`` (spherical.py:249-283, 412-424)."""
    _, coords = broadcast_coordinates(coordinates)
    points = tuple(_lib.f64(np.atleast_1d(p).ravel()) for p in points[:3])
    lib = _lib.ensure_init()
    jac = np.empty((coords[0].size, points[0].size), dtype=np.float64)
    _lib.check(
        lib.hb200_eqs_jacobian_spherical(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), coords[0].size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), points[0].size,
            _lib.ptr(jac),
        )  # fmt: skip
    )
    return jac.astype(dtype, copy=False)


def _fit_inputs(coordinates, points, data, weights, coordinate_system):
    if coordinate_system not in ("cartesian", "spherical"):
        raise ValueError(f"Coordinate system {coordinate_system} not recognized.")
    coords = tuple(_lib.f64(np.atleast_1d(c).ravel()) for c in coordinates[:3])
    points = tuple(_lib.f64(np.atleast_1d(p).ravel()) for p in points[:3])
    data = _lib.f64(np.atleast_1d(data).ravel())
    if any(c.size != data.size for c in coords):
        raise ValueError("Coordinate and data arrays must have the same size.")
    if weights is not None:
        weights = _lib.f64(np.atleast_1d(weights).ravel())
        if weights.size != data.size:
            raise ValueError("Weights must have the same size as the data array.")
    return coords, points, data, weights


def eqs_fit(coordinates, points, data, weights=None, damping=None, *,
            coordinate_system="cartesian", return_solver_path=False):  # fmt: skip
    """
    Coefficients of the point sources ``points`` that fit ``data`` on ``coordinates``: the body
    of ``EquivalentSources.fit`` (``cartesian.py:277-280``), i.e. ``jacobian`` +
    ``verde.base.least_squares(jacobian, data, weights, damping)``, entirely on the device
    (``hb200_eqs_fit``; the Jacobian never leaves HBM). ``damping=None`` gives the minimum-norm
    least-squares solution (scipy's ``lstsq`` behind sklearn's ``LinearRegression``), a number
    the ridge solution by Cholesky (sklearn's ``Ridge``), both on the column-scaled system.
    """
    coords, points, data, weights = _fit_inputs(coordinates, points, data, weights, coordinate_system)
    if data.size < points[0].size:
        warnings.warn(
            f"Under-determined problem detected (ndata, nparams)={(data.size, points[0].size)}.",
            stacklevel=2,
        )
    lib = _lib.ensure_init()
    coefs = np.empty(points[0].size, dtype=np.float64)
    path = ctypes.c_int(-1)
    _lib.check(
        lib.hb200_eqs_fit(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), data.size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), coefs.size,
            _lib.ptr(data), _lib.ptr(weights) if weights is not None else None,
            float("nan") if damping is None else float(damping),
            int(coordinate_system == "spherical"), _lib.ptr(coefs), ctypes.byref(path),
        )  # fmt: skip
    )
    return (coefs, path.value) if return_solver_path else coefs


def eqs_fit_gradient_boosted(coordinates, points, data, weights, damping, source_windows,
                             data_windows, *, coordinate_system="cartesian"):  # fmt: skip
    """
    ``EquivalentSourcesGB._gradient_boosting`` (``gradient_boosted.py:244-293``) on the device
    (``hb200_eqs_fit_gb``): the windows are visited in the given order; returns the summed
    coefficients and the reference's ``rmse_per_iteration_``.
    """
    coords, points, data, weights = _fit_inputs(coordinates, points, data, weights, coordinate_system)
    if len(source_windows) != len(data_windows):
        raise ValueError("source_windows and data_windows must have the same length.")
    n_windows = len(source_windows)

    def pack(windows, size):
        offsets = np.zeros(n_windows + 1, dtype=np.int64)
        for k, w in enumerate(windows):
            offsets[k + 1] = offsets[k] + np.size(w)
        index = (np.concatenate([np.asarray(w, dtype=np.int64).ravel() for w in windows])
                 if n_windows else np.zeros(0, dtype=np.int64))  # fmt: skip
        if index.size and (index.min() < 0 or index.max() >= size):
            raise IndexError("window index out of range")
        return np.ascontiguousarray(index), offsets

    src_index, src_offset = pack(source_windows, points[0].size)
    data_index, data_offset = pack(data_windows, data.size)
    lib = _lib.ensure_init()
    coefs = np.zeros(points[0].size, dtype=np.float64)
    rmse = np.zeros(n_windows + 1, dtype=np.float64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    _lib.check(
        lib.hb200_eqs_fit_gb(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), data.size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), coefs.size,
            _lib.ptr(data), _lib.ptr(weights) if weights is not None else None,
            float("nan") if damping is None else float(damping),
            int(coordinate_system == "spherical"), n_windows,
            src_index.ctypes.data_as(i64p), src_offset.ctypes.data_as(i64p),
            data_index.ctypes.data_as(i64p), data_offset.ctypes.data_as(i64p),
            _lib.ptr(coefs), _lib.ptr(rmse),
        )  # fmt: skip
    )
    return coefs, rmse


def _check_fit_input(coordinates, data, weights):
    """``verde.base.check_fit_input`` for a single data component."""
    coordinates = tuple(np.asarray(c) for c in coordinates)
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
    return coordinates, data, weights


class EquivalentSources:
    """
    ``harmonica.EquivalentSources`` (``cartesian.py:33-644``) on the GPU: ``fit`` (:236-281)
    builds the Jacobian and solves the (damped) least-squares problem on the device
    (:func:`eqs_fit`), ``predict`` (:353-383) is the pair kernel. Same constructor signature,
    attributes (``points_``, ``coefs_``, ``depth_``, ``region_``) and messages as the reference.

    Not provided here: ``grid`` / ``scatter`` / ``profile`` (verde's ``BaseGridder`` returning
    xarray / pandas objects); :meth:`from_fitted` evaluates sources fitted elsewhere.
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

    def estimate_required_memory(self, coordinates):
        """Bytes of the Jacobian a fit on ``coordinates`` builds (``cartesian.py:199-234``)."""
        coordinates = _gridding.n_1d_arrays(coordinates, 3)
        points = self._build_points(coordinates)
        return coordinates[0].size * points[0].size * np.dtype(self.dtype).itemsize

    def _build_points(self, coordinates):
        """Relative-depth sources below the (block-averaged) data points (``cartesian.py:283-324``)."""
        if self.block_size is not None:
            coordinates = self._block_average_coordinates(coordinates)
        if isinstance(self.depth, str):
            # 4.5 x the mean distance to the first neighbour
            self.depth_ = 4.5 * np.mean(_gridding.neighbor_distance(coordinates[:2]))
        else:
            self.depth_ = self.depth
        return coordinates[0], coordinates[1], coordinates[2] - self.depth_

    def _block_average_coordinates(self, coordinates):
        """Block-median of the observation points (``cartesian.py:326-351``)."""
        return _gridding.block_average_coordinates(coordinates, self.block_size)

    def _prepare_fit(self, coordinates, data, weights):
        """``cartesian.py:264-276``: checks, casts, region, sources."""
        coordinates, data, weights = _check_fit_input(coordinates[:3], data, weights)
        # utils.py:16-39 (cast_fit_input)
        coordinates = tuple(c.astype(self.dtype) for c in coordinates)
        data = data.astype(self.dtype)
        if weights is not None:
            weights = weights.astype(self.dtype).ravel()
        self.region_ = _gridding.get_region(coordinates[:2])
        coordinates = _gridding.n_1d_arrays(coordinates, 3)
        if self.points is None:
            self.points_ = tuple(p.astype(self.dtype) for p in self._build_points(coordinates))
        else:
            self.depth_ = None
            self.points_ = tuple(
                p.astype(self.dtype) for p in _gridding.n_1d_arrays(self.points, 3)
            )
        return coordinates, data.ravel(), weights

    def fit(self, coordinates, data, weights=None):
        """Fit the coefficients of the equivalent sources (``cartesian.py:236-281``)."""
        coordinates, data, weights = self._prepare_fit(coordinates, data, weights)
        self.coefs_ = eqs_fit(coordinates, self.points_, data, weights, self.damping,
                              coordinate_system=self.coordinate_system)  # fmt: skip
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
        """``cartesian.py:385-415`` / ``spherical.py:249-283``: the matrix on the host."""
        if self.coordinate_system == "spherical":
            return eqs_jacobian_spherical(coordinates, points, dtype=dtype)
        return eqs_jacobian(coordinates, points, dtype=dtype)


class EquivalentSourcesGB(EquivalentSources):
    """
    ``harmonica.EquivalentSourcesGB`` (``gradient_boosted.py:23-397``): gradient-boosted
    equivalent sources fitted window by window. The whole boosting loop (window Jacobian, solve,
    prediction of the window's sources on every data point, residue update) runs on the device
    (:func:`eqs_fit_gradient_boosted`); the windows themselves are built on the host.
    """

    # gradient_boosted.py:96-97: 50 % overlap between adjacent windows
    overlapping = 0.5

    def __init__(self, damping=None, points=None, depth="default", block_size=None,
                 window_size="default", parallel=True, random_state=None, dtype="float64"):  # fmt: skip
        if isinstance(window_size, str) and window_size != "default":
            raise ValueError(
                f"Found invalid 'window_size' value equal to '{window_size}'."
                "It should be 'default' or a numeric value."
            )
        super().__init__(damping=damping, points=points, depth=depth, block_size=block_size,
                         parallel=parallel, dtype=dtype)  # fmt: skip
        self.random_state = random_state
        self.window_size = window_size

    def estimate_required_memory(self, coordinates):
        """Bytes of the largest window Jacobian (``gradient_boosted.py:142-182``)."""
        coordinates = _gridding.n_1d_arrays(coordinates, 3)
        self.points_ = self._build_points(coordinates)
        source_windows, data_windows = self._create_windows(coordinates)
        sizes = [s.size * d.size for s, d in zip(source_windows, data_windows)]
        return max(sizes) * np.dtype(self.dtype).itemsize

    def fit(self, coordinates, data, weights=None):
        """``gradient_boosted.py:184-242``."""
        coordinates, data, weights = self._prepare_fit(coordinates, data, weights)
        self.coefs_ = np.zeros_like(self.points_[0])
        self._gradient_boosting(coordinates, data, weights)
        return self

    def _gradient_boosting(self, coordinates, data, weights):
        """``gradient_boosted.py:244-293``, one device call for all windows."""
        point_windows, data_windows = self._create_windows(coordinates)
        coefs, rmse = eqs_fit_gradient_boosted(coordinates, self.points_, data, weights,
                                               self.damping, point_windows, data_windows)  # fmt: skip
        self.coefs_ = self.coefs_ + coefs.astype(self.coefs_.dtype)
        self.rmse_per_iteration_ = rmse

    def _create_windows(self, coordinates, shuffle=True):
        """Indices of sources and data points per overlapping window (``:295-374``)."""
        region = _get_region_data_sources(coordinates, self.points_)
        if isinstance(self.window_size, str):
            area = (region[1] - region[0]) * (region[3] - region[2])
            ndata = coordinates[0].size
            if ndata <= 5e3:
                warnings.warn(
                    f"Found {ndata} number of coordinates (<= 5e3). Only one window will be used.",
                    stacklevel=1,
                )
                self.window_size_ = None
                return [np.arange(self.points_[0].size)], [np.arange(ndata)]
            self.window_size_ = np.sqrt(5e3 / (ndata / area))
        else:
            self.window_size_ = self.window_size
        kwargs = {"region": region, "window_size": self.window_size_, "overlap": self.overlapping}
        source_windows = _gridding.rolling_windows(self.points_[:2], **kwargs)
        data_windows = _gridding.rolling_windows(coordinates[:2], **kwargs)
        if shuffle:
            source_windows, data_windows = _gridding.shuffle_together(
                source_windows, data_windows, random_state=self.random_state
            )
        keep = [k for k in range(len(source_windows))
                if source_windows[k].size > 0 and data_windows[k].size > 0]  # fmt: skip
        return [source_windows[k] for k in keep], [data_windows[k] for k in keep]


def _get_region_data_sources(coordinates, points):
    """Region that holds every data point and every source (``gradient_boosted.py:377-397``)."""
    data_region = _gridding.get_region(coordinates)
    sources_region = _gridding.get_region(points)
    return (
        min(data_region[0], sources_region[0]),
        max(data_region[1], sources_region[1]),
        min(data_region[2], sources_region[2]),
        max(data_region[3], sources_region[3]),
    )


class EquivalentSourcesSph(EquivalentSources):
    """
    ``harmonica.EquivalentSourcesSph`` (``spherical.py:29-424``): coordinates are
    (longitude, latitude, radius) with angles in degrees; ``fit`` (:168-216) and ``predict``
    (:219-248) run on the device. The prediction has the dtype of the coordinates (:241-244).
    """

    coordinate_system = "spherical"

    def __init__(self, damping=None, points=None, relative_depth=500, parallel=True):
        # the reference's spherical class (spherical.py:112-125) does not validate the depth:
        # relative_depth == 0 is accepted here as well (coincident sources then divide by zero
        # in fit / predict, as in the reference)
        super().__init__(damping=damping, points=points, depth="default", parallel=parallel)
        self.depth = relative_depth
        self.relative_depth = relative_depth
        self.greens_function = greens_func_spherical

    @classmethod
    def from_fitted(cls, points, coefs, dtype="float64"):
        self = cls(points=points)
        self.points_ = tuple(np.asarray(p).astype(dtype).ravel() for p in points[:3])
        self.coefs_ = np.asarray(coefs).ravel()
        return self

    def estimate_required_memory(self, coordinates):
        """``spherical.py:131-166``."""
        coordinates = _gridding.n_1d_arrays(coordinates, 3)
        n_points = coordinates[0].size if self.points is None else np.size(self.points[0])
        return coordinates[0].size * n_points * np.asarray(coordinates[0]).dtype.itemsize

    def fit(self, coordinates, data, weights=None):
        """``spherical.py:168-216``: sources at ``radius - relative_depth`` below the data."""
        coordinates, data, weights = _check_fit_input(coordinates[:3], data, weights)
        self.region_ = _gridding.get_region(coordinates[:2])
        coordinates = _gridding.n_1d_arrays(coordinates, 3)
        if self.points is None:
            self.points_ = (coordinates[0], coordinates[1], coordinates[2] - self.relative_depth)
        else:
            self.points_ = _gridding.n_1d_arrays(self.points, 3)
        self.coefs_ = eqs_fit(coordinates, self.points_, data.ravel(),
                              None if weights is None else weights.ravel(), self.damping,
                              coordinate_system="spherical")  # fmt: skip
        return self

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
