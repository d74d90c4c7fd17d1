"""
``tesseroid_gravity``: forward model of tesseroids (spherical prisms) on the GPU.

Drop-in for ``harmonica.tesseroid_gravity`` (``harmonica/_forward/tesseroid_gravity.py:36-224``)
for constant densities: same signature, checks, error messages, units and signs. The pair loop
(``jit_tesseroid_gravity``, :236-339: adaptive discretisation + Gauss-Legendre point masses) runs
in ``libharmonica_b200.so`` (``hb200_tesseroid_gravity``); the O(N x P) "points outside
tesseroids" check (``_tesseroid_utils.py:398-454``) is a device scan as well.

``density`` may also be a function of the radius (variable-density tesseroids, [Soler2019]_):
see ``_tesseroid_density.py``. Not provided: a density function together with
``radial_adaptive_discretization=True`` (the leaves then have their own radial nodes, at which a
Python callable cannot be evaluated from inside the CUDA kernel).
"""

import ctypes

import numpy as np

from . import _lib
from ._tesseroid_density import _evaluate, density_at_radial_nodes, density_based_discretization
from ._utils import broadcast_coordinates, observer_chunks, progress

_FIELDS = {"potential": 0, "g_z": 3}


def _check_tesseroids(tesseroids):
    """``_tesseroid_utils.py:303-395``: boundary checks (same order, same messages)."""
    west, east, south, north, bottom, top = tuple(tesseroids[:, i] for i in range(6))
    err_msg = "Invalid tesseroid or tesseroids. "

    def fail(invalid, text):
        msg = err_msg + text
        for tess in tesseroids[invalid]:
            msg += f"\tInvalid tesseroid: {tess}\n"
        raise ValueError(msg)

    invalid = np.logical_or(
        np.logical_or(south < -90, south > 90), np.logical_or(north < -90, north > 90)
    )
    if invalid.any():
        fail(invalid, "The latitudinal boundaries must be inside the [-90, 90] degrees interval.\n")
    invalid = south > north
    if invalid.any():
        fail(invalid, "The south boundary can't be greater than the north one.\n")
    invalid = np.logical_or(bottom < 0, top < 0)
    if invalid.any():
        fail(invalid, "The bottom and top radii should be positive or zero.\n")
    invalid = bottom > top
    if invalid.any():
        fail(invalid, "The bottom radius boundary can't be greater than the top one.\n")
    invalid = np.logical_or(
        np.logical_or(west < -180, west > 360), np.logical_or(east < -180, east > 360)
    )
    if invalid.any():
        fail(invalid, "The longitudinal boundaries must be inside the [-180, 360] degrees interval.\n")
    if (west > east).any():
        tesseroids = _longitude_continuity(tesseroids)
        west, east = tesseroids[:, 0], tesseroids[:, 1]
    invalid = west > east
    if invalid.any():
        fail(invalid, "The west boundary can't be greater than the east one.\n")
    invalid = east - west > 360
    if invalid.any():
        fail(
            invalid,
            "The difference between east and west boundaries cannot be greater than "
            "one turn around the globe.\n",
        )
    return tesseroids


def _longitude_continuity(tesseroids):
    """``_tesseroid_utils.py:457-486``: west > east tesseroids are moved to [-180, 180)."""
    tesseroids = tesseroids.copy()
    west, east = tesseroids[:, 0], tesseroids[:, 1]
    change = west > east
    east[change] = ((east[change] + 180) % 360) - 180
    west[change] = ((west[change] + 180) % 360) - 180
    return tesseroids


def _discard_null_tesseroids(tesseroids, density):
    """``_tesseroid_utils.py:489-538``: zero volume or zero density."""
    west, east, south, north, bottom, top = tuple(tesseroids[:, i] for i in range(6))
    null = (west == east) | (south == north) | (bottom == top)
    null[density == 0] = True
    keep = np.logical_not(null)
    return tesseroids[keep, :], density[keep]


def _conflicting_pairs(coordinates, tesseroids, block=4096):
    """The (point, tesseroid) index pairs of ``_check_points_outside_tesseroids`` (:431-454),
    point-major like the reference's loop; only run once the device scan has found a conflict."""
    longitude, latitude, radius = coordinates
    west, east, south, north, bottom, top = (tesseroids[:, i][None, :] for i in range(6))
    pairs = []
    for start in range(0, longitude.size, block):
        lon = longitude[start:start + block, None]
        lat = latitude[start:start + block, None]
        rad = radius[start:start + block, None]
        lon_360 = lon % 360
        lon_180 = ((lon + 180) % 360) - 180
        inside = (
            (((west < lon_180) & (lon_180 < east)) | ((west < lon_360) & (lon_360 < east)))
            & (south < lat) & (lat < north) & (bottom < rad) & (rad < top)
        )  # fmt: skip
        ii, jj = np.nonzero(inside)
        pairs.extend(zip((ii + start).tolist(), jj.tolist()))
    return pairs


def check_points_outside_tesseroids(coordinates, tesseroids):
    """
    Raise ``ValueError`` if a computation point lies inside a tesseroid
    (``_tesseroid_utils.py:398-428``); the pair scan runs on the device.
    """
    coordinates = tuple(_lib.f64(np.atleast_1d(c).ravel()) for c in coordinates[:3])
    tesseroids = _lib.f64(np.atleast_2d(tesseroids))
    if coordinates[0].size == 0 or tesseroids.shape[0] == 0:
        return
    lib = _lib.ensure_init()
    flags = ctypes.c_uint32(0)
    _lib.check(
        lib.hb200_tesseroid_inside_scan(
            _lib.ptr(coordinates[0]), _lib.ptr(coordinates[1]), _lib.ptr(coordinates[2]),
            coordinates[0].size, _lib.ptr(tesseroids), tesseroids.shape[0], ctypes.byref(flags),
        )  # fmt: skip
    )
    if not flags.value & _lib.FLAG_TESS_INSIDE:
        return
    longitude, latitude, radius = coordinates
    err_msg = (
        "Found computation point(s) inside tesseroid(s). "
        "Computation points must be outside of tesseroids.\n"
    )
    for i, j in _conflicting_pairs(coordinates, tesseroids):
        west, east, south, north, bottom, top = tesseroids[j, :]
        err_msg += (
            f" - Computation point '({longitude[i]}, {latitude[i]}, {radius[i]})' "
            "inside tesseroid "
            f"'({west}, {east}, {south}, {north}, {bottom}, {top})'.\n"
        )
    raise ValueError(err_msg)


def _spread_table():
    """8 bits interleaved with zeros, for every byte value"""
    v = np.arange(256, dtype=np.uint16)
    v = (v | (v << np.uint16(4))) & np.uint16(0x0F0F)
    v = (v | (v << np.uint16(2))) & np.uint16(0x3333)
    v = (v | (v << np.uint16(1))) & np.uint16(0x5555)
    return v


_SPREAD = _spread_table()


def _locality_order(longitude, latitude):
    """
    Permutation that puts computation points that are close on the sphere next to each other
    (Morton order of longitude / latitude quantised to 8 bits each: cells of 1.4 x 0.7 degrees,
    ties in the caller's order; 16-bit keys sort in linear time). The 32 observers of a warp then
    split the same tesseroids, so their discretisation walks run in lockstep; the results do not
    depend on the order of the observers.
    """
    with np.errstate(invalid="ignore"):
        lon = longitude - 360.0 * np.floor(longitude * (1.0 / 360.0))
        qx = (lon * (255.999 / 360.0)).astype(np.int32) & 255  # NaN: any cell
        qy = ((np.clip(latitude, -90.0, 90.0) + 90.0) * (255.999 / 180.0)).astype(np.int32) & 255
    key = _SPREAD[qx] | (_SPREAD[qy] << np.uint16(1))
    return np.argsort(key, kind="stable")


def _already_local(longitude, latitude):
    """True when consecutive computation points are neighbours anyway (grids, profiles): the
    ordering would then only cost time (measured: -10 % on a row-major global grid). Looks at
    the jumps between consecutive points inside 64 runs of 64 points."""
    n = longitude.size
    starts = np.linspace(0, max(n - 64, 0), 64).astype(np.int64)
    index = (starts[:, None] + np.arange(min(64, n))[None, :]).ravel()
    lon, lat = longitude[index].reshape(64, -1), latitude[index].reshape(64, -1)
    dlon = np.abs(np.diff(lon, axis=1))
    jump = np.abs(np.diff(lat, axis=1)) + np.minimum(dlon, 360.0 - dlon)
    # random points on a sphere jump by ~120 degrees, a grid by its spacing (plus one row change)
    return bool(np.nanmean(jump) < 15.0) if jump.size else True


def tesseroid_gravity(
    coordinates,
    tesseroids,
    density,
    field,
    parallel=True,
    radial_adaptive_discretization=False,
    dtype=np.float64,
    progressbar=False,
    disable_checks=False,
    *,
    shard="auto",
    sort_observers=True,
):
    """
    Gravitational potential (J/kg) or downward acceleration ``g_z`` (mGal) of tesseroids on
    computation points given as (longitude, latitude, radius) in degrees and metres.

    ``sort_observers`` (extension): hand the computation points to the device in a locality
    preserving order (and return the result in the caller's order); it only affects speed.

    ``parallel`` is accepted for signature compatibility (the device is always parallel). With
    ``progressbar=True`` the computation points are processed in ~20 chunks and a ``tqdm`` bar
    advances per chunk. All arithmetic is float64 and the result is cast to ``dtype`` at the end.
    """
    if field not in _FIELDS:
        raise ValueError(f"Gravitational field {field} not recognized")
    shape, coords = broadcast_coordinates(coordinates)
    tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=np.float64))
    if not disable_checks:
        tesseroids = _check_tesseroids(tesseroids)
        check_points_outside_tesseroids(coords, tesseroids)
    if callable(density):
        # tesseroid_gravity.py:182-183: radial pieces in which the density is close to linear,
        # then density(radius_p) at the two radial quadrature nodes of every piece
        tesseroids = density_based_discretization(tesseroids, density)
        density_function = density
        density, density_upper = density_at_radial_nodes(tesseroids, density)
    else:
        density = np.atleast_1d(density).ravel()
        if not disable_checks and density.size != tesseroids.shape[0]:
            raise ValueError(
                f"Number of elements in density ({density.size}) "
                + f"mismatch the number of tesseroids ({tesseroids.shape[0]})"
            )
        tesseroids, density = _discard_null_tesseroids(tesseroids, density)
        density_upper = None
        density_function = None
    tesseroids, density = _lib.f64(tesseroids), _lib.f64(density)
    lib = _lib.ensure_init()
    order = None
    if sort_observers and coords[0].size >= 2048 and not _already_local(coords[0], coords[1]):
        order = _locality_order(coords[0], coords[1])
        coords = tuple(np.ascontiguousarray(c[order]) for c in coords)
    out = np.empty(coords[0].size, dtype=np.float64)
    all_flags = 0
    if density_upper is not None:
        density_upper = _lib.f64(density_upper)
    # tesseroid_gravity.py:191-207: the reference's bar advances per computation point from inside
    # the jitted loop; here the call is split in ~20 chunks of computation points
    if density_upper is not None and radial_adaptive_discretization:
        # a density function with the 3-D discretisation: the leaves have their own radial nodes,
        # which the library reports back (in batches) for the function to be evaluated at
        with progress(coords[0].size, progressbar) as proxy:
            all_flags = _density_function_call(lib, coords, tesseroids, density, density_upper,
                                               density_function, _FIELDS[field], out)
            if proxy is not None:
                proxy.update(coords[0].size)
        return _finish(out, all_flags, order, dtype, shape)
    with progress(coords[0].size, progressbar) as proxy:
        for lo, hi in observer_chunks(coords[0].size, proxy):
            whole = lo == 0 and hi == coords[0].size
            part = coords if whole else tuple(np.ascontiguousarray(c[lo:hi]) for c in coords)
            part_out = out if whole else np.empty(hi - lo, dtype=np.float64)
            flags = ctypes.c_uint32(0)
            if density_upper is None:
                rc = lib.hb200_tesseroid_gravity(
                    _lib.ptr(part[0]), _lib.ptr(part[1]), _lib.ptr(part[2]), hi - lo,
                    _lib.ptr(tesseroids), _lib.ptr(density), tesseroids.shape[0], _FIELDS[field],
                    int(bool(radial_adaptive_discretization)), _lib.shard_mode(shard),
                    _lib.ptr(part_out), ctypes.byref(flags),
                )  # fmt: skip
            else:
                rc = lib.hb200_tesseroid_gravity_variable_density(
                    _lib.ptr(part[0]), _lib.ptr(part[1]), _lib.ptr(part[2]), hi - lo,
                    _lib.ptr(tesseroids), _lib.ptr(density), _lib.ptr(density_upper),
                    tesseroids.shape[0], _FIELDS[field], _lib.shard_mode(shard), _lib.ptr(part_out),
                    ctypes.byref(flags),
                )  # fmt: skip
            _lib.check(rc)
            if not whole:
                out[lo:hi] = part_out
            all_flags |= flags.value
            if proxy is not None:
                proxy.update(hi - lo)
    return _finish(out, all_flags, order, dtype, shape)


def _density_function_call(lib, coords, tesseroids, density_lower, density_upper, density_function,
                           field_id, out):
    """hb200_tesseroid_gravity_density_function with ``density_function`` behind a ctypes callback
    (vectorised when the function accepts arrays); an exception inside it is re-raised here."""
    errors = []

    @_lib.DENSITY_FN
    def callback(radius_p, density_p, n, _user):
        values = np.ctypeslib.as_array(density_p, shape=(n,))
        try:
            values[:] = _evaluate(density_function, np.ctypeslib.as_array(radius_p, shape=(n,)).copy())
        except BaseException as error:  # noqa: BLE001 - must not propagate through the C frames
            errors.append(error)
            values[:] = np.nan

    flags = ctypes.c_uint32(0)
    rc = lib.hb200_tesseroid_gravity_density_function(
        _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), coords[0].size,
        _lib.ptr(tesseroids), _lib.ptr(density_lower), _lib.ptr(density_upper), tesseroids.shape[0],
        field_id, ctypes.cast(callback, ctypes.c_void_p), None, _lib.ptr(out), ctypes.byref(flags),
    )  # fmt: skip
    if errors:
        raise errors[0]
    _lib.check(rc)
    return flags.value


def _finish(out, all_flags, order, dtype, shape):
    """the reference's exceptions, the caller's order, dtype and shape"""
    flags = ctypes.c_uint32(all_flags)
    # the reference raises from inside the jitted loop: numba's float division raises on a zero
    # divisor (a computation point on a corner that 3-D discretisation splits without end, or
    # on a quadrature node); _tesseroid_utils.py:192-207 raise OverflowError
    if flags.value & _lib.FLAG_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    if flags.value & _lib.FLAG_TESS_STACK:
        raise OverflowError("Stack Overflow. Try to increase the stack size.")
    if flags.value & _lib.FLAG_TESS_LEAVES:
        raise OverflowError(
            "Exceeded maximum discretizations. Please increase the MAX_DISCRETIZATIONS."
        )
    if order is not None:
        unsorted = np.empty_like(out)
        unsorted[order] = out
        out = unsorted
    return out.astype(dtype, copy=False).reshape(shape)
