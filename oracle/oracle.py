"""
oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/liboracle.so`` (built from oracle/choclo_port.c by
oracle/Makefile) plus a restatement of the HOST semantics of the reference's
wrappers (validation order, null-source discard, sign and unit conventions,
reshape). Only tests/, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of bench.py may import this module; the product
package ``harmonica_b200`` never does.

Reference lines each function follows:
  prism_gravity        src/harmonica/_forward/prisms/gravity.py:196-236, 452-486
  prism_magnetic       src/harmonica/_forward/prisms/magnetic.py:102-136, 190-200, 259-272, 430-440
  point_gravity        src/harmonica/_forward/point.py:231-261, 285-316
  prism_layer_gravity  src/harmonica/_forward/prisms/layer.py:376-433
  eqs_predict          src/harmonica/_equivalent_sources/cartesian.py:374-383
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIELD_IDS = {
    "potential": 0, "g_e": 1, "g_n": 2, "g_z": 3,
    "g_ee": 4, "g_nn": 5, "g_zz": 6, "g_en": 7, "g_ez": 8, "g_nz": 9,
}  # fmt: skip
POINT_ALIASES = {"g_ne": "g_en", "g_ze": "g_ez", "g_zn": "g_nz"}
MAG_DEFAULT_FLAGS = 3  # NaN on edges + outside-limit face rule (see choclo_port.c K4)

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
DENSITY_FN = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double)  # density(radius), tesseroids


def build():
    """Compile oracle/liboracle.so (idempotent)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        sources = [os.path.join(_HERE, f) for f in ("choclo_port.c", "tesseroid_port.c")]
        if not os.path.exists(path) or os.path.getmtime(path) < max(map(os.path.getmtime, sources)):
            build()
        L = ctypes.CDLL(path)
        L.hbo_safe_atan2.restype = ctypes.c_double
        L.hbo_safe_atan2.argtypes = [ctypes.c_double] * 2
        L.hbo_safe_log.restype = ctypes.c_double
        L.hbo_safe_log.argtypes = [ctypes.c_double] * 4
        L.hbo_prism_gravity.restype = ctypes.c_double
        L.hbo_prism_gravity.argtypes = [ctypes.c_int] + [ctypes.c_double] * 10
        L.hbo_prism_magnetic_field.restype = None
        L.hbo_prism_magnetic_field.argtypes = [ctypes.c_double] * 12 + [ctypes.c_int, _dp]
        L.hbo_point_gravity.restype = ctypes.c_double
        L.hbo_point_gravity.argtypes = (
            [ctypes.c_int] + [ctypes.c_double] * 7 + [ctypes.POINTER(ctypes.c_int)]
        )
        L.hbo_max_threads.restype = ctypes.c_int
        L.hbo_prism_gravity_loop.restype = None
        L.hbo_prism_gravity_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, ctypes.c_int, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_prism_any_singular.restype = ctypes.c_int
        L.hbo_prism_any_singular.argtypes = [ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp]
        L.hbo_prism_magnetic_loop.restype = None
        L.hbo_prism_magnetic_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, ctypes.c_int, _dp,
            ctypes.c_int,
        ]  # fmt: skip
        L.hbo_prism_layer_loop.restype = None
        L.hbo_prism_layer_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _i64, _dp, _dp, _dp, _dp, _dp,
            ctypes.c_double, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_point_cartesian_loop.restype = ctypes.c_int
        L.hbo_point_cartesian_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_point_spherical_loop.restype = ctypes.c_int
        L.hbo_point_spherical_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_eqs_predict_loop.restype = ctypes.c_int
        L.hbo_eqs_predict_loop.argtypes = [
            _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_eqs_jacobian_loop.restype = None
        L.hbo_eqs_jacobian_loop.argtypes = [
            _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_dipole_magnetic_loop.restype = ctypes.c_int
        L.hbo_dipole_magnetic_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_eqs_predict_spherical_loop.restype = ctypes.c_int
        L.hbo_eqs_predict_spherical_loop.argtypes = [
            _i64, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_tesseroid_loop.restype = ctypes.c_int
        L.hbo_tesseroid_loop.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, _dp, ctypes.c_double, ctypes.c_int, _dp,
            ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
        ]  # fmt: skip
        L.hbo_tesseroid_loop_variable_density.restype = ctypes.c_int
        L.hbo_tesseroid_loop_variable_density.argtypes = [
            ctypes.c_int, _i64, _dp, _dp, _dp, _i64, _dp, DENSITY_FN, ctypes.c_double, ctypes.c_int, _dp,
        ]  # fmt: skip
        L.hbo_adaptive_discretization.restype = ctypes.c_int64
        L.hbo_adaptive_discretization.argtypes = [
            _dp, _dp, ctypes.c_double, _dp, ctypes.c_int, _dp, _i64, ctypes.c_int,
        ]  # fmt: skip
        L.hbo_tesseroid_dimensions.restype = None
        L.hbo_tesseroid_dimensions.argtypes = [_dp, _dp, _dp, _dp]
        L.hbo_distance_tesseroid_point.restype = ctypes.c_double
        L.hbo_distance_tesseroid_point.argtypes = [_dp, _dp]
        _LIB = L
    return _LIB


def max_threads():
    return int(lib().hbo_max_threads())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _coords(coordinates):
    cast = np.broadcast(*coordinates[:3])
    arrs = tuple(_f64(np.broadcast_to(np.asarray(c, dtype=np.float64), cast.shape).ravel())
                 for c in coordinates[:3])
    return cast, arrs


def _scale(result, field):
    if field in ("g_z", "g_ez", "g_nz"):
        result *= -1
    if field in ("g_e", "g_n", "g_z"):
        result *= 1e5
    if field in ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz"):
        result *= 1e9
    return result


def discard_null_prisms(prisms, density):
    """gravity.py:478-486."""
    w, e, s, n, b, t = (prisms[:, i] for i in range(6))
    null = (w == e) | (s == n) | (b == t)
    null[density == 0] = True
    return _f64(prisms[~null, :]), _f64(density[~null])


def prism_gravity_si(coords, prisms, density, field, nthreads=0, f32acc=False):
    """The raw jitted loop (gravity.py:524-537), SI units, choclo sign."""
    oe, on, ou = coords
    out = np.zeros(oe.size, dtype=np.float64)
    lib().hbo_prism_gravity_loop(
        FIELD_IDS[field], oe.size, _p(oe), _p(on), _p(ou), prisms.shape[0], _p(prisms),
        _p(density), _p(out), int(f32acc), nthreads,
    )  # fmt: skip
    return out


def prism_gravity(coordinates, prisms, density, field, dtype="float64", nthreads=0):
    """Restatement of the reference's ``prism_gravity`` (no checks/warnings)."""
    if field not in FIELD_IDS:
        raise ValueError(f"Gravitational field {field} not recognized")
    cast, coords = _coords(coordinates)
    prisms = np.atleast_2d(np.asarray(prisms, dtype=np.float64))
    density = np.atleast_1d(np.asarray(density, dtype=np.float64)).ravel()
    prisms, density = discard_null_prisms(prisms, density)
    f32 = np.dtype(dtype) == np.float32
    out = prism_gravity_si(coords, prisms, density, field, nthreads, f32acc=f32)
    out = out.astype(dtype)
    return _scale(out, field).reshape(cast.shape)


def any_singular(coordinates, prisms, field):
    """gravity.py:239-269 (True where the reference would warn)."""
    if FIELD_IDS[field] < 4:
        return False
    _, (oe, on, ou) = _coords(coordinates)
    prisms = _f64(np.atleast_2d(prisms))
    return bool(
        lib().hbo_prism_any_singular(
            FIELD_IDS[field], oe.size, _p(oe), _p(on), _p(ou), prisms.shape[0], _p(prisms)
        )
    )


def prism_magnetic(coordinates, prisms, magnetization, field, nthreads=0, flags=MAG_DEFAULT_FLAGS):
    """Restatement of the reference's ``prism_magnetic`` (float64)."""
    if field not in ("b", "b_e", "b_n", "b_u"):
        raise ValueError(f"Invalid field '{field}'. Please choose one of 'b,b_e,b_n,b_u'.")
    cast, (oe, on, ou) = _coords(coordinates)
    prisms = np.atleast_2d(np.asarray(prisms, dtype=np.float64))
    me, mn, mu = (np.atleast_1d(np.asarray(m, dtype=np.float64)).ravel() for m in magnetization)
    w, e, s, n, b, t = (prisms[:, i] for i in range(6))
    null = (w == e) | (s == n) | (b == t)
    null[(me == 0) & (mn == 0) & (mu == 0)] = True
    prisms = _f64(prisms[~null, :])
    me, mn, mu = (_f64(m[~null]) for m in (me, mn, mu))
    comp = {"b": -1, "b_e": 0, "b_n": 1, "b_u": 2}[field]
    out = np.zeros((3 if comp < 0 else 1) * oe.size, dtype=np.float64)
    lib().hbo_prism_magnetic_loop(
        comp, oe.size, _p(oe), _p(on), _p(ou), prisms.shape[0], _p(prisms), _p(me), _p(mn), _p(mu),
        flags, _p(out), nthreads,
    )  # fmt: skip
    out *= 1e9
    if comp < 0:
        return tuple(out[i * oe.size:(i + 1) * oe.size].reshape(cast.shape) for i in range(3))
    return out.reshape(cast.shape)


def point_gravity(coordinates, points, masses, field, coordinate_system="cartesian", nthreads=0):
    """Restatement of the reference's ``point_gravity`` (float64)."""
    if coordinate_system not in ("cartesian", "spherical"):
        raise ValueError(f"Coordinate system {coordinate_system} not recognized.")
    cast, (oe, on, ou) = _coords(coordinates)
    pe, pn, pu = (_f64(np.atleast_1d(p).ravel()) for p in points[:3])
    masses = _f64(np.atleast_1d(masses).ravel())
    out = np.zeros(oe.size, dtype=np.float64)
    if coordinate_system == "cartesian":
        base = POINT_ALIASES.get(field, field)
        if base not in FIELD_IDS:
            raise ValueError(f"Gravitational field '{field}' not recognized")
        zd = lib().hbo_point_cartesian_loop(
            FIELD_IDS[base], oe.size, _p(oe), _p(on), _p(ou), pe.size, _p(pe), _p(pn), _p(pu),
            _p(masses), _p(out), nthreads,
        )  # fmt: skip
    else:
        if field in ("g_n", "g_e"):
            raise NotImplementedError
        if field not in ("potential", "g_z"):
            raise ValueError(f"Gravitational field '{field}' not recognized")
        scratch = np.empty(3 * (oe.size + pe.size), dtype=np.float64)
        zd = lib().hbo_point_spherical_loop(
            FIELD_IDS[field], oe.size, _p(oe), _p(on), _p(ou), pe.size, _p(pe), _p(pn), _p(pu),
            _p(masses), _p(out), _p(scratch), nthreads,
        )  # fmt: skip
    if zd:
        raise ZeroDivisionError("division by zero")
    if field in ("g_z", "g_ez", "g_ze", "g_nz", "g_zn"):
        out *= -1
    if field in ("g_e", "g_n", "g_z"):
        out *= 1e5
    if field in ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz", "g_ne", "g_ze", "g_zn"):
        out *= 1e9
    return out.reshape(cast.shape)


def prism_layer_gravity(coordinates, easting, northing, bottom, top, density, field,
                        thickness_threshold=None, nthreads=0):
    """Restatement of ``DatasetAccessorPrismLayer.gravity`` on raw arrays."""
    if field not in FIELD_IDS:
        raise ValueError(f"Gravitational field '{field}' not recognized.")
    cast, (oe, on, ou) = _coords(coordinates)
    easting, northing = _f64(easting), _f64(northing)
    bottom, top, density = _f64(bottom), _f64(top), _f64(density)
    thr = 0.0 if thickness_threshold is None else float(thickness_threshold)
    out = np.zeros(oe.size, dtype=np.float64)
    lib().hbo_prism_layer_loop(
        FIELD_IDS[field], oe.size, _p(oe), _p(on), _p(ou), easting.size, northing.size,
        _p(easting), _p(northing), _p(bottom), _p(top), _p(density), thr, _p(out), nthreads,
    )  # fmt: skip
    return _scale(out, field).reshape(cast.shape)


def eqs_predict(coordinates, points, coefs, nthreads=0):
    """Restatement of ``EquivalentSources.predict`` (float64)."""
    cast, (oe, on, ou) = _coords(coordinates)
    pe, pn, pu = (_f64(np.atleast_1d(p).ravel()) for p in points[:3])
    coefs = _f64(np.atleast_1d(coefs).ravel())
    out = np.zeros(oe.size, dtype=np.float64)
    zd = lib().hbo_eqs_predict_loop(
        oe.size, _p(oe), _p(on), _p(ou), pe.size, _p(pe), _p(pn), _p(pu), _p(coefs), _p(out),
        nthreads,
    )  # fmt: skip
    if zd:
        raise ZeroDivisionError("division by zero")
    return out.reshape(cast.shape)


def eqs_jacobian(coordinates, points, nthreads=0):
    _, (oe, on, ou) = _coords(coordinates)
    pe, pn, pu = (_f64(np.atleast_1d(p).ravel()) for p in points[:3])
    jac = np.empty((oe.size, pe.size), dtype=np.float64)
    lib().hbo_eqs_jacobian_loop(
        oe.size, _p(oe), _p(on), _p(ou), pe.size, _p(pe), _p(pn), _p(pu), _p(jac), nthreads
    )
    return jac


def dipole_magnetic(coordinates, dipoles, magnetic_moments, field, nthreads=0):
    """Restatement of the reference's ``dipole_magnetic`` (dipole.py:97-137, 186-200, 252-271)."""
    if field not in ("b", "b_e", "b_n", "b_u"):
        raise ValueError(f"Invalid field '{field}'. Please choose one of 'b, b_e, b_n, b_u'.")
    cast, (oe, on, ou) = _coords(coordinates)
    pe, pn, pu = (_f64(np.atleast_1d(p).ravel()) for p in dipoles[:3])
    me, mn, mu = (_f64(np.atleast_1d(m).ravel()) for m in magnetic_moments)
    comp = {"b": -1, "b_e": 0, "b_n": 1, "b_u": 2}[field]
    out = np.zeros((3 if comp < 0 else 1) * oe.size, dtype=np.float64)
    zd = lib().hbo_dipole_magnetic_loop(
        comp, oe.size, _p(oe), _p(on), _p(ou), pe.size, _p(pe), _p(pn), _p(pu), _p(me), _p(mn),
        _p(mu), _p(out), nthreads,
    )  # fmt: skip
    if zd:
        raise ZeroDivisionError("division by zero")
    out *= 1e9
    if comp < 0:
        return tuple(out[i * oe.size:(i + 1) * oe.size].reshape(cast.shape) for i in range(3))
    return out.reshape(cast.shape)


def eqs_predict_spherical(coordinates, points, coefs, nthreads=0):
    """Restatement of ``EquivalentSourcesSph.predict`` (spherical.py:219-248), float64."""
    cast, (oe, on, ou) = _coords(coordinates)
    pe, pn, pu = (_f64(np.atleast_1d(p).ravel()) for p in points[:3])
    coefs = _f64(np.atleast_1d(coefs).ravel())
    out = np.zeros(oe.size, dtype=np.float64)
    scratch = np.empty(3 * (oe.size + pe.size), dtype=np.float64)
    zd = lib().hbo_eqs_predict_spherical_loop(
        oe.size, _p(oe), _p(on), _p(ou), pe.size, _p(pe), _p(pn), _p(pu), _p(coefs), _p(out),
        _p(scratch), nthreads,
    )  # fmt: skip
    if zd:
        raise ZeroDivisionError("division by zero")
    return out.reshape(cast.shape)


# ----------------------------------------------------------------- tesseroids
TESSEROID_RATII = {"potential": 1.0, "g_z": 2.5}  # tesseroid_gravity.py:33


def longitude_continuity(tesseroids):
    """_tesseroid_utils.py:457-486: move west > east tesseroids to [-180, 180)."""
    tesseroids = np.array(tesseroids, dtype=np.float64)
    west, east = tesseroids[:, 0], tesseroids[:, 1]
    change = west > east
    east[change] = ((east[change] + 180) % 360) - 180
    west[change] = ((west[change] + 180) % 360) - 180
    return tesseroids


def discard_null_tesseroids(tesseroids, density):
    """_tesseroid_utils.py:489-538."""
    west, east, south, north, bottom, top = (tesseroids[:, i] for i in range(6))
    null = (west == east) | (south == north) | (bottom == top)
    null[density == 0] = True
    return tesseroids[~null], density[~null]


def adaptive_discretization(coordinates, tesseroid, distance_size_ratio, radial=False,
                            stack_size=100, max_small=100000):
    """The leaves of ``_adaptive_discretization`` (_tesseroid_utils.py:136-217) as an array;
    raises OverflowError where the reference does."""
    coordinates = _f64(np.asarray(coordinates, dtype=float).ravel())
    tesseroid = _f64(np.asarray(tesseroid, dtype=float).ravel())
    stack = np.empty((stack_size, 6))
    small = np.empty((max_small, 6))
    n = lib().hbo_adaptive_discretization(_p(coordinates), _p(tesseroid), distance_size_ratio,
                                          _p(stack), stack_size, _p(small), max_small, int(radial))
    if n == -1:
        raise OverflowError("Stack Overflow. Try to increase the stack size.")
    if n == -2:
        raise OverflowError("Exceeded maximum discretizations. Please increase the MAX_DISCRETIZATIONS.")
    if n == -4:
        raise ZeroDivisionError("division by zero")
    return small[:n].copy()


def tesseroid_gravity(coordinates, tesseroids, density, field, radial_adaptive_discretization=False,
                      nthreads=0, return_counts=False):
    """Restatement of ``tesseroid_gravity`` (tesseroid_gravity.py:36-224) for constant densities
    and valid models (the input checks are restated in the product and tested against the
    reference's expectations directly)."""
    if field not in TESSEROID_RATII:
        raise ValueError(f"Gravitational field {field} not recognized")
    cast, (lon, lat, rad) = _coords(coordinates)
    tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=np.float64))
    if (tesseroids[:, 0] > tesseroids[:, 1]).any():
        tesseroids = longitude_continuity(tesseroids)
    density = np.atleast_1d(np.asarray(density, dtype=np.float64)).ravel()
    tesseroids, density = discard_null_tesseroids(tesseroids, density)
    tesseroids, density = _f64(tesseroids), _f64(density)
    out = np.zeros(lon.size, dtype=np.float64)
    counts = np.zeros((lon.size, tesseroids.shape[0]), dtype=np.int64) if return_counts else None
    status = lib().hbo_tesseroid_loop(
        FIELD_IDS[field], lon.size, _p(lon), _p(lat), _p(rad), tesseroids.shape[0], _p(tesseroids),
        _p(density), TESSEROID_RATII[field], int(bool(radial_adaptive_discretization)), _p(out),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if return_counts else None, nthreads,
    )  # fmt: skip
    if status & 1:
        raise OverflowError("Stack Overflow. Try to increase the stack size.")
    if status & 2:
        raise OverflowError("Exceeded maximum discretizations. Please increase the MAX_DISCRETIZATIONS.")
    if status & 4:
        raise ZeroDivisionError("division by zero")  # numba's error model for float division
    if field == "g_z":
        out *= -1
        out *= 1e5
    out = out.reshape(cast.shape)
    return (out, counts) if return_counts else out


# ---- tesseroids with a density function (_forward/_tesseroid_variable_density.py)
DELTA_RATIO = 0.1  # _tesseroid_variable_density.py:17


def density_minmax(density, bottom, top):
    """:159-199: extrema of the density inside [bottom, top] (bounded scalar minimisation, and
    the values at the two ends)."""
    from scipy.optimize import minimize_scalar

    ends = sorted([density(bottom), density(top)])
    kwargs = {"bounds": [bottom, top], "method": "bounded"}
    minimum = min(minimize_scalar(density, **kwargs).fun, ends[0])
    maximum = max(-minimize_scalar(lambda radius: -density(radius), **kwargs).fun, ends[1])
    return minimum, maximum


def straight_line(radius, normalized_density, bottom, top):
    """:238-259: the chord of the normalised density between the two ends."""
    at_bottom, at_top = normalized_density(bottom), normalized_density(top)
    slope = (at_top - at_bottom) / (top - bottom)
    return slope * (radius - bottom) + at_bottom


def maximum_absolute_diff(normalized_density, bottom, top):
    """:202-235: where the normalised density is farthest from its chord, and by how much."""
    from scipy.optimize import minimize_scalar

    result = minimize_scalar(
        lambda radius: -np.abs(normalized_density(radius) - straight_line(radius, normalized_density, bottom, top)),
        bounds=[bottom, top], method="bounded",
    )  # fmt: skip
    return result.x, -result.fun


def density_based_discretization_single(tesseroid, density):
    """:125-157: radial splits of one tesseroid until the density is close to linear in each."""
    w, e, s, n, bottom, top = tesseroid[:]
    density_min, density_max = density_minmax(density, bottom, top)
    if np.isclose(density_min, density_max):
        return [tesseroid]

    def normalized_density(radius):
        return (density(radius) - density_min) / (density_max - density_min)

    size_original = top - bottom
    pending, done = [tesseroid], []
    while pending:
        bottom, top = pending.pop(0)[-2:]
        radius_split, max_diff = maximum_absolute_diff(normalized_density, bottom, top)
        if max_diff * ((top - bottom) / size_original) > DELTA_RATIO:
            pending.append([w, e, s, n, radius_split, top])
            pending.append([w, e, s, n, bottom, radius_split])
        else:
            done.append([w, e, s, n, bottom, top])
    return done


def density_based_discretization(tesseroids, density):
    """:108-122."""
    out = []
    for tesseroid in tesseroids:
        out.extend(density_based_discretization_single(tesseroid, density))
    return np.atleast_2d(out)


def tesseroid_gravity_variable_density(coordinates, tesseroids, density, field,
                                       radial_adaptive_discretization=False):
    """Restatement of ``tesseroid_gravity`` for a callable ``density(radius)``
    (tesseroid_gravity.py:182-183, 342-445): density-based radial discretisation, then the pair
    loop with the density evaluated at every radial quadrature node (through a C callback)."""
    if field not in TESSEROID_RATII:
        raise ValueError(f"Gravitational field {field} not recognized")
    cast, (lon, lat, rad) = _coords(coordinates)
    tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=np.float64))
    if (tesseroids[:, 0] > tesseroids[:, 1]).any():
        tesseroids = longitude_continuity(tesseroids)
    tesseroids = _f64(density_based_discretization(tesseroids, density))
    out = np.zeros(lon.size, dtype=np.float64)
    callback = DENSITY_FN(lambda radius: float(density(radius)))
    status = lib().hbo_tesseroid_loop_variable_density(
        FIELD_IDS[field], lon.size, _p(lon), _p(lat), _p(rad), tesseroids.shape[0], _p(tesseroids),
        callback, TESSEROID_RATII[field], int(bool(radial_adaptive_discretization)), _p(out),
    )  # fmt: skip
    if status & 3:
        raise OverflowError("adaptive discretisation overflow")
    if status & 4:
        raise ZeroDivisionError("division by zero")
    if field == "g_z":
        out *= -1
        out *= 1e5
    return out.reshape(cast.shape)
