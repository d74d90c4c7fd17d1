"""
``point_gravity``: drop-in for ``harmonica.point_gravity``.

Host logic restated from ``harmonica/_forward/point.py:30-316``; the loops
``point_mass_cartesian`` / ``point_mass_spherical`` (:357-454) and the choclo
point kernels run in ``libharmonica_b200.so``.
"""

import ctypes

import numpy as np

from . import _lib
from ._utils import broadcast_coordinates, check_coordinate_system

_ALIASES = {"g_ne": "g_en", "g_ze": "g_ez", "g_zn": "g_nz"}
_SPHERICAL = {"potential": True, "g_z": True, "g_n": False, "g_e": False}


def _field_id(coordinate_system, field):
    """point.py:281-316 (``get_kernel``): same errors for unknown / unimplemented fields."""
    if coordinate_system == "cartesian":
        base = _ALIASES.get(field, field)
        if base not in _lib.FIELD_IDS:
            raise ValueError(f"Gravitational field '{field}' not recognized")
        return _lib.FIELD_IDS[base]
    if field not in _SPHERICAL:
        raise ValueError(f"Gravitational field '{field}' not recognized")
    if not _SPHERICAL[field]:
        raise NotImplementedError
    return _lib.FIELD_IDS[field]


def point_gravity(
    coordinates,
    points,
    masses,
    field,
    coordinate_system="cartesian",
    parallel=True,
    dtype="float64",
    *,
    shard="auto",
):
    """
    Gravitational fields of point masses (Cartesian or geocentric spherical).

    Same signature, units and signs as ``harmonica.point_gravity``. A
    computation point that coincides with a point mass raises
    ``ZeroDivisionError`` like the reference's jitted loop does.
    """
    check_coordinate_system(coordinate_system, valid_coord_systems=("cartesian", "spherical"))
    shape, coords = broadcast_coordinates(coordinates)
    points = tuple(_lib.f64(np.atleast_1d(p).ravel()) for p in points[:3])
    masses = _lib.f64(np.atleast_1d(masses).ravel())
    if masses.size != points[0].size:
        raise ValueError(
            f"Number of elements in masses ({masses.size}) "
            + f"mismatch the number of points ({points[0].size})"
        )
    field_id = _field_id(coordinate_system, field)
    lib = _lib.ensure_init()
    out = np.empty(coords[0].size, dtype=np.float64)
    flags = ctypes.c_uint32(0)
    _lib.check(
        lib.hb200_point_gravity(
            _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]), coords[0].size,
            _lib.ptr(points[0]), _lib.ptr(points[1]), _lib.ptr(points[2]), _lib.ptr(masses),
            masses.size, 1 << field_id, int(coordinate_system == "spherical"),
            _lib.shard_mode(shard), _lib.ptr(out), ctypes.byref(flags),
        )  # fmt: skip
    )
    if flags.value & _lib.FLAG_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    return out.astype(dtype, copy=False).reshape(shape)
