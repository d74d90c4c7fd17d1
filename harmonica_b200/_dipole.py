"""
``dipole_magnetic``: drop-in for ``harmonica.dipole_magnetic``.

Host logic restated from ``harmonica/_forward/dipole.py:27-137, 274-289``; the loops
``_jit_dipole_magnetic_field_cartesian`` / ``_jit_dipole_magnetic_component_cartesian``
(:292-415) and choclo's dipole kernels run in ``libharmonica_b200.so``.
"""

import ctypes

import numpy as np

from . import _lib
from ._utils import broadcast_coordinates, observer_chunks, progress

VALID_FIELDS = ("b", "b_e", "b_n", "b_u")
_COMPONENT_MASK = {"b": 7, "b_e": 1, "b_n": 2, "b_u": 4}


def _check_dipoles_and_magnetic_moments(dipoles, magnetic_moments):
    """dipole.py:274-289."""
    if (size := len(magnetic_moments)) != 3:
        raise ValueError(
            f"Invalid magnetic moments with '{size}' elements."
            " Magnetic moments vectors should have 3 components."
        )
    if magnetic_moments[0].size != dipoles[0].size:
        raise ValueError(
            f"Number of elements in magnetic_moments ({magnetic_moments[0].size})"
            f" mismatch the number of dipoles ({dipoles[0].size})."
        )


def dipole_magnetic(
    coordinates,
    dipoles,
    magnetic_moments,
    field,
    parallel=True,
    dtype="float64",
    progressbar=False,
    disable_checks=False,
    *,
    shard="auto",
):
    """
    Magnetic field (nT) of dipoles in Cartesian coordinates.

    Same signature as ``harmonica.dipole_magnetic``: ``dipoles`` is a tuple of the three
    coordinate arrays, ``magnetic_moments`` a tuple of the three moment components (A m^2),
    ``field`` one of ``"b"`` (tuple ``(b_e, b_n, b_u)`` from one fused pass), ``"b_e"``,
    ``"b_n"``, ``"b_u"``. An observation point that coincides with a dipole raises
    ``ZeroDivisionError`` like the reference's jitted loop.
    """
    if field not in VALID_FIELDS:
        raise ValueError(
            f"Invalid field '{field}'. Please choose one of '{', '.join(VALID_FIELDS)}'."
        )
    shape, coords = broadcast_coordinates(coordinates)
    dipoles = tuple(_lib.f64(np.atleast_1d(i).ravel()) for i in dipoles[:3])
    magnetic_moments = tuple(_lib.f64(np.atleast_1d(m).ravel()) for m in magnetic_moments)
    if not disable_checks:
        _check_dipoles_and_magnetic_moments(dipoles, magnetic_moments)
    lib = _lib.ensure_init()
    n_fields = 3 if field == "b" else 1
    n_obs = coords[0].size
    out = np.empty((n_fields, n_obs), dtype=np.float64)
    zero_div = False
    with progress(n_obs, progressbar) as proxy:
        for lo, hi in observer_chunks(n_obs, proxy):
            sub = tuple(np.ascontiguousarray(c[lo:hi]) for c in coords)
            res = np.empty((n_fields, hi - lo), dtype=np.float64)
            flags = ctypes.c_uint32(0)
            _lib.check(
                lib.hb200_dipole_magnetic(
                    _lib.ptr(sub[0]), _lib.ptr(sub[1]), _lib.ptr(sub[2]), hi - lo,
                    _lib.ptr(dipoles[0]), _lib.ptr(dipoles[1]), _lib.ptr(dipoles[2]),
                    _lib.ptr(magnetic_moments[0]), _lib.ptr(magnetic_moments[1]),
                    _lib.ptr(magnetic_moments[2]), dipoles[0].size, _COMPONENT_MASK[field],
                    _lib.shard_mode(shard), _lib.ptr(res), ctypes.byref(flags),
                )  # fmt: skip
            )
            out[:, lo:hi] = res
            zero_div = zero_div or bool(flags.value & _lib.FLAG_ZERO_DIV)
            if proxy is not None:
                proxy.update(hi - lo)
    if zero_div:
        raise ZeroDivisionError("division by zero")
    if field == "b":
        return tuple(out[i].astype(dtype, copy=False).reshape(shape) for i in range(3))
    return out[0].astype(dtype, copy=False).reshape(shape)
