"""
``prism_magnetic``: drop-in for ``harmonica.prism_magnetic``.

Host logic restated from ``harmonica/_forward/prisms/magnetic.py:28-136,
403-457``; the loops ``_jit_prism_magnetic_field`` / ``_jit_prism_magnetic_component``
(:275-400) and choclo's ``magnetic_*`` kernels run in ``libharmonica_b200.so``.
"""

import ctypes

import numpy as np

from . import _lib
from ._utils import broadcast_coordinates, check_prisms, observer_chunks, progress

VALID_FIELDS = ("b", "b_e", "b_n", "b_u")
_COMPONENT_MASK = {"b": 7, "b_e": 1, "b_n": 2, "b_u": 4}


def _run_sanity_checks(prisms, magnetization):
    """magnetic.py:443-457."""
    if (size := len(magnetization)) != 3:
        raise ValueError(
            f"Invalid magnetization vectors with '{size}' elements. "
            + "Magnetization vectors should have only 3 elements."
        )
    if magnetization[0].size != prisms.shape[0]:
        raise ValueError(
            f"Number of magnetization vectors ({magnetization[0].size}) "
            + f"mismatch the number of prisms ({prisms.shape[0]})"
        )
    check_prisms(prisms)


def prism_magnetic(
    coordinates,
    prisms,
    magnetization,
    field,
    parallel=True,
    dtype=np.float64,
    progressbar=False,
    disable_checks=False,
    *,
    shard="auto",
    rules=_lib.MAG_DEFAULT_RULES,
):
    """
    Magnetic field (nT) of right-rectangular prisms in Cartesian coordinates.

    Same signature as ``harmonica.prism_magnetic``: ``magnetization`` is a tuple
    of three 1-D arrays (A/m), ``field`` one of ``"b"`` (returns the tuple
    ``(b_e, b_n, b_u)`` from one fused pass), ``"b_e"``, ``"b_n"``, ``"b_u"``.

    ``rules`` (extension) selects the singular-point behaviour of the kernels,
    which depends on the installed choclo version in the reference: bit 0 = NaN
    on prism edges and vertices, bit 1 = outside limit on east/north/top faces.
    """
    if field not in VALID_FIELDS:
        raise ValueError(
            f"Invalid field '{field}'. Please choose one of '{','.join(VALID_FIELDS)}'."
        )
    shape, coords = broadcast_coordinates(coordinates)
    prisms = np.atleast_2d(np.asarray(prisms, dtype=np.float64))
    magnetization = tuple(
        np.atleast_1d(np.asarray(m, dtype=np.float64)).ravel() for m in magnetization
    )
    if not disable_checks:
        _run_sanity_checks(prisms, magnetization)
    # magnetic.py:403-440: zero volume or all three components zero
    mag_e, mag_n, mag_u = magnetization
    null = (
        (prisms[:, 0] == prisms[:, 1])
        | (prisms[:, 2] == prisms[:, 3])
        | (prisms[:, 4] == prisms[:, 5])
        | ((mag_e == 0) & (mag_n == 0) & (mag_u == 0))
    )
    prisms = _lib.f64(prisms[~null])
    mag_e, mag_n, mag_u = (_lib.f64(m[~null]) for m in (mag_e, mag_n, mag_u))
    lib = _lib.ensure_init()
    mask = _COMPONENT_MASK[field]
    n_fields = 3 if field == "b" else 1
    n_obs = coords[0].size
    out = np.empty((n_fields, n_obs), dtype=np.float64)
    with progress(n_obs, progressbar) as proxy:
        for lo, hi in observer_chunks(n_obs, proxy):
            sub = tuple(np.ascontiguousarray(c[lo:hi]) for c in coords)
            res = np.empty((n_fields, hi - lo), dtype=np.float64)
            flags = ctypes.c_uint32(0)
            _lib.check(
                lib.hb200_prism_magnetic(
                    _lib.ptr(sub[0]), _lib.ptr(sub[1]), _lib.ptr(sub[2]), hi - lo,
                    _lib.ptr(prisms), _lib.ptr(mag_e), _lib.ptr(mag_n), _lib.ptr(mag_u),
                    prisms.shape[0], mask, int(rules), _lib.shard_mode(shard), _lib.ptr(res),
                    ctypes.byref(flags),
                )  # fmt: skip
            )
            out[:, lo:hi] = res
            if proxy is not None:
                proxy.update(hi - lo)
    if field == "b":
        return tuple(out[i].astype(dtype, copy=False).reshape(shape) for i in range(3))
    return out[0].astype(dtype, copy=False).reshape(shape)
