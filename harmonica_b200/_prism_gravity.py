"""
``prism_gravity``: drop-in for ``harmonica.prism_gravity``.

Host logic restated from ``harmonica/_forward/prisms/gravity.py:51-236, 239-269,
452-486`` (validation order and messages, null-prism discard, singular-point
warning, shape/dtype of the result); the pair loop ``jit_prism_gravity``
(:489-545) and the choclo kernels run in ``libharmonica_b200.so``.
"""

import ctypes
import warnings

import numpy as np

from . import _lib
from ._utils import (broadcast_coordinates, cartesian_locality_order, check_prisms, observer_chunks,
                     progress)

#: available fields (same keys as the reference's ``FIELDS``, gravity.py:37-48)
FIELDS = tuple(_lib.FIELD_IDS)
TENSOR_FIELDS = ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz")


def _discard_null_prisms(prisms, density):
    """Drop zero-volume and zero-density prisms (gravity.py:452-486)."""
    null = (
        (prisms[:, 0] == prisms[:, 1])
        | (prisms[:, 2] == prisms[:, 3])
        | (prisms[:, 4] == prisms[:, 5])
        | (density == 0)
    )
    return prisms[~null], density[~null], prisms[null]


def _run(lib, coords, prisms, density, mask, shard, n_fields, progress_proxy=None):
    """Call hb200_prism_gravity (in observer chunks when a progress bar is shown)."""
    n_obs = coords[0].size
    out = np.empty((n_fields, n_obs), dtype=np.float64)
    flags_all = 0
    # the C entry point takes at most 6 fields per call: potential + accelerations, then tensor
    masks = [mask] if n_fields <= 6 else [mask & 0x00F, mask & _lib.MASK_TENSOR]
    for lo, hi in observer_chunks(n_obs, progress_proxy):
        sub = tuple(np.ascontiguousarray(c[lo:hi]) for c in coords)
        row = 0
        for part in masks:
            n_part = bin(part).count("1")
            res = np.empty((n_part, hi - lo), dtype=np.float64)
            flags = ctypes.c_uint32(0)
            _lib.check(
                lib.hb200_prism_gravity(
                    _lib.ptr(sub[0]), _lib.ptr(sub[1]), _lib.ptr(sub[2]), hi - lo,
                    _lib.ptr(prisms), _lib.ptr(density), prisms.shape[0], part, shard,
                    _lib.ptr(res), ctypes.byref(flags),
                )  # fmt: skip
            )
            out[row:row + n_part, lo:hi] = res
            row += n_part
            flags_all |= flags.value
        if progress_proxy is not None:
            progress_proxy.update(hi - lo)
    return out, flags_all


def prism_gravity(
    coordinates,
    prisms,
    density,
    field,
    parallel=True,
    dtype="float64",
    progressbar=False,
    disable_checks=False,
    *,
    shard="auto",
):
    """
    Gravitational fields of right-rectangular prisms in Cartesian coordinates.

    Same signature, units and sign conventions as ``harmonica.prism_gravity``:
    ``potential`` in J/kg, ``g_e``/``g_n``/``g_z`` in mGal (``g_z`` positive
    downward), tensor components in Eotvos (``g_ez``, ``g_nz`` with z down).
    Tensor components are NaN on their singular points and take the outside
    limit on faces normal to a diagonal component.

    ``parallel`` is accepted for compatibility (the GPU path is always
    parallel). ``shard`` ("auto", "observers", "sources") is an extension that
    selects how the work is split over the visible B200s.

    ``field`` may also be a tuple/list of field names: they are computed in one
    fused pass and returned as a tuple (extension; e.g. all six tensor
    components share one pass over the prisms).
    """
    multi = not isinstance(field, str)
    fields = tuple(field) if multi else (field,)
    for f in fields:
        if f not in _lib.FIELD_IDS:
            raise ValueError(f"Gravitational field {f} not recognized")
    shape, coords = broadcast_coordinates(coordinates)
    prisms = np.atleast_2d(np.asarray(prisms, dtype=np.float64))
    density = np.atleast_1d(np.asarray(density, dtype=np.float64)).ravel()
    shard_mode = _lib.shard_mode(shard)
    if not disable_checks:
        if density.size != prisms.shape[0]:
            raise ValueError(
                f"Number of elements in density ({density.size}) "
                + f"mismatch the number of prisms ({prisms.shape[0]})"
            )
        check_prisms(prisms)
    prisms, density, null_prisms = _discard_null_prisms(prisms, density)
    prisms = _lib.f64(prisms)
    density = _lib.f64(density)
    mask = 0
    for f in fields:
        mask |= 1 << _lib.FIELD_IDS[f]
    order = sorted(fields, key=lambda f: _lib.FIELD_IDS[f])
    if len(set(fields)) != len(fields):
        raise ValueError("Repeated fields")
    lib = _lib.ensure_init()
    # the potential and the accelerations spend most of their time in log / atan sequences whose
    # length depends on the distance: scattered observers are handed over in a local order
    # (measured on B200, 20 000 prisms x 262 144 random observers: potential 20 -> 27 G pair/s,
    # fused accelerations 24 -> 29 G; the tensor kernels gain nothing and would pay the sort)
    perm = None
    if any(f in ("potential", "g_e", "g_n", "g_z") for f in fields) and prisms.shape[0] >= 4096:
        perm = cartesian_locality_order(coords[0], coords[1], prisms.shape[0])
    if perm is not None:  # neighbouring observers share a warp; results do not depend on the order
        coords = tuple(np.ascontiguousarray(c[perm]) for c in coords)
    with progress(coords[0].size, progressbar) as proxy:
        out, flags = _run(lib, coords, prisms, density, mask, shard_mode, len(fields), proxy)
    if perm is not None:
        unsorted = np.empty_like(out)
        unsorted[:, perm] = out
        out = unsorted
    # gravity.py:239-269: the reference scans ALL prisms (before the null
    # discard) for the tensor fields unless checks are disabled
    if not disable_checks and any(f in TENSOR_FIELDS for f in fields):
        singular = bool(flags & _lib.FLAG_SINGULAR)
        if not singular and null_prisms.shape[0] > 0:
            null_prisms = _lib.f64(null_prisms)
            for f in fields:
                if f not in TENSOR_FIELDS:
                    continue
                fl = ctypes.c_uint32(0)
                _lib.check(
                    lib.hb200_prism_singular_scan(
                        _lib.ptr(coords[0]), _lib.ptr(coords[1]), _lib.ptr(coords[2]),
                        coords[0].size, _lib.ptr(null_prisms), null_prisms.shape[0],
                        _lib.FIELD_IDS[f], ctypes.byref(fl),
                    )  # fmt: skip
                )
                singular = singular or bool(fl.value & _lib.FLAG_SINGULAR)
        if singular:
            warnings.warn(
                "Found observation point on singular point of a prism.",
                UserWarning,
                stacklevel=2,
            )
    results = {f: out[i].astype(dtype, copy=False).reshape(shape) for i, f in enumerate(order)}
    if multi:
        return tuple(results[f] for f in fields)
    return results[field]
