"""Host-side checks shared by the forward-modelling wrappers."""

import contextlib

import numpy as np


def check_prisms(prisms):
    """
    Raise ``ValueError`` if any prism has inverted boundaries.

    Same contract and messages as the reference's
    ``harmonica/_forward/prisms/utils.py:12-43`` (zero-volume prisms pass).
    """
    columns = ("west", "east"), ("south", "north"), ("bottom", "top")
    labels = ("prism", "prism", "tesseroid")  # the reference's wording, kept verbatim
    for (low_name, high_name), (low, high), label in zip(
        columns, ((0, 1), (2, 3), (4, 5)), labels
    ):
        bad = prisms[:, low] > prisms[:, high]
        if bad.any():
            msg = (
                "Invalid prism or prisms. "
                f"The {low_name} boundary can't be greater than the {high_name} one.\n"
            )
            msg += "".join(f"\tInvalid {label}: {p}\n" for p in prisms[bad])
            raise ValueError(msg)


def check_coordinate_system(
    coordinate_system, valid_coord_systems=("cartesian", "spherical", "geodetic")
):
    """``harmonica/_forward/utils.py:71-87``."""
    if coordinate_system not in valid_coord_systems:
        raise ValueError(f"Coordinate system {coordinate_system} not recognized.")


def broadcast_coordinates(coordinates):
    """
    Shape of the result and the three raveled float64 coordinate arrays.

    The reference ravels each array separately (``gravity.py:200-203``), which
    only works when all three have the same size; like it, only the first
    three entries are used.
    """
    cast = np.broadcast(*coordinates[:3])
    arrays = tuple(
        np.ascontiguousarray(np.atleast_1d(np.asarray(c, dtype=np.float64)).ravel())
        for c in coordinates[:3]
    )
    sizes = {a.size for a in arrays}
    if len(sizes) != 1:
        # the reference's jitted loop would index out of bounds here; broadcast instead
        arrays = tuple(
            np.ascontiguousarray(
                np.broadcast_to(np.asarray(c, dtype=np.float64), cast.shape).ravel()
            )
            for c in coordinates[:3]
        )
    return cast.shape, arrays


class _Progress:
    """Minimal stand-in for ``numba_progress.ProgressBar`` (``update(n)``)."""

    def __init__(self, total):
        from tqdm import tqdm  # noqa: PLC0415

        self._bar = tqdm(total=total)

    def update(self, n):
        self._bar.update(n)

    def close(self):
        self._bar.close()


@contextlib.contextmanager
def progress(total, use_progressbar):
    """
    Context manager yielding a progress proxy or None.

    The reference (``_forward/utils.py:333-392``) needs ``numba_progress`` and
    updates once per observer from inside the jitted loop; here the GPU call is
    split in observer chunks and the bar advances once per chunk.
    """
    if not use_progressbar:
        yield None
        return
    try:
        bar = _Progress(total)
    except ModuleNotFoundError as original:  # pragma: no cover
        raise ImportError(
            "Cannot import the optional dependency 'tqdm'. "
            "It must be installed to be able to show a progressbar."
        ) from original
    try:
        yield bar
    finally:
        bar.close()


def observer_chunks(n_obs, progress_proxy, n_chunks=20):
    """Observer index ranges: one range without a progress bar, ~20 with."""
    if progress_proxy is None or n_obs < 2 * n_chunks:
        return [(0, n_obs)]
    edges = np.linspace(0, n_obs, n_chunks + 1).astype(np.int64)
    return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def _spread_bits8():
    v = np.arange(256, dtype=np.uint16)
    v = (v | (v << np.uint16(4))) & np.uint16(0x0F0F)
    v = (v | (v << np.uint16(2))) & np.uint16(0x3333)
    v = (v | (v << np.uint16(1))) & np.uint16(0x5555)
    return v


_SPREAD8 = _spread_bits8()


def cartesian_locality_order(easting, northing, n_sources):
    """
    Permutation that puts observation points that are close in the horizontal plane next to each
    other (Morton order, 8 bits per axis over the bounding box; 16-bit keys sort in linear time),
    or None when it would not pay: few pairs, or consecutive points are neighbours already (grids,
    profiles).

    The prism kernels pick the length of their log / atan sequences per observer from the
    observer's distance to the prism; the 32 observers of a warp that lie next to each other pick
    the same one, scattered ones make the warp run through several (measured on B200: 22 of 32
    lanes active for random observers above a 100 km model). The value of every (observer, prism)
    pair depends on that pair only, so the order of the observers changes no result.
    """
    n = easting.size
    if n < 4096 or float(n) * n_sources < 5e8:
        return None
    with np.errstate(invalid="ignore"):
        lo_e, hi_e = np.nanmin(easting), np.nanmax(easting)
        lo_n, hi_n = np.nanmin(northing), np.nanmax(northing)
    extent = (hi_e - lo_e) + (hi_n - lo_n)
    if not np.isfinite(extent) or extent <= 0:
        return None
    starts = np.linspace(0, n - 64, 64).astype(np.int64)
    index = (starts[:, None] + np.arange(64)[None, :]).ravel()
    jump = np.abs(np.diff(easting[index].reshape(64, 64), axis=1)) + np.abs(
        np.diff(northing[index].reshape(64, 64), axis=1)
    )
    if np.nanmean(jump) < 0.05 * extent:  # random points jump by ~ extent / 3
        return None
    with np.errstate(invalid="ignore"):
        qx = ((easting - lo_e) * (255.999 / max(hi_e - lo_e, 1e-300))).astype(np.int32) & 255
        qy = ((northing - lo_n) * (255.999 / max(hi_n - lo_n, 1e-300))).astype(np.int32) & 255
    key = _SPREAD8[qx] | (_SPREAD8[qy] << np.uint16(1))
    return np.argsort(key, kind="stable")
