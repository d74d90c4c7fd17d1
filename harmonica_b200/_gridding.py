"""
Host-side coordinate helpers the equivalent-sources classes take from verde / bordado in the
reference. Neither package is installed here (nor vendored in ``/root/reference``), so their
published behaviour is restated and pinned on the values the reference's own tests hold:

* ``block_average_coordinates``: ``verde.BlockReduce(spacing, reduction=np.median,
  drop_coords=False).filter`` as called at ``_equivalent_sources/cartesian.py:345-351``
  (golden values: ``test/test_eq_sources_cartesian.py:216-254``);
* ``rolling_windows``: ``bordado.rolling_window(coordinates, region=, window_size=, overlap=)``
  as called at ``gradient_boosted.py:351-360`` (golden window populations:
  ``test/test_gradient_boosted_eqs.py:263-295``);
* ``shuffle_together``: ``sklearn.utils.shuffle(a, b, random_state=)`` (``:363-366``);
* ``neighbor_distance``: ``bordado.neighbor_distance_statistics(coordinates, "median", k=1)``
  (``cartesian.py:315-317``).

None of this is on the pairwise hot path; it is plain numpy / scipy.
"""

import numpy as np


def get_region(coordinates):
    """``bordado.get_region``: (W, E, S, N) of the first two coordinate arrays."""
    east, north = coordinates[:2]
    return (np.min(east), np.max(east), np.min(north), np.max(north))


def n_1d_arrays(arrays, n):
    """``verde.base.n_1d_arrays``: the first ``n`` arrays, raveled."""
    return tuple(np.atleast_1d(i).ravel() for i in arrays[:n])


def neighbor_distance(coordinates):
    """Distance of every point to its nearest neighbour (horizontal coordinates)."""
    from scipy.spatial import cKDTree  # noqa: PLC0415

    xy = np.transpose(n_1d_arrays(coordinates, 2))
    return cKDTree(xy).query(xy, k=2)[0][:, 1]


def _block_centres(start, stop, spacing):
    """Pixel-registered centres of ``round(length / spacing)`` equal blocks (spacing adjusted)."""
    n_blocks = max(int(round((stop - start) / spacing)), 1)
    width = (stop - start) / n_blocks
    return start + (np.arange(n_blocks) + 0.5) * width


def block_average_coordinates(coordinates, block_size):
    """
    Median of every coordinate array inside blocks of ``block_size`` (a number or
    ``(size_north, size_east)``); blocks span the data region with the spacing adjusted to fit
    it, points go to the nearest block centre (the lower one on a tie), blocks come out
    northing-major / easting-minor and empty blocks are dropped.
    """
    arrays = tuple(np.atleast_1d(np.asarray(c)).ravel() for c in coordinates)
    if np.ndim(block_size) == 0:
        size_north = size_east = float(block_size)
    else:
        size_north, size_east = (float(v) for v in block_size)
    west, east, south, north = get_region(arrays)
    centres_e = _block_centres(west, east, size_east)
    centres_n = _block_centres(south, north, size_north)
    # nearest centre per axis; a point midway between two centres belongs to the lower block
    col = np.searchsorted((centres_e[:-1] + centres_e[1:]) / 2, arrays[0], side="left")
    row = np.searchsorted((centres_n[:-1] + centres_n[1:]) / 2, arrays[1], side="left")
    label = row * centres_e.size + col
    order = np.argsort(label, kind="stable")
    sorted_labels = label[order]
    starts = np.flatnonzero(np.r_[True, sorted_labels[1:] != sorted_labels[:-1]])
    groups = np.split(order, starts[1:])
    return tuple(np.array([np.median(a[g]) for g in groups]) for a in arrays)


def _window_centres(start, stop, spacing):
    """Centres on [start, stop] with the spacing adjusted so that both ends are centres."""
    length = stop - start
    if length <= 0:
        return np.array([start], dtype=float)
    n = int(round(length / spacing)) + 1
    n = max(n, 2)
    return np.linspace(start, stop, n)


def rolling_windows(coordinates, region, window_size, overlap):
    """
    Indices of the points inside every square window of ``window_size`` rolled over ``region``
    with the given fractional ``overlap`` (adjusted so that the windows fit the region exactly).
    Windows are closed (points on the edge belong to the window) and come out northing-major /
    easting-minor; each entry is a 1-D integer array, possibly empty.
    """
    from scipy.spatial import cKDTree  # noqa: PLC0415

    east, north = n_1d_arrays(coordinates, 2)
    west, east_b, south, north_b = (float(v) for v in region)
    if min(east_b - west, north_b - south) < window_size:
        raise ValueError(
            f"Window size '{window_size}' is larger than dimensions of the region "
            f"'{(west, east_b, south, north_b)}'."
        )
    step = (1 - overlap) * window_size
    centres_e = _window_centres(west + window_size / 2, east_b - window_size / 2, step)
    centres_n = _window_centres(south + window_size / 2, north_b - window_size / 2, step)
    grid_e, grid_n = np.meshgrid(centres_e, centres_n)
    tree = cKDTree(np.transpose([east, north]))
    found = tree.query_ball_point(
        np.transpose([grid_e.ravel(), grid_n.ravel()]), r=window_size / 2, p=np.inf
    )
    return [np.sort(np.asarray(idx, dtype=np.int64)) for idx in found]


def shuffle_together(first, second, random_state=None):
    """``sklearn.utils.shuffle`` of two equally long lists with one permutation."""
    if random_state is None or isinstance(random_state, (int, np.integer)):
        rng = np.random.RandomState(random_state)
    else:
        rng = random_state
    indices = np.arange(len(first))
    rng.shuffle(indices)
    return [first[i] for i in indices], [second[i] for i in indices]
