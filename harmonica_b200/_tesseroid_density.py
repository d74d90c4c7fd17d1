"""
Tesseroids whose density is a function of the radius: the host side of
``harmonica/_forward/_tesseroid_variable_density.py``.

The reference splits every tesseroid radially until its density function is close to linear in
each piece (``density_based_discretization``, :108-157) and then evaluates ``density(radius_p)`` at
the radial Gauss-Legendre nodes of every leaf of the adaptive discretisation (:55-58). With the
default horizontal discretisation a leaf keeps the radial bounds of its tesseroid, so those nodes
are the SAME two radii for all leaves of one tesseroid: the density function is evaluated here,
on the host, at the two nodes of every (radially split) tesseroid, and the CUDA kernel receives
two densities per tesseroid (``hb200_tesseroid_gravity_variable_density``). The function can be
any Python callable of a scalar radius (numba-jitted ones included).
"""

import numpy as np

DELTA_RATIO = 0.1  # _tesseroid_variable_density.py:17
GLQ_NODE = 0.5773502691896257  # numpy.polynomial.legendre.leggauss(2)


def _bounded_minimum(function, bottom, top):
    from scipy.optimize import minimize_scalar  # noqa: PLC0415

    return minimize_scalar(function, bounds=[bottom, top], method="bounded")


def density_minmax(density, bottom, top):
    """Smallest and largest density between ``bottom`` and ``top`` (:159-199): bounded scalar
    searches for an interior extremum, compared with the values at the two ends."""
    low_end, high_end = np.sort([density(bottom), density(top)])
    interior_min = _bounded_minimum(density, bottom, top).fun
    interior_max = -_bounded_minimum(lambda radius: -density(radius), bottom, top).fun
    return np.min((interior_min, low_end)), np.max((interior_max, high_end))


def straight_line(radius, normalized_density, bottom, top):
    """Chord of the normalised density through its values at ``bottom`` and ``top`` (:238-259)."""
    value_bottom = normalized_density(bottom)
    value_top = normalized_density(top)
    slope = (value_top - value_bottom) / (top - bottom)
    return slope * (radius - bottom) + value_bottom


def maximum_absolute_diff(normalized_density, bottom, top):
    """Radius at which the normalised density departs most from its chord, and that departure
    (:202-235)."""
    def negative_departure(radius):
        return -np.abs(normalized_density(radius) - straight_line(radius, normalized_density, bottom, top))

    found = _bounded_minimum(negative_departure, bottom, top)
    return found.x, -found.fun


def _density_based_discretization(tesseroid, density):
    """Radial pieces of one tesseroid (:125-157), in the reference's (breadth-first) order."""
    w, e, s, n, bottom, top = tesseroid[:]
    density_min, density_max = density_minmax(density, bottom, top)
    if np.isclose(density_min, density_max):
        return [tesseroid]

    def normalized_density(radius):
        return (density(radius) - density_min) / (density_max - density_min)

    full_size = top - bottom
    queue, pieces = [tesseroid], []
    while queue:
        bottom, top = queue.pop(0)[-2:]
        radius_split, max_diff = maximum_absolute_diff(normalized_density, bottom, top)
        if max_diff * (top - bottom) / full_size > DELTA_RATIO:
            queue.append([w, e, s, n, radius_split, top])
            queue.append([w, e, s, n, bottom, radius_split])
        else:
            pieces.append([w, e, s, n, bottom, top])
    return pieces


def density_based_discretization(tesseroids, density):
    """All tesseroids split radially according to ``density`` (:108-122)."""
    pieces = []
    for tesseroid in tesseroids:
        pieces.extend(_density_based_discretization(tesseroid, density))
    return np.atleast_2d(pieces)


def _evaluate(density, radii):
    """``density`` at every radius: one vectorised call when the function accepts arrays (checked
    against scalar calls at both ends), otherwise one call per radius (numba-jitted scalar
    functions, functions with branches)."""
    if radii.size > 64:
        try:
            values = np.asarray(density(radii), dtype=np.float64)
            if (values.shape == radii.shape and values[0] == density(float(radii[0]))
                    and values[-1] == density(float(radii[-1]))):  # fmt: skip
                return values
        except Exception:  # noqa: BLE001, S110 - any failure means "not vectorisable"
            pass
    return np.array([density(r) for r in radii.tolist()], dtype=np.float64)


def density_at_radial_nodes(tesseroids, density):
    """``density(radius_p)`` at the two radial Gauss-Legendre nodes of every tesseroid
    (``radius_p`` as in :55-57)."""
    bottom, top = tesseroids[:, 4], tesseroids[:, 5]
    values = []
    for node in (-GLQ_NODE, GLQ_NODE):
        radius_p = 0.5 * (top - bottom) * node + 0.5 * (top + bottom)
        values.append(_evaluate(density, radius_p))
    return values[0], values[1]
