"""
oracle/numba_loops.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Numba ``parallel=True`` restatement of the reference's jitted double loops, over
the restated choclo kernels of ``oracle/choclo_numba.py``. This is the CPU arm
``bench.py`` times as ``cpu_baseline.kind == "numba"`` / ``--impl reference`` on
the GPU box, where ``/root/reference`` does not exist (in the build container
the reference's UNMODIFIED loops run through ``oracle/ref_shim.py`` and
``tests/test_oracle_golden.py`` asserts that these loops give bit-identical
results). Same technology as the reference (Numba, ``prange`` over the
observers, serial over the sources, no fastmath), same loop shape:

  prism_gravity_loop   src/harmonica/_forward/prisms/gravity.py:489-545
  prism_layer_loop     src/harmonica/_forward/prisms/layer.py:522-633
  prism_magnetic_loop  src/harmonica/_forward/prisms/magnetic.py:275-335
  eqs_predict_loop     src/harmonica/_equivalent_sources/utils.py:77-101 with
                       greens_func_cartesian (cartesian.py:634-644)

Only ``bench.py`` (CPU legs) and ``tests/`` import this module.
"""

import numpy as np
from numba import jit, prange

import choclo_numba as C

_GRAVITY = {
    "potential": C.gravity_pot, "g_e": C.gravity_e, "g_n": C.gravity_n, "g_z": C.gravity_u,
    "g_ee": C.gravity_ee, "g_nn": C.gravity_nn, "g_zz": C.gravity_uu, "g_en": C.gravity_en,
    "g_ez": C.gravity_eu, "g_nz": C.gravity_nu,
}  # fmt: skip
_LOOPS = {}


def _prism_loop_for(field):
    """One jitted loop per field (the kernel is a compile-time constant, like the
    reference's dispatch through a first-class function argument)."""
    if ("prism", field) in _LOOPS:
        return _LOOPS["prism", field]
    forward = _GRAVITY[field]

    @jit(nopython=True, parallel=True)
    def loop(easting, northing, upward, prisms, density, out):
        for i in prange(easting.size):
            for j in range(prisms.shape[0]):
                out[i] += forward(
                    easting[i], northing[i], upward[i], prisms[j, 0], prisms[j, 1], prisms[j, 2],
                    prisms[j, 3], prisms[j, 4], prisms[j, 5], density[j])  # fmt: skip

    _LOOPS["prism", field] = loop
    return loop


def _layer_loop_for(field):
    if ("layer", field) in _LOOPS:
        return _LOOPS["layer", field]
    forward = _GRAVITY[field]

    @jit(nopython=True, parallel=True)
    def loop(easting, northing, upward, prisms_easting, prisms_northing, top, bottom, density,
             thickness_threshold, out):
        half_e = (prisms_easting[1] - prisms_easting[0]) / 2
        half_n = (prisms_northing[1] - prisms_northing[0]) / 2
        for i in prange(easting.size):
            # easting outer, northing inner; skip rules in the reference's order
            for j in range(prisms_easting.size):
                west, east = prisms_easting[j] - half_e, prisms_easting[j] + half_e
                for k in range(prisms_northing.size):
                    rho = density[k, j]
                    if rho == 0 or np.isnan(rho):
                        continue
                    b, t = bottom[k, j], top[k, j]
                    if t - b < thickness_threshold:
                        continue
                    if np.isnan(t) or np.isnan(b):
                        continue
                    out[i] += forward(
                        easting[i], northing[i], upward[i], west, east,
                        prisms_northing[k] - half_n, prisms_northing[k] + half_n, b, t, rho)  # fmt: skip

    _LOOPS["layer", field] = loop
    return loop


@jit(nopython=True, parallel=True)
def _magnetic_field_loop(easting, northing, upward, prisms, me, mn, mu, be, bn, bu):
    for i in prange(easting.size):
        for j in range(prisms.shape[0]):
            e, n, u = C.magnetic_field(
                easting[i], northing[i], upward[i], prisms[j, 0], prisms[j, 1], prisms[j, 2],
                prisms[j, 3], prisms[j, 4], prisms[j, 5], me[j], mn[j], mu[j])  # fmt: skip
            be[i] += e
            bn[i] += n
            bu[i] += u


@jit(nopython=True, parallel=True)
def _eqs_predict_loop(easting, northing, upward, pe, pn, pu, coefs, out):
    for i in prange(easting.size):
        for j in range(pe.size):
            de, dn, du = easting[i] - pe[j], northing[i] - pn[j], upward[i] - pu[j]
            out[i] += coefs[j] * (1 / np.sqrt(de * de + dn * dn + du * du))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def prism_gravity(coordinates, prisms, density, field):
    """SI, choclo's sign (upward positive): the loop only, like ``jit_prism_gravity``."""
    e, n, u = (_c(c).ravel() for c in coordinates[:3])
    out = np.zeros(e.size)
    _prism_loop_for(field)(e, n, u, _c(prisms), _c(density), out)
    return out


def prism_layer_gravity(coordinates, prisms_easting, prisms_northing, bottom, top, density, field,
                        thickness_threshold=0.0):
    e, n, u = (_c(c).ravel() for c in coordinates[:3])
    out = np.zeros(e.size)
    _layer_loop_for(field)(e, n, u, _c(prisms_easting), _c(prisms_northing), _c(top), _c(bottom),
                           _c(density), float(thickness_threshold), out)
    return out


def prism_magnetic_field(coordinates, prisms, magnetization):
    e, n, u = (_c(c).ravel() for c in coordinates[:3])
    be, bn, bu = np.zeros(e.size), np.zeros(e.size), np.zeros(e.size)
    me, mn, mu = (_c(m) for m in magnetization)
    _magnetic_field_loop(e, n, u, _c(prisms), me, mn, mu, be, bn, bu)
    return be, bn, bu


def eqs_predict(coordinates, points, coefs):
    e, n, u = (_c(c).ravel() for c in coordinates[:3])
    out = np.zeros(e.size)
    _eqs_predict_loop(e, n, u, _c(points[0]), _c(points[1]), _c(points[2]), _c(coefs), out)
    return out


def num_threads():
    import numba  # noqa: PLC0415

    return int(numba.get_num_threads())
