"""
oracle/choclo_numba.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Numba restatement of the `choclo` functions the reference imports on the hot
path (choclo is a third-party dependency pinned only as ``choclo >= 0.1`` in
/root/reference/pyproject.toml:41; it is not vendored and not installed here).
It exists for ONE purpose: to be injected as a fake ``choclo`` package by
``oracle/ref_shim.py`` so that the reference's UNMODIFIED wrappers and jitted
loops (src/harmonica/_forward/prisms/{gravity,magnetic,layer}.py,
src/harmonica/_forward/point.py) can be executed in the build container, to
validate ``oracle/choclo_port.c`` and to generate ``tests/golden/*.npz``.

The arithmetic is the same statement-for-statement as oracle/choclo_port.c
(same operation order, no fastmath), so the two agree to the last bit wherever
libm's log/atan agree.

Import sites this mirrors (reference file:line):
  prisms/gravity.py:14-30   choclo.prism.gravity_*, choclo.prism._utils.is_point_on_*_edge
  prisms/magnetic.py:14     choclo.prism.magnetic_{e,n,u,field}
  point.py:12-24            choclo.constants.GRAVITATIONAL_CONST, choclo.point.gravity_*
"""

import sys
import types

import numpy as np
from numba import jit

GRAVITATIONAL_CONST = 6.6743e-11
VACUUM_MAGNETIC_PERMEABILITY = 4 * np.pi * 1e-7


@jit(nopython=True)
def safe_atan2(y, x):
    if x != 0:
        return np.arctan(y / x)
    if y > 0:
        return np.pi / 2
    if y < 0:
        return -np.pi / 2
    return 0.0


@jit(nopython=True)
def safe_log(x, y, z, r):
    if r == 0:
        return 0.0
    if x < 0:
        if r == -x:
            return -np.log(-2 * x)
        return np.log((y * y + z * z) / (r - x))
    return np.log(x + r)


@jit(nopython=True)
def kernel_pot(e, n, u, r):
    return (
        e * n * safe_log(u, e, n, r)
        + n * u * safe_log(e, n, u, r)
        + e * u * safe_log(n, e, u, r)
        - 0.5 * (e * e) * safe_atan2(u * n, e * r)
        - 0.5 * (n * n) * safe_atan2(u * e, n * r)
        - 0.5 * (u * u) * safe_atan2(e * n, u * r)
    )


@jit(nopython=True)
def kernel_e(e, n, u, r):
    return -(n * safe_log(u, e, n, r) + u * safe_log(n, e, u, r) - e * safe_atan2(n * u, e * r))


@jit(nopython=True)
def kernel_n(e, n, u, r):
    return -(u * safe_log(e, n, u, r) + e * safe_log(u, e, n, r) - n * safe_atan2(u * e, n * r))


@jit(nopython=True)
def kernel_u(e, n, u, r):
    return -(e * safe_log(n, e, u, r) + n * safe_log(e, n, u, r) - u * safe_atan2(e * n, u * r))


@jit(nopython=True)
def kernel_ee(e, n, u, r):
    return -safe_atan2(n * u, e * r)


@jit(nopython=True)
def kernel_nn(e, n, u, r):
    return -safe_atan2(e * u, n * r)


@jit(nopython=True)
def kernel_uu(e, n, u, r):
    return -safe_atan2(e * n, u * r)


@jit(nopython=True)
def kernel_en(e, n, u, r):
    return safe_log(u, e, n, r)


@jit(nopython=True)
def kernel_eu(e, n, u, r):
    return safe_log(n, e, u, r)


@jit(nopython=True)
def kernel_nu(e, n, u, r):
    return safe_log(e, n, u, r)


@jit(nopython=True)
def _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel):
    result = 0.0
    for i in range(2):
        se = (e if i == 0 else w) - E
        se2 = se * se
        for j in range(2):
            sn = (n if j == 0 else s) - N
            sn2 = sn * sn
            for k in range(2):
                su = (t if k == 0 else b) - U
                su2 = su * su
                r = np.sqrt(se2 + sn2 + su2)
                sign = -1.0 if (i + j + k) % 2 else 1.0
                result += sign * kernel(se, sn, su, r)
    return result


@jit(nopython=True)
def is_point_on_easting_edge(E, N, U, w, e, s, n, b, t):
    return (w <= E <= e) and (N == s or N == n) and (U == b or U == t)


@jit(nopython=True)
def is_point_on_northing_edge(E, N, U, w, e, s, n, b, t):
    return (s <= N <= n) and (E == w or E == e) and (U == b or U == t)


@jit(nopython=True)
def is_point_on_upward_edge(E, N, U, w, e, s, n, b, t):
    return (b <= U <= t) and (E == w or E == e) and (N == s or N == n)


@jit(nopython=True)
def is_point_on_edge(E, N, U, w, e, s, n, b, t):
    return (
        is_point_on_easting_edge(E, N, U, w, e, s, n, b, t)
        or is_point_on_northing_edge(E, N, U, w, e, s, n, b, t)
        or is_point_on_upward_edge(E, N, U, w, e, s, n, b, t)
    )


@jit(nopython=True)
def is_point_on_east_face(E, N, U, w, e, s, n, b, t):
    return E == e and (s <= N <= n) and (b <= U <= t)


@jit(nopython=True)
def is_point_on_north_face(E, N, U, w, e, s, n, b, t):
    return N == n and (w <= E <= e) and (b <= U <= t)


@jit(nopython=True)
def is_point_on_top_face(E, N, U, w, e, s, n, b, t):
    return U == t and (w <= E <= e) and (s <= N <= n)


def _plain(kernel):
    @jit(nopython=True)
    def gravity(E, N, U, w, e, s, n, b, t, density):
        return GRAVITATIONAL_CONST * density * _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel)

    return gravity


gravity_pot = _plain(kernel_pot)
gravity_e = _plain(kernel_e)
gravity_n = _plain(kernel_n)
gravity_u = _plain(kernel_u)


@jit(nopython=True)
def gravity_ee(E, N, U, w, e, s, n, b, t, density):
    if is_point_on_northing_edge(E, N, U, w, e, s, n, b, t) or is_point_on_upward_edge(
        E, N, U, w, e, s, n, b, t
    ):
        return np.nan
    result = _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel_ee)
    if is_point_on_east_face(E, N, U, w, e, s, n, b, t):
        result += 4 * np.pi
    return GRAVITATIONAL_CONST * density * result


@jit(nopython=True)
def gravity_nn(E, N, U, w, e, s, n, b, t, density):
    if is_point_on_easting_edge(E, N, U, w, e, s, n, b, t) or is_point_on_upward_edge(
        E, N, U, w, e, s, n, b, t
    ):
        return np.nan
    result = _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel_nn)
    if is_point_on_north_face(E, N, U, w, e, s, n, b, t):
        result += 4 * np.pi
    return GRAVITATIONAL_CONST * density * result


@jit(nopython=True)
def gravity_uu(E, N, U, w, e, s, n, b, t, density):
    if is_point_on_easting_edge(E, N, U, w, e, s, n, b, t) or is_point_on_northing_edge(
        E, N, U, w, e, s, n, b, t
    ):
        return np.nan
    result = _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel_uu)
    if is_point_on_top_face(E, N, U, w, e, s, n, b, t):
        result += 4 * np.pi
    return GRAVITATIONAL_CONST * density * result


@jit(nopython=True)
def gravity_en(E, N, U, w, e, s, n, b, t, density):
    if is_point_on_upward_edge(E, N, U, w, e, s, n, b, t):
        return np.nan
    return GRAVITATIONAL_CONST * density * _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel_en)


@jit(nopython=True)
def gravity_eu(E, N, U, w, e, s, n, b, t, density):
    if is_point_on_northing_edge(E, N, U, w, e, s, n, b, t):
        return np.nan
    return GRAVITATIONAL_CONST * density * _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel_eu)


@jit(nopython=True)
def gravity_nu(E, N, U, w, e, s, n, b, t, density):
    if is_point_on_easting_edge(E, N, U, w, e, s, n, b, t):
        return np.nan
    return GRAVITATIONAL_CONST * density * _evaluate_kernel(E, N, U, w, e, s, n, b, t, kernel_nu)


@jit(nopython=True)
def magnetic_field(E, N, U, w, e, s, n, b, t, me, mn, mu):
    if is_point_on_edge(E, N, U, w, e, s, n, b, t):
        return np.nan, np.nan, np.nan
    be, bn, bu = 0.0, 0.0, 0.0
    for i in range(2):
        se = (e if i == 0 else w) - E
        se2 = se * se
        for j in range(2):
            sn = (n if j == 0 else s) - N
            sn2 = sn * sn
            for k in range(2):
                su = (t if k == 0 else b) - U
                su2 = su * su
                r = np.sqrt(se2 + sn2 + su2)
                sign = -1.0 if (i + j + k) % 2 else 1.0
                ee = kernel_ee(se, sn, su, r)
                nn = kernel_nn(se, sn, su, r)
                uu = kernel_uu(se, sn, su, r)
                en = kernel_en(se, sn, su, r)
                eu = kernel_eu(se, sn, su, r)
                nu = kernel_nu(se, sn, su, r)
                be += sign * (me * ee + mn * en + mu * eu)
                bn += sign * (me * en + mn * nn + mu * nu)
                bu += sign * (me * eu + mn * nu + mu * uu)
    if is_point_on_east_face(E, N, U, w, e, s, n, b, t):
        be += me * (4 * np.pi)
    if is_point_on_north_face(E, N, U, w, e, s, n, b, t):
        bn += mn * (4 * np.pi)
    if is_point_on_top_face(E, N, U, w, e, s, n, b, t):
        bu += mu * (4 * np.pi)
    cm = VACUUM_MAGNETIC_PERMEABILITY / 4 / np.pi
    return cm * be, cm * bn, cm * bu


@jit(nopython=True)
def magnetic_e(E, N, U, w, e, s, n, b, t, me, mn, mu):
    return magnetic_field(E, N, U, w, e, s, n, b, t, me, mn, mu)[0]


@jit(nopython=True)
def magnetic_n(E, N, U, w, e, s, n, b, t, me, mn, mu):
    return magnetic_field(E, N, U, w, e, s, n, b, t, me, mn, mu)[1]


@jit(nopython=True)
def magnetic_u(E, N, U, w, e, s, n, b, t, me, mn, mu):
    return magnetic_field(E, N, U, w, e, s, n, b, t, me, mn, mu)[2]


# ---------------------------------------------------------------- point masses
def _point(which):
    @jit(nopython=True)
    def gravity(E, N, U, eq, nq, uq, mass):
        de, dn, du = E - eq, N - nq, U - uq
        d = np.sqrt(de * de + dn * dn + du * du)
        if which == 0:
            k = 1 / d
        elif which == 1:
            k = -de / (d * d * d)
        elif which == 2:
            k = -dn / (d * d * d)
        elif which == 3:
            k = -du / (d * d * d)
        elif which == 4:
            k = 3 * de * de / (d * d * d * d * d) - 1 / (d * d * d)
        elif which == 5:
            k = 3 * dn * dn / (d * d * d * d * d) - 1 / (d * d * d)
        elif which == 6:
            k = 3 * du * du / (d * d * d * d * d) - 1 / (d * d * d)
        elif which == 7:
            k = 3 * de * dn / (d * d * d * d * d)
        elif which == 8:
            k = 3 * de * du / (d * d * d * d * d)
        else:
            k = 3 * dn * du / (d * d * d * d * d)
        return GRAVITATIONAL_CONST * mass * k

    return gravity


# -------------------------------------------------------------------- dipoles
@jit(nopython=True)
def dipole_magnetic_field(E, N, U, eq, nq, uq, me, mn, mu):
    re, rn, ru = E - eq, N - nq, U - uq
    d = np.sqrt(re * re + rn * rn + ru * ru)
    dot = me * re + mn * rn + mu * ru
    cm = VACUUM_MAGNETIC_PERMEABILITY / 4 / np.pi
    d3 = d * d * d
    d5 = d3 * d * d
    return (
        cm * (3 * dot * re / d5 - me / d3),
        cm * (3 * dot * rn / d5 - mn / d3),
        cm * (3 * dot * ru / d5 - mu / d3),
    )


@jit(nopython=True)
def dipole_magnetic_e(E, N, U, eq, nq, uq, me, mn, mu):
    return dipole_magnetic_field(E, N, U, eq, nq, uq, me, mn, mu)[0]


@jit(nopython=True)
def dipole_magnetic_n(E, N, U, eq, nq, uq, me, mn, mu):
    return dipole_magnetic_field(E, N, U, eq, nq, uq, me, mn, mu)[1]


@jit(nopython=True)
def dipole_magnetic_u(E, N, U, eq, nq, uq, me, mn, mu):
    return dipole_magnetic_field(E, N, U, eq, nq, uq, me, mn, mu)[2]


def install_fake_choclo():
    """Register fake ``choclo`` modules in sys.modules (idempotent)."""
    if "choclo" in sys.modules and getattr(sys.modules["choclo"], "__hb200_fake__", False):
        return sys.modules["choclo"]
    choclo = types.ModuleType("choclo")
    choclo.__hb200_fake__ = True
    choclo.__path__ = []
    constants = types.ModuleType("choclo.constants")
    constants.GRAVITATIONAL_CONST = GRAVITATIONAL_CONST
    constants.VACUUM_MAGNETIC_PERMEABILITY = VACUUM_MAGNETIC_PERMEABILITY
    prism = types.ModuleType("choclo.prism")
    prism.__path__ = []
    for name in ("pot", "e", "n", "u", "ee", "nn", "uu", "en", "eu", "nu"):
        setattr(prism, f"gravity_{name}", globals()[f"gravity_{name}"])
    for name in ("field", "e", "n", "u"):
        setattr(prism, f"magnetic_{name}", globals()[f"magnetic_{name}"])
    utils = types.ModuleType("choclo.prism._utils")
    for name in (
        "is_point_on_easting_edge",
        "is_point_on_northing_edge",
        "is_point_on_upward_edge",
        "is_point_on_edge",
        "safe_log",
        "safe_atan2",
    ):
        setattr(utils, name, globals()[name])
    prism._utils = utils
    point = types.ModuleType("choclo.point")
    for idx, name in enumerate(("pot", "e", "n", "u", "ee", "nn", "uu", "en", "eu", "nu")):
        setattr(point, f"gravity_{name}", _point(idx))
    dipole = types.ModuleType("choclo.dipole")
    dipole.magnetic_field = dipole_magnetic_field
    dipole.magnetic_e = dipole_magnetic_e
    dipole.magnetic_n = dipole_magnetic_n
    dipole.magnetic_u = dipole_magnetic_u
    choclo.constants, choclo.prism, choclo.point, choclo.dipole = constants, prism, point, dipole
    sys.modules.update(
        {
            "choclo": choclo,
            "choclo.constants": constants,
            "choclo.prism": prism,
            "choclo.prism._utils": utils,
            "choclo.point": point,
            "choclo.dipole": dipole,
        }
    )
    return choclo
