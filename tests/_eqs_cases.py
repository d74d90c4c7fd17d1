"""
Shared bodies of the equivalent-sources fit tests. Every case takes ``hb`` (the package, or
the package with its device calls substituted by the checker for the CPU run of the host
logic) and checks it against the oracle's Jacobian / prediction loops plus the scikit-learn
calls ``verde.base.least_squares`` makes, and against the acceptance criteria of the reference's
own tests (``test/test_eq_sources_cartesian.py``, ``test_gradient_boosted_eqs.py``,
``test_eq_sources_spherical.py``).
"""

import warnings

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import TOL, max_rel
from test_eqs_fit_host import verde_least_squares


REGION = (-3e3, -1e3, 5e3, 7e3)


def grid(shape, upward, region=REGION):
    e, n = np.meshgrid(np.linspace(region[0], region[1], shape[1]),
                       np.linspace(region[2], region[3], shape[0]))  # fmt: skip
    return e, n, np.full_like(e, float(upward))


def checkerboard_masses(pts, region=REGION, amplitude=1e13):
    """verde.synthetic.CheckerBoard(amplitude, region).predict(points)"""
    w_e, w_n = (region[1] - region[0]) / 2, (region[3] - region[2]) / 2
    return amplitude * np.sin((2 * np.pi / w_e) * (pts[0] - region[0])) * np.cos(
        (2 * np.pi / w_n) * (pts[1] - region[2]))  # fmt: skip


def make_sample(hb):
    pts = grid((6, 6), -1e3)
    masses = checkerboard_masses(pts)
    coords = grid((40, 40), 0)
    small = grid((8, 8), 0)
    return {
        "points": pts, "masses": masses, "coordinates": coords,
        "data": hb.point_gravity(coords, pts, masses, "g_z"),
        "coordinates_small": small,
        "data_small": hb.point_gravity(small, pts, masses, "g_z"),
    }  # fmt: skip


def quiet(fn, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


# ------------------------------------------------------------------ the solve itself
def case_eqs_fit_against_verde_least_squares(hb, shape, damping, weighted):
    rng = np.random.default_rng(5)
    n, p = shape
    coords = (rng.uniform(0, 5e3, n), rng.uniform(0, 5e3, n), rng.uniform(0, 100, n))
    points = (rng.uniform(0, 5e3, p), rng.uniform(0, 5e3, p), np.full(p, -600.0))
    jac = O.eqs_jacobian(coords, points)
    data = jac @ rng.normal(size=p) * 1e3
    weights = rng.uniform(0.5, 2.0, n) if weighted else None
    got, path = quiet(hb.eqs_fit, coords, points, data, weights, damping, return_solver_path=True)
    assert path == (1 if damping is None else 0)
    want = quiet(verde_least_squares, jac, data, weights, damping)
    if damping is None and n <= p:
        sw = np.ones(n) if weights is None else np.sqrt(weights)
        npt.assert_allclose(sw * (jac @ got), sw * (jac @ want), atol=1e-4 * np.abs(data).max())
    else:
        npt.assert_allclose(got, want, rtol=1e-5, atol=1e-8 * np.abs(want).max())
    if n < p:
        with pytest.warns(UserWarning, match="Under-determined problem"):
            hb.eqs_fit(coords, points, data, weights, damping)


def case_eqs_jacobian_spherical_against_greens_function(hb):
    rng = np.random.default_rng(6)
    obs = (rng.uniform(-40, 40, 70), rng.uniform(-60, 60, 70), rng.uniform(6.4e6, 6.5e6, 70))
    pts = (rng.uniform(-45, 45, 50), rng.uniform(-65, 65, 50), rng.uniform(6.2e6, 6.3e6, 50))
    want = hb._eqs.greens_func_spherical(obs[0][:, None], obs[1][:, None], obs[2][:, None],
                                         pts[0][None, :], pts[1][None, :], pts[2][None, :])  # fmt: skip
    got = hb.eqs_jacobian_spherical(obs, pts)
    assert got.shape == (70, 50)
    assert max_rel(got, want) <= TOL
    # each column is the prediction of a unit source: same kernel as the oracle's predict loop
    col = O.eqs_predict_spherical(obs, pts, np.eye(50)[7])
    assert max_rel(got[:, 7], col) <= TOL


# ------------------------------------------------------------------ EquivalentSources
def case_equivalent_sources_small_data(hb, sample, weights):
    """test/test_eq_sources_cartesian.py:157-183"""
    coords, data = sample["coordinates_small"], sample["data_small"]
    eqs = hb.EquivalentSources(depth=500).fit(coords, data, weights=weights)
    npt.assert_allclose(data, eqs.predict(coords), rtol=1e-5)
    npt.assert_allclose([c.ravel() for c in coords[:2]], eqs.points_[:2], rtol=1e-5)
    npt.assert_allclose(coords[2].ravel() - 500, eqs.points_[2], rtol=1e-5)
    up = grid((8, 8), 20)
    true = hb.point_gravity(up, sample["points"], sample["masses"], "g_z")
    npt.assert_allclose(true, eqs.predict(up), rtol=0.08)


def case_equivalent_sources_cartesian(hb, sample):
    """test/test_eq_sources_cartesian.py:111-154: 40 x 40 data, interpolate onto 60 x 60"""
    coords, data = sample["coordinates"], sample["data"]
    atol = 1e-3 * np.abs(data).max()
    eqs = hb.EquivalentSources(depth=500).fit(coords, data)
    npt.assert_allclose(data, eqs.predict(coords), atol=atol)
    dense = grid((60, 60), 0)
    true = hb.point_gravity(dense, sample["points"], sample["masses"], "g_z")
    npt.assert_allclose(true, eqs.predict(dense), atol=atol)


def case_equivalent_sources_block_averaged_and_damped(hb, sample):
    """test/test_eq_sources_cartesian.py:239-254, 340-375"""
    coords, data = sample["coordinates"], sample["data"]
    eqs = hb.EquivalentSources(depth=500, block_size=500, damping=1e-6).fit(coords, data)
    assert eqs.points_[0].size == 16  # 2 km / 500 m = 4 blocks per axis
    assert eqs.coefs_.shape == (16,)
    jac = O.eqs_jacobian(tuple(c.ravel() for c in coords), eqs.points_)
    want = verde_least_squares(jac, data.ravel(), None, 1e-6)
    npt.assert_allclose(eqs.coefs_, want, rtol=1e-6)
    for dtype in ("float64", "float32"):
        eqs = hb.EquivalentSources(depth=500, damping=1e-3, dtype=dtype).fit(coords, data)
        assert eqs.coefs_.dtype == np.float64  # like the reference: the solve is float64
        assert eqs.predict(coords).dtype == np.dtype(dtype)
        assert all(p.dtype == np.dtype(dtype) for p in eqs.points_)


# ------------------------------------------------------------------ EquivalentSourcesSph
def case_equivalent_sources_spherical(hb):
    """test/test_eq_sources_spherical.py:26-67: fit point-mass data, interpolate upwards"""
    region = (-70, -60, -40, -30)
    radius = 6400e3
    lon_p, lat_p = np.meshgrid(np.linspace(region[0], region[1], 6), np.linspace(region[2], region[3], 6))
    points = (lon_p, lat_p, np.full_like(lon_p, radius - 500e3))
    masses = checkerboard_masses(points, region=region)
    lon, lat = np.meshgrid(np.linspace(region[0], region[1], 30), np.linspace(region[2], region[3], 30))
    coords = (lon, lat, np.full_like(lon, radius))
    data = hb.point_gravity(coords, points, masses, "g_z", coordinate_system="spherical")
    atol = 1e-3 * np.abs(data).max()
    eqs = hb.EquivalentSourcesSph(relative_depth=500e3).fit(coords, data)
    npt.assert_allclose(data, eqs.predict(coords), atol=atol)
    npt.assert_allclose(eqs.points_[2], radius - 500e3)
    upward = (lon, lat, np.full_like(lon, radius + 2e3))
    true = hb.point_gravity(upward, points, masses, "g_z", coordinate_system="spherical")
    npt.assert_allclose(true, eqs.predict(upward), atol=atol)
    # custom points (:112-150) and the damped solve against the checker
    src = (lon_p.ravel(), lat_p.ravel(), np.full(36, radius - 300e3))
    eqs = hb.EquivalentSourcesSph(points=src, damping=1e-6).fit(coords, data)
    jac = hb.eqs_jacobian_spherical(tuple(c.ravel() for c in coords), src)
    npt.assert_allclose(eqs.coefs_, verde_least_squares(jac, data.ravel(), None, 1e-6), rtol=1e-6)


# ------------------------------------------------------------------ EquivalentSourcesGB
def _gb_checker(coords, points, data, weights, damping, source_windows, data_windows):
    """gradient_boosted.py:244-293 with the oracle's loops and verde's solve."""
    coefs = np.zeros(points[0].size)
    residue = data.copy()
    errors = [np.sqrt(np.mean(data**2))]
    for pw, dw in zip(source_windows, data_windows):
        pts = tuple(p[pw] for p in points)
        cds = tuple(c[dw] for c in coords)
        jac = O.eqs_jacobian(cds, pts)
        chunk = quiet(verde_least_squares, jac, residue[dw], None if weights is None else weights[dw], damping)
        residue -= O.eqs_predict(coords, pts, chunk)
        errors.append(np.sqrt(np.mean(residue**2)))
        coefs[pw] += chunk
    return coefs, np.array(errors)


def case_gradient_boosting_loop_against_checker(hb, sample, weighted):
    coords = tuple(c.ravel() for c in sample["coordinates"])
    data = sample["data"].ravel()
    weights = np.random.default_rng(3).uniform(0.5, 2, data.size) if weighted else None
    eqs = hb.EquivalentSourcesGB(depth=1e3, damping=1e-1, window_size=1e3, random_state=42)
    eqs.fit(coords, data, weights=weights)
    source_windows, data_windows = eqs._create_windows(coords)  # same random_state, same order
    assert len(source_windows) == 9
    want_coefs, want_rmse = _gb_checker(coords, eqs.points_, data, weights, 1e-1, source_windows, data_windows)
    # The RMSE history and the predicted field are well conditioned. The coefficients are not:
    # every later window fits a residue that has lost digits by cancellation (data - predicted,
    # relative rounding ~1e-13 once the residue is 1e-3 of the data) and the window systems have
    # condition numbers ~1e8, so two correct implementations with different rounding (the CPU
    # checker / the GPU) agree on them only to ~1e-4 of the largest coefficient (measured 6e-5).
    assert eqs.rmse_per_iteration_.shape == (10,)
    npt.assert_allclose(eqs.rmse_per_iteration_, want_rmse, rtol=1e-5)
    predicted = O.eqs_predict(coords, eqs.points_, eqs.coefs_)
    want_predicted = O.eqs_predict(coords, eqs.points_, want_coefs)
    npt.assert_allclose(predicted, want_predicted, rtol=0, atol=1e-7 * np.abs(data).max())
    npt.assert_allclose(eqs.coefs_, want_coefs, rtol=0, atol=1e-3 * np.abs(want_coefs).max())


def case_gb_eqs_small_data(hb, sample, weights):
    """test/test_gradient_boosted_eqs.py:157-176"""
    coords, data = sample["coordinates_small"], sample["data_small"]
    eqs = hb.EquivalentSourcesGB(depth=1e3, damping=None, window_size=1e3, random_state=42)
    eqs.fit(coords, data, weights=weights)
    npt.assert_allclose(data, eqs.predict(coords), rtol=0, atol=0.05 * np.abs(data).max())


def case_gradient_boosted_eqs_single_window_and_predictions(hb, sample):
    """test/test_gradient_boosted_eqs.py:178-221"""
    coords, data = sample["coordinates"], sample["data"]
    dense = grid((60, 60), 0)
    true = hb.point_gravity(dense, sample["points"], sample["masses"], "g_z")
    eqs = hb.EquivalentSourcesGB(depth=500, window_size=REGION[1] - REGION[0], damping=1e-18)
    eqs.fit(coords, data)
    npt.assert_allclose(data, eqs.predict(coords), rtol=1e-5, atol=5e-8)
    npt.assert_allclose(true, eqs.predict(dense), rtol=1e-3, atol=5e-8)
    eqs = hb.EquivalentSourcesGB(window_size=1e3, depth=1e3, damping=1e-24, random_state=42)
    eqs.fit(coords, data)
    npt.assert_allclose(data, eqs.predict(coords), rtol=0, atol=0.02 * np.abs(data).max())
    npt.assert_allclose(true, eqs.predict(dense), rtol=0, atol=0.02 * np.abs(true).max())
    # same random_state, same coefficients (:223-235); custom points (:113-127)
    a = hb.EquivalentSourcesGB(window_size=500, random_state=0, damping=1e-6).fit(coords, data)
    b = hb.EquivalentSourcesGB(window_size=500, random_state=0, damping=1e-6).fit(coords, data)
    npt.assert_allclose(a.coefs_, b.coefs_)
    custom = grid((3, 3), -550)
    eqs = hb.EquivalentSourcesGB(points=custom, window_size=500, depth=500)
    eqs.fit(sample["coordinates_small"], sample["data_small"])
    assert eqs.depth_ is None
    npt.assert_allclose([p.ravel() for p in custom], eqs.points_)
