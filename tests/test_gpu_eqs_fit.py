"""
GPU tests of the device-resident equivalent-sources fits: ``eqs_fit`` /
``EquivalentSources.fit`` / ``EquivalentSourcesSph.fit`` / ``EquivalentSourcesGB.fit`` through the
public API -> ctypes -> C ABI (``hb200_eqs_fit``, ``hb200_eqs_fit_gb``, ``hb200_eqs_jacobian*``).
The case bodies live in ``_eqs_cases.py`` (they also run on the CPU with the device calls
substituted by the checker, ``test_eqs_classes_host.py``).
"""

import numpy as np
import pytest

import _eqs_cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sample(hb):
    return C.make_sample(hb)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("shape", [(300, 120), (120, 300), (200, 200)])
@pytest.mark.parametrize("damping", [None, 1e-3])
def test_eqs_fit_against_verde_least_squares(hb, shape, damping, weighted):
    C.case_eqs_fit_against_verde_least_squares(hb, shape, damping, weighted)


def test_eqs_jacobian_spherical_against_greens_function(hb):
    C.case_eqs_jacobian_spherical_against_greens_function(hb)


@pytest.mark.parametrize("weights", [None, np.ones((8, 8))], ids=["none", "ones"])
def test_equivalent_sources_small_data(hb, sample, weights):
    C.case_equivalent_sources_small_data(hb, sample, weights)


def test_equivalent_sources_cartesian(hb, sample):
    C.case_equivalent_sources_cartesian(hb, sample)


def test_equivalent_sources_block_averaged_and_damped(hb, sample):
    C.case_equivalent_sources_block_averaged_and_damped(hb, sample)


def test_equivalent_sources_spherical(hb):
    C.case_equivalent_sources_spherical(hb)


@pytest.mark.parametrize("weighted", [False, True])
def test_gradient_boosting_loop_against_checker(hb, sample, weighted):
    C.case_gradient_boosting_loop_against_checker(hb, sample, weighted)


@pytest.mark.parametrize("weights", [None, np.ones((8, 8))], ids=["none", "ones"])
def test_gb_eqs_small_data(hb, sample, weights):
    C.case_gb_eqs_small_data(hb, sample, weights)


def test_gradient_boosted_eqs_single_window_and_predictions(hb, sample):
    C.case_gradient_boosted_eqs_single_window_and_predictions(hb, sample)


def test_fit_launches_kernels_and_keeps_the_jacobian_on_the_device(hb, sample):
    lib = hb._lib.load()
    before = lib.hb200_launch_count()
    hb.EquivalentSources(depth=500, damping=1e-6).fit(sample["coordinates"], sample["data"])
    assert lib.hb200_launch_count() - before >= 5  # jacobian, scaling x2, diagonal, unscale


def test_fit_rcond_matches_the_installed_sklearn(hb):
    """hb200_set_fit_rcond(1e-6): the truncated minimum-norm solution of an UNDER-determined
    undamped system equals LinearRegression(tol=1e-6)'s (scikit-learn >= 1.7) coefficient by
    coefficient; the default (machine epsilon) fits the data more closely."""
    import warnings

    from test_eqs_fit_host import _system, verde_least_squares

    lib = hb._lib.load()
    rng = np.random.default_rng(7)
    n, p = 120, 300
    obs = rng.uniform(0, 5e3, (n, 3)) * [1, 1, 0.02]
    src = rng.uniform(0, 5e3, (p, 3)) * [1, 1, 0.0] - [0, 0, 600.0]
    jac = 1 / np.sqrt(((obs[:, None, :] - src[None, :, :]) ** 2).sum(-1))
    data = jac @ rng.normal(size=p) * 1e3
    coords, points = tuple(obs.T.copy()), tuple(src.T.copy())
    try:
        hb._lib.check(lib.hb200_set_fit_rcond(1e-6))
        got = hb.eqs_fit(coords, points, data)
    finally:
        hb._lib.check(lib.hb200_set_fit_rcond(np.finfo(float).eps))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = verde_least_squares(jac, data, None, None)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6 * np.abs(want).max())
    tight = hb.eqs_fit(coords, points, data)
    assert np.linalg.norm(jac @ tight - data) <= np.linalg.norm(jac @ got - data) * (1 + 1e-6)
    assert _system is not None


def test_coincident_source_raises_zero_division(hb):
    """The reference's jitted Jacobian / predict loops raise ZeroDivisionError when a data point
    coincides with a source (numba error_model='python'); so do the fits here."""
    rng = np.random.default_rng(8)
    coords = tuple(rng.uniform(0, 1e3, (3, 50)))
    points = tuple(c.copy() for c in coords)  # sources ON the data points
    data = rng.normal(size=50)
    with pytest.raises(ZeroDivisionError):
        hb.eqs_jacobian(coords, points)
    with pytest.raises(ZeroDivisionError):
        hb.eqs_fit(coords, points, data, damping=1e-3)
    with pytest.raises(ZeroDivisionError):
        hb.EquivalentSources(points=points, damping=1e-3).fit(coords, data)
    # the spherical class accepts relative_depth = 0 (the reference does not validate it) ...
    sph = hb.EquivalentSourcesSph(relative_depth=0)
    lon, lat = rng.uniform(-10, 10, 30), rng.uniform(-10, 10, 30)
    with pytest.raises(ZeroDivisionError):  # ... and then divides by zero in the fit
        sph.fit((lon, lat, np.full(30, 6371e3)), rng.normal(size=30))
