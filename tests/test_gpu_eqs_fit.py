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
