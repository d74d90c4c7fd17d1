"""
CPU run of the equivalent-sources HOST logic (``EquivalentSources`` / ``EquivalentSourcesGB`` /
``EquivalentSourcesSph``: input checks, source placement, block averaging, windows, dtype rules)
with the package's device calls substituted by the checker: the oracle's loops for the pair
sums and the replay of the device's dense-solve sequence (``test_eqs_fit_host``). The same case
bodies run against the real CUDA path in ``test_gpu_eqs_fit.py``; here they prove that the
expectations hold for the algorithm the device executes, without a GPU.
"""

import types

import numpy as np
import pytest

import _eqs_cases as C
import oracle as O
from test_eqs_fit_host import replay_dense_least_squares


def _jacobian(coords, points, system):
    coords = tuple(np.asarray(c, dtype=float).ravel() for c in coords[:3])
    points = tuple(np.asarray(p, dtype=float).ravel() for p in points[:3])
    if system == "cartesian":
        return O.eqs_jacobian(coords, points)
    import harmonica_b200._eqs as E

    return E.greens_func_spherical(coords[0][:, None], coords[1][:, None], coords[2][:, None],
                                   points[0][None, :], points[1][None, :], points[2][None, :])  # fmt: skip


@pytest.fixture()
def hb(monkeypatch):
    import harmonica_b200 as package
    import harmonica_b200._eqs as E

    def eqs_fit(coordinates, points, data, weights=None, damping=None, *,
                coordinate_system="cartesian", return_solver_path=False):  # fmt: skip
        coords, points, data, weights = E._fit_inputs(coordinates, points, data, weights, coordinate_system)
        if data.size < points[0].size:
            import warnings

            warnings.warn(f"Under-determined problem detected (ndata, nparams)={(data.size, points[0].size)}.",
                          stacklevel=2)  # fmt: skip
        coefs, path = replay_dense_least_squares(_jacobian(coords, points, coordinate_system), data, weights, damping)
        return (coefs, path) if return_solver_path else coefs

    def eqs_predict(coordinates, points, coefs, dtype="float64", *, coordinate_system="cartesian", shard="auto"):
        shape = np.broadcast(*coordinates[:3]).shape
        fn = O.eqs_predict if coordinate_system == "cartesian" else O.eqs_predict_spherical
        coords = tuple(np.asarray(c, dtype=float).ravel() for c in coordinates[:3])
        return np.asarray(fn(coords, points, coefs)).astype(dtype).reshape(shape)

    def eqs_fit_gradient_boosted(coordinates, points, data, weights, damping, source_windows,
                                 data_windows, *, coordinate_system="cartesian"):  # fmt: skip
        coords, points, data, weights = E._fit_inputs(coordinates, points, data, weights, coordinate_system)
        coefs, residue = np.zeros(points[0].size), data.copy()
        errors = [np.sqrt(np.mean(data**2))]
        for pw, dw in zip(source_windows, data_windows):
            pts, cds = tuple(p[pw] for p in points), tuple(c[dw] for c in coords)
            chunk, _ = replay_dense_least_squares(_jacobian(cds, pts, coordinate_system), residue[dw],
                                                  None if weights is None else weights[dw], damping)  # fmt: skip
            residue -= O.eqs_predict(coords, pts, chunk)
            errors.append(np.sqrt(np.mean(residue**2)))
            coefs[pw] += chunk
        return coefs, np.array(errors)

    def point_gravity(coordinates, points, masses, field, coordinate_system="cartesian", **kwargs):
        return O.point_gravity(coordinates, points, masses, field, coordinate_system=coordinate_system)

    for name, fn in [("eqs_fit", eqs_fit), ("eqs_predict", eqs_predict),
                     ("eqs_fit_gradient_boosted", eqs_fit_gradient_boosted)]:  # fmt: skip
        monkeypatch.setattr(E, name, fn)
    monkeypatch.setattr(E, "eqs_jacobian_spherical", lambda c, p, dtype="float64": _jacobian(c, p, "spherical"))
    ns = types.SimpleNamespace(
        EquivalentSources=package.EquivalentSources, EquivalentSourcesGB=package.EquivalentSourcesGB,
        EquivalentSourcesSph=package.EquivalentSourcesSph, eqs_fit=eqs_fit, point_gravity=point_gravity,
        eqs_jacobian_spherical=lambda c, p: _jacobian(c, p, "spherical"), _eqs=E,
    )  # fmt: skip
    return ns


@pytest.fixture()
def sample(hb):
    return C.make_sample(hb)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("shape", [(300, 120), (120, 300), (200, 200)])
@pytest.mark.parametrize("damping", [None, 1e-3])
def test_eqs_fit_against_verde_least_squares(hb, shape, damping, weighted):
    C.case_eqs_fit_against_verde_least_squares(hb, shape, damping, weighted)


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
