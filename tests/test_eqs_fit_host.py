"""
CPU checks of the equivalent-sources FIT host logic (no GPU):

* the dense-solve call sequence of ``dense_least_squares`` (``csrc/hb200_fit_host.cuh``) replayed
  on flat buffers with column-major BLAS / LAPACK semantics (same routine names, same
  ``m, n, lda`` arguments) against the scikit-learn calls ``verde.base.least_squares`` makes -
  this pins the transposition conventions and the four solver branches;
* the verde / bordado helpers restated in ``harmonica_b200/_gridding.py`` against the golden
  values the reference's tests hold.
"""

import warnings

import numpy as np
import numpy.testing as npt
import pytest

from harmonica_b200 import _gridding


# ------------------------------------------------------------------ column-major mini BLAS
def _cm(buf, rows, cols, ld):
    """Column-major (rows x cols, leading dimension ld) view of a flat buffer."""
    return np.lib.stride_tricks.as_strided(buf, (rows, cols), (buf.itemsize, ld * buf.itemsize))


def dsyrk_lower(trans, n, k, a, lda, c, ldc):
    op = _cm(a, n, k, lda) if trans == "N" else _cm(a, k, n, lda).T
    full = op @ op.T
    view = _cm(c, n, n, ldc)
    view[np.tril_indices(n)] = full[np.tril_indices(n)]  # only the lower triangle is written


def dgemv(trans, m, n, a, lda, x, y):
    mat = _cm(a, m, n, lda)
    y[: (m if trans == "N" else n)] = (mat if trans == "N" else mat.T) @ x[: (n if trans == "N" else m)]


def dgeam_transpose(m, n, a, lda, c, ldc):
    _cm(c, m, n, ldc)[:] = _cm(a, n, m, lda).T


def dpotrf_lower(n, a, lda):
    view = _cm(a, n, n, lda)
    sym = np.tril(view) + np.tril(view, -1).T
    try:
        view[np.tril_indices(n)] = np.linalg.cholesky(sym)[np.tril_indices(n)]
    except np.linalg.LinAlgError:
        return 1
    return 0


def dpotrs_lower(n, a, lda, b):
    import scipy.linalg as sl

    low = np.tril(_cm(a, n, n, lda))
    b[:n] = sl.cho_solve((low, True), b[:n])


def dgesvd_ss(m, n, a, lda, s, u, ldu, vt, ldvt):
    assert m >= n, "cusolverDnDgesvd needs m >= n"
    uu, ss, vv = np.linalg.svd(_cm(a, m, n, lda), full_matrices=False)
    s[:n] = ss
    _cm(u, m, n, ldu)[:] = uu
    _cm(vt, n, n, ldvt)[:] = vv


#: hb200_set_fit_rcond(1e-6) == LinearRegression(tol=1e-6) of the installed scikit-learn (>= 1.7);
#: the library's default is machine epsilon (cond=None of older releases)
SKLEARN_RCOND = 1e-6


def replay_dense_least_squares(jacobian, data, weights, damping, force_svd_fallback=False,
                               rcond=np.finfo(float).eps):
    """The steps of dense_least_squares(), in its order, on a row-major n x p buffer."""
    n, p = jacobian.shape
    jac = np.array(jacobian, dtype=np.float64).ravel()  # row-major n x p == column-major p x n
    # column_scale_kernel
    mean = jacobian.mean(axis=0)
    var = ((jacobian - mean) ** 2).mean(axis=0)
    eps = np.finfo(float).eps
    bound = n * eps * var + (n * mean * eps) ** 2
    scale = np.where(var <= bound, 1.0, np.sqrt(var))
    # scale_system_kernel
    view = jac.reshape(n, p)
    view /= scale
    y = np.array(data, dtype=np.float64)
    if weights is not None:
        view *= np.sqrt(weights)[:, None]
        y = y * np.sqrt(weights)
    x = np.zeros(max(n, p))
    damped = damping is not None
    need_svd, mode, param, path = not damped, 0, rcond, 0
    if damped:
        primal = p <= n
        k = p if primal else n
        g = np.zeros(k * k)
        if primal:
            dsyrk_lower("N", p, n, jac, p, g, p)
            dgemv("N", p, n, jac, p, y, x)
        else:
            dsyrk_lower("T", n, p, jac, p, g, n)
            x[:n] = y
        g[:: k + 1] += damping
        info = 1 if force_svd_fallback else dpotrf_lower(k, g, k)
        if info == 0:  # diagonal_minmax_kernel + the pivot-ratio guard
            pivots = g[:: k + 1]
            if not (pivots.min() / pivots.max()) ** 2 > eps:
                info = k + 1
        if info == 0:
            dpotrs_lower(k, g, k, x)
            if not primal:
                tmp = x[:n].copy()
                dgemv("N", p, n, jac, p, tmp, x)
        else:
            need_svd, mode, param, path = True, 1, damping, 2
    if need_svd:
        if not damped:
            path = 1
        tall_m = p >= n
        m, k = (p, n) if tall_m else (n, p)
        a = jac
        if not tall_m:
            a = np.zeros(n * p)
            dgeam_transpose(n, p, jac, p, a, n)
        s, u, vt, t = np.zeros(k), np.zeros(m * k), np.zeros(k * k), np.zeros(k)
        dgesvd_ss(m, k, a, m, s, u, m, vt, k)
        if tall_m:
            dgemv("N", k, k, vt, k, y, t)
        else:
            dgemv("T", m, k, u, m, y, t)
        # singular_filter_kernel
        if mode == 0:
            t *= np.where(s > param * s[0], 1.0 / s, 0.0)
        else:
            t *= np.where(s > 1e-15, s / (s * s + param), 0.0)
        if tall_m:
            dgemv("N", m, k, u, m, t, x)
        else:
            dgemv("T", k, k, vt, k, t, x)
    return x[:p] / scale, path


def verde_least_squares(jacobian, data, weights, damping):
    """What the reference calls (verde.base.least_squares): the scikit-learn sequence."""
    from sklearn.linear_model import LinearRegression, Ridge
    from sklearn.preprocessing import StandardScaler

    scaler = StandardScaler(copy=False, with_mean=False, with_std=True)
    jacobian = scaler.fit_transform(jacobian.copy())
    regr = (LinearRegression(fit_intercept=False) if damping is None
            else Ridge(alpha=damping, fit_intercept=False))  # fmt: skip
    regr.fit(jacobian, data.ravel(), sample_weight=weights)
    return regr.coef_ / scaler.scale_


def _system(rng, n, p, depth=600.0):
    obs = rng.uniform(0, 5e3, (n, 3)) * [1, 1, 0.02]
    src = rng.uniform(0, 5e3, (p, 3)) * [1, 1, 0.0] - [0, 0, depth]
    jac = 1 / np.sqrt(((obs[:, None, :] - src[None, :, :]) ** 2).sum(-1))
    data = jac @ rng.normal(size=p) * 1e3
    return jac, data


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("shape", [(60, 25), (25, 60), (40, 40)])
@pytest.mark.parametrize("damping", [None, 1e-3])
def test_dense_solve_sequence_matches_verde(shape, damping, weighted):
    rng = np.random.default_rng(7)
    jac, data = _system(rng, *shape)
    weights = rng.uniform(0.5, 2.0, shape[0]) if weighted else None
    got, path = replay_dense_least_squares(jac, data, weights, damping, rcond=SKLEARN_RCOND)
    assert path == (1 if damping is None else 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = verde_least_squares(jac, data, weights, damping)
    if damping is None:
        # same singular-value cutoff as the installed scikit-learn's (cond = tol = 1e-6): the
        # truncated minimum-norm solutions agree, also for (under)determined ill-conditioned systems
        npt.assert_allclose(got, want, rtol=1e-5, atol=1e-7 * np.abs(want).max())
        # the library's default cutoff (machine epsilon) fits the data more closely
        tight, _ = replay_dense_least_squares(jac, data, weights, damping)
        sw = np.ones(shape[0]) if weights is None else np.sqrt(weights)
        misfit = lambda c: np.linalg.norm(sw * (jac @ c - data))  # noqa: E731
        assert misfit(tight) <= misfit(got) * (1 + 1e-9) + 1e-9 * np.abs(data).max()
    else:
        npt.assert_allclose(got, want, rtol=2e-6, atol=1e-9 * np.abs(want).max())


@pytest.mark.parametrize("shape", [(60, 25), (25, 60)])
def test_dense_solve_svd_ridge_fallback(shape):
    rng = np.random.default_rng(8)
    jac, data = _system(rng, *shape)
    got, path = replay_dense_least_squares(jac, data, None, 1e-3, force_svd_fallback=True)
    assert path == 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = verde_least_squares(jac, data, None, 1e-3)
    npt.assert_allclose(got, want, rtol=2e-6, atol=1e-9 * np.abs(want).max())


# ------------------------------------------------------------------ verde / bordado helpers
REGION = (-3e3, -1e3, 5e3, 7e3)


def _grid(region, shape=None, spacing=None):
    if spacing is not None:
        east = np.arange(region[0], region[1] + spacing / 2, spacing)
        north = np.arange(region[2], region[3] + spacing / 2, spacing)
    else:
        east = np.linspace(region[0], region[1], shape[1])
        north = np.linspace(region[2], region[3], shape[0])
    return np.meshgrid(east, north)


@pytest.mark.parametrize("block_size", [750, (750, 1e3)])
def test_block_average_coordinates_golden(block_size):
    """test/test_eq_sources_cartesian.py:216-236"""
    easting, northing = _grid(REGION, shape=(9, 9))
    upward = np.arange(81, dtype=float).reshape(9, 9)
    if np.ndim(block_size):
        expected = (
            [-2500.0, -1375.0, -2500.0, -1375.0, -2500.0, -1375.0],
            [5250.0, 5250.0, 6000.0, 6000.0, 6750.0, 6750.0],
            [11.0, 15.5, 38.0, 42.5, 65.0, 69.5],
        )
    else:
        expected = (
            [-2750, -2000, -1250, -2750, -2000, -1250, -2750, -2000, -1250],
            [5250, 5250, 5250, 6000, 6000, 6000, 6750, 6750, 6750],
            [10.0, 13.0, 16.0, 37.0, 40.0, 43.0, 64.0, 67.0, 70.0],
        )
    npt.assert_allclose(expected, _gridding.block_average_coordinates((easting, northing, upward), block_size))


def test_rolling_windows_golden():
    """test/test_gradient_boosted_eqs.py:263-295: 3 x 3 windows of size 1 over (1, 3, 1, 3)"""
    coordinates = _grid((1, 3, 1, 3), spacing=1)
    points = _grid((1, 2, 1, 3), spacing=1)
    kwargs = {"region": (1, 3, 1, 3), "window_size": 1, "overlap": 0.5}
    data_windows = _gridding.rolling_windows(coordinates, **kwargs)
    source_windows = _gridding.rolling_windows(points, **kwargs)
    assert len(data_windows) == len(source_windows) == 9
    assert [w.size for w in data_windows] == [4, 2, 4, 2, 1, 2, 4, 2, 4]
    assert [w.size for w in source_windows] == [4, 2, 2, 2, 1, 1, 4, 2, 2]
    # a single window that covers the whole region (:179-196)
    (single,) = _gridding.rolling_windows(coordinates, region=(1, 3, 1, 3), window_size=2, overlap=0.5)
    assert single.size == 9
    with pytest.raises(ValueError, match="larger than dimensions"):
        _gridding.rolling_windows(coordinates, region=(1, 3, 1, 3), window_size=2.5, overlap=0.5)


def test_shuffle_together_is_sklearns_shuffle():
    from sklearn.utils import shuffle

    a = [np.arange(k) for k in range(12)]
    b = [np.arange(k) + 100 for k in range(12)]
    want_a, want_b = shuffle(a, b, random_state=42)
    got_a, got_b = _gridding.shuffle_together(a, b, random_state=42)
    assert all(np.array_equal(x, y) for x, y in zip(want_a, got_a))
    assert all(np.array_equal(x, y) for x, y in zip(want_b, got_b))


def test_gb_window_creation_follows_reference_tests():
    """test/test_gradient_boosted_eqs.py:238-295, 404-437 (no device work)"""
    import harmonica_b200 as hb

    coordinates = tuple(c.ravel() for c in _grid((1, 3, 1, 3), spacing=1)) + (np.zeros(9),)
    eqs = hb.EquivalentSourcesGB(window_size=1)
    pe, pn = _grid((1, 2, 1, 3), spacing=1)
    eqs.points_ = (pe.ravel(), pn.ravel(), np.full(pe.size, -10.0))
    source_windows, data_windows = eqs._create_windows(coordinates, shuffle=False)
    assert [w.size for w in data_windows] == [4, 2, 4, 2, 1, 2, 4, 2, 4]
    assert [w.size for w in source_windows] == [4, 2, 2, 2, 1, 1, 4, 2, 2]
    # default window size: ~5000 data points per window
    region = (0, 10e3, -5e3, 5e3)
    east, north = _grid(region, shape=(100, 100))
    grid_coords = (east.ravel(), north.ravel(), np.zeros(east.size))
    eqs = hb.EquivalentSourcesGB()
    eqs.points_ = eqs._build_points(grid_coords)
    eqs._create_windows(grid_coords)
    npt.assert_allclose(eqs.window_size_, np.sqrt(5e3 / (100**2 / 10e3**2)))
    # <= 5000 points: one window, with a warning
    east, north = _grid(region, shape=(50, 50))
    small = (east.ravel(), north.ravel(), np.zeros(east.size))
    eqs = hb.EquivalentSourcesGB()
    eqs.points_ = eqs._build_points(small)
    with pytest.warns(UserWarning, match="Only one window will be used"):
        source_windows, data_windows = eqs._create_windows(small)
    assert eqs.window_size_ is None and len(source_windows) == len(data_windows) == 1
    assert source_windows[0].size == data_windows[0].size == 2500
    with pytest.raises(ValueError, match="Found invalid 'window_size' value equal to"):
        hb.EquivalentSourcesGB(window_size="Chuckie took my soul!")


def test_memory_estimation_follows_reference_tests():
    """test/test_gradient_boosted_eqs.py:130-150 and test_eq_sources_cartesian.py (no device)"""
    import harmonica_b200 as hb

    region = (-1e4, 1e4, -1e4, 1e4)
    for spacing, window_size in [(100, 1e3), (100, 2e3), (200, 4e3)]:
        east, north = _grid(region, spacing=spacing)
        coordinates = (east, north, np.zeros_like(east))
        per_window = (int(window_size / spacing) + 1) ** 2
        for dtype, itemsize in [("float64", 8), ("float32", 4)]:
            eqs = hb.EquivalentSourcesGB(window_size=window_size, dtype=dtype)
            assert eqs.estimate_required_memory(coordinates) == per_window**2 * itemsize
    east, north = _grid(region, spacing=1000)
    eqs = hb.EquivalentSources(depth=100)
    assert eqs.estimate_required_memory((east, north, np.zeros_like(east))) == east.size**2 * 8
