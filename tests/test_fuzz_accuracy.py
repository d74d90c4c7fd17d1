"""
Randomised geometry fuzz of the product's per-pair math (host build, kernel variant 2) against the
oracle, CPU only: model sizes from centimetres to 100 km, prism aspect ratios up to 1:10^4,
observers anywhere / almost on vertices, edges and faces / far away / exactly in face planes.

Bar: max|ours - oracle| <= 1e-9 * max|field|. Where that fails the reference's OWN rounding noise
is the cause (thin or distant prisms: e*log(...) terms many orders above the field); the test then
compares both with a 40-digit evaluation of the same closed form at the observer where the two
differ most and demands that the product is either within the bar of that truth or no further
from it than 4x the oracle's own error there (the observer is picked where the product
deviates most, which biases the comparison against the product).
"""

import numpy as np
import pytest

import oracle as O
from _common import GRAVITY_FIELDS, TOL, harness_prism

G = 6.6743e-11


def _mp_kernel_sum(field, E, N, U, prism, rho):
    import mpmath as mp

    def s_atan2(y, x):
        if x != 0:
            return mp.atan(y / x)
        return mp.pi / 2 if y > 0 else (-mp.pi / 2 if y < 0 else mp.mpf(0))

    def s_log(x, y, z, r):
        if r == 0:
            return mp.mpf(0)
        if x < 0:
            if y == 0 and z == 0:
                return -mp.log(-2 * x)
            return mp.log((y * y + z * z) / (r - x))
        return mp.log(x + r)

    w, e, s, n, b, t = [mp.mpf(float(v)) for v in prism]
    E, N, U = mp.mpf(float(E)), mp.mpf(float(N)), mp.mpf(float(U))
    tot = mp.mpf(0)
    for i, x in enumerate((e - E, w - E)):
        for j, y in enumerate((n - N, s - N)):
            for k, z in enumerate((t - U, b - U)):
                r = mp.sqrt(x * x + y * y + z * z)
                v = {
                    "potential": lambda: (x * y * s_log(z, x, y, r) + y * z * s_log(x, y, z, r)
                                          + x * z * s_log(y, x, z, r) - x * x / 2 * s_atan2(z * y, x * r)
                                          - y * y / 2 * s_atan2(z * x, y * r) - z * z / 2 * s_atan2(x * y, z * r)),
                    "g_e": lambda: -(y * s_log(z, x, y, r) + z * s_log(y, x, z, r) - x * s_atan2(y * z, x * r)),
                    "g_n": lambda: -(z * s_log(x, y, z, r) + x * s_log(z, x, y, r) - y * s_atan2(z * x, y * r)),
                    "g_z": lambda: -(x * s_log(y, x, z, r) + y * s_log(x, y, z, r) - z * s_atan2(x * y, z * r)),
                    "g_ee": lambda: -s_atan2(y * z, x * r),
                    "g_nn": lambda: -s_atan2(x * z, y * r),
                    "g_zz": lambda: -s_atan2(x * y, z * r),
                    "g_en": lambda: s_log(z, x, y, r),
                    "g_ez": lambda: s_log(y, x, z, r),
                    "g_nz": lambda: s_log(x, y, z, r),
                }[field]()
                tot += (-1) ** (i + j + k) * v
    return mp.mpf("6.6743e-11") * mp.mpf(float(rho)) * tot


def _case(rng, trial):
    scale = 10.0 ** rng.uniform(-2, 5)
    aspect = 10.0 ** rng.uniform(-2, 2, 3)
    P, N = 30, 40
    c = rng.uniform(-1, 1, (P, 3)) * scale
    h = rng.uniform(0.01, 0.3, (P, 3)) * scale * aspect / np.max(aspect)
    prisms = np.stack([c[:, 0] - h[:, 0], c[:, 0] + h[:, 0], c[:, 1] - h[:, 1], c[:, 1] + h[:, 1],
                       c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], 1)  # fmt: skip
    density = rng.uniform(-3000, 3000, P)
    mode = trial % 4
    if mode == 0:  # anywhere, including inside prisms
        obs = rng.uniform(-1.3, 1.3, (N, 3)) * scale
    elif mode == 1:  # offsets of 1e-12 .. 1e-3 of the model size from vertices / edges / faces
        k = rng.integers(0, P, N)
        obs = np.empty((N, 3))
        for q in range(N):
            corner = prisms[k[q], [rng.integers(0, 2), 2 + rng.integers(0, 2), 4 + rng.integers(0, 2)]]
            obs[q] = corner + rng.choice([-1, 1], 3) * scale * 10.0 ** rng.uniform(-12, -3, 3) * rng.integers(0, 2, 3)
    elif mode == 2:  # far field, 3 to 1000 model sizes away
        obs = rng.uniform(-1, 1, (N, 3)) * scale * 10.0 ** rng.uniform(0.5, 3)
    else:  # exactly in the plane of a face of some prism
        k = rng.integers(0, P, N)
        obs = rng.uniform(-1.3, 1.3, (N, 3)) * scale
        for q in range(N):
            ax = rng.integers(0, 3)
            obs[q, ax] = prisms[k[q], 2 * ax + rng.integers(0, 2)]
    return (obs[:, 0].copy(), obs[:, 1].copy(), obs[:, 2].copy()), prisms, density, mode


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_geometry_against_oracle_and_high_precision(seed):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(seed)
    n_noise_cases = 0
    worse = []
    for trial in range(12):
        coords, prisms, density, mode = _case(rng, trial)
        prm = np.zeros((prisms.shape[0], 3))
        prm[:, 0] = G * density
        fused = {"acc3": GRAVITY_FIELDS[1:4], "tensor6": GRAVITY_FIELDS[4:]}
        results = {f: harness_prism(f, 2, coords, prisms, prm)[0][0] for f in GRAVITY_FIELDS}
        for name, members in fused.items():
            out = harness_prism(name, 2, coords, prisms, prm)[0]
            for k, f in enumerate(members):
                results[f"{name}:{f}"] = out[k]
        for key, got in results.items():
            f = key.split(":")[-1]
            want = O.prism_gravity_si(coords, prisms, density, f)
            assert np.array_equal(np.isnan(got), np.isnan(want)), (trial, key)
            if not np.isfinite(want).any():
                continue
            scale = np.nanmax(np.abs(want))
            diff = np.where(np.isnan(want), 0.0, np.abs(got - want))
            if np.max(diff) <= TOL * scale:
                continue
            # beyond the bar: the reference's own rounding noise must be the reason
            n_noise_cases += 1
            iw = int(np.argmax(diff))
            truth = sum(_mp_kernel_sum(f, coords[0][iw], coords[1][iw], coords[2][iw], prisms[j], density[j])
                        for j in range(prisms.shape[0]))
            err_ours = float(abs(mp.mpf(float(got[iw])) - truth))
            err_oracle = float(abs(mp.mpf(float(want[iw])) - truth))
            worse.append(err_ours > err_oracle)
            assert err_ours <= max(4.0 * err_oracle, TOL * scale), (trial, mode, key, err_ours, err_oracle, float(scale))
    assert n_noise_cases < 200  # sanity: most cases meet the bar outright
    print(f"noise cases: {n_noise_cases}, product further from truth than the oracle in {sum(worse)}")
