"""First timings of the rows added late in round 1 (tesseroids, device-resident EQS fits).
Run on a B200: python scripts/time_new_rows.py >> gpurun_out/new_rows_timing.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import harmonica_b200 as hb  # noqa: E402

hb.init([0])
lib = hb._lib.load()


def timed(fn, repeat=2):
    fn()
    best = 1e30
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def emit(**kw):
    print(json.dumps(kw), flush=True)


def tesseroids():
    for variant in (1, 0):
        lib.hb200_set_tesseroid_variant(variant)
        _tesseroids(variant)
    lib.hb200_set_tesseroid_variant(1)


def _tesseroids(variant):
    import bench

    for n_obs in (4096, 32768):
        wl = bench.make_workload("tess_gz", n_obs, 0, 0)
        for field in ("g_z", "potential"):
            for radial in (False, True):
                dt, _ = timed(lambda: hb.tesseroid_gravity(wl["coords"], wl["tesseroids"], wl["density"], field,
                                                           radial_adaptive_discretization=radial,
                                                           disable_checks=True), repeat=1)  # fmt: skip
                emit(row="tesseroid_gravity", kernel_variant=variant, field=field, radial=radial, n_obs=n_obs, n_tess=wl["n_src"],
                     seconds=dt, pairs_per_s=n_obs * wl["n_src"] / dt, api="numpy host API, e2e")  # fmt: skip
    # observers ON the surface of a regional model: deep splitting
    rng = np.random.default_rng(1)
    R = 6371008.771415059
    lon_c, lat_c = np.meshgrid(np.arange(-9.75, 10, 0.5), np.arange(-9.75, 10, 0.5))
    tess = np.stack([lon_c.ravel() - 0.25, lon_c.ravel() + 0.25, lat_c.ravel() - 0.25, lat_c.ravel() + 0.25,
                     np.full(lon_c.size, R - 5e3), np.full(lon_c.size, R)], axis=1)  # fmt: skip
    coords = (rng.uniform(-10, 10, 8192), rng.uniform(-10, 10, 8192), np.full(8192, R + 10.0))
    dt, _ = timed(lambda: hb.tesseroid_gravity(coords, tess, np.full(lon_c.size, 2670.0), "g_z",
                                               disable_checks=True), repeat=1)  # fmt: skip
    emit(row="tesseroid_gravity", kernel_variant=variant, case="observers 10 m above a 0.5 degree regional layer", n_obs=8192,
         n_tess=int(lon_c.size), seconds=dt, pairs_per_s=8192 * lon_c.size / dt)  # fmt: skip


def fits():
    rng = np.random.default_rng(2)
    for n in (2048, 8192):
        coords = (rng.uniform(0, 50e3, n), rng.uniform(0, 50e3, n), rng.uniform(0, 500, n))
        pts = (coords[0], coords[1], coords[2] - 1500.0)
        data = hb.eqs_predict(coords, (rng.uniform(0, 50e3, 50), rng.uniform(0, 50e3, 50), np.full(50, -5e3)),
                              rng.normal(size=50))  # fmt: skip
        for damping in (1e-3, None):
            if damping is None and n > 4096:
                continue
            dt, (coefs, path) = timed(lambda: hb.eqs_fit(coords, pts, data, None, damping, return_solver_path=True),
                                      repeat=1)  # fmt: skip
            misfit = float(np.max(np.abs(hb.eqs_predict(coords, pts, coefs) - data)) / np.max(np.abs(data)))
            emit(row="eqs_fit", n_data=n, n_sources=n, damping=damping, solver_path=path, seconds=dt, misfit=misfit)
    n = 40000
    side = int(np.sqrt(n))
    e, nn = np.meshgrid(np.linspace(0, 100e3, side), np.linspace(0, 100e3, side))
    coords = (e.ravel(), nn.ravel(), np.zeros(e.size))
    data = hb.eqs_predict(coords, (rng.uniform(0, 100e3, 80), rng.uniform(0, 100e3, 80), np.full(80, -8e3)),
                          rng.normal(size=80))  # fmt: skip
    eqs = hb.EquivalentSourcesGB(depth=3e3, damping=1e-2, window_size=20e3, random_state=0)
    t0 = time.perf_counter()
    eqs.fit(coords, data)
    dt = time.perf_counter() - t0
    emit(row="EquivalentSourcesGB.fit", n_data=int(e.size), windows=int(eqs.rmse_per_iteration_.size - 1),
         seconds=dt, rmse_first=float(eqs.rmse_per_iteration_[0]), rmse_last=float(eqs.rmse_per_iteration_[-1]))  # fmt: skip


if __name__ == "__main__":
    which = sys.argv[1:] or ["tesseroids", "fits"]
    for name in which:
        try:
            {"tesseroids": tesseroids, "fits": fits}[name]()
        except Exception as exc:  # keep going: this is a measurement script
            emit(row=name, error=f"{type(exc).__name__}: {exc}")
