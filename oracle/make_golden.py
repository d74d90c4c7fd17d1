"""
oracle/make_golden.py -- TEST INFRASTRUCTURE. Run in the build container only:

    python oracle/make_golden.py

Generates tests/golden/*.npz by executing the reference's UNMODIFIED wrappers
and jitted loops from /root/reference through oracle/ref_shim.py (restated
choclo kernels injected), and by reading the golden vectors the reference's
own tests hold (test/data/sample_point_gravity.csv). The fixtures are small
(float64 arrays of a few hundred values) and travel to the GPU box, where
/root/reference does not exist.
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
GRAVITY_FIELDS = ("potential", "g_e", "g_n", "g_z", "g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz")


def random_prisms(rng, n, region, zrange, half):
    c = np.stack(
        [rng.uniform(region[0], region[1], n), rng.uniform(region[2], region[3], n),
         rng.uniform(zrange[0], zrange[1], n)], axis=1)  # fmt: skip
    h = rng.uniform(half[0], half[1], (n, 3))
    return np.stack([c[:, 0] - h[:, 0], c[:, 0] + h[:, 0], c[:, 1] - h[:, 1], c[:, 1] + h[:, 1],
                     c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], axis=1)  # fmt: skip


def singular_suite(prism):
    """Observers on the 8 vertices, 12 edge mid-points, 6 face centres, 6 off-centre
    face points, points on edge extensions and generic points around one prism."""
    w, e, s, n, b, t = prism
    xs, ys, zs = (w, e), (s, n), (b, t)
    xm, ym, zm = (w + e) / 2, (s + n) / 2, (b + t) / 2
    pts = []
    pts += [(x, y, z) for x in xs for y in ys for z in zs]                  # vertices
    pts += [(xm, y, z) for y in ys for z in zs]                              # easting edges
    pts += [(x, ym, z) for x in xs for z in zs]                              # northing edges
    pts += [(x, y, zm) for x in xs for y in ys]                              # upward edges
    pts += [(x, ym, zm) for x in xs] + [(xm, y, zm) for y in ys] + [(xm, ym, z) for z in zs]
    qx, qy, qz = w + 0.3 * (e - w), s + 0.7 * (n - s), b + 0.2 * (t - b)
    pts += [(x, qy, qz) for x in xs] + [(qx, y, qz) for y in ys] + [(qx, qy, z) for z in zs]
    ext = 0.5 * (e - w)
    pts += [(e + ext, n, t), (w - ext, s, b), (e, n + ext, t), (w, s - ext, b),
            (e, n, t + ext), (w, s, b - ext)]                               # edge extensions
    pts += [(xm, ym, t + ext), (xm, ym, zm), (e + ext, ym, zm), (qx, qy, qz),  # above, inside
            (e + ext, n + ext, t), (xm, n + ext, t), (e, ym, t + ext)]          # face planes outside
    a = np.array(pts, dtype=np.float64)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def main():
    ref = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # ---- 1. the reference's own golden CSV for point potential
    csv = "/root/reference/test/data/sample_point_gravity.csv"
    e, n, u, pot = np.loadtxt(csv, delimiter=",", unpack=True)
    np.savez(os.path.join(OUT, "point_potential_csv.npz"), easting=e, northing=n, upward=u,
             potential=pot, point=np.array([0.0, 0.0, 0.0]), mass=np.array([5000.0]))

    # ---- 2. prism gravity, random model, observers above and anywhere
    prisms = random_prisms(rng, 60, (-5e3, 5e3, -5e3, 5e3), (-3e3, -300), (50, 600))
    density = rng.uniform(-600, 600, 60)
    density[7] = 0.0                      # null prisms exercise the discard
    prisms[11, 1] = prisms[11, 0]
    obs_above = (rng.uniform(-5e3, 5e3, 80), rng.uniform(-5e3, 5e3, 80), rng.uniform(0, 1e3, 80))
    obs_any = (rng.uniform(-5e3, 5e3, 80), rng.uniform(-5e3, 5e3, 80), rng.uniform(-4e3, 1e3, 80))
    data = dict(prisms=prisms, density=density,
                above_e=obs_above[0], above_n=obs_above[1], above_u=obs_above[2],
                any_e=obs_any[0], any_n=obs_any[1], any_u=obs_any[2])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in GRAVITY_FIELDS:
            data[f"above_{f}"] = ref.gravity.prism_gravity(obs_above, prisms, density, f)
            data[f"any_{f}"] = ref.gravity.prism_gravity(obs_any, prisms, density, f)
    np.savez(os.path.join(OUT, "prism_gravity_random.npz"), **data)

    # ---- 3. singular suite around two prisms (one shares a face with the other)
    p0 = np.array([-30.0, 50.0, -20.0, 40.0, -80.0, -10.0])
    p1 = np.array([50.0, 120.0, -20.0, 40.0, -80.0, -10.0])   # east neighbour of p0
    two = np.stack([p0, p1])
    rho2 = np.array([2670.0, -300.0])
    se, sn, su = singular_suite(p0)
    data = dict(prisms=two, density=rho2, easting=se, northing=sn, upward=su)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in GRAVITY_FIELDS:
            data[f"one_{f}"] = ref.gravity.prism_gravity((se, sn, su), p0, rho2[0], f)
            data[f"two_{f}"] = ref.gravity.prism_gravity((se, sn, su), two, rho2, f)
        mag = (np.array([1.3, -0.4]), np.array([0.2, 2.1]), np.array([-0.7, 0.9]))
        data["mag"] = np.stack(mag)
        b = ref.magnetic.prism_magnetic((se, sn, su), two, mag, "b")
        data["two_b"] = np.stack(b)
        for f in ("b_e", "b_n", "b_u"):
            data[f"two_{f}"] = ref.magnetic.prism_magnetic((se, sn, su), two, mag, f)
    np.savez(os.path.join(OUT, "prism_singular_suite.npz"), **data)

    # ---- 4. prism magnetic, random
    M = tuple(rng.normal(size=60) for _ in range(3))
    M[0][3] = M[1][3] = M[2][3] = 0.0     # null magnetization
    data = dict(prisms=prisms, mag=np.stack(M), easting=obs_any[0], northing=obs_any[1],
                upward=obs_any[2])
    data["b"] = np.stack(ref.magnetic.prism_magnetic(obs_any, prisms, M, "b"))
    for f in ("b_e", "b_n", "b_u"):
        data[f] = ref.magnetic.prism_magnetic(obs_any, prisms, M, f)
    np.savez(os.path.join(OUT, "prism_magnetic_random.npz"), **data)

    # ---- 5. point gravity, cartesian (all fields + aliases) and spherical
    pts = (rng.uniform(-5e3, 5e3, 70), rng.uniform(-5e3, 5e3, 70), rng.uniform(-3e3, -100, 70))
    masses = rng.uniform(1e6, 1e9, 70)
    data = dict(points=np.stack(pts), masses=masses, easting=obs_above[0], northing=obs_above[1],
                upward=obs_above[2])
    for f in GRAVITY_FIELDS + ("g_ne", "g_ze", "g_zn"):
        data[f] = ref.point.point_gravity(obs_above, pts, masses, f)
    lon = rng.uniform(-20, 20, 70)
    lat = rng.uniform(-30, 30, 70)
    rad = rng.uniform(6.2e6, 6.35e6, 70)
    olon, olat, orad = rng.uniform(-25, 25, 80), rng.uniform(-35, 35, 80), rng.uniform(6.4e6, 6.6e6, 80)
    data.update(sph_points=np.stack([lon, lat, rad]), sph_obs=np.stack([olon, olat, orad]))
    for f in ("potential", "g_z"):
        data[f"sph_{f}"] = ref.point.point_gravity((olon, olat, orad), (lon, lat, rad), masses, f,
                                                   coordinate_system="spherical")
    np.savez(os.path.join(OUT, "point_gravity_random.npz"), **data)

    # ---- 6. prism layer: the reference's jitted layer loop on raw arrays
    ne, nn = 9, 7
    east_c = np.linspace(-2e3, 2e3, ne)
    north_c = np.linspace(-1.5e3, 1.5e3, nn)
    surface = rng.uniform(-200, 600, (nn, ne))
    reference = 0.0
    top = np.maximum(surface, reference)
    bottom = np.minimum(surface, reference)
    dens = np.where(surface >= 0, 2670.0, -1630.0)
    dens[2, 3] = np.nan
    dens[4, 1] = 0.0
    top[5, 5] = np.nan
    bottom[1, 6] = np.nan
    top[3, 3] = bottom[3, 3] + 5.0           # thin prism (threshold test)
    lobs = (rng.uniform(-2.5e3, 2.5e3, 50), rng.uniform(-2e3, 2e3, 50), rng.uniform(700, 1500, 50))
    data = dict(east_c=east_c, north_c=north_c, top=top, bottom=bottom, density=dens,
                easting=lobs[0], northing=lobs[1], upward=lobs[2])
    for thr_name, thr in (("thr0", 0.0), ("thr10", 10.0)):
        for f in GRAVITY_FIELDS:
            res = np.zeros(50)
            ref.layer._forward_gravity_prism_layer_parallel(
                lobs, east_c, north_c, bottom, top, dens, ref.gravity.FIELDS[f], res, thr, None)
            if f in ("g_z", "g_ez", "g_nz"):
                res *= -1
            if f in ("g_e", "g_n", "g_z"):
                res *= 1e5
            if f in ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz"):
                res *= 1e9
            data[f"{thr_name}_{f}"] = res
    np.savez(os.path.join(OUT, "prism_layer.npz"), **data)

    # ---- 7. equivalent sources predict: the reference's own loop + distance
    greens = ref_shim.greens_func_cartesian()
    coefs = rng.normal(size=70)
    res = np.zeros(80)
    ref.eqs_utils.predict_numba_parallel(obs_above, pts, coefs, res, greens)
    jac = np.zeros((80, 70))
    ref.eqs_utils.jacobian_numba_parallel(obs_above, pts, jac, greens)
    np.savez(os.path.join(OUT, "eqs_predict.npz"), points=np.stack(pts), coefs=coefs,
             easting=obs_above[0], northing=obs_above[1], upward=obs_above[2], predicted=res,
             jacobian=jac)
    # ---- 8. dipole magnetic (reference wrapper + loops, restated choclo.dipole kernels)
    dip = (rng.uniform(-5e3, 5e3, 70), rng.uniform(-5e3, 5e3, 70), rng.uniform(-3e3, -100, 70))
    moments = tuple(rng.normal(size=70) * 1e6 for _ in range(3))
    data = dict(dipoles=np.stack(dip), moments=np.stack(moments), easting=obs_above[0],
                northing=obs_above[1], upward=obs_above[2])
    data["b"] = np.stack(ref.dipole.dipole_magnetic(obs_above, dip, moments, "b"))
    for f in ("b_e", "b_n", "b_u"):
        data[f] = ref.dipole.dipole_magnetic(obs_above, dip, moments, f)
    np.savez(os.path.join(OUT, "dipole_magnetic.npz"), **data)

    # ---- 9. spherical equivalent sources predict: the reference's loop + distance_spherical
    greens_sph = ref_shim.greens_func_spherical()
    res = np.zeros(80)
    ref.eqs_utils.predict_numba_parallel((olon, olat, orad), (lon, lat, rad), coefs, res, greens_sph)
    np.savez(os.path.join(OUT, "eqs_predict_spherical.npz"), points=np.stack([lon, lat, rad]),
             coefs=coefs, obs=np.stack([olon, olat, orad]), predicted=res)
    print("golden fixtures written to", OUT)
    for name in sorted(os.listdir(OUT)):
        print(" ", name, os.path.getsize(os.path.join(OUT, name)), "bytes")


if __name__ == "__main__":
    main()
