#!/usr/bin/env python
"""
bench.py -- throughput of the pairwise forward-modelling hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload layer_gz|c1_gz|tensor|mag_b|eqs] [--scaling weak|strong]

A "step" is one pass of the hot path over one batch of synthetic input. The
default workload is BASELINE.json configs[1]: `prism_layer.gravity()` (g_z) of
a 500x500 topography layer (250k prisms) on a 500x500 observation grid at
1 km height = 6.25e10 prism-observer pairs per step and per GPU.

Prints ONE JSON line (rank 0):
  value        pair evaluations / s, whole job, inputs resident in HBM, CUDA
               events on the launching stream, max over ranks
  e2e          the same metric through the public API on host (numpy) buffers,
               host<->device copies inside the timed region
  roofline     FP64-vector-pipe roofline of the dominant kernel: algorithmic
               flops (2 x I_pair of SURVEY 8d) per second over the FP64 FMA
               peak measured on this device by the library's DFMA probe
  cpu_baseline the CPU oracle (port of the reference's loop) on the host cores
`--impl reference` times the reference's CPU algorithm (oracle port, OpenMP on
all host cores, the reference's prange-over-observers loop shape) on a bounded
observer sample of the same workload.
"""

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# FP64-pipe instruction equivalents per pair of the REFERENCE algorithm
I_PAIR = {"layer_gz": 1068, "c1_gz": 1068, "tensor": 2020, "mag_b": 2100, "eqs": 26.5,
          # tesseroids: one stack pop (7 sin/cos, 2 acos, the distance to the
          # centre, 3 divisions) ~ 390 and one 2x2x2 quadrature leaf ~ 456 FP64 instructions of the
          # reference algorithm; an unsplit pair is one pop + one leaf, near pairs cost more (the
          # measured leaves per pair of the CPU sample are reported next to the rate)
          "tess_gz": 390 + 456}
NOMINAL_FP64_FLOPS = 148 * 64 * 2 * 1.965e9  # 37.2 TFLOP/s
# dram__bytes_read.sum + dram__bytes_write.sum of ONE prism_kernel launch of the default workload
# at full size (profiles/r1_traffic_layer_gz_full_launch.csv; 22.6 MB read + 1.5 MB written;
# algorithmic: 6 MB observers + 16 MB packed prisms + 2 MB result). L2 serves 31.8 GB of tile loads.
DRAM_TRAFFIC_PER_LAUNCH = {"layer_gz": 22625024 + 1515776}
# executed thread-instructions per pair of the CURRENT kernels, from the committed ncu source
# pages (profiles/r1_opmix_*_final.txt): (FP64 pipe, all other pipes)
EXECUTED_PER_PAIR = {"layer_gz": (224.7, 147.4), "c1_gz": (224.7, 147.4), "tensor": (284.1, 171.6),
                     "eqs": (12.0, 5.1)}


# ------------------------------------------------------------------ workloads
def make_workload(name, n_obs, n_src, rank):
    """Synthetic inputs (SURVEY 8d seeds). Returns dict of host float64 arrays + meta."""
    from _common import config1, layer_config2, random_prisms

    if name == "layer_gz":
        coords, east_c, north_c, bottom, top, density = layer_config2()
        # weak scaling: every rank owns its own 500x500 observation grid (a different height)
        coords = (coords[0], coords[1], coords[2] + 25.0 * rank)
        if n_obs:  # profiling runs: a strided subset of the observation grid
            pick = np.linspace(0, coords[0].size - 1, n_obs).astype(np.int64)
            coords = tuple(np.ascontiguousarray(c[pick]) for c in coords)
        return dict(kind="layer", coords=coords, east_c=east_c, north_c=north_c, bottom=bottom,
                    top=top, density=density, n_src=east_c.size * north_c.size, mask=1 << 3,
                    nf=1, desc="prism_layer.gravity g_z, 500x500 layer (250k prisms, 1% NaN, "
                    "1% zero density) x 500x500 grid at 1 km")  # fmt: skip
    if name == "c1_gz":
        coords, prisms, density = config1(n_src or 10_000, n_obs or 10_000, seed=1)
        return dict(kind="prism", coords=coords, prisms=prisms, density=density,
                    n_src=prisms.shape[0], mask=1 << 3, nf=1,
                    desc="prism_gravity g_z, 10k random prisms x 10k observers")
    if name == "tensor":
        n_src, n_obs = n_src or 1_000_000, n_obs or 65_536
        coords, prisms, density = config1(n_src, n_obs, seed=3 + 100 * rank, scale=10.0)
        return dict(kind="prism", coords=coords, prisms=prisms, density=density, n_src=n_src,
                    mask=0x3F0, nf=6,
                    desc=f"prism_gravity 6 tensor components fused, {n_src} prisms x {n_obs} observers")
    if name == "mag_b":
        n_src, n_obs = n_src or 200_000, n_obs or 131_072
        coords, prisms, _ = config1(n_src, n_obs, seed=4 + 100 * rank, scale=4.0)
        rng = np.random.default_rng(4)
        mag = tuple(rng.normal(size=n_src) for _ in range(3))
        return dict(kind="mag", coords=coords, prisms=prisms, mag=mag, n_src=n_src, mask=7, nf=3,
                    desc=f"prism_magnetic b, {n_src} prisms x {n_obs} observers")
    if name == "eqs":
        n_src, n_obs = n_src or 4_000_000, n_obs or 262_144
        rng = np.random.default_rng(5 + 100 * rank)
        side = int(np.ceil(np.sqrt(n_src)))
        gx, gy = np.meshgrid(np.arange(side), np.arange(side))
        pe = (gx.ravel()[:n_src] + rng.uniform(-0.3, 0.3, n_src)) * 500.0
        pn = (gy.ravel()[:n_src] + rng.uniform(-0.3, 0.3, n_src)) * 500.0
        pu = np.full(n_src, -3000.0)
        coefs = rng.normal(size=n_src)
        coords = (rng.uniform(0, side * 500.0, n_obs), rng.uniform(0, side * 500.0, n_obs),
                  rng.uniform(0, 500.0, n_obs))  # fmt: skip
        return dict(kind="eqs", coords=coords, points=(pe, pn, pu), coefs=coefs, n_src=n_src,
                    mask=1, nf=1,
                    desc=f"EquivalentSources.predict (sum coef/r), {n_src} sources x {n_obs} observers")
    if name == "tess_gz":
        # a 2 x 2 degree global layer of tesseroids (topography-like tops) seen from 10 km above
        # the reference sphere: most pairs are far (one leaf), the ones below each observer split
        n_obs = n_obs or 65_536
        rng = np.random.default_rng(6 + 100 * rank)
        R = 6371008.771415059
        lon_c, lat_c = np.meshgrid(np.arange(-179.0, 180.0, 2.0), np.arange(-89.0, 90.0, 2.0))
        top = R + 2e3 * np.sin(np.radians(3 * lon_c)) * np.cos(np.radians(2 * lat_c)) - 3e3
        tess = np.stack([lon_c.ravel() - 1, lon_c.ravel() + 1, lat_c.ravel() - 1, lat_c.ravel() + 1,
                         np.full(lon_c.size, R - 30e3), top.ravel()], axis=1)  # fmt: skip
        density = rng.uniform(2500, 3300, lon_c.size)
        coords = (rng.uniform(-180, 180, n_obs), np.degrees(np.arcsin(rng.uniform(-1, 1, n_obs))),
                  np.full(n_obs, R + 10e3))  # fmt: skip
        return dict(kind="tess", coords=coords, tesseroids=np.ascontiguousarray(tess),
                    density=density, n_src=tess.shape[0], mask=1 << 3, nf=1,
                    desc=f"tesseroid_gravity g_z, {tess.shape[0]} tesseroids (2x2 degree global layer) "
                    f"x {n_obs} observers at 10 km")  # fmt: skip
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------- clock sampling
class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}  # fmt: skip


# ------------------------------------------------------------------ reference arm
def cpu_rate(wl, n_obs_sample, nthreads):
    """Time the oracle (port of the reference's CPU loop) on the first n_obs_sample observers."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O  # bench.py's cpu_baseline / reference legs may execute the oracle

    sub = tuple(np.ascontiguousarray(c[:n_obs_sample]) for c in wl["coords"])
    t0 = time.perf_counter()
    if wl["kind"] == "layer":
        O.prism_layer_gravity(sub, wl["east_c"], wl["north_c"], wl["bottom"], wl["top"],
                              wl["density"], "g_z", nthreads=nthreads)
    elif wl["kind"] == "prism":
        fields = ("g_z",) if wl["nf"] == 1 else ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz")
        for f in fields:  # the reference computes one field per call
            O.prism_gravity(sub, wl["prisms"], wl["density"], f, nthreads=nthreads)
    elif wl["kind"] == "mag":
        O.prism_magnetic(sub, wl["prisms"], wl["mag"], "b", nthreads=nthreads)
    elif wl["kind"] == "tess":
        O.tesseroid_gravity(sub, wl["tesseroids"], wl["density"], "g_z", nthreads=nthreads)
    else:
        O.eqs_predict(sub, wl["points"], wl["coefs"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return n_obs_sample * wl["n_src"] / dt, dt


def size_cpu_sample(wl, target_s, nthreads):
    """Observer sample sized for ~target_s seconds of CPU work (probe first)."""
    n_total = wl["coords"][0].size
    probe = max(nthreads, min(n_total, int(4e7 / wl["n_src"]) + 1))
    rate, _ = cpu_rate(wl, probe, nthreads)  # also warms the thread pool
    n = int(rate * target_s / wl["n_src"])
    n = max(nthreads, min(n_total, n))
    return max(1, n // nthreads * nthreads) if n >= nthreads else n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    O.build()
    wl = make_workload(args.workload, args.n_obs, args.n_src, 0)
    nthreads = O.max_threads()
    n_sample = size_cpu_sample(wl, args.cpu_seconds, nthreads)
    for _ in range(args.warmup):
        cpu_rate(wl, n_sample, nthreads)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_rate(wl, n_sample, nthreads)
        t_total += dt
    value = args.steps * n_sample * wl["n_src"] / t_total
    line = {
        "impl": "reference", "metric": "prism-observer pair evals/sec", "value": value,
        "unit": "pair/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload},
        "cpu_baseline": {
            "value": value, "unit": "pair/s", "cores": nthreads, "kind": "port",
            "sample": f"first {n_sample} of {wl['coords'][0].size} observers x all {wl['n_src']} "
                      "sources per step; oracle/choclo_port.c (C restatement of the reference's "
                      "numba prange-over-observers loop + choclo kernels), OpenMP, all host threads",
        },
        "e2e": {"value": value, "unit": "pair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import harmonica_b200 as hb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    lib = hb._lib.load()
    hb.init([local])

    wl = make_workload(args.workload, args.n_obs, args.n_src, 0 if args.shard == "sources" else rank)
    coords = wl["coords"]
    if args.scaling == "strong" and world > 1:
        n = coords[0].size
        lo, hi = n * rank // world, n * (rank + 1) // world
        coords = tuple(np.ascontiguousarray(c[lo:hi]) for c in coords)
    n_obs, n_src, nf = coords[0].size, wl["n_src"], wl["nf"]
    pairs_per_step_rank = float(n_obs) * float(n_src)

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    oe, on, ou = (t(c) for c in coords)
    out = torch.empty((nf, n_obs), dtype=torch.float64, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    src_sharded = args.shard == "sources" and world > 1
    if src_sharded:
        if wl["kind"] != "eqs":
            raise SystemExit("--shard sources is implemented for the eqs workload")
        # BASELINE config 5: every rank owns a slice of the sources and ALL observers; the
        # partial fields are summed with an NCCL reduce (float64) inside the timed region
        s_lo, s_hi = n_src * rank // world, n_src * (rank + 1) // world
        wl["points"] = tuple(np.ascontiguousarray(p[s_lo:s_hi]) for p in wl["points"])
        wl["coefs"] = np.ascontiguousarray(wl["coefs"][s_lo:s_hi])
        n_src = s_hi - s_lo
        pairs_per_step_rank = float(n_obs) * float(n_src)
    if wl["kind"] == "eqs":
        ws_bytes = lib.hb200_point_ws_bytes(n_obs, n_src)
    elif wl["kind"] == "tess":
        ws_bytes = lib.hb200_tesseroid_ws_bytes(n_obs, n_src)
    else:
        ws_bytes = lib.hb200_prism_ws_bytes(n_obs, n_src, nf)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731

    if wl["kind"] == "layer":
        d = {k: t(wl[k]) for k in ("east_c", "north_c", "bottom", "top", "density")}

        def step_dev():
            return lib.hb200_prism_layer_gravity_dev(
                P(oe), P(on), P(ou), n_obs, P(d["east_c"]), wl["east_c"].size, P(d["north_c"]),
                wl["north_c"].size, P(d["bottom"]), P(d["top"]), P(d["density"]), 0.0, wl["mask"],
                P(out), P(flags), P(ws), ws_bytes, stream)  # fmt: skip

        def step_host():
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return hb.prism_layer_gravity(coords, wl["east_c"], wl["north_c"], wl["bottom"],
                                              wl["top"], wl["density"], "g_z")
        h2d = 8 * (3 * n_obs + wl["east_c"].size + wl["north_c"].size + 3 * n_src)
        launches_per_step = 2  # pack_layer_kernel + prism_kernel
    elif wl["kind"] == "prism":
        pr, rho = t(wl["prisms"]), t(wl["density"])

        def step_dev():
            return lib.hb200_prism_gravity_dev(P(oe), P(on), P(ou), n_obs, P(pr), P(rho), n_src,
                                               wl["mask"], P(out), P(flags), P(ws), ws_bytes, stream)
        fields = "g_z" if nf == 1 else ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz")

        def step_host():
            return hb.prism_gravity(coords, wl["prisms"], wl["density"], fields, disable_checks=True)
        h2d = 8 * (3 * n_obs + 7 * n_src)
        launches_per_step = 2  # pack_prisms_kernel + prism_kernel (+1 reduce when sources are chunked)
    elif wl["kind"] == "tess":
        ts, rho = t(wl["tesseroids"]), t(wl["density"])

        def step_dev():
            return lib.hb200_tesseroid_gravity_dev(P(oe), P(on), P(ou), n_obs, P(ts), P(rho), n_src,
                                                   3, 0, P(out), P(flags), P(ws), ws_bytes, stream)

        def step_host():
            return hb.tesseroid_gravity(coords, wl["tesseroids"], wl["density"], "g_z",
                                        disable_checks=True)
        h2d = 8 * (3 * n_obs + 7 * n_src)
        launches_per_step = 2  # pack_tesseroids_kernel + tesseroid_kernel
    elif wl["kind"] == "mag":
        pr = t(wl["prisms"])
        m = [t(x) for x in wl["mag"]]

        def step_dev():
            return lib.hb200_prism_magnetic_dev(P(oe), P(on), P(ou), n_obs, P(pr), P(m[0]), P(m[1]),
                                                P(m[2]), n_src, 7, 3, P(out), P(flags), P(ws),
                                                ws_bytes, stream)

        def step_host():
            return hb.prism_magnetic(coords, wl["prisms"], wl["mag"], "b", disable_checks=True)
        h2d = 8 * (3 * n_obs + 9 * n_src)
        launches_per_step = 2
    else:
        pts = [t(x) for x in wl["points"]]
        cf = t(wl["coefs"])

        def step_dev():
            rc = lib.hb200_point_gravity_dev(P(oe), P(on), P(ou), n_obs, P(pts[0]), P(pts[1]),
                                             P(pts[2]), P(cf), n_src, 1, 0, 0, P(out), P(flags),
                                             P(ws), ws_bytes, stream)
            if src_sharded:  # reduce-sum of the partial fields over NVLink (NCCL)
                dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
            return rc

        def step_host():
            res = hb.eqs_predict(coords, wl["points"], wl["coefs"])
            if src_sharded:
                part = torch.from_numpy(res).to(dev)
                dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)
                res = part.cpu().numpy()
            return res
        h2d = 8 * (3 * n_obs + 4 * n_src)
        launches_per_step = 2 + (1 if n_obs < 512 * 148 * 120 else 0)  # + chunk reduce
    d2h = 8 * nf * n_obs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 FMA peak of this device, measured now (MEASURED_PEAKS.json has no FP64 entry)
    flops, secs = ctypes.c_double(0), ctypes.c_double(0)
    hb._lib.check(lib.hb200_fp64_peak(args.peak_iters, ctypes.byref(flops), ctypes.byref(secs)))
    fp64_peak = flops.value

    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    for _ in range(args.warmup):
        hb._lib.check(step_dev())
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    launches_before = lib.hb200_launch_count()
    for k in range(args.steps):
        l2_flush.fill_(k)  # flush L2 between timed iterations (outside the events)
        starts[k].record()
        hb._lib.check(step_dev())
        ends[k].record()
    barrier()
    gpu_launches = int(lib.hb200_launch_count() - launches_before)  # counted by the library
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    if int(flags.item()) & 2:
        raise SystemExit("zero-distance pair in the synthetic workload")

    # end to end through the public API: host numpy buffers in, host result out
    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O

        O.build()
        nthreads = O.max_threads()
        n_sample = size_cpu_sample(wl, args.cpu_seconds, nthreads)
        rate, dt = cpu_rate(wl, n_sample, nthreads)
        cpu_baseline = {
            "value": rate, "unit": "pair/s", "cores": nthreads, "kind": "port",
            "sample": f"first {n_sample} of {wl['coords'][0].size} observers x all {n_src} sources "
                      f"({dt:.1f} s); oracle/*_port.c (C restatement of the reference's loop), OpenMP "
                      "over observers like the reference's numba prange",
        }  # fmt: skip

    if rank == 0:
        pairs_total = pairs_per_step_rank * world * args.steps
        value = pairs_total / total_s
        per_gpu = value / world
        f_pair = 2.0 * I_PAIR[args.workload]
        achieved = per_gpu * f_pair / 1e12
        executed = None
        if args.workload in EXECUTED_PER_PAIR and clocks and clocks.get("sm_mhz"):
            f64, other = EXECUTED_PER_PAIR[args.workload]
            slots = 148 * 4 * clocks["sm_mhz"] * 1e6 * 32  # thread-level issue slots / s
            executed = {
                "fp64_instr_per_pair": f64, "other_instr_per_pair": other,
                "source": "ncu source-page counts of this kernel build, profiles/r1_opmix_*",
                "fp64_pipe_frac": per_gpu * f64 * 2 / fp64_peak,
                "issue_bound_frac": per_gpu * (2 * f64 + other) / slots,
                "note": "an FP64 warp instruction occupies 2 issue slots on this part: time ~ "
                        "(2*FP64 + other) / issue rate (DESIGN.md section 4); the slot rate uses "
                        "nvidia-smi's SM clock, ncu's cycle counter runs ~1.3 % faster (1.99 GHz), "
                        "so ~1.0 means: at the issue bound",
            }
        line = {
            "metric": "prism-observer pair evals/sec", "value": value, "unit": "pair/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": "strong" if src_sharded else args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "pairs_per_step_per_gpu":
                       pairs_per_step_rank, "l2": "flushed between timed iterations (256 MiB write)",
                       "kernel_variant": int(lib.hb200_get_variant()),
                       "sharding": "sources + NCCL reduce-sum" if src_sharded else
                       ("observers, no collective" if world > 1 else "single GPU")},
            "e2e": {"value": pairs_total / e2e_s, "unit": "pair/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "harmonica_b200 public API on numpy buffers (ctypes -> C ABI), blocking"},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved * 1e12 / fp64_peak,
                "frac_of_nominal": achieved * 1e12 / NOMINAL_FP64_FLOPS,
                "peak_source": "measured on this device in this run: hb200_fp64_peak DFMA probe "
                               "(MEASURED_PEAKS.json carries no FP64 entry; nominal 37.2 TFLOP/s)",
                "flops_per_pair": f_pair,
                "note": "algorithmic flops = 2 x I_pair of the REFERENCE algorithm (SURVEY 8d); "
                        "the merged-transcendental kernel executes fewer instructions per pair, "
                        "pipe utilisation is in profiles/",
                "traffic": DRAM_TRAFFIC_PER_LAUNCH.get(args.workload) if not args.n_obs else None,
                "executed": executed,
            },
            "cpu_baseline": cpu_baseline,
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="layer_gz", choices=sorted(I_PAIR))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--shard", default="observers", choices=["observers", "sources"])
    ap.add_argument("--n-obs", type=int, default=0)
    ap.add_argument("--n-src", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--peak-iters", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.cpu_seconds = min(args.cpu_seconds, 8.0)
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
