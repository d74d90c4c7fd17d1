#!/usr/bin/env python
"""
bench.py -- throughput of the pairwise forward-modelling hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload layer_gz|c1_gz|tensor|mag_b|eqs|tess_gz]
                    [--scaling strong|weak] [--shard observers|sources]

A "step" is one pass of the hot path over one batch of synthetic input. The
default workload is BASELINE.json configs[1]: `prism_layer.gravity()` (g_z) of
a 500x500 topography layer (250k prisms) on a 500x500 observation grid at
1 km height. With N > 1 ranks (torchrun, one process per GPU) the default is
STRONG scaling of that fixed problem through the repo's own multi-GPU API
(harmonica_b200/distributed.py): observers sharded over the ranks, sources
replicated, the result gathered on rank 0 over NCCL inside the timed region.

Prints ONE JSON line (rank 0):
  value        pair evaluations / s, whole job, inputs resident in HBM, CUDA
               events on the launching stream, max over ranks (N > 1: kernels +
               NCCL gather)
  e2e          the same metric through the public API on host (numpy) buffers,
               host<->device copies (and the gather) inside the timed region
  roofline     FP64-vector-pipe roofline of the dominant kernel: EXECUTED FP64
               instructions (ncu source-page counts of this very kernel build,
               profiles/executed_per_pair.json) x 2 flop over the FP64 FMA peak
               measured on this device by the library's DFMA probe;
               `algorithmic` carries the SURVEY 8d convention (instructions of
               the REFERENCE algorithm), which exceeds 1 because the merged
               kernels execute fewer instructions per pair
  also         short runs of the other BASELINE configs (tensor, mag_b, eqs)
  north_star   BASELINE configs[2]: the six tensor components of 1M prisms on 1M
               observers, end to end through the (sharded) public API, 1 step
  source_sharded (N > 1) BASELINE configs[4]: EQS predict with the 4M sources
               sharded over the ranks and an NCCL float64 reduce-sum, short run
  cpu_baseline the reference's CPU path on the host cores: the Numba
               parallel=True restatement (oracle/numba_loops.py) and the
               C/OpenMP port (oracle/choclo_port.c), both reported
`--impl reference` times the reference's CPU algorithm (Numba prange over
observers, all host cores, OMP_NUM_THREADS of torchrun ignored) on a bounded
observer sample of the same workload.
"""

import argparse
import ctypes
import hashlib
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

METRIC = "prism-observer pair evals/sec"
# FP64-pipe instruction equivalents per pair of the REFERENCE algorithm
I_PAIR = {"layer_gz": 1068, "c1_gz": 1068, "tensor": 2020, "mag_b": 2100, "eqs": 26.5,
          # tesseroids: one stack pop ~ 390 and one 2x2x2 quadrature leaf
          # ~ 456 FP64 instructions of the reference algorithm per unsplit pair
          "tess_gz": 390 + 456}  # fmt: skip
NOMINAL_FP64_FLOPS = 148 * 64 * 2 * 1.965e9  # 37.2 TFLOP/s
EXECUTED_FILE = os.path.join(ROOT, "profiles", "executed_per_pair.json")
CSRC = os.path.join(ROOT, "harmonica_b200", "csrc")


def host_cores():
    """Host threads this process may use (torchrun's OMP_NUM_THREADS=1 is ignored on purpose)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def csrc_sha16(files):
    h = hashlib.sha256()
    for name in sorted(files):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def executed_for(workload):
    """ncu-measured executed instructions per pair of the CURRENT kernel build, or a note saying
    that the committed numbers belong to another build (then nothing is printed as measured)."""
    try:
        with open(EXECUTED_FILE) as f:
            table = json.load(f)
    except OSError:
        return None, "profiles/executed_per_pair.json missing"
    entry = table.get(workload)
    if not entry:
        return None, f"no ncu capture recorded for workload {workload}"
    sha = csrc_sha16(entry["files"])
    if sha != entry["sha16"]:
        return None, (f"stale: {entry['source']} was captured on kernel sources {entry['sha16']}, "
                      f"this build is {sha}")
    return entry, None


# ------------------------------------------------------------------ workloads
def make_workload(name, n_obs=0, n_src=0, shift=0):
    """Synthetic inputs (SURVEY 8d seeds). Returns dict of host float64 arrays + meta."""
    from _common import config1, layer_config2

    if name == "layer_gz":
        coords, east_c, north_c, bottom, top, density = layer_config2()
        coords = (coords[0], coords[1], coords[2] + 25.0 * shift)
        if n_obs:  # profiling runs: a strided subset of the observation grid
            pick = np.linspace(0, coords[0].size - 1, n_obs).astype(np.int64)
            coords = tuple(np.ascontiguousarray(c[pick]) for c in coords)
        skipped = (density == 0) | np.isnan(density) | np.isnan(top) | np.isnan(bottom) | (top - bottom < 0)
        n_eval = int(east_c.size * north_c.size - skipped.sum())
        return dict(kind="prism_layer", coords=coords, n_src=east_c.size * north_c.size, n_eval=n_eval,
                    sources=dict(easting=east_c, northing=north_c, bottom=bottom, top=top, density=density),
                    fields="g_z", nf=1, launches=2,
                    desc="prism_layer.gravity g_z, 500x500 layer (250k prisms, 1% NaN, 1% zero density) "
                         "x 500x500 grid at 1 km")  # fmt: skip
    if name == "c1_gz":
        coords, prisms, density = config1(n_src or 10_000, n_obs or 10_000, seed=1)
        return dict(kind="prism_gravity", coords=coords, sources=dict(prisms=prisms, density=density),
                    n_src=prisms.shape[0], n_eval=prisms.shape[0], fields="g_z", nf=1, launches=3,
                    desc="prism_gravity g_z, 10k random prisms x 10k observers")
    if name in ("tensor", "tensor_1m"):
        n_src = n_src or 1_000_000
        n_obs = n_obs or (1_000_000 if name == "tensor_1m" else 65_536)
        coords, prisms, density = config1(n_src, n_obs, seed=3 + 100 * shift, scale=10.0)
        return dict(kind="prism_gravity", coords=coords, sources=dict(prisms=prisms, density=density),
                    n_src=n_src, n_eval=n_src, nf=6, launches=2,
                    fields=("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz"),
                    desc=f"prism_gravity 6 tensor components fused, {n_src} prisms x {n_obs} observers")
    if name == "mag_b":
        n_src, n_obs = n_src or 200_000, n_obs or 131_072
        coords, prisms, _ = config1(n_src, n_obs, seed=4 + 100 * shift, scale=4.0)
        rng = np.random.default_rng(4)
        mag = tuple(rng.normal(size=n_src) for _ in range(3))
        return dict(kind="prism_magnetic", coords=coords, sources=dict(prisms=prisms, magnetization=mag),
                    n_src=n_src, n_eval=n_src, fields="b", nf=3, launches=2,
                    desc=f"prism_magnetic b, {n_src} prisms x {n_obs} observers")
    if name == "eqs":
        n_src, n_obs = n_src or 4_000_000, n_obs or 262_144
        rng = np.random.default_rng(5 + 100 * shift)
        side = int(np.ceil(np.sqrt(n_src)))
        gx, gy = np.meshgrid(np.arange(side), np.arange(side))
        pe = (gx.ravel()[:n_src] + rng.uniform(-0.3, 0.3, n_src)) * 500.0
        pn = (gy.ravel()[:n_src] + rng.uniform(-0.3, 0.3, n_src)) * 500.0
        pu = np.full(n_src, -3000.0)
        coefs = rng.normal(size=n_src)
        coords = (rng.uniform(0, side * 500.0, n_obs), rng.uniform(0, side * 500.0, n_obs),
                  rng.uniform(0, 500.0, n_obs))  # fmt: skip
        return dict(kind="eqs_predict", coords=coords, sources=dict(points=(pe, pn, pu), coefs=coefs),
                    n_src=n_src, n_eval=n_src, fields="potential", nf=1, launches=3,
                    desc=f"EquivalentSources.predict (sum coef/r), {n_src} sources x {n_obs} observers")
    if name == "tess_gz":
        # a 2 x 2 degree global layer of tesseroids (topography-like tops) seen from 10 km above
        # the reference sphere: most pairs are far (one leaf), the ones below each observer split
        n_obs = n_obs or 65_536
        rng = np.random.default_rng(6 + 100 * shift)
        R = 6371008.771415059
        lon_c, lat_c = np.meshgrid(np.arange(-179.0, 180.0, 2.0), np.arange(-89.0, 90.0, 2.0))
        top = R + 2e3 * np.sin(np.radians(3 * lon_c)) * np.cos(np.radians(2 * lat_c)) - 3e3
        tess = np.stack([lon_c.ravel() - 1, lon_c.ravel() + 1, lat_c.ravel() - 1, lat_c.ravel() + 1,
                         np.full(lon_c.size, R - 30e3), top.ravel()], axis=1)  # fmt: skip
        density = rng.uniform(2500, 3300, lon_c.size)
        coords = (rng.uniform(-180, 180, n_obs), np.degrees(np.arcsin(rng.uniform(-1, 1, n_obs))),
                  np.full(n_obs, R + 10e3))  # fmt: skip
        return dict(kind="tess", coords=coords, tesseroids=np.ascontiguousarray(tess), density=density,
                    n_src=tess.shape[0], n_eval=tess.shape[0], fields="g_z", nf=1, launches=3,
                    desc=f"tesseroid_gravity g_z, {tess.shape[0]} tesseroids (2x2 degree global layer) "
                         f"x {n_obs} observers at 10 km")  # fmt: skip
    raise SystemExit(f"unknown workload {name}")


def config_for(wl, name, args, world):
    """The `config` object: identical in the B200 arm and the reference arm."""
    if world == 1:
        sharding = "single GPU"
    elif args.shard == "sources":
        sharding = f"sources sharded over {world} ranks, float64 reduce-sum to rank 0 (NCCL)"
    elif args.scaling == "strong":
        sharding = f"observers sharded over {world} ranks, sources replicated, gather to rank 0 (NCCL)"
    else:
        sharding = f"{world} independent replicas (weak), no collective"
    n_obs = wl["coords"][0].size
    replicas = world if (world > 1 and args.scaling == "weak" and args.shard != "sources") else 1
    return {"workload": wl["desc"], "name": name, "observers": n_obs, "sources": wl["n_src"],
            "pairs_per_step": float(n_obs) * wl["n_eval"] * replicas,
            "pairs_counted": "observers x sources that pass the reference's skip rules "
                             f"({wl['n_eval']} of {wl['n_src']})",
            "sharding": sharding,
            "l2": "GPU arm: flushed between timed iterations (256 MiB write); CPU arm: n/a"}  # fmt: skip


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


# ------------------------------------------------------------------ CPU arms
def _oracle_modules():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O  # bench.py's cpu_baseline / reference legs may execute the oracle

    O.build()
    return O


def _numba_loops(nthreads):
    """The Numba restatement, on `nthreads` threads whatever torchrun exported."""
    os.environ["NUMBA_NUM_THREADS"] = str(nthreads)
    os.environ["OMP_NUM_THREADS"] = str(nthreads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numba  # noqa: PLC0415
    import numba_loops as NL  # noqa: PLC0415

    numba.set_num_threads(min(nthreads, numba.config.NUMBA_NUM_THREADS))
    return NL


def cpu_rate(wl, n_sample, nthreads, kind):
    """Time one CPU pass (Numba loops or the C/OpenMP port) on the first n_sample observers."""
    sub = tuple(np.ascontiguousarray(c[:n_sample]) for c in wl["coords"])
    s = wl.get("sources", {})
    if kind == "numba":
        NL = _numba_loops(nthreads)
        t0 = time.perf_counter()
        if wl["kind"] == "prism_layer":
            NL.prism_layer_gravity(sub, s["easting"], s["northing"], s["bottom"], s["top"], s["density"], "g_z")
        elif wl["kind"] == "prism_gravity":
            fields = (wl["fields"],) if isinstance(wl["fields"], str) else wl["fields"]
            for f in fields:  # the reference computes one field per call
                NL.prism_gravity(sub, s["prisms"], s["density"], f)
        elif wl["kind"] == "prism_magnetic":
            NL.prism_magnetic_field(sub, s["prisms"], s["magnetization"])
        elif wl["kind"] == "eqs_predict":
            NL.eqs_predict(sub, s["points"], s["coefs"])
        else:
            raise ValueError("no Numba restatement of this workload")
    else:
        O = _oracle_modules()
        t0 = time.perf_counter()
        if wl["kind"] == "prism_layer":
            O.prism_layer_gravity(sub, s["easting"], s["northing"], s["bottom"], s["top"], s["density"],
                                  "g_z", nthreads=nthreads)
        elif wl["kind"] == "prism_gravity":
            fields = (wl["fields"],) if isinstance(wl["fields"], str) else wl["fields"]
            for f in fields:
                O.prism_gravity(sub, s["prisms"], s["density"], f, nthreads=nthreads)
        elif wl["kind"] == "prism_magnetic":
            O.prism_magnetic(sub, s["prisms"], s["magnetization"], "b", nthreads=nthreads)
        elif wl["kind"] == "tess":
            O.tesseroid_gravity(sub, wl["tesseroids"], wl["density"], "g_z", nthreads=nthreads)
        else:
            O.eqs_predict(sub, s["points"], s["coefs"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return n_sample * wl["n_eval"] / dt, dt


def size_cpu_sample(wl, target_s, nthreads, kind):
    """Observer sample sized for ~target_s seconds of CPU work (probe first; the probe also
    JIT-compiles the Numba loops / warms the thread pool)."""
    n_total = wl["coords"][0].size
    probe = max(nthreads, min(n_total, int(4e7 / wl["n_src"]) + 1))
    cpu_rate(wl, probe, nthreads, kind)  # compile / warm up
    rate, _ = cpu_rate(wl, probe, nthreads, kind)
    n = int(rate * target_s / wl["n_eval"])
    n = max(nthreads, min(n_total, n))
    return max(1, n // nthreads * nthreads) if n >= nthreads else n


def cpu_kinds(wl):
    kinds = []
    if wl["kind"] != "tess":
        try:
            import numba  # noqa: F401, PLC0415

            kinds.append("numba")
        except ImportError:
            pass
    kinds.append("port")
    return kinds


CPU_DESCRIPTION = {
    "numba": "oracle/numba_loops.py: Numba parallel=True restatement of the reference's jitted loop "
             "(prange over observers) over the restated choclo kernels; bit-identical to the "
             "reference's unmodified loops run through oracle/ref_shim.py",
    "port": "oracle/*_port.c: C restatement of the same loop, OpenMP over observers",
}  # fmt: skip


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = make_workload(args.workload, args.n_obs, args.n_src)
    nthreads = host_cores()
    kinds = cpu_kinds(wl)
    n_sample = None
    while kinds:  # Numba first; if it cannot run here (no compiler cache directory, ...) the C port
        kind = kinds[0]
        try:
            n_sample = size_cpu_sample(wl, args.cpu_seconds, nthreads, kind)
            break
        except Exception as error:  # noqa: BLE001
            print(f"reference arm: {kind} failed: {type(error).__name__}: {error}", file=sys.stderr)
            kinds = kinds[1:]
    if n_sample is None:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU implementation could run"}), flush=True)
        return
    for _ in range(args.warmup):
        cpu_rate(wl, n_sample, nthreads, kind)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_rate(wl, n_sample, nthreads, kind)
        t_total += dt
    value = args.steps * n_sample * wl["n_eval"] / t_total
    other = None
    if len(kinds) > 1:  # the other CPU implementation, one pass, for the record
        n_other = size_cpu_sample(wl, min(args.cpu_seconds, 4.0), nthreads, kinds[1])
        rate, dt = cpu_rate(wl, n_other, nthreads, kinds[1])
        other = {"value": rate, "unit": "pair/s", "cores": nthreads, "kind": kinds[1],
                 "sample": f"first {n_other} observers, one pass ({dt:.1f} s); {CPU_DESCRIPTION[kinds[1]]}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "pair/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.shard == "sources" else args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(wl, args.workload, args, world),
        "cpu_baseline": {
            "value": value, "unit": "pair/s", "cores": nthreads, "kind": kind,
            "sample": f"first {n_sample} of {wl['coords'][0].size} observers x all {wl['n_src']} "
                      f"sources per step; {CPU_DESCRIPTION[kind]}; {nthreads} host threads "
                      f"(os.sched_getaffinity; OMP_NUM_THREADS of the launcher ignored)",
            "also": other,
        },
        "e2e": {"value": value, "unit": "pair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm
class Stepper:
    """step_dev(): one pass with device-resident inputs (async on torch's current stream,
    collective included); step_host(): the end-to-end call on host numpy buffers."""

    def __init__(self, wl, rank, world, scaling, shard):
        import torch

        import harmonica_b200 as hb
        from harmonica_b200 import distributed as hbd

        self.wl, self.hb, self.world = wl, hb, world
        lib = hb._lib.load()
        self.lib = lib
        sharded = world > 1 and (scaling == "strong" or shard == "sources")
        self.sharded = sharded
        n_obs = wl["coords"][0].size
        if wl["kind"] == "tess":
            dev = torch.device("cuda", torch.cuda.current_device())
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
            self.t = [t(c) for c in wl["coords"]] + [t(wl["tesseroids"]), t(wl["density"])]
            self.out = torch.empty((1, n_obs), dtype=torch.float64, device=dev)
            self.flags_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self.ws_bytes = lib.hb200_tesseroid_ws_bytes(n_obs, wl["n_src"])
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
            self.job = None
            self.h2d = 8 * (3 * n_obs + 7 * wl["n_src"])
            self.d2h = 8 * n_obs
            self.pairs_per_step = float(n_obs) * wl["n_eval"] * (world if not sharded else 1)
            return
        group = None if sharded else "local"
        self.job = hbd.ShardedJob(wl["kind"], wl["coords"], wl["sources"], wl["fields"],
                                  shard if sharded else "observers", group=group, dst=0)
        self.job.upload()
        self.h2d = self.job.h2d_bytes
        self.d2h = self.job.d2h_bytes if (not sharded or rank == 0) else 0
        self.pairs_per_step = float(n_obs) * wl["n_eval"] * (world if not sharded else 1)

    def step_dev(self):
        import torch

        if self.job is not None:
            self.job.launch()
            return
        P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
        t = self.t
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.hb._lib.check(self.lib.hb200_tesseroid_gravity_dev(
            P(t[0]), P(t[1]), P(t[2]), t[0].numel(), P(t[3]), P(t[4]), self.wl["n_src"], 3, 0,
            P(self.out), P(self.flags_dev), P(self.ws), self.ws_bytes, st))  # fmt: skip

    def step_host(self):
        import warnings

        hb, wl, s = self.hb, self.wl, self.wl.get("sources", {})
        if self.sharded:
            return self.job.run()  # upload + kernels + NCCL collective + download on rank 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if wl["kind"] == "prism_layer":
                return hb.prism_layer_gravity(wl["coords"], s["easting"], s["northing"], s["bottom"],
                                              s["top"], s["density"], "g_z")
            if wl["kind"] == "prism_gravity":
                return hb.prism_gravity(wl["coords"], s["prisms"], s["density"], wl["fields"],
                                        disable_checks=True)
            if wl["kind"] == "prism_magnetic":
                return hb.prism_magnetic(wl["coords"], s["prisms"], s["magnetization"], "b",
                                         disable_checks=True)
            if wl["kind"] == "tess":
                return hb.tesseroid_gravity(wl["coords"], wl["tesseroids"], wl["density"], "g_z",
                                            disable_checks=True)
            return hb.eqs_predict(wl["coords"], s["points"], s["coefs"])

    def flags(self):
        if self.job is not None:
            return int(self.job.d["flags"].item())
        return int(self.flags_dev.item())


def measure(stepper, steps, warmup, rank, world, local, sample_clocks, e2e_steps=None):
    """Device-timed steps (CUDA events, L2 flushed in between, max over ranks) + e2e steps."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", local)
    lib = stepper.lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    for _ in range(warmup):
        stepper.step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    barrier()
    launches_before = lib.hb200_launch_count()
    for k in range(steps):
        l2_flush.fill_(k)  # flush L2 between timed iterations (outside the events)
        starts[k].record()
        stepper.step_dev()
        ends[k].record()
    barrier()
    gpu_launches = int(lib.hb200_launch_count() - launches_before)  # counted by the library
    clocks = sampler.stop() if sampler else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    if stepper.flags() & 2:
        raise SystemExit("zero-distance pair in the synthetic workload")
    del l2_flush

    # end to end through the public API: host numpy buffers in, host result out
    e2e_steps = e2e_steps or steps
    for _ in range(min(warmup, 2)):
        stepper.step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        stepper.step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    pairs = stepper.pairs_per_step
    return {"value": pairs * steps / total_s, "ms_per_step": 1e3 * total_s / steps,
            "step_ms_min_max": [min(step_ms), max(step_ms)],
            "e2e_value": pairs * e2e_steps / e2e_s, "e2e_ms_per_step": 1e3 * e2e_s / e2e_steps,
            "gpu_launches": gpu_launches, "clocks": clocks}  # fmt: skip


def roofline_for(name, per_gpu_rate, fp64_peak, clocks, full_size):
    f_alg = 2.0 * I_PAIR[name]
    alg = per_gpu_rate * f_alg
    entry, why = executed_for(name)
    roof = {"bound": "fp64", "unit": "TFLOP/s", "peak": fp64_peak / 1e12,
            "peak_source": "measured on this device in this run: hb200_fp64_peak DFMA probe "
                           "(MEASURED_PEAKS.json carries no FP64 entry; nominal 148 SM x 64 lanes x 2 x "
                           "1.965 GHz = 37.2 TFLOP/s)"}  # fmt: skip
    if entry:
        f64, other = entry["fp64"], entry["other"]
        achieved = per_gpu_rate * 2.0 * f64
        roof.update({
            "achieved": achieved / 1e12, "frac": achieved / fp64_peak,
            "frac_of_nominal": achieved / NOMINAL_FP64_FLOPS,
            "definition": "EXECUTED FP64-pipe thread instructions per pair (ncu source page of this "
                          "kernel build) x 2 flop x pairs/s per GPU / peak: hardware utilisation of "
                          "the FP64 vector pipe",
            "executed": {"fp64_instr_per_pair": f64, "other_instr_per_pair": other,
                         "kernel": entry["kernel"], "source": entry["source"], "sha16": entry["sha16"]},
            "traffic": entry.get("dram_bytes_per_launch") if full_size else None,
        })  # fmt: skip
        if clocks and clocks.get("sm_mhz"):
            slots = 148 * 4 * clocks["sm_mhz"] * 1e6 * 32  # thread-level issue slots / s
            roof["executed"]["issue_bound_frac"] = per_gpu_rate * (2 * f64 + other) / slots
            roof["executed"]["note"] = ("an FP64 warp instruction occupies 2 issue slots: time ~ "
                                        "(2 FP64 + other) / issue rate (DESIGN.md section 4)")
    else:
        roof.update({"achieved": None, "frac": None, "traffic": None, "executed": None, "stale": why})
    roof["algorithmic"] = {
        "flops_per_pair": f_alg, "achieved": alg / 1e12, "speedup_over_peak": alg / fp64_peak,
        "note": "SURVEY 8d convention: 2 x FP64-pipe instructions of the REFERENCE algorithm per pair; "
                "> 1 means the merged-transcendental kernel needs fewer instructions than the "
                "reference formulation would at 100 % of peak. NOT a hardware fraction."}  # fmt: skip
    return roof


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
    lib = hb._lib.load()
    hb.init([local])
    if args.tile_mode >= 0:
        hb._lib.check(lib.hb200_set_tile_mode(args.tile_mode))
    sharded = world > 1 and (args.scaling == "strong" or args.shard == "sources")

    wl = make_workload(args.workload, args.n_obs, args.n_src, 0 if sharded else rank)
    if wl["kind"] == "tess" and sharded:
        raise SystemExit("tess_gz: use --scaling weak (replicas) with several ranks")
    stepper = Stepper(wl, rank, world, args.scaling, args.shard)

    # FP64 FMA peak of this device, measured now (MEASURED_PEAKS.json has no FP64 entry)
    flops, secs = ctypes.c_double(0), ctypes.c_double(0)
    hb._lib.check(lib.hb200_fp64_peak(args.peak_iters, ctypes.byref(flops), ctypes.byref(secs)))
    fp64_peak = flops.value

    main = measure(stepper, args.steps, args.warmup, rank, world, local, True)
    h2d, d2h = stepper.h2d, stepper.d2h
    del stepper
    torch.cuda.empty_cache()

    also = []
    if world == 1 and not args.no_also and args.workload == "layer_gz" and not args.n_obs:
        for name in ("tensor", "mag_b", "eqs"):
            try:  # a side line must never cost the main one
                w2 = make_workload(name)
                st2 = Stepper(w2, 0, 1, "strong", "observers")
                m2 = measure(st2, 3, 3, 0, 1, local, False, e2e_steps=2)
                roof2 = roofline_for(name, m2["value"], fp64_peak, main["clocks"], True)
                also.append({"name": name, "workload": w2["desc"], "value": m2["value"], "unit": "pair/s",
                             "steps": 3, "warmup": 3, "ms_per_step": m2["ms_per_step"],
                             "e2e": m2["e2e_value"], "gpu_launches": m2["gpu_launches"],
                             "roofline_frac": roof2.get("frac"),
                             "fp64_instr_per_pair": (roof2.get("executed") or {}).get("fp64_instr_per_pair"),
                             "algorithmic_speedup_over_peak": roof2["algorithmic"]["speedup_over_peak"]})
                del st2, w2
            except Exception as error:  # noqa: BLE001
                also.append({"name": name, "error": f"{type(error).__name__}: {error}"})
            torch.cuda.empty_cache()

    north_star = None
    if not args.no_north_star and args.workload == "layer_gz" and not args.n_obs:
        # BASELINE configs[2]: six tensor components, 1M prisms x 1M observers, observers sharded
        # over the ranks, result gathered on rank 0; end to end on host buffers, 1 step
        w3 = make_workload("tensor_1m")
        st3 = Stepper(w3, rank, world, "strong", "observers")
        m3 = measure(st3, 1, 1, rank, world, local, False, e2e_steps=1)
        north_star = {"workload": w3["desc"], "scaling": "strong", "n_gpus": world, "steps": 1,
                      "warmup": 1, "value": m3["value"], "unit": "pair/s",
                      "ms_per_step": m3["ms_per_step"], "e2e": m3["e2e_value"],
                      "e2e_ms_per_step": m3["e2e_ms_per_step"],
                      "sharding": "single GPU" if world == 1 else
                      f"observers over {world} ranks, gather to rank 0 (NCCL) inside the timed region"}
        del st3, w3
        torch.cuda.empty_cache()

    source_sharded = None
    if world > 1 and not args.no_also and args.workload == "layer_gz" and not args.n_obs:
        # BASELINE configs[4] (262 144 of its 4M observers): sources sharded over the ranks, every rank
        # a full-length partial field, ONE float64 reduce-sum to rank 0 over NCCL; short run
        w5 = make_workload("eqs")
        st5 = Stepper(w5, rank, world, "strong", "sources")
        m5 = measure(st5, 3, 3, rank, world, local, False, e2e_steps=2)
        source_sharded = {"workload": w5["desc"], "scaling": "strong", "n_gpus": world, "steps": 3,
                          "warmup": 3, "value": m5["value"], "unit": "pair/s",
                          "ms_per_step": m5["ms_per_step"], "e2e": m5["e2e_value"],
                          "gpu_launches": m5["gpu_launches"],
                          "sharding": f"sources over {world} ranks, torch.distributed NCCL reduce(sum, f64) "
                                      "to rank 0 inside the timed region"}
        del st5, w5
        torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nthreads = host_cores()
        for kind in cpu_kinds(wl):
            try:
                n_sample = size_cpu_sample(wl, args.cpu_seconds, nthreads, kind)
                rate, dt = cpu_rate(wl, n_sample, nthreads, kind)
            except Exception as error:  # noqa: BLE001 - e.g. no Numba cache directory: the C port follows
                print(f"cpu_baseline[{kind}] failed: {type(error).__name__}: {error}", file=sys.stderr)
                continue
            entry = {"value": rate, "unit": "pair/s", "cores": nthreads, "kind": kind,
                     "sample": f"first {n_sample} of {wl['coords'][0].size} observers x all "
                               f"{wl['n_src']} sources ({dt:.1f} s); {CPU_DESCRIPTION[kind]}"}
            if cpu_baseline is None:
                cpu_baseline = entry
            else:
                cpu_baseline["also"] = entry

    if rank == 0:
        value = main["value"]
        per_gpu = value / world
        full_size = not args.n_obs and not args.n_src
        line = {
            "metric": METRIC, "value": value, "unit": "pair/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if sharded else args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_for(wl, args.workload, args, world),
            "run": {"kernel_variant": int(lib.hb200_get_variant()),
                    "tile_mode": int(lib.hb200_get_tile_mode()),
                    "tesseroid_variant": int(lib.hb200_get_tesseroid_variant()),
                    "step_ms_min_max": main["step_ms_min_max"],
                    "collective": ("none" if not sharded else
                                   "torch.distributed NCCL gather of the observer slices" if
                                   args.shard != "sources" else "torch.distributed NCCL reduce(sum, f64)"),
                    "in_library_multi_gpu": "hb200_init(devices): one process, peer copies + fixed-order "
                                            "reduce kernel instead of NCCL (SURVEY 8e alternative)"},
            "e2e": {"value": main["e2e_value"], "unit": "pair/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": main["e2e_ms_per_step"],
                    "api": ("harmonica_b200 public API on numpy buffers (ctypes -> C ABI), blocking"
                            if not sharded else
                            "harmonica_b200.distributed.ShardedJob.run() on numpy buffers: H2D of the "
                            "rank's shard, kernels, NCCL collective, D2H on rank 0")},
            "gpu_launches": main["gpu_launches"],
            "clocks": main["clocks"],
            "roofline": roofline_for(args.workload, per_gpu, fp64_peak, main["clocks"], full_size),
            "cpu_baseline": cpu_baseline,
            "also": also,
            "north_star": north_star,
            "source_sharded": source_sharded,
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
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--shard", default="observers", choices=["observers", "sources"])
    ap.add_argument("--n-obs", type=int, default=0)
    ap.add_argument("--n-src", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--peak-iters", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--no-north-star", action="store_true")
    ap.add_argument("--tile-mode", type=int, default=-1, help="experiments: 0 per-CTA, 1 per-warp tiles")
    args = ap.parse_args()
    if args.impl == "reference":
        # a bounded sample per step: the whole --steps K --warmup W run stays within ~2 minutes
        args.cpu_seconds = max(0.5, min(args.cpu_seconds, 6.0, 120.0 / max(1, args.steps + args.warmup)))
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
