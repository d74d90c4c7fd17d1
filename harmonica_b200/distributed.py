"""
One-process-per-GPU sharding of the pairwise forward models over
``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU test-suite).

The path shards without any data-path exchange when observers are split: every
rank owns the observers ``[rank * N / W, (rank + 1) * N / W)``, the sources are
replicated, the output slices are disjoint and the only communication is the
gather of the result on ``dst`` (BASELINE configs 2-4). When the
sources are split instead (config 5: very many sources) every rank computes a
full-length partial field over its source slice and ONE collective combines
them: a float64 reduce-sum.

Everything between the upload of this rank's shard and the download of the
result on ``dst`` stays on the device: the kernels are launched through the
``*_dev`` entry points of the C ABI on torch's current stream, the collective
runs on the device tensors. PyTorch is the plumbing only (device memory,
streams, the process group). The in-process multi-GPU path of the C library
(``hb200_init`` with several devices) does the same from ONE process with peer
copies; this module is for ``torchrun`` jobs.

    torchrun --nproc-per-node 8 job.py
        dist.init_process_group("nccl"); torch.cuda.set_device(LOCAL_RANK)
        g_z = hbd.prism_layer_gravity(coords, easting, northing, bottom, top, density, "g_z")
        # numpy array on rank 0, None elsewhere (dst=None: on every rank)
"""

import ctypes

import numpy as np

from . import _lib
from ._lib import FIELD_IDS


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) of ``n`` units for ``rank`` of ``world``."""
    return n * rank // world, n * (rank + 1) // world


def _group_info(group):
    import torch.distributed as dist  # noqa: PLC0415

    if isinstance(group, str) and group == "local":  # this process alone, whatever is initialised
        return 0, 1, None
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1, None
    return dist.get_rank(group), dist.get_world_size(group), dist.get_backend(group)


def _device_for(backend):
    import torch  # noqa: PLC0415

    if backend == "gloo":
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------- collectives
def gather_observer_slices(local, n_obs, group=None, dst=0):
    """
    ``local``: this rank's ``(n_fields, n_local)`` float64 tensor (on the
    device for NCCL). Returns the ``(n_fields, n_obs)`` tensor on ``dst``
    (``None`` on the other ranks), or on every rank when ``dst is None``.
    Shards may be ragged: slots are padded to ``ceil(n_obs / world)``.
    """
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    rank, world, _ = _group_info(group)
    if world == 1:
        return local
    n_fields = local.shape[0]
    slot = (n_obs + world - 1) // world
    send = local
    if local.shape[1] != slot:
        send = torch.zeros((n_fields, slot), dtype=local.dtype, device=local.device)
        send[:, : local.shape[1]] = local
    send = send.contiguous()
    if dst is None:
        recv = torch.empty((world, n_fields, slot), dtype=local.dtype, device=local.device)
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    else:
        recv = None
        if rank == dst:
            recv = torch.empty((world, n_fields, slot), dtype=local.dtype, device=local.device)
        dist.gather(send, list(recv.unbind(0)) if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    if n_obs == slot * world:
        return recv.permute(1, 0, 2).reshape(n_fields, n_obs)
    out = torch.empty((n_fields, n_obs), dtype=local.dtype, device=local.device)
    for r in range(world):
        a, b = shard_bounds(n_obs, r, world)
        out[:, a:b] = recv[r, :, : b - a]
    return out


def reduce_source_partials(partial, group=None, dst=0):
    """Float64 sum of the per-rank partial fields: ``reduce`` to ``dst`` (in place on ``dst``;
    the other ranks return ``None``) or ``all_reduce`` when ``dst is None``."""
    import torch.distributed as dist  # noqa: PLC0415

    rank, world, _ = _group_info(group)
    if world == 1:
        return partial
    if dst is None:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
        return partial
    dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return partial if rank == dst else None


# ------------------------------------------------ callback-based generic helpers
def observer_sharded(compute, coordinates, n_fields=1, group=None, dst=None):
    """
    ``compute(sub_coordinates) -> array (n_fields, n_local)`` on this rank's
    observer slice; returns the full ``(n_fields, n_obs)`` numpy result (on
    every rank by default, on ``dst`` only when given).
    """
    import torch  # noqa: PLC0415

    rank, world, backend = _group_info(group)
    coords = tuple(np.ascontiguousarray(np.asarray(c, dtype=np.float64).ravel()) for c in coordinates[:3])
    n_obs = coords[0].size
    lo, hi = shard_bounds(n_obs, rank, world)
    local = np.asarray(compute(tuple(c[lo:hi] for c in coords)), dtype=np.float64).reshape(n_fields, hi - lo)
    t = torch.from_numpy(np.ascontiguousarray(local)).to(_device_for(backend))
    full = gather_observer_slices(t, n_obs, group, dst)
    return None if full is None else full.cpu().numpy()


def source_sharded(compute_partial, n_sources, n_obs, n_fields=1, group=None, dst=None):
    """
    ``compute_partial(lo, hi) -> array (n_fields, n_obs)``: the field of sources
    ``[lo, hi)`` on ALL observers (linear units). Returns the summed field.
    """
    import torch  # noqa: PLC0415

    rank, world, backend = _group_info(group)
    lo, hi = shard_bounds(n_sources, rank, world)
    part = np.asarray(compute_partial(lo, hi), dtype=np.float64).reshape(n_fields, n_obs)
    t = torch.from_numpy(np.ascontiguousarray(part)).to(_device_for(backend))
    full = reduce_source_partials(t, group, dst)
    return None if full is None else full.cpu().numpy()


# ------------------------------------------------------------ device-resident jobs
def _field_mask(fields):
    names = (fields,) if isinstance(fields, str) else tuple(fields)
    for f in names:
        if f not in FIELD_IDS:
            raise ValueError(f"Gravitational field {f} not recognized")
    ids = [FIELD_IDS[f] for f in names]
    if sorted(ids) != ids or len(set(ids)) != len(ids):
        raise ValueError("fields must be distinct and in the order of the FIELDS table")
    mask = 0
    for i in ids:
        mask |= 1 << i
    return mask, len(ids)


class ShardedJob:
    """
    One forward model sharded over the ranks of a process group, with all of its buffers
    resident on this rank's GPU.

        job = ShardedJob("prism_layer", coords, dict(easting=..., northing=..., bottom=...,
                         top=..., density=...), "g_z")
        job.upload()            # H2D of this rank's shard (observer slice + replicated sources)
        dev = job.launch()      # kernels + collective, asynchronous on torch's current stream
        out = job.result()      # D2H on dst: numpy (n_fields, n_obs); None elsewhere

    kinds and their ``sources`` dicts (float64 numpy arrays, identical on every rank):
      prism_layer     easting, northing, bottom, top, density [, thickness_threshold]
      prism_gravity   prisms (P, 6), density
      prism_magnetic  prisms (P, 6), magnetization = (M_e, M_n, M_u); fields "b" or "b_e" ...
      eqs_predict     points = (e, n, u), coefs       (sum coef / distance)
      point_gravity   points = (e, n, u), masses      (Cartesian)
    ``shard``: "observers" (gather) or "sources" (reduce-sum; eqs_predict / point_gravity /
    prism_gravity / prism_magnetic).
    """

    def __init__(self, kind, coordinates, sources, fields, shard="observers", group=None, dst=0):
        import torch  # noqa: PLC0415

        if shard not in ("observers", "sources"):
            raise ValueError(f"Invalid shard '{shard}'. Choose 'observers' or 'sources'.")
        if shard == "sources" and kind == "prism_layer":
            raise ValueError("a prism layer is replicated; shard its observers")
        self.kind, self.shard, self.group, self.dst = kind, shard, group, dst
        self.rank, self.world, backend = _group_info(group)
        if backend == "gloo":
            raise _lib.HarmonicaB200Error("ShardedJob launches CUDA kernels: it needs the nccl backend "
                                          "(harmonica_b200 has no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.coords = tuple(_lib.f64(np.asarray(c).ravel()) for c in coordinates[:3])
        self.n_obs = self.coords[0].size
        self.sources = dict(sources)
        if kind == "prism_magnetic":
            comps = {"b": 7, "b_e": 1, "b_n": 2, "b_u": 4}
            if fields not in comps:
                raise ValueError(f"Invalid field '{fields}'. Please choose one of 'b,b_e,b_n,b_u'.")
            self.mask, self.nf = comps[fields], 3 if fields == "b" else 1
        elif kind == "eqs_predict":
            self.mask, self.nf = 1, 1
        else:
            self.mask, self.nf = _field_mask(fields)
        if kind == "prism_layer":
            self.n_src = self.sources["easting"].size * self.sources["northing"].size
        elif kind in ("prism_gravity", "prism_magnetic"):
            self.n_src = np.asarray(self.sources["prisms"]).shape[0]
        else:
            self.n_src = np.asarray(self.sources["points"][0]).size
        if shard == "observers":
            self.olo, self.ohi = shard_bounds(self.n_obs, self.rank, self.world)
            self.slo, self.shi = 0, self.n_src
        else:
            self.olo, self.ohi = 0, self.n_obs
            self.slo, self.shi = shard_bounds(self.n_src, self.rank, self.world)
        self.n_local, self.n_src_local = self.ohi - self.olo, self.shi - self.slo
        self.h2d_bytes = 0
        self.d = None

    # -- host -> device
    def _put(self, a):
        import torch  # noqa: PLC0415

        a = _lib.f64(a)
        self.h2d_bytes += a.nbytes
        return torch.from_numpy(a).to(self.device, non_blocking=False)

    def upload(self):
        import torch  # noqa: PLC0415

        self.h2d_bytes = 0
        s, d = self.sources, {}
        d["obs"] = [self._put(c[self.olo:self.ohi]) for c in self.coords]
        sl = slice(self.slo, self.shi)
        if self.kind == "prism_layer":
            for k in ("easting", "northing", "bottom", "top", "density"):
                d[k] = self._put(s[k])
        elif self.kind == "prism_gravity":
            d["prisms"] = self._put(np.asarray(s["prisms"])[sl])
            d["density"] = self._put(np.asarray(s["density"])[sl])
        elif self.kind == "prism_magnetic":
            d["prisms"] = self._put(np.asarray(s["prisms"])[sl])
            d["mag"] = [self._put(np.asarray(m)[sl]) for m in s["magnetization"]]
        else:
            d["points"] = [self._put(np.asarray(p)[sl]) for p in s["points"]]
            d["weights"] = self._put(np.asarray(s["coefs" if self.kind == "eqs_predict" else "masses"])[sl])
        if self.d is None or self.d["out"].shape != (self.nf, self.n_local):
            lib = self.lib
            if self.kind in ("eqs_predict", "point_gravity"):
                ws_bytes = lib.hb200_point_ws_bytes(self.n_local, self.n_src_local)
            else:
                ws_bytes = lib.hb200_prism_ws_bytes(self.n_local, self.n_src_local, self.nf)
            d["ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            d["out"] = torch.empty((self.nf, self.n_local), dtype=torch.float64, device=self.device)
            d["flags"] = torch.zeros(1, dtype=torch.int32, device=self.device)
        else:
            for k in ("ws", "out", "flags"):
                d[k] = self.d[k]
        self.d = d
        return self

    # -- kernels + collective (asynchronous)
    def launch_local(self):
        """This rank's kernels only (no collective): fills the local output tensor."""
        import torch  # noqa: PLC0415

        d, lib = self.d, self.lib
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        o = d["obs"]
        ws, wsb = P(d["ws"]), d["ws"].numel()
        if self.n_local == 0:
            return d["out"]
        if self.kind == "prism_layer":
            rc = lib.hb200_prism_layer_gravity_dev(
                P(o[0]), P(o[1]), P(o[2]), self.n_local, P(d["easting"]), d["easting"].numel(),
                P(d["northing"]), d["northing"].numel(), P(d["bottom"]), P(d["top"]), P(d["density"]),
                float(self.sources.get("thickness_threshold") or 0.0), self.mask, P(d["out"]),
                P(d["flags"]), ws, wsb, st)  # fmt: skip
        elif self.kind == "prism_gravity":
            rc = lib.hb200_prism_gravity_dev(
                P(o[0]), P(o[1]), P(o[2]), self.n_local, P(d["prisms"]), P(d["density"]),
                self.n_src_local, self.mask, P(d["out"]), P(d["flags"]), ws, wsb, st)
        elif self.kind == "prism_magnetic":
            m = d["mag"]
            rc = lib.hb200_prism_magnetic_dev(
                P(o[0]), P(o[1]), P(o[2]), self.n_local, P(d["prisms"]), P(m[0]), P(m[1]), P(m[2]),
                self.n_src_local, self.mask, _lib.MAG_DEFAULT_RULES, P(d["out"]), P(d["flags"]), ws,
                wsb, st)
        else:
            p = d["points"]
            rc = lib.hb200_point_gravity_dev(
                P(o[0]), P(o[1]), P(o[2]), self.n_local, P(p[0]), P(p[1]), P(p[2]), P(d["weights"]),
                self.n_src_local, self.mask, 0, 0 if self.kind == "eqs_predict" else 1, P(d["out"]),
                P(d["flags"]), ws, wsb, st)
        _lib.check(rc)
        return d["out"]

    def launch(self):
        """Kernels + the gather (observer shards) or reduce-sum (source shards)."""
        local = self.launch_local()
        if self.shard == "observers":
            self.full = gather_observer_slices(local, self.n_obs, self.group, self.dst)
        else:
            # the unit / sign scaling of the *_dev entry points is linear: partial fields add up
            self.full = reduce_source_partials(local, self.group, self.dst)
        return self.full

    # -- device -> host
    def result(self):
        import torch  # noqa: PLC0415

        torch.cuda.current_stream().synchronize()
        self.flags = int(self.d["flags"].item())
        if self.full is None:
            return None
        return self.full.cpu().numpy()

    def run(self):
        """upload + launch + result: the end-to-end call on host buffers."""
        self.upload()
        self.launch()
        return self.result()

    @property
    def d2h_bytes(self):
        return 8 * self.nf * self.n_obs


def _finish(job, shape, squeeze):
    out = job.run()
    if job.flags & _lib.FLAG_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    if out is None:
        return None
    out = out.reshape((job.nf, *shape))
    return out[0] if squeeze else tuple(out)


def prism_layer_gravity(coordinates, easting, northing, bottom, top, density, field,
                        thickness_threshold=None, group=None, dst=0):
    """``harmonica_b200.prism_layer_gravity`` with the observers sharded over the ranks."""
    shape = np.broadcast(*coordinates[:3]).shape
    src = dict(easting=_lib.f64(easting), northing=_lib.f64(northing), bottom=_lib.f64(bottom),
               top=_lib.f64(top), density=_lib.f64(density), thickness_threshold=thickness_threshold)
    job = ShardedJob("prism_layer", coordinates, src, field, "observers", group, dst)
    return _finish(job, shape, isinstance(field, str))


def prism_gravity(coordinates, prisms, density, field, shard="observers", group=None, dst=0):
    """``harmonica_b200.prism_gravity`` (checks disabled) sharded over the ranks."""
    from ._prism_gravity import _discard_null_prisms  # noqa: PLC0415

    shape = np.broadcast(*coordinates[:3]).shape
    prisms = np.atleast_2d(np.asarray(prisms, dtype=np.float64))
    density = np.atleast_1d(np.asarray(density, dtype=np.float64)).ravel()
    prisms, density, _ = _discard_null_prisms(prisms, density)  # gravity.py:452-486
    job = ShardedJob("prism_gravity", coordinates, dict(prisms=prisms, density=density), field,
                     shard, group, dst)
    return _finish(job, shape, isinstance(field, str))


def prism_magnetic(coordinates, prisms, magnetization, field, shard="observers", group=None, dst=0):
    """``harmonica_b200.prism_magnetic`` (checks disabled) sharded over the ranks."""
    shape = np.broadcast(*coordinates[:3]).shape
    prisms = np.atleast_2d(np.asarray(prisms, dtype=np.float64))
    magnetization = tuple(np.atleast_1d(np.asarray(m, dtype=np.float64)).ravel() for m in magnetization)
    null = ((prisms[:, 0] == prisms[:, 1]) | (prisms[:, 2] == prisms[:, 3]) | (prisms[:, 4] == prisms[:, 5])
            | ((magnetization[0] == 0) & (magnetization[1] == 0) & (magnetization[2] == 0)))
    prisms = prisms[~null]  # magnetic.py:403-440
    magnetization = tuple(m[~null] for m in magnetization)
    job = ShardedJob("prism_magnetic", coordinates, dict(prisms=prisms, magnetization=magnetization),
                     field, shard, group, dst)
    return _finish(job, shape, field != "b")


def eqs_predict(coordinates, points, coefs, shard="sources", group=None, dst=0):
    """``harmonica_b200.eqs_predict``; by default the SOURCES are sharded and the partial
    fields reduce-summed over NCCL (BASELINE config 5)."""
    shape = np.broadcast(*coordinates[:3]).shape
    job = ShardedJob("eqs_predict", coordinates, dict(points=points, coefs=coefs), "potential",
                     shard, group, dst)
    return _finish(job, shape, True)


def point_gravity(coordinates, points, masses, field, shard="observers", group=None, dst=0):
    """``harmonica_b200.point_gravity`` (Cartesian) sharded over the ranks."""
    shape = np.broadcast(*coordinates[:3]).shape
    job = ShardedJob("point_gravity", coordinates, dict(points=points, masses=masses), field,
                     shard, group, dst)
    return _finish(job, shape, isinstance(field, str))
