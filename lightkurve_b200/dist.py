"""Multi-GPU plumbing: shard a collection BY TARGET across ranks, one NCCL all-gather of the power
arrays.  One process per GPU (``torchrun``).  The all-gather runs either through
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo for the CPU tests of the host logic - the
default) or through the C ABI's own communicator (``lkb_nccl_init`` / ``lkb_allgather_f32``,
``via="abi"``), which needs no torch process group: ``init_abi_communicator`` shows the bootstrap.

The path has no other exchange step: every light curve's periodogram is independent, shared
inputs (frequency / period grids, a shared design matrix) are replicated.
"""
import numpy as np

__all__ = ["shard_by_length", "allgather_rows", "ls_power_sharded", "init_abi_communicator", "ShardedLombScargle"]


def shard_by_length(lengths, world_size):
    """Sort targets by cadence count (descending) and deal them round-robin: balances ragged
    batches (BASELINE config 5).  Returns a list of index arrays, one per rank."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    return [order[r::world_size] for r in range(world_size)]


def init_abi_communicator(rank, world_size, exchange):
    """Bootstrap the C ABI's NCCL communicator.  `exchange(id_bytes_or_None) -> id_bytes` is the side channel
    of the host program: rank 0 passes the fresh 128-byte id in and every rank gets rank 0's id back
    (e.g. ``torch.distributed.broadcast_object_list`` on any backend, an MPI broadcast, a pipe)."""
    from . import engine
    uid = exchange(engine.nccl_unique_id() if rank == 0 else None)
    engine.nccl_init(rank, world_size, uid)


def allgather_rows(local_rows, shards, n_total, group=None, via="torch"):
    """All-gather per-rank row blocks [n_local, F] and restore the original target order.
    `local_rows` is a torch tensor (CUDA for NCCL, CPU for gloo); returns [n_total, F] on every rank.
    ``via="abi"``: the gather is ``lkb_allgather_f32`` on the C ABI's communicator (CUDA float32 only)."""
    import torch
    import torch.distributed as dist
    if via == "abi":
        from . import engine
        world = engine.nccl_rank_world()[1]
        if world != len(shards):
            raise ValueError("the C-ABI communicator has %d ranks but there are %d shards" % (world, len(shards)))
    else:
        world = dist.get_world_size(group)
    F = local_rows.shape[1]
    n_max = max(len(s) for s in shards)
    pad = torch.zeros((n_max, F), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    if via == "abi":
        gathered = engine.allgather_f32(pad)
    else:
        gathered = torch.empty((world * n_max, F), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(gathered, pad, group=group)
    out = torch.empty((n_total, F), dtype=local_rows.dtype, device=local_rows.device)
    for r, s in enumerate(shards):
        if len(s):
            idx = torch.as_tensor(np.asarray(s), device=local_rows.device, dtype=torch.long)
            out[idx] = gathered[r * n_max: r * n_max + len(s)]
    return out


def ls_power_sharded(times, fluxes, frequency, normalization="amplitude", norm_scale=None, compute=None,
                     device=None, group=None, via="torch"):
    """Lomb-Scargle of a ragged collection sharded by target over the ranks of `group`.

    Every rank passes the SAME full lists (cheap host metadata); each computes the power of its
    own shard with the ragged CUDA kernel (K1) and one all-gather reassembles [B, F] on every
    rank.  `compute(times, fluxes, frequency, normalization, norm_scale) -> [n, F] array` can be
    injected (the gloo CPU tests use the oracle there; the product default is the CUDA engine).
    """
    import torch
    import torch.distributed as dist
    if via == "abi":
        from . import engine
        rank, world = engine.nccl_rank_world()
        if world <= 0:
            raise ValueError("via='abi' needs the C-ABI communicator (init_abi_communicator) on every rank")
    else:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    shards = shard_by_length([len(t) for t in times], world)
    mine = shards[rank]
    if compute is None:
        from . import engine
        compute = engine.ls_power_ragged
    ns = None if norm_scale is None else [norm_scale[i] for i in mine]
    F = len(frequency)
    if len(mine):
        local = np.asarray(compute([times[i] for i in mine], [fluxes[i] for i in mine], frequency,
                                   normalization, ns), dtype=np.float32).reshape(len(mine), F)
    else:
        local = np.zeros((0, F), dtype=np.float32)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if via == "abi" or dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
    return allgather_rows(torch.as_tensor(local, device=device), shards, len(times), group, via=via)


class ShardedLombScargle:
    """Device-resident, chunk-pipelined form of ``ls_power_sharded`` (BASELINE config 5: a ragged collection sharded
    by target over the GPUs of one box, the power rows reassembled on every rank by all-gathers over NVLink).

    ``ShardedLombScargle(times, fluxes, frequency, ...)`` deals the targets (sorted by length, round-robin), uploads
    THIS rank's shard once (`upload()`; call it again to time the host->device copy) and splits it into `chunks`
    pieces.  ``run()`` computes piece c on the current stream and hands its power block to an asynchronous
    ``all_gather_into_tensor`` while piece c + 1 is being computed - the collective of all but the last piece is
    hidden behind the kernels (gather per tile, overlapped).  Returns the [B, F] float32 CUDA
    tensor in the ORIGINAL target order on every rank.  Nothing goes through the host inside ``run()``.

    `compute(t_cat, y_cat, offsets, frequency, normalization, norm_scale, out)` can be injected (the gloo CPU test
    uses the oracle on CPU tensors); the product default is ``engine.ls_power_ragged_device``."""

    def __init__(self, times, fluxes, frequency, normalization="amplitude", norm_scale=None, chunks=4, device=None,
                 group=None, compute=None, algo="auto"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if device is None:
            nccl = self.world > 1 and dist.get_backend(group) == "nccl"
            device = torch.device("cuda", torch.cuda.current_device()) if (nccl or torch.cuda.is_available()) \
                else torch.device("cpu")
        self.device = device
        self.normalization, self.algo = normalization, algo
        self.n_total = len(times)
        self.F = len(frequency)
        self.shards = shard_by_length([len(t) for t in times], self.world)
        n_max = max(len(s) for s in self.shards)
        self.chunks = max(1, min(int(chunks), n_max))
        # piece c of every rank holds rows [lo_c, hi_c) of its (length-sorted) shard; pieces have the same row count
        # on every rank (the last rows of a short shard are padding: zero power rows that are dropped again)
        bounds = np.linspace(0, n_max, self.chunks + 1).round().astype(int)
        self.bounds = [(int(bounds[c]), int(bounds[c + 1])) for c in range(self.chunks)]
        mine = self.shards[self.rank]
        self._host = []
        for lo, hi in self.bounds:
            ids = mine[lo:min(hi, len(mine))]
            ts = [np.asarray(times[i], dtype=np.float64) for i in ids]
            ydt = np.float32 if all(np.asarray(fluxes[i]).dtype == np.float32 for i in ids) else np.float64
            ys = [np.asarray(fluxes[i], dtype=ydt) for i in ids]
            off = np.zeros(len(ids) + 1, dtype=np.int64)
            np.cumsum([len(t) for t in ts], out=off[1:])
            ns = None if norm_scale is None else np.asarray([norm_scale[i] for i in ids], dtype=np.float64)
            self._host.append(dict(t=np.concatenate(ts) if ts else np.zeros(0), y=np.concatenate(ys) if ys else
                                   np.zeros(0, ydt), off=off, ns=ns, n=len(ids)))
        self._freq_host = np.ascontiguousarray(frequency, dtype=np.float64)
        if compute is None:
            from . import engine
            compute = lambda t, y, off, f, norm, ns, out: engine.ls_power_ragged_device(t, y, off, f, norm, ns,
                                                                                        algo=self.algo, out=out)
        self._compute = compute
        self._pinned = None
        self._dev = None
        self._blocks = [torch.zeros((hi - lo, self.F), dtype=torch.float32, device=device) for lo, hi in self.bounds]
        self._gathered = [torch.empty((self.world * (hi - lo), self.F), dtype=torch.float32, device=device)
                          for lo, hi in self.bounds]
        self._out = torch.empty((self.n_total, self.F), dtype=torch.float32, device=device)
        # where the rows of every gathered piece go in the original order
        self._dst = []
        for lo, hi in self.bounds:
            idx = np.full(self.world * (hi - lo), -1, dtype=np.int64)
            for r, s in enumerate(self.shards):
                ids = s[lo:min(hi, len(s))]
                idx[r * (hi - lo): r * (hi - lo) + len(ids)] = ids
            keep = np.flatnonzero(idx >= 0)
            self._dst.append((torch.as_tensor(keep, device=device), torch.as_tensor(idx[keep], device=device)))

    @property
    def h2d_bytes(self):
        return int(sum(h["t"].nbytes + h["y"].nbytes for h in self._host) + self._freq_host.nbytes)

    def upload(self):
        """Host -> device copy of this rank's shard (pinned staging on CUDA)."""
        torch = self.torch
        cuda = self.device.type == "cuda"
        if self._pinned is None:
            pin = (lambda a: torch.from_numpy(a).pin_memory()) if cuda else torch.from_numpy
            self._pinned = [dict(t=pin(h["t"]), y=pin(h["y"]), ns=None if h["ns"] is None else pin(h["ns"]))
                            for h in self._host]
            self._pinned_f = pin(self._freq_host)
        self._dev = [dict(t=p["t"].to(self.device, non_blocking=True), y=p["y"].to(self.device, non_blocking=True),
                          ns=None if p["ns"] is None else p["ns"].to(self.device, non_blocking=True))
                     for p in self._pinned]
        self._dev_f = self._pinned_f.to(self.device, non_blocking=True)

    def run(self):
        if self._dev is None:
            self.upload()
        works = []
        for c, (h, d) in enumerate(zip(self._host, self._dev)):
            blk = self._blocks[c]
            if h["n"]:
                self._compute(d["t"], d["y"], h["off"], self._dev_f, self.normalization, d["ns"], blk[: h["n"]])
            if self.world > 1:
                works.append(self.dist.all_gather_into_tensor(self._gathered[c], blk, group=self.group, async_op=True))
            else:
                self._gathered[c] = blk
        for w in works:
            w.wait()
        for c, (src, dst) in enumerate(self._dst):
            self._out.index_copy_(0, dst, self._gathered[c].index_select(0, src))
        return self._out
