"""Multi-GPU plumbing: shard a collection BY TARGET across ranks, one NCCL all-gather of the power
arrays.  One process per GPU (``torchrun``).  The all-gather runs either through
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo for the CPU tests of the host logic - the
default) or through the C ABI's own communicator (``lkb_nccl_init`` / ``lkb_allgather_f32``,
``via="abi"``), which needs no torch process group: ``init_abi_communicator`` shows the bootstrap.

The path has no other exchange step: every light curve's periodogram is independent, shared
inputs (frequency / period grids, a shared design matrix) are replicated.
"""
import numpy as np

__all__ = ["shard_by_length", "allgather_rows", "ls_power_sharded", "init_abi_communicator"]


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
