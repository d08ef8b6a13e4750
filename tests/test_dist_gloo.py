"""world_size-2 gloo test (CPU) of the N>1 plumbing: shard by target, compute, all-gather, restore
order.  The compute function is injected (the oracle) because CUDA kernels cannot run here; on the
GPU box the same code path runs with NCCL and the CUDA engine (tests/test_gpu_dist.py)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed=5, n_lc=7):
    rng = np.random.default_rng(seed)
    times, fluxes = [], []
    for _ in range(n_lc):
        n = int(rng.integers(50, 400))
        t = np.sort(rng.uniform(0, 30, n))
        times.append(t)
        fluxes.append(1 + 0.01 * np.sin(2 * np.pi * t / 2.5) + 1e-3 * rng.normal(size=n))
    return times, fluxes, np.linspace(0.05, 3, 64)


def _oracle_compute(times, fluxes, frequency, normalization, norm_scale):
    from oracle import ls as ols
    return np.stack([np.sqrt(ols.ls_slow_psd(t, y, frequency)) * np.sqrt(4.0 / len(t)) for t, y in zip(times, fluxes)])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lightkurve_b200.dist import ls_power_sharded
    times, fluxes, freq = _make()
    out = ls_power_sharded(times, fluxes, freq, "amplitude", compute=_oracle_compute)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _oracle_compute_cat(t_cat, y_cat, off, freq, normalization, norm_scale, out):
    """Stand-in for engine.ls_power_ragged_device on CPU tensors (test infrastructure)."""
    t, y, f = t_cat.numpy(), y_cat.numpy(), freq.numpy()
    times = [t[off[b]:off[b + 1]] for b in range(len(off) - 1)]
    fluxes = [y[off[b]:off[b + 1]] for b in range(len(off) - 1)]
    out.copy_(torch.from_numpy(_oracle_compute(times, fluxes, f, normalization, None).astype(np.float32)))


def _worker_pipelined(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lightkurve_b200.dist import ShardedLombScargle
    times, fluxes, freq = _make()
    job = ShardedLombScargle(times, fluxes, freq, "amplitude", chunks=3, device=torch.device("cpu"),
                             compute=_oracle_compute_cat)
    out = job.run()
    again = job.run().clone()                      # a second step through the same buffers
    assert torch.equal(out, again)
    q.put((rank, out.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_chunk_pipelined_sharded_ls_matches_single_process():
    """ShardedLombScargle (device-resident shards, one asynchronous all-gather per piece): 7 ragged light curves
    over 2 ranks in 3 pieces (pieces of 2 / 1 / 1 rows, the short shard padded) = the single-process result."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_pipelined, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    times, fluxes, freq = _make()
    ref = _oracle_compute(times, fluxes, freq, "amplitude", None).astype(np.float32)
    for r in range(world):
        np.testing.assert_allclose(results[r], ref, rtol=1e-6)


def test_sharded_ls_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    times, fluxes, freq = _make()
    ref = _oracle_compute(times, fluxes, freq, "amplitude", None).astype(np.float32)
    for r in range(world):
        np.testing.assert_allclose(results[r], ref, rtol=1e-6)
