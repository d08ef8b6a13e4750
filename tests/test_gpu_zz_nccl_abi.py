"""2-rank test of the C ABI's own NCCL exchange step (lkb_nccl_unique_id / lkb_nccl_init /
lkb_allgather_f32; skipped on a 1-GPU box).  No torch.distributed process group is created: the 128-byte
id travels from rank 0 to rank 1 through a multiprocessing queue, exactly the side channel a non-torch
host program would provide.  Named *_zz_* so that it runs after every single-GPU parity test."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _get(q, procs, timeout=240.0):
    """q.get that gives up as soon as a worker has died (instead of sitting out the whole timeout)."""
    import queue
    import time
    t0 = time.time()
    while True:
        try:
            return q.get(timeout=1.0)
        except queue.Empty:
            dead = [p for p in procs if not p.is_alive() and p.exitcode not in (0, None)]
            assert not dead, "worker exited with code %r" % [p.exitcode for p in dead]
            assert time.time() - t0 < timeout, "timed out waiting for the workers"


def _make(seed=9, n_lc=11):
    rng = np.random.default_rng(seed)
    times, fluxes = [], []
    for _ in range(n_lc):
        n = int(rng.integers(200, 3000))
        t = np.sort(rng.uniform(0, 27, n))
        times.append(t)
        fluxes.append(1 + 0.01 * np.sin(2 * np.pi * t / 2.5) + 1e-3 * rng.normal(size=n))
    return times, fluxes, np.linspace(0.05, 20, 500)


def _worker(rank, world, id_queues, out_q):
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    from lightkurve_b200 import engine
    from lightkurve_b200.dist import init_abi_communicator, ls_power_sharded
    engine.init(rank)

    def exchange(uid):
        if rank == 0:
            for q in id_queues[1:]:
                q.put(uid)
            return uid
        return id_queues[rank].get(timeout=120)

    init_abi_communicator(rank, world, exchange)
    assert engine.nccl_rank_world() == (rank, world)
    # plain gather: rank r contributes rows filled with r+1
    loc = torch.full((3, 5), float(rank + 1), device="cuda", dtype=torch.float32)
    got = engine.allgather_f32(loc)
    torch.cuda.synchronize()
    plain_ok = bool(torch.equal(got.cpu(), torch.arange(1, world + 1, dtype=torch.float32).repeat_interleave(3)[:, None]
                                .expand(-1, 5)))
    times, fluxes, freq = _make()
    out = ls_power_sharded(times, fluxes, freq, "amplitude", via="abi")
    torch.cuda.synchronize()
    out_q.put((rank, plain_ok, out.cpu().numpy()))
    engine.nccl_shutdown()
    assert engine.nccl_rank_world() == (-1, 0)


def test_abi_nccl_allgather_and_sharded_ls():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import ls as ols
    world = 2
    ctx = mp.get_context("spawn")
    id_queues = [ctx.Queue() for _ in range(world)]
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, id_queues, out_q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, plain_ok, arr = _get(out_q, procs)
            assert plain_ok
            results[rank] = arr
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:                                  # never leave a rank behind (a hung collective would block the exit)
        for p in procs:
            if p.is_alive():
                p.kill()
    times, fluxes, freq = _make()
    assert np.array_equal(results[0], results[1])                     # every rank holds the same [B, F]
    for b in (0, 5, 10):
        ref = np.sqrt(ols.ls_slow_psd(times[b], fluxes[b], freq)) * np.sqrt(4.0 / len(times[b]))
        got = results[0][b].astype(np.float64)
        assert np.all(np.abs(got - ref) <= 1e-5 * ref.max() + 1e-4 * ref)


def _single_rank_worker(out_q):
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(0)
    from lightkurve_b200 import engine
    from lightkurve_b200.dist import init_abi_communicator, ls_power_sharded
    engine.init(0)
    init_abi_communicator(0, 1, lambda uid: uid)
    ok = engine.nccl_rank_world() == (0, 1) and engine.nccl_version() > 0
    loc = torch.arange(12, device="cuda", dtype=torch.float32).reshape(3, 4)
    got = engine.allgather_f32(loc)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(got, loc))
    times, fluxes, freq = _make(n_lc=4)
    out = ls_power_sharded(times, fluxes, freq, "amplitude", via="abi")
    ref = engine.ls_power_ragged(times, fluxes, freq, "amplitude")
    ok = ok and np.array_equal(out.cpu().numpy(), np.asarray(ref))
    engine.nccl_shutdown()
    ok = ok and engine.nccl_rank_world() == (-1, 0)
    out_q.put(bool(ok))


def test_abi_nccl_single_rank_communicator():
    """World size 1 on one GPU: exercises the run-time NCCL binding, the id / init / all-gather / shutdown calls
    and the sharded path on top of them (a child process, so that the communicator never outlives the test)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    p = ctx.Process(target=_single_rank_worker, args=(out_q,), daemon=True)
    p.start()
    try:
        assert _get(out_q, [p]) is True
        p.join(timeout=60)
        assert p.exitcode == 0
    finally:
        if p.is_alive():
            p.kill()
