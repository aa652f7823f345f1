"""Multi-GPU host logic on CPU: frame-wise sharding, the one-frame window halo, result-slab gather and the
max-over-ranks timing rule, with torch.distributed (gloo, world_size 2 and 3).  The per-block "compute" here
is a deterministic stand-in that fills the same fixed-capacity slabs fe_pipeline_batch fills -- the GPU
kernels are covered by the -m gpu tests; this file checks that N shards concatenate to the 1-shard result."""
import os
import socket

import numpy as np
import pytest

from front_end_b200 import KPOINT, MATCH, shard


def test_shard_ranges_partition_everything():
    for n in (0, 1, 5, 96, 1024, 1023):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            for i in range(n):
                r = shard.shard_of(i, n, world)
                assert blocks[r][0] <= i < blocks[r][1]
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_window_shards_cover_consecutive_matches_with_halo():
    # BASELINE config 4: 10-frame window -> 9 consecutive-frame matches (src/WindowMatcher.cpp:104-157)
    for frames in (1, 2, 10, 11):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                m0, m1, f0, f1 = shard.window_shards(frames, r, world)
                if m1 > m0:
                    assert f0 == m0 - 1 and f1 == m1      # needs its block plus the frame before it
                seen += list(range(m0, m1))
            assert seen == list(range(1, frames))


def _fake_block(start, stop, cap=64):
    """Deterministic stand-in for FrontEnd.pipeline_batch on pairs [start, stop)."""
    n = stop - start
    out = dict(kps=np.zeros((2 * n, cap), KPOINT), desc=np.zeros((2 * n, cap, 32), np.uint8),
               n_kps=np.zeros(2 * n, np.int32), matches_a=np.zeros((n, cap), MATCH), n_a=np.zeros(n, np.int32),
               matches_b=np.zeros((n, cap), MATCH), n_b=np.zeros(n, np.int32))
    for p in range(start, stop):
        rng = np.random.default_rng(p)
        i = p - start
        for e in range(2):
            k = int(rng.integers(1, cap))
            out["n_kps"][2 * i + e] = k
            out["kps"]["x"][2 * i + e, :k] = rng.random(k, dtype=np.float32) * 1000
            out["kps"]["y"][2 * i + e, :k] = np.sort(rng.random(k, dtype=np.float32) * 700)
            out["desc"][2 * i + e, :k] = rng.integers(0, 256, (k, 32), dtype=np.uint8)
        for key, cnt in (("matches_a", "n_a"), ("matches_b", "n_b")):
            m = int(rng.integers(0, cap))
            out[cnt][i] = m
            out[key]["queryIdx"][i, :m] = np.arange(m)
            out[key]["trainIdx"][i, :m] = rng.integers(0, cap, m)
            out[key]["distance"][i, :m] = rng.integers(0, 256, m)
    return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop, out = shard.run_sharded(n_pairs, rank, world, _fake_block)
        if out is None:
            out = _fake_block(0, 0)
        res = shard.gather_results(out, start, stop, KPOINT, MATCH, dst=0)
        t = shard.max_over_ranks_ms(10.0 + rank)
        if rank == 0:
            ok = True
            want = _fake_block(0, n_pairs)
            covered = []
            for (s, e), o in sorted(res.items()):
                covered += list(range(s, e))
                for k in want:
                    lo, hi = (2 * s, 2 * e) if k in ("kps", "desc", "n_kps") else (s, e)
                    ok &= bool(np.array_equal(o[k], want[k][lo:hi]))
            ok &= covered == list(range(n_pairs))
            ok &= t == 10.0 + world - 1
            q.put(ok)
        else:
            assert res is None
            q.put(t == 10.0 + world - 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs", [(2, 7), (3, 2), (2, 1)])
def test_sharded_blocks_gather_to_the_single_rank_result(world, n_pairs):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def _phased_worker(rank, world, port, n_steps, q):
    """Every rank appends (phase, step, rank) records to a shared log; the phases of different ranks must never interleave."""
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hb = shard.HostBarrier("test_%d" % port, rank, world, create=(rank == 0)) if rank == 0 else None
        dist.barrier()                                   # rank 0 has created the file
        if hb is None:
            hb = shard.HostBarrier("test_%d" % port, rank, world, create=False)
        log = []

        def stamp(kind, i):
            t0 = time.perf_counter()
            time.sleep(0.002 * (1 + rank))               # ranks take different times inside a phase
            log.append((kind, i, t0, time.perf_counter()))
        shard.phased_steps(n_steps, lambda i: stamp("h2d", i), lambda i: None, lambda i: stamp("d2h", i), hb.wait)
        hb.wait()
        hb.close()
        q.put((rank, log))
    finally:
        dist.destroy_process_group()


def test_phased_copy_schedule_keeps_directions_apart():
    """shard.phased_steps + shard.HostBarrier (world size 2, CPU): every rank uploads steps 0 .. n-1 and downloads each exactly
    once and in order; no rank's host->device interval of step i overlaps any rank's device->host interval of step i - 1
    (the phase barrier), except for the drain of the last step which has no barrier after it."""
    import torch.multiprocessing as mp
    world, n_steps = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_phased_worker, args=(r, world, port, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    logs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert [(k, i) for k, i, _, _ in logs[r] if k == "h2d"] == [("h2d", i) for i in range(n_steps)]
        assert [(k, i) for k, i, _, _ in logs[r] if k == "d2h"] == [("d2h", i) for i in range(n_steps)]
    # CLOCK_MONOTONIC is shared by the processes of one host: intervals are comparable across ranks
    ups = {i: [(t0, t1) for r in logs for k, j, t0, t1 in logs[r] if k == "h2d" and j == i] for i in range(n_steps)}
    downs = {i: [(t0, t1) for r in logs for k, j, t0, t1 in logs[r] if k == "d2h" and j == i] for i in range(n_steps)}
    for i in range(1, n_steps):
        assert max(t1 for _, t1 in ups[i]) <= min(t0 for t0, _ in downs[i - 1])          # D2H(i-1) starts after every H2D(i)
        if i + 1 < n_steps:
            assert max(t1 for _, t1 in downs[i - 1]) <= min(t0 for t0, _ in ups[i + 1])  # and ends before any H2D(i+1)
    assert not os.path.exists("/dev/shm/fe_b200_barrier_test_%d" % port)
