"""Frame-wise sharding of independent stereo pairs over the GPUs of one box.

Every stereo pair is independent (detect, describe and stereo matching touch only that pair), so the
batch is partitioned into contiguous blocks, one process (one fe_ctx) per GPU, and there is **no
collective on the data path**.  The only exchange step is optional: gathering the fixed-capacity result
slabs on one rank (`gather_results`), and -- for WindowMatcher sequences (src/WindowMatcher.cpp:104-157,
consecutive frames only) -- a one-frame halo so that the F-1 consecutive matches can be split between
ranks (`window_shards`).

The functions here are host logic only: they work with any torch.distributed backend (NCCL on the GPU
box; gloo in the CPU tests) and never touch descriptors or pixels themselves.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block of rank `rank` when n_items are split as evenly as possible: the first
    n_items % world ranks get one extra item.  Returns (start, stop)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_of(index, n_items, world):
    """Inverse of shard_range: the rank that owns item `index`."""
    base, extra = divmod(int(n_items), world)
    pivot = extra * (base + 1)
    if index < pivot:
        return index // (base + 1)
    return extra + (index - pivot) // max(base, 1)


def window_shards(n_frames, rank, world):
    """WindowMatcher matches frame f against frame f-1 for f = 1 .. n_frames-1.  Rank r gets a contiguous
    block of those matches and needs frames [first_match - 1, last_match] -- i.e. its block plus a one-frame
    halo at the front (the boundary frame's features are duplicated on two ranks; that is the only
    "exchange").  Returns (match_start, match_stop, frame_start, frame_stop); empty blocks have
    match_start == match_stop."""
    m0, m1 = shard_range(max(n_frames - 1, 0), rank, world)
    m0 += 1
    m1 += 1
    if m0 >= m1:
        return m0, m0, m0, m0
    return m0, m1, m0 - 1, m1


def run_sharded(n_pairs, rank, world, process_block):
    """Calls process_block(start, stop) for this rank's block of pairs and returns (start, stop, result)."""
    start, stop = shard_range(n_pairs, rank, world)
    return start, stop, (process_block(start, stop) if stop > start else None)


def pack_result_slab(out, start, stop):
    """Compacts a pipeline_batch output dict (fixed-capacity slabs) for pairs [start, stop) into one flat
    uint8 buffer + a small int64 header, the unit that travels in gather_results."""
    n = stop - start
    header = np.array([start, stop, out["kps"].shape[1], out["desc"].shape[2], out["desc"].dtype.itemsize], np.int64)
    parts = [np.ascontiguousarray(out[k][:2 * n if k in ("kps", "desc", "n_kps") else n]).view(np.uint8).reshape(-1)
             for k in ("n_kps", "n_a", "n_b", "kps", "desc", "matches_a", "matches_b")]
    return header, np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def unpack_result_slab(header, flat, kp_dtype, match_dtype):
    start, stop, cap, dw, isz = (int(v) for v in header)
    n = stop - start
    ddt = np.uint8 if isz == 1 else np.float32
    spec = [("n_kps", (2 * n,), np.int32), ("n_a", (n,), np.int32), ("n_b", (n,), np.int32),
            ("kps", (2 * n, cap), kp_dtype), ("desc", (2 * n, cap, dw), ddt),
            ("matches_a", (n, cap), match_dtype), ("matches_b", (n, cap), match_dtype)]
    out, off = {}, 0
    for name, shape, dt in spec:
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
        out[name] = flat[off:off + nbytes].view(dt).reshape(shape)
        off += nbytes
    return start, stop, out


def gather_results(out, start, stop, kp_dtype, match_dtype, dst=0, group=None, device=None):
    """Optional gather of per-rank result slabs on rank `dst` (NVLink when the backend is NCCL and `device`
    is a CUDA device; gloo/CPU otherwise).  Returns {(start, stop): out_dict} on dst, None elsewhere.  Not on
    the timed hot path: results are ~0.3-0.6 MB per pair and each rank can also just keep its own."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    header, flat = pack_result_slab(out, start, stop)
    dev = device if device is not None else "cpu"
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([flat.size], dtype=torch.int64, device=dev), group=group)
    max_size = max(int(s.item()) for s in sizes)
    hdrs = [torch.zeros(5, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(hdrs, torch.from_numpy(header).to(dev), group=group)
    padded = torch.zeros(max(max_size, 1), dtype=torch.uint8, device=dev)
    padded[:flat.size] = torch.from_numpy(flat).to(dev)
    bufs = [torch.zeros_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    res = {}
    for r in range(world):
        h = hdrs[r].cpu().numpy()
        if h[1] <= h[0]:
            continue
        s, e, o = unpack_result_slab(h, bufs[r].cpu().numpy()[:int(sizes[r].item())], kp_dtype, match_dtype)
        res[(s, e)] = o
    return res


def max_over_ranks_ms(local_ms, device=None, group=None):
    """bench.py's timing rule: the step time of a multi-GPU run is the max over ranks."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class HostBarrier:
    """Spin barrier between the ranks of ONE host, through per-rank epoch counters in a /dev/shm file.

    Used by the direction-phased copy schedule (`phased_steps`): with eight GPUs behind one host fabric, running all
    host->device copies together and then all device->host copies together moves more bytes per second than letting
    the two directions compete (measured on the 8-GPU boxes: 237 GB/s H2D alone, 120 GB/s D2H alone, 77 + 77 GB/s
    when mixed).  A phase change needs every rank on the host to agree on "now", tens of thousands of times per second
    less often than a collective would be worth: each rank owns one cache line it alone writes (its epoch), and waits
    until every line has reached that epoch.  No atomics, no GPU work, microseconds per wait.

    Rank 0 creates (and at close() removes) the file; the other ranks must attach after rank 0 has created it
    (callers put a torch.distributed barrier between the two)."""
    _SLOT = 8          # int64 words per rank = one 64-byte line

    def __init__(self, name, rank, world, create):
        import os
        self.path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", "fe_b200_barrier_%s" % name)
        self.rank, self.world, self.owner = int(rank), int(world), bool(create)
        if create:
            with open(self.path, "wb") as fh:
                fh.write(b"\0" * (8 * self._SLOT * self.world))
        self.slots = np.memmap(self.path, dtype=np.int64, mode="r+", shape=(self.world, self._SLOT))
        self.epoch = 0

    def wait(self, timeout_s=120.0):
        import time
        self.epoch += 1
        self.slots[self.rank, 0] = self.epoch
        col, ep = self.slots[:, 0], self.epoch
        spins, t0 = 0, None
        while int(col.min()) < ep:
            spins += 1
            if spins & 0x3FF == 0:
                t0 = t0 or time.perf_counter()
                if time.perf_counter() - t0 > timeout_s:
                    raise TimeoutError("HostBarrier: a rank did not arrive (epoch %d, slots %s)" % (ep, col.tolist()))

    def close(self):
        import os
        del self.slots
        if self.owner:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def phased_steps(n_steps, upload, run, download, barrier):
    """Direction-phased schedule of `n_steps` independent batches on one rank with two contexts (step i uses context
    i % 2):

        H2D phase i   : upload(i)  -- every rank copies host->device at the same time -- then run(i) starts (asynchronous)
        D2H phase i-1 : download(i - 1) -- every rank copies device->host at the same time; run(i) overlaps it

    `barrier()` is entered exactly 2 * n_steps times by every rank, so ranks with the same n_steps stay in lock step.
    upload / run / download take the step index; download(i) must wait for run(i) itself (fe_batch_download does)."""
    for i in range(n_steps):
        barrier()
        upload(i)
        run(i)
        barrier()
        if i > 0:
            download(i - 1)
    if n_steps > 0:
        download(n_steps - 1)
