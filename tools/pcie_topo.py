"""Host <-> device copy bandwidth of an N-GPU box with all GPUs copying at once (context for bench.py's e2e number at
N > 1): run under torchrun, one rank per GPU.  Each rank measures H2D, D2H and duplex from pinned memory (a) wherever
the scheduler put the process and (b) after binding the process to the CPUs of its GPU's NUMA node and re-allocating.
usage: python -m torch.distributed.run --nproc-per-node N tools/pcie_topo.py"""
import os
import time

import torch
import torch.distributed as dist


def gpu_numa(idx):
    bus = torch.cuda.get_device_properties(idx)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        busid = pynvml.nvmlDeviceGetPciInfo(h).busId
        busid = busid.decode() if isinstance(busid, bytes) else busid
        busid = busid.lower()[-12:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % busid).read())
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % busid).read().strip()
        return busid, node, cpus
    except Exception as e:  # noqa: BLE001
        return str(e), -1, ""


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def measure(n=192 << 20, reps=6):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1); h2.fill_(2)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def t(fn):
        fn(); torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        dist.barrier()
        return n / dt / 1e9

    a = t(lambda: d.copy_(h, non_blocking=True))
    b = t(lambda: h2.copy_(d2, non_blocking=True))

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    c = t(both)
    return a, b, c


def main():
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo")
    busid, node, cpus = gpu_numa(rank)
    aff0 = sorted(os.sched_getaffinity(0))
    r0 = measure()
    bound = False
    if cpus:
        try:
            os.sched_setaffinity(0, set(parse_cpulist(cpus)) & set(aff0) or set(aff0))
            bound = True
        except OSError:
            pass
    r1 = measure()
    res = [None] * dist.get_world_size()
    dist.all_gather_object(res, (rank, busid, node, cpus, len(aff0), bound, r0, r1))
    if rank == 0:
        tot0 = [sum(r[6][i] for r in res) for i in range(3)]
        tot1 = [sum(r[7][i] for r in res) for i in range(3)]
        for r in res:
            print("gpu %d bus %s numa %d cpus %s affinity0 %d bound %s | unbound H2D %.1f D2H %.1f duplex %.1f | bound H2D %.1f D2H %.1f duplex %.1f"
                  % (r[0], r[1], r[2], r[3], r[4], r[5], *r[6], *r[7]))
        print("aggregate GB/s unbound: H2D %.1f D2H %.1f duplex(each way) %.1f | bound: H2D %.1f D2H %.1f duplex %.1f" % (*tot0, *tot1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
