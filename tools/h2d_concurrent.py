"""Concurrent host-to-device copy ceiling of one box: every rank copies pinned host buffers to its own GPU at the same time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/h2d_concurrent.py

Prints one JSON line (rank 0): per-rank GB/s alone (ranks take turns) and all together, for 256 MB buffers (the size of one
config-2 step's input) -- the ceiling the `e2e` figure of bench.py can reach at N GPUs.  Also tries NUMA-local pinned
memory when libnuma's policy can be set through `numactl`-less means (os.sched_setaffinity to the GPU's local cores
before allocating), and reports both."""
import json
import os
import time

import torch
import torch.distributed as dist


def gbs(nbytes, ms):
    return nbytes / ms / 1e6


def local_cpus(dev):
    """cores local to the GPU's PCIe root (sysfs), or None"""
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, devid)
        txt = open(path).read().strip()
        cpus = []
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus += list(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        return cpus or None
    except Exception:
        return None


def measure(host, devbuf, stream, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        devbuf.copy_(host, non_blocking=True)
        stream.synchronize()
        e0.record(stream)
        for _ in range(reps):
            devbuf.copy_(host, non_blocking=True)
        e1.record(stream)
    stream.synchronize()
    return gbs(host.numel() * host.element_size() * reps, e0.elapsed_time(e1))


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 64 * 1024 * 1024                                   # 256 MB of fp32
    stream = torch.cuda.Stream()
    devbuf = torch.empty(n, dtype=torch.float32, device="cuda")
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.fill_(1.0)
    cpus = local_cpus(local)
    host_local = None
    if cpus:
        try:
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, set(cpus) & old or old)
            host_local = torch.empty(n, dtype=torch.float32).pin_memory()      # first touch on a GPU-local core
            host_local.fill_(1.0)
            os.sched_setaffinity(0, old)
        except Exception:
            host_local = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    alone = [0.0] * world
    for r in range(world):                                 # one rank at a time
        barrier()
        if r == rank:
            alone[r] = measure(host, devbuf, stream)
        barrier()
    barrier()
    together = measure(host, devbuf, stream)               # all ranks at once
    barrier()
    together_local = measure(host_local, devbuf, stream) if host_local is not None else None
    barrier()
    res = torch.tensor([alone[rank], together, together_local if together_local is not None else -1.0], device="cuda")
    if world > 1:
        out = [torch.zeros_like(res) for _ in range(world)]
        dist.all_gather(out, res)
    else:
        out = [res]
    if rank == 0:
        rows = [o.tolist() for o in out]
        print(json.dumps({
            "what": "pinned host -> device copies of 256 MB, GB/s per rank",
            "n_gpus": world,
            "alone": [round(r[0], 1) for r in rows],
            "all_ranks_at_once": [round(r[1], 1) for r in rows],
            "all_ranks_at_once_numa_local_first_touch": [round(r[2], 1) for r in rows] if rows[0][2] >= 0 else None,
            "sum_all_at_once": round(sum(r[1] for r in rows), 1),
            "host_cores": os.cpu_count(),
            "local_cpulist_rank0": (cpus[:4] + ["..."] + cpus[-2:]) if cpus and len(cpus) > 6 else cpus,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
