#!/usr/bin/env python
"""What can this box do with N concurrent device->host streams into pinned memory?

    python tools/d2h_probe.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py

Every rank copies a device buffer into a page-locked host buffer back to back for a few seconds (cudaMemcpyAsync,
one copy of `--mb` MiB at a time, on its own stream), all ranks at once; rank 0 prints ONE JSON line with the
per-rank and aggregate GB/s -- the ceiling of bench.py's e2e leg (8 B per cell-timestep have to cross PCIe).
Variants: default pinned allocation; pinned allocation after binding the process to the GPU's local CPUs (NUMA
first touch); H2D for comparison.  Topology (PCI bus id, NUMA node and local CPU list from sysfs) is reported."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def sysfs(bus_id, name):
    try:
        return open("/sys/bus/pci/devices/%s/%s" % (bus_id.lower(), name)).read().strip()
    except Exception:
        return None


def parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part.strip():
            cpus.add(int(part))
    return cpus


def measure(dev_buf, host_buf, seconds, d2h=True):
    stream = torch.cuda.Stream()
    n = 0
    with torch.cuda.stream(stream):
        for _ in range(2):
            (host_buf if d2h else dev_buf).copy_(dev_buf if d2h else host_buf, non_blocking=True)
        stream.synchronize()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                (host_buf if d2h else dev_buf).copy_(dev_buf if d2h else host_buf, non_blocking=True)
            stream.synchronize()
            n += 4
        dt = time.perf_counter() - t0
    return n * dev_buf.numel() * dev_buf.element_size() / dt / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=3.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    props = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    numa, cpulist = sysfs(bus, "numa_node"), sysfs(bus, "local_cpulist")
    n = args.mb << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    res = {}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    sync()
    res["d2h_default"] = measure(dev, host, args.seconds)
    sync()
    res["h2d_default"] = measure(dev, host, args.seconds, d2h=False)
    del host
    bound = False
    cpus = parse_cpulist(cpulist)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
            bound = True
        except Exception:
            bound = False
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)            # first touch on the bound CPUs
    sync()
    res["d2h_numa_bound"] = measure(dev, host, args.seconds)
    sync()
    info = {"rank": rank, "bus": bus, "numa_node": numa, "local_cpulist": cpulist, "bound": bound, **res}
    if world > 1:
        allinfo = [None] * world
        dist.all_gather_object(allinfo, info)
    else:
        allinfo = [info]
    if rank == 0:
        out = {"probe": "d2h", "n_gpus": world, "copy_mib": args.mb, "host_cpus": os.cpu_count(), "ranks": allinfo}
        for k in ("d2h_default", "h2d_default", "d2h_numa_bound"):
            out["aggregate_" + k + "_gbs"] = sum(r[k] for r in allinfo)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
