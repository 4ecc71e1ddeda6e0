#!/usr/bin/env python
"""Bare pinned-memory D2H / H2D bandwidth with R ranks copying AT THE SAME TIME (one process per GPU, torchrun), i.e. the
ceiling of the e2e (host-buffer) numbers of bench.py at R GPUs.  Copy size = the observation block of one e2e step
(2^20 envs x 157 B = 164.6 MB), in the 4 chunks mgb_step_host uses and as one copy.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port 29533 profiles/tools/pcie_bw_ranks.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = (1 << 20) * 157
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    streams = [torch.cuda.Stream(dev) for _ in range(3)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run(kind, chunks, reps=10):
        per = (n + chunks - 1) // chunks
        for it in range(2 + reps):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            for c in range(chunks):
                lo, hi = c * per, min(n, (c + 1) * per)
                with torch.cuda.stream(streams[c % 3]):
                    if kind == "d2h":
                        h[lo:hi].copy_(d[lo:hi], non_blocking=True)
                    else:
                        d[lo:hi].copy_(h[lo:hi], non_blocking=True)
            torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return reps * n / float(t.item()) / 1e9          # GB/s per rank at the pace of the slowest rank

    out = {"ranks": world, "bytes_per_copy": n}
    for kind in ("d2h", "h2d"):
        for chunks in (1, 4):
            out["%s_%dchunk_gbs_per_rank" % (kind, chunks)] = run(kind, chunks)
    out["d2h_aggregate_gbs"] = world * out["d2h_4chunk_gbs_per_rank"]
    try:
        out["numa_nodes_online"] = open("/sys/devices/system/node/online").read().strip()
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        out["gpu_numa_node"] = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        out["host_cpus"] = os.cpu_count()
    except Exception as e:
        out["topology_error"] = repr(e)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
