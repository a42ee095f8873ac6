"""Host -> device feed under contention (run under torchrun, one rank per GPU): what bounds `e2e` at 4-8 GPUs?

Per rank: NUMA placement of the GPU and of the process, then the pinned H2D rate of a 1 GB buffer with k = 1, 2, 4, ..., N
ranks copying at once, for three placements of the pinned pages: default (first touch under the process' CPU affinity),
interleaved over all NUMA nodes (set_mempolicy MPOL_INTERLEAVE), bound to the GPU's own node (MPOL_BIND); plus the host
memory copy bandwidth with all ranks copying.  Rank 0 prints one JSON document.
"""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import dist as mdist  # noqa: E402

MPOL_DEFAULT, MPOL_BIND, MPOL_INTERLEAVE = 0, 2, 3
SYS_set_mempolicy = 238                      # x86_64


def set_mempolicy(mode, nodes):
    mask = 0
    for n in nodes:
        mask |= 1 << n
    arr = (ctypes.c_ulong * 16)(*([mask & (2 ** 64 - 1)] + [0] * 15))
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(arr) if nodes else None, 1024 if nodes else 0)
    return rc, ctypes.get_errno()


def online_nodes():
    try:
        txt = open("/sys/devices/system/node/online").read().strip()
    except OSError:
        return [0]
    out = []
    for part in txt.split(","):
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def main():
    rank, world, local = mdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    props = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    try:
        gpu_node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
    except (OSError, ValueError):
        gpu_node = None
    nodes = online_nodes()
    info = {"rank": rank, "pci": bus, "gpu_numa_node": gpu_node, "cpus_allowed": sorted(os.sched_getaffinity(0))[:4] +
            ["...", len(os.sched_getaffinity(0))]}
    n_elems = 1 << 27                                           # 1 GB of doubles
    dst = torch.empty(n_elems, dtype=torch.float64, device=dev)

    def h2d_rate(src, active):
        """GB/s of this rank with the ranks < active copying at once (others idle)."""
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        gbs = 0.0
        if rank < active:
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                if rep:
                    gbs = max(gbs, 2 * n_elems * 8 / (time.perf_counter() - t0) / 1e9)
        if world > 1:
            td.barrier()
        t = torch.tensor([gbs if rank < active else 1e9], dtype=torch.float64, device=dev)
        s = torch.tensor([gbs if rank < active else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MIN)
            td.all_reduce(s, op=td.ReduceOp.SUM)
        return float(t.item()), float(s.item())

    results = {}
    placements = [("default", MPOL_DEFAULT, [])]
    if len(nodes) > 1:
        placements.append(("interleave_all_nodes", MPOL_INTERLEAVE, nodes))
        if gpu_node is not None and gpu_node >= 0:
            placements.append(("bind_gpu_node", MPOL_BIND, [gpu_node]))
    for name, mode, node_set in placements:
        rc, err = set_mempolicy(mode, node_set)
        src = torch.empty(n_elems, dtype=torch.float64).pin_memory()
        src.fill_(1.0)
        set_mempolicy(MPOL_DEFAULT, [])
        res = {"set_mempolicy_rc": rc, "errno": err}
        k = 1
        while k <= world:
            slowest, total = h2d_rate(src, k)
            res["ranks_%d" % k] = {"slowest_gbs": slowest, "total_gbs": total}
            k *= 2
        results[name] = res
        del src
    # the storage feed itself (mlmc_b200.sample_storage.stream_levels: double-buffered chunks on a copy stream), all ranks
    # at once, for several chunk sizes
    from mlmc_b200.sample_storage import stream_levels
    src = torch.empty((50_000_000, 1, 1), dtype=torch.float64).pin_memory()      # 400 MB, like one cfg2 step
    src.fill_(1.0)
    chunked = {}
    for mb in (8, 32, 128, 400):
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            if world > 1:
                td.barrier()
            t0 = time.perf_counter()
            for _lvl, _rows in stream_levels([(0, src)], dev, mb << 20):
                pass
            torch.cuda.synchronize()
            if rep:
                best = min(best, time.perf_counter() - t0)
        t = torch.tensor([src.numel() * 8 / best / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MIN)
        chunked["chunk_%d_MB_slowest_gbs" % mb] = float(t.item())
    results["stream_levels_all_ranks"] = chunked
    del src
    # host memory copy bandwidth, all ranks at once (threads = allowed CPUs / world)
    threads = max(1, len(os.sched_getaffinity(0)) // world)
    torch.set_num_threads(threads)
    a = torch.empty(1 << 26, dtype=torch.float64)
    b = torch.empty_like(a)
    a.fill_(1.0)
    b.copy_(a)
    if world > 1:
        td.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        b.copy_(a)
    host_gbs = 3 * 2 * a.numel() * 8 / (time.perf_counter() - t0) / 1e9
    t = torch.tensor([host_gbs], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    infos = [None] * world
    if world > 1:
        td.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        mem_total = None
        try:
            for line in open("/proc/meminfo"):
                if line.startswith("MemTotal"):
                    mem_total = line.split()[1] + " kB"
        except OSError:
            pass
        print(json.dumps({"world": world, "numa_nodes_online": nodes, "mem_total": mem_total, "ranks": infos,
                          "h2d": results, "host_copy_total_gbs_read_plus_write": float(t.item()),
                          "host_copy_threads_per_rank": threads}, indent=1))
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
