"""Sample-axis sharding across GPUs.

The reference has no distributed estimation path at all (its only parallelism is task-parallel sample
PRODUCTION, ``mlmc/sampling_pool.py``).  Samples of a level are i.i.d. rows and every statistic on the path is a
sum over rows, so each rank reduces a contiguous row range of every level and ONE small all-reduce (SUM, fp64)
of the packed level accumulators ``[L, 2 + 2K]`` combines them: counts ride in the same buffer (exact below
2^53), sums of a plain SUM all-reduce are exact up to fp64 rounding of P <= 8 partials.

One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch on the box, gloo in the CPU tests) is the
plumbing.  Sharding is OFF unless ``enable()`` was called (``bench.py`` does under torchrun), so a plain
single-process script never needs a process group.
"""
import os

import torch

_state = {"enabled": False, "rank": 0, "world": 1, "group": None, "peer": None, "peer_fallbacks": 0}


def enable(rank=None, world=None, group=None):
    """Turn row sharding on.  rank / world default to the initialised default process group."""
    import torch.distributed as td
    if rank is None or world is None:
        if not td.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        rank, world = td.get_rank(group), td.get_world_size(group)
    _state.update(enabled=world > 1, rank=int(rank), world=int(world), group=group)


def disable():
    _state.update(enabled=False, rank=0, world=1, group=None)


def enable_peer_reduce(slot_doubles=16384):
    """Set up the NVLink peer-memory exchange buffers of ``mlmcb200_allreduce_finalize_levels`` (one per rank, mapped
    into every other rank through cudaIpc handles; all ranks of the default group must sit on ONE node).  Collective.
    Afterwards ``estimate_mean`` adds the level sums and finalizes in a single launch whenever the accumulators fit a
    slot; wider ones keep the NCCL all-reduce.  Returns True on success, False (and stays on NCCL) otherwise."""
    import ctypes
    import torch.distributed as td
    from . import _native
    if _state["peer"] is not None:
        return True
    if world_size() < 2 or world_size() > 16 or not torch.cuda.is_available():
        return False
    lib = _native.load()
    world, rank_ = _state["world"], _state["rank"]
    ok = True
    own = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    n_bytes = lib.mlmcb200_peer_buffer_bytes(world, slot_doubles)
    if n_bytes < 0 or lib.mlmcb200_peer_alloc(n_bytes, ctypes.byref(own), handle) != 0:
        ok = False
    handles = [None] * world
    td.all_gather_object(handles, bytes(handle) if ok else None, group=_state["group"])
    ptrs, opened = [], []
    if ok and all(h is not None for h in handles):
        for r in range(world):
            if r == rank_:
                ptrs.append(own.value)
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            if lib.mlmcb200_peer_open(buf, ctypes.byref(p)) != 0:
                ok = False
                break
            ptrs.append(p.value)
            opened.append(p.value)
    else:
        ok = False
    flags = [None] * world
    td.all_gather_object(flags, ok, group=_state["group"])                  # all or nobody
    if not all(flags):
        for p in opened:
            lib.mlmcb200_peer_close(ctypes.c_void_p(p))
        if own.value:
            lib.mlmcb200_peer_free(own)
        return False
    device = torch.device("cuda", torch.cuda.current_device())
    _state["peer"] = {"rank": rank_, "world": world, "slot": int(slot_doubles), "own": own.value, "opened": opened,
                      "ptrs": torch.tensor(ptrs, dtype=torch.int64, device=device)}
    return True


def disable_peer_reduce():
    """Back to the NCCL all-reduce (the exchange buffers stay mapped until the process exits)."""
    _state["peer"] = None


def peer_state(n_doubles=None):
    """The peer-reduce set-up if it is active (and ``n_doubles`` fit a slot), else None."""
    peer = _state["peer"] if _state["enabled"] else None
    if peer is not None and n_doubles is not None and n_doubles > peer["slot"]:
        return None
    return peer


def note_peer_fallback():
    """Count an estimate whose fused peer reduce timed out and was redone with NCCL (``peer_fallbacks()``)."""
    _state["peer_fallbacks"] += 1


def peer_fallbacks():
    return _state["peer_fallbacks"]


def peer_error():
    """True if a fused reduce gave up waiting for a peer (its results are NaN)."""
    import ctypes
    from . import _native
    peer = _state["peer"]
    if peer is None:
        return False
    err = ctypes.c_int32(0)
    _native.load().mlmcb200_peer_error(ctypes.c_void_p(peer["own"]), peer["world"], peer["slot"], ctypes.byref(err))
    return err.value != 0


def world_size():
    return _state["world"] if _state["enabled"] else 1


def rank():
    return _state["rank"] if _state["enabled"] else 0


def shard_range(n, rank_=None, world=None):
    """Contiguous row range [lo, hi) of ``n`` rows owned by a rank: rank r gets [r n / P, (r+1) n / P)."""
    r = rank() if rank_ is None else rank_
    p = world_size() if world is None else world
    return (r * n) // p, ((r + 1) * n) // p


def all_reduce_sum(tensor):
    """In-place SUM all-reduce of a packed accumulator over the active group (no-op when not sharded)."""
    if world_size() == 1:
        return tensor
    import torch.distributed as td
    td.all_reduce(tensor, op=td.ReduceOp.SUM, group=_state["group"])
    return tensor


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (pinned staging buffers are then allocated
    next to the PCIe root of the GPU they feed).  Best effort: silently does nothing where sysfs does not tell."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT) and enable sharding.  Returns (rank, world, local_rank)."""
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank_ = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            bind_to_gpu_numa_node(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        td.init_process_group(backend=backend, rank=rank_, world_size=world, **kw)
    if world > 1:
        enable(rank_, world)
    return rank_, world, local
