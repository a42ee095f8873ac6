"""CPU, world_size 2 (gloo): the multi-rank host logic of the sharded estimate -- contiguous row shards per level,
packed level accumulators [L, 2 + 2K], one SUM all-reduce -- reproduces the un-sharded level sums.
Per-shard sums come from the oracle here (no GPU in this test); on the box the same ``mlmc_b200.dist`` calls carry
the CUDA accumulators over NCCL (tests/test_multi_gpu.py, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _level_sums(levels, basis, ranges):
    """Packed accumulator [L, 2 + 2K] of the rows in ``ranges`` (oracle arithmetic)."""
    from oracle import mlmc_oracle as orc
    K = basis.out_size
    acc = np.zeros((len(levels), 2 + 2 * K))
    for l, rows in enumerate(levels):
        lo, hi = ranges[l]
        part = rows[lo:hi]
        if len(part) == 0:
            continue
        x = part.transpose(2, 0, 1)
        if l == 0:
            x = x[:, :, :1]
        y, n_bad = orc.drop_nan_samples(orc.moments_chunk(basis, x))
        d = y[:, :, 0] if l == 0 else y[:, :, 0] - y[:, :, 1]
        acc[l, 0], acc[l, 1] = y.shape[1], n_bad
        acc[l, 2:2 + K] = d.sum(axis=1)
        acc[l, 2 + K:] = (d ** 2).sum(axis=1)
    return acc


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from mlmc_b200 import dist
    from oracle import mlmc_oracle as orc
    r, w, _local = dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dist.world_size() == world and dist.rank() == rank
    g = np.load(os.path.join(ROOT, "tests", "golden", "estimates.npz"))
    levels = [g["A_rows%d" % l] for l in range(3)]
    basis = orc.Basis("legendre", 12, tuple(g["A_domain"]))
    ranges = [dist.shard_range(len(rows)) for rows in levels]
    acc = torch.from_numpy(_level_sums(levels, basis, ranges))
    dist.all_reduce_sum(acc)
    np.save(os.path.join(out_dir, "acc%d.npy" % rank), acc.numpy())
    import torch.distributed as td
    td.barrier()
    td.destroy_process_group()


def test_sharded_level_sums_equal_unsharded(tmp_path, golden):
    from oracle import mlmc_oracle as orc
    world = 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    basis = orc.Basis("legendre", 12, tuple(g["A_domain"]))
    whole = _level_sums(levels, basis, [(0, len(r)) for r in levels])
    accs = [np.load(os.path.join(str(tmp_path), "acc%d.npy" % r)) for r in range(world)]
    assert np.array_equal(accs[0], accs[1])                                   # every rank holds the reduced result
    assert np.array_equal(accs[0][:, :2], whole[:, :2])                       # counts are exact
    assert np.allclose(accs[0], whole, rtol=1e-13, atol=1e-12)
    # finalised statistics from the reduced sums = the reference's
    n = accs[0][:, 0:1]
    K = 12
    l_means = accs[0][:, 2:2 + K] / n
    l_vars = (accs[0][:, 2 + K:] - accs[0][:, 2:2 + K] ** 2 / n) / (n - 1)
    assert np.allclose(l_means, g["A_leg_l_means"], rtol=1e-10, atol=1e-15)
    assert np.allclose(l_vars, g["A_leg_l_vars"], rtol=1e-10, atol=1e-15)
