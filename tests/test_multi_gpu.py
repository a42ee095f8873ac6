"""GPU, world size 2 (NCCL): sharded ``estimate_mean`` through the public API equals the single-GPU estimate.
Skipped on boxes with one GPU (the CPU/gloo test ``test_dist_gloo.py`` covers the host logic everywhere)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), MLMCB200_PEER_TIMEOUT_MS="300")
    import torch
    from mlmc_b200 import dist
    from mlmc_b200.moments import Legendre
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.quantity import quantity_estimate as qe
    dist.init_from_env(backend="nccl")
    g = np.load(os.path.join(ROOT, "tests", "golden", "estimates.npz"))
    levels = [g["A_rows%d" % l] for l in range(3)]
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    storage = Memory.from_arrays(levels, level_parameters=[[h] for h in g["A_steps"]], result_format=spec)
    value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
    qm = qe.estimate_mean(qe.moments(value, Legendre(12, tuple(g["A_domain"]))))
    cm = qe.estimate_mean(qe.covariance(value, Legendre(8, tuple(g["A_domain"]))))
    # bootstrap: replicates are sharded over the ranks, every rank ends up with all of them
    bs = qe.bootstrap_moments(value, Legendre(6, tuple(g["A_domain"])), [800, 400, 200], 5, seed=4, method="gather")
    # ... also on the one-pass weighted path (multiplicities x moment differences on DMMA tiles), 11 replicates
    bsw = qe.bootstrap_moments(value, Legendre(6, tuple(g["A_domain"])), [800, 400, 200], 11, seed=4, method="weighted")
    bs1 = bsw1 = None
    if rank == 0:                                   # the same calls on one rank (sharding off) for comparison
        dist.disable()
        bs1 = qe.bootstrap_moments(value, Legendre(6, tuple(g["A_domain"])), [800, 400, 200], 5, seed=4, method="gather")
        bsw1 = qe.bootstrap_moments(value, Legendre(6, tuple(g["A_domain"])), [800, 400, 200], 11, seed=4,
                                    method="weighted")
        dist.enable(rank, world)
    # the same estimates with the sum over the ranks fused into the finalize launch (NVLink peer memory)
    peer_ok = dist.enable_peer_reduce()
    peer = {}
    if peer_ok:
        for rep in range(3):                        # several epochs: the two exchange buffers alternate
            qm2 = qe.estimate_mean(qe.moments(value, Legendre(12, tuple(g["A_domain"]))))
            cm2 = qe.estimate_mean(qe.covariance(value, Legendre(8, tuple(g["A_domain"]))))
        peer = dict(p_l_means=qm2.l_means, p_l_vars=qm2.l_vars, p_n=qm2.n_samples, p_cov=cm2.mean, p_cov_var=cm2.var,
                    p_err=int(dist.peer_error()), p_fallbacks=dist.peer_fallbacks())
        # skewed ranks: rank 1 reaches the fused reduce 1.5 s after rank 0 gave up on it (the time-out is 300 ms here):
        # both must notice, repeat the reduction with NCCL and return the right numbers
        import time
        torch.cuda.synchronize()
        import torch.distributed as td_
        td_.barrier()
        if rank == 1:
            time.sleep(1.5)
        qm3 = qe.estimate_mean(qe.moments(value, Legendre(12, tuple(g["A_domain"]))))
        qm4 = qe.estimate_mean(qe.moments(value, Legendre(12, tuple(g["A_domain"]))))       # next epoch works again
        peer.update(s_l_means=qm3.l_means, s_l_vars=qm3.l_vars, s_n=qm3.n_samples, s_fallbacks=dist.peer_fallbacks(),
                    s2_l_means=qm4.l_means, s2_fallbacks=dist.peer_fallbacks())
    # vector quantity (6 components, 4 levels): fused covariance and transformed moments, rows sharded over the ranks
    levels_c = [g["C_rows%d" % l] for l in range(4)]
    spec_c = [QuantitySpec(name="v", unit="", shape=(6, 1), times=[0.0], locations=["0"])]
    storage_c = Memory.from_arrays(levels_c, result_format=spec_c)
    vec = make_root_quantity(storage_c, spec_c)["v"][0.0]["0"]
    vcov = qe.estimate_mean(qe.covariance(vec, Legendre(3, tuple(g["C_domain"]))))
    from mlmc_b200.moments import TransformedMoments
    l_mat = np.array([[1.0, 0.0, 0.0, 0.0, 0.0], [0.3, -1.2, 0.5, 0.0, 0.1], [0.0, 0.4, 0.0, -0.7, 2.0]])
    vtm = qe.estimate_mean(qe.moments(vec, TransformedMoments(Legendre(5, tuple(g["C_domain"])), l_mat)))
    peer.update(v_cov=vcov.mean, v_cov_var=vcov.var, v_tm_mean=vtm.mean, v_tm_var=vtm.var)
    # linearised covariance means (moment sums of the 2R-1 basis, all-reduce, C applied once)
    lin = qe.estimate_mean(qe.covariance(value, Legendre(8, tuple(g["A_domain"]))), variance=False)
    peer["lin_cov"] = lin.mean
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), l_means=qm.l_means, l_vars=qm.l_vars, n=qm.n_samples,
             n_rm=qm.n_rm_samples, cov=cm.mean, cov_var=cm.var, bs_l_means=bs["l_means"], bs_n=bs["n_samples"],
             bs1_l_means=bs1["l_means"] if bs1 is not None else np.zeros(0), bsw_l_vars=bsw["l_vars"], bsw_n=bsw["n_samples"],
             bsw1_l_vars=bsw1["l_vars"] if bsw1 is not None else np.zeros(0), peer_ok=int(peer_ok), **peer)
    import torch.distributed as td
    td.barrier()
    td.destroy_process_group()


def test_two_gpu_sharded_estimate(tmp_path, golden):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    g = golden("estimates")
    for r in range(world):
        out = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(out["n"], g["A_leg_n"]) and np.array_equal(out["n_rm"], g["A_leg_n_rm"])
        assert np.allclose(out["l_means"], g["A_leg_l_means"], rtol=1e-10, atol=1e-15)
        assert np.allclose(out["l_vars"], g["A_leg_l_vars"], rtol=1e-10, atol=1e-15)
        assert np.allclose(out["cov"], g["A_cov_mean"], rtol=1e-8, atol=1e-14)
        assert np.allclose(out["cov_var"], g["A_cov_var"], rtol=1e-8, atol=1e-16)
        assert np.allclose(out["lin_cov"], g["A_cov_mean"], rtol=1e-8, atol=1e-14)
        assert np.allclose(out["v_cov"], g["C_cov_mean"], rtol=1e-8, atol=1e-14)
        assert np.allclose(out["v_cov_var"], g["C_cov_var"], rtol=1e-8, atol=1e-16)
        # transformed moments of the vector quantity against the oracle on all rows
        from oracle import mlmc_oracle as orc
        l_mat = np.array([[1.0, 0.0, 0.0, 0.0, 0.0], [0.3, -1.2, 0.5, 0.0, 0.1], [0.0, 0.4, 0.0, -0.7, 2.0]])
        want = orc.estimate_moments([g["C_rows%d" % l] for l in range(4)],
                                    orc.Basis("legendre", 5, tuple(g["C_domain"]), matrix=l_mat))
        assert np.allclose(np.ravel(out["v_tm_mean"]), want.mean, rtol=1e-9, atol=1e-14)
        assert np.allclose(np.ravel(out["v_tm_var"]), want.var, rtol=1e-8, atol=1e-16)
    a, b = (np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(2))
    assert a["bs_l_means"].shape == (5, 3, 6) and np.array_equal(a["bs_l_means"], b["bs_l_means"])
    assert np.array_equal(a["bs_n"], b["bs_n"]) and np.all(a["bs_n"].sum(axis=1) > 0)
    # replicates are keyed by (seed, global replicate number): sharded over two ranks or not, the same five replicates
    assert np.array_equal(a["bs_l_means"], a["bs1_l_means"])
    # weighted path: the replicates of a seed do not depend on the world size; the CTA partials are summed in CTA order,
    # so even the bits agree
    assert a["bsw_l_vars"].shape[0] == 11 and np.array_equal(a["bsw_l_vars"], a["bsw1_l_vars"])
    assert np.all(a["bsw_n"].sum(axis=1) > 0)
    assert np.all(np.abs(a["bs_l_means"][:, :, 0].sum(axis=1) - 1.0) < 1e-12)
    # fused peer-memory reduce: same numbers as the NCCL route (two ranks: a + b either way), no time-out
    assert int(a["peer_ok"]) == int(b["peer_ok"])
    if int(a["peer_ok"]):
        for out in (a, b):
            assert int(out["p_err"]) == 0
            assert np.array_equal(out["p_n"], out["n"])
            assert np.array_equal(out["p_l_means"], out["l_means"]) and np.array_equal(out["p_l_vars"], out["l_vars"])
            assert np.array_equal(out["p_cov"], out["cov"]) and np.array_equal(out["p_cov_var"], out["cov_var"])
            # the skewed call: timed out on rank 0, poisoned for rank 1, both redone with NCCL -> same numbers
            assert int(out["p_fallbacks"]) == 0 and int(out["s_fallbacks"]) == 1 and int(out["s2_fallbacks"]) == 1
            assert np.array_equal(out["s_n"], out["n"])
            assert np.array_equal(out["s_l_means"], out["l_means"]) and np.array_equal(out["s_l_vars"], out["l_vars"])
            assert np.array_equal(out["s2_l_means"], out["l_means"])
    else:
        print("peer-memory reduce not available on this box (cudaIpc); NCCL route tested only")
