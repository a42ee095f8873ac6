"""BASELINE.json's full sizes, checked through size-independent properties (the oracle would need minutes to hours):

* cfg2 (3 levels x 1e7 pairs, Legendre 50) through the public API: equals an independent device evaluation (bit-exact
  basis tables from ``mlmcb200_basis_eval`` + torch fp64 sums, chunk by chunk), splits of the levels add up, moment 0 is
  exact, sample counts add up.
* cfg3 (1.25e8 samples = one GPU's share of 1e9, Legendre 100 covariance): exactly symmetric, first row = the level sums
  of the moments kernel (test/test_quantity_concept.py:613), two halves add up.
"""
import numpy as np
import pytest
import torch

from oracle import mlmc_oracle as orc
from test_kernels_gpu import dev, native, rel_close, to_struct

pytestmark = pytest.mark.gpu


def _synth(n, h_f, h_c, seed):
    g = torch.Generator(device=dev()).manual_seed(seed)
    x = torch.randn(n, generator=g, device=dev(), dtype=torch.float64)
    root = torch.sqrt(1e-4 + x.abs())
    fine = x + h_f * root
    coarse = torch.zeros_like(x) if h_c is None else x + h_c * root
    return torch.stack([fine, coarse], dim=1).unsqueeze(2).contiguous()          # [n, 2, 1]


def test_cfg2_full_size_properties():
    from mlmc_b200.moments import Legendre
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.quantity import quantity_estimate as qe
    nat = native()
    n, R = 10_000_000, 50
    steps = orc.level_steps(3, (0.5, 0.005))
    domain = (-3.719016485455709, 3.719016485455709)
    levels = [_synth(n, steps[l], steps[l - 1] if l else None, 1234 + 1000 * l) for l in range(3)]
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    storage = Memory.from_arrays([lv.cpu() for lv in levels], level_parameters=[[h] for h in steps], result_format=spec)
    value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
    fn = Legendre(R, domain)
    qm = qe.estimate_mean(qe.moments(value, fn))
    assert qm.mean[0] == 1.0 and qm.var[0] == 0.0
    assert all(a + b == n for a, b in zip(qm.n_samples, qm.n_rm_samples))
    assert sum(qm.n_rm_samples) > 0                                   # the 1e-4 tails fall outside the domain

    # independent evaluation: table kernel (bit-exact with numpy) + torch fp64 sums
    basis = fn.basis_struct()
    l_means, l_vars = [], []
    for l, rows in enumerate(levels):
        s = torch.zeros(R, dtype=torch.float64, device=dev())
        sq = torch.zeros(R, dtype=torch.float64, device=dev())
        cnt = 0
        for lo in range(0, n, 1_000_000):
            part = rows[lo:lo + 1_000_000]
            pf = nat.basis_eval(basis, part[:, 0, 0].contiguous(), R)
            d = pf if l == 0 else pf - nat.basis_eval(basis, part[:, 1, 0].contiguous(), R)
            ok = ~torch.isnan(d).any(dim=1)
            d = d[ok]
            cnt += int(ok.sum())
            s += d.sum(dim=0)
            sq += (d * d).sum(dim=0)
        assert cnt == qm.n_samples[l]
        l_means.append((s / cnt).cpu().numpy())
        l_vars.append(((sq - s * s / cnt) / (cnt - 1)).cpu().numpy())
    rel_close(qm.l_means, np.array(l_means), rtol=1e-10, atol_scale=1e-13, per_level=True)
    rel_close(qm.l_vars, np.array(l_vars), rtol=1e-10, atol_scale=1e-13, per_level=True)

    # linearity: the level sums of two halves add up to the sums of the whole
    acc_w = nat.LevelAccumulator(1, R, dev())
    acc_h = nat.LevelAccumulator(1, R, dev())
    x = levels[2].permute(2, 0, 1)
    nat.moments_accumulate(basis, x, acc_w.level(0))
    nat.moments_accumulate(basis, x[:, : n // 3], acc_h.level(0))
    nat.moments_accumulate(basis, x[:, n // 3:], acc_h.level(0))
    assert torch.equal(acc_w.acc[:, :2], acc_h.acc[:, :2])
    assert torch.allclose(acc_w.acc, acc_h.acc, rtol=1e-11, atol=1e-9)


def test_cfg3_full_size_properties():
    nat = native()
    n, R = 125_000_000, 100
    domain = (-3.719016485455709, 3.719016485455709)
    rows = _synth(n, 0.05, 0.5, 77)                                    # 2 GB
    x = rows.permute(2, 0, 1)
    basis = to_struct(orc.Basis("legendre", R, domain))
    cov = nat.LevelAccumulator(1, R * R, dev())
    nat.gram_accumulate(basis, x, cov.level(0), mode=0, want_var=False)
    mom = nat.LevelAccumulator(1, R, dev())
    nat.moments_accumulate(basis, x, mom.level(0))
    c = cov.acc[0].cpu().numpy()
    m = mom.acc[0].cpu().numpy()
    assert c[0] == m[0] and c[1] == m[1] and c[0] + c[1] == n          # same samples kept / dropped
    g = c[2:2 + R * R].reshape(R, R)
    assert np.array_equal(g, g.T)
    # phi_0 = 1: the first row of sum(phi_i phi_j [fine] - phi_i phi_j [coarse]) is sum(phi_j(f) - phi_j(c))
    rel_close(g[0], m[2:2 + R], rtol=1e-9, atol_scale=1e-12)
    assert g[0, 0] == 0.0
    half = nat.LevelAccumulator(1, R * R, dev())
    nat.gram_accumulate(basis, x[:, : n // 2], half.level(0), mode=0, want_var=False)
    nat.gram_accumulate(basis, x[:, n // 2:], half.level(0), mode=0, want_var=False)
    assert torch.equal(cov.acc[0, :2], half.acc[0, :2])
    scale = float(cov.acc[0, 2:2 + R * R].abs().max())
    assert float((cov.acc[0, 2:2 + R * R] - half.acc[0, 2:2 + R * R]).abs().max()) < 1e-10 * scale


def test_cfg2_size_weighted_bootstrap_identity_and_linearity():
    """One-pass bootstrap kernel (csrc/bootstrap.cu) at cfg2's level size, size-independent properties: multiplicity 1 for
    every row reproduces the plain level sums of the fused moments kernel (counts exactly, sums to 1e-10); multiplicity 2
    gives exactly twice that, bit for bit (scaling by two is exact in every partial sum); multiplicity 0 gives zeros;
    a replicate that keeps only the first half of the rows equals the plain sums of that half."""
    nat = native()
    n, R = 10_000_000, 50
    g = torch.Generator(device=dev()).manual_seed(7)
    x0 = torch.randn(n, generator=g, device=dev(), dtype=torch.float64)
    rows = torch.stack([x0, x0 + 0.05 * torch.randn(n, generator=g, device=dev(), dtype=torch.float64)], dim=1).contiguous()
    x = rows.reshape(n, 2, 1).permute(2, 0, 1)
    basis = to_struct(orc.Basis("legendre", R, (-3.719016485455709, 3.719016485455709)))
    counts = torch.zeros((4, n), dtype=torch.uint8, device=dev())
    counts[0] = 1
    counts[1] = 2
    counts[3, :n // 2] = 1
    acc = torch.zeros((4, 2 + 2 * R), dtype=torch.float64, device=dev())
    nat.moments_accumulate_weighted(basis, x, counts, acc)
    plain = torch.zeros(2 + 2 * R, dtype=torch.float64, device=dev())
    nat.moments_accumulate(basis, x, plain)
    half = torch.zeros(2 + 2 * R, dtype=torch.float64, device=dev())
    nat.moments_accumulate(basis, x[:, :n // 2], half)
    a, p, h = acc.cpu().numpy(), plain.cpu().numpy(), half.cpu().numpy()
    assert a[0, 0] == p[0] and a[0, 1] == p[1] and p[0] + p[1] == n and p[1] > 0
    rel_close(a[0, 2:], p[2:], rtol=1e-10, atol_scale=1e-12)
    assert np.array_equal(a[1], 2.0 * a[0])
    assert not a[2].any()
    assert a[3, 0] == h[0] and a[3, 1] == h[1]
    rel_close(a[3, 2:], h[2:], rtol=1e-10, atol_scale=1e-12)
