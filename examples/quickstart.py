"""Drop-in use of the B200 estimation path: the same calls a GeoMop/MLMC post-processing script makes
(docs/source/examples_postprocessing.rst:45-82 of the reference), with ``mlmc`` replaced by ``mlmc_b200``.

    python examples/quickstart.py            # needs one CUDA device
"""
import os
import sys

import numpy as np
import scipy.stats as stats

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200.moments import Legendre
from mlmc_b200.sample_storage import Memory
from mlmc_b200.quantity.quantity import make_root_quantity
from mlmc_b200.quantity.quantity_spec import QuantitySpec
from mlmc_b200.estimator import Estimate, estimate_n_samples_for_target_variance, determine_level_parameters

# --- synthetic 3-level samples in the reference's storage row order [N, 2, M] (fine row, coarse row) ---
rng = np.random.default_rng(1234)
steps = [p[0] for p in determine_level_parameters(3, (0.5, 0.005))]
levels = []
for l, n in enumerate([200_000, 50_000, 10_000]):
    x = rng.normal(size=n)
    root = np.sqrt(1e-4 + np.abs(x))
    coarse = x + steps[l - 1] * root if l else np.zeros(n)
    levels.append(np.stack([x + steps[l] * root, coarse], axis=1)[:, :, None])
spec = [QuantitySpec(name="value", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
n_ops = [(1 / h) ** 2 * np.log(max(1 / h, 2.0)) for h in steps]
storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], n_ops=n_ops, result_format=spec)

# --- the reference's post-processing calls ---
root_quantity = make_root_quantity(storage, spec)
value = root_quantity["value"][0.0]["0"][0, 0]
domain = Estimate.estimate_domain(value, storage, quantile=0.001)
moments_fn = Legendre(20, domain)
estimator = Estimate(value, storage, moments_fn)

means, variances = estimator.estimate_moments()
print("moment means     :", np.round(means[:5], 6))
print("estimate variance:", variances[:5])

reg_vars, ops = estimator.estimate_diff_vars_regression(storage.get_n_collected())
print("n_samples for target variance 1e-5:", estimate_n_samples_for_target_variance(1e-5, reg_vars, ops, n_levels=3))

cov_mean, cov_var = estimator.estimate_covariance()
print("covariance matrix:", cov_mean.shape, "symmetric:", np.array_equal(cov_mean, cov_mean.T))

distr, info, result, orth_moments = estimator.construct_density(tol=1e-8, orth_moments_tol=1e-4)
xs = np.linspace(domain[0], domain[1], 7)
print("max-ent pdf :", np.round(distr.density(xs), 4))
print("normal pdf  :", np.round(stats.norm.pdf(xs), 4))
